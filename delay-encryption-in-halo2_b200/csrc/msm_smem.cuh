// msm_smem.cuh — XYZZ points in shared memory and the tree sums over them (one lane or four lanes per addition, ec_quad.cuh):
// shared by the bucket merges (msm.cuh, msm.cu) and the bucket reduction (msm_reduce.cuh, msm_reduce.cu).
#pragma once
#include "ec_quad.cuh"

namespace de {

__device__ __forceinline__ XYZZ shfl_down_xyzz(const XYZZ& v, int delta) {
    XYZZ r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, v.x.l[i], delta);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, v.y.l[i], delta);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, v.zz.l[i], delta);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, v.zzz.l[i], delta);
    }
    return r;
}

// XYZZ points in shared memory as 8 planes of 16-byte words (plane p, slot i at planes[p * N + i]): consecutive threads touch
// consecutive 16-byte words, where an array of 128-byte structures would put every thread of a quarter-warp on the same banks
template <int N>
struct SmemPoints {
    uint4 w[8 * N];
    __device__ __forceinline__ void put(unsigned int i, const XYZZ& v) {
        const Fq* f = &v.x;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            w[(2 * q) * N + i] = make_uint4(f[q].l[0], f[q].l[1], f[q].l[2], f[q].l[3]);
            w[(2 * q + 1) * N + i] = make_uint4(f[q].l[4], f[q].l[5], f[q].l[6], f[q].l[7]);
        }
    }
    __device__ __forceinline__ XYZZ get(unsigned int i) const {
        XYZZ v;
        Fq* f = &v.x;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 a = w[(2 * q) * N + i], b = w[(2 * q + 1) * N + i];
            f[q].l[0] = a.x; f[q].l[1] = a.y; f[q].l[2] = a.z; f[q].l[3] = a.w;
            f[q].l[4] = b.x; f[q].l[5] = b.y; f[q].l[6] = b.z; f[q].l[7] = b.w;
        }
        return v;
    }
    // one coordinate (role 0 X, 1 Y, 2 ZZ, 3 ZZZ) of slot i: what a lane of a quad holds (ec_quad.cuh)
    __device__ __forceinline__ Fq coord(unsigned int i, unsigned int role) const {
        const uint4 a = w[(2 * role) * N + i], b = w[(2 * role + 1) * N + i];
        Fq f;
        f.l[0] = a.x; f.l[1] = a.y; f.l[2] = a.z; f.l[3] = a.w;
        f.l[4] = b.x; f.l[5] = b.y; f.l[6] = b.z; f.l[7] = b.w;
        return f;
    }
    __device__ __forceinline__ void put_coord(unsigned int i, unsigned int role, const Fq& f) {
        w[(2 * role) * N + i] = make_uint4(f.l[0], f.l[1], f.l[2], f.l[3]);
        w[(2 * role + 1) * N + i] = make_uint4(f.l[4], f.l[5], f.l[6], f.l[7]);
    }
};
// one level of `adds` independent additions slot[lhs(j)] += slot[rhs(j)], j < adds, by quads (4 * adds <= threads; whole warps
// only: a warp none of whose quads has an addition skips the level)
template <int N, class Lhs, class Rhs>
__device__ __forceinline__ void smem_quad_level(SmemPoints<N>& s, unsigned int tid, unsigned int adds, Lhs lhs, Rhs rhs) {
    if ((tid & ~31u) >= 4 * adds) return;
    const unsigned int j = tid >> 2, role = tid & 3;
    const bool active = j < adds;
    const Fq a = active ? s.coord(lhs(j), role) : Fq::zero();
    const Fq b = active ? s.coord(rhs(j), role) : Fq::zero();
    const Fq r = quad_add(a, b, role);
    if (active) s.put_coord(lhs(j), role, r);
}
template <int N>
__device__ __forceinline__ void smem_tree_sum(SmemPoints<N>& s, unsigned int tid, unsigned int len) {
    // slot 0 <- sum of slots [0, len), len a power of two <= blockDim (a multiple of 32); ends with a barrier.  Levels with at
    // most blockDim / 4 additions run four lanes per addition (6 multiplication latencies instead of 14)
    for (unsigned int d = len >> 1; d >= 1; d >>= 1) {
        __syncthreads();
        if (4 * d <= blockDim.x) {
            smem_quad_level(s, tid, d, [](unsigned int j) { return j; }, [d](unsigned int j) { return j + d; });
        } else if (tid < d) {
            XYZZ a = s.get(tid);
            XYZZ b = s.get(tid + d);
            xyzz_add(a, b);
            s.put(tid, a);
        }
    }
    __syncthreads();
}
// slot base <- sum of the 32 slots [base, base + 32) by ONE warp (no CTA barrier): 16 one-lane additions, then four quad levels
template <int N>
__device__ __forceinline__ void smem_warp_tree_sum(SmemPoints<N>& s, unsigned int base, unsigned int lane) {
    __syncwarp();
    if (lane < 16) {
        XYZZ a = s.get(base + lane);
        XYZZ b = s.get(base + lane + 16);
        xyzz_add(a, b);
        s.put(base + lane, a);
    }
    for (unsigned int d = 8; d >= 1; d >>= 1) {
        __syncwarp();
        smem_quad_level(s, lane, d, [base](unsigned int j) { return base + j; }, [base, d](unsigned int j) { return base + j + d; });
    }
    __syncwarp();
}

}  // namespace de
