// msm.cuh — Pippenger multi-scalar multiplication over BN254 G1 for sm_100a (replaces
// halo2_proofs::arithmetic::best_multiexp and ParamsKZG::commit / commit_lagrange; SURVEY.md section 8 rows a3, a8;
// reference behaviour Appendix B.1).  The result is the unique group element sum_i s_i * P_i, so any window layout
// gives the reference's answer; this one is chosen for the GPU:
//
//   1. k_msm_digits     canonical scalar (one Montgomery reduction) -> signed c-bit digits; one (key, value) entry per
//                       non-zero digit: key = bucket id, value = base-table index | sign
//   2. counting sort    histogram (fused into k_msm_digits) -> exclusive scan -> scatter (k_msm_scatter): entries grouped by bucket
//   3. k_msm_accumulate one thread per task (a run of <= CH entries of one bucket): gathers 64-byte affine bases with
//                       16-byte loads and adds them into an XYZZ accumulator (8M + 2S per point); heavy buckets are
//                       split into several tasks whose partial sums are merged warp-cooperatively (k_msm_merge)
//   4. k_msm_reduce_*   sum_b (b+1) * B_b per bucket set by chunked running sums
//   5. k_msm_combine    sum over bucket sets of 2^(c*u) * R_u
//
// With tables 2^(c*nsets*t) * P_i precomputed per ParamsKZG (k_msm_precompute) all windows of a scalar share one
// bucket set (nsets = 1): one reduction, no doublings.
#pragma once
#include "ec_quad.cuh"

namespace de {

#define DE_MSM_INVALID 0xffffffffu

struct MsmShape {
    unsigned int c;         // window bits
    unsigned int W;         // number of windows, c * W >= 255
    unsigned int nsets;     // bucket sets per polynomial; window w uses set w % nsets and table w / nsets
    unsigned int NB;        // buckets per set = 2^(c-1)
    unsigned int count;     // polynomials in the batch
    unsigned long long n;   // scalars per polynomial
    unsigned long long table_stride;  // elements between consecutive base tables
    unsigned long long base_offset;   // first base of this (sharded) range inside each table
    unsigned int alt_first;           // polynomials >= alt_first take their bases alt_delta elements further on (the other
    long long alt_delta;              // basis of the same ParamsKZG): one launch sequence for commitments over both bases
};

// ---- 1. digits -------------------------------------------------------------------------------------------------
// AGG: the histogram's atomics are aggregated per warp (lanes with equal keys elect a leader that adds their count once and
// hands out consecutive ranks).  Neighbouring rows of a column often carry the SAME scalar - the constant tail of a grand product
// behind the used rows, runs of 0 / 1 in a witness column - and then all 32 lanes of a warp hit one counter 16 times over.
template <bool AGG>
__global__ void k_msm_digits(const Fr* scalars, unsigned long long stride, MsmShape sh, unsigned int* keys, unsigned int* vals,
                             unsigned int* ranks, unsigned int* counts) {
    unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = gid < sh.n * sh.count;
    if (!AGG && !live) return;
    if (!live) gid = sh.n * sh.count - 1;  // AGG: the whole warp stays for the match; this lane repeats a scalar and writes nothing
    unsigned int b = (unsigned int)(gid / sh.n);
    unsigned long long i = gid % sh.n;
    Fr s = from_mont(load(&scalars[b * stride + i]));
    const unsigned int c = sh.c;
    const unsigned int half = 1u << (c - 1);
    unsigned int carry = 0;
    for (unsigned int w = 0; w < sh.W; w++) {
        unsigned int bit = w * c;
        unsigned int limb = bit >> 5, off = bit & 31;
        unsigned long long lo = limb < 8 ? s.l[limb] : 0;
        unsigned long long hi = (limb + 1) < 8 ? s.l[limb + 1] : 0;
        unsigned int raw = (unsigned int)(((lo | (hi << 32)) >> off) & ((1ull << c) - 1));
        unsigned int v = raw + carry;
        unsigned int neg = 0, mag = v;
        carry = 0;
        if (v > half) {
            mag = (1u << c) - v;
            neg = 1;
            carry = 1;
        }
        unsigned long long e = ((unsigned long long)b * sh.W + w) * sh.n + i;
        unsigned int set = w % sh.nsets, table = w / sh.nsets;
        if (AGG) {
            const unsigned int key = (live && mag) ? (b * sh.nsets + set) * sh.NB + (mag - 1) : DE_MSM_INVALID;
            const unsigned int lane = threadIdx.x & 31u;
            // the match is worth its latency only where keys repeat, and repeated keys come in RUNS (equal scalars in neighbouring
            // rows): one shuffle and a vote decide per warp and window; uniform columns take the plain atomics
            const unsigned int next_key = __shfl_down_sync(0xffffffffu, key, 1);
            const bool runs = __any_sync(0xffffffffu, lane < 31 && key == next_key && key != DE_MSM_INVALID);
            unsigned int rank = 0;
            if (runs) {
                const unsigned int peers = __match_any_sync(0xffffffffu, key);
                const unsigned int leader = __ffs(peers) - 1;
                unsigned int first = 0;
                if (lane == leader && key != DE_MSM_INVALID) first = atomicAdd(&counts[key], __popc(peers));
                first = __shfl_sync(0xffffffffu, first, leader);
                rank = first + __popc(peers & ((1u << lane) - 1u));
            } else if (key != DE_MSM_INVALID) {
                rank = atomicAdd(&counts[key], 1u);
            }
            if (live) {
                keys[e] = key;
                if (key != DE_MSM_INVALID) {
                    unsigned long long tb = table * sh.table_stride + sh.base_offset + i;
                    if (b >= sh.alt_first) tb = (unsigned long long)((long long)tb + sh.alt_delta);
                    vals[e] = (unsigned int)tb | (neg << 31);
                    if (ranks) ranks[e] = rank;
                }
            }
        } else if (mag == 0) {
            keys[e] = DE_MSM_INVALID;
        } else {
            const unsigned int key = (b * sh.nsets + set) * sh.NB + (mag - 1);
            keys[e] = key;
            unsigned long long tb = table * sh.table_stride + sh.base_offset + i;
            if (b >= sh.alt_first) tb = (unsigned long long)((long long)tb + sh.alt_delta);
            vals[e] = (unsigned int)tb | (neg << 31);
            // bucket histogram of the counting sort, fused here; the value the atomic returns is this entry's rank inside its
            // bucket, which spares the scatter an atomic of its own (ranks == nullptr: the scatter takes its own, below)
            if (ranks) ranks[e] = atomicAdd(&counts[key], 1u);
            else atomicAdd(&counts[key], 1u);
        }
    }
}

// ---- 2. counting sort ------------------------------------------------------------------------------------------
// Two forms.  With ranks: position = bucket offset + the rank k_msm_digits captured, no atomics here (proof-sized batches, 8.4 M
// entries over 262 144 buckets: 105 -> ~60 us per round, 149 -> 155 proofs/s; mod_pow k = 17: 69.5 -> 79 proofs/s; witness-like
// scalars gain at every size).  With a cursor: the histogram's atomics stay fire-and-forget and the scatter takes its position
// from an atomic on a cursor array.  Measured on ONE commitment of uniform scalars (DE_SCATTER_RANKS=0/1): equal up to 2^20
// (3.38 vs 3.40 ms), the cursor form ahead from 2^21 (5.87 vs 6.15 ms, 2^22: 11.0 vs 12.0, 2^24: 42.4 vs 43.7) - msm.cu switches
// at n = 2^20 scalars per polynomial.
__global__ void k_msm_scatter(const unsigned int* keys, const unsigned int* vals, const unsigned int* ranks, unsigned long long E,
                              const unsigned int* offsets, unsigned int* sorted) {
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    unsigned int k = keys[e];
    if (k == DE_MSM_INVALID) return;
    sorted[offsets[k] + ranks[e]] = vals[e];
}
__global__ void k_msm_scatter_cursor(const unsigned int* keys, const unsigned int* vals, unsigned long long E, unsigned int* cursor,
                                     unsigned int* sorted) {
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    unsigned int k = keys[e];
    if (k == DE_MSM_INVALID) return;
    unsigned int pos = atomicAdd(&cursor[k], 1u);
    sorted[pos] = vals[e];
}

// exclusive scan of n u32 values in three kernels (4096 items per block; n <= 4096 * 4096).  out has n + 1 entries.
#define DE_SCAN_ITEMS 4
#define DE_SCAN_THREADS 1024
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total, unsigned int* sm) {
    // sm: 32 words
    const unsigned int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) sm[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned int s = sm[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        sm[lane] = s;
    }
    __syncthreads();
    unsigned int base = wid ? sm[wid - 1] : 0;
    *total = sm[31];
    __syncthreads();
    return base + x - v;
}
__global__ void __launch_bounds__(DE_SCAN_THREADS) k_scan_blocks(const unsigned int* in, unsigned long long n, unsigned int* out,
                                                                   unsigned int* block_sums) {
    __shared__ unsigned int sm[32];
    unsigned long long base = ((unsigned long long)blockIdx.x * DE_SCAN_THREADS + threadIdx.x) * DE_SCAN_ITEMS;
    unsigned int v[DE_SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < DE_SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    unsigned int total;
    unsigned int ex = block_exclusive_scan(sum, &total, sm);
#pragma unroll
    for (int k = 0; k < DE_SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(DE_SCAN_THREADS) k_scan_tops(unsigned int* block_sums, unsigned int nblocks, unsigned int* grand_total) {
    __shared__ unsigned int sm[32];
    unsigned int base = threadIdx.x * DE_SCAN_ITEMS;
    unsigned int v[DE_SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < DE_SCAN_ITEMS; k++) {
        v[k] = (base + k < nblocks) ? block_sums[base + k] : 0;
        sum += v[k];
    }
    unsigned int total;
    unsigned int ex = block_exclusive_scan(sum, &total, sm);
#pragma unroll
    for (int k = 0; k < DE_SCAN_ITEMS; k++) {
        if (base + k < nblocks) block_sums[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) *grand_total = total;
}
__global__ void k_scan_add(unsigned int* out, unsigned long long n, const unsigned int* block_sums) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] += block_sums[i / (DE_SCAN_THREADS * DE_SCAN_ITEMS)];
}

// ---- 3. bucket accumulation ------------------------------------------------------------------------------------
// A task is a run of <= CH entries of one bucket.  Tasks are binned by exact length and laid out longest-first, so that
// the 32 tasks of a warp have (almost always) the same trip count and the longest tasks start first (LPT order).
#define DE_MSM_MAX_CH 128

// pass 1: tasks per bucket, multi-task bucket lists (few partials: one thread merges; many: one warp), length histogram
// task length for this launch, from the ACTUAL number of entries (scal[0], known after the scan): enough tasks to fill the chip
// several times over, each no longer than CH_max entries.  Witness-like columns produce far fewer entries than the worst case
// the host can bound, and a task length sized for the worst case would leave most SMs idle.  One thread; writes scal[4].
__global__ void k_msm_choose_ch(unsigned int* scal, unsigned int nbuckets, unsigned int ch_max, unsigned int target_tasks) {
    if (blockIdx.x || threadIdx.x) return;
    const unsigned int entries = scal[0];
    unsigned int ch = 8;
    while (ch < ch_max && entries / (ch * 2) >= target_tasks) ch *= 2;
    const unsigned int avg = (entries + nbuckets - 1) / nbuckets;
    if (nbuckets >= target_tasks / 2)
        while (ch < ch_max && ch < 2 * avg) ch *= 2;
    scal[4] = ch;
}

__global__ void __launch_bounds__(256) k_msm_task_counts(const unsigned int* counts, unsigned int nbuckets, unsigned int* ntasks,
                                                         unsigned int* multi_small, unsigned int* multi_large, unsigned int* scal,
                                                         unsigned int* len_bins) {
    __shared__ unsigned int sbins[DE_MSM_MAX_CH + 1];
    const unsigned int CH = scal[4];
    for (unsigned int i = threadIdx.x; i <= CH; i += blockDim.x) sbins[i] = 0;
    __syncthreads();
    unsigned int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbuckets) {
        unsigned int cnt = counts[b];
        unsigned int full = cnt / CH, rem = cnt % CH;
        unsigned int t = full + (rem ? 1 : 0);
        ntasks[b] = t;
        if (t > 1) {
            if (t <= 8) multi_small[atomicAdd(&scal[2], 1u)] = b;
            else multi_large[atomicAdd(&scal[3], 1u)] = b;
        }
        if (full) atomicAdd(&sbins[CH], full);
        if (rem) atomicAdd(&sbins[rem], 1u);
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i <= CH; i += blockDim.x)
        if (sbins[i]) atomicAdd(&len_bins[i], sbins[i]);
}
// bin start offsets, longest first: start[len] = sum of bins of greater length.  One block.
__global__ void k_msm_bin_starts(const unsigned int* len_bins, const unsigned int* scal, unsigned int* bin_cursor) {
    if (threadIdx.x == 0) {
        const unsigned int CH = scal[4];
        unsigned int run = 0;
        for (int l = (int)CH; l >= 1; l--) {
            bin_cursor[l] = run;
            run += len_bins[l];
        }
        bin_cursor[0] = run;
    }
}
// pass 2: every bucket writes its tasks (bucket, local index) into the slot range of their length bin
__global__ void __launch_bounds__(256) k_msm_task_fill(const unsigned int* counts, unsigned int nbuckets, const unsigned int* scal,
                                                       unsigned int* bin_cursor, uint2* task_list) {
    __shared__ unsigned int scount[DE_MSM_MAX_CH + 1];
    __shared__ unsigned int sbase[DE_MSM_MAX_CH + 1];
    const unsigned int CH = scal[4];
    for (unsigned int i = threadIdx.x; i <= CH; i += blockDim.x) scount[i] = 0;
    __syncthreads();
    unsigned int b = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int cnt = 0, full = 0, rem = 0, off_full = 0, off_rem = 0;
    if (b < nbuckets) {
        cnt = counts[b];
        full = cnt / CH;
        rem = cnt % CH;
        if (full) off_full = atomicAdd(&scount[CH], full);
        if (rem) off_rem = atomicAdd(&scount[rem], 1u);
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i <= CH; i += blockDim.x)
        if (scount[i]) sbase[i] = atomicAdd(&bin_cursor[i], scount[i]);
    __syncthreads();
    if (b < nbuckets) {
        for (unsigned int k = 0; k < full; k++) task_list[sbase[CH] + off_full + k] = make_uint2(b, k);
        if (rem) task_list[sbase[rem] + off_rem] = make_uint2(b, full);
    }
}

__device__ __forceinline__ Affine msm_fetch(const Affine* bases, unsigned int v) {
    Affine p = load_affine(&bases[v & 0x7fffffffu]);
    if (v >> 31) p.y = neg(p.y);
    return p;
}

// 4 CTAs per SM (128 registers, 8 bytes of stack): with the dedicated squaring the loop fits, and 16 resident warps instead of 12
// (130 registers) measure 0.556 against 0.572 ms per launch, 161.3 against 159.3 proofs/s; before the squaring the same bound
// spilled more and measured slower (0.614 against 0.600 ms)
__global__ void __launch_bounds__(128, 4) k_msm_accumulate(const unsigned int* sorted, const unsigned int* offsets, const unsigned int* counts,
                                                        const unsigned int* task_off, const uint2* task_list, const unsigned int* scal,
                                                        const Affine* bases, XYZZ* buckets, XYZZ* partials) {
    unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= scal[1]) return;
    const unsigned int CH = scal[4];
    const uint2 task = task_list[t];
    const unsigned int b = task.x, local = task.y;
    const unsigned int cnt = counts[b];
    const unsigned int start = offsets[b] + local * CH;
    unsigned int len = cnt - local * CH;
    if (len > CH) len = CH;
    XYZZ acc = xyzz_identity();
    Affine next = msm_fetch(bases, sorted[start]);
    for (unsigned int k = 0; k < len; k++) {
        Affine cur = next;
        if (k + 1 < len) next = msm_fetch(bases, sorted[start + k + 1]);
        xyzz_madd(acc, cur);
    }
    if (cnt <= CH) store_xyzz(&buckets[b], acc);
    else store_xyzz(&partials[task_off[b] + local], acc);
}

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
// buckets split into 2..8 tasks: a quad adds the partial sums (<= 7 dependent additions of 6 multiplication latencies each,
// ec_quad.cuh; one thread per bucket took 14 per addition)
__global__ void __launch_bounds__(128) k_msm_merge_small(const unsigned int* multi_small, const unsigned int* scal, const unsigned int* task_off,
                                                         const XYZZ* partials, XYZZ* buckets) {
    const unsigned int total = scal[2];
    const unsigned int role = threadIdx.x & 3u, quad = (threadIdx.x & 31u) >> 2;
    const unsigned int quads = gridDim.x * blockDim.x / 4;
    // every warp walks blocks of 8 buckets; the loop bound is warp-uniform (the shuffles inside need all 32 lanes)
    for (unsigned int base = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) / 4; base < total; base += quads) {
        const unsigned int m = base + quad;
        const bool active = m < total;
        unsigned int b = 0, first = 0, n = 0;
        if (active) {
            b = multi_small[m];
            first = task_off[b];
            n = task_off[b + 1] - first;
        }
        Fq acc = active ? quad_load(&partials[first], role) : Fq::zero();
        const unsigned int longest = __reduce_max_sync(0xffffffffu, n);
        for (unsigned int i = 1; i < longest; i++) {
            const Fq v = i < n ? quad_load(&partials[first + i], role) : Fq::zero();  // identity: the sum is unchanged
            acc = quad_add(acc, v, role);
        }
        if (active) quad_store(&buckets[b], role, acc);
    }
}
// heavy buckets (> 8 tasks).  More than 32 partial sums: one 128-thread CTA per bucket - the threads stride over the partial sums
// (a witness column's 0 / 1 digits put thousands of partials into one bucket: 128 lanes keep that chain at count / 128
// additions), then a tree in shared memory whose levels of <= 32 additions run four lanes per addition (smem_tree_sum,
// ec_quad.cuh).  9 .. 32 partial sums - the hundreds of small-value buckets of a witness column at task length 8 - take ONE warp
// each, four buckets per CTA at a time (WARP_MEDIUM; with a whole CTA per such bucket three of its four warps only held registers).
template <bool WARP_MEDIUM>
__global__ void __launch_bounds__(128) k_msm_merge_large(const unsigned int* multi_large, const unsigned int* scal, const unsigned int* task_off,
                                                         const XYZZ* partials, XYZZ* buckets) {
    __shared__ SmemPoints<128> sm;
    const unsigned int tid = threadIdx.x;
    const unsigned int total = scal[3];
    for (unsigned int m = blockIdx.x; m < total; m += gridDim.x) {
        const unsigned int b = multi_large[m];
        const unsigned int first = task_off[b], last = task_off[b + 1];
        const unsigned int cnt = last - first;
        if (WARP_MEDIUM && cnt <= 32) continue;             // CTA-uniform
        const unsigned int stride = cnt > 32 ? 128u : 32u;  // CTA-uniform
        XYZZ acc = xyzz_identity();
        if (tid < stride)
            for (unsigned int p = first + tid; p < last; p += stride) {
                XYZZ v = load_xyzz(&partials[p]);
                xyzz_add(acc, v);
            }
        if (tid < stride) sm.put(tid, acc);
        smem_tree_sum(sm, tid, stride);
        if (tid == 0) store_xyzz(&buckets[b], sm.get(0));
        __syncthreads();  // sm is reused by the next bucket of this CTA
    }
    if (!WARP_MEDIUM) return;
    const unsigned int lane = tid & 31u, wid = tid >> 5;
    for (unsigned int m = blockIdx.x * 4 + wid; m < total; m += gridDim.x * 4) {  // warp-uniform
        const unsigned int b = multi_large[m];
        const unsigned int first = task_off[b];
        const unsigned int cnt = task_off[b + 1] - first;
        if (cnt > 32) continue;
        XYZZ acc = xyzz_identity();
        if (lane < cnt) acc = load_xyzz(&partials[first + lane]);
        sm.put(32 * wid + lane, acc);
        smem_warp_tree_sum(sm, 32 * wid, lane);
        if (lane == 0) store_xyzz(&buckets[b], sm.get(32 * wid));
        __syncwarp();  // this warp's slots are reused by its next bucket
    }
}

// ---- 4. bucket reduction ---------------------------------------------------------------------------------------
// sum_b (b + 1) * B_b without long serial chains: write b in radix-32 digits d_j (bit offset 5j, the top digit narrower).
//   sum_b (b+1) B_b = sum_b B_b + sum_j 2^(5j) * sum_v v * D[j][v],   D[j][v] = sum of the buckets whose j-th digit is v.
// Every D[j][v] is a PLAIN sum (a CTA: 8 serial adds per thread, then a tree), so the only dependent chain left is the
// 32-element weighted sum per digit, done by one warp with two shuffle scans.
__global__ void __launch_bounds__(128, 4) k_msm_digit_sums(const XYZZ* buckets, unsigned int NB, unsigned int cm1 /* c - 1 */, XYZZ* dsums) {
    // grid: x = digit slot (j * 32 + v), y = bucket set
    __shared__ XYZZ sm[4];
    const unsigned int j = blockIdx.x >> 5, v = blockIdx.x & 31;
    const unsigned int off = 5 * j;
    const unsigned int width = (cm1 - off) < 5 ? (cm1 - off) : 5;
    const unsigned int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    XYZZ acc = xyzz_identity();
    if (v < (1u << width)) {
        const XYZZ* B = buckets + (unsigned long long)blockIdx.y * NB;
        const unsigned int m_count = NB >> width;
        for (unsigned int m = tid; m < m_count; m += blockDim.x) {
            unsigned int lo = m & ((1u << off) - 1), hi = m >> off;
            unsigned int b = (hi << (off + width)) | (v << off) | lo;
            XYZZ x = load_xyzz(&B[b]);
            xyzz_add(acc, x);
        }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ o = shfl_down_xyzz(acc, d);
        if (lane + d < 32) xyzz_add(acc, o);
    }
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (tid == 0) {
        for (unsigned int w = 1; w < (blockDim.x >> 5); w++) {
            XYZZ o = sm[w];
            xyzz_add(acc, o);
        }
        store_xyzz(&dsums[(unsigned long long)blockIdx.y * gridDim.x + blockIdx.x], acc);
    }
}
// one CTA per bucket set, one warp per digit: W_j = sum_v v * D[j][v] by two shuffle scans (all digits in parallel), then
// thread 0 folds the digits: result = (..(W_top * 32 + W_{top-1}) * 32 ..) + W_0 + total
__global__ void __launch_bounds__(128) k_msm_digit_final(const XYZZ* dsums, unsigned int ndigits, XYZZ* set_out) {
    __shared__ XYZZ sw[4];
    __shared__ XYZZ stotal;
    const unsigned int set = blockIdx.x, lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const XYZZ* D = dsums + (unsigned long long)set * ndigits * 32;
    if (j < ndigits) {
        XYZZ r = load_xyzz(&D[j * 32 + lane]);
        // suffix sums R_v = sum_{u >= v} X_u
        for (int d = 1; d < 32; d <<= 1) {
            XYZZ o = shfl_down_xyzz(r, d);
            if (lane + d < 32) xyzz_add(r, o);
        }
        if (lane == 0) {
            if (j == 0) stotal = r;  // R_0 of digit 0 = sum of all buckets
            r = xyzz_identity();
        }
        // sum_{v >= 1} R_v = sum_v v * X_v
        for (int d = 16; d >= 1; d >>= 1) {
            XYZZ o = shfl_down_xyzz(r, d);
            if (lane + d < 32) xyzz_add(r, o);
        }
        if (lane == 0) sw[j] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        XYZZ result = sw[ndigits - 1];
        for (int k = (int)ndigits - 2; k >= 0; k--) {
            for (int t = 0; t < 5; t++) result = xyzz_dbl(result);
            XYZZ w = sw[k];
            xyzz_add(result, w);
        }
        XYZZ tot = stotal;
        xyzz_add(result, tot);
        store_xyzz(&set_out[set], result);
    }
}

// ---- 4b. two-digit reduction (c >= 11) ---------------------------------------------------------------------------
// Split the bucket index b = u * V0 + v (v: low w0 bits, u: high w1 bits).  Then
//   sum_b (b+1) B_b = T + sum_v v * D0[v] + 2^w0 * sum_u u * D1[u],   D0[v] = sum_u B[u][v],  D1[u] = sum_v B[u][v],  T = sum D0.
// D0 / D1 are plain sums (2 additions per bucket in total instead of 3), formed by lane-efficient segmented sums: every
// thread adds <= 8 terms serially, level after level; the two short weighted sums reuse the radix-32 digit kernels.
// out[i] = sum_{k < seg} in[i * seg + k]
__global__ void __launch_bounds__(128, 4) k_xyzz_segsum(const XYZZ* in, XYZZ* out, unsigned long long n_out, unsigned int seg) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const XYZZ* p = in + i * seg;
    XYZZ acc = load_xyzz(&p[0]);
    for (unsigned int k = 1; k < seg; k++) {
        XYZZ v = load_xyzz(&p[k]);
        xyzz_add(acc, v);
    }
    store_xyzz(&out[i], acc);
}
// column partial sums: out[(set * V0 + v) * Q + q] = sum_{k < seg} B[set][(q * seg + k) * V0 + v]
__global__ void __launch_bounds__(128, 4) k_xyzz_colsum(const XYZZ* buckets, XYZZ* out, unsigned int NB, unsigned int w0, unsigned int Q,
                                                         unsigned int seg, unsigned long long n_out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    // consecutive threads take consecutive v (coalesced 128-byte bucket reads): decode i as (set, q, v) for the loads
    const unsigned int V0 = 1u << w0;
    const unsigned int v = (unsigned int)(i & (V0 - 1));
    const unsigned int q = (unsigned int)((i >> w0) % Q);
    const unsigned long long set = (i >> w0) / Q;
    const XYZZ* B = buckets + set * NB + (unsigned long long)q * seg * V0 + v;
    XYZZ acc = load_xyzz(&B[0]);
    for (unsigned int k = 1; k < seg; k++) {
        XYZZ x = load_xyzz(&B[(unsigned long long)k * V0]);
        xyzz_add(acc, x);
    }
    store_xyzz(&out[(set * V0 + v) * Q + q], acc);
}
// one CTA per bucket set, warp (a, j) = weighted sum of digit j of array a; thread 0 folds:
//   result = T + W0 + 2^w0 * W1,  W_a = W_{a,1} * 32 + W_{a,0}
__global__ void __launch_bounds__(128) k_msm_digit_final2(const XYZZ* dsums0, unsigned int nd0, const XYZZ* dsums1, unsigned int nd1,
                                                          unsigned int w0, XYZZ* set_out) {
    __shared__ XYZZ sw[4];
    __shared__ XYZZ stotal;
    const unsigned int set = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int a = wid >> 1, j = wid & 1;
    const unsigned int nd = a ? nd1 : nd0;
    const XYZZ* D = (a ? dsums1 : dsums0) + (unsigned long long)set * nd * 32;
    XYZZ r = xyzz_identity();
    if (j < nd) r = load_xyzz(&D[j * 32 + lane]);
    for (int d = 1; d < 32; d <<= 1) {
        XYZZ o = shfl_down_xyzz(r, d);
        if (lane + d < 32) xyzz_add(r, o);
    }
    if (lane == 0) {
        if (wid == 0) stotal = r;
        r = xyzz_identity();
    }
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ o = shfl_down_xyzz(r, d);
        if (lane + d < 32) xyzz_add(r, o);
    }
    if (lane == 0) sw[wid] = r;
    __syncthreads();
    if (threadIdx.x < 64 && lane == 0) {
        // warp 0 lane 0 folds array 0, warp 1 lane 0 folds array 1 (in parallel), results back through shared memory
        const unsigned int arr = threadIdx.x >> 5;
        XYZZ w = sw[arr * 2 + 1];
        for (int t = 0; t < 5; t++) w = xyzz_dbl(w);
        XYZZ lo = sw[arr * 2];
        xyzz_add(w, lo);
        if (arr == 1)
            for (unsigned int t = 0; t < w0; t++) w = xyzz_dbl(w);
        sw[arr * 2] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        XYZZ result = sw[0];
        XYZZ w1 = sw[2];
        xyzz_add(result, w1);
        XYZZ tot = stotal;
        xyzz_add(result, tot);
        store_xyzz(&set_out[set], result);
    }
}

// ---- 4c. latency-oriented two-digit reduction for c <= 16 (the proof-sized MSMs) ---------------------------------------
// The same decomposition as 4b (D0[v] = column sums, D1[u] = row sums of the 2^w1 x 2^w0 bucket array), arranged so that the
// chain of DEPENDENT point additions is short and no CTA occupies an SM for long (a dependent XYZZ addition costs ~6 us of
// latency on one warp, whatever the occupancy):
//   k_bucket_rowcol   one launch, 64-thread CTAs: a CTA per bucket row (<= 4 serial additions per thread + a 6-level tree in
//                     shared memory) and a CTA per 4 columns (16 row groups, <= 8 serial additions + a 4-level tree);
//                     tree levels of <= 16 additions run four lanes per addition (ec_quad.cuh)
//   k_bucket_bitsums  sum_v v * D[v] = sum_j 2^j * S_j with S_j = sum of the D[v] whose index has bit j set: one 64-thread
//                     CTA per (array, bit) forms S_j by a tree; one more forms T = sum of all buckets
//   k_bucket_bits_final  one warp per bucket set: lane s doubles its term s times (bit j of D1 weighs 2^(w0 + j) and sits in
//                     slot w0 + j), then a 4-level tree adds the <= 16 terms.  result = T + sum_s 2^s * S_s
#define DE_RC_THREADS 64
#define DE_RC_COLS 4  // 2 columns x 32 row groups (a shorter serial phase, twice the CTAs) measured 1 % slower per proof
__global__ void __launch_bounds__(DE_RC_THREADS, 8) k_bucket_rowcol(const XYZZ* buckets, unsigned int NB, unsigned int w0, unsigned int w1, XYZZ* D0,
                                                                    XYZZ* D1) {
    __shared__ SmemPoints<DE_RC_THREADS> sm;
    const unsigned int V0 = 1u << w0, V1 = 1u << w1;
    const unsigned int tid = threadIdx.x;
    const unsigned long long set = blockIdx.y;
    const XYZZ* B = buckets + set * NB;
    if (blockIdx.x < V1) {
        // row sum: D1[u] = sum_v B[u][v] over V0 contiguous buckets; thread t takes v = t, t + 64, ...
        const unsigned int u = blockIdx.x;
        XYZZ acc = xyzz_identity();
        for (unsigned int v = tid; v < V0; v += DE_RC_THREADS) {
            XYZZ x = load_xyzz(&B[(unsigned long long)u * V0 + v]);
            xyzz_add(acc, x);
        }
        sm.put(tid, acc);
        smem_tree_sum(sm, tid, DE_RC_THREADS);
        if (tid == 0) store_xyzz(&D1[set * V1 + u], sm.get(0));
    } else {
        // column sums of DE_RC_COLS adjacent columns: thread (g, cv) adds rows g, g + G, ... of column v0 + cv (G = 64 /
        // DE_RC_COLS row groups), then a tree over g; its levels of at most 16 additions run four lanes per addition
        constexpr unsigned int G = DE_RC_THREADS / DE_RC_COLS;
        const unsigned int v0 = (blockIdx.x - V1) * DE_RC_COLS;
        const unsigned int g = tid / DE_RC_COLS, cv = tid % DE_RC_COLS;
        XYZZ acc = xyzz_identity();
        for (unsigned int u = g; u < V1; u += G) {
            XYZZ x = load_xyzz(&B[(unsigned long long)u * V0 + v0 + cv]);
            xyzz_add(acc, x);
        }
        // sm[cv * G + g]: each column's G partials are contiguous; tree over g inside every group
        sm.put(cv * G + g, acc);
        for (unsigned int d = G / 2; d >= 1; d >>= 1) {
            __syncthreads();
            if (4 * DE_RC_COLS * d <= DE_RC_THREADS) {
                smem_quad_level(sm, tid, DE_RC_COLS * d, [d](unsigned int j) { return (j / d) * G + j % d; },
                                [d](unsigned int j) { return (j / d) * G + j % d + d; });
                continue;
            }
            const unsigned int c = tid / G, gg = tid % G;
            if (gg < d) {
                XYZZ a = sm.get(c * G + gg);
                XYZZ b = sm.get(c * G + gg + d);
                xyzz_add(a, b);
                sm.put(c * G + gg, a);
            }
        }
        __syncthreads();
        if (tid < DE_RC_COLS) store_xyzz(&D0[set * V0 + v0 + tid], sm.get(tid * G));
    }
}
// grid.x = slot: [0, w0) bit j of D0, [w0, w0 + w1) bit (slot - w0) of D1, w0 + w1: the plain total of D0.  grid.y = set.
__global__ void __launch_bounds__(DE_RC_THREADS, 8) k_bucket_bitsums(const XYZZ* D0, const XYZZ* D1, unsigned int w0, unsigned int w1, XYZZ* S) {
    __shared__ SmemPoints<DE_RC_THREADS> sm;
    const unsigned int tid = threadIdx.x, slot = blockIdx.x;
    const unsigned long long set = blockIdx.y;
    const unsigned int nslots = w0 + w1 + 1;
    XYZZ acc = xyzz_identity();
    if (slot == w0 + w1) {
        const XYZZ* D = D0 + set * (1ull << w0);
        for (unsigned int v = tid; v < (1u << w0); v += DE_RC_THREADS) {
            XYZZ x = load_xyzz(&D[v]);
            xyzz_add(acc, x);
        }
    } else {
        const bool second = slot >= w0;
        const unsigned int j = second ? slot - w0 : slot;
        const unsigned int w = second ? w1 : w0;
        const XYZZ* D = second ? D1 + set * (1ull << w1) : D0 + set * (1ull << w0);
        for (unsigned int m = tid; m < (1u << (w - 1)); m += DE_RC_THREADS) {
            const unsigned int v = ((m >> j) << (j + 1)) | (1u << j) | (m & ((1u << j) - 1));  // m with a 1 inserted at bit j
            XYZZ x = load_xyzz(&D[v]);
            xyzz_add(acc, x);
        }
    }
    sm.put(tid, acc);
    smem_tree_sum(sm, tid, DE_RC_THREADS);
    if (tid == 0) store_xyzz(&S[set * nslots + slot], sm.get(0));
}
__global__ void __launch_bounds__(64) k_bucket_bits_final(const XYZZ* S, unsigned int w0, unsigned int w1, XYZZ* set_out) {
    // 16 terms x 4 lanes (ec_quad.cuh): term s is doubled s times - 4 multiplication latencies per doubling - then a 4-level tree
    __shared__ SmemPoints<16> sm;
    const unsigned int tid = threadIdx.x, role = tid & 3, term = tid >> 2;
    const unsigned long long set = blockIdx.x;
    const unsigned int nslots = w0 + w1 + 1;  // <= 16 for c <= 16; the last slot is the plain total (no doublings)
    Fq c = Fq::zero();
    unsigned int doublings = 0;
    if (term < nslots) {
        c = quad_load(&S[set * nslots + term], role);
        doublings = term == nslots - 1 ? 0 : term;
    }
    const unsigned int warp_doublings = __reduce_max_sync(0xffffffffu, doublings);
    for (unsigned int d = 0; d < warp_doublings; d++) {
        const Fq twice = quad_dbl(c, role);
        if (d < doublings) c = twice;
    }
    sm.put_coord(term, role, c);
    smem_tree_sum(sm, tid, 16);
    if (term == 0) quad_store(&set_out[set], role, sm.coord(0, role));
}

// ---- 5. combine bucket sets: out[b] = sum_u 2^(c*u) * R[b][u], written as Jacobian -------------------------------
__global__ void __launch_bounds__(32) k_msm_combine(const XYZZ* set_in, unsigned int nsets, unsigned int c, Jac* out) {
    const unsigned int b = blockIdx.x, lane = threadIdx.x;
    // nsets <= 32 (c >= 8): lane u scales set u by 2^(c*u), then a shuffle tree adds the lanes
    XYZZ acc = xyzz_identity();
    if (lane < nsets) {
        acc = load_xyzz(&set_in[(unsigned long long)b * nsets + lane]);
        for (unsigned int d = 0; d < c * lane; d++) acc = xyzz_dbl(acc);
    }
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ o = shfl_down_xyzz(acc, d);
        if (lane + d < 32) xyzz_add(acc, o);
    }
    if (lane == 0) {
        Jac j = xyzz_to_jac(acc);
        store(&out[b].x, j.x);
        store(&out[b].y, j.y);
        store(&out[b].z, j.z);
    }
}

// ---- one-time per ParamsKZG: tables[t][i] = 2^(shift * t) * base[i], affine --------------------------------------
__global__ void __launch_bounds__(128) k_msm_precompute(const Affine* bases, unsigned long long n, unsigned int shift, unsigned int ntables,
                                                        unsigned long long table_stride, Affine* tables) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine p = load_affine(&bases[i]);
    store_affine(&tables[i], p);
    XYZZ cur = xyzz_from_affine(p);
    for (unsigned int t = 1; t < ntables; t++) {
        for (unsigned int d = 0; d < shift; d++) cur = xyzz_dbl(cur);
        Affine a;
        if (is_identity(cur)) {
            a.x = Fq::zero();
            a.y = Fq::zero();
        } else {
            // 1/ZZZ by Fermat; 1/ZZ = ZZ^2 / ZZZ^2
            Fq iz3 = inv(cur.zzz);
            Fq iz2 = mul(sqr(cur.zz), sqr(iz3));
            a.x = mul(cur.x, iz2);
            a.y = mul(cur.y, iz3);
        }
        store_affine(&tables[t * table_stride + i], a);
    }
}

// sum of `count` Jacobian points (multi-GPU partial combine); single warp
__global__ void __launch_bounds__(32) k_g1_sum(const Jac* pts, unsigned int count, Jac* out) {
    const unsigned int lane = threadIdx.x;
    XYZZ acc = xyzz_identity();
    for (unsigned int i = lane; i < count; i += 32) {
        Jac p;
        p.x = load(&pts[i].x); p.y = load(&pts[i].y); p.z = load(&pts[i].z);
        if (!p.z.is_zero()) {
            // Jacobian (X, Y, Z) -> XYZZ (X, Y, Z^2, Z^3)
            XYZZ v;
            v.x = p.x; v.y = p.y; v.zz = sqr(p.z); v.zzz = mul(v.zz, p.z);
            xyzz_add(acc, v);
        }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ o = shfl_down_xyzz(acc, d);
        if (lane + d < 32) xyzz_add(acc, o);
    }
    if (lane == 0) {
        Jac j = xyzz_to_jac(acc);
        store(&out->x, j.x); store(&out->y, j.y); store(&out->z, j.z);
    }
}

// Jacobian -> affine, one thread per point (Fermat inversion)
__global__ void k_g1_normalize(const Jac* pts, unsigned int count, Affine* out) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Jac p;
    p.x = load(&pts[i].x); p.y = load(&pts[i].y); p.z = load(&pts[i].z);
    Affine a;
    if (p.z.is_zero()) {
        a.x = Fq::zero();
        a.y = Fq::zero();
    } else {
        Fq zi = inv(p.z);
        Fq zi2 = sqr(zi);
        a.x = mul(p.x, zi2);
        a.y = mul(p.y, mul(zi2, zi));
    }
    store_affine(&out[i], a);
}

// out[i] = [scalars[i]] base, affine (fixed-base scalar multiplication by double-and-add; ParamsKZG::setup's
// g[i] = [s^i] G and synthetic SRS bases for the size sweeps).  One thread per scalar.
__global__ void __launch_bounds__(128) k_g1_mul_base(Affine base, const Fr* scalars, unsigned long long n, Affine* out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr s = from_mont(load(&scalars[i]));
    int top = 255;
    while (top >= 0 && !((s.l[top >> 5] >> (top & 31)) & 1)) top--;
    XYZZ acc = xyzz_identity();
    for (int b = top; b >= 0; b--) {
        acc = xyzz_dbl(acc);
        if ((s.l[b >> 5] >> (b & 31)) & 1) xyzz_madd(acc, base);
    }
    Affine a;
    if (is_identity(acc)) {
        a.x = Fq::zero();
        a.y = Fq::zero();
    } else {
        Fq iz3 = inv(acc.zzz);
        Fq iz2 = mul(sqr(acc.zz), sqr(iz3));
        a.x = mul(acc.x, iz2);
        a.y = mul(acc.y, iz3);
    }
    store_affine(&out[i], a);
}

}  // namespace de
