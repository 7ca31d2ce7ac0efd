// msm_reduce.cuh — the bucket reduction of the Pippenger MSM (steps 4 and 5 of msm.cuh's pipeline): sum_b (b + 1) * B_b per bucket
// set and the combination of the sets.  Its own translation unit (msm_reduce.cu) so that the two halves of the MSM compile side
// by side; the kernels are launched by msm_reduce_enqueue only.
#pragma once
#include "msm_smem.cuh"

namespace de {

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

}  // namespace de
