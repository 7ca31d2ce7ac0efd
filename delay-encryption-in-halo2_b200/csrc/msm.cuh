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
//   4. bucket reduction sum_b (b+1) * B_b per bucket set            } msm_reduce.cuh / msm_reduce.cu (their own translation unit;
//   5. k_msm_combine    sum over bucket sets of 2^(c*u) * R_u       } shared-memory point trees: msm_smem.cuh)
//
// With tables 2^(c*nsets*t) * P_i precomputed per ParamsKZG (k_msm_precompute) all windows of a scalar share one
// bucket set (nsets = 1): one reduction, no doublings.
#pragma once
#include "msm_smem.cuh"

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
    // the sorted index runs two entries ahead of the addition and the 64-byte point one: the point gather never waits for the
    // index load it depends on (and the loop then fits 128 registers without its 8 bytes of stack)
    Affine next = msm_fetch(bases, sorted[start]);
    unsigned int idx_next = len > 1 ? sorted[start + 1] : 0;
    for (unsigned int k = 0; k < len; k++) {
        Affine cur = next;
        if (k + 1 < len) next = msm_fetch(bases, idx_next);
        if (k + 2 < len) idx_next = sorted[start + k + 2];
        xyzz_madd(acc, cur);
    }
    if (cnt <= CH) store_xyzz(&buckets[b], acc);
    else store_xyzz(&partials[task_off[b] + local], acc);
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
