// ntt.cuh — Fr number-theoretic transform for sm_100a (replaces halo2_proofs::arithmetic::best_fft and the
// transforms inside poly::EvaluationDomain; SURVEY.md section 8 rows a4-a7, behavioural spec Appendix B.2/B.3).
//
// Algorithm: natural order in, natural order out, A[j] = sum_i a[i] w^(ij), computed as a 1-, 2- or 3-pass
// "four-step" decomposition N = L1*L2*L3 (every Lk = 2^Sk <= 1024).  A pass is ONE kernel: a CTA stages a tile of
// T independent Lk-point sub-transforms in shared memory (T chosen so that the T*32 B they share in global memory
// are contiguous), runs radix-8 decimation-in-frequency rounds on registers with a shared-memory exchange between
// rounds, and writes the tile back transposed, multiplying by the inter-pass twiddle w^(c*j) on the way out.
// Passes 1..p-1 work on strided columns in place; the last pass reads contiguous rows and scatters them to natural
// order.  The zeta-coset scaling and zero padding of coeff_to_extended are fused into the first pass' load, the
// 1/N and zeta^-i scalings of lagrange_to_coeff / extended_to_coeff into the last pass' store.
#pragma once
#include "field.cuh"

namespace de {

struct NttPassParams {
    const Fr* in;
    Fr* out;
    unsigned long long in_batch_stride, out_batch_stride;  // elements between consecutive polynomials of a batch
    unsigned int G;                                         // tiles per group
    unsigned long long in_grp, in_blk, in_tt, in_el;        // element strides: group, tile, tile member, sub-transform index
    unsigned long long out_grp, out_blk, out_tt, out_el;
    const uint4* wl_planes;  // powers of the Lk-th root of unity w_L^i, i < L/2, ALREADY in the kernel's shared-memory layout: the
                             // low 16 bytes of every power, then the high 16 bytes - one TMA bulk copy stages the table
    int tw_mode;   // 0: none, 1: full table tw_full[e] = w_N^e (e < N), 2: two-level tw_hi[e >> lo_bits] * tw_lo[e & mask],
                   // 3: tw_pass[o] = the twiddle of output position o of THIS pass (resident per plan and pass; read with the
                   //    same coalesced pattern as the store, one multiplication per element at any N)
    const Fr* tw_pass;
    const Fr* tw_full;
    const Fr* tw_hi;
    const Fr* tw_lo;
    unsigned int tw_lo_bits;
    unsigned long long tw_mul;  // exponent = column * j * tw_mul
    int in_mode;                // 0: plain, 1: element i is a[i] * zeta^(i mod 3) for i < n_in and 0 beyond (coeff_to_extended)
    unsigned long long n_in;
    int out_mode;               // 0: none, 1: multiply output j by oscale[j mod 3]
    Fr zeta[2];                 // zeta, zeta^2
    Fr oscale[3];
};

// Extra arguments of the multi-GPU transform's first stage (DIST = true): the last pass of the local N/W-point transform stores
// column j straight into the exchange buffer of the rank that owns it - a peer-memory store over NVLink, so the all-to-all of the
// four-step transform costs no pass of its own.  (The inter-stage twiddle w_N^(rank * j) is applied by the receiving side in
// k_ntt_cross, whose integer pipe is idle while it waits for NVLink.)
template <bool DIST>
struct NttDistArgs {};
template <>
struct NttDistArgs<true> {
    Fr* peer[8];                   // exchange buffers of the ranks (peer-mapped device pointers), world <= 8
    unsigned int col_bits;         // log2 C, C = N / world^2 columns per destination rank
    unsigned long long row_off;    // rank * C: this rank's row inside every exchange buffer
    unsigned int block_off;        // the pass is launched in chunks of CTAs (pipelined exchange): first CTA of this launch
};

__device__ __forceinline__ void sm_put(uint4* lo, uint4* hi, int slot, const Fr& v) {
    lo[slot] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[slot] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr sm_get(const uint4* lo, const uint4* hi, int slot) {
    uint4 a = lo[slot], b = hi[slot];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// ---- TMA (bulk asynchronous copy) of the twiddle table into shared memory, completion tracked by an mbarrier -------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

template <int LT>
__device__ __forceinline__ int ntt_slot(int pos, int tt) {
    int i = (pos << LT) | tt;
    return i + ((i >> (3 + LT)) << LT);  // one tile-row of padding every 8 rows: keeps the last round conflict-free
}
template <int S, int LT>
struct NttShape {
    static constexpr int L = 1 << S;
    static constexpr int T = 1 << LT;
    static constexpr int M = L * T;
    static constexpr int MPAD = M + (M >> 3);
    static constexpr int NTHREADS = (M / 8) < 32 ? 32 : (M / 8);
    static constexpr int WN = (L / 2) < 1 ? 1 : (L / 2);
    static constexpr size_t SMEM = (size_t)(2 * MPAD + 2 * WN) * sizeof(uint4);
    static constexpr int MINBLOCKS = 512 / NTHREADS;  // two 256-thread CTAs (16 warps) per SM: caps registers at 128
};

// One radix-2^R decimation-in-frequency round starting at stage Q of an L = 2^S point transform.
template <int S, int LT, int Q, int R>
__device__ __forceinline__ void ntt_round(uint4* lo, uint4* hi, const uint4* wlo, const uint4* whi, int tid, int nthreads) {
    constexpr int T = 1 << LT;
    constexpr int LOWBITS = S - Q - R;
    constexpr int NG = (T << S) >> R;
    for (int g = tid; g < NG; g += nthreads) {
        int tt = g & (T - 1);
        int gg = g >> LT;
        int base_low = gg & ((1 << LOWBITS) - 1);
        int base = ((gg >> LOWBITS) << (LOWBITS + R)) | base_low;
        Fr x[1 << R];
#pragma unroll
        for (int e = 0; e < (1 << R); e++) x[e] = sm_get(lo, hi, ntt_slot<LT>(base + (e << LOWBITS), tt));
#pragma unroll
        for (int u = 0; u < R; u++) {
            constexpr int dummy = 0;
            (void)dummy;
            const int half = 1 << (R - 1 - u);
#pragma unroll
            for (int e = 0; e < (1 << R); e++) {
                if (e & half) continue;
                Fr a = x[e], b = x[e + half];
                x[e] = add(a, b);
                Fr d = sub(a, b);
                const int e_low = e & (half - 1);
                if (LOWBITS == 0 && e_low == 0) {
                    x[e + half] = d;  // twiddle is w^0
                } else {
                    int ex = (base_low << (Q + u)) + (e_low << (S - R + u));
                    x[e + half] = mul(d, sm_get(wlo, whi, ex));
                }
            }
        }
#pragma unroll
        for (int e = 0; e < (1 << R); e++) sm_put(lo, hi, ntt_slot<LT>(base + (e << LOWBITS), tt), x[e]);
    }
    __syncthreads();
}

template <int S, int LT, int Q>
__device__ __forceinline__ void ntt_rounds(uint4* lo, uint4* hi, const uint4* wlo, const uint4* whi, int tid, int nthreads) {
    if constexpr (Q < S) {
        constexpr int R = (S - Q) >= 3 ? 3 : (S - Q);
        ntt_round<S, LT, Q, R>(lo, hi, wlo, whi, tid, nthreads);
        ntt_rounds<S, LT, Q + R>(lo, hi, wlo, whi, tid, nthreads);
    }
}

__device__ __forceinline__ Fr ntt_twiddle(const NttPassParams& p, unsigned long long e) {
    if (p.tw_mode == 1) return load(&p.tw_full[e]);
    Fr h = load(&p.tw_hi[e >> p.tw_lo_bits]);
    Fr l = load(&p.tw_lo[e & ((1ull << p.tw_lo_bits) - 1)]);
    return mul(h, l);
}

template <int S, int LT, bool DIST = false>
__global__ void __launch_bounds__(NttShape<S, LT>::NTHREADS, NttShape<S, LT>::MINBLOCKS)
    k_ntt_pass(const __grid_constant__ NttPassParams p, const __grid_constant__ NttDistArgs<DIST> dx) {
    using Sh = NttShape<S, LT>;
    constexpr int T = Sh::T, M = Sh::M, L = Sh::L;
    extern __shared__ uint4 smem[];
    uint4* lo = smem;
    uint4* hi = lo + Sh::MPAD;
    uint4* wlo = hi + Sh::MPAD;
    uint4* whi = wlo + Sh::WN;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    unsigned int bx = blockIdx.x;
    if constexpr (DIST) bx += dx.block_off;
    const unsigned int grp = bx / p.G, tile = bx % p.G;
    const Fr* in = p.in + (unsigned long long)blockIdx.y * p.in_batch_stride;
    Fr* out = p.out + (unsigned long long)blockIdx.y * p.out_batch_stride;
    const unsigned long long in_base = grp * p.in_grp + tile * p.in_blk;
    const unsigned long long out_base = grp * p.out_grp + tile * p.out_blk;

    // the sub-transform's twiddle table arrives by TMA while the threads gather the tile
    __shared__ __align__(8) unsigned long long tw_bar;
    if constexpr (L >= 2) {
        if (tid == 0) mbar_init(&tw_bar, 1);
        __syncthreads();
        if (tid == 0) {
            constexpr unsigned int bytes = 2u * Sh::WN * sizeof(uint4);
            mbar_expect_tx(&tw_bar, bytes);
            tma_bulk_g2s(wlo, p.wl_planes, bytes, &tw_bar);
        }
    }
    for (int idx = tid; idx < M; idx += nthreads) {
        int tt = idx & (T - 1), t = idx >> LT;
        unsigned long long gi = in_base + tt * p.in_tt + t * p.in_el;
        Fr v;
        if (p.in_mode == 1) {
            if (gi < p.n_in) {
                v = load(&in[gi]);
                unsigned int r3 = (unsigned int)(gi % 3ull);
                if (r3 != 0) v = mul(v, p.zeta[r3 - 1]);
            } else {
                v = Fr::zero();
            }
        } else {
            v = load(&in[gi]);
        }
        sm_put(lo, hi, ntt_slot<LT>(t, tt), v);
    }
    __syncthreads();
    if constexpr (L >= 2) mbar_wait(&tw_bar, 0);

    ntt_rounds<S, LT, 0>(lo, hi, wlo, whi, tid, nthreads);

    for (int idx = tid; idx < M; idx += nthreads) {
        int tt = idx & (T - 1), j = idx >> LT;
        int pos = (S == 0) ? 0 : (int)(__brev((unsigned)j) >> (32 - (S == 0 ? 1 : S)));
        Fr v = sm_get(lo, hi, ntt_slot<LT>(pos, tt));
        unsigned long long go = out_base + tt * p.out_tt + (unsigned long long)j * p.out_el;
        if (p.tw_mode == 3) {
            v = mul(v, load(&p.tw_pass[go]));
        } else if (p.tw_mode != 0) {
            unsigned long long c = (unsigned long long)tile * T + tt;
            unsigned long long e = c * (unsigned long long)j * p.tw_mul;
            if (e != 0) v = mul(v, ntt_twiddle(p, e));
        }
        if constexpr (DIST) {
            // go = column j2 of the local transform -> row `rank` of its owner's exchange buffer
            Fr* dst = dx.peer[go >> dx.col_bits];
            store256(&dst[dx.row_off + (go & ((1ull << dx.col_bits) - 1))], v);
        } else {
            if (p.out_mode == 1) v = mul(v, p.oscale[(unsigned int)(go % 3ull)]);
            store(&out[go], v);
        }
    }
}

// Pass-ordered twiddles (tw_mode 3) of a strided-column pass over groups of L*M elements: output position o = g L M + j M + c
// gets w_N^(c j Lprod).  One-time per plan and pass.
__global__ void k_pass_twiddles(Fr* out, unsigned long long n, unsigned long long M, unsigned long long LM, unsigned long long lprod,
                                const Fr* hi, const Fr* lo, unsigned int bits) {
    unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    unsigned long long gi = o % LM;
    unsigned long long e = (gi % M) * (gi / M) * lprod;
    store(&out[o], e == 0 ? Fr::one() : mul(load(&hi[e >> bits]), load(&lo[e & ((1ull << bits) - 1)])));
}

// Inter-stage twiddles of rank q, resident per plan: out[(i1 - 1) C + c] = w_N^(i1 (q C + c)) for 1 <= i1 < W, c < C, from the
// two-level tables hi[e >> bits] * lo[e & mask] (one-time per plan and rank)
__global__ void k_dist_twiddles(Fr* out, unsigned long long C, unsigned int rows, unsigned long long col0, const Fr* hi, const Fr* lo,
                                unsigned int bits) {
    unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= C * rows) return;
    unsigned long long i1 = id / C + 1, c = id % C;
    unsigned long long e = i1 * (col0 + c);
    store(&out[id], mul(load(&hi[e >> bits]), load(&lo[e & ((1ull << bits) - 1)])));
}

// Second stage of the multi-GPU transform: a W-point transform ACROSS the ranks for every column of this rank's exchange
// buffer z[W][C] (row i1 came from rank i1, still to be multiplied by w_N^(i1 j2), j2 = rank C + c: table tw), root
// w_W = w_N^(N/W).  Output j1 of column c is element
// A[j1 * (N/W) + rank * C + c] of the result and is stored into rank j1's output block - the second exchange, again as
// peer stores (one thread per column: a warp writes 1 KiB contiguous per destination).
template <int LW>
struct NttCrossArgs {
    const Fr* z;
    const Fr* tw;  // (W - 1) x C inter-stage twiddles (k_dist_twiddles)
    Fr* peer_out[1 << LW];
    unsigned long long C;
    unsigned long long out_off;  // rank * C
    // pipelined exchange: this launch covers the columns s * period + chunk_off + u, s < C / period, u < chunk_len - the columns
    // a contiguous range of CTAs of every rank's exchange pass has produced (period = M / L of that pass).  Unchunked:
    // period = chunk_len = C, chunk_off = 0.
    unsigned long long period, chunk_off, chunk_len;
    Fr w[(1 << LW) / 2 < 1 ? 1 : (1 << LW) / 2];  // w_W^i, i < W/2
};
template <int LW>
__global__ void __launch_bounds__(128) k_ntt_cross(const __grid_constant__ NttCrossArgs<LW> a) {
    constexpr int W = 1 << LW;
    const unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long srow = id / a.chunk_len;  // sub-row of the exchange pass' output inside this rank's column range
    const unsigned long long c = srow * a.period + a.chunk_off + (id - srow * a.chunk_len);
    if (srow * a.period >= a.C) return;
    Fr x[W];
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = load256(&a.z[(unsigned long long)i * a.C + c]);
#pragma unroll
    for (int i = 1; i < W; i++) x[i] = mul(x[i], load256(&a.tw[(unsigned long long)(i - 1) * a.C + c]));
    // decimation in frequency; position p ends up holding output bit-reverse(p)
#pragma unroll
    for (int u = 0; u < LW; u++) {
        const int half = W >> (u + 1);
#pragma unroll
        for (int e = 0; e < W; e++) {
            if (e & half) continue;
            Fr s = add(x[e], x[e + half]);
            Fr d = sub(x[e], x[e + half]);
            const int ex = (e & (half - 1)) << u;
            x[e] = s;
            x[e + half] = ex == 0 ? d : mul(d, a.w[ex]);
        }
    }
#pragma unroll
    for (int p = 0; p < W; p++) {
        int j1 = 0;
#pragma unroll
        for (int b = 0; b < LW; b++) j1 |= ((p >> b) & 1) << (LW - 1 - b);
        store256(&a.peer_out[j1][a.out_off + c], x[p]);
    }
}

// ---- ordering of the exchange stages across PROCESSES (one process per GPU): flags in peer memory ----------------------------
// Every rank owns an array of u32 flags that the others can store into (CUDA IPC mapping).  slot(k, r) = k * 8 + r holds the
// epoch (a per-call counter, the same on all ranks) up to which rank r has completed step k: steps 0 .. K - 1 are the chunks
// of the exchange pass, step DE_DIST_DONE_STEP the whole cross stage.  A signal is a system-scope release store after a system
// fence (the data stores of the preceding kernels on the same stream have completed); a wait spins on acquire loads, gives
// up after ~2 s and raises *error instead of hanging the GPU.
#define DE_DIST_MAX_CHUNKS 8
#define DE_DIST_DONE_STEP DE_DIST_MAX_CHUNKS
#define DE_DIST_FLAG_WORDS ((DE_DIST_MAX_CHUNKS + 1) * 8)
struct NttFlagPeers {
    unsigned int* flags[8];
};
__global__ void k_flag_signal(const __grid_constant__ NttFlagPeers peers, unsigned int world, unsigned int slot, unsigned int epoch) {
    if (threadIdx.x >= world) return;
    __threadfence_system();
    unsigned int* f = peers.flags[threadIdx.x] + slot;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__global__ void k_flag_wait(const unsigned int* flags, unsigned int world, unsigned int step, unsigned int epoch, unsigned int* error) {
    if (threadIdx.x >= world) return;
    const unsigned int* f = flags + step * 8 + threadIdx.x;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int)(v - epoch) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) {  // 2 s: a peer never arrived
            atomicExch(error, 1u);
            break;
        }
        __nanosleep(200);
    }
}

// Deal a contiguous block of the natural-order vector round-robin to the ranks (host entry point de_ntt_sharded: every GPU
// receives the block a[rank M, (rank + 1) M) over its own PCIe link and forwards element rank M + u to rank u mod W, slot
// rank C + u / W of its cyclic input slice): thread (q, v) reads stage[v W + q], stores peer_x[q][rank C + v] (coalesced stores).
struct NttDealArgs {
    const Fr* stage;
    Fr* peer_x[8];
    unsigned int log_w;
    unsigned long long C, M, row_off;
};
__global__ void __launch_bounds__(256) k_ntt_deal(const __grid_constant__ NttDealArgs a) {
    unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.M) return;
    unsigned long long q = id / a.C, v = id - q * a.C;
    store256(&a.peer_x[q][a.row_off + v], load256(&a.stage[(v << a.log_w) + q]));
}

// out[i] = base^(i * step) for i < n  (twiddle tables; one-time per plan)
__global__ void k_pow_table(Fr* out, unsigned long long n, Fr base, unsigned long long step) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // exponent i*step can exceed 64 bits only if the caller misuses it; tables here keep it < 2^60
    unsigned long long e = i * step;
    Fr acc = Fr::one(), cur = base;
    while (e) {
        if (e & 1) acc = mul(acc, cur);
        cur = sqr(cur);
        e >>= 1;
    }
    store(&out[i], acc);
}

// Fr table -> the split-plane layout of NttShape's shared-memory twiddle area: [low 16 B of every entry | high 16 B]
__global__ void k_split_planes(const Fr* in, uint4* out, unsigned int n) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = load(&in[i]);
    out[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    out[n + i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// element-wise helpers used by the domain operations
__global__ void k_scale_periodic(Fr* a, unsigned long long n, const Fr* table, unsigned int period_mask) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store(&a[i], mul(load(&a[i]), load(&table[i & period_mask])));
}

}  // namespace de
