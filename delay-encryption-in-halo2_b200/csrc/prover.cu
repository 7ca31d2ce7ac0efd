// prover.cu — halo2_proofs::plonk::create_proof for one circuit instance (KZG commitments, GWC multi-open, Blake2b
// transcript), with every polynomial resident in HBM from the moment the advice columns are uploaded until the proof
// bytes are complete (reached from the reference at benches/delay_enc.rs:123-131,
// benches/mod_pow.rs:201-209, benches/pose_enc.rs:127-135).
//
// What runs where:
//   host    the transcript (csrc/transcript.hpp), a few dozen scalar Fr operations on challenges, sequencing
//   device  blinding rows, every commitment (msm.cu), theta-compression of the lookup expressions (the GraphEvaluator
//           interpreter of eval.cuh on the n lagrange rows), lookup::prover::permute_expression_pair (bitonic sort of the
//           canonical values + binary-search matching + compaction), the permutation and lookup grand products (fractions,
//           batched inversion, prefix-product scan), all transforms and the quotient (ntt.cu, eval.cu), the opening
//           evaluations and Kate divisions (poly.cu), the linear combinations of the multi-open
// The host reads back only what the transcript hashes: 64 bytes per commitment and 32 bytes per evaluation.
//
// The order of commitments, challenges, evaluations, opening queries and RNG draws follows create_proof as recorded in
// SURVEY.md Appendix B / E; tests/test_gpu_prover.py compares the proof bytes with the CPU restatement of the same algorithm.
// Fr::random(rng) draws are an INPUT (`randoms`, in draw order): the GPU never touches an RNG.
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "ec.cuh"
#include "eval.cuh"
#include "poly.cuh"
#include "transcript.hpp"

namespace de {

// msm.cu
int commit_canonical_dev(de_params* p, int basis, const Fr* d_scalars, size_t stride, size_t n, size_t count, uint8_t* out_xy);
int commit_canonical_mixed_dev(de_params* p, const Fr* d_scalars, size_t stride, size_t n, size_t count_lagrange, size_t count_coeff, uint8_t* out_xy);
int scan_u32(de_ctx* ctx, const unsigned int* in, unsigned long long n, unsigned int* out, unsigned int* block_sums, unsigned int* grand_total);
de_ctx* params_ctx(de_params* p);
size_t params_n(de_params* p);

#define DE_SORT_TILE 1024
#define DE_INV_CHUNK 16
#define DE_PP_CHUNK 32
#define DE_PP_SCAN_THREADS 512
#define DE_MAX_PERM_COLS 16
#define DE_MAX_SETS 8
#define DE_MAX_TAILS 48

// ---- blinding rows: dst[i] = src[i] for a list of short runs ----------------------------------------------------------------
struct TailDescs {
    Fr* dst[DE_MAX_TAILS];
    const Fr* src[DE_MAX_TAILS];
    unsigned int count[DE_MAX_TAILS];
};
__global__ void k_write_tails(const __grid_constant__ TailDescs d) {
    const unsigned int t = blockIdx.x;
    for (unsigned int i = threadIdx.x; i < d.count[t]; i += blockDim.x) store(&d.dst[t][i], load(&d.src[t][i]));
}

// ---- theta-compressed lookup expressions on the n lagrange rows (plonk::evaluation::evaluate with rot_scale = 1) --------
__global__ void __launch_bounds__(128) k_graph_rows(const __grid_constant__ EvalParams p, const DevGraph* graphs, Fr* out) {
    const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.ext_n) return;
    RowCtx c;
    c.p = &p;
    c.idx = idx;
    c.prev = Fr::zero();
    const Fr v = run_graph(c, graphs[blockIdx.y]);
    store(&out[(unsigned long long)blockIdx.y * p.ext_n + idx], v);
}

// ---- sort of canonical field values (Fr's Ord = order of the canonical integers) ---------------------------------------
__device__ __forceinline__ bool lt256(const Fr& a, const Fr& b) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] != b.l[i]) return a.l[i] < b.l[i];
    }
    return false;
}
// rows >= usable are padded with 2^256 - 1 (above every canonical value) so that they sort to the end
__global__ void k_sort_prepare(const Fr* src, unsigned long long src_stride, Fr* dst, unsigned long long npad, unsigned long long usable) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    Fr v;
    if (i < usable) {
        v = from_mont(load(&src[blockIdx.y * src_stride + i]));
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) v.l[k] = 0xffffffffu;
    }
    store(&dst[blockIdx.y * npad + i], v);
}
__device__ __forceinline__ void cmp_swap(Fr* s, unsigned int i, unsigned int partner, bool ascending) {
    Fr a = load(&s[i]), b = load(&s[partner]);
    const bool swap = ascending ? lt256(b, a) : lt256(a, b);
    if (swap) {
        store(&s[i], b);
        store(&s[partner], a);
    }
}
// mode 0: full bitonic sort of each tile (k = 2 .. tile); mode 1: the in-tile tail (j = tile/2 .. 1) of merge step k_outer
__global__ void __launch_bounds__(DE_SORT_TILE / 2) k_bitonic_tile(Fr* data, unsigned long long npad, unsigned int tile, int mode,
                                                                  unsigned long long k_outer) {
    __shared__ Fr s[DE_SORT_TILE];
    Fr* base = data + blockIdx.y * npad + (unsigned long long)blockIdx.x * tile;
    const unsigned long long gbase = (unsigned long long)blockIdx.x * tile;
    const unsigned int t = threadIdx.x;
    for (unsigned int i = t; i < tile; i += blockDim.x) store(&s[i], load(&base[i]));
    __syncthreads();
    for (unsigned long long k = mode ? k_outer : 2; k <= (mode ? k_outer : (unsigned long long)tile); k <<= 1) {
        const unsigned int j0 = (k >> 1) > (tile >> 1) ? (tile >> 1) : (unsigned int)(k >> 1);
        for (unsigned int j = j0; j > 0; j >>= 1) {
            if (t < tile / 2) {
                const unsigned int i = (t / j) * 2 * j + (t % j);
                cmp_swap(s, i, i + j, ((gbase + i) & k) == 0);
            }
            __syncthreads();
        }
    }
    for (unsigned int i = t; i < tile; i += blockDim.x) store(&base[i], load(&s[i]));
}
__global__ void k_bitonic_global(Fr* data, unsigned long long npad, unsigned long long j, unsigned long long k) {
    unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npad / 2) return;
    Fr* base = data + blockIdx.y * npad;
    const unsigned long long i = (t / j) * 2 * j + (t % j);
    Fr a = load(&base[i]), b = load(&base[i + j]);
    const bool ascending = (i & k) == 0;
    const bool swap = ascending ? lt256(b, a) : lt256(a, b);
    if (swap) {
        store(&base[i], b);
        store(&base[i + j], a);
    }
}

// ---- lookup::prover::permute_expression_pair on the sorted arrays ------------------------------------------------------------
// index of the first element >= v in sorted[0 .. len)
__device__ __forceinline__ unsigned int lower_bound256(const Fr* sorted, unsigned int len, const Fr& v) {
    unsigned int lo = 0, hi = len;
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (lt256(load(&sorted[mid]), v)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// sorted: [inputs of all lookups | tables of all lookups], npad apart.  rep[i] = 1 when sorted input row i repeats row i - 1;
// left[j] = 1 when sorted table element j is NOT consumed by the first occurrence of an input value.
__global__ void k_lookup_flags(const Fr* sorted, unsigned long long npad, unsigned int usable, unsigned int n_lookups, unsigned int* rep,
                               unsigned int* left, int* err) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= usable) return;
    const unsigned int l = blockIdx.y;
    const Fr* A = sorted + (unsigned long long)l * npad;
    const Fr* T = sorted + (unsigned long long)(n_lookups + l) * npad;
    const Fr a = load(&A[i]);
    const bool first_a = i == 0 || a != load(&A[i - 1]);
    rep[(unsigned long long)l * npad + i] = first_a ? 0u : 1u;
    if (first_a) {
        const unsigned int pos = lower_bound256(T, usable, a);
        if (pos >= usable || load(&T[pos]) != a) atomicExch(err, 1);  // Error::ConstraintSystemFailure in the reference
    }
    const Fr tv = load(&T[i]);
    const bool first_t = i == 0 || tv != load(&T[i - 1]);
    bool consumed = false;
    if (first_t) {
        const unsigned int pos = lower_bound256(A, usable, tv);
        consumed = pos < usable && load(&A[pos]) == tv;
    }
    left[(unsigned long long)l * npad + i] = consumed ? 0u : 1u;
}
// The flags of all arrays ([rep of every lookup | left of every lookup], npad apart) are scanned as ONE vector (scan_u32); this
// turns the global exclusive scan into per-array compaction indices and totals: idx[i] -= idx[array start].
// The array heads idx[a * npad] are only READ here (every block of array a, and of array a - 1 for its total, needs them);
// k_segment_zero_heads rewrites them afterwards.
__global__ void k_segment_fixup(unsigned int* idx, unsigned long long npad, unsigned int n_arrays, const unsigned int* d_grand, unsigned int* totals) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int a = blockIdx.y;
    const unsigned int start = idx[a * npad];
    const unsigned int next = (a + 1 < n_arrays) ? idx[(a + 1) * npad] : *d_grand;
    if (i == 0) totals[a] = next - start;
    if (i < npad && i > 0) idx[a * npad + i] -= start;
}
__global__ void k_segment_zero_heads(unsigned int* idx, unsigned long long npad, unsigned int n_arrays) {
    const unsigned int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_arrays) idx[a * npad] = 0;
}
// permuted input = sorted input; permuted table row = the input value at first occurrences; repeated rows are listed in R
__global__ void k_lookup_fill_first(const Fr* sorted, unsigned long long npad, unsigned int usable, const unsigned int* rep,
                                    const unsigned int* rep_idx, Fr* a_out, Fr* s_out, unsigned long long out_stride, unsigned int* R) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= usable) return;
    const unsigned int l = blockIdx.y;
    const Fr v = to_mont(load(&sorted[(unsigned long long)l * npad + i]));
    store(&a_out[(unsigned long long)l * out_stride + i], v);
    if (rep[(unsigned long long)l * npad + i]) R[(unsigned long long)l * npad + rep_idx[(unsigned long long)l * npad + i]] = i;
    else store(&s_out[(unsigned long long)l * out_stride + i], v);
}
// leftover table elements in ascending order fill the repeated rows from the LAST repeated row backwards (the reference pops
// them off the end of its list of repeated rows)
__global__ void k_lookup_fill_rest(const Fr* sorted, unsigned long long npad, unsigned int usable, unsigned int n_lookups,
                                   const unsigned int* left, const unsigned int* left_idx, const unsigned int* totals, const unsigned int* R,
                                   Fr* s_out, unsigned long long out_stride) {
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= usable) return;
    const unsigned int l = blockIdx.y;
    if (!left[(unsigned long long)l * npad + j]) return;
    const unsigned int m = totals[l];  // number of repeated rows
    const unsigned int p = left_idx[(unsigned long long)l * npad + j];
    if (p >= m) return;  // only reachable after a flagged constraint failure
    const unsigned int row = R[(unsigned long long)l * npad + (m - 1 - p)];
    store(&s_out[(unsigned long long)l * out_stride + row], to_mont(load(&sorted[(unsigned long long)(n_lookups + l) * npad + j])));
}

// ---- grand products ------------------------------------------------------------------------------------------------------------
struct PermFracParams {
    const Fr* col[DE_MAX_PERM_COLS];    // lagrange values of the permutation columns, in order
    const Fr* sigma[DE_MAX_PERM_COLS];  // lagrange values of the permutation polynomials
    Fr delta_set_start[DE_MAX_SETS];    // delta^(set * chunk_len)
    const Fr* omega_pows;               // extended_omega^i table of the pk; omega^i = entry i << omega_shift
    unsigned int omega_shift, n_cols, chunk_len;
    unsigned long long n;
    Fr beta, gamma, delta;
    Fr* num;                            // set s at num + s * n
    Fr* den;
};
__global__ void __launch_bounds__(128) k_perm_fractions(const __grid_constant__ PermFracParams p) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const unsigned int s = blockIdx.y;
    const unsigned int c0 = s * p.chunk_len;
    const unsigned int c1 = (c0 + p.chunk_len < p.n_cols) ? c0 + p.chunk_len : p.n_cols;
    Fr dw = mul(mul(p.delta_set_start[s], load(&p.omega_pows[i << p.omega_shift])), p.beta);  // delta^c * omega^i * beta
    Fr num = Fr::one(), den = Fr::one();
    for (unsigned int c = c0; c < c1; c++) {
        const Fr v = load(&p.col[c][i]);
        den = mul(den, add(add(mul(p.beta, load(&p.sigma[c][i])), p.gamma), v));
        num = mul(num, add(add(dw, p.gamma), v));
        dw = mul(dw, p.delta);
    }
    store(&p.num[(unsigned long long)s * p.n + i], num);
    store(&p.den[(unsigned long long)s * p.n + i], den);
}
// lookup l: num = (A + beta)(S + gamma) from the compressed expressions, den = (a' + beta)(s' + gamma) from the permuted ones
__global__ void __launch_bounds__(128) k_lookup_fractions(const Fr* comp, const Fr* a_perm, const Fr* s_perm, unsigned long long n,
                                                          unsigned int n_lookups, Fr beta, Fr gamma, Fr* num, Fr* den) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int l = blockIdx.y;
    const Fr ci = load(&comp[(unsigned long long)l * n + i]), ct = load(&comp[(unsigned long long)(n_lookups + l) * n + i]);
    const Fr a = load(&a_perm[(unsigned long long)l * n + i]), s = load(&s_perm[(unsigned long long)l * n + i]);
    store(&num[(unsigned long long)l * n + i], mul(add(ci, beta), add(ct, gamma)));
    store(&den[(unsigned long long)l * n + i], mul(add(beta, a), add(gamma, s)));
}
// num[i] <- num[i] / den[i] (zero denominators give zero, as ff::BatchInvert leaves them).  Montgomery's trick at CTA scope:
// every thread multiplies up DE_INV_CHUNK denominators, a prefix and a suffix product scan over the 256 thread totals run in the
// same loop, ONE thread inverts the CTA's total (a single active lane: no divergence in the Euclidean loop), and each thread
// gets the inverse of its own total as total^-1 * (product of the totals before it) * (product of those after it).
#define DE_INV_THREADS 256
__global__ void __launch_bounds__(DE_INV_THREADS) k_frac_finish(Fr* num, const Fr* den, unsigned long long total) {
    __shared__ Fr s_pre[DE_INV_THREADS];
    __shared__ Fr s_suf[DE_INV_THREADS];
    __shared__ Fr s_inv;
    const unsigned int tid = threadIdx.x;
    const unsigned long long lo = ((unsigned long long)blockIdx.x * DE_INV_THREADS + tid) * DE_INV_CHUNK;
    const unsigned int cnt = lo >= total ? 0u : (unsigned int)((total - lo) < DE_INV_CHUNK ? (total - lo) : DE_INV_CHUNK);
    Fr pre[DE_INV_CHUNK];
    Fr acc = Fr::one();
    for (unsigned int k = 0; k < cnt; k++) {
        pre[k] = acc;
        const Fr d = load(&den[lo + k]);
        if (!d.is_zero()) acc = mul(acc, d);
    }
    store(&s_pre[tid], acc);
    store(&s_suf[tid], acc);
    __syncthreads();
    // inclusive prefix products in s_pre, inclusive suffix products in s_suf
    for (unsigned int d = 1; d < DE_INV_THREADS; d <<= 1) {
        Fr a = Fr::one(), b = Fr::one();
        const bool hp = tid >= d, hs = tid + d < DE_INV_THREADS;
        if (hp) a = load(&s_pre[tid - d]);
        if (hs) b = load(&s_suf[tid + d]);
        __syncthreads();
        if (hp) store(&s_pre[tid], mul(load(&s_pre[tid]), a));
        if (hs) store(&s_suf[tid], mul(load(&s_suf[tid]), b));
        __syncthreads();
    }
    if (tid == 0) store(&s_inv, inv(load(&s_pre[DE_INV_THREADS - 1])));
    __syncthreads();
    if (cnt == 0) return;
    Fr ai = load(&s_inv);
    if (tid > 0) ai = mul(ai, load(&s_pre[tid - 1]));
    if (tid + 1 < DE_INV_THREADS) ai = mul(ai, load(&s_suf[tid + 1]));
    // ai = 1 / (this thread's product of non-zero denominators)
    for (int k = (int)cnt - 1; k >= 0; k--) {
        const Fr d = load(&den[lo + k]);
        Fr r = Fr::zero();
        if (!d.is_zero()) {
            r = mul(mul(ai, pre[k]), load(&num[lo + k]));
            ai = mul(ai, d);
        }
        store(&num[lo + k], r);
    }
}
// prefix products: z[0] = carry, z[i] = carry * prod_{r < i} frac[r]
__global__ void __launch_bounds__(128) k_pp_chunks(const Fr* frac, unsigned long long n, unsigned long long nchunks, Fr* cp) {
    const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const Fr* f = frac + blockIdx.y * n + c * DE_PP_CHUNK;
    const unsigned int cnt = (unsigned int)((n - c * DE_PP_CHUNK) < DE_PP_CHUNK ? (n - c * DE_PP_CHUNK) : DE_PP_CHUNK);
    Fr acc = load(&f[0]);
    for (unsigned int k = 1; k < cnt; k++) acc = mul(acc, load(&f[k]));
    store(&cp[blockIdx.y * nchunks + c], acc);
}
// exclusive prefix product of the chunk products, one CTA per column
__global__ void __launch_bounds__(DE_PP_SCAN_THREADS) k_pp_scan(const Fr* cp, unsigned long long nchunks, Fr* cpfx) {
    __shared__ Fr sm[DE_PP_SCAN_THREADS];
    const Fr* src = cp + blockIdx.x * nchunks;
    Fr* dst = cpfx + blockIdx.x * nchunks;
    const unsigned int tid = threadIdx.x;
    const unsigned long long seg = (nchunks + DE_PP_SCAN_THREADS - 1) / DE_PP_SCAN_THREADS;
    const unsigned long long lo = tid * seg;
    unsigned long long hi = lo + seg;
    if (hi > nchunks) hi = nchunks;
    Fr tot = Fr::one();
    for (unsigned long long i = lo; i < hi; i++) tot = mul(tot, load(&src[i]));
    store(&sm[tid], tot);
    __syncthreads();
    for (unsigned int d = 1; d < DE_PP_SCAN_THREADS; d <<= 1) {
        Fr o = Fr::one();
        const bool has = tid >= d;
        if (has) o = load(&sm[tid - d]);
        __syncthreads();
        if (has) store(&sm[tid], mul(load(&sm[tid]), o));
        __syncthreads();
    }
    Fr run = tid ? load(&sm[tid - 1]) : Fr::one();
    for (unsigned long long i = lo; i < hi; i++) {
        const Fr v = load(&src[i]);
        store(&dst[i], run);
        run = mul(run, v);
    }
}
// carry[col]: the permutation sets are chained (z of set s starts at the value of set s - 1 at row u = n - blinding - 1); the
// lookup products start at one.  One thread.
__global__ void k_pp_carries(const Fr* frac, const Fr* cpfx, unsigned long long n, unsigned long long nchunks, unsigned int n_sets,
                             unsigned int n_cols, unsigned long long u, Fr* carry) {
    if (blockIdx.x || threadIdx.x) return;
    Fr c = Fr::one();
    for (unsigned int s = 0; s < n_cols; s++) {
        if (s >= n_sets) c = Fr::one();
        store(&carry[s], c);
        if (s + 1 < n_sets) {
            const unsigned long long cu = u / DE_PP_CHUNK;
            Fr v = mul(c, load(&cpfx[s * nchunks + cu]));
            for (unsigned long long r = cu * DE_PP_CHUNK; r < u; r++) v = mul(v, load(&frac[s * n + r]));
            c = v;  // z_s[u]
        }
    }
}
struct ZDest {
    Fr* z[DE_MAX_SETS + 16];
};
__global__ void __launch_bounds__(128) k_pp_write(const Fr* frac, const Fr* cpfx, const Fr* carry, unsigned long long n, unsigned long long nchunks,
                                                  const __grid_constant__ ZDest dst) {
    const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const Fr* f = frac + blockIdx.y * n + c * DE_PP_CHUNK;
    Fr* z = dst.z[blockIdx.y] + c * DE_PP_CHUNK;
    const unsigned int cnt = (unsigned int)((n - c * DE_PP_CHUNK) < DE_PP_CHUNK ? (n - c * DE_PP_CHUNK) : DE_PP_CHUNK);
    Fr run = mul(load(&cpfx[blockIdx.y * nchunks + c]), load(&carry[blockIdx.y]));
    for (unsigned int k = 0; k < cnt; k++) {
        store(&z[k], run);
        run = mul(run, load(&f[k]));
    }
}

// ---- vanishing argument / multi-open helpers --------------------------------------------------------------------------------------
// out[j] = sum_i xn^i * h[i * n + j]   (h(X) as one polynomial of degree < n in X after substituting X^n -> xn)
__global__ void k_fold_pieces(const Fr* h, unsigned long long n, unsigned int pieces, Fr xn, Fr* out) {
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Fr acc = load(&h[(unsigned long long)(pieces - 1) * n + j]);
    for (int i = (int)pieces - 2; i >= 0; i--) acc = add(mul(acc, xn), load(&h[(unsigned long long)i * n + j]));
    store(&out[j], acc);
}
// out[j] = sum_i v^i * polys[i][j]  (Horner from the last polynomial)
__global__ void __launch_bounds__(128) k_lincomb(const Fr* const* polys, unsigned int count, Fr v, unsigned long long n, Fr* out) {
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Fr acc = load(&polys[count - 1][j]);
    for (int i = (int)count - 2; i >= 0; i--) acc = add(mul(acc, v), load(&polys[i][j]));
    store(&out[j], acc);
}
}  // namespace de

using namespace de;
using host::HFr;

struct de_prover {
    de_ctx* ctx;
    de_params* params;
    de_pk* pk;
    de_domain* dom;
    uint32_t A, I, Z, L, F, P, bf, chunk, deg, n_cols;
    size_t n, ext_n, usable, npad;
    std::vector<uint32_t> aq_col, fq_col;
    std::vector<int32_t> aq_rot, fq_rot;
    std::vector<uint32_t> perm_kind, perm_index;
    uint8_t transcript_repr[32];
    HFr omega, omega_inv, delta;
    // device state
    Fr *lag, *coef, *fixed_lag, *comp, *sorted, *frac_num, *frac_den, *cp, *cpfx, *carry, *randoms, *h, *hx, *open_acc, *open_q, *kate_scratch;
    Fr *d_points, *d_evals;
    unsigned int *rep, *rep_idx, *left, *left_idx, *totals, *R, *scan_scratch;
    int* d_err;
    DevGraph* d_lookup_graphs;  // inputs of all lookups, then tables
    const Fr** d_eval_polys;
    unsigned int* d_eval_pidx;
    const Fr** d_open_polys;    // all sets concatenated
    std::vector<int32_t> rots;                 // distinct rotations in first-appearance order of the opening queries
    std::vector<uint32_t> open_first, open_count;  // per point set: slice of d_open_polys
    size_t n_evals;
    size_t n_random;
    std::vector<void*> allocs;
    std::string err;
    // second stream: lagrange -> coefficient -> coset transforms of a column block run while the block is being committed
    cudaStream_t st_b = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    bool overlap = true;
};

namespace {
struct StreamSwap {  // the NTT / evaluator entry points launch on ctx->stream
    de_ctx* c;
    cudaStream_t old;
    StreamSwap(de_ctx* ctx, cudaStream_t s) : c(ctx), old(ctx->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = old; }
};
}  // namespace

namespace {

// column order of the per-proof blocks `lag` / `coef` and of the pk's coset workspace: [advice | instance | a' | s' | permz |
// lookup z] (+ one spare column in `lag`)
size_t off_advice(const de_prover*) { return 0; }
size_t off_instance(const de_prover* p) { return p->A; }
size_t off_lookup_a(const de_prover* p) { return p->A + p->I; }
size_t off_lookup_s(const de_prover* p) { return p->A + p->I + p->L; }
size_t off_permz(const de_prover* p) { return p->A + p->I + 2 * (size_t)p->L; }
size_t off_lookup_z(const de_prover* p) { return p->A + p->I + 2 * (size_t)p->L + p->Z; }

HFr to_hfr(const de_fr& v) {
    HFr r;
    memcpy(r.l, v.l, 32);
    return r;
}
de_fr to_defr(const HFr& v) {
    de_fr r;
    memcpy(r.l, v.l, 32);
    return r;
}
Fr dev_fr(const HFr& v) { return fr_from_host(to_defr(v)); }

// x * omega^rot
HFr rotate_omega(const de_prover* p, const HFr& x, int32_t rot) {
    if (rot == 0) return x;
    const HFr w = rot > 0 ? host::fr_pow(p->omega, (uint64_t)rot) : host::fr_pow(p->omega_inv, (uint64_t)(-(int64_t)rot));
    return host::fr_mul(x, w);
}

int bitonic_sort(de_ctx* ctx, Fr* data, size_t npad, unsigned int arrays) {
    const unsigned int tile = npad < DE_SORT_TILE ? (unsigned int)npad : DE_SORT_TILE;
    const dim3 tgrid((unsigned int)(npad / tile), arrays);
    k_bitonic_tile<<<tgrid, DE_SORT_TILE / 2, 0, ctx->stream>>>(data, npad, tile, 0, 0);
    DE_CHECK_LAUNCH(ctx);
    for (unsigned long long k = 2ull * tile; k <= npad; k <<= 1) {
        for (unsigned long long j = k >> 1; j >= tile; j >>= 1) {
            k_bitonic_global<<<dim3((unsigned int)((npad / 2 + 255) / 256), arrays), 256, 0, ctx->stream>>>(data, npad, j, k);
            DE_CHECK_LAUNCH(ctx);
        }
        k_bitonic_tile<<<tgrid, DE_SORT_TILE / 2, 0, ctx->stream>>>(data, npad, tile, 1, k);
        DE_CHECK_LAUNCH(ctx);
    }
    return DE_OK;
}

}  // namespace

extern "C" {

int de_prover_free(de_prover* p) {
    if (!p) return DE_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (void* a : p->allocs) cudaFree(a);
    if (p->st_b) cudaStreamDestroy(p->st_b);
    if (p->ev_a) cudaEventDestroy(p->ev_a);
    if (p->ev_b) cudaEventDestroy(p->ev_b);
    delete p;
    return DE_OK;
}

size_t de_prover_random_count(de_prover* p) { return p ? p->n_random : 0; }
size_t de_prover_proof_size(de_prover* p) {
    if (!p) return 0;
    const size_t points = p->A + 3 * (size_t)p->L + p->Z + 1 + (p->deg - 1) + p->rots.size();
    return 32 * (points + p->n_evals);
}

int de_prover_create(de_params* params, de_pk* pk, const de_prover_desc* desc, de_prover** out) {
    if (!params || !pk) return DE_ERR_ARG;
    de_ctx* ctx = pk->ctx;
    if (!desc || !out) return fail(ctx, DE_ERR_ARG, "de_prover_create: null pointer");
    *out = nullptr;
    if (params_ctx(params) != ctx) return fail(ctx, DE_ERR_ARG, "de_prover_create: params and pk must belong to the same context");
    if (params_n(params) != pk->n) return fail(ctx, DE_ERR_ARG, "de_prover_create: params.k differs from the domain's k");
    if (pk->n_perm_cols > DE_MAX_PERM_COLS || pk->n_sets > DE_MAX_SETS || pk->n_sets + pk->n_lookups > DE_MAX_SETS + 16)
        return fail(ctx, DE_ERR_UNSUPPORTED, "de_prover_create: too many permutation columns / lookups");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    de_prover* p = new de_prover();
    p->ctx = ctx; p->params = params; p->pk = pk; p->dom = pk->dom;
    const de_domain_view* dv = (const de_domain_view*)pk->dom;
    p->A = pk->n_advice; p->I = pk->n_instance; p->Z = pk->n_sets; p->L = pk->n_lookups; p->F = pk->n_fixed; p->P = pk->n_perm_cols;
    p->bf = pk->blinding; p->chunk = pk->chunk_len; p->deg = dv->j;
    p->n = pk->n; p->ext_n = pk->ext_n; p->usable = pk->n - (pk->blinding + 1);
    p->npad = pk->n;
    p->n_cols = p->A + p->I + p->Z + 3 * p->L;
    p->omega = to_hfr(dv->omega);
    p->omega_inv = to_hfr(dv->omega_inv);
    p->delta = to_hfr(fr_to_host(pk->delta));
    {
        const HFr c = host::fr_from_mont(to_hfr(desc->transcript_repr));
        memcpy(p->transcript_repr, c.l, 32);
    }
    auto bail = [&](int rc) { de_prover_free(p); return rc; };
    {
        const char* e = getenv("DE_PROVER_OVERLAP");
        p->overlap = !(e && e[0] == '0');
        if (cudaStreamCreateWithFlags(&p->st_b, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_b, cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(ctx, DE_ERR_CUDA, "de_prover_create: stream / event creation failed"));
    }
    for (uint32_t i = 0; i < desc->n_advice_queries; i++) {
        if (desc->advice_query_column[i] >= p->A) return bail(fail(ctx, DE_ERR_ARG, "de_prover_create: advice query column out of range"));
        p->aq_col.push_back(desc->advice_query_column[i]);
        p->aq_rot.push_back(desc->advice_query_rotation[i]);
    }
    for (uint32_t i = 0; i < desc->n_fixed_queries; i++) {
        if (desc->fixed_query_column[i] >= p->F) return bail(fail(ctx, DE_ERR_ARG, "de_prover_create: fixed query column out of range"));
        p->fq_col.push_back(desc->fixed_query_column[i]);
        p->fq_rot.push_back(desc->fixed_query_rotation[i]);
    }
    if (p->L && (!desc->lookup_input_graphs || !desc->lookup_table_graphs))
        return bail(fail(ctx, DE_ERR_ARG, "de_prover_create: lookup graphs missing"));
    const size_t n = p->n;
    p->n_random = (size_t)p->A * (p->bf + 2) + (size_t)p->L * (2 * (p->bf + 1) + 2) + (size_t)p->Z * (p->bf + 1) + (size_t)p->L * (p->bf + 1) + n + 1 +
                  (p->deg - 1);
    auto dmalloc = [&](void** ptr, size_t bytes) -> int {
        DE_CUDA(ctx, cudaMalloc(ptr, bytes ? bytes : 16));
        p->allocs.push_back(*ptr);
        return DE_OK;
    };
    const size_t ncol = p->n_cols, zl = (size_t)p->Z + p->L;
    const size_t nchunks = (n + DE_PP_CHUNK - 1) / DE_PP_CHUNK;
    int rc;
#define DE_PALLOC(field, type, count) \
    if ((rc = dmalloc((void**)&p->field, sizeof(type) * (count))) != DE_OK) return bail(rc)
    DE_PALLOC(lag, Fr, (ncol + 1) * n);  // + one column: the random polynomial rides along with the grand products' commit
    DE_PALLOC(coef, Fr, ncol * n);
    DE_PALLOC(fixed_lag, Fr, ((size_t)p->F + p->P) * n + 1);
    DE_PALLOC(comp, Fr, 2 * (size_t)p->L * n + 1);
    DE_PALLOC(sorted, Fr, 2 * (size_t)p->L * p->npad + 1);
    DE_PALLOC(frac_num, Fr, zl * n + 1);
    DE_PALLOC(frac_den, Fr, zl * n + 1);
    DE_PALLOC(cp, Fr, zl * nchunks + 1);
    DE_PALLOC(cpfx, Fr, zl * nchunks + 1);
    DE_PALLOC(carry, Fr, zl + 1);
    DE_PALLOC(randoms, Fr, p->n_random);
    DE_PALLOC(h, Fr, p->ext_n);
    DE_PALLOC(hx, Fr, n);
    // flags / compaction indices: [rep of every lookup | left of every lookup], contiguous so that one scan covers them
    DE_PALLOC(rep, unsigned int, 2 * (size_t)p->L * p->npad + 1);
    DE_PALLOC(rep_idx, unsigned int, 2 * (size_t)p->L * p->npad + 1);
    p->left = p->rep + (size_t)p->L * p->npad;
    p->left_idx = p->rep_idx + (size_t)p->L * p->npad;
    DE_CUDA(ctx, cudaMemsetAsync(p->rep, 0, sizeof(unsigned int) * (2 * (size_t)p->L * p->npad + 1), ctx->stream));  // rows >= usable stay 0
    DE_PALLOC(scan_scratch, unsigned int, 4096 + 64);
    if (2ull * p->L * p->npad > 4096ull * 4096ull) return bail(fail(ctx, DE_ERR_UNSUPPORTED, "de_prover_create: lookup columns too long for the compaction scan"));
    DE_PALLOC(R, unsigned int, (size_t)p->L * p->npad + 1);
    DE_PALLOC(totals, unsigned int, 2 * (size_t)p->L + 1);
    DE_PALLOC(d_err, int, 1);
    // lagrange values of the fixed columns and of the permutation polynomials (pk.fixed_values, pk.permutation.permutations)
    if (p->F + p->P) {
        DE_CUDA(ctx, cudaMemcpyAsync(p->fixed_lag, pk->coeff, sizeof(Fr) * n * ((size_t)p->F + p->P), cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = de_coeff_to_lagrange_dev(p->dom, (de_fr*)p->fixed_lag, n, (size_t)p->F + p->P)) != DE_OK) return bail(rc);
    }
    p->perm_kind.resize(p->P);
    p->perm_index.resize(p->P);
    if (p->P) {
        DE_CUDA(ctx, cudaMemcpyAsync(p->perm_kind.data(), pk->d_perm_kind, sizeof(uint32_t) * p->P, cudaMemcpyDeviceToHost, ctx->stream));
        DE_CUDA(ctx, cudaMemcpyAsync(p->perm_index.data(), pk->d_perm_index, sizeof(uint32_t) * p->P, cudaMemcpyDeviceToHost, ctx->stream));
        DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // compiled theta-compression programs
    {
        std::vector<DevGraph> g(2 * (size_t)p->L);
        for (uint32_t l = 0; l < p->L; l++) {
            if ((rc = upload_graph(ctx, desc->lookup_input_graphs[l], &g[l], p->allocs)) != DE_OK) return bail(rc);
            if ((rc = upload_graph(ctx, desc->lookup_table_graphs[l], &g[p->L + l], p->allocs)) != DE_OK) return bail(rc);
        }
        DE_PALLOC(d_lookup_graphs, DevGraph, g.size() + 1);
        if (!g.empty()) {
            DE_CUDA(ctx, cudaMemcpyAsync(p->d_lookup_graphs, g.data(), sizeof(DevGraph) * g.size(), cudaMemcpyHostToDevice, ctx->stream));
            DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    // ---- opening plan: the queries of create_proof in order, grouped by rotation in first-appearance order (GWC's
    // construct_intermediate_sets), and the evaluation plan in the order the evaluations are written to the transcript
    std::vector<std::vector<const Fr*>> sets;
    auto rot_slot = [&](int32_t rot) -> size_t {
        for (size_t i = 0; i < p->rots.size(); i++)
            if (p->rots[i] == rot) return i;
        p->rots.push_back(rot);
        sets.emplace_back();
        return p->rots.size() - 1;
    };
    auto open = [&](int32_t rot, const Fr* poly) { sets[rot_slot(rot)].push_back(poly); };
    const int32_t last_rot = -((int32_t)p->bf + 1);
    const Fr* c_adv = p->coef + off_advice(p) * n;
    const Fr* c_permz = p->coef + off_permz(p) * n;
    const Fr* c_lz = p->coef + off_lookup_z(p) * n;
    const Fr* c_la = p->coef + off_lookup_a(p) * n;
    const Fr* c_ls = p->coef + off_lookup_s(p) * n;
    const Fr* c_fixed = pk->coeff;
    const Fr* c_sigma = pk->coeff + (size_t)p->F * n;
    const Fr* random_poly = p->randoms + (p->n_random - (p->deg - 1) - 1 - n);
    for (size_t i = 0; i < p->aq_col.size(); i++) open(p->aq_rot[i], c_adv + (size_t)p->aq_col[i] * n);
    for (uint32_t s = 0; s < p->Z; s++) {
        open(0, c_permz + (size_t)s * n);
        open(1, c_permz + (size_t)s * n);
    }
    for (int s = (int)p->Z - 2; s >= 0; s--) open(last_rot, c_permz + (size_t)s * n);
    for (uint32_t l = 0; l < p->L; l++) {
        open(0, c_lz + (size_t)l * n);
        open(0, c_la + (size_t)l * n);
        open(0, c_ls + (size_t)l * n);
        open(-1, c_la + (size_t)l * n);
        open(1, c_lz + (size_t)l * n);
    }
    for (size_t i = 0; i < p->fq_col.size(); i++) open(p->fq_rot[i], c_fixed + (size_t)p->fq_col[i] * n);
    for (uint32_t c = 0; c < p->P; c++) open(0, c_sigma + (size_t)c * n);
    open(0, p->hx);
    open(0, random_poly);
    if (p->rots.size() > 16) return bail(fail(ctx, DE_ERR_UNSUPPORTED, "de_prover_create: more than 16 distinct rotations"));
    std::vector<const Fr*> flat;
    for (auto& s : sets) {
        p->open_first.push_back((uint32_t)flat.size());
        p->open_count.push_back((uint32_t)s.size());
        flat.insert(flat.end(), s.begin(), s.end());
    }
    std::vector<const Fr*> ev_polys;
    std::vector<unsigned int> ev_pidx;
    auto ev = [&](const Fr* poly, int32_t rot) {
        ev_polys.push_back(poly);
        ev_pidx.push_back((unsigned int)rot_slot(rot));
    };
    for (size_t i = 0; i < p->aq_col.size(); i++) ev(c_adv + (size_t)p->aq_col[i] * n, p->aq_rot[i]);
    for (size_t i = 0; i < p->fq_col.size(); i++) ev(c_fixed + (size_t)p->fq_col[i] * n, p->fq_rot[i]);
    ev(random_poly, 0);
    for (uint32_t c = 0; c < p->P; c++) ev(c_sigma + (size_t)c * n, 0);
    for (uint32_t s = 0; s < p->Z; s++) {
        ev(c_permz + (size_t)s * n, 0);
        ev(c_permz + (size_t)s * n, 1);
        if (s + 1 < p->Z) ev(c_permz + (size_t)s * n, last_rot);
    }
    for (uint32_t l = 0; l < p->L; l++) {
        ev(c_lz + (size_t)l * n, 0);
        ev(c_lz + (size_t)l * n, 1);
        ev(c_la + (size_t)l * n, 0);
        ev(c_la + (size_t)l * n, -1);
        ev(c_ls + (size_t)l * n, 0);
    }
    p->n_evals = ev_polys.size();
    const size_t n_sets_open = p->rots.size();
    DE_PALLOC(open_acc, Fr, n_sets_open * n);
    DE_PALLOC(open_q, Fr, n_sets_open * n);
    DE_PALLOC(kate_scratch, Fr, kate_scratch_elems(n, n_sets_open) + n_sets_open * 2);
    DE_PALLOC(d_points, Fr, 16);
    DE_PALLOC(d_evals, Fr, p->n_evals + 1);
    DE_PALLOC(d_eval_polys, const Fr*, p->n_evals + 1);
    DE_PALLOC(d_eval_pidx, unsigned int, p->n_evals + 1);
    DE_PALLOC(d_open_polys, const Fr*, flat.size() + n_sets_open + 1);
#undef DE_PALLOC
    // after the concatenated query lists: one pointer per set to its accumulator (the Kate division's inputs)
    for (size_t s = 0; s < n_sets_open; s++) flat.push_back(p->open_acc + s * n);
    DE_CUDA(ctx, cudaMemcpyAsync(p->d_eval_polys, ev_polys.data(), sizeof(const Fr*) * ev_polys.size(), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(p->d_eval_pidx, ev_pidx.data(), sizeof(unsigned int) * ev_pidx.size(), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(p->d_open_polys, flat.data(), sizeof(const Fr*) * flat.size(), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return DE_OK;
}

// the proof once the advice columns and the random stream are in p->lag / p->randoms
static int prove_core(de_prover* p, const de_fr* const* instances, const size_t* instance_lens, uint8_t* proof_out, size_t proof_cap,
                      size_t* proof_len) {
    de_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    const size_t n = p->n, bf = p->bf, usable = p->usable;
    const uint32_t A = p->A, I = p->I, Z = p->Z, L = p->L;
    de_pk* pk = p->pk;
    host::TranscriptWriter tr;
    std::vector<uint8_t> xy(64 * 64);
    // DE_PROVER_TRACE=1: wall-clock time of each phase on stderr (every phase below ends in a stream sync, except where noted)
    static const bool trace = getenv("DE_PROVER_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_prev = trace ? now() : 0.0;
    std::string trace_line;
    auto mark = [&](const char* name) {
        if (!trace) return;
        const double t = now();
        char buf[64];
        snprintf(buf, sizeof(buf), " %s=%.3f", name, t - t_prev);
        trace_line += buf;
        t_prev = t;
    };

    tr.common_scalar(p->transcript_repr);
    for (uint32_t i = 0; i < I; i++) {
        if (instance_lens[i] > usable) return fail(ctx, DE_ERR_ARG, "de_create_proof: InstanceTooLarge");
        Fr* col = p->lag + (off_instance(p) + i) * n;
        DE_CUDA(ctx, cudaMemsetAsync(col, 0, sizeof(Fr) * n, st));
        if (instance_lens[i]) DE_CUDA(ctx, cudaMemcpyAsync(col, instances[i], sizeof(Fr) * instance_lens[i], cudaMemcpyHostToDevice, st));
        for (size_t r = 0; r < instance_lens[i]; r++) {
            const HFr c = host::fr_from_mont(to_hfr(instances[i][r]));
            tr.common_scalar((const uint8_t*)c.l);
        }
    }
    // ---- random stream cursor (draw order of create_proof; blinds are drawn by the reference but unused by KZG)
    size_t rpos = 0;
    TailDescs tails;
    unsigned int n_tails = 0;
    int* tail_rc_p = nullptr;
    auto flush_tails = [&]() -> int {
        if (tail_rc_p && *tail_rc_p != DE_OK) return *tail_rc_p;  // an earlier launch of this round failed
        if (n_tails) {
            k_write_tails<<<n_tails, 32, 0, st>>>(tails);
            DE_CHECK_LAUNCH(ctx);
        }
        n_tails = 0;
        return DE_OK;
    };
    int tail_rc = DE_OK;  // a failed intermediate flush is reported by the flush that ends the round (DE_TRY(flush_tails()))
    tail_rc_p = &tail_rc;
    auto add_tail = [&](Fr* dst, size_t count) {
        if (n_tails == DE_MAX_TAILS) {
            const int rc = flush_tails();
            if (rc != DE_OK && tail_rc == DE_OK) tail_rc = rc;
        }
        tails.dst[n_tails] = dst;
        tails.src[n_tails] = p->randoms + rpos;
        tails.count[n_tails] = (unsigned int)count;
        n_tails++;
        rpos += count;
    };
    auto commit = [&](int basis, const Fr* d, size_t count) -> int {
        if (count > 64) return fail(ctx, DE_ERR_UNSUPPORTED, "de_create_proof: more than 64 commitments in one round");
        DE_TRY(commit_canonical_dev(p->params, basis, d, n, n, count, xy.data()));
        for (size_t i = 0; i < count; i++)
            if (!tr.write_point(xy.data() + 64 * i)) return fail(ctx, DE_ERR_ARG, "de_create_proof: a commitment is the point at infinity");
        return DE_OK;
    };

    // columns [c0, c0 + cnt) are final in lagrange form on `st`: coefficient form and extended coset on the second stream
    // (neither depends on a transcript challenge), overlapping the same columns' commitment
    // In throughput mode the blocks are deferred and all columns go through ONE batched transform pair after the last block
    // (23 columns keep the NTT kernels at 88 % of the multiply peak, batches of 7 / 10 / 7 at 70-75 %); in latency mode each
    // block starts at once so that its transforms hide under the block's commitment.
    const bool defer = ctx->mode == DE_MODE_THROUGHPUT;
    auto to_cosets = [&](size_t c0, size_t cnt, bool last_block) -> int {
        if (defer) {
            if (!last_block) return DE_OK;
            c0 = 0;
            cnt = p->n_cols;
        }
        const size_t work_c0 = c0;
        if (!cnt) return DE_OK;
        cudaStream_t sb = p->overlap ? p->st_b : st;
        if (p->overlap) {
            DE_CUDA(ctx, cudaEventRecord(p->ev_a, st));
            DE_CUDA(ctx, cudaStreamWaitEvent(sb, p->ev_a, 0));
        }
        StreamSwap sw(ctx, sb);
        DE_CUDA(ctx, cudaMemcpyAsync(p->coef + c0 * n, p->lag + c0 * n, sizeof(Fr) * n * cnt, cudaMemcpyDeviceToDevice, sb));
        DE_TRY(de_lagrange_to_coeff_dev(p->dom, (de_fr*)(p->coef + c0 * n), n, cnt));
        DE_TRY(de_coeff_to_extended_dev(p->dom, (const de_fr*)(p->coef + c0 * n), n, (de_fr*)(pk->work + work_c0 * p->ext_n), p->ext_n, cnt));
        return DE_OK;
    };

    // ---- advice: blinding rows, commitments
    for (uint32_t a = 0; a < A; a++) add_tail(p->lag + (off_advice(p) + a) * n + usable, bf + 1);
    rpos += A;  // advice blinds
    DE_TRY(flush_tails());
    DE_TRY(to_cosets(off_advice(p), (size_t)A + I, false));
    DE_TRY(commit(1, p->lag + off_advice(p) * n, A));
    mark("advice_commit");
    const HFr theta = tr.squeeze_challenge();

    // ---- lookups: compress, permute, commit
    Fr* a_perm = p->lag + off_lookup_a(p) * n;
    Fr* s_perm = p->lag + off_lookup_s(p) * n;
    if (L) {
        EvalParams ep;
        memset(&ep, 0, sizeof(ep));
        ep.ext_mask = (uint32_t)(n - 1);
        ep.rot_scale = 1;
        ep.ext_n = n;
        ep.fixed = p->fixed_lag;
        ep.advice = p->lag + off_advice(p) * n;
        ep.instance = p->lag + off_instance(p) * n;
        ep.theta = dev_fr(theta);
        k_graph_rows<<<dim3((unsigned int)((n + 127) / 128), 2 * L), 128, 0, st>>>(ep, p->d_lookup_graphs, p->comp);
        DE_CHECK_LAUNCH(ctx);
        k_sort_prepare<<<dim3((unsigned int)((p->npad + 255) / 256), 2 * L), 256, 0, st>>>(p->comp, n, p->sorted, p->npad, usable);
        DE_CHECK_LAUNCH(ctx);
        DE_TRY(bitonic_sort(ctx, p->sorted, p->npad, 2 * L));
        DE_CUDA(ctx, cudaMemsetAsync(p->d_err, 0, sizeof(int), st));
        const dim3 ugrid((unsigned int)((usable + 255) / 256), L);
        k_lookup_flags<<<ugrid, 256, 0, st>>>(p->sorted, p->npad, (unsigned int)usable, L, p->rep, p->left, p->d_err);
        DE_CHECK_LAUNCH(ctx);
        // compaction indices of the 2L flag arrays by one global scan + per-array fix-up
        DE_TRY(scan_u32(ctx, p->rep, 2ull * L * p->npad, p->rep_idx, p->scan_scratch, p->scan_scratch + 4096));
        k_segment_fixup<<<dim3((unsigned int)((p->npad + 255) / 256), 2 * L), 256, 0, st>>>(p->rep_idx, p->npad, 2 * L, p->scan_scratch + 4096, p->totals);
        DE_CHECK_LAUNCH(ctx);
        k_segment_zero_heads<<<1, 64, 0, st>>>(p->rep_idx, p->npad, 2 * L);
        DE_CHECK_LAUNCH(ctx);
        k_lookup_fill_first<<<ugrid, 256, 0, st>>>(p->sorted, p->npad, (unsigned int)usable, p->rep, p->rep_idx, a_perm, s_perm, n, p->R);
        DE_CHECK_LAUNCH(ctx);
        k_lookup_fill_rest<<<ugrid, 256, 0, st>>>(p->sorted, p->npad, (unsigned int)usable, L, p->left, p->left_idx, p->totals, p->R, s_perm, n);
        DE_CHECK_LAUNCH(ctx);
        for (uint32_t l = 0; l < L; l++) {
            add_tail(a_perm + (size_t)l * n + usable, bf + 1);
            add_tail(s_perm + (size_t)l * n + usable, bf + 1);
            rpos += 2;  // permuted input / table blinds
        }
        DE_TRY(flush_tails());
        int h_err = 0;
        DE_CUDA(ctx, cudaMemcpyAsync(&h_err, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
        DE_TRY(to_cosets(off_lookup_a(p), 2 * (size_t)L, false));
        // one launch sequence for the 2L permuted columns (a' block then s' block, adjacent in HBM); the transcript takes
        // them interleaved: a'_0, s'_0, a'_1, s'_1, ...
        std::vector<uint8_t> pas(64 * 2 * (size_t)L);
        DE_TRY(commit_canonical_dev(p->params, 1, a_perm, n, n, 2 * (size_t)L, pas.data()));
        if (h_err) return fail(ctx, DE_ERR_ARG, "de_create_proof: a lookup input is not in its table (ConstraintSystemFailure)");
        for (uint32_t l = 0; l < L; l++)
            if (!tr.write_point(pas.data() + 64 * l) || !tr.write_point(pas.data() + 64 * ((size_t)L + l)))
                return fail(ctx, DE_ERR_ARG, "de_create_proof: a commitment is the point at infinity");
    }
    mark("lookup_permute_commit");
    const HFr beta = tr.squeeze_challenge();
    const HFr gamma = tr.squeeze_challenge();

    // ---- grand products: permutation sets, then lookups
    const size_t zl = (size_t)Z + L;
    if (zl) {
        const size_t nchunks = (n + DE_PP_CHUNK - 1) / DE_PP_CHUNK;
        if (Z) {
            PermFracParams pp;
            memset(&pp, 0, sizeof(pp));
            const std::vector<uint32_t>&kinds = p->perm_kind, &idxs = p->perm_index;
            for (uint32_t c = 0; c < p->P; c++) {
                pp.col[c] = kinds[c] == DE_VAL_ADVICE ? p->lag + (off_advice(p) + idxs[c]) * n
                            : kinds[c] == DE_VAL_INSTANCE ? p->lag + (off_instance(p) + idxs[c]) * n
                                                          : p->fixed_lag + (size_t)idxs[c] * n;
                pp.sigma[c] = p->fixed_lag + ((size_t)p->F + c) * n;
            }
            for (uint32_t s = 0; s < Z; s++) pp.delta_set_start[s] = dev_fr(host::fr_pow(p->delta, (uint64_t)s * p->chunk));
            pp.omega_pows = pk->resident + ((size_t)p->F + p->P + 3) * p->ext_n;
            pp.omega_shift = pk->ek - pk->k;
            pp.n_cols = p->P;
            pp.chunk_len = p->chunk;
            pp.n = n;
            pp.beta = dev_fr(beta); pp.gamma = dev_fr(gamma); pp.delta = dev_fr(p->delta);
            pp.num = p->frac_num;
            pp.den = p->frac_den;
            k_perm_fractions<<<dim3((unsigned int)((n + 127) / 128), Z), 128, 0, st>>>(pp);
            DE_CHECK_LAUNCH(ctx);
        }
        if (L) {
            k_lookup_fractions<<<dim3((unsigned int)((n + 127) / 128), L), 128, 0, st>>>(p->comp, a_perm, s_perm, n, L, dev_fr(beta), dev_fr(gamma),
                                                                                       p->frac_num + (size_t)Z * n, p->frac_den + (size_t)Z * n);
            DE_CHECK_LAUNCH(ctx);
        }
        const size_t total = zl * n;
        const size_t inv_threads = (total + DE_INV_CHUNK - 1) / DE_INV_CHUNK;
        k_frac_finish<<<(unsigned int)((inv_threads + DE_INV_THREADS - 1) / DE_INV_THREADS), DE_INV_THREADS, 0, st>>>(p->frac_num, p->frac_den, total);
        DE_CHECK_LAUNCH(ctx);
        const dim3 cgrid((unsigned int)((nchunks + 127) / 128), (unsigned int)zl);
        k_pp_chunks<<<cgrid, 128, 0, st>>>(p->frac_num, n, nchunks, p->cp);
        DE_CHECK_LAUNCH(ctx);
        k_pp_scan<<<(unsigned int)zl, DE_PP_SCAN_THREADS, 0, st>>>(p->cp, nchunks, p->cpfx);
        DE_CHECK_LAUNCH(ctx);
        k_pp_carries<<<1, 32, 0, st>>>(p->frac_num, p->cpfx, n, nchunks, Z, (unsigned int)zl, usable, p->carry);
        DE_CHECK_LAUNCH(ctx);
        ZDest zd;
        memset(&zd, 0, sizeof(zd));
        for (uint32_t s = 0; s < Z; s++) zd.z[s] = p->lag + (off_permz(p) + s) * n;
        for (uint32_t l = 0; l < L; l++) zd.z[Z + l] = p->lag + (off_lookup_z(p) + l) * n;
        k_pp_write<<<cgrid, 128, 0, st>>>(p->frac_num, p->cpfx, p->carry, n, nchunks, zd);
        DE_CHECK_LAUNCH(ctx);
        for (uint32_t s = 0; s < Z; s++) {
            add_tail(p->lag + (off_permz(p) + s) * n + (n - bf), bf);
            rpos += 1;  // blind
        }
        for (uint32_t l = 0; l < L; l++) {
            add_tail(p->lag + (off_lookup_z(p) + l) * n + (n - bf), bf);
            rpos += 1;
        }
        DE_TRY(flush_tails());
    }
    DE_TRY(to_cosets(off_permz(p), zl, true));  // the last block: in throughput mode this transforms every column at once

    // ---- vanishing argument: the random polynomial.  Its commitment follows the grand products' in the transcript with no
    // challenge in between, so the two rounds share one launch sequence: zl polynomials over g_lagrange and one over g.
    const Fr* random_poly = p->randoms + rpos;
    rpos += n + 1;  // coefficients + blind
    if (rpos + (p->deg - 1) != p->n_random) return fail(ctx, DE_ERR_ARG, "de_create_proof: internal error: random-draw layout out of step");
    {
        // column order in `lag` is [advice | instance | a' | s' | permz | lookup z | spare]: the spare column after the z block
        // takes the random polynomial, so the batch is contiguous
        Fr* batch = p->lag + off_permz(p) * n;
        DE_CUDA(ctx, cudaMemcpyAsync(batch + zl * n, random_poly, sizeof(Fr) * n, cudaMemcpyDeviceToDevice, st));
        if (zl + 1 > 64) return fail(ctx, DE_ERR_UNSUPPORTED, "de_create_proof: more than 64 commitments in one round");
        DE_TRY(commit_canonical_mixed_dev(p->params, batch, n, n, zl, 1, xy.data()));
        for (size_t i = 0; i < zl + 1; i++)
            if (!tr.write_point(xy.data() + 64 * i)) return fail(ctx, DE_ERR_ARG, "de_create_proof: a commitment is the point at infinity");
    }
    mark("products_random_commit");
    const HFr y = tr.squeeze_challenge();

    // ---- quotient: every coset is (being) produced on the second stream
    if (p->overlap) {
        DE_CUDA(ctx, cudaEventRecord(p->ev_b, p->st_b));
        DE_CUDA(ctx, cudaStreamWaitEvent(st, p->ev_b, 0));
    }
    de_challenges ch;
    memset(&ch, 0, sizeof(ch));
    ch.y = to_defr(y); ch.beta = to_defr(beta); ch.gamma = to_defr(gamma); ch.theta = to_defr(theta);
    DE_TRY(de_evaluate_h_rows_dev(pk, &ch, (de_fr*)p->h));
    DE_TRY(de_divide_by_vanishing_dev(p->dom, (de_fr*)p->h, p->ext_n, 1));
    size_t h_len = 0;
    DE_TRY(de_extended_to_coeff_dev(p->dom, (de_fr*)p->h, p->ext_n, 1, &h_len));
    const uint32_t pieces = p->deg - 1;
    rpos += pieces;  // h blinds
    DE_TRY(commit(0, p->h, pieces));
    mark("quotient_commit");
    const HFr x = tr.squeeze_challenge();
    const HFr xn = host::fr_pow(x, (uint64_t)n);

    // ---- evaluations
    std::vector<HFr> points(p->rots.size());
    {
        Fr hp[16];
        for (size_t i = 0; i < p->rots.size(); i++) {
            points[i] = rotate_omega(p, x, p->rots[i]);
            hp[i] = dev_fr(points[i]);
        }
        DE_CUDA(ctx, cudaMemcpyAsync(p->d_points, hp, sizeof(Fr) * p->rots.size(), cudaMemcpyHostToDevice, st));
    }
    k_fold_pieces<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(p->h, n, pieces, dev_fr(xn), p->hx);
    DE_CHECK_LAUNCH(ctx);
    DE_TRY(eval_polynomials_dev(ctx, p->d_eval_polys, n, p->d_eval_pidx, p->d_points, p->n_evals, nullptr, p->d_evals));
    std::vector<uint8_t> evals(32 * p->n_evals);
    DE_CUDA(ctx, cudaMemcpyAsync(evals.data(), p->d_evals, 32 * p->n_evals, cudaMemcpyDeviceToHost, st));
    DE_CUDA(ctx, stream_wait(ctx, st));
    for (size_t i = 0; i < p->n_evals; i++) tr.write_scalar(evals.data() + 32 * i);

    mark("evaluations");
    // ---- ProverGWC: one witness polynomial per distinct point
    const HFr v = tr.squeeze_challenge();
    const size_t n_open = p->rots.size();
    size_t flat_total = 0;
    for (size_t s = 0; s < n_open; s++) flat_total += p->open_count[s];
    std::vector<de_fr> bs(n_open);
    for (size_t s = 0; s < n_open; s++) {
        k_lincomb<<<(unsigned int)((n + 127) / 128), 128, 0, st>>>(p->d_open_polys + p->open_first[s], p->open_count[s], dev_fr(v), n, p->open_acc + s * n);
        DE_CHECK_LAUNCH(ctx);
        bs[s] = to_defr(points[s]);
    }
    // kate_division never reads the constant coefficient, so subtracting the combined evaluation is a no-op for the quotient
    DE_TRY(kate_division_dev(ctx, p->d_open_polys + flat_total, n, bs.data(), n_open, p->open_q, n, p->kate_scratch));
    DE_TRY(commit(0, p->open_q, n_open));

    mark("openings_commit");
    if (trace) fprintf(stderr, "[de_prover]%s\n", trace_line.c_str());
    if (tr.proof.size() > proof_cap) return fail(ctx, DE_ERR_ARG, "de_create_proof: proof buffer too small");
    memcpy(proof_out, tr.proof.data(), tr.proof.size());
    *proof_len = tr.proof.size();
    return DE_OK;
}

// an early error return may leave transforms queued on the second stream: drain both before the buffers are reused
static int finish(de_prover* p, int rc) {
    if (rc != DE_OK) {
        cudaStreamSynchronize(p->st_b);
        cudaStreamSynchronize(p->ctx->stream);
        cudaGetLastError();
    }
    return rc;
}

static int check_args(de_prover* p, const void* advice, const de_fr* const* instances, const size_t* instance_lens, const void* randoms,
                      size_t n_randoms, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    de_ctx* ctx = p->ctx;
    if ((p->A && !advice) || (p->I && (!instances || !instance_lens)) || !randoms || !proof_out || !proof_len)
        return fail(ctx, DE_ERR_ARG, "de_create_proof: null pointer");
    if (n_randoms < p->n_random) return fail(ctx, DE_ERR_ARG, "de_create_proof: not enough random field elements (see de_prover_random_count)");
    if (proof_cap < de_prover_proof_size(p)) return fail(ctx, DE_ERR_ARG, "de_create_proof: proof buffer too small (see de_prover_proof_size)");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return DE_OK;
}

int de_create_proof(de_prover* p, const de_fr* const* advice, const de_fr* const* instances, const size_t* instance_lens, const de_fr* randoms,
                    size_t n_randoms, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!p) return DE_ERR_ARG;
    de_ctx* ctx = p->ctx;
    DE_TRY(check_args(p, advice, instances, instance_lens, randoms, n_randoms, proof_out, proof_cap, proof_len));
    DE_CUDA(ctx, cudaMemcpyAsync(p->randoms, randoms, sizeof(Fr) * p->n_random, cudaMemcpyHostToDevice, ctx->stream));
    for (uint32_t a = 0; a < p->A; a++)
        if (!advice[a]) return fail(ctx, DE_ERR_ARG, "de_create_proof: null advice column");
    // columns that are adjacent in host memory (a staging buffer of A x n elements, what de_circuit_witness fills) go up in ONE
    // copy: a 2 MB copy runs at a third of the link rate (17 of 54 GB/s measured), five of them cost 0.6 ms where one costs 0.2
    for (uint32_t a = 0; a < p->A;) {
        uint32_t run = 1;
        while (a + run < p->A && advice[a + run] == advice[a] + (size_t)run * p->n) run++;
        DE_CUDA(ctx, cudaMemcpyAsync(p->lag + (off_advice(p) + a) * p->n, advice[a], sizeof(Fr) * p->n * run, cudaMemcpyHostToDevice, ctx->stream));
        a += run;
    }
    return finish(p, prove_core(p, instances, instance_lens, proof_out, proof_cap, proof_len));
}

int de_create_proof_dev(de_prover* p, const de_fr* d_advice, size_t advice_stride, const de_fr* const* instances, const size_t* instance_lens,
                        const de_fr* d_randoms, size_t n_randoms, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!p) return DE_ERR_ARG;
    de_ctx* ctx = p->ctx;
    DE_TRY(check_args(p, d_advice, instances, instance_lens, d_randoms, n_randoms, proof_out, proof_cap, proof_len));
    DE_CUDA(ctx, cudaMemcpyAsync(p->randoms, d_randoms, sizeof(Fr) * p->n_random, cudaMemcpyDeviceToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpy2DAsync(p->lag + off_advice(p) * p->n, sizeof(Fr) * p->n, d_advice, sizeof(Fr) * advice_stride, sizeof(Fr) * p->n, p->A,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    return finish(p, prove_core(p, instances, instance_lens, proof_out, proof_cap, proof_len));
}

}  // extern "C"
