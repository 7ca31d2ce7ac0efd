// poly.cu — opening-phase polynomial kernels (the callers' side of the hot path):
//   halo2_proofs::arithmetic::eval_polynomial(poly, point)   -> de_eval_polynomial      (k_eval_polynomial)
//   halo2_proofs::arithmetic::kate_division(a, b)            -> de_kate_division        (k_kate_*)
// create_proof evaluates every queried polynomial at x * omega^rot (58 Horner evaluations for the RSA shape,
// benches/delay_enc.rs:123 -> plonk::prover) and the GWC multi-open divides the combined polynomial of each point by
// (X - point) before committing to the quotient.  prover.cu drives the batched device-side entry points declared in
// poly.cuh; the C ABI functions here are the single-polynomial parity surface.
#include "poly.cuh"

#include "transcript.hpp"

namespace de {

#define DE_POLY_THREADS 512
#define DE_KATE_CHUNK 32
#define DE_KATE_SCAN 512

// tree sum of Fr values in shared memory (adds only); result in sm[0]
__device__ __forceinline__ void block_sum(Fr* sm, int tid, int nthreads) {
    for (int d = nthreads >> 1; d >= 1; d >>= 1) {
        __syncthreads();
        if (tid < d) store(&sm[tid], add(load(&sm[tid]), load(&sm[tid + d])));
    }
    __syncthreads();
}

// grid (evaluation, segment): the SEGS x 512 threads of an evaluation Horner-evaluate contiguous chunks of `chunk` coefficients
// (r_g for global thread g); with X = x^chunk the value is sum_g r_g X^g.  Inside a CTA that sum is a tree whose level d folds
// slot t + d into slot t with ONE multiplication by X^d (the powers X^(2^k) come from a squaring chain of thread 0) - no
// per-thread x^(g * chunk).  A segment leaves its sum in `partial`; the CTA that finishes last (counter `done`, zeroed by the
// host before the launch) folds the SEGS partial sums by Horner in X^512.  58 evaluations of a proof thus occupy 232 SMs' worth
// of CTAs instead of 58, each with a quarter of the serial chain (one proof in flight: 186 -> ~60 us).
// polys[e] points at n coefficients; the result is written in Montgomery form and, when out_canonical != nullptr, also as
// the canonical integer (Fr::to_repr), which is what the transcript hashes.
#define DE_POLY_SEGS 4
__device__ __forceinline__ Fr load_cg(const Fr* p) {  // bypasses L1: written by another CTA of the same launch
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p)), b = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__global__ void __launch_bounds__(DE_POLY_THREADS) k_eval_polynomial(const Fr* const* polys, unsigned long long n, const unsigned int* point_index,
                                                                     const Fr* points, Fr* out, Fr* out_canonical, Fr* partial,
                                                                     unsigned int* done) {
    __shared__ Fr sm[DE_POLY_THREADS];
    __shared__ Fr pw[10];  // pw[k] = X^(2^k), X = x^chunk; pw[9] = X^512 steps from one segment to the next
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const unsigned int e = blockIdx.x, seg = blockIdx.y, segs = gridDim.y;
    const Fr* p = polys[e];
    const Fr x = load(&points[point_index ? point_index[e] : e]);
    const unsigned long long chunk = (n + (unsigned long long)segs * DE_POLY_THREADS - 1) / ((unsigned long long)segs * DE_POLY_THREADS);
    const unsigned long long lo = ((unsigned long long)seg * DE_POLY_THREADS + tid) * chunk;
    unsigned long long hi = lo + chunk;
    if (hi > n) hi = n;
    Fr acc = Fr::zero();
    for (unsigned long long i = hi; i > lo; i--) acc = add(mul(acc, x), load(&p[i - 1]));
    store(&sm[tid], acc);
    if (tid == 0) {
        Fr X = Fr::one(), base = x;  // x^chunk, LSB first
        for (unsigned long long b = chunk; b; b >>= 1) {
            if (b & 1) X = mul(X, base);
            base = sqr(base);
        }
        store(&pw[0], X);
        for (int k = 1; k < 10; k++) {
            X = sqr(X);
            store(&pw[k], X);
        }
    }
    __syncthreads();
    for (int d = DE_POLY_THREADS >> 1, k = 8; d >= 1; d >>= 1, k--) {
        if (tid < d) store(&sm[tid], add(load(&sm[tid]), mul(load(&sm[tid + d]), load(&pw[k]))));
        __syncthreads();
    }
    if (tid == 0) {
        is_last = true;
        if (segs > 1) {
            store(&partial[(unsigned long long)e * segs + seg], load(&sm[0]));
            __threadfence();
            is_last = atomicAdd(&done[e], 1u) == segs - 1;
        }
    }
    __syncthreads();
    if (!is_last || tid != 0) return;
    Fr r = load(&sm[0]);
    if (segs > 1) {
        __threadfence();
        const Fr Y = load(&pw[9]);
        r = load_cg(&partial[(unsigned long long)e * segs + segs - 1]);
        for (int s = (int)segs - 2; s >= 0; s--) r = add(mul(r, Y), load_cg(&partial[(unsigned long long)e * segs + s]));
    }
    if (out) store(&out[e], r);
    if (out_canonical) store(&out_canonical[e], from_mont(r));
}

// kate_division: with t_i = q[i-1] the quotient satisfies t_i = a[i] + b * t_(i+1), t_n = 0.
// pass 1: per chunk c = [lo, hi) of a: V_c = sum_{i in chunk} a[i] * b^(i - lo), so that t_lo = V_c + b^(hi-lo) * t_hi.
// grid.y = polynomial of the batch; b per polynomial.
__global__ void k_kate_chunk_values(const Fr* const* as, unsigned long long n, const Fr* bs, Fr* vals, unsigned long long nchunks) {
    unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const Fr* a = as[blockIdx.y];
    const Fr b = load(&bs[blockIdx.y]);
    unsigned long long lo = c * DE_KATE_CHUNK, hi = lo + DE_KATE_CHUNK;
    if (hi > n) hi = n;
    Fr acc = Fr::zero();
    for (unsigned long long i = hi; i > lo; i--) acc = add(mul(acc, b), load(&a[i - 1]));
    store(&vals[blockIdx.y * nchunks + c], acc);
}
// pass 2 (one CTA per polynomial): carry[c] = t_(hi_c) = V_(c+1) + B * carry[c+1] with B = b^CHUNK, carry[last] = 0: a suffix
// scan of affine maps y -> a + B y that all share the multiplier B.  Two levels: thread t composes the maps of its G =
// ceil(nchunks / DE_KATE_SCAN) consecutive chunks (Horner, G steps), one Hillis-Steele sweep over the threads follows in which
// step d needs ONE multiplication, by (B^G)^d - a squaring chain every thread keeps for itself -, then the thread replays its G
// chunks from the carry entering them.  G + 9 + G dependent steps (n = 2^16: 17) where sweeps of DE_KATE_SCAN single chunks with
// both map components in shared memory took 9 * nchunks / DE_KATE_SCAN + ... (36): 105 -> ~50 us of a single proof's last round.
__global__ void __launch_bounds__(DE_KATE_SCAN) k_kate_carries(const Fr* vals, unsigned long long nchunks, const Fr* b_pow_chunks, Fr* carries) {
    __shared__ Fr s_add[DE_KATE_SCAN];
    const int tid = threadIdx.x;
    vals += blockIdx.x * nchunks;
    carries += blockIdx.x * nchunks;
    const Fr B = load(&b_pow_chunks[blockIdx.x]);
    const unsigned long long G = (nchunks + DE_KATE_SCAN - 1) / DE_KATE_SCAN;
    // element e of the descending order is chunk c = nchunks - 1 - e; its addend is V_(c+1) (zero for the top chunk)
    const unsigned long long e0 = (unsigned long long)tid * G;
    Fr A = Fr::zero();
    for (unsigned long long g = 0; g < G && e0 + g < nchunks; g++) {
        const unsigned long long c = nchunks - 1 - (e0 + g);
        const Fr a = c + 1 < nchunks ? load(&vals[c + 1]) : Fr::zero();
        A = add(a, mul(B, A));
    }
    store(&s_add[tid], A);
    Fr Md = Fr::one(), base = B;  // (B^G)^d, d = 1 to start with
    for (unsigned long long g = G; g; g >>= 1) {
        if (g & 1) Md = mul(Md, base);
        base = sqr(base);
    }
    __syncthreads();
    // inclusive scan: afterwards s_add[t] is the carry LEAVING thread t's chunks (the carry entering thread 0 is zero).  A thread
    // with fewer than G chunks sits at the end of the order: its own entry is wrong then, and nobody reads it.
    for (int d = 1; d < DE_KATE_SCAN; d <<= 1) {
        Fr pa = Fr::zero();
        const bool has = tid >= d;
        if (has) pa = load(&s_add[tid - d]);
        __syncthreads();
        if (has) store(&s_add[tid], add(load(&s_add[tid]), mul(Md, pa)));
        Md = sqr(Md);
        __syncthreads();
    }
    Fr cur = tid ? load(&s_add[tid - 1]) : Fr::zero();
    for (unsigned long long g = 0; g < G && e0 + g < nchunks; g++) {
        const unsigned long long c = nchunks - 1 - (e0 + g);
        const Fr a = c + 1 < nchunks ? load(&vals[c + 1]) : Fr::zero();
        cur = add(a, mul(B, cur));
        store(&carries[c], cur);
    }
}
// pass 3: per chunk, run the recurrence downwards from the incoming carry and write q (n - 1 coefficients; q[n-1] := 0 is
// also written so that the quotient can be committed as an n-coefficient polynomial)
__global__ void k_kate_write(const Fr* const* as, unsigned long long n, const Fr* bs, const Fr* carries, Fr* q, unsigned long long q_stride,
                             unsigned long long nchunks) {
    unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const Fr* a = as[blockIdx.y];
    const Fr b = load(&bs[blockIdx.y]);
    Fr* qq = q + blockIdx.y * q_stride;
    unsigned long long lo = c * DE_KATE_CHUNK, hi = lo + DE_KATE_CHUNK;
    if (hi > n) hi = n;
    Fr cur = load(&carries[blockIdx.y * nchunks + c]);  // = t_hi = q[hi - 1] (zero for the top chunk)
    for (unsigned long long i = hi; i > lo; i--) {
        store(&qq[i - 1], cur);
        cur = add(load(&a[i - 1]), mul(cur, b));
    }
}

int eval_polynomials_dev(de_ctx* ctx, const Fr* const* d_polys, size_t n, const unsigned int* d_point_index, const Fr* d_points, size_t count,
                         Fr* d_out, Fr* d_out_canonical) {
    if (count == 0) return DE_OK;
    static const char* seg_env = getenv("DE_POLY_SEGS");  // A/B switch for measurements
    const unsigned int segs = n >= 8192 ? (seg_env ? (unsigned int)atoi(seg_env) : DE_POLY_SEGS) : 1;
    if (segs < 1 || segs > 64) return fail(ctx, DE_ERR_ARG, "eval_polynomial: DE_POLY_SEGS out of range");
    Fr* partial = nullptr;
    unsigned int* done = nullptr;
    if (segs > 1) {
        DE_WS(ctx, scratch, Fr, WS_POLY_SCRATCH, (sizeof(Fr) * segs + sizeof(unsigned int)) * count);
        partial = scratch;
        done = (unsigned int*)(scratch + (size_t)segs * count);
        DE_CUDA(ctx, cudaMemsetAsync(done, 0, sizeof(unsigned int) * count, ctx->stream));
    }
    k_eval_polynomial<<<dim3((unsigned int)count, segs), DE_POLY_THREADS, 0, ctx->stream>>>(d_polys, n, d_point_index, d_points, d_out, d_out_canonical,
                                                                                           partial, done);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

size_t kate_scratch_elems(size_t n, size_t count) {
    const size_t nchunks = (n + DE_KATE_CHUNK - 1) / DE_KATE_CHUNK;
    return 2 * nchunks * count + 2 * count;
}

// d_as[i]: n coefficients; host_bs[i]: the divisor points (Montgomery); d_q: count quotients, q_stride apart, n entries each
// (the top one zero); d_scratch: kate_scratch_elems(n, count) elements.
int kate_division_dev(de_ctx* ctx, const Fr* const* d_as, size_t n, const de_fr* host_bs, size_t count, Fr* d_q, size_t q_stride,
                      Fr* d_scratch) {
    if (count == 0) return DE_OK;
    if (count > 64) return fail(ctx, DE_ERR_UNSUPPORTED, "kate_division: batch larger than 64");
    const size_t nchunks = (n + DE_KATE_CHUNK - 1) / DE_KATE_CHUNK;
    Fr* vals = d_scratch;
    Fr* carries = vals + nchunks * count;
    Fr* d_b = carries + nchunks * count;
    Fr* d_bpow = d_b + count;
    Fr hb[128];
    for (size_t i = 0; i < count; i++) {
        host::HFr b;
        memcpy(b.l, host_bs[i].l, 32);
        host::HFr bp = host::fr_pow(b, DE_KATE_CHUNK);
        de_fr t;
        memcpy(t.l, bp.l, 32);
        hb[i] = fr_from_host(host_bs[i]);
        hb[count + i] = fr_from_host(t);
    }
    // small pageable copy: staged by the runtime before the call returns, so the stack buffer may go out of scope
    DE_CUDA(ctx, cudaMemcpyAsync(d_b, hb, sizeof(Fr) * 2 * count, cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid((unsigned int)((nchunks + 127) / 128), (unsigned int)count);
    k_kate_chunk_values<<<grid, 128, 0, ctx->stream>>>(d_as, n, d_b, vals, nchunks);
    DE_CHECK_LAUNCH(ctx);
    k_kate_carries<<<(unsigned int)count, DE_KATE_SCAN, 0, ctx->stream>>>(vals, nchunks, d_bpow, carries);
    DE_CHECK_LAUNCH(ctx);
    k_kate_write<<<grid, 128, 0, ctx->stream>>>(d_as, n, d_b, carries, d_q, q_stride, nchunks);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

// ---- roofline denominator measured in the run (bench.py `roofline.peak`): dependent chains of Fr Montgomery products, the
// instruction mix every hot kernel of this library is made of (136 IMAD.WIDE.U32(.X) per product).  ILP independent chains
// per thread; the result is stored so that nothing is eliminated.
template <int ILP, bool SQUARE = false>
__global__ void __launch_bounds__(128) k_int_peak(Fr* out, const Fr* in, int iters) {
    Fr x[ILP];
    const Fr y = load(&in[threadIdx.x & 31]);
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = load(&in[(threadIdx.x + i) & 63]);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = SQUARE ? sqr(x[i]) : mul(x[i], y);
    }
    Fr s = x[0];
#pragma unroll
    for (int i = 1; i < ILP; i++) s = add(s, x[i]);
    store(&out[(size_t)blockIdx.x * blockDim.x + threadIdx.x], s);
}

}  // namespace de

using namespace de;

extern "C" {

static int int_peak(de_ctx* ctx, bool square, double* gmul_per_s);
int de_int_peak(de_ctx* ctx, double* gmul_per_s) { return int_peak(ctx, false, gmul_per_s); }
int de_int_peak_sqr(de_ctx* ctx, double* gsqr_per_s) { return int_peak(ctx, true, gsqr_per_s); }
static int int_peak(de_ctx* ctx, bool square, double* gmul_per_s) {
    if (!ctx) return DE_ERR_ARG;
    if (!gmul_per_s) return fail(ctx, DE_ERR_ARG, "de_int_peak: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    const int tpb = 128, warps_per_sm = 16, iters = 512, ILP = 2;
    const int blocks = ctx->sm_count * warps_per_sm * 32 / tpb;
    const size_t nthreads = (size_t)blocks * tpb;
    DE_WS(ctx, buf, Fr, WS_IO_A, sizeof(Fr) * (nthreads + 64));
    Fr* in = buf + nthreads;
    Fr h[64];
    for (int i = 0; i < 64; i++)
        for (int k = 0; k < 8; k++) h[i].l[k] = (k == 7) ? (0x0fffffffu - i) : (0x9e3779b9u * (i * 8 + k + 1));
    DE_CUDA(ctx, cudaMemcpyAsync(in, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t e0, e1;
    DE_CUDA(ctx, cudaEventCreate(&e0));
    DE_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {  // the first two launches warm the clocks up
        cudaEventRecord(e0, ctx->stream);
        if (square) k_int_peak<ILP, true><<<blocks, tpb, 0, ctx->stream>>>(buf, in, iters);
        else k_int_peak<ILP, false><<<blocks, tpb, 0, ctx->stream>>>(buf, in, iters);
        ctx->launches++;
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return fail(ctx, DE_ERR_CUDA, std::string("de_int_peak: ") + cudaGetErrorString(e));
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gmul_per_s = (double)nthreads * iters * ILP / (best * 1e-3) / 1e9;
    return DE_OK;
}

int de_eval_polynomial(de_ctx* ctx, const de_fr* poly, size_t n, const de_fr* point, de_fr* out) {
    if (!ctx) return DE_ERR_ARG;
    if (!point || !out || (n && !poly)) return fail(ctx, DE_ERR_ARG, "de_eval_polynomial: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_WS(ctx, d, Fr, WS_IO_A, sizeof(Fr) * (n + 4) + 64);
    Fr* d_point = d + n;
    Fr* d_out = d + n + 1;
    const Fr** d_ptr = (const Fr**)(d + n + 2);
    const Fr* hp = d;
    if (n) DE_CUDA(ctx, cudaMemcpyAsync(d, poly, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(d_point, point, sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync((void*)d_ptr, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(eval_polynomials_dev(ctx, d_ptr, n, nullptr, d_point, 1, d_out, nullptr));
    DE_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

int de_kate_division(de_ctx* ctx, const de_fr* a, size_t n, const de_fr* b, de_fr* q) {
    if (!ctx) return DE_ERR_ARG;
    if (!a || !b || !q) return fail(ctx, DE_ERR_ARG, "de_kate_division: null pointer");
    if (n < 2) return fail(ctx, DE_ERR_ARG, "de_kate_division: polynomial needs at least 2 coefficients");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_WS(ctx, d, Fr, WS_IO_A, sizeof(Fr) * (2 * n + kate_scratch_elems(n, 1) + 2));
    Fr* d_q = d + n;
    Fr* scratch = d_q + n;
    const Fr** d_ptr = (const Fr**)(scratch + kate_scratch_elems(n, 1));
    const Fr* hp = d;
    DE_CUDA(ctx, cudaMemcpyAsync(d, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync((void*)d_ptr, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(kate_division_dev(ctx, d_ptr, n, b, 1, d_q, n, scratch));
    DE_CUDA(ctx, cudaMemcpyAsync(q, d_q, sizeof(Fr) * (n - 1), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

}  // extern "C"
