// batch_affine_experiment.cu — MEASURES the alternative to XYZZ bucket accumulation that DESIGN.md section 4.6 had rejected on
// an estimate (VERDICT r01 "weak" #11): affine additions with a shared inversion (Montgomery's trick), 5M + 1S per addition plus
// the batch overhead, against the 8M + 2S of the mixed XYZZ addition the MSM uses today.
//
//   xyzz_chain     the inner loop of k_msm_accumulate: every thread adds M affine points into an XYZZ accumulator
//   affine_tree    one round of a pairwise reduction tree, the best case for batched affine: N independent pairs (P_i, Q_i),
//                  every thread owns G of them, multiplies up the denominators x2 - x1, the lanes of a warp (or the threads of a
//                  CTA) combine their products with prefix / suffix scans, ONE thread inverts, everybody back-substitutes and
//                  writes P_i + Q_i.  No bucket bookkeeping, no exceptional cases: an UPPER bound on what an MSM could gain.
//
// Both read their points with the same coalesced pattern and are timed with CUDA events at the occupancy each one reaches.
// Output: additions per second and field multiplications per addition (counted).  Build / run (see profiles/r02_batch_affine.md):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I delay-encryption-in-halo2_b200/csrc tools/batch_affine_experiment.cu -o tools/batch_affine_experiment
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#include "ec.cuh"
using namespace de;

// distinct curve points: P_i = [i + 1] G computed by a chain per thread block is overkill here; any x with a valid y is not
// needed either - the arithmetic cost does not depend on the points being on the curve, only on x1 != x2.  Pseudo-random
// field elements below p serve as coordinates.
__global__ void k_fill(Affine* pts, size_t n, unsigned int seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine a;
    unsigned int s = seed + (unsigned int)i * 2654435761u;
    for (int k = 0; k < 8; k++) {
        s = s * 1664525u + 1013904223u;
        a.x.l[k] = s;
        s = s * 1664525u + 1013904223u;
        a.y.l[k] = s;
    }
    a.x.l[7] &= 0x1fffffffu;
    a.y.l[7] &= 0x1fffffffu;
    store_affine(&pts[i], a);
}

template <int M>
__global__ void __launch_bounds__(128) k_xyzz_chain(const Affine* pts, size_t stride, XYZZ* out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ acc = xyzz_identity();
    Affine next = load_affine(&pts[t]);
    for (int k = 0; k < M; k++) {
        Affine cur = next;
        if (k + 1 < M) next = load_affine(&pts[(size_t)(k + 1) * stride + t]);
        xyzz_madd(acc, cur);
    }
    store_xyzz(&out[t], acc);
}

__device__ __forceinline__ Fq shfl_fq(const Fq& v, int src) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}
__device__ __forceinline__ Fq shfl_up_fq(const Fq& v, int d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
__device__ __forceinline__ Fq shfl_down_fq(const Fq& v, int d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}

// SCOPE 0: one inversion per warp; SCOPE 1: one inversion per CTA (THREADS threads)
template <int G, int SCOPE, int THREADS>
__global__ void __launch_bounds__(THREADS) k_affine_tree(const Affine* p, const Affine* q, size_t stride, Affine* out) {
    __shared__ Fq s_pre[THREADS / 32], s_suf[THREADS / 32];
    __shared__ Fq s_inv;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Fq pre[G];
    Fq acc = Fq::one();
#pragma unroll
    for (int k = 0; k < G; k++) {
        const Fq x1 = load(&p[(size_t)k * stride + t].x), x2 = load(&q[(size_t)k * stride + t].x);
        pre[k] = acc;
        acc = mul(acc, sub(x2, x1));
    }
    // inclusive prefix (incl) and suffix (sfx) products of `acc` over the lanes of the warp
    Fq incl = acc, sfx = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Fq a = shfl_up_fq(incl, d), b = shfl_down_fq(sfx, d);
        if (lane >= (unsigned)d) incl = mul(incl, a);
        if (lane + d < 32) sfx = mul(sfx, b);
    }
    Fq others;  // product of every OTHER thread's acc inside the scope, times the inverse of the scope's total
    if (SCOPE == 0) {
        Fq total = shfl_fq(incl, 31);
        Fq tinv = Fq::zero();
        if (lane == 31) tinv = inv(total);
        tinv = shfl_fq(tinv, 31);
        Fq excl_pre = shfl_up_fq(incl, 1), excl_suf = shfl_down_fq(sfx, 1);
        others = tinv;
        if (lane > 0) others = mul(others, excl_pre);
        if (lane < 31) others = mul(others, excl_suf);
    } else {
        constexpr int W = THREADS / 32;
        if (lane == 31) store(&s_pre[wid], incl);
        __syncthreads();
        if (threadIdx.x == 0) {
            // W warp totals: prefix / suffix products and the single inversion, serially (W <= 16)
            Fq tot[W], pf[W], sf[W];
            for (int w = 0; w < W; w++) tot[w] = load(&s_pre[w]);
            Fq a = Fq::one();
            for (int w = 0; w < W; w++) { pf[w] = a; a = mul(a, tot[w]); }
            Fq b = Fq::one();
            for (int w = W - 1; w >= 0; w--) { sf[w] = b; b = mul(b, tot[w]); }
            store(&s_inv, inv(a));
            for (int w = 0; w < W; w++) { store(&s_pre[w], pf[w]); store(&s_suf[w], sf[w]); }
        }
        __syncthreads();
        Fq excl_pre = shfl_up_fq(incl, 1), excl_suf = shfl_down_fq(sfx, 1);
        others = mul(load(&s_inv), mul(load(&s_pre[wid]), load(&s_suf[wid])));
        if (lane > 0) others = mul(others, excl_pre);
        if (lane < 31) others = mul(others, excl_suf);
    }
    // others = 1 / acc (this thread's own product); walk back through the thread's G pairs
    Fq ai = others;
#pragma unroll
    for (int k = G - 1; k >= 0; k--) {
        const Affine P = load_affine(&p[(size_t)k * stride + t]), Q = load_affine(&q[(size_t)k * stride + t]);
        const Fq dinv = mul(ai, pre[k]);
        ai = mul(ai, sub(Q.x, P.x));
        const Fq lam = mul(sub(Q.y, P.y), dinv);
        Affine r;
        r.x = sub(sub(sqr(lam), P.x), Q.x);
        r.y = sub(mul(lam, sub(P.x, r.x)), P.y);
        store_affine(&out[(size_t)k * stride + t], r);
    }
}

template <class F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

template <int G, int SCOPE, int THREADS>
static void run_tree(const Affine* p, const Affine* q, Affine* out, size_t threads, const char* scope) {
    const int blocks = (int)(threads / THREADS);
    float ms = time_ms([&] { k_affine_tree<G, SCOPE, THREADS><<<blocks, THREADS>>>(p, q, threads, out); });
    cudaError_t e = cudaGetLastError();
    const double adds = (double)threads * G;
    const int batch = SCOPE == 0 ? 32 * G : THREADS * G;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_affine_tree<G, SCOPE, THREADS>);
    printf("{\"kernel\": \"affine_tree\", \"pairs_per_thread\": %d, \"inversion_scope\": \"%s\", \"additions_per_inversion\": %d, \"registers\": %d, "
           "\"local_bytes\": %zu, \"gadds_s\": %.3f, \"muls_per_add_counted\": %.2f, \"err\": \"%s\"}\n",
           G, scope, batch, fa.numRegs, fa.localSizeBytes, adds / ms / 1e6, 6.0 + (SCOPE == 0 ? 12.0 : 14.0) / G, cudaGetErrorString(e));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t threads = (size_t)sms * 2048;  // every variant gets the same number of threads: 16 resident warps/SM x 4 waves
    const int GMAX = 16;
    Affine *p, *q, *out;
    XYZZ* xo;
    cudaMalloc((void**)&p, sizeof(Affine) * threads * GMAX);
    cudaMalloc((void**)&q, sizeof(Affine) * threads * GMAX);
    cudaMalloc((void**)&out, sizeof(Affine) * threads * GMAX);
    cudaMalloc((void**)&xo, sizeof(XYZZ) * threads);
    k_fill<<<(unsigned)((threads * GMAX + 255) / 256), 256>>>(p, threads * GMAX, 1u);
    k_fill<<<(unsigned)((threads * GMAX + 255) / 256), 256>>>(q, threads * GMAX, 77u);
    cudaDeviceSynchronize();
    printf("{\"device\": \"%s\", \"sms\": %d, \"threads\": %zu}\n", prop.name, sms, threads);
    {
        const int blocks = (int)(threads / 128);
        float ms = time_ms([&] { k_xyzz_chain<16><<<blocks, 128>>>(p, threads, xo); });
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_xyzz_chain<16>);
        printf("{\"kernel\": \"xyzz_chain\", \"points_per_thread\": 16, \"registers\": %d, \"gadds_s\": %.3f, \"muls_per_add_counted\": 10.0}\n", fa.numRegs,
               (double)threads * 16 / ms / 1e6);
        ms = time_ms([&] { k_xyzz_chain<32><<<blocks / 2, 128>>>(p, threads / 2, xo); });
        printf("{\"kernel\": \"xyzz_chain\", \"points_per_thread\": 32, \"registers\": %d, \"gadds_s\": %.3f, \"muls_per_add_counted\": 10.0}\n", fa.numRegs,
               (double)(blocks / 2) * 128 * 32 / ms / 1e6);
    }
    run_tree<4, 0, 128>(p, q, out, threads, "warp");
    run_tree<8, 0, 128>(p, q, out, threads, "warp");
    run_tree<16, 0, 128>(p, q, out, threads, "warp");
    run_tree<4, 1, 256>(p, q, out, threads, "cta256");
    run_tree<8, 1, 256>(p, q, out, threads, "cta256");
    run_tree<16, 1, 256>(p, q, out, threads, "cta256");
    run_tree<8, 1, 512>(p, q, out, threads, "cta512");
    run_tree<16, 1, 512>(p, q, out, threads, "cta512");
    return 0;
}
