// ntt.cu — host side of the NTT / EvaluationDomain entry points: transform plans (pass geometry + twiddle tables
// resident in HBM), kernel dispatch, and the C ABI functions de_ntt*, de_domain_*, de_coeff_to_extended*, ...
// Reference semantics: halo2_proofs::arithmetic::best_fft and poly::EvaluationDomain.
#include <atomic>
#include <functional>
#include <stdlib.h>
#include <string.h>

#include <thread>

#include "common.cuh"
#include "ntt.cuh"

namespace de {

thread_local std::string g_create_error;

struct NttPlan {
    uint32_t log_n = 0;
    de_fr omega;
    int npass = 0;
    int S[3] = {0, 0, 0};
    Fr* wl[3] = {nullptr, nullptr, nullptr};
    uint4* wl_planes[3] = {nullptr, nullptr, nullptr};  // the same tables in the kernel's shared-memory layout (TMA source)
    Fr* tw_full = nullptr;
    Fr* tw_hi = nullptr;
    Fr* tw_lo = nullptr;
    uint32_t tw_lo_bits = 0;
    Fr* tw_pass[2] = {nullptr, nullptr};  // log_n > 22: pass-ordered twiddles of the strided-column passes (N entries each)
    ~NttPlan() {
        for (int i = 0; i < 2; i++)
            if (tw_pass[i]) cudaFree(tw_pass[i]);
        for (int i = 0; i < 3; i++)
            if (wl[i]) cudaFree(wl[i]);
        for (int i = 0; i < 3; i++)
            if (wl_planes[i]) cudaFree(wl_planes[i]);
        if (tw_full) cudaFree(tw_full);
        if (tw_hi) cudaFree(tw_hi);
        if (tw_lo) cudaFree(tw_lo);
    }
};

// Tables of the multi-GPU transform (de_ntt_dist_*): w_N^e two-level, from which the resident inter-stage twiddles of a rank are
// built, the root of the local N/W-point transform and the W/2 powers of the cross-rank root.
struct NttDistPlan {
    uint32_t log_n = 0, log_w = 0;
    de_fr omega, omega_local;
    Fr* tw_hi = nullptr;
    Fr* tw_lo = nullptr;
    uint32_t tw_lo_bits = 0;
    Fr* tw_rank = nullptr;   // inter-stage twiddles (k_dist_twiddles) of the rank this context last ran stage 2 as
    uint32_t tw_rank_of = 0;
    Fr wcross[4];
    ~NttDistPlan() {
        if (tw_hi) cudaFree(tw_hi);
        if (tw_lo) cudaFree(tw_lo);
        if (tw_rank) cudaFree(tw_rank);
    }
};

void ntt_free_plans(de_ctx* ctx) {
    for (auto* p : ctx->plans) delete p;
    ctx->plans.clear();
    for (auto* p : ctx->dist_plans) delete p;
    ctx->dist_plans.clear();
    if (ctx->dist_stream) {
        cudaStreamDestroy(ctx->dist_stream);
        ctx->dist_stream = nullptr;
    }
    if (ctx->dist_error) {
        cudaFree(ctx->dist_error);
        ctx->dist_error = nullptr;
    }
    for (auto& e : ctx->dist_ev)
        if (e) {
            cudaEventDestroy(e);
            e = nullptr;
        }
}

static int pow_table(de_ctx* ctx, Fr** out, unsigned long long n, const Fr& base, unsigned long long step) {
    DE_CUDA(ctx, cudaMalloc((void**)out, sizeof(Fr) * (n ? n : 1)));
    unsigned int threads = 256;
    unsigned long long blocks = (n + threads - 1) / threads;
    k_pow_table<<<(unsigned int)blocks, threads, 0, ctx->stream>>>(*out, n, base, step);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

static int get_plan(de_ctx* ctx, const de_fr& omega, uint32_t log_n, NttPlan** out) {
    for (auto* p : ctx->plans)
        if (p->log_n == log_n && memcmp(&p->omega, &omega, sizeof(de_fr)) == 0) {
            *out = p;
            return DE_OK;
        }
    if (log_n < 1 || log_n > 28) return fail(ctx, DE_ERR_ARG, "ntt: log_n must be in 1..28");
    NttPlan* p = new NttPlan();
    p->log_n = log_n;
    p->omega = omega;
    if (log_n <= 10) {
        p->npass = 1;
        p->S[0] = (int)log_n;
    } else if (log_n <= 20) {
        p->npass = 2;
        p->S[0] = (int)(log_n + 1) / 2;
        p->S[1] = (int)log_n - p->S[0];
    } else {
        p->npass = 3;
        p->S[0] = (int)(log_n + 2) / 3;
        p->S[1] = (int)(log_n - p->S[0] + 1) / 2;
        p->S[2] = (int)log_n - p->S[0] - p->S[1];
    }
    Fr w = fr_from_host(omega);
    unsigned long long N = 1ull << log_n;
    int rc = DE_OK;
    for (int k = 0; k < p->npass && rc == DE_OK; k++) {
        unsigned long long L = 1ull << p->S[k];
        rc = pow_table(ctx, &p->wl[k], L / 2, w, N / L);
        if (rc == DE_OK && L >= 2) {
            const unsigned int half = (unsigned int)(L / 2);
            if (cudaMalloc((void**)&p->wl_planes[k], sizeof(uint4) * 2 * half) != cudaSuccess) {
                cudaGetLastError();
                rc = fail(ctx, DE_ERR_OOM, "ntt: twiddle table allocation failed");
            } else {
                k_split_planes<<<(half + 255) / 256, 256, 0, ctx->stream>>>(p->wl[k], p->wl_planes[k], half);
                ctx->launches++;
            }
        }
    }
    if (rc == DE_OK && p->npass > 1) {
        if (log_n <= 22) {
            rc = pow_table(ctx, &p->tw_full, N, w, 1);
        } else {
            p->tw_lo_bits = 12;
            rc = pow_table(ctx, &p->tw_lo, 1ull << p->tw_lo_bits, w, 1);
            if (rc == DE_OK) rc = pow_table(ctx, &p->tw_hi, N >> p->tw_lo_bits, w, 1ull << p->tw_lo_bits);
            // 180 GB of HBM: keep every inter-pass twiddle resident in the order its pass reads it (2 x 32 N bytes per plan) instead of
            // rebuilding it from the two-level tables with a second multiplication per element.  Budget per plan: DE_NTT_TABLE_MB
            // (default 16 GiB, i.e. up to N = 2^28); beyond it, or if the allocation fails, the two-level path stays.
            size_t budget = (size_t)16 << 30;
            if (const char* e = getenv("DE_NTT_TABLE_MB")) budget = (size_t)atoll(e) << 20;
            if (rc == DE_OK && sizeof(Fr) * N * (p->npass - 1) <= budget) {
                unsigned long long lprod = 1;
                for (int k = 0; k + 1 < p->npass; k++) {
                    const unsigned long long L = 1ull << p->S[k];
                    const unsigned long long M = N / (lprod * L);
                    if (cudaMalloc((void**)&p->tw_pass[k], sizeof(Fr) * N) != cudaSuccess) {
                        cudaGetLastError();
                        for (int i = 0; i < 2; i++) {
                            if (p->tw_pass[i]) cudaFree(p->tw_pass[i]);
                            p->tw_pass[i] = nullptr;
                        }
                        break;
                    }
                    k_pass_twiddles<<<(unsigned int)((N + 255) / 256), 256, 0, ctx->stream>>>(p->tw_pass[k], N, M, L * M, lprod, p->tw_hi, p->tw_lo,
                                                                                          p->tw_lo_bits);
                    ctx->launches++;
                    lprod *= L;
                }
            }
        }
    }
    if (rc != DE_OK) {
        delete p;
        return rc;
    }
    ctx->plans.push_back(p);
    *out = p;
    return DE_OK;
}

template <int S, int LT, bool DIST>
static int launch_pass_t(de_ctx* ctx, const NttPassParams& prm, const NttDistArgs<DIST>& dx, unsigned int blocks, unsigned int batch) {
    using Sh = NttShape<S, LT>;
    // the attribute is per (kernel, device); several host threads (one per in-flight proof) may get here at once: setting it
    // twice is harmless, so a relaxed atomic flag per device is enough.  Devices beyond the table set it on every launch.
    static std::atomic<bool> configured[64];
    const bool tracked = ctx->device >= 0 && ctx->device < 64;
    if (!tracked || !configured[ctx->device].load(std::memory_order_acquire)) {
        DE_CUDA(ctx, cudaFuncSetAttribute(k_ntt_pass<S, LT, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sh::SMEM));
        if (tracked) configured[ctx->device].store(true, std::memory_order_release);
    }
    dim3 grid(blocks, batch);
    DE_TIMED(ctx, DIST ? "k_ntt_pass_dist" : "k_ntt_pass", (double)blocks * batch * Sh::M,
             (k_ntt_pass<S, LT, DIST><<<grid, Sh::NTHREADS, Sh::SMEM, ctx->stream>>>(prm, dx)));
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

// tile width used for sub-transform size 2^S when tiling is on
static int tile_log(int S) {
    int lt = 11 - S;
    if (lt > 3) lt = 3;
    if (lt < 0) lt = 0;
    return lt;
}

static int launch_pass(de_ctx* ctx, int S, int LT, const NttPassParams& prm, unsigned int blocks, unsigned int batch) {
    const NttDistArgs<false> none{};
#define CASE(s, lt) \
    if (S == s && LT == lt) return launch_pass_t<s, lt, false>(ctx, prm, none, blocks, batch);
    CASE(1, 0) CASE(2, 0) CASE(3, 0) CASE(4, 0) CASE(5, 0) CASE(6, 0) CASE(7, 0) CASE(8, 0) CASE(9, 0) CASE(10, 0)
    CASE(1, 3) CASE(2, 3) CASE(3, 3) CASE(4, 3) CASE(5, 3) CASE(6, 3) CASE(7, 3) CASE(8, 3) CASE(9, 2) CASE(10, 1)
#undef CASE
    return fail(ctx, DE_ERR_UNSUPPORTED, "ntt: no kernel for this pass shape");
}

// last pass of the local transform of the multi-GPU NTT (peer-store epilogue); local sizes 2^11 .. 2^25 end in one of these
static int launch_pass_dist(de_ctx* ctx, int S, int LT, const NttPassParams& prm, const NttDistArgs<true>& dx, unsigned int blocks) {
#define CASE(s, lt) \
    if (S == s && LT == lt) return launch_pass_t<s, lt, true>(ctx, prm, dx, blocks, 1);
    CASE(5, 3) CASE(6, 3) CASE(7, 3) CASE(8, 3) CASE(9, 2) CASE(10, 1)
#undef CASE
    return fail(ctx, DE_ERR_UNSUPPORTED, "ntt (multi-GPU): no peer-store kernel for this pass shape");
}

// Pipelined exchange of the multi-GPU transform: the peer-store pass is launched as `chunks` contiguous ranges of CTAs and
// after_chunk(k) runs on the host after range k has been enqueued (it records an event or enqueues a flag signal).  On return
// `period` = the number of columns one sub-row of the pass spans (M / L) and `chunks` the count actually used.
struct NttDistChunks {
    int chunks = 1;
    std::function<int(int)> after_chunk;
    unsigned long long period = 0;
};

int ntt_run(de_ctx* ctx, const de_fr& omega, uint32_t log_n, const Fr* d_src, size_t src_stride, Fr* d_dst, size_t dst_stride,
            size_t batch, int in_mode, size_t n_in, const Fr* zeta2, int out_mode, const Fr* oscale3, const NttDistArgs<true>* dist,
            NttDistChunks* chunks) {
    if (batch == 0) return DE_OK;
    if (dist && (batch != 1 || log_n < 11)) return fail(ctx, DE_ERR_ARG, "ntt (multi-GPU): local transform must be one vector of >= 2^11");
    if (batch > 65535) return fail(ctx, DE_ERR_ARG, "ntt: batch too large");
    if (log_n == 0) {
        if (in_mode != 0 || out_mode != 0) return fail(ctx, DE_ERR_ARG, "ntt: fused scaling needs log_n >= 1");
        if (d_src != d_dst)
            DE_CUDA(ctx, cudaMemcpy2DAsync(d_dst, dst_stride * sizeof(Fr), d_src, src_stride * sizeof(Fr), sizeof(Fr), batch,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
        return DE_OK;
    }
    NttPlan* plan = nullptr;
    DE_TRY(get_plan(ctx, omega, log_n, &plan));
    const unsigned long long N = 1ull << log_n;
    const int P = plan->npass;
    Fr* scratch = nullptr;
    if (P > 1) {
        scratch = (Fr*)ctx->ws[WS_NTT_SCRATCH].ensure(sizeof(Fr) * N * batch);
        if (!scratch) return fail(ctx, DE_ERR_OOM, "ntt: scratch allocation failed");
    }
    unsigned long long Lprod = 1;  // L_1 * ... * L_{k-1}
    for (int k = 0; k < P; k++) {
        const int S = plan->S[k];
        const unsigned long long L = 1ull << S;
        const bool last = (k == P - 1);
        NttPassParams prm;
        memset(&prm, 0, sizeof(prm));
        prm.wl_planes = plan->wl_planes[k];
        // source / destination of this pass
        if (k == 0) {
            prm.in = d_src;
            prm.in_batch_stride = src_stride;
            prm.in_mode = in_mode;
            prm.n_in = n_in;
            if (in_mode == 1) {
                prm.zeta[0] = zeta2[0];
                prm.zeta[1] = zeta2[1];
            }
        } else {
            prm.in = scratch;
            prm.in_batch_stride = N;
        }
        if (last) {
            prm.out = d_dst;
            prm.out_batch_stride = dst_stride;
            prm.out_mode = out_mode;
            if (out_mode == 1)
                for (int i = 0; i < 3; i++) prm.oscale[i] = oscale3[i];
        } else {
            prm.out = scratch;
            prm.out_batch_stride = N;
        }
        int LT;
        unsigned long long blocks;
        if (!last) {
            // strided-column pass over segments of length L * M
            const unsigned long long M = N / (Lprod * L);
            LT = tile_log(S);
            const unsigned long long T = 1ull << LT;
            prm.G = (unsigned int)(M / T);
            prm.in_grp = prm.out_grp = L * M;
            prm.in_blk = prm.out_blk = T;
            prm.in_tt = prm.out_tt = 1;
            prm.in_el = prm.out_el = M;
            prm.tw_mul = Lprod;  // w_{L*M} = w_N^(N / (L*M)) = w_N^(L_1..L_{k-1})
            if (plan->tw_full) {
                prm.tw_mode = 1;
                prm.tw_full = plan->tw_full;
            } else if (plan->tw_pass[k]) {
                prm.tw_mode = 3;
                prm.tw_pass = plan->tw_pass[k];
            } else {
                prm.tw_mode = 2;
                prm.tw_hi = plan->tw_hi;
                prm.tw_lo = plan->tw_lo;
                prm.tw_lo_bits = plan->tw_lo_bits;
            }
            blocks = N / (T * L);
        } else if (P == 1) {
            LT = 0;
            prm.G = 1;
            prm.in_el = prm.out_el = 1;
            blocks = 1;
        } else {
            // contiguous rows of length L; tile over j1 so the transposed store is T*32 B contiguous
            const unsigned long long L1 = 1ull << plan->S[0];
            const unsigned long long M1 = N / L1;
            LT = tile_log(S);
            const unsigned long long T = 1ull << LT;
            prm.G = (unsigned int)(L1 / T);
            prm.in_grp = L;
            prm.in_blk = T * M1;
            prm.in_tt = M1;
            prm.in_el = 1;
            prm.out_grp = L1;
            prm.out_blk = T;
            prm.out_tt = 1;
            prm.out_el = N / L;
            blocks = N / (T * L);
        }
        if (last && dist) {
            // a contiguous range of CTAs covers the same contiguous range of column offsets c' = blockIdx.x * T + tt inside each
            // of the L sub-rows (length N / L) of the pass' output
            NttDistArgs<true> dx = *dist;
            int K = chunks ? chunks->chunks : 1;
            while (K > 1 && (blocks % K != 0 || blocks / K < 2ull * ctx->sm_count)) K >>= 1;
            if (chunks) {
                chunks->chunks = K;
                chunks->period = N / L;
            }
            const unsigned int per = (unsigned int)(blocks / K);
            for (int c = 0; c < K; c++) {
                dx.block_off = (unsigned int)c * per;
                DE_TRY(launch_pass_dist(ctx, S, LT, prm, dx, per));
                if (chunks && chunks->after_chunk) DE_TRY(chunks->after_chunk(c));
            }
        } else {
            DE_TRY(launch_pass(ctx, S, LT, prm, (unsigned int)blocks, (unsigned int)batch));
        }
        Lprod *= L;
    }
    return DE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// small device helpers (one-time constants; keeps every field operation of the library on the GPU)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_fr_pow(Fr base, unsigned long long e, Fr* out) {
    Fr acc = Fr::one(), cur = base;
    while (e) {
        if (e & 1) acc = mul(acc, cur);
        cur = sqr(cur);
        e >>= 1;
    }
    store(out, acc);
}

// base^e on the device, result read back (plan constants; synchronises the context's stream)
int fr_host_pow(de_ctx* ctx, const de_fr& base, uint64_t e, de_fr* out) {
    DE_WS(ctx, d, Fr, WS_IO_B, sizeof(Fr));
    k_fr_pow<<<1, 1, 0, ctx->stream>>>(fr_from_host(base), e, d);
    DE_CHECK_LAUNCH(ctx);
    Fr h;
    DE_CUDA(ctx, cudaMemcpyAsync(&h, d, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = fr_to_host(h);
    return DE_OK;
}

template <class P>
__device__ Fp<P> dev_inv(const Fp<P>& a) {
    Fp<P> acc = Fp<P>::one();
    for (int i = 7; i >= 0; i--) {
        uint32_t w = P::p(i);
        if (i == 0) w -= 2;
        for (int b = 31; b >= 0; b--) {
            acc = sqr(acc);
            if ((w >> b) & 1) acc = mul(acc, a);
        }
    }
    return acc;
}

// EvaluationDomain::new constants.  out: [0] ext_omega [1] ext_omega_inv [2] omega [3] omega_inv [4] zeta [5] zeta^2
// [6] 1/n [7] 1/ext_n [8..8+2^(ek-k)) t_evaluations.  One thread per constant.
__global__ void k_domain_consts(unsigned int k, unsigned int ek, Fr* out) {
    const unsigned int t = threadIdx.x;
    const unsigned int nt = 1u << (ek - k);
    if (t >= 8 + nt) return;
    Fr root = Fr::zero();  // ROOT_OF_UNITY = 7^((r-1)/2^28), canonical limbs -> Montgomery
    {
        const uint32_t c[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
        for (int i = 0; i < 8; i++) root.l[i] = c[i];
        root = to_mont(root);
    }
    Fr zeta = Fr::zero();
    {
        const uint32_t c[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
        for (int i = 0; i < 8; i++) zeta.l[i] = c[i];
        zeta = to_mont(zeta);
    }
    Fr ext_omega = root;
    for (unsigned int i = ek; i < 28; i++) ext_omega = sqr(ext_omega);
    Fr omega = ext_omega;
    for (unsigned int i = k; i < ek; i++) omega = sqr(omega);
    Fr r;
    if (t == 0) r = ext_omega;
    else if (t == 1) r = dev_inv(ext_omega);
    else if (t == 2) r = omega;
    else if (t == 3) r = dev_inv(omega);
    else if (t == 4) r = zeta;
    else if (t == 5) r = sqr(zeta);
    else if (t == 6 || t == 7) {
        // 1 / 2^m = (1/2)^m
        Fr two = add(Fr::one(), Fr::one());
        Fr half = dev_inv(two);
        unsigned int m = (t == 6) ? k : ek;
        r = Fr::one();
        for (unsigned int i = 0; i < m; i++) r = mul(r, half);
    } else {
        // t_evaluations[i] = 1 / ((zeta * ext_omega^i)^n - 1)
        unsigned int i = t - 8;
        Fr cur = zeta;
        for (unsigned int s = 0; s < i; s++) cur = mul(cur, ext_omega);
        for (unsigned int s = 0; s < k; s++) cur = sqr(cur);
        r = dev_inv(sub(cur, Fr::one()));
    }
    store(&out[t], r);
}

template <class F>
__global__ void k_vec_op(int op, const F* a, const F* b, F* out, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = load(&a[i]);
    F r;
    if (op == DE_OP_MUL) r = mul(x, load(&b[i]));
    else if (op == DE_OP_ADD) r = add(x, load(&b[i]));
    else if (op == DE_OP_SUB) r = sub(x, load(&b[i]));
    else if (op == DE_OP_FROM_MONT) r = from_mont(x);
    else if (op == DE_OP_INV) r = inv(x);
    else if (op == DE_OP_SQR) r = sqr(x);
    else r = to_mont(x);
    store(&out[i], r);
}

template <class F>
static int vec_op(de_ctx* ctx, int op, const void* a, const void* b, void* out, size_t n) {
    if (!ctx) return DE_ERR_ARG;
    if (op < DE_OP_MUL || op > DE_OP_SQR) return fail(ctx, DE_ERR_ARG, "vec_op: unknown op");
    if (n == 0) return DE_OK;
    bool binary = op <= DE_OP_SUB;
    if (!a || !out || (binary && !b)) return fail(ctx, DE_ERR_ARG, "vec_op: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_WS(ctx, da, F, WS_IO_A, sizeof(F) * n * 2);
    DE_WS(ctx, dout, F, WS_IO_B, sizeof(F) * n);
    F* db = da + n;
    DE_CUDA(ctx, cudaMemcpyAsync(da, a, sizeof(F) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (binary) DE_CUDA(ctx, cudaMemcpyAsync(db, b, sizeof(F) * n, cudaMemcpyHostToDevice, ctx->stream));
    unsigned int threads = 128;
    k_vec_op<F><<<(unsigned int)((n + threads - 1) / threads), threads, 0, ctx->stream>>>(op, da, db, dout, n);
    DE_CHECK_LAUNCH(ctx);
    DE_CUDA(ctx, cudaMemcpyAsync(out, dout, sizeof(F) * n, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

}  // namespace de

using namespace de;

struct de_domain {
    de_ctx* ctx;
    uint32_t j, k, ek;
    size_t n, ext_n, qdeg;
    de_fr omega, omega_inv, ext_omega, ext_omega_inv;
    Fr zeta[2];          // zeta, zeta^2
    Fr ifft_scale[3];    // 1/n three times
    Fr ext_out_scale[3]; // (1/ext_n) * {1, zeta^2, zeta}
    Fr* d_t_evals;       // 2^(ek-k) entries
    uint32_t t_len;
};

extern "C" {

const char* de_version(void) { return "de_b200 0.1 (sm_100a)"; }

int de_ctx_create(int device, de_ctx** out) {
    if (!out) return fail(nullptr, DE_ERR_ARG, "de_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, DE_ERR_CUDA, std::string("no CUDA device available (") +
                                              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                              "); this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(nullptr, DE_ERR_ARG, "de_ctx_create: device index out of range");
    de_ctx* ctx = new de_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        std::string msg = cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return fail(nullptr, DE_ERR_CUDA, "de_ctx_create: " + msg);
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return DE_OK;
}

int de_ctx_destroy(de_ctx* ctx) {
    if (!ctx) return DE_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ntt_free_plans(ctx);
    for (auto& g : ctx->msm_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& w : ctx->ws) w.release();
    ctx->pinned.release();
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return DE_OK;
}

int de_ctx_set_stream(de_ctx* ctx, void* s) {
    if (!ctx) return DE_ERR_ARG;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return DE_OK;
}

int de_ctx_set_mode(de_ctx* ctx, int mode) {
    if (!ctx) return DE_ERR_ARG;
    if (mode != DE_MODE_LATENCY && mode != DE_MODE_THROUGHPUT) return fail(ctx, DE_ERR_ARG, "de_ctx_set_mode: unknown mode");
    ctx->mode = mode;
    return DE_OK;
}

int de_ctx_sync(de_ctx* ctx) {
    if (!ctx) return DE_ERR_ARG;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

const char* de_last_error(de_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

// ---- per-kernel timing (CUDA events on the context's stream) ------------------------------------------------
static void timing_collect(de_ctx* ctx) {
    for (auto& t : ctx->timed) {
        float ms = 0;
        if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            KernelStat* st = nullptr;
            for (auto& s : ctx->stats)
                if (s.name == t.name) st = &s;
            if (!st) {
                ctx->stats.push_back(KernelStat());
                st = &ctx->stats.back();
                st->name = t.name;
            }
            st->ms += ms;
            st->units += t.units;
            st->launches++;
        }
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    cudaGetLastError();
    ctx->timed.clear();
}
int de_timing_enable(de_ctx* ctx, int on) {
    if (!ctx) return DE_ERR_ARG;
    timing_collect(ctx);
    ctx->timing = on != 0;
    return DE_OK;
}
int de_timing_reset(de_ctx* ctx) {
    if (!ctx) return DE_ERR_ARG;
    timing_collect(ctx);
    ctx->stats.clear();
    return DE_OK;
}
int de_timing_get(de_ctx* ctx, const char* kernel, double* total_ms, double* total_units, uint64_t* launches) {
    if (!ctx || !kernel) return DE_ERR_ARG;
    timing_collect(ctx);
    double ms = 0, units = 0;
    uint64_t n = 0;
    for (auto& s : ctx->stats)
        if (s.name == kernel) {
            ms = s.ms;
            units = s.units;
            n = s.launches;
        }
    if (total_ms) *total_ms = ms;
    if (total_units) *total_units = units;
    if (launches) *launches = n;
    return DE_OK;
}
uint64_t de_launch_count(de_ctx* ctx) { return ctx ? ctx->launches : 0; }

int de_fr_vec_op(de_ctx* ctx, int op, const de_fr* a, const de_fr* b, de_fr* out, size_t n) { return vec_op<Fr>(ctx, op, a, b, out, n); }
int de_fq_vec_op(de_ctx* ctx, int op, const de_fq* a, const de_fq* b, de_fq* out, size_t n) { return vec_op<Fq>(ctx, op, a, b, out, n); }

// ---- best_fft -------------------------------------------------------------------------------------------------
int de_ntt_dev(de_ctx* ctx, de_fr* d_a, const de_fr* omega, uint32_t log_n, size_t batch, size_t stride) {
    if (!ctx) return DE_ERR_ARG;
    if (!d_a || !omega) return fail(ctx, DE_ERR_ARG, "de_ntt_dev: null pointer");
    if (log_n > 28) return fail(ctx, DE_ERR_ARG, "de_ntt_dev: log_n > 28 exceeds the 2-adicity of Fr");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, *omega, log_n, (const Fr*)d_a, stride, (Fr*)d_a, stride, batch, 0, 0, nullptr, 0, nullptr);
}

int de_ntt(de_ctx* ctx, de_fr* a, const de_fr* omega, uint32_t log_n) {
    if (!ctx) return DE_ERR_ARG;
    if (!a || !omega) return fail(ctx, DE_ERR_ARG, "de_ntt: null pointer");
    if (log_n > 28) return fail(ctx, DE_ERR_ARG, "de_ntt: log_n > 28 exceeds the 2-adicity of Fr");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)1 << log_n;
    DE_WS(ctx, d, Fr, WS_IO_A, sizeof(Fr) * n);
    DE_CUDA(ctx, cudaMemcpyAsync(d, a, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(ntt_run(ctx, *omega, log_n, d, n, d, n, 1, 0, 0, nullptr, 0, nullptr));
    DE_CUDA(ctx, cudaMemcpyAsync(a, d, sizeof(Fr) * n, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

}  // extern "C"

// ---- multi-GPU best_fft (SURVEY.md section 8e, row "NTT: single huge vector") --------------------------------
// N = W * M (W ranks).  Rank r holds x_r[t] = a[r + W t] (cyclic).  With i = i1 + W i2 and j = j2 + M j1:
//   A[j2 + M j1] = sum_i1 w_W^(i1 j1) * [ w_N^(i1 j2) * sum_i2 a[i1 + W i2] w_M^(i2 j2) ]
// stage 1 (rank i1): local M-point transform, peer-store of column j2 into row i1 of the exchange buffer of rank j2 / C
// (C = M / W);  stage 2 (rank q): twiddle, W-point transform down every column of its exchange buffer, peer-store of output j1
// into rank j1's block at q C + c.  Result: rank j1 holds A[j1 M .. (j1 + 1) M) - natural order, contiguous blocks.
static int get_dist_plan(de_ctx* ctx, const de_fr& omega, uint32_t log_n, uint32_t world, NttDistPlan** out) {
    uint32_t lw = 0;
    while ((1u << lw) < world) lw++;
    if ((1u << lw) != world || world > 8) return fail(ctx, DE_ERR_ARG, "ntt (multi-GPU): world must be 1, 2, 4 or 8");
    if (log_n > 28 || log_n < 11 + lw) return fail(ctx, DE_ERR_ARG, "ntt (multi-GPU): need 11 + log2(world) <= log_n <= 28");
    for (auto* p : ctx->dist_plans)
        if (p->log_n == log_n && p->log_w == lw && memcmp(&p->omega, &omega, sizeof(de_fr)) == 0) {
            *out = p;
            return DE_OK;
        }
    NttDistPlan* p = new NttDistPlan();
    p->log_n = log_n;
    p->log_w = lw;
    p->omega = omega;
    const unsigned long long N = 1ull << log_n;
    Fr w = fr_from_host(omega);
    p->tw_lo_bits = log_n < 12 ? log_n : 12;
    int rc = pow_table(ctx, &p->tw_lo, 1ull << p->tw_lo_bits, w, 1);
    if (rc == DE_OK) rc = pow_table(ctx, &p->tw_hi, N >> p->tw_lo_bits, w, 1ull << p->tw_lo_bits);
    if (rc == DE_OK) rc = fr_host_pow(ctx, omega, world, &p->omega_local);
    for (uint32_t i = 0; rc == DE_OK && i < 4; i++) {
        // w_W^i = w_N^(i * N / W); entries beyond W/2 are unused
        de_fr t;
        rc = fr_host_pow(ctx, omega, (uint64_t)i * (N >> lw), &t);
        p->wcross[i] = fr_from_host(t);
    }
    if (rc != DE_OK) {
        delete p;
        return rc;
    }
    ctx->dist_plans.push_back(p);
    *out = p;
    return DE_OK;
}

// inter-stage twiddles of `rank` (k_dist_twiddles), resident per plan; built on `stream` when missing (one-time: allocates)
static int ensure_rank_twiddles(de_ctx* ctx, NttDistPlan* plan, uint32_t rank, cudaStream_t stream) {
    const uint32_t LW = plan->log_w;
    if (LW == 0 || (plan->tw_rank && plan->tw_rank_of == rank)) return DE_OK;
    const unsigned long long C = (1ull << (plan->log_n - LW)) >> LW;
    const unsigned long long cnt = C * ((1u << LW) - 1);
    if (!plan->tw_rank && cudaMalloc((void**)&plan->tw_rank, sizeof(Fr) * cnt) != cudaSuccess) {
        cudaGetLastError();
        plan->tw_rank = nullptr;
        return fail(ctx, DE_ERR_OOM, "ntt (multi-GPU): twiddle table allocation failed");
    }
    k_dist_twiddles<<<(unsigned int)((cnt + 255) / 256), 256, 0, stream>>>(plan->tw_rank, C, (1u << LW) - 1, (unsigned long long)rank * C, plan->tw_hi,
                                                                           plan->tw_lo, plan->tw_lo_bits);
    DE_CHECK_LAUNCH(ctx);
    plan->tw_rank_of = rank;
    return DE_OK;
}

// chunk `ck` of `nck` (nck = 1: the whole exchange buffer) of the cross stage on `stream`; period = columns per sub-row of the
// exchange pass (ignored when nck = 1)
template <int LW>
static int launch_cross(de_ctx* ctx, NttDistPlan* plan, const Fr* d_z, de_fr* const* d_out_peers, uint32_t rank, cudaStream_t stream,
                        int ck = 0, int nck = 1, unsigned long long period = 0) {
    NttCrossArgs<LW> a;
    memset(&a, 0, sizeof(a));
    const unsigned long long M = 1ull << (plan->log_n - LW);
    a.z = d_z;
    a.C = M >> LW;
    DE_TRY(ensure_rank_twiddles(ctx, plan, rank, stream));
    a.tw = plan->tw_rank;
    a.out_off = (unsigned long long)rank * a.C;
    if (nck <= 1 || period == 0 || period > a.C) {
        a.period = a.chunk_len = a.C;
        a.chunk_off = 0;
    } else {
        a.period = period;
        a.chunk_len = period / (unsigned long long)nck;
        a.chunk_off = (unsigned long long)ck * a.chunk_len;
    }
    for (int i = 0; i < (1 << LW); i++) a.peer_out[i] = (Fr*)d_out_peers[i];
    for (int i = 0; i < ((1 << LW) / 2 < 1 ? 1 : (1 << LW) / 2); i++) a.w[i] = plan->wcross[i];
    const unsigned int threads = 128;
    const unsigned long long cols = (a.C / a.period) * a.chunk_len;
    if (stream == ctx->stream) {
        DE_TIMED(ctx, "k_ntt_cross", (double)(cols << LW), (k_ntt_cross<LW><<<(unsigned int)((cols + threads - 1) / threads), threads, 0, stream>>>(a)));
    } else {
        k_ntt_cross<LW><<<(unsigned int)((cols + threads - 1) / threads), threads, 0, stream>>>(a);
    }
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}
static int launch_cross_any(de_ctx* ctx, NttDistPlan* plan, const Fr* d_z, de_fr* const* d_out_peers, uint32_t rank, cudaStream_t stream, int ck,
                            int nck, unsigned long long period) {
    switch (plan->log_w) {
        case 0: return launch_cross<0>(ctx, plan, d_z, d_out_peers, rank, stream, ck, nck, period);
        case 1: return launch_cross<1>(ctx, plan, d_z, d_out_peers, rank, stream, ck, nck, period);
        case 2: return launch_cross<2>(ctx, plan, d_z, d_out_peers, rank, stream, ck, nck, period);
        default: return launch_cross<3>(ctx, plan, d_z, d_out_peers, rank, stream, ck, nck, period);
    }
}
// the context's second stream (cross stage, high priority so that its small CTAs are placed as soon as pass CTAs retire), the
// events and the error word of the flag waits
static int dist_resources(de_ctx* c) {
    DE_CUDA(c, cudaSetDevice(c->device));
    if (!c->dist_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        DE_CUDA(c, cudaStreamCreateWithPriority(&c->dist_stream, cudaStreamNonBlocking, hi));
    }
    for (auto& e : c->dist_ev)
        if (!e) DE_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!c->dist_error) {
        DE_CUDA(c, cudaMalloc((void**)&c->dist_error, sizeof(unsigned int)));
        DE_CUDA(c, cudaMemsetAsync(c->dist_error, 0, sizeof(unsigned int), c->stream));
    }
    return DE_OK;
}

static int dist_args_ok(de_ctx* ctx, const void* a, const de_fr* omega, const void* const* peers, uint32_t world, uint32_t rank) {
    if (!a || !omega || !peers) return fail(ctx, DE_ERR_ARG, "ntt (multi-GPU): null pointer");
    if (world == 0 || world > 8 || rank >= world) return fail(ctx, DE_ERR_ARG, "ntt (multi-GPU): need rank < world <= 8");
    for (uint32_t i = 0; i < world; i++)
        if (!peers[i]) return fail(ctx, DE_ERR_ARG, "ntt (multi-GPU): null peer pointer");
    return DE_OK;
}

extern "C" {

int de_ntt_dist_stage1(de_ctx* ctx, const de_fr* d_x, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank,
                       de_fr* const* d_z_peers) {
    if (!ctx) return DE_ERR_ARG;
    DE_TRY(dist_args_ok(ctx, d_x, omega, (const void* const*)d_z_peers, world, rank));
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    NttDistPlan* plan = nullptr;
    DE_TRY(get_dist_plan(ctx, *omega, log_n, world, &plan));
    const uint32_t m = log_n - plan->log_w;
    NttDistArgs<true> dx;
    memset(&dx, 0, sizeof(dx));
    for (uint32_t i = 0; i < world; i++) dx.peer[i] = (Fr*)d_z_peers[i];
    dx.col_bits = m - plan->log_w;
    dx.row_off = (unsigned long long)rank << dx.col_bits;
    const size_t M = (size_t)1 << m;
    return ntt_run(ctx, plan->omega_local, m, (const Fr*)d_x, M, nullptr, M, 1, 0, 0, nullptr, 0, nullptr, &dx);
}

int de_ntt_dist_stage2(de_ctx* ctx, const de_fr* d_z, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank,
                       de_fr* const* d_out_peers) {
    if (!ctx) return DE_ERR_ARG;
    DE_TRY(dist_args_ok(ctx, d_z, omega, (const void* const*)d_out_peers, world, rank));
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    NttDistPlan* plan = nullptr;
    DE_TRY(get_dist_plan(ctx, *omega, log_n, world, &plan));
    return launch_cross_any(ctx, plan, (const Fr*)d_z, d_out_peers, rank, ctx->stream, 0, 1, 0);
}

// One call of the multi-GPU transform on THIS rank with the exchange pipelined in up to `chunks` column ranges and the stages
// ordered across processes by flags in peer memory (d_flag_peers[r]: rank r's DE_DIST_FLAG_WORDS u32 words, zero-initialised
// once; `epoch`: 1, 2, 3, ... the same on every rank for the same call):
//   main stream   local passes, then per range k: peer-store pass over range k, signal (k, rank) to every rank
//   cross stream  per range k: wait for (k, *) from every rank, cross stage of range k (peer stores into the output blocks);
//                 then signal (done, rank)
//   main stream   joins the cross stream and waits for (done, *): every rank's stores into this rank's block have landed and
//                 every rank has finished reading its exchange buffer (the next call may overwrite it)
// The cross stage of range k (NVLink-bound) thus runs under the pass of range k + 1 (multiply-bound).  A rank that never
// arrives makes the waits give up after ~2 s: de_ntt_dist_error() then reports it.
// Everything of de_ntt_dist_run that allocates or synchronises (plan tables, the local transform's plan and scratch, streams,
// events, this rank's inter-stage twiddles), done ahead of time.  de_ntt_dist_run does it itself on first use; a caller that
// drives SEVERAL ranks of one device from one thread must prepare all of them before the first run, because an allocation made
// while another rank's wait kernel is spinning on that device may not return before the wait gives up.
int de_ntt_dist_prepare(de_ctx* ctx, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank) {
    if (!ctx) return DE_ERR_ARG;
    if (!omega || world == 0 || world > 8 || rank >= world) return fail(ctx, DE_ERR_ARG, "de_ntt_dist_prepare: bad arguments");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    NttDistPlan* plan = nullptr;
    DE_TRY(get_dist_plan(ctx, *omega, log_n, world, &plan));
    DE_TRY(dist_resources(ctx));
    DE_TRY(ensure_rank_twiddles(ctx, plan, rank, ctx->stream));
    const uint32_t m = log_n - plan->log_w;
    NttPlan* local = nullptr;
    DE_TRY(get_plan(ctx, plan->omega_local, m, &local));
    if (local->npass > 1 && !ctx->ws[WS_NTT_SCRATCH].ensure(sizeof(Fr) << m)) return fail(ctx, DE_ERR_OOM, "de_ntt_dist_prepare: scratch allocation failed");
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

int de_ntt_dist_run(de_ctx* ctx, const de_fr* d_x, const de_fr* omega, uint32_t log_n, uint32_t world, uint32_t rank, de_fr* const* d_z_peers,
                    de_fr* const* d_out_peers, uint32_t* const* d_flag_peers, uint32_t epoch, uint32_t chunks) {
    if (!ctx) return DE_ERR_ARG;
    DE_TRY(dist_args_ok(ctx, d_x, omega, (const void* const*)d_z_peers, world, rank));
    DE_TRY(dist_args_ok(ctx, d_x, omega, (const void* const*)d_out_peers, world, rank));
    DE_TRY(dist_args_ok(ctx, d_x, omega, (const void* const*)d_flag_peers, world, rank));
    if (epoch == 0 || chunks == 0 || chunks > DE_DIST_MAX_CHUNKS) return fail(ctx, DE_ERR_ARG, "de_ntt_dist_run: epoch must be >= 1, 1 <= chunks <= 8");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    NttDistPlan* plan = nullptr;
    DE_TRY(get_dist_plan(ctx, *omega, log_n, world, &plan));
    DE_TRY(dist_resources(ctx));
    DE_TRY(ensure_rank_twiddles(ctx, plan, rank, ctx->stream));  // before anything that waits on a peer is queued
    if (!ctx->dist_peers_checked) {
        // ranks of the same process on other devices: their buffers are plain device pointers and need peer access from this
        // device (buffers of other processes arrive IPC-mapped and already belong to this device's address space)
        for (uint32_t i = 0; i < world; i++) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, d_z_peers[i]) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                    return fail(ctx, DE_ERR_CUDA, std::string("de_ntt_dist_run: no peer access to the device of rank ") + std::to_string(i));
                }
            }
            cudaGetLastError();
        }
        ctx->dist_peers_checked = true;
    }
    const uint32_t m = log_n - plan->log_w;
    const size_t M = (size_t)1 << m;
    NttDistArgs<true> dx;
    memset(&dx, 0, sizeof(dx));
    for (uint32_t i = 0; i < world; i++) dx.peer[i] = (Fr*)d_z_peers[i];
    dx.col_bits = m - plan->log_w;
    dx.row_off = (unsigned long long)rank << dx.col_bits;
    NttFlagPeers fp;
    memset(&fp, 0, sizeof(fp));
    for (uint32_t i = 0; i < world; i++) fp.flags[i] = (unsigned int*)d_flag_peers[i];
    const unsigned int* my_flags = fp.flags[rank];
    cudaStream_t sa = ctx->stream, sb = ctx->dist_stream;
    // the cross stream starts behind whatever the main stream holds now (the caller's input, the twiddle tables of a first call)
    DE_CUDA(ctx, cudaEventRecord(ctx->dist_ev[10], sa));
    DE_CUDA(ctx, cudaStreamWaitEvent(sb, ctx->dist_ev[10], 0));
    NttDistChunks ch;
    ch.chunks = (int)chunks;
    ch.after_chunk = [&](int k) -> int {
        k_flag_signal<<<1, 32, 0, sa>>>(fp, world, (unsigned int)k * 8 + rank, epoch);
        DE_CHECK_LAUNCH(ctx);
        return DE_OK;
    };
    DE_TRY(ntt_run(ctx, plan->omega_local, m, (const Fr*)d_x, M, nullptr, M, 1, 0, 0, nullptr, 0, nullptr, &dx, &ch));
    // every rank derives the same chunk count from the same shape, so the flag steps match
    const Fr* d_z = (const Fr*)d_z_peers[rank];
    for (int k = 0; k < ch.chunks; k++) {
        k_flag_wait<<<1, 32, 0, sb>>>(my_flags, world, (unsigned int)k, epoch, ctx->dist_error);
        DE_CHECK_LAUNCH(ctx);
        DE_TRY(launch_cross_any(ctx, plan, d_z, d_out_peers, rank, sb, k, ch.chunks, ch.period));
    }
    k_flag_signal<<<1, 32, 0, sb>>>(fp, world, (unsigned int)DE_DIST_DONE_STEP * 8 + rank, epoch);
    DE_CHECK_LAUNCH(ctx);
    DE_CUDA(ctx, cudaEventRecord(ctx->dist_ev[8], sb));
    DE_CUDA(ctx, cudaStreamWaitEvent(sa, ctx->dist_ev[8], 0));
    k_flag_wait<<<1, 32, 0, sa>>>(my_flags, world, (unsigned int)DE_DIST_DONE_STEP, epoch, ctx->dist_error);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

// 1 when a flag wait of de_ntt_dist_run gave up on this context since the last call of this function (synchronises the stream)
int de_ntt_dist_error(de_ctx* ctx, int* timed_out) {
    if (!ctx || !timed_out) return DE_ERR_ARG;
    *timed_out = 0;
    if (!ctx->dist_error) return DE_OK;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned int v = 0;
    DE_CUDA(ctx, cudaMemcpyAsync(&v, ctx->dist_error, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (v) DE_CUDA(ctx, cudaMemsetAsync(ctx->dist_error, 0, sizeof(v), ctx->stream));
    *timed_out = (int)v;
    return DE_OK;
}

int de_ntt_sharded_dev(de_ctx* const* ctxs, int n_gpus, const de_fr* const* d_x, de_fr* const* d_out, const de_fr* omega, uint32_t log_n) {
    if (!ctxs || n_gpus < 1 || !ctxs[0]) return DE_ERR_ARG;
    de_ctx* c0 = ctxs[0];
    if (n_gpus > 8 || (n_gpus & (n_gpus - 1))) return fail(c0, DE_ERR_ARG, "de_ntt_sharded_dev: n_gpus must be 1, 2, 4 or 8");
    if (!d_x || !d_out || !omega) return fail(c0, DE_ERR_ARG, "de_ntt_sharded_dev: null pointer");
    uint32_t lw = 0;
    while ((1 << lw) < n_gpus) lw++;
    if (log_n > 28 || log_n < 11 + lw) return fail(c0, DE_ERR_ARG, "de_ntt_sharded_dev: need 11 + log2(n_gpus) <= log_n <= 28");
    const size_t M = (size_t)1 << (log_n - lw);
    de_fr* z[8];
    for (int r = 0; r < n_gpus; r++) {
        de_ctx* c = ctxs[r];
        if (!c || !d_x[r] || !d_out[r]) return fail(c0, DE_ERR_ARG, "de_ntt_sharded_dev: null context or buffer");
        for (int q = 0; q < r; q++)
            if (ctxs[q] == c) return fail(c0, DE_ERR_ARG, "de_ntt_sharded_dev: contexts must be distinct (one per rank)");
        DE_CUDA(c0, cudaSetDevice(c->device));
        for (int q = 0; q < n_gpus; q++) {
            // every rank stores into every other rank's buffers
            const int peer = ctxs[q] ? ctxs[q]->device : c->device;
            if (peer == c->device) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                return fail(c0, DE_ERR_CUDA, std::string("de_ntt_sharded_dev: no peer access between the GPUs: ") + cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
        z[r] = (de_fr*)c->ws[WS_NTT_DIST].ensure(sizeof(Fr) * M);
        if (!z[r]) return fail(c0, DE_ERR_OOM, "de_ntt_sharded_dev: exchange buffer allocation failed");
        int rc = dist_resources(c);
        if (rc != DE_OK) return fail(c0, rc, std::string(c->err));
    }
    // Pipelined exchange (the single-process form of de_ntt_dist_run, ordered by CUDA events instead of flags): every rank's
    // peer-store pass runs as K ranges of CTAs on its main stream; the cross stage of range k runs on the rank's second stream as
    // soon as EVERY rank has finished range k, i.e. under the passes of the later ranges; the call is complete on a rank's main
    // stream when every rank's cross stage has finished (which also orders the next call's stage 1 behind the readers of the
    // exchange buffers).  Once the first kernel is queued an error must not return while peers may still be storing into buffers
    // the caller is about to reuse or free: drain() waits for every participating stream first.
    auto drain = [&](int rc, const std::string& msg) -> int {
        for (int r = 0; r < n_gpus; r++) {
            cudaSetDevice(ctxs[r]->device);
            cudaStreamSynchronize(ctxs[r]->stream);
            if (ctxs[r]->dist_stream) cudaStreamSynchronize(ctxs[r]->dist_stream);
        }
        cudaGetLastError();
        return fail(c0, rc, msg);
    };
#define DE_DIST_CUDA(expr)                                                                                   \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return drain(DE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
    int K = 0;
    unsigned long long period = 0;
    NttDistPlan* plans[8] = {};
    for (int r = 0; r < n_gpus; r++) {
        de_ctx* c = ctxs[r];
        DE_DIST_CUDA(cudaSetDevice(c->device));
        int rc = get_dist_plan(c, *omega, log_n, (uint32_t)n_gpus, &plans[r]);
        if (rc != DE_OK) return drain(rc, std::string(c->err));
        DE_DIST_CUDA(cudaEventRecord(c->dist_ev[10], c->stream));
        DE_DIST_CUDA(cudaStreamWaitEvent(c->dist_stream, c->dist_ev[10], 0));
        const uint32_t m = log_n - lw;
        NttDistArgs<true> dx;
        memset(&dx, 0, sizeof(dx));
        for (int i = 0; i < n_gpus; i++) dx.peer[i] = (Fr*)z[i];
        dx.col_bits = m - lw;
        dx.row_off = (unsigned long long)r << dx.col_bits;
        NttDistChunks ch;
        // ranges of the pipelined exchange: 1 (not pipelined) by default - measured on 8 B200s, 2 and 4 ranges are SLOWER (2^27:
        // 4.42 ms with 1, 4.49 with 2, 4.67 with 4): both stages push their peer stores through the same NVLink egress, and the
        // pass kernel's CTAs hold the whole register file, so the cross stage's CTAs only get in as the pass drains
        ch.chunks = 1;
        if (const char* e = getenv("DE_NTT_DIST_CHUNKS")) ch.chunks = atoi(e) >= 1 && atoi(e) <= DE_DIST_MAX_CHUNKS ? atoi(e) : 1;
        ch.after_chunk = [&](int k) -> int {
            cudaError_t e = cudaEventRecord(c->dist_ev[k], c->stream);
            return e == cudaSuccess ? DE_OK : fail(c, DE_ERR_CUDA, std::string("cudaEventRecord: ") + cudaGetErrorString(e));
        };
        rc = ntt_run(c, plans[r]->omega_local, m, (const Fr*)d_x[r], M, nullptr, M, 1, 0, 0, nullptr, 0, nullptr, &dx, &ch);
        if (rc != DE_OK) return drain(rc, std::string(c->err));
        K = ch.chunks;  // the same on every rank: it depends on the shape only
        period = ch.period;
    }
    for (int k = 0; k < K; k++)
        for (int q = 0; q < n_gpus; q++) {
            de_ctx* c = ctxs[q];
            DE_DIST_CUDA(cudaSetDevice(c->device));
            for (int r = 0; r < n_gpus; r++) DE_DIST_CUDA(cudaStreamWaitEvent(c->dist_stream, ctxs[r]->dist_ev[k], 0));
            int rc = launch_cross_any(c, plans[q], (const Fr*)z[q], d_out, (uint32_t)q, c->dist_stream, k, K, period);
            if (rc != DE_OK) return drain(rc, std::string(c->err));
        }
    for (int q = 0; q < n_gpus; q++) {
        DE_DIST_CUDA(cudaSetDevice(ctxs[q]->device));
        DE_DIST_CUDA(cudaEventRecord(ctxs[q]->dist_ev[8], ctxs[q]->dist_stream));
    }
    for (int q = 0; q < n_gpus; q++) {
        de_ctx* c = ctxs[q];
        DE_DIST_CUDA(cudaSetDevice(c->device));
        for (int r = 0; r < n_gpus; r++) DE_DIST_CUDA(cudaStreamWaitEvent(c->stream, ctxs[r]->dist_ev[8], 0));
    }
#undef DE_DIST_CUDA
    return DE_OK;
}

// best_fft(a, omega, log_n) on a HOST vector in natural order, spread over the GPUs of `ctxs`: block r goes up GPU r's own PCIe
// link (one host thread per GPU, as in de_commit_sharded), is dealt round-robin to the ranks' cyclic slices by peer stores,
// transformed by de_ntt_sharded_dev, and block r of the result comes back down the same link.
int de_ntt_sharded(de_ctx* const* ctxs, int n_gpus, de_fr* a, const de_fr* omega, uint32_t log_n) {
    if (!ctxs || n_gpus < 1 || !ctxs[0]) return DE_ERR_ARG;
    de_ctx* c0 = ctxs[0];
    if (n_gpus > 8 || (n_gpus & (n_gpus - 1))) return fail(c0, DE_ERR_ARG, "de_ntt_sharded: n_gpus must be 1, 2, 4 or 8");
    if (!a || !omega) return fail(c0, DE_ERR_ARG, "de_ntt_sharded: null pointer");
    uint32_t lw = 0;
    while ((1 << lw) < n_gpus) lw++;
    if (log_n > 28 || log_n < 11 + lw) return fail(c0, DE_ERR_ARG, "de_ntt_sharded: need 11 + log2(n_gpus) <= log_n <= 28");
    for (int r = 0; r < n_gpus; r++) {
        if (!ctxs[r]) return fail(c0, DE_ERR_ARG, "de_ntt_sharded: null context");
        for (int q = 0; q < r; q++)
            if (ctxs[q] == ctxs[r]) return fail(c0, DE_ERR_ARG, "de_ntt_sharded: contexts must be distinct");
    }
    const size_t M = (size_t)1 << (log_n - lw), C = M >> lw;
    Fr* stage[8];
    Fr* x[8];
    for (int r = 0; r < n_gpus; r++) {
        de_ctx* c = ctxs[r];
        DE_CUDA(c0, cudaSetDevice(c->device));
        for (int q = 0; q < n_gpus; q++) {
            if (ctxs[q]->device == c->device) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                return fail(c0, DE_ERR_CUDA, std::string("de_ntt_sharded: no peer access between the GPUs: ") + cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
        stage[r] = (Fr*)c->ws[WS_IO_A].ensure(sizeof(Fr) * M);
        x[r] = (Fr*)c->ws[WS_NTT_DIST_X].ensure(sizeof(Fr) * M);
        if (!stage[r] || !x[r]) return fail(c0, DE_ERR_OOM, "de_ntt_sharded: staging allocation failed");
        if (!c->dist_ev[9]) DE_CUDA(c0, cudaEventCreateWithFlags(&c->dist_ev[9], cudaEventDisableTiming));
    }
    std::vector<int> rc(n_gpus, DE_OK);
    auto each_gpu = [&](auto fn) {
        std::vector<std::thread> th;
        for (int r = 0; r < n_gpus; r++) th.emplace_back([&, r]() { rc[r] = fn(r); });
        for (auto& t : th) t.join();
        for (int r = 0; r < n_gpus; r++)
            if (rc[r] != DE_OK) return fail(c0, rc[r], std::string("de_ntt_sharded: GPU ") + std::to_string(r) + ": " + std::string(ctxs[r]->err));
        return (int)DE_OK;
    };
    DE_TRY(each_gpu([&](int r) -> int {
        de_ctx* c = ctxs[r];
        DE_CUDA(c, cudaSetDevice(c->device));
        DE_CUDA(c, cudaMemcpyAsync(stage[r], a + (size_t)r * M, sizeof(Fr) * M, cudaMemcpyHostToDevice, c->stream));
        NttDealArgs d;
        memset(&d, 0, sizeof(d));
        d.stage = stage[r];
        for (int q = 0; q < n_gpus; q++) d.peer_x[q] = x[q];
        d.log_w = lw;
        d.C = C;
        d.M = M;
        d.row_off = (unsigned long long)r * C;
        k_ntt_deal<<<(unsigned int)((M + 255) / 256), 256, 0, c->stream>>>(d);
        DE_CHECK_LAUNCH(c);
        DE_CUDA(c, cudaEventRecord(c->dist_ev[9], c->stream));
        return DE_OK;
    }));
    for (int q = 0; q < n_gpus; q++) {
        DE_CUDA(c0, cudaSetDevice(ctxs[q]->device));
        for (int r = 0; r < n_gpus; r++)
            if (r != q) DE_CUDA(c0, cudaStreamWaitEvent(ctxs[q]->stream, ctxs[r]->dist_ev[9], 0));
    }
    DE_TRY(de_ntt_sharded_dev(ctxs, n_gpus, (const de_fr* const*)x, (de_fr* const*)stage, omega, log_n));
    return each_gpu([&](int r) -> int {
        de_ctx* c = ctxs[r];
        DE_CUDA(c, cudaSetDevice(c->device));
        DE_CUDA(c, cudaMemcpyAsync(a + (size_t)r * M, stage[r], sizeof(Fr) * M, cudaMemcpyDeviceToHost, c->stream));
        DE_CUDA(c, cudaStreamSynchronize(c->stream));
        return DE_OK;
    });
}

// ---- device buffers that can be mapped into another process (one process per GPU: the ranks exchange these handles once,
// e.g. with torch.distributed.all_gather_object, and pass the mapped pointers to de_ntt_dist_stage1 / 2) -------------------------
int de_dev_alloc(de_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx) return DE_ERR_ARG;
    if (!d_ptr || bytes == 0) return fail(ctx, DE_ERR_ARG, "de_dev_alloc: bad argument");
    *d_ptr = nullptr;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_CUDA(ctx, cudaMalloc(d_ptr, bytes));
    return DE_OK;
}
int de_dev_free(de_ctx* ctx, void* d_ptr) {
    if (!ctx) return DE_ERR_ARG;
    if (!d_ptr) return DE_OK;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_CUDA(ctx, cudaFree(d_ptr));
    return DE_OK;
}
int de_dev_copy(de_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    if (!ctx) return DE_ERR_ARG;
    if (bytes == 0) return DE_OK;
    if (!d_dst || !d_src) return fail(ctx, DE_ERR_ARG, "de_dev_copy: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return DE_OK;
}
int de_ipc_export(de_ctx* ctx, void* d_ptr, uint8_t handle[64]) {
    if (!ctx) return DE_ERR_ARG;
    if (!d_ptr || !handle) return fail(ctx, DE_ERR_ARG, "de_ipc_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    DE_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return DE_OK;
}
int de_ipc_import(de_ctx* ctx, const uint8_t handle[64], void** d_ptr) {
    if (!ctx) return DE_ERR_ARG;
    if (!d_ptr || !handle) return fail(ctx, DE_ERR_ARG, "de_ipc_import: null pointer");
    *d_ptr = nullptr;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    DE_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DE_OK;
}
int de_ipc_release(de_ctx* ctx, void* d_ptr) {
    if (!ctx) return DE_ERR_ARG;
    if (!d_ptr) return DE_OK;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    DE_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return DE_OK;
}

// ---- EvaluationDomain -----------------------------------------------------------------------------------------
int de_domain_create(de_ctx* ctx, uint32_t j, uint32_t k, de_domain** out) {
    if (!ctx) return DE_ERR_ARG;
    if (!out) return fail(ctx, DE_ERR_ARG, "de_domain_create: out is NULL");
    *out = nullptr;
    if (j < 2) return fail(ctx, DE_ERR_ARG, "de_domain_create: cs.degree() must be >= 2");
    uint32_t ek = k;
    while (ek < 64 && ((uint64_t)1 << ek) < ((uint64_t)1 << k) * (j - 1)) ek++;
    if (k < 1 || ek > 28) return fail(ctx, DE_ERR_ARG, "de_domain_create: need 1 <= k and extended_k <= 28 (Fr::S)");
    if (ek - k > 4) return fail(ctx, DE_ERR_ARG, "de_domain_create: extension factor above 16 is not supported");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t t_len = 1u << (ek - k);
    Fr* d_consts = nullptr;
    DE_CUDA(ctx, cudaMalloc((void**)&d_consts, sizeof(Fr) * (8 + t_len)));
    k_domain_consts<<<1, 32, 0, ctx->stream>>>(k, ek, d_consts);
    ctx->launches++;
    Fr h[8 + 16];
    cudaError_t e = cudaMemcpyAsync(h, d_consts, sizeof(Fr) * (8 + t_len), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d_consts);
        return fail(ctx, DE_ERR_CUDA, std::string("de_domain_create: ") + cudaGetErrorString(e));
    }
    de_domain* d = new de_domain();
    d->ctx = ctx;
    d->j = j; d->k = k; d->ek = ek;
    d->n = (size_t)1 << k;
    d->ext_n = (size_t)1 << ek;
    d->qdeg = j - 1;
    d->ext_omega = fr_to_host(h[0]);
    d->ext_omega_inv = fr_to_host(h[1]);
    d->omega = fr_to_host(h[2]);
    d->omega_inv = fr_to_host(h[3]);
    d->zeta[0] = h[4];
    d->zeta[1] = h[5];
    d->t_len = t_len;
    // keep t_evaluations on the device; reuse the constants buffer
    d->d_t_evals = d_consts;  // entries [8, 8 + t_len)
    for (int i = 0; i < 3; i++) d->ifft_scale[i] = h[6];
    // extended_to_coeff multiplies by ext_ifft_divisor and zeta^-(i mod 3) = {1, zeta^2, zeta}; done on the device
    // by a 3-thread kernel to keep host code free of field arithmetic
    {
        Fr* tmp = d_consts;  // overwrite slots 0..2 (already copied to host)
        Fr a3[3] = {Fr::one(), h[5], h[4]};
        Fr b3[3] = {h[7], h[7], h[7]};
        Fr* d_ab = nullptr;
        if (cudaMalloc((void**)&d_ab, sizeof(Fr) * 6) != cudaSuccess) {
            cudaFree(d_consts);
            delete d;
            return fail(ctx, DE_ERR_OOM, "de_domain_create: allocation failed");
        }
        cudaMemcpyAsync(d_ab, a3, sizeof(a3), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_ab + 3, b3, sizeof(b3), cudaMemcpyHostToDevice, ctx->stream);
        k_vec_op<Fr><<<1, 32, 0, ctx->stream>>>(DE_OP_MUL, d_ab, d_ab + 3, tmp, 3);
        ctx->launches++;
        cudaMemcpyAsync(d->ext_out_scale, tmp, sizeof(Fr) * 3, cudaMemcpyDeviceToHost, ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_ab);
        if (e != cudaSuccess) {
            cudaFree(d_consts);
            delete d;
            return fail(ctx, DE_ERR_CUDA, std::string("de_domain_create: ") + cudaGetErrorString(e));
        }
    }
    *out = d;
    return DE_OK;
}

int de_domain_free(de_domain* d) {
    if (!d) return DE_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    if (d->d_t_evals) cudaFree(d->d_t_evals);
    delete d;
    return DE_OK;
}

int de_domain_info(de_domain* d, uint32_t* extended_k, de_fr consts[4]) {
    if (!d) return DE_ERR_ARG;
    if (extended_k) *extended_k = d->ek;
    if (consts) {
        consts[0] = d->omega;
        consts[1] = d->omega_inv;
        consts[2] = d->ext_omega;
        consts[3] = d->ext_omega_inv;
    }
    return DE_OK;
}

int de_coeff_to_extended_dev(de_domain* d, const de_fr* d_coeff, size_t in_stride, de_fr* d_ext, size_t out_stride, size_t batch) {
    if (!d) return DE_ERR_ARG;
    de_ctx* ctx = d->ctx;
    if (!d_coeff || !d_ext) return fail(ctx, DE_ERR_ARG, "de_coeff_to_extended_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, d->ext_omega, d->ek, (const Fr*)d_coeff, in_stride, (Fr*)d_ext, out_stride, batch, 1, d->n, d->zeta, 0, nullptr);
}
int de_extended_to_coeff_dev(de_domain* d, de_fr* d_ext, size_t stride, size_t batch, size_t* out_len) {
    if (!d) return DE_ERR_ARG;
    de_ctx* ctx = d->ctx;
    if (!d_ext) return fail(ctx, DE_ERR_ARG, "de_extended_to_coeff_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (out_len) *out_len = d->n * d->qdeg;
    return ntt_run(ctx, d->ext_omega_inv, d->ek, (const Fr*)d_ext, stride, (Fr*)d_ext, stride, batch, 0, 0, nullptr, 1, d->ext_out_scale);
}
int de_lagrange_to_coeff_dev(de_domain* d, de_fr* d_a, size_t stride, size_t batch) {
    if (!d) return DE_ERR_ARG;
    de_ctx* ctx = d->ctx;
    if (!d_a) return fail(ctx, DE_ERR_ARG, "de_lagrange_to_coeff_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, d->omega_inv, d->k, (const Fr*)d_a, stride, (Fr*)d_a, stride, batch, 0, 0, nullptr, 1, d->ifft_scale);
}
int de_coeff_to_lagrange_dev(de_domain* d, de_fr* d_a, size_t stride, size_t batch) {
    if (!d) return DE_ERR_ARG;
    de_ctx* ctx = d->ctx;
    if (!d_a) return fail(ctx, DE_ERR_ARG, "de_coeff_to_lagrange_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return ntt_run(ctx, d->omega, d->k, (const Fr*)d_a, stride, (Fr*)d_a, stride, batch, 0, 0, nullptr, 0, nullptr);
}
int de_divide_by_vanishing_dev(de_domain* d, de_fr* d_ext, size_t stride, size_t batch) {
    if (!d) return DE_ERR_ARG;
    de_ctx* ctx = d->ctx;
    if (!d_ext) return fail(ctx, DE_ERR_ARG, "de_divide_by_vanishing_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    for (size_t b = 0; b < batch; b++) {
        unsigned int threads = 256;
        k_scale_periodic<<<(unsigned int)((d->ext_n + threads - 1) / threads), threads, 0, ctx->stream>>>(
            (Fr*)d_ext + b * stride, d->ext_n, d->d_t_evals + 8, d->t_len - 1);
        DE_CHECK_LAUNCH(ctx);
    }
    return DE_OK;
}

// host-pointer wrappers: stage through the context's I/O workspace
static int host_roundtrip(de_domain* d, const de_fr* in, size_t n_in, de_fr* out, size_t n_out, int which) {
    de_ctx* ctx = d->ctx;
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t cap = n_in > n_out ? n_in : n_out;
    if (which == 0) cap = d->n + d->ext_n;
    DE_WS(ctx, buf, Fr, WS_IO_A, sizeof(Fr) * cap);
    DE_CUDA(ctx, cudaMemcpyAsync(buf, in, sizeof(Fr) * n_in, cudaMemcpyHostToDevice, ctx->stream));
    Fr* res = buf;
    switch (which) {
        case 0:
            res = buf + d->n;
            DE_TRY(de_coeff_to_extended_dev(d, (const de_fr*)buf, d->n, (de_fr*)res, d->ext_n, 1));
            break;
        case 1: DE_TRY(de_extended_to_coeff_dev(d, (de_fr*)buf, d->ext_n, 1, nullptr)); break;
        case 2: DE_TRY(de_lagrange_to_coeff_dev(d, (de_fr*)buf, d->n, 1)); break;
        case 3: DE_TRY(de_coeff_to_lagrange_dev(d, (de_fr*)buf, d->n, 1)); break;
        case 4: DE_TRY(de_divide_by_vanishing_dev(d, (de_fr*)buf, d->ext_n, 1)); break;
    }
    DE_CUDA(ctx, cudaMemcpyAsync(out, res, sizeof(Fr) * n_out, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

int de_coeff_to_extended(de_domain* d, const de_fr* coeff_n, de_fr* ext_out) {
    if (!d) return DE_ERR_ARG;
    if (!coeff_n || !ext_out) return fail(d->ctx, DE_ERR_ARG, "de_coeff_to_extended: null pointer");
    return host_roundtrip(d, coeff_n, d->n, ext_out, d->ext_n, 0);
}
int de_extended_to_coeff(de_domain* d, de_fr* a, size_t* out_len) {
    if (!d) return DE_ERR_ARG;
    if (!a) return fail(d->ctx, DE_ERR_ARG, "de_extended_to_coeff: null pointer");
    if (out_len) *out_len = d->n * d->qdeg;
    return host_roundtrip(d, a, d->ext_n, a, d->ext_n, 1);
}
int de_lagrange_to_coeff(de_domain* d, de_fr* a) {
    if (!d) return DE_ERR_ARG;
    if (!a) return fail(d->ctx, DE_ERR_ARG, "de_lagrange_to_coeff: null pointer");
    return host_roundtrip(d, a, d->n, a, d->n, 2);
}
int de_coeff_to_lagrange(de_domain* d, de_fr* a) {
    if (!d) return DE_ERR_ARG;
    if (!a) return fail(d->ctx, DE_ERR_ARG, "de_coeff_to_lagrange: null pointer");
    return host_roundtrip(d, a, d->n, a, d->n, 3);
}
int de_divide_by_vanishing(de_domain* d, de_fr* a) {
    if (!d) return DE_ERR_ARG;
    if (!a) return fail(d->ctx, DE_ERR_ARG, "de_divide_by_vanishing: null pointer");
    return host_roundtrip(d, a, d->ext_n, a, d->ext_n, 4);
}

}  // extern "C"
