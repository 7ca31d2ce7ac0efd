// common.cuh — host-side plumbing shared by the translation units of libde_b200.so: the context object, grow-only
// device workspaces, error capture.  No arithmetic lives here.
#pragma once
#include <cuda_runtime.h>
#include <sched.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/de_b200.h"
#include "field.cuh"

namespace de {

struct NttPlan;
struct NttDistPlan;
template <bool DIST>
struct NttDistArgs;
struct NttDistChunks;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // returns nullptr on allocation failure
    void* ensure(size_t bytes) {
        if (bytes <= cap && p) return p;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes < 256 ? 256 : bytes;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return nullptr;
        }
        cap = want;
        return p;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes) {
        if (bytes <= cap && p) return p;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes < 256 ? 256 : bytes;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return nullptr;
        }
        cap = want;
        return p;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

enum { WS_IO_A = 0, WS_IO_B, WS_NTT_SCRATCH, WS_MSM_KEYS, WS_MSM_VALS, WS_MSM_SORTED, WS_MSM_COUNTS, WS_MSM_BUCKETS,
       WS_MSM_PARTIALS, WS_MSM_MISC, WS_MSM_RED, WS_MSM_OUT, WS_EVAL_A, WS_EVAL_B, WS_EVAL_C, WS_NTT_DIST, WS_NTT_DIST_X, WS_POLY_SCRATCH, WS_COUNT };

}  // namespace de

namespace de {
// optional per-kernel CUDA-event timing on the context's stream (bench.py's roofline section)
struct TimedLaunch {
    const char* name;
    cudaEvent_t e0, e1;
    double units;  // algorithmic units processed by this launch (points, elements, rows)
};
struct KernelStat {
    std::string name;
    double ms = 0, units = 0;
    uint64_t launches = 0;
};
// one captured launch sequence of a commitment round (msm.cu): replayed while every pointer, size and workspace address of the
// call is the same as at capture time
struct MsmGraph {
    std::vector<uint64_t> key;
    cudaGraphExec_t exec = nullptr;
    uint32_t seen = 0;       // calls with this key so far (the second one captures)
    bool eager_only = false; // a capture of this key failed once: never again
    uint64_t launches = 0;   // kernels inside the graph (de_launch_count keeps counting them on replay)
    uint64_t last_used = 0;
};
}  // namespace de

struct de_ctx {
    bool timing = false;
    std::vector<de::TimedLaunch> timed;
    std::vector<de::KernelStat> stats;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int sm_count = 148;
    int mode = 0;  // DE_MODE_LATENCY / DE_MODE_THROUGHPUT (de_ctx_set_mode)
    std::string err;
    uint64_t launches = 0;
    de::DevBuf ws[de::WS_COUNT];
    std::vector<de::MsmGraph> msm_graphs;
    uint64_t msm_graph_clock = 0;
    de::PinnedBuf pinned;
    std::vector<de::NttPlan*> plans;
    std::vector<de::NttDistPlan*> dist_plans;           // multi-GPU transform tables (ntt.cu)
    // multi-GPU transform (ntt.cu): events [0 .. 7] = exchange-pass chunk k complete, [8] = cross stage complete, [9] = deal
    // complete, [10] = scratch; a second, high-priority stream for the cross stage; the error word of the flag waits
    cudaEvent_t dist_ev[11] = {};
    cudaStream_t dist_stream = nullptr;
    unsigned int* dist_error = nullptr;
    bool dist_peers_checked = false;  // peer access towards the devices of the other ranks' buffers enabled (same-process ranks)
};

namespace de {

extern thread_local std::string g_create_error;

inline int fail(de_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

#define DE_CUDA(ctx, expr)                                                                                  \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            cudaGetLastError();                                                                             \
            return de::fail((ctx), e__ == cudaErrorMemoryAllocation ? DE_ERR_OOM : DE_ERR_CUDA,             \
                            std::string(#expr) + ": " + cudaGetErrorString(e__));                           \
        }                                                                                                   \
    } while (0)

#define DE_CHECK_LAUNCH(ctx)                                                                                \
    do {                                                                                                    \
        (ctx)->launches++;                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                               \
        if (e__ != cudaSuccess)                                                                             \
            return de::fail((ctx), DE_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__) +  \
                                                    " at " + __FILE__ + ":" + std::to_string(__LINE__));    \
    } while (0)

// DE_TIMED(ctx, "kernel", units, launch-statement): wraps the launch in a CUDA event pair when timing is enabled
#define DE_TIMED(ctx, kname, nunits, stmt)                                      \
    do {                                                                        \
        de::TimedLaunch tl__ = {kname, nullptr, nullptr, (double)(nunits)};     \
        if ((ctx)->timing) {                                                    \
            cudaEventCreate(&tl__.e0);                                          \
            cudaEventCreate(&tl__.e1);                                          \
            cudaEventRecord(tl__.e0, (ctx)->stream);                            \
        }                                                                       \
        stmt;                                                                   \
        if ((ctx)->timing) {                                                    \
            cudaEventRecord(tl__.e1, (ctx)->stream);                            \
            (ctx)->timed.push_back(tl__);                                       \
        }                                                                       \
    } while (0)

// the same for a sequence of launches: timing_begin / timing_end bracket them with one event pair
inline TimedLaunch timing_begin(de_ctx* ctx, const char* name, double units) {
    TimedLaunch tl = {name, nullptr, nullptr, units};
    if (ctx->timing) {
        cudaEventCreate(&tl.e0);
        cudaEventCreate(&tl.e1);
        cudaEventRecord(tl.e0, ctx->stream);
    }
    return tl;
}
inline void timing_end(de_ctx* ctx, TimedLaunch& tl) {
    if (ctx->timing) {
        cudaEventRecord(tl.e1, ctx->stream);
        ctx->timed.push_back(tl);
    }
}

// Host wait for everything queued on `st`.  DE_MODE_LATENCY spins inside the driver (cudaStreamSynchronize: lowest wake-up
// latency, one proof owns the machine).  DE_MODE_THROUGHPUT polls and YIELDS between polls: with 8 provers per GPU (x 8 GPUs on a
// 32-core host) every prover thread waits several times per proof, and 64 spinning waiters would hold the cores the witness
// passes of the end-to-end path need; sched_yield returns at once while cores are idle, so nothing is lost at one GPU.  (A
// cudaEventBlockingSync wait was measured too: +2 % end to end on 8 GPUs but -2 % on one, from the interrupt wake-up latency.)
inline cudaError_t stream_wait(de_ctx* ctx, cudaStream_t st) {
    static const bool spin = getenv("DE_WAIT_SPIN") != nullptr;  // A/B switch for measurements
    if (ctx->mode != DE_MODE_THROUGHPUT || spin) return cudaStreamSynchronize(st);
    for (;;) {
        cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
        sched_yield();
    }
}

#define DE_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != DE_OK) return rc__; \
    } while (0)

#define DE_WS(ctx, var, type, slot, bytes)                                                      \
    type* var = (type*)(ctx)->ws[slot].ensure(bytes);                                           \
    if (!var) return de::fail((ctx), DE_ERR_OOM, std::string("device workspace allocation of ") + \
                                                     std::to_string((size_t)(bytes)) + " bytes failed")

inline Fr fr_from_host(const de_fr& v) {
    Fr r;
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)v.l[i];
        r.l[2 * i + 1] = (uint32_t)(v.l[i] >> 32);
    }
    return r;
}
inline de_fr fr_to_host(const Fr& v) {
    de_fr r;
    for (int i = 0; i < 4; i++) r.l[i] = (uint64_t)v.l[2 * i] | ((uint64_t)v.l[2 * i + 1] << 32);
    return r;
}

// ntt.cu
int ntt_run(de_ctx* ctx, const de_fr& omega, uint32_t log_n, const Fr* d_src, size_t src_stride, Fr* d_dst, size_t dst_stride,
            size_t batch, int in_mode, size_t n_in, const Fr* zeta2, int out_mode, const Fr* oscale3,
            const NttDistArgs<true>* dist = nullptr, NttDistChunks* chunks = nullptr);
void ntt_free_plans(de_ctx* ctx);
// device-side Fr helpers running single-thread kernels, used for domain constants (ntt.cu)
int fr_host_pow(de_ctx* ctx, const de_fr& base, uint64_t e, de_fr* out);

}  // namespace de
