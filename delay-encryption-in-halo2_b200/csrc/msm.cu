// msm.cu — host side of the MSM entry points: window-shape selection, workspace layout, the launch sequence of
// msm.cuh, and the C ABI functions de_msm*, de_params_*, de_commit*, de_g1_sum.
// Reference semantics: halo2_proofs::arithmetic::best_multiexp and ParamsKZG::{commit, commit_lagrange}
// (SURVEY.md Appendix B.1 / B.4).
#include <stdlib.h>
#include <string.h>

#include <thread>

#include "common.cuh"
#include "msm.cuh"
#include "transcript.hpp"

namespace de {

// msm_reduce.cu: bucket sets -> one Jacobian point per polynomial at d_out (steps 4 and 5 of the pipeline)
int msm_reduce_enqueue(de_ctx* ctx, cudaStream_t st, unsigned int c, unsigned int nsets, unsigned int NB, size_t count, size_t n, XYZZ* buckets,
                       XYZZ* red, XYZZ* red2, Jac* d_out);

struct MsmCfg {
    unsigned int c, W, nsets, ntables;
};

static unsigned int windows_for(unsigned int c) { return (255 + c - 1) / c; }

// cost model in field multiplications: 10 per mixed add in the bucket fill, ~40 per bucket in the reduction
static MsmCfg choose_cfg(unsigned long long n, bool precomputed, unsigned long long max_tables) {
    MsmCfg best = {0, 0, 0, 0};
    double best_cost = 1e300;
    const char* env = getenv("DE_MSM_C");
    unsigned int forced = env ? (unsigned int)atoi(env) : 0;
    for (unsigned int c = 8; c <= 20; c++) {
        if (forced && c != forced) continue;
        unsigned int W = windows_for(c);
        unsigned int ntables = 1, nsets = W;
        if (precomputed) {
            ntables = W;
            if (ntables > max_tables) ntables = (unsigned int)max_tables;
            if (ntables < 1) ntables = 1;
            nsets = (W + ntables - 1) / ntables;
            ntables = (W + nsets - 1) / nsets;
        }
        double cost = 10.0 * (double)n * W + 30.0 * (double)nsets * (double)(1u << (c - 1));  // 2 full additions (14 muls) per bucket
        if (cost < best_cost) {
            best_cost = cost;
            best = {c, W, nsets, ntables};
        }
    }
    return best;
}

// exclusive scan of n u32 values (three launches); out has n entries, *grand_total receives the sum.  block_sums: scratch of
// DE_SCAN_THREADS * DE_SCAN_ITEMS words.  Also used by prover.cu for the lookup compaction indices.
int scan_u32(de_ctx* ctx, const unsigned int* in, unsigned long long n, unsigned int* out, unsigned int* block_sums,
             unsigned int* grand_total) {
    const unsigned long long per_block = DE_SCAN_THREADS * DE_SCAN_ITEMS;
    unsigned long long nblocks = (n + per_block - 1) / per_block;
    if (nblocks > per_block) return fail(ctx, DE_ERR_UNSUPPORTED, "msm: too many buckets for the scan");
    k_scan_blocks<<<(unsigned int)nblocks, DE_SCAN_THREADS, 0, ctx->stream>>>(in, n, out, block_sums);
    DE_CHECK_LAUNCH(ctx);
    k_scan_tops<<<1, DE_SCAN_THREADS, 0, ctx->stream>>>(block_sums, (unsigned int)nblocks, grand_total);
    DE_CHECK_LAUNCH(ctx);
    k_scan_add<<<(unsigned int)((n + 255) / 256), 256, 0, ctx->stream>>>(out, n, block_sums);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// d_scalars: `count` polynomials of n Montgomery scalars, `stride` elements apart.  d_tables: cfg.ntables base tables,
// table_stride elements apart.  Queues the whole launch sequence on the context's stream and leaves `count` Jacobian points at
// d_out (device); *d_entries (device, may be null) receives the number of bucket additions.
static int msm_enqueue(de_ctx* ctx, const Fr* d_scalars, size_t stride, size_t n, size_t count, const Affine* d_tables,
                       size_t table_stride, size_t base_offset, const MsmCfg& cfg, Jac* d_out, unsigned int alt_first, long long alt_delta,
                       unsigned int* d_entries) {
    if ((unsigned long long)cfg.ntables * table_stride >= (1ull << 31) ||
        (alt_first != 0xffffffffu && 2ull * cfg.ntables * table_stride >= (1ull << 31)))
        return fail(ctx, DE_ERR_UNSUPPORTED, "msm: base table exceeds 2^31 points");
    MsmShape sh;
    sh.c = cfg.c; sh.W = cfg.W; sh.nsets = cfg.nsets; sh.NB = 1u << (cfg.c - 1);
    sh.count = (unsigned int)count; sh.n = n; sh.table_stride = table_stride; sh.base_offset = base_offset;
    sh.alt_first = alt_first; sh.alt_delta = alt_delta;
    const unsigned long long E = (unsigned long long)count * sh.W * n;
    const unsigned long long nbuckets = (unsigned long long)count * sh.nsets * sh.NB;
    if (E >= (1ull << 32) || nbuckets >= (1ull << 31)) return fail(ctx, DE_ERR_UNSUPPORTED, "msm: batch too large for 32-bit entry indices");
    // task length CH is chosen on the device from the actual entry count (k_msm_choose_ch); the host only bounds the task
    // count for the allocation and the grid: tasks <= E / 8 + nbuckets
    const unsigned int target_tasks = (unsigned int)ctx->sm_count * 1024;
    const unsigned long long max_tasks = E / 8 + nbuckets + 1;
    const unsigned int ndigits = (sh.c - 1 + 4) / 5;
    const unsigned int nsets_total = (unsigned int)(count * sh.nsets);

    static const char* force_ranks = getenv("DE_SCATTER_RANKS");  // A/B switch for measurements: "0" / "1"
    const bool use_ranks = force_ranks ? force_ranks[0] == '1' : n <= (1ull << 20);  // measured crossover, see k_msm_scatter
    DE_WS(ctx, keys, unsigned int, WS_MSM_KEYS, sizeof(unsigned int) * (use_ranks ? 2 : 1) * E);
    unsigned int* ranks = use_ranks ? keys + E : nullptr;  // rank of every entry inside its bucket
    DE_WS(ctx, vals, unsigned int, WS_MSM_VALS, sizeof(unsigned int) * E);
    DE_WS(ctx, sorted, unsigned int, WS_MSM_SORTED, sizeof(unsigned int) * (E + 2) + sizeof(uint2) * max_tasks);
    uint2* task_list = (uint2*)(sorted + ((E + 1) & ~1ull));
    // u32 arrays of nbuckets + 1 entries each: counts, offsets, cursor, ntasks, task_off, multi_small, multi_large; then scan
    // block sums, scalars and the task-length bins
    const size_t nb1 = align_up(nbuckets + 1, 64);
    const size_t misc_words = nb1 * 7 + 2 * (DE_SCAN_THREADS * DE_SCAN_ITEMS) + 64 + 2 * 256;
    DE_WS(ctx, misc, unsigned int, WS_MSM_COUNTS, sizeof(unsigned int) * misc_words);
    unsigned int* counts = misc;
    unsigned int* offsets = counts + nb1;
    unsigned int* cursor = offsets + nb1;
    unsigned int* ntasks = cursor + nb1;
    unsigned int* task_off = ntasks + nb1;
    unsigned int* multi_small = task_off + nb1;
    unsigned int* multi_large = multi_small + nb1;
    unsigned int* block_sums = multi_large + nb1;
    unsigned int* block_sums2 = block_sums + DE_SCAN_THREADS * DE_SCAN_ITEMS;
    unsigned int* scalars_u32 = block_sums2 + DE_SCAN_THREADS * DE_SCAN_ITEMS;  // [0] entries, [1] tasks, [2] multi_small, [3] multi_large
    unsigned int* len_bins = scalars_u32 + 64;  // 256 words
    unsigned int* bin_cursor = len_bins + 256;  // 256 words
    DE_WS(ctx, buckets, XYZZ, WS_MSM_BUCKETS, sizeof(XYZZ) * nbuckets);
    DE_WS(ctx, partials, XYZZ, WS_MSM_PARTIALS, sizeof(XYZZ) * max_tasks);
    DE_WS(ctx, red, XYZZ, WS_MSM_MISC, sizeof(XYZZ) * ((size_t)ndigits * 32 * nsets_total + nsets_total));
    // scratch of the two-digit reduction: two ping-pong partial buffers, D0 / D1 and their digit sums
    DE_WS(ctx, red2, XYZZ, WS_MSM_RED, sizeof(XYZZ) * ((size_t)nsets_total * sh.NB + 4096 + (size_t)nsets_total * 4 * 2048));

    cudaStream_t st = ctx->stream;
    DE_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(unsigned int) * nb1, st));
    DE_CUDA(ctx, cudaMemsetAsync(ntasks, 0, sizeof(unsigned int) * nb1, st));
    DE_CUDA(ctx, cudaMemsetAsync(scalars_u32, 0, sizeof(unsigned int) * (64 + 512), st));
    DE_CUDA(ctx, cudaMemsetAsync(buckets, 0, sizeof(XYZZ) * nbuckets, st));

    const unsigned long long nscal = (unsigned long long)n * count;
    static const char* force_agg = getenv("DE_DIGITS_AGG");  // A/B switch for measurements: "0" / "1"
    if (force_agg ? force_agg[0] == '1' : true)
        k_msm_digits<true><<<(unsigned int)((nscal + 127) / 128), 128, 0, st>>>(d_scalars, stride, sh, keys, vals, ranks, counts);
    else
        k_msm_digits<false><<<(unsigned int)((nscal + 127) / 128), 128, 0, st>>>(d_scalars, stride, sh, keys, vals, ranks, counts);
    DE_CHECK_LAUNCH(ctx);
    DE_TRY(scan_u32(ctx, counts, nbuckets + 1, offsets, block_sums, &scalars_u32[0]));
    if (use_ranks) {
        k_msm_scatter<<<(unsigned int)((E + 255) / 256), 256, 0, st>>>(keys, vals, ranks, E, offsets, sorted);
    } else {
        DE_CUDA(ctx, cudaMemcpyAsync(cursor, offsets, sizeof(unsigned int) * (nbuckets + 1), cudaMemcpyDeviceToDevice, st));
        k_msm_scatter_cursor<<<(unsigned int)((E + 255) / 256), 256, 0, st>>>(keys, vals, E, cursor, sorted);
    }
    DE_CHECK_LAUNCH(ctx);
    k_msm_choose_ch<<<1, 32, 0, st>>>(scalars_u32, (unsigned int)nbuckets, DE_MSM_MAX_CH, target_tasks);
    DE_CHECK_LAUNCH(ctx);
    k_msm_task_counts<<<(unsigned int)((nbuckets + 255) / 256), 256, 0, st>>>(counts, (unsigned int)nbuckets, ntasks, multi_small, multi_large,
                                                                             scalars_u32, len_bins);
    DE_CHECK_LAUNCH(ctx);
    DE_TRY(scan_u32(ctx, ntasks, nbuckets + 1, task_off, block_sums2, &scalars_u32[1]));
    k_msm_bin_starts<<<1, 32, 0, st>>>(len_bins, scalars_u32, bin_cursor);
    DE_CHECK_LAUNCH(ctx);
    k_msm_task_fill<<<(unsigned int)((nbuckets + 255) / 256), 256, 0, st>>>(counts, (unsigned int)nbuckets, scalars_u32, bin_cursor, task_list);
    DE_CHECK_LAUNCH(ctx);
    DE_TIMED(ctx, "k_msm_accumulate", (double)n * count,
             (k_msm_accumulate<<<(unsigned int)((max_tasks + 127) / 128), 128, 0, st>>>(sorted, offsets, counts, task_off, task_list,
                                                                                       scalars_u32, d_tables, buckets, partials)));
    DE_CHECK_LAUNCH(ctx);
    k_msm_merge_small<<<ctx->sm_count * 4, 128, 0, st>>>(multi_small, scalars_u32, task_off, partials, buckets);
    DE_CHECK_LAUNCH(ctx);
    // A/B switch for measurements: "0" / "1".  A warp per 9 .. 32-partial bucket measured neutral (8.98 vs 9.02 ms per proof,
    // 158.8 vs 158.7 proofs/s): off by default
    static const char* merge_env = getenv("DE_MERGE_WARP");
    if (merge_env ? merge_env[0] == '1' : false)
        k_msm_merge_large<true><<<ctx->sm_count * 4, 128, 0, st>>>(multi_large, scalars_u32, task_off, partials, buckets);
    else
        k_msm_merge_large<false><<<ctx->sm_count * 4, 128, 0, st>>>(multi_large, scalars_u32, task_off, partials, buckets);
    DE_CHECK_LAUNCH(ctx);
    DE_TRY(msm_reduce_enqueue(ctx, st, sh.c, sh.nsets, sh.NB, count, n, buckets, red, red2, d_out));
    if (d_entries) DE_CUDA(ctx, cudaMemcpyAsync(d_entries, &scalars_u32[0], sizeof(unsigned int), cudaMemcpyDeviceToDevice, st));
    return DE_OK;
}

// The launch sequence of msm_enqueue as a CUDA graph.  A proof repeats the same commitment rounds on the same buffers, and at
// the small circuits (pose_enc, k = 11: ~20 kernels of 3-6 us per round) the rounds are bound by launch and dependency latency,
// not by work.  The first call with a given key runs eagerly (and sizes the grow-only workspaces), the second one is captured,
// later ones replay.  The key holds every pointer, size and workspace address the sequence bakes in, so a reallocated workspace or
// another prover's columns simply miss.  DE_MODE_LATENCY only: measured on the B200, one proof in flight gains (pose_enc 1.86 ->
// 1.67 ms, mod_pow 15.45 -> 15.26, delay_enc 8.39 -> 8.35), while eight provers replaying graphs from eight host threads LOSE
// (pose_enc 1678 -> 1575 proofs/s, delay_enc 158.1 -> 156.4): a graph launch is one unit of work to the driver and the streams
// interleave more coarsely.  DE_MSM_GRAPH=0 switches it off, =2 forces it in throughput mode too; per-kernel timing
// (de_timing_enable) bypasses it.
static int msm_enqueue_graphed(de_ctx* ctx, const Fr* d_scalars, size_t stride, size_t n, size_t count, const Affine* d_tables,
                               size_t table_stride, size_t base_offset, const MsmCfg& cfg, Jac* d_out, unsigned int alt_first, long long alt_delta) {
    static const char* env = getenv("DE_MSM_GRAPH");
    const bool enabled = !(env && env[0] == '0') && !ctx->timing && n <= (1ull << 17) &&
                         (ctx->mode == DE_MODE_LATENCY || (env && env[0] == '2'));
    if (!enabled) return msm_enqueue(ctx, d_scalars, stride, n, count, d_tables, table_stride, base_offset, cfg, d_out, alt_first, alt_delta, nullptr);
    std::vector<uint64_t> key = {(uint64_t)d_scalars, stride, n, count, (uint64_t)d_tables, table_stride, base_offset,
                                 ((uint64_t)cfg.c << 48) | ((uint64_t)cfg.W << 32) | ((uint64_t)cfg.nsets << 16) | cfg.ntables, (uint64_t)d_out,
                                 alt_first, (uint64_t)alt_delta, (uint64_t)ctx->mode, (uint64_t)ctx->sm_count};
    for (int s = WS_MSM_KEYS; s <= WS_MSM_OUT; s++) key.push_back((uint64_t)ctx->ws[s].p);
    MsmGraph* slot = nullptr;
    for (auto& g : ctx->msm_graphs)
        if (g.key == key) slot = &g;
    ctx->msm_graph_clock++;
    if (!slot) {
        if (ctx->msm_graphs.size() >= 32) {  // evict the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < ctx->msm_graphs.size(); i++)
                if (ctx->msm_graphs[i].last_used < ctx->msm_graphs[lru].last_used) lru = i;
            if (ctx->msm_graphs[lru].exec) cudaGraphExecDestroy(ctx->msm_graphs[lru].exec);
            ctx->msm_graphs.erase(ctx->msm_graphs.begin() + lru);
        }
        ctx->msm_graphs.emplace_back();
        slot = &ctx->msm_graphs.back();
        slot->key = key;
    }
    slot->last_used = ctx->msm_graph_clock;
    slot->seen++;
    if (slot->exec) {
        DE_CUDA(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
        ctx->launches += slot->launches;
        return DE_OK;
    }
    if (slot->seen < 2 || slot->eager_only) {
        const int rc = msm_enqueue(ctx, d_scalars, stride, n, count, d_tables, table_stride, base_offset, cfg, d_out, alt_first, alt_delta, nullptr);
        // the eager run may have grown a workspace: the key to match next time carries the addresses as they are now
        for (int s = WS_MSM_KEYS; s <= WS_MSM_OUT; s++) slot->key[13 + (s - WS_MSM_KEYS)] = (uint64_t)ctx->ws[s].p;
        return rc;
    }
    const uint64_t launches_before = ctx->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        // e.g. the caller is capturing this stream itself: leave its capture alone and launch eagerly from now on
        cudaGetLastError();
        slot->eager_only = true;
        return msm_enqueue(ctx, d_scalars, stride, n, count, d_tables, table_stride, base_offset, cfg, d_out, alt_first, alt_delta, nullptr);
    }
    const int rc = msm_enqueue(ctx, d_scalars, stride, n, count, d_tables, table_stride, base_offset, cfg, d_out, alt_first, alt_delta, nullptr);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    const uint64_t in_graph = ctx->launches - launches_before;
    ctx->launches = launches_before;
    cudaGraphExec_t exec = nullptr;
    if (rc == DE_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
        slot->exec = exec;
        slot->launches = in_graph;
        DE_CUDA(ctx, cudaGraphLaunch(exec, ctx->stream));
        ctx->launches += in_graph;
        return DE_OK;
    }
    // capture failed: nothing was executed; clear the error state and run eagerly, now and whenever this key comes back
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    slot->eager_only = true;
    return msm_enqueue(ctx, d_scalars, stride, n, count, d_tables, table_stride, base_offset, cfg, d_out, alt_first, alt_delta, nullptr);
}

// Writes `count` points to host_out.  out_mode 0: Jacobian points (de_g1, Montgomery); 1: canonical affine x || y, 64 bytes per
// point (transcript form).  `dense` = the caller expects full-width scalars in every polynomial (the grand products of a proof).
// DE_MSM_SPLIT=<m> (measurement switch, off by default): with one proof in flight a dense batch of more than m proof-sized
// polynomials goes through the launch sequence in two halves.  Tried because the sort of 8 x 2^16 x 16 entries works on 134 MB of
// keys, values, ranks and sorted values - more than the L2 holds; measured on the B200 it LOSES 0.33 ms per proof (9.02 vs 8.69 ms):
// the second half's latency-bound reduction tail costs more than the smaller working set saves.
static int msm_core(de_ctx* ctx, const Fr* d_scalars, size_t stride, size_t n, size_t count, const Affine* d_tables,
                    size_t table_stride, size_t base_offset, const MsmCfg& cfg, void* host_out, int out_mode = 0,
                    unsigned int alt_first = 0xffffffffu, long long alt_delta = 0, bool dense = false) {
    if (count == 0) return DE_OK;
    if (n == 0) {
        memset(host_out, 0, (out_mode ? sizeof(de_g1_affine) : sizeof(de_g1)) * count);
        return DE_OK;
    }
    static const char* split_env = getenv("DE_MSM_SPLIT");  // largest batch that is not split; 0 = never split
    const size_t split_above = split_env ? (size_t)atoll(split_env) : 0;
    const bool split = dense && split_above && count > split_above && ctx->mode == DE_MODE_LATENCY && n <= (1ull << 18);
    DE_WS(ctx, d_out_all, Jac, WS_MSM_OUT, sizeof(Jac) * count + 2 * sizeof(unsigned int));
    unsigned int* d_entries = (unsigned int*)(d_out_all + count);
    cudaStream_t st = ctx->stream;
    const size_t first = split ? (count + 1) / 2 : count;  // the larger half first: the grow-only workspaces are sized once
    for (size_t c0 = 0, part = 0; c0 < count; c0 += first, part++) {
        const size_t cnt = c0 + first <= count ? first : count - c0;
        unsigned int af = alt_first == 0xffffffffu ? alt_first : alt_first > c0 ? (unsigned int)(alt_first - c0) : 0u;
        if (ctx->timing)
            DE_TRY(msm_enqueue(ctx, d_scalars + c0 * stride, stride, n, cnt, d_tables, table_stride, base_offset, cfg, d_out_all + c0, af, alt_delta,
                               d_entries + part));
        else
            DE_TRY(msm_enqueue_graphed(ctx, d_scalars + c0 * stride, stride, n, cnt, d_tables, table_stride, base_offset, cfg, d_out_all + c0, af,
                                       alt_delta));
    }
    Jac* d_out = d_out_all;
    std::vector<de_g1> jac_tmp;
    if (out_mode == 1) jac_tmp.resize(count);
    DE_CUDA(ctx, cudaMemcpyAsync(out_mode == 1 ? (void*)jac_tmp.data() : host_out, d_out, sizeof(Jac) * count, cudaMemcpyDeviceToHost, st));
    unsigned int entries_h[2] = {0, 0};
    if (ctx->timing) DE_CUDA(ctx, cudaMemcpyAsync(entries_h, d_entries, sizeof(unsigned int) * (split ? 2 : 1), cudaMemcpyDeviceToHost, st));
    DE_CUDA(ctx, stream_wait(ctx, st));
    const unsigned int total_entries = entries_h[0] + entries_h[1];
    // transcript form: one batched inversion on the host for the round's handful of points
    if (out_mode == 1) host::g1_jacobian_to_canonical((const uint64_t*)jac_tmp.data(), count, (uint8_t*)host_out);
    if (ctx->timing) {
        // bucket additions actually performed (non-zero signed digits): the work figure behind the int-pipe fraction
        KernelStat* stat = nullptr;
        for (auto& k : ctx->stats)
            if (k.name == "msm_bucket_adds") stat = &k;
        if (!stat) {
            ctx->stats.push_back(KernelStat());
            stat = &ctx->stats.back();
            stat->name = "msm_bucket_adds";
        }
        stat->units += total_entries;
        stat->launches += split ? 2 : 1;  // one k_msm_accumulate launch per half
    }
    return DE_OK;
}

}  // namespace de

using namespace de;

struct de_params {
    de_ctx* ctx;
    uint32_t k;
    size_t n;
    MsmCfg cfg;
    Affine* tables[2];  // [0] g, [1] g_lagrange; each cfg.ntables tables of n points, carved out of ONE allocation
    Affine* block;
};

extern "C" {

int de_msm_dev(de_ctx* ctx, const de_fr* d_scalars, const de_g1_affine* d_bases, size_t n, de_g1* out) {
    if (!ctx) return DE_ERR_ARG;
    if (!out || (n && (!d_scalars || !d_bases))) return fail(ctx, DE_ERR_ARG, "de_msm_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    MsmCfg cfg = choose_cfg(n, false, 1);
    return msm_core(ctx, (const Fr*)d_scalars, n, n, 1, (const Affine*)d_bases, n, 0, cfg, out);
}

int de_msm(de_ctx* ctx, const de_fr* scalars, const de_g1_affine* bases, size_t n, de_g1* out) {
    if (!ctx) return DE_ERR_ARG;
    if (!out || (n && (!scalars || !bases))) return fail(ctx, DE_ERR_ARG, "de_msm: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) {
        memset(out, 0, sizeof(*out));
        return DE_OK;
    }
    DE_WS(ctx, ds, Fr, WS_IO_A, sizeof(Fr) * n);
    DE_WS(ctx, db, Affine, WS_IO_B, sizeof(Affine) * n);
    DE_CUDA(ctx, cudaMemcpyAsync(ds, scalars, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(db, bases, sizeof(Affine) * n, cudaMemcpyHostToDevice, ctx->stream));
    return de_msm_dev(ctx, (const de_fr*)ds, (const de_g1_affine*)db, n, out);
}

int de_params_upload(de_ctx* ctx, uint32_t k, const de_g1_affine* g, const de_g1_affine* g_lagrange, de_params** out) {
    if (!ctx) return DE_ERR_ARG;
    if (!out) return fail(ctx, DE_ERR_ARG, "de_params_upload: out is NULL");
    *out = nullptr;
    if (k > 26) return fail(ctx, DE_ERR_ARG, "de_params_upload: k > 26 not supported");
    if (!g && !g_lagrange) return fail(ctx, DE_ERR_ARG, "de_params_upload: both bases NULL");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << k;
    // table budget per basis: 4 GiB (override with DE_MSM_TABLE_MB)
    size_t budget = (size_t)4 << 30;
    if (const char* e = getenv("DE_MSM_TABLE_MB")) budget = (size_t)atoll(e) << 20;
    unsigned long long max_tables = budget / (sizeof(Affine) * n);
    if (max_tables < 1) max_tables = 1;
    de_params* p = new de_params();
    p->ctx = ctx;
    p->k = k;
    p->n = n;
    p->cfg = choose_cfg(n, true, max_tables);
    p->tables[0] = p->tables[1] = nullptr;
    p->block = nullptr;
    const de_g1_affine* src[2] = {g, g_lagrange};
    {
        const size_t per_basis = n * p->cfg.ntables;
        const size_t nb = (g ? 1 : 0) + (g_lagrange ? 1 : 0);
        if (cudaMalloc((void**)&p->block, sizeof(Affine) * per_basis * nb) != cudaSuccess) {
            cudaGetLastError();
            delete p;
            return fail(ctx, DE_ERR_OOM, "de_params_upload: device allocation failed");
        }
        size_t off = 0;
        for (int b = 0; b < 2; b++)
            if (src[b]) {
                p->tables[b] = p->block + off;
                off += per_basis;
            }
    }
    for (int b = 0; b < 2; b++) {
        if (!src[b]) continue;
        cudaError_t e = cudaSuccess;
        Affine* staging = (Affine*)ctx->ws[WS_IO_B].ensure(sizeof(Affine) * n);
        if (!staging) {
            cudaGetLastError();
            de_params_free(p);
            return fail(ctx, DE_ERR_OOM, "de_params_upload: device allocation failed");
        }
        e = cudaMemcpyAsync(staging, src[b], sizeof(Affine) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            k_msm_precompute<<<(unsigned int)((n + 127) / 128), 128, 0, ctx->stream>>>(staging, n, p->cfg.c * p->cfg.nsets, p->cfg.ntables, n,
                                                                                     p->tables[b]);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            de_params_free(p);
            return fail(ctx, DE_ERR_CUDA, std::string("de_params_upload: ") + cudaGetErrorString(e));
        }
    }
    *out = p;
    return DE_OK;
}

int de_params_free(de_params* p) {
    if (!p) return DE_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->block) cudaFree(p->block);
    delete p;
    return DE_OK;
}

int de_commit_batch_dev(de_params* p, int basis, const de_fr* d_scalars, size_t stride, size_t n, size_t count, de_g1* out) {
    if (!p) return DE_ERR_ARG;
    de_ctx* ctx = p->ctx;
    if (basis < 0 || basis > 1 || !p->tables[basis]) return fail(ctx, DE_ERR_ARG, "de_commit: basis not uploaded");
    if (n > p->n) return fail(ctx, DE_ERR_ARG, "de_commit: polynomial longer than the SRS (n > 2^k)");
    if (!out || (n && count && !d_scalars)) return fail(ctx, DE_ERR_ARG, "de_commit: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return msm_core(ctx, (const Fr*)d_scalars, stride, n, count, p->tables[basis], p->n, 0, p->cfg, out);
}

int de_commit_batch_canonical_dev(de_params* p, int basis, const de_fr* d_scalars, size_t stride, size_t n, size_t count, uint8_t* out_xy) {
    if (!p) return DE_ERR_ARG;
    de_ctx* ctx = p->ctx;
    if (basis < 0 || basis > 1 || !p->tables[basis]) return fail(ctx, DE_ERR_ARG, "de_commit: basis not uploaded");
    if (n > p->n) return fail(ctx, DE_ERR_ARG, "de_commit: polynomial longer than the SRS (n > 2^k)");
    if (!out_xy || (n && count && !d_scalars)) return fail(ctx, DE_ERR_ARG, "de_commit: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return msm_core(ctx, (const Fr*)d_scalars, stride, n, count, p->tables[basis], p->n, 0, p->cfg, out_xy, 1);
}

int de_commit_batch(de_params* p, int basis, const de_fr* const* scalars, size_t n, size_t count, de_g1* out) {
    if (!p) return DE_ERR_ARG;
    de_ctx* ctx = p->ctx;
    if (!out || (count && !scalars)) return fail(ctx, DE_ERR_ARG, "de_commit_batch: null pointer");
    if (n > p->n) return fail(ctx, DE_ERR_ARG, "de_commit: polynomial longer than the SRS (n > 2^k)");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return DE_OK;
    if (n == 0) {
        memset(out, 0, sizeof(de_g1) * count);
        return DE_OK;
    }
    DE_WS(ctx, ds, Fr, WS_IO_A, sizeof(Fr) * n * count);
    for (size_t i = 0; i < count; i++) {
        if (!scalars[i]) return fail(ctx, DE_ERR_ARG, "de_commit_batch: null polynomial");
        DE_CUDA(ctx, cudaMemcpyAsync(ds + i * n, scalars[i], sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    return de_commit_batch_dev(p, basis, (const de_fr*)ds, n, n, count, out);
}

int de_commit(de_params* p, int basis, const de_fr* scalars, size_t n, de_g1* out) {
    const de_fr* arr[1] = {scalars};
    return de_commit_batch(p, basis, arr, n, 1, out);
}

int de_commit_range(de_params* p, int basis, const de_fr* scalars, size_t lo, size_t hi, de_g1* out_partial) {
    if (!p) return DE_ERR_ARG;
    de_ctx* ctx = p->ctx;
    if (basis < 0 || basis > 1 || !p->tables[basis]) return fail(ctx, DE_ERR_ARG, "de_commit_range: basis not uploaded");
    if (lo > hi || hi > p->n) return fail(ctx, DE_ERR_ARG, "de_commit_range: bad range");
    if (!out_partial || (hi > lo && !scalars)) return fail(ctx, DE_ERR_ARG, "de_commit_range: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = hi - lo;
    if (n == 0) {
        memset(out_partial, 0, sizeof(de_g1));
        return DE_OK;
    }
    DE_WS(ctx, ds, Fr, WS_IO_A, sizeof(Fr) * n);
    DE_CUDA(ctx, cudaMemcpyAsync(ds, scalars + lo, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    return msm_core(ctx, ds, n, n, 1, p->tables[basis], p->n, lo, p->cfg, out_partial);
}

// Base-range sharding inside ONE process (the shape a Rust prover has: one process, several GPUs).  shards[i] is a
// ParamsKZG staged on its own context / GPU from the base slice [shard_lo[i], shard_lo[i] + shard_len[i]) of the SRS; one
// host thread per shard commits its slice of `scalars` concurrently, the partial points are summed on shards[0]'s GPU.
int de_commit_sharded(de_params* const* shards, const size_t* shard_lo, const size_t* shard_len, int n_shards, int basis, const de_fr* scalars,
                      de_g1* out) {
    if (!shards || n_shards < 1 || !shards[0]) return DE_ERR_ARG;
    de_ctx* ctx0 = shards[0]->ctx;
    if (!shard_lo || !shard_len || !scalars || !out) return fail(ctx0, DE_ERR_ARG, "de_commit_sharded: null pointer");
    if (n_shards > 64) return fail(ctx0, DE_ERR_ARG, "de_commit_sharded: more than 64 shards");
    for (int i = 0; i < n_shards; i++) {
        if (!shards[i]) return fail(ctx0, DE_ERR_ARG, "de_commit_sharded: null shard");
        if (shard_len[i] > shards[i]->n) return fail(ctx0, DE_ERR_ARG, "de_commit_sharded: slice longer than the shard's bases");
        for (int j = 0; j < i; j++)
            if (shards[j]->ctx == shards[i]->ctx) return fail(ctx0, DE_ERR_ARG, "de_commit_sharded: shards must use distinct contexts");
    }
    std::vector<de_g1> partial(n_shards);
    std::vector<int> rc(n_shards, DE_OK);
    std::vector<std::thread> threads;
    for (int i = 0; i < n_shards; i++)
        threads.emplace_back([&, i]() { rc[i] = de_commit(shards[i], basis, scalars + shard_lo[i], shard_len[i], &partial[i]); });
    for (auto& t : threads) t.join();
    for (int i = 0; i < n_shards; i++)
        if (rc[i] != DE_OK) return fail(ctx0, rc[i], std::string("de_commit_sharded: shard ") + std::to_string(i) + ": " + de_last_error(shards[i]->ctx));
    return de_g1_sum(ctx0, partial.data(), (size_t)n_shards, out);
}

int de_g1_sum(de_ctx* ctx, const de_g1* points, size_t count, de_g1* out) {
    if (!ctx) return DE_ERR_ARG;
    if (!out || (count && !points)) return fail(ctx, DE_ERR_ARG, "de_g1_sum: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) {
        memset(out, 0, sizeof(*out));
        return DE_OK;
    }
    DE_WS(ctx, d, Jac, WS_MSM_OUT, sizeof(Jac) * (count + 1));
    DE_CUDA(ctx, cudaMemcpyAsync(d + 1, points, sizeof(Jac) * count, cudaMemcpyHostToDevice, ctx->stream));
    k_g1_sum<<<1, 32, 0, ctx->stream>>>(d + 1, (unsigned int)count, d);
    DE_CHECK_LAUNCH(ctx);
    DE_CUDA(ctx, cudaMemcpyAsync(out, d, sizeof(Jac), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

int de_g1_mul_base_dev(de_ctx* ctx, const de_g1_affine* base, const de_fr* d_scalars, size_t n, de_g1_affine* d_out) {
    if (!ctx) return DE_ERR_ARG;
    if (!base || (n && (!d_scalars || !d_out))) return fail(ctx, DE_ERR_ARG, "de_g1_mul_base_dev: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return DE_OK;
    Affine b;
    memcpy(&b, base, sizeof(b));
    k_g1_mul_base<<<(unsigned int)((n + 127) / 128), 128, 0, ctx->stream>>>(b, (const Fr*)d_scalars, n, (Affine*)d_out);
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

int de_g1_batch_normalize(de_ctx* ctx, const de_g1* points, size_t count, de_g1_affine* out) {
    if (!ctx) return DE_ERR_ARG;
    if (count && (!points || !out)) return fail(ctx, DE_ERR_ARG, "de_g1_batch_normalize: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return DE_OK;
    DE_WS(ctx, d, Jac, WS_MSM_OUT, sizeof(Jac) * count + sizeof(Affine) * count);
    Affine* da = (Affine*)(d + count);
    DE_CUDA(ctx, cudaMemcpyAsync(d, points, sizeof(Jac) * count, cudaMemcpyHostToDevice, ctx->stream));
    k_g1_normalize<<<(unsigned int)((count + 63) / 64), 64, 0, ctx->stream>>>(d, (unsigned int)count, da);
    DE_CHECK_LAUNCH(ctx);
    DE_CUDA(ctx, cudaMemcpyAsync(out, da, sizeof(Affine) * count, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

}  // extern "C"

// internal entry points for prover.cu
namespace de {
int commit_canonical_dev(de_params* p, int basis, const Fr* d_scalars, size_t stride, size_t n, size_t count, uint8_t* out_xy) {
    return de_commit_batch_canonical_dev(p, basis, (const de_fr*)d_scalars, stride, n, count, out_xy);
}
// `count_lagrange` polynomials committed over g_lagrange followed by `count_coeff` over g, one launch sequence
int commit_canonical_mixed_dev(de_params* p, const Fr* d_scalars, size_t stride, size_t n, size_t count_lagrange, size_t count_coeff,
                               uint8_t* out_xy) {
    de_ctx* ctx = p->ctx;
    if (!p->tables[0] || !p->tables[1]) return fail(ctx, DE_ERR_ARG, "de_commit: both bases are needed");
    if (n > p->n) return fail(ctx, DE_ERR_ARG, "de_commit: polynomial longer than the SRS (n > 2^k)");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    // indices are offsets from the start of the shared allocation, so that both bases are reachable with non-negative indices
    return msm_core(ctx, d_scalars, stride, n, count_lagrange + count_coeff, p->block, p->n, (size_t)(p->tables[1] - p->block), p->cfg, out_xy, 1,
                    (unsigned int)count_lagrange, (long long)(p->tables[0] - p->tables[1]), true);
}
de_ctx* params_ctx(de_params* p) { return p->ctx; }
size_t params_n(de_params* p) { return p->n; }
}  // namespace de
