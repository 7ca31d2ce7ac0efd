// eval.cuh — pieces of the quotient evaluator shared with the prover (prover.cu): the device image of a compiled
// GraphEvaluator, its interpreter, and the ProvingKey object (see eval.cu for the reference mapping).
#pragma once
#include <string.h>

#include <vector>

#include "common.cuh"

struct de_domain_view {  // layout prefix of de_domain (ntt.cu) that this unit reads
    de_ctx* ctx;
    uint32_t j, k, ek;
    size_t n, ext_n, qdeg;
    de_fr omega, omega_inv, ext_omega, ext_omega_inv;
};

namespace de {

#define DE_CALC_HORNER_FROM_ZERO 8  // internal: DE_CALC_HORNER with a zero start value (set by upload_graph)
#define DE_MAX_INTER 24   // live intermediates per row AFTER slot allocation (upload_graph); local memory per thread
#define DE_MAX_ROT 16

struct DevSrc { uint32_t kind, index, rot; };
struct DevCalc { uint32_t op; DevSrc a, b; uint32_t hfirst, hlen, target; };
struct DevGraph {
    const Fr* constants;
    const int* rotations;
    const DevCalc* calcs;
    const DevSrc* hparts;
    uint32_t n_rot, n_calcs;
};

struct EvalParams {
    uint32_t ext_mask, rot_scale;
    unsigned long long ext_n;
    // resident pk cosets (column-major, ext_n apart)
    const Fr* fixed; const Fr* sigma; const Fr* l0; const Fr* l_last; const Fr* l_active; const Fr* omega_pows;
    // per-proof cosets
    const Fr* advice; const Fr* instance; const Fr* permz; const Fr* lookup_z; const Fr* lookup_a; const Fr* lookup_s;
    const Fr* challenges;
    const Fr* ypows;   // y^0 .. y^n_terms (device); n_terms = constraints folded after the custom gates
    uint32_t n_terms;
    Fr y, beta, gamma, theta, delta, delta_start;  // delta_start holds ZETA (Montgomery)
    DevGraph gates;
    const DevGraph* lookups;
    uint32_t n_lookups;
    // permutation
    uint32_t n_perm_cols, chunk_len, n_sets;
    int last_rotation;
    const uint32_t* perm_kind; const uint32_t* perm_index;
    Fr* out;
};

struct RowCtx {
    const EvalParams* p;
    unsigned int idx;
    unsigned int rot_idx[DE_MAX_ROT];
    Fr inter[DE_MAX_INTER];
    Fr prev;
};

__device__ __forceinline__ unsigned int rot_index(unsigned int idx, int rot, unsigned int rot_scale, unsigned int mask) {
    return (unsigned int)((int)idx + rot * (int)rot_scale) & mask;
}

__device__ __forceinline__ Fr fetch(const RowCtx& c, const DevGraph& g, const DevSrc& s) {
    const EvalParams& p = *c.p;
    switch (s.kind) {
        case DE_VAL_CONSTANT: return load(&g.constants[s.index]);
        case DE_VAL_INTERMEDIATE: return c.inter[s.index];
        case DE_VAL_FIXED: return load(&p.fixed[(unsigned long long)s.index * p.ext_n + c.rot_idx[s.rot]]);
        case DE_VAL_ADVICE: return load(&p.advice[(unsigned long long)s.index * p.ext_n + c.rot_idx[s.rot]]);
        case DE_VAL_INSTANCE: return load(&p.instance[(unsigned long long)s.index * p.ext_n + c.rot_idx[s.rot]]);
        case DE_VAL_CHALLENGE: return load(&p.challenges[s.index]);
        case DE_VAL_BETA: return p.beta;
        case DE_VAL_GAMMA: return p.gamma;
        case DE_VAL_THETA: return p.theta;
        case DE_VAL_Y: return p.y;
        default: return c.prev;  // DE_VAL_PREVIOUS
    }
}

__device__ inline Fr run_graph(RowCtx& c, const DevGraph& g) {
    const EvalParams& p = *c.p;
    for (uint32_t r = 0; r < g.n_rot; r++) c.rot_idx[r] = rot_index(c.idx, g.rotations[r], p.rot_scale, p.ext_mask);
    Fr last = Fr::zero();
    for (uint32_t i = 0; i < g.n_calcs; i++) {
        const DevCalc cc = g.calcs[i];
        Fr a = fetch(c, g, cc.a);
        Fr r;
        switch (cc.op) {
            case DE_CALC_ADD: r = add(a, fetch(c, g, cc.b)); break;
            case DE_CALC_SUB: r = sub(a, fetch(c, g, cc.b)); break;
            case DE_CALC_MUL: r = mul(a, fetch(c, g, cc.b)); break;
            case DE_CALC_SQUARE: r = sqr(a); break;
            case DE_CALC_DOUBLE: r = dbl(a); break;
            case DE_CALC_NEGATE: r = neg(a); break;
            case DE_CALC_HORNER: {
                Fr f = fetch(c, g, cc.b);
                r = a;
                for (uint32_t k = 0; k < cc.hlen; k++) r = add(mul(r, f), fetch(c, g, g.hparts[cc.hfirst + k]));
                break;
            }
            case DE_CALC_HORNER_FROM_ZERO: {  // Horner whose start value is zero: 0 * f + part_0 = part_0, one multiplication less
                Fr f = fetch(c, g, cc.b);
                r = fetch(c, g, g.hparts[cc.hfirst]);
                for (uint32_t k = 1; k < cc.hlen; k++) r = add(mul(r, f), fetch(c, g, g.hparts[cc.hfirst + k]));
                break;
            }
            default: r = a; break;  // DE_CALC_STORE
        }
        c.inter[cc.target] = r;
        last = r;
    }
    return last;
}

__device__ __forceinline__ const Fr* perm_column(const EvalParams& p, uint32_t col) {
    const uint32_t kind = p.perm_kind[col], index = p.perm_index[col];
    const Fr* base = kind == DE_VAL_ADVICE ? p.advice : (kind == DE_VAL_FIXED ? p.fixed : p.instance);
    return base + (unsigned long long)index * p.ext_n;
}

}  // namespace de

using namespace de;  // internal header of libde_b200.so

struct de_pk {
    de_domain* dom;
    de_ctx* ctx;
    uint32_t n_fixed, n_advice, n_instance, n_perm_cols, chunk_len, n_sets, n_lookups, blinding;
    size_t n, ext_n;
    uint32_t k, ek;
    Fr* resident;  // fixed | sigma | l0 | l_last | l_active | omega_pows, ext_n apart
    Fr* coeff;     // fixed | sigma polynomials in coefficient form, n apart (the prover opens and evaluates them)
    Fr* work;      // per-proof cosets, ext_n apart, in the prover's column order: advice | instance | a' | s' | permz | lookup z
    Fr* d_challenges;
    uint32_t challenges_cap;
    Fr* d_ypows;
    uint32_t ypows_cap;
    uint32_t *d_perm_kind, *d_perm_index;
    DevGraph gates;
    DevGraph* d_lookups;
    std::vector<void*> allocs;
    Fr delta, zeta;
};

inline int upload_graph(de_ctx* ctx, const de_graph& g, DevGraph* out, std::vector<void*>& allocs) {
    if (g.n_rotations > DE_MAX_ROT) return fail(ctx, DE_ERR_UNSUPPORTED, "evaluator: too many distinct rotations in a graph");
    std::vector<DevCalc> calcs(g.n_calcs);
    auto conv = [](const de_value_source& s) { return DevSrc{s.kind, s.index, s.rotation}; };
    for (uint32_t i = 0; i < g.n_calcs; i++) {
        const de_calculation& c = g.calcs[i];
        if (c.target >= g.n_intermediates) return fail(ctx, DE_ERR_ARG, "evaluator: calculation target out of range");
        if (c.op > DE_CALC_STORE) return fail(ctx, DE_ERR_ARG, "evaluator: unknown calculation");
        if (c.op == DE_CALC_HORNER && (uint64_t)c.horner_first + c.horner_len > g.n_horner_parts)
            return fail(ctx, DE_ERR_ARG, "evaluator: horner parts out of range");
        calcs[i] = DevCalc{c.op, conv(c.a), conv(c.b), c.horner_first, c.horner_len, c.target};
    }
    std::vector<DevSrc> parts(g.n_horner_parts);
    for (uint32_t i = 0; i < g.n_horner_parts; i++) parts[i] = conv(g.horner_parts[i]);
    // ---- device-side program optimisation (results unchanged: same operations on the same values, in the same order)
    // 1. Store(x) only copies a column / challenge value into an intermediate: forward x into the consumers instead, so that
    //    the value is loaded where it is used and never parked in thread-local memory.
    {
        std::vector<int> store_of(g.n_intermediates, -1);
        for (uint32_t i = 0; i < calcs.size(); i++)
            if (calcs[i].op == DE_CALC_STORE && i + 1 != calcs.size()) store_of[calcs[i].target] = (int)i;
        auto fwd = [&](DevSrc& s) {
            int guard = 0;
            while (s.kind == DE_VAL_INTERMEDIATE && s.index < store_of.size() && store_of[s.index] >= 0 && guard++ < 64) s = calcs[store_of[s.index]].a;
        };
        // a Store's own source may be an earlier Store's target: resolve in program order first
        for (auto& c : calcs)
            if (c.op == DE_CALC_STORE) fwd(c.a);
        for (auto& c : calcs) {
            if (c.op == DE_CALC_STORE) continue;
            fwd(c.a);
            fwd(c.b);
        }
        for (auto& s : parts) fwd(s);
        std::vector<DevCalc> kept;
        for (uint32_t i = 0; i < calcs.size(); i++)
            if (!(calcs[i].op == DE_CALC_STORE && i + 1 != calcs.size())) kept.push_back(calcs[i]);
        calcs.swap(kept);
    }
    // 1b. Horner(start, f, parts) with start = the constant 0 (the theta-compressions of the lookup arguments) or PreviousValue
    //     (zero in every kernel that runs these programs: the gates' fold starts from nothing): 0 * f + part_0 is part_0
    {
        auto is_zero_const = [&](const DevSrc& s) {
            if (s.kind == DE_VAL_PREVIOUS) return true;
            if (s.kind != DE_VAL_CONSTANT || s.index >= g.n_constants) return false;
            const de_fr& v = g.constants[s.index];
            return (v.l[0] | v.l[1] | v.l[2] | v.l[3]) == 0;
        };
        for (auto& c : calcs)
            if (c.op == DE_CALC_HORNER && c.hlen >= 1 && is_zero_const(c.a)) c.op = DE_CALC_HORNER_FROM_ZERO;
    }
    // 2. slot allocation by liveness: an intermediate's slot is reused after its last consumer
    {
        const uint32_t nc = (uint32_t)calcs.size();
        std::vector<int> last_use(g.n_intermediates, -1);
        auto use = [&](const DevSrc& s, int at) {
            if (s.kind == DE_VAL_INTERMEDIATE && s.index < last_use.size()) last_use[s.index] = at;
        };
        for (uint32_t i = 0; i < nc; i++) {
            use(calcs[i].a, (int)i);
            use(calcs[i].b, (int)i);
            if (calcs[i].op == DE_CALC_HORNER || calcs[i].op == DE_CALC_HORNER_FROM_ZERO)
                for (uint32_t k = 0; k < calcs[i].hlen; k++) use(parts[calcs[i].hfirst + k], (int)i);
        }
        std::vector<int> slot_of(g.n_intermediates, -1);
        std::vector<int> free_at(DE_MAX_INTER, -1);  // calc index after which the slot is free again; -1 = free now
        auto remap = [&](DevSrc& s) {
            if (s.kind == DE_VAL_INTERMEDIATE && s.index < slot_of.size() && slot_of[s.index] >= 0) s.index = (uint32_t)slot_of[s.index];
        };
        std::vector<char> part_done(parts.size(), 0);
        for (uint32_t i = 0; i < nc; i++) {
            remap(calcs[i].a);
            remap(calcs[i].b);
            if (calcs[i].op == DE_CALC_HORNER || calcs[i].op == DE_CALC_HORNER_FROM_ZERO)
                for (uint32_t k = 0; k < calcs[i].hlen; k++)
                    if (!part_done[calcs[i].hfirst + k]) {
                        remap(parts[calcs[i].hfirst + k]);
                        part_done[calcs[i].hfirst + k] = 1;
                    }
            // operands were read before the target is written, so a slot whose last use is this very calculation may be reused
            int slot = -1;
            for (int s = 0; s < DE_MAX_INTER; s++)
                if (free_at[s] <= (int)i) {
                    slot = s;
                    break;
                }
            if (slot < 0) return fail(ctx, DE_ERR_UNSUPPORTED, "evaluator: more live intermediates than DE_MAX_INTER");
            const uint32_t t = calcs[i].target;
            slot_of[t] = slot;
            free_at[slot] = last_use[t] < 0 ? (int)i : last_use[t];  // never used again: free right after
            calcs[i].target = (uint32_t)slot;
        }
    }
    auto up = [&](const void* src, size_t bytes, const void** dst) -> int {
        void* d = nullptr;
        DE_CUDA(ctx, cudaMalloc(&d, bytes ? bytes : 16));
        allocs.push_back(d);
        if (bytes) DE_CUDA(ctx, cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        *dst = d;
        return DE_OK;
    };
    DE_TRY(up(g.constants, sizeof(Fr) * g.n_constants, (const void**)&out->constants));
    DE_TRY(up(g.rotations, sizeof(int) * g.n_rotations, (const void**)&out->rotations));
    DE_TRY(up(calcs.data(), sizeof(DevCalc) * calcs.size(), (const void**)&out->calcs));
    DE_TRY(up(parts.data(), sizeof(DevSrc) * parts.size(), (const void**)&out->hparts));
    out->n_rot = g.n_rotations;
    out->n_calcs = (uint32_t)calcs.size();
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    return DE_OK;
}

