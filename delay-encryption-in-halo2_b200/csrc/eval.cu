// eval.cu — quotient-polynomial evaluator on the extended domain (replaces
// halo2_proofs::plonk::evaluation::Evaluator::evaluate_h; reached from the
// reference through create_proof at benches/delay_enc.rs:123, benches/mod_pow.rs:201, benches/pose_enc.rs:127).
//
// ProvingKey staging (de_pk_upload): fixed, sigma, l0, l_last, l_active_row cosets and the coset points are computed on
// the device once and stay resident.  Per proof: ONE batched coeff_to_extended for advice / instance / z / lookup
// polynomials, then ONE fused kernel: each thread owns an extended-domain row, keeps `value` in registers and runs
//   custom gates (compiled GraphEvaluator program, interpreted; intermediates in thread-local memory)
//   -> permutation argument constraints -> every lookup's five constraints
// folding with y exactly in the reference's order, so the output equals evaluate_h's bit for bit.
#include "eval.cuh"
#include "transcript.hpp"

namespace de {

// The reference folds every constraint into one value by Horner's rule in y (value = value * y + constraint), each constraint
// already multiplied by its l0 / l_last / l_active factor: two multiplications per constraint.  The field is exact, so any
// arrangement of the same polynomial identity gives the same element: here constraint j (of T, in the reference's order) is
// multiplied by the precomputed power y^(T-1-j) and added to the running sum of ITS indicator column, and the three sums meet
// their indicator once at the end:
//   value = gates * y^T + l0 * S0 + l_last * SL + l_active * SA          (one multiplication per constraint + 4)
__global__ void __launch_bounds__(128) k_eval_h(const __grid_constant__ EvalParams p) {
    const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.ext_n) return;
    RowCtx c;
    c.p = &p;
    c.idx = idx;
    const Fr one = Fr::one();
    // ---- custom gates: Horner(previous = 0, gate polynomials, y) inside the compiled program
    c.prev = Fr::zero();
    const Fr gates = run_graph(c, p.gates);
    const unsigned int r_next = rot_index(idx, 1, p.rot_scale, p.ext_mask);
    const unsigned int r_prev = rot_index(idx, -1, p.rot_scale, p.ext_mask);
    const Fr* yp = p.ypows + p.n_terms;  // yp[-1 - j] = y^(T-1-j): the power constraint j carries
    Fr s0 = Fr::zero(), sl = Fr::zero(), sa = Fr::zero();
    int j = 0;
#define DE_TERM(acc, expr)                               \
    do {                                                 \
        acc = add(acc, mul((expr), load(&yp[-1 - j])));  \
        j++;                                             \
    } while (0)
    // ---- permutation argument
    if (p.n_sets > 0) {
        const unsigned int r_last = rot_index(idx, p.last_rotation, p.rot_scale, p.ext_mask);
        const Fr z_first = load(&p.permz[idx]);
        DE_TERM(s0, sub(one, z_first));
        const Fr z_lastset = load(&p.permz[(unsigned long long)(p.n_sets - 1) * p.ext_n + idx]);
        DE_TERM(sl, sub(sqr(z_lastset), z_lastset));
        for (uint32_t s = 1; s < p.n_sets; s++) {
            Fr zi = load(&p.permz[(unsigned long long)s * p.ext_n + idx]);
            Fr zp = load(&p.permz[(unsigned long long)(s - 1) * p.ext_n + r_last]);
            DE_TERM(s0, sub(zi, zp));
        }
        Fr current_delta = mul(mul(p.beta, p.delta_start), load(&p.omega_pows[idx]));  // beta * ZETA * extended_omega^idx
        for (uint32_t s = 0; s < p.n_sets; s++) {
            const uint32_t c0 = s * p.chunk_len;
            const uint32_t c1 = (c0 + p.chunk_len < p.n_perm_cols) ? c0 + p.chunk_len : p.n_perm_cols;
            Fr left = load(&p.permz[(unsigned long long)s * p.ext_n + r_next]);
            Fr right = load(&p.permz[(unsigned long long)s * p.ext_n + idx]);
            for (uint32_t col = c0; col < c1; col++) {
                Fr v = load(&perm_column(p, col)[idx]);
                Fr sg = load(&p.sigma[(unsigned long long)col * p.ext_n + idx]);
                left = mul(left, add(add(v, mul(p.beta, sg)), p.gamma));
                right = mul(right, add(add(v, current_delta), p.gamma));
                current_delta = mul(current_delta, p.delta);
            }
            DE_TERM(sa, sub(left, right));
        }
    }
    // ---- lookups
    for (uint32_t n = 0; n < p.n_lookups; n++) {
        c.prev = Fr::zero();
        const Fr table_value = run_graph(c, p.lookups[n]);
        const Fr* zc = p.lookup_z + (unsigned long long)n * p.ext_n;
        const Fr* ac = p.lookup_a + (unsigned long long)n * p.ext_n;
        const Fr* sc = p.lookup_s + (unsigned long long)n * p.ext_n;
        const Fr z = load(&zc[idx]), a = load(&ac[idx]), s = load(&sc[idx]);
        const Fr a_minus_s = sub(a, s);
        DE_TERM(s0, sub(one, z));
        DE_TERM(sl, sub(sqr(z), z));
        DE_TERM(sa, sub(mul(mul(load(&zc[r_next]), add(a, p.beta)), add(s, p.gamma)), mul(z, table_value)));
        DE_TERM(s0, a_minus_s);
        DE_TERM(sa, mul(a_minus_s, sub(a, load(&ac[r_prev]))));
    }
#undef DE_TERM
    Fr value = mul(gates, load(&p.ypows[p.n_terms]));
    value = add(value, mul(s0, load(&p.l0[idx])));
    value = add(value, mul(sl, load(&p.l_last[idx])));
    value = add(value, mul(sa, load(&p.l_active[idx])));
    store(&p.out[idx], value);
}

// lagrange indicator columns for l0 / l_last / l_blind (written in lagrange form, then iNTT + coset NTT)
__global__ void k_fill_indicators(Fr* l0, Fr* l_last, Fr* l_blind, unsigned long long n, unsigned int blinding) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr one = Fr::one(), zero = Fr::zero();
    store(&l0[i], i == 0 ? one : zero);
    store(&l_last[i], i == n - blinding - 1 ? one : zero);
    store(&l_blind[i], i >= n - blinding ? one : zero);
}
// l_active = 1 - (l_last + l_blind) on the extended domain
__global__ void k_l_active(const Fr* l_last, const Fr* l_blind, Fr* out, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store(&out[i], sub(Fr::one(), add(load(&l_last[i]), load(&l_blind[i]))));
}
__global__ void k_pow_seq(Fr* out, unsigned long long n, Fr base) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long e = i;
    Fr acc = Fr::one(), cur = base;
    while (e) {
        if (e & 1) acc = mul(acc, cur);
        cur = sqr(cur);
        e >>= 1;
    }
    store(&out[i], acc);
}

struct HostGraph {
    DevGraph dev;
    std::vector<void*> allocs;
};

}  // namespace de

using namespace de;

extern "C" {

int de_pk_free(de_pk* pk) {
    if (!pk) return DE_OK;
    cudaSetDevice(pk->ctx->device);
    cudaStreamSynchronize(pk->ctx->stream);
    for (void* a : pk->allocs) cudaFree(a);
    delete pk;
    return DE_OK;
}

int de_pk_upload(de_domain* dom, const de_pk_desc* desc, de_pk** out) {
    if (!dom) return DE_ERR_ARG;
    const de_domain_view* dv = (const de_domain_view*)dom;
    de_ctx* ctx = dv->ctx;
    if (!desc || !out) return fail(ctx, DE_ERR_ARG, "de_pk_upload: null pointer");
    *out = nullptr;
    if (desc->chunk_len == 0 && desc->n_perm_columns) return fail(ctx, DE_ERR_ARG, "de_pk_upload: chunk_len is 0");
    if (desc->blinding_factors + 1 >= dv->n) return fail(ctx, DE_ERR_ARG, "de_pk_upload: blinding_factors too large for n");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    de_pk* pk = new de_pk();
    pk->dom = dom; pk->ctx = ctx;
    pk->n_fixed = desc->n_fixed; pk->n_advice = desc->n_advice; pk->n_instance = desc->n_instance;
    pk->n_perm_cols = desc->n_perm_columns; pk->chunk_len = desc->chunk_len; pk->n_lookups = desc->n_lookups;
    pk->n_sets = desc->n_perm_columns ? (desc->n_perm_columns + desc->chunk_len - 1) / desc->chunk_len : 0;
    pk->blinding = desc->blinding_factors;
    pk->n = dv->n; pk->ext_n = dv->ext_n; pk->k = dv->k; pk->ek = dv->ek;
    pk->delta = fr_from_host(desc->delta);
    pk->d_challenges = nullptr; pk->challenges_cap = 0;
    pk->d_ypows = nullptr; pk->ypows_cap = 0;
    const size_t n = dv->n, ext_n = dv->ext_n;
    const size_t n_res = (size_t)desc->n_fixed + desc->n_perm_columns + 4;  // + l0, l_last, l_active, omega_pows
    const size_t n_work = (size_t)desc->n_advice + desc->n_instance + pk->n_sets + 3 * (size_t)desc->n_lookups;
    auto bail = [&](int rc) { de_pk_free(pk); return rc; };
    auto dmalloc = [&](void** p, size_t bytes) -> int {
        DE_CUDA(ctx, cudaMalloc(p, bytes ? bytes : 16));
        pk->allocs.push_back(*p);
        return DE_OK;
    };
    int rc;
    if ((rc = dmalloc((void**)&pk->resident, sizeof(Fr) * ext_n * n_res)) != DE_OK) return bail(rc);
    if ((rc = dmalloc((void**)&pk->work, sizeof(Fr) * ext_n * (n_work ? n_work : 1))) != DE_OK) return bail(rc);
    // stage coefficient-form fixed + sigma polynomials, then one batched coeff_to_extended into the resident block
    const size_t n_polys = (size_t)desc->n_fixed + desc->n_perm_columns;
    Fr* staging = (Fr*)ctx->ws[WS_EVAL_A].ensure(sizeof(Fr) * n * (n_polys + 3));
    if (!staging) return bail(fail(ctx, DE_ERR_OOM, "de_pk_upload: staging allocation failed"));
    for (size_t i = 0; i < n_polys; i++) {
        const de_fr* src = i < desc->n_fixed ? desc->fixed_coeff[i] : desc->sigma_coeff[i - desc->n_fixed];
        if (!src) return bail(fail(ctx, DE_ERR_ARG, "de_pk_upload: null polynomial"));
        cudaError_t e = cudaMemcpyAsync(staging + i * n, src, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return bail(fail(ctx, DE_ERR_CUDA, cudaGetErrorString(e)));
    }
    if ((rc = dmalloc((void**)&pk->coeff, sizeof(Fr) * n * (n_polys ? n_polys : 1))) != DE_OK) return bail(rc);
    if (n_polys) {
        cudaError_t e = cudaMemcpyAsync(pk->coeff, staging, sizeof(Fr) * n * n_polys, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) return bail(fail(ctx, DE_ERR_CUDA, cudaGetErrorString(e)));
    }
    Fr* ind = staging + n_polys * n;  // l0, l_last, l_blind in lagrange form
    k_fill_indicators<<<(unsigned int)((n + 255) / 256), 256, 0, ctx->stream>>>(ind, ind + n, ind + 2 * n, n, desc->blinding_factors);
    ctx->launches++;
    if ((rc = de_lagrange_to_coeff_dev(dom, (de_fr*)ind, n, 3)) != DE_OK) return bail(rc);
    Fr* res_fixed = pk->resident;
    Fr* res_l0 = pk->resident + n_polys * ext_n;
    Fr* res_l_last = res_l0 + ext_n;
    Fr* res_l_active = res_l_last + ext_n;
    Fr* res_omega = res_l_active + ext_n;
    if (n_polys && (rc = de_coeff_to_extended_dev(dom, (const de_fr*)staging, n, (de_fr*)res_fixed, ext_n, n_polys)) != DE_OK) return bail(rc);
    // l0, l_last -> their slots; l_blind -> the omega slot temporarily
    if ((rc = de_coeff_to_extended_dev(dom, (const de_fr*)ind, n, (de_fr*)res_l0, ext_n, 2)) != DE_OK) return bail(rc);
    if ((rc = de_coeff_to_extended_dev(dom, (const de_fr*)(ind + 2 * n), n, (de_fr*)res_omega, ext_n, 1)) != DE_OK) return bail(rc);
    k_l_active<<<(unsigned int)((ext_n + 255) / 256), 256, 0, ctx->stream>>>(res_l_last, res_omega, res_l_active, ext_n);
    ctx->launches++;
    k_pow_seq<<<(unsigned int)((ext_n + 255) / 256), 256, 0, ctx->stream>>>(res_omega, ext_n, fr_from_host(dv->ext_omega));
    ctx->launches++;
    // permutation column descriptors
    if (desc->n_perm_columns) {
        if ((rc = dmalloc((void**)&pk->d_perm_kind, sizeof(uint32_t) * desc->n_perm_columns)) != DE_OK) return bail(rc);
        if ((rc = dmalloc((void**)&pk->d_perm_index, sizeof(uint32_t) * desc->n_perm_columns)) != DE_OK) return bail(rc);
        for (uint32_t i = 0; i < desc->n_perm_columns; i++) {
            uint32_t kind = desc->perm_column_kind[i], index = desc->perm_column_index[i];
            uint32_t lim = kind == DE_VAL_ADVICE ? desc->n_advice : kind == DE_VAL_FIXED ? desc->n_fixed : kind == DE_VAL_INSTANCE ? desc->n_instance : 0;
            if (index >= lim) return bail(fail(ctx, DE_ERR_ARG, "de_pk_upload: permutation column out of range"));
        }
        cudaMemcpyAsync(pk->d_perm_kind, desc->perm_column_kind, sizeof(uint32_t) * desc->n_perm_columns, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(pk->d_perm_index, desc->perm_column_index, sizeof(uint32_t) * desc->n_perm_columns, cudaMemcpyHostToDevice, ctx->stream);
    } else {
        pk->d_perm_kind = pk->d_perm_index = nullptr;
    }
    if ((rc = upload_graph(ctx, desc->gates, &pk->gates, pk->allocs)) != DE_OK) return bail(rc);
    std::vector<DevGraph> lg(desc->n_lookups);
    for (uint32_t i = 0; i < desc->n_lookups; i++)
        if ((rc = upload_graph(ctx, desc->lookups[i], &lg[i], pk->allocs)) != DE_OK) return bail(rc);
    if ((rc = dmalloc((void**)&pk->d_lookups, sizeof(DevGraph) * (lg.size() ? lg.size() : 1))) != DE_OK) return bail(rc);
    if (!lg.empty()) cudaMemcpyAsync(pk->d_lookups, lg.data(), sizeof(DevGraph) * lg.size(), cudaMemcpyHostToDevice, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bail(fail(ctx, DE_ERR_CUDA, std::string("de_pk_upload: ") + cudaGetErrorString(e)));
    *out = pk;
    return DE_OK;
}

int de_pk_extend_dev(de_pk* pk, const de_fr* d_advice, const de_fr* d_instance, const de_fr* d_permz, const de_fr* d_lookup,
                     size_t stride) {
    if (!pk) return DE_ERR_ARG;
    de_ctx* ctx = pk->ctx;
    if ((pk->n_advice && !d_advice) || (pk->n_instance && !d_instance) || (pk->n_sets && !d_permz) || (pk->n_lookups && !d_lookup))
        return fail(ctx, DE_ERR_ARG, "de_pk_extend_dev: missing polynomial block");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ext_n = pk->ext_n;
    const size_t L = pk->n_lookups;
    Fr* w_adv = pk->work;
    Fr* w_inst = w_adv + (size_t)pk->n_advice * ext_n;
    Fr* w_la = w_inst + (size_t)pk->n_instance * ext_n;  // a' block, then s' block
    Fr* w_permz = w_la + 2 * L * ext_n;
    Fr* w_lz = w_permz + (size_t)pk->n_sets * ext_n;
    if (pk->n_advice) DE_TRY(de_coeff_to_extended_dev(pk->dom, d_advice, stride, (de_fr*)w_adv, ext_n, pk->n_advice));
    if (pk->n_instance) DE_TRY(de_coeff_to_extended_dev(pk->dom, d_instance, stride, (de_fr*)w_inst, ext_n, pk->n_instance));
    if (pk->n_sets) DE_TRY(de_coeff_to_extended_dev(pk->dom, d_permz, stride, (de_fr*)w_permz, ext_n, pk->n_sets));
    if (L) {
        // the API's lookup block is [all z | all a' | all s']
        DE_TRY(de_coeff_to_extended_dev(pk->dom, d_lookup, stride, (de_fr*)w_lz, ext_n, L));
        DE_TRY(de_coeff_to_extended_dev(pk->dom, d_lookup + L * stride, stride, (de_fr*)w_la, ext_n, 2 * L));
    }
    return DE_OK;
}

int de_evaluate_h_rows_dev(de_pk* pk, const de_challenges* ch, de_fr* d_h_ext) {
    if (!pk) return DE_ERR_ARG;
    de_ctx* ctx = pk->ctx;
    if (!ch || !d_h_ext) return fail(ctx, DE_ERR_ARG, "de_evaluate_h: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t ext_n = pk->ext_n;
    Fr* w_adv = pk->work;
    Fr* w_inst = w_adv + (size_t)pk->n_advice * ext_n;
    Fr* w_la = w_inst + (size_t)pk->n_instance * ext_n;
    Fr* w_ls = w_la + (size_t)pk->n_lookups * ext_n;
    Fr* w_permz = w_ls + (size_t)pk->n_lookups * ext_n;
    Fr* w_lz = w_permz + (size_t)pk->n_sets * ext_n;
    if (ch->n_challenges > pk->challenges_cap) {
        Fr* d = nullptr;
        DE_CUDA(ctx, cudaMalloc((void**)&d, sizeof(Fr) * ch->n_challenges));
        pk->allocs.push_back(d);
        pk->d_challenges = d;
        pk->challenges_cap = ch->n_challenges;
    }
    if (ch->n_challenges) DE_CUDA(ctx, cudaMemcpyAsync(pk->d_challenges, ch->challenges, sizeof(Fr) * ch->n_challenges, cudaMemcpyHostToDevice, ctx->stream));
    // powers of y for the constraints behind the custom gates (k_eval_h): y^0 .. y^T, T = permutation + lookup constraints
    const uint32_t n_terms = (pk->n_sets ? 2 + (pk->n_sets - 1) + pk->n_sets : 0) + 5 * pk->n_lookups;
    {
        std::vector<host::HFr> yp(n_terms + 1);
        host::HFr yv;
        memcpy(yv.l, ch->y.l, 32);
        yp[0] = host::fr_one();
        for (uint32_t i = 1; i <= n_terms; i++) yp[i] = host::fr_mul(yp[i - 1], yv);
        if (!pk->d_ypows || pk->ypows_cap < n_terms + 1) {
            Fr* d = nullptr;
            DE_CUDA(ctx, cudaMalloc((void**)&d, sizeof(Fr) * (n_terms + 1)));
            pk->allocs.push_back(d);
            pk->d_ypows = d;
            pk->ypows_cap = n_terms + 1;
        }
        // pageable source: the runtime stages it before cudaMemcpyAsync returns, so the vector may go out of scope
        DE_CUDA(ctx, cudaMemcpyAsync(pk->d_ypows, yp.data(), sizeof(Fr) * (n_terms + 1), cudaMemcpyHostToDevice, ctx->stream));
    }
    EvalParams p;
    memset(&p, 0, sizeof(p));
    p.ypows = pk->d_ypows;
    p.n_terms = n_terms;
    p.ext_mask = (uint32_t)(ext_n - 1);
    p.rot_scale = 1u << (pk->ek - pk->k);
    p.ext_n = ext_n;
    const size_t n_polys = (size_t)pk->n_fixed + pk->n_perm_cols;
    p.fixed = pk->resident;
    p.sigma = pk->resident + (size_t)pk->n_fixed * ext_n;
    p.l0 = pk->resident + n_polys * ext_n;
    p.l_last = p.l0 + ext_n;
    p.l_active = p.l_last + ext_n;
    p.omega_pows = p.l_active + ext_n;
    p.advice = w_adv; p.instance = w_inst; p.permz = w_permz; p.lookup_z = w_lz; p.lookup_a = w_la; p.lookup_s = w_ls;
    p.challenges = pk->d_challenges;
    p.y = fr_from_host(ch->y); p.beta = fr_from_host(ch->beta); p.gamma = fr_from_host(ch->gamma); p.theta = fr_from_host(ch->theta);
    p.delta = pk->delta;
    p.gates = pk->gates;
    p.lookups = pk->d_lookups;
    p.n_lookups = pk->n_lookups;
    p.n_perm_cols = pk->n_perm_cols; p.chunk_len = pk->chunk_len; p.n_sets = pk->n_sets;
    p.last_rotation = -((int)pk->blinding + 1);
    p.perm_kind = pk->d_perm_kind; p.perm_index = pk->d_perm_index;
    p.out = (Fr*)d_h_ext;
    {
        // ZETA in Montgomery form (a field constant); delta_start = beta * ZETA is formed per thread in the kernel
        const uint32_t zm[8] = {0x55fcd653u, 0x0363f299u, 0x5fc1e200u, 0x73e7950bu, 0x576d9d24u, 0xc5fce83eu, 0xa1c3a4d4u, 0x059c805du};
        for (int i = 0; i < 8; i++) p.delta_start.l[i] = zm[i];
    }
    DE_TIMED(ctx, "k_eval_h", (double)ext_n, (k_eval_h<<<(unsigned int)((ext_n + 127) / 128), 128, 0, ctx->stream>>>(p)));
    DE_CHECK_LAUNCH(ctx);
    return DE_OK;
}

int de_evaluate_h_dev(de_pk* pk, const de_fr* d_advice, const de_fr* d_instance, const de_challenges* ch, const de_fr* d_permz,
                      const de_fr* d_lookup, size_t stride, de_fr* d_h_ext) {
    if (!pk) return DE_ERR_ARG;
    if (!ch || !d_h_ext) return fail(pk->ctx, DE_ERR_ARG, "de_evaluate_h: null pointer");
    DE_TRY(de_pk_extend_dev(pk, d_advice, d_instance, d_permz, d_lookup, stride));
    return de_evaluate_h_rows_dev(pk, ch, d_h_ext);
}

int de_evaluate_h(de_pk* pk, const de_fr* const* advice_coeff, const de_fr* const* instance_coeff, const de_challenges* ch,
                  const de_fr* const* perm_z_coeff, const de_fr* const* lookup_coeff, de_fr* h_ext_out) {
    if (!pk) return DE_ERR_ARG;
    de_ctx* ctx = pk->ctx;
    if (!ch || !h_ext_out) return fail(ctx, DE_ERR_ARG, "de_evaluate_h: null pointer");
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = pk->n, ext_n = pk->ext_n;
    const size_t counts[4] = {pk->n_advice, pk->n_instance, pk->n_sets, 3 * (size_t)pk->n_lookups};
    const de_fr* const* srcs[4] = {advice_coeff, instance_coeff, perm_z_coeff, lookup_coeff};
    const size_t total = counts[0] + counts[1] + counts[2] + counts[3];
    DE_WS(ctx, staging, Fr, WS_EVAL_A, sizeof(Fr) * n * (total ? total : 1));
    DE_WS(ctx, d_h, Fr, WS_EVAL_C, sizeof(Fr) * ext_n);
    Fr* blocks[4];
    size_t off = 0;
    for (int b = 0; b < 4; b++) {
        blocks[b] = staging + off * n;
        if (counts[b] && !srcs[b]) return fail(ctx, DE_ERR_ARG, "de_evaluate_h: missing polynomial array");
        for (size_t i = 0; i < counts[b]; i++) {
            if (!srcs[b][i]) return fail(ctx, DE_ERR_ARG, "de_evaluate_h: null polynomial");
            DE_CUDA(ctx, cudaMemcpyAsync(blocks[b] + i * n, srcs[b][i], sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
        }
        off += counts[b];
    }
    DE_TRY(de_evaluate_h_dev(pk, (const de_fr*)blocks[0], (const de_fr*)blocks[1], ch, (const de_fr*)blocks[2], (const de_fr*)blocks[3], n,
                             (de_fr*)d_h));
    DE_CUDA(ctx, cudaMemcpyAsync(h_ext_out, d_h, sizeof(Fr) * ext_n, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
}

}  // extern "C"
