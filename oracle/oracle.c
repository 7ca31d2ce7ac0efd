/*
 * oracle.c — CPU restatement (plain C, pthreads) of the hot path under the delay-encryption halo2 circuits.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path (libde_b200.so) never links it.
 *
 * PARITY UNPINNED at the MSM/NTT boundary: the algorithms restated here live in the un-vendored crates
 * halo2_proofs (tag v2023_04_20, /root/reference/Cargo.toml:17) and halo2curves 0.3.x; no source and no Rust
 * toolchain exist in this environment and the reference holds no golden vector for them.
 * Call sites in the reference that reach these functions: create_proof at benches/delay_enc.rs:123,
 * benches/mod_pow.rs:201, benches/pose_enc.rs:127; keygen_vk/keygen_pk at benches/delay_enc.rs:86,103.
 * The restatement follows SURVEY.md Appendix B (behavioural spec) and is cross-checked against
 * oracle/pyoracle.py, which is pinned by the reference's Poseidon known-answer vectors.
 *
 * Data conventions (identical to the C ABI in include/de_b200.h): Fr/Fq = 4 x u64 little-endian limbs in
 * Montgomery form (R = 2^256); G1Affine = {x, y} (identity = all zero); G1 = {x, y, z} Jacobian (identity z = 0).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef struct { u64 l[4]; } fe;
typedef struct { fe x, y; } aff;
typedef struct { fe x, y, z; } jac;

typedef struct { u64 p[4]; u64 inv; fe r; fe r2; } field_t;

static const field_t FRF = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
    {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}}};
static const field_t FQF = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
    {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}}};

/* ---------------------------------------------------------------- field arithmetic (halo2curves a1) */
static inline int fe_geq(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline void fe_sub_raw(u64 r[4], const u64 a[4], const u64 b[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)t;
        borrow = (t >> 64) & 1;
    }
}
static inline void f_add(const field_t* F, fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    u64 t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (u64)c;
        c >>= 64;
    }
    if (c || fe_geq(t, F->p)) fe_sub_raw(t, t, F->p);
    memcpy(r->l, t, 32);
}
static inline void f_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
    u64 t[4];
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        t[i] = (u64)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + F->p[i];
            t[i] = (u64)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, 32);
}
static inline int f_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int f_eq(const fe* a, const fe* b) { return memcmp(a, b, 32) == 0; }
static inline void f_neg(const field_t* F, fe* r, const fe* a) {
    if (f_is_zero(a)) { *r = *a; return; }
    fe_sub_raw(r->l, F->p, a->l);
}
/* Montgomery CIOS, 4 x 64-bit limbs */
static inline void f_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (u64)c;
        t[5] = (u64)(c >> 64);
        u64 m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (u64)c;
        t[4] = t[5] + (u64)(c >> 64);
    }
    if (t[4] || fe_geq(t, F->p)) fe_sub_raw(t, t, F->p);
    memcpy(r->l, t, 32);
}
static inline void f_sqr(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static void f_pow(const field_t* F, fe* r, const fe* a, const u64 e[4]) {
    fe acc = F->r, base = *a;
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) f_mul(F, &acc, &acc, &base);
        f_sqr(F, &base, &base);
    }
    *r = acc;
}
static void f_inv(const field_t* F, fe* r, const fe* a) {
    u64 e[4];
    u64 two[4] = {2, 0, 0, 0};
    fe_sub_raw(e, F->p, two);
    f_pow(F, r, a, e);
}
static inline void f_from_mont(const field_t* F, fe* r, const fe* a) {
    fe one = {{1, 0, 0, 0}};
    f_mul(F, r, a, &one);
}
static inline void f_to_mont(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, &F->r2); }
static inline void f_from_u64(const field_t* F, fe* r, u64 v) {
    fe t = {{v, 0, 0, 0}};
    f_to_mont(F, r, &t);
}
static inline void f_dbl(const field_t* F, fe* r, const fe* a) { f_add(F, r, a, a); }

/* exported element-wise vector ops, used by the field-kernel fuzz tests */
#define VEC2(name, FLD, op)                                                         \
    void name(const u64* a, const u64* b, u64* out, size_t n) {                     \
        for (size_t i = 0; i < n; i++) op(&FLD, (fe*)(out + 4 * i), (const fe*)(a + 4 * i), (const fe*)(b + 4 * i)); \
    }
VEC2(orc_fr_mul, FRF, f_mul)
VEC2(orc_fr_add, FRF, f_add)
VEC2(orc_fr_sub, FRF, f_sub)
VEC2(orc_fq_mul, FQF, f_mul)
VEC2(orc_fq_add, FQF, f_add)
VEC2(orc_fq_sub, FQF, f_sub)
void orc_fr_from_mont(const u64* a, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) f_from_mont(&FRF, (fe*)(out + 4 * i), (const fe*)(a + 4 * i));
}
void orc_fr_to_mont(const u64* a, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) f_to_mont(&FRF, (fe*)(out + 4 * i), (const fe*)(a + 4 * i));
}
void orc_fq_from_mont(const u64* a, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) f_from_mont(&FQF, (fe*)(out + 4 * i), (const fe*)(a + 4 * i));
}
void orc_fq_to_mont(const u64* a, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) f_to_mont(&FQF, (fe*)(out + 4 * i), (const fe*)(a + 4 * i));
}
void orc_fr_inv(const u64* a, u64* out, size_t n) {
    for (size_t i = 0; i < n; i++) f_inv(&FRF, (fe*)(out + 4 * i), (const fe*)(a + 4 * i));
}

/* ---------------------------------------------------------------- G1 (halo2curves a2): y^2 = x^3 + 3 */
static inline int aff_is_id(const aff* p) { return f_is_zero(&p->x) && f_is_zero(&p->y); }
static inline void jac_set_id(jac* p) { memset(p, 0, sizeof(*p)); }
static inline int jac_is_id(const jac* p) { return f_is_zero(&p->z); }
static inline void jac_from_aff(jac* r, const aff* p) {
    if (aff_is_id(p)) { jac_set_id(r); return; }
    r->x = p->x; r->y = p->y; r->z = FQF.r;
}
static void jac_double(jac* r, const jac* p) {
    const field_t* F = &FQF;
    if (jac_is_id(p)) { *r = *p; return; }
    fe a, b, c, d, e, f, t, x3, y3, z3;
    f_sqr(F, &a, &p->x);
    f_sqr(F, &b, &p->y);
    f_sqr(F, &c, &b);
    f_add(F, &t, &p->x, &b);
    f_sqr(F, &t, &t);
    f_sub(F, &t, &t, &a);
    f_sub(F, &t, &t, &c);
    f_dbl(F, &d, &t);
    f_dbl(F, &e, &a);
    f_add(F, &e, &e, &a);
    f_sqr(F, &f, &e);
    f_dbl(F, &t, &d);
    f_sub(F, &x3, &f, &t);
    f_sub(F, &t, &d, &x3);
    f_mul(F, &y3, &e, &t);
    f_dbl(F, &t, &c); f_dbl(F, &t, &t); f_dbl(F, &t, &t);
    f_sub(F, &y3, &y3, &t);
    f_mul(F, &z3, &p->y, &p->z);
    f_dbl(F, &z3, &z3);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(jac* r, const jac* p, const jac* q) {
    const field_t* F = &FQF;
    if (jac_is_id(p)) { *r = *q; return; }
    if (jac_is_id(q)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
    f_sqr(F, &z1z1, &p->z);
    f_sqr(F, &z2z2, &q->z);
    f_mul(F, &u1, &p->x, &z2z2);
    f_mul(F, &u2, &q->x, &z1z1);
    f_mul(F, &s1, &p->y, &q->z); f_mul(F, &s1, &s1, &z2z2);
    f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
    if (f_eq(&u1, &u2)) {
        if (f_eq(&s1, &s2)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    f_sub(F, &h, &u2, &u1);
    f_sub(F, &rr, &s2, &s1);
    f_sqr(F, &hh, &h);
    f_mul(F, &hhh, &h, &hh);
    f_mul(F, &v, &u1, &hh);
    f_sqr(F, &x3, &rr);
    f_sub(F, &x3, &x3, &hhh);
    f_dbl(F, &t, &v);
    f_sub(F, &x3, &x3, &t);
    f_sub(F, &t, &v, &x3);
    f_mul(F, &y3, &rr, &t);
    f_mul(F, &t, &s1, &hhh);
    f_sub(F, &y3, &y3, &t);
    f_mul(F, &z3, &p->z, &q->z);
    f_mul(F, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_aff(jac* r, const jac* p, const aff* q) {
    const field_t* F = &FQF;
    if (aff_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { jac_from_aff(r, q); return; }
    fe z1z1, u2, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
    f_sqr(F, &z1z1, &p->z);
    f_mul(F, &u2, &q->x, &z1z1);
    f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
    if (f_eq(&p->x, &u2)) {
        if (f_eq(&p->y, &s2)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    f_sub(F, &h, &u2, &p->x);
    f_sub(F, &rr, &s2, &p->y);
    f_sqr(F, &hh, &h);
    f_mul(F, &hhh, &h, &hh);
    f_mul(F, &v, &p->x, &hh);
    f_sqr(F, &x3, &rr);
    f_sub(F, &x3, &x3, &hhh);
    f_dbl(F, &t, &v);
    f_sub(F, &x3, &x3, &t);
    f_sub(F, &t, &v, &x3);
    f_mul(F, &y3, &rr, &t);
    f_mul(F, &t, &p->y, &hhh);
    f_sub(F, &y3, &y3, &t);
    f_mul(F, &z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_to_aff(aff* r, const jac* p) {
    const field_t* F = &FQF;
    if (jac_is_id(p)) { memset(r, 0, sizeof(*r)); return; }
    fe zi, zi2, zi3;
    f_inv(F, &zi, &p->z);
    f_sqr(F, &zi2, &zi);
    f_mul(F, &zi3, &zi2, &zi);
    f_mul(F, &r->x, &p->x, &zi2);
    f_mul(F, &r->y, &p->y, &zi3);
}
void orc_g1_to_affine(const u64* j, u64* a, size_t n) {
    for (size_t i = 0; i < n; i++) jac_to_aff((aff*)(a + 8 * i), (const jac*)(j + 12 * i));
}
void orc_g1_add(const u64* p, const u64* q, u64* r) { jac_add((jac*)r, (const jac*)p, (const jac*)q); }
void orc_g1_double(const u64* p, u64* r) { jac_double((jac*)r, (const jac*)p); }
int orc_g1_on_curve(const u64* a) {
    const aff* p = (const aff*)a;
    if (aff_is_id(p)) return 1;
    fe y2, x3, three;
    f_sqr(&FQF, &y2, &p->y);
    f_sqr(&FQF, &x3, &p->x);
    f_mul(&FQF, &x3, &x3, &p->x);
    f_from_u64(&FQF, &three, 3);
    f_add(&FQF, &x3, &x3, &three);
    return f_eq(&y2, &x3);
}
/* scalar given as canonical (non-Montgomery) 4 x u64 */
static void jac_mul_canon(jac* r, const aff* p, const u64 k[4]) {
    jac acc;
    jac_set_id(&acc);
    for (int i = 255; i >= 0; i--) {
        jac_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) jac_add_aff(&acc, &acc, p);
    }
    *r = acc;
}
void orc_g1_mul(const u64* base_aff, const u64* scalar_mont, u64* out_jac) {
    fe k;
    f_from_mont(&FRF, &k, (const fe*)scalar_mont);
    jac_mul_canon((jac*)out_jac, (const aff*)base_aff, k.l);
}

/* ---------------------------------------------------------------- thread helper */
typedef void (*range_fn)(void* ctx, size_t lo, size_t hi, int tid);
typedef struct { range_fn fn; void* ctx; size_t lo, hi; int tid; } job_t;
static void* job_tramp(void* p) {
    job_t* j = (job_t*)p;
    j->fn(j->ctx, j->lo, j->hi, j->tid);
    return NULL;
}
/* split [0,n) into `threads` contiguous ranges (halo2's `parallelize`) */
static void parallel_for(size_t n, int threads, range_fn fn, void* ctx) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(ctx, 0, n, 0); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
    size_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        jobs[t] = (job_t){fn, ctx, lo, hi, t};
        pthread_create(&th[t], NULL, job_tramp, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* out[i] = [scalars[i]] base as affine points (ParamsKZG::setup's g / g_lagrange for a known secret), threaded */
typedef struct { const aff* base; const fe* scalars; aff* out; } mulmany_ctx;
static void mulmany_job(void* v, size_t lo, size_t hi, int tid) {
    (void)tid;
    mulmany_ctx* c = (mulmany_ctx*)v;
    for (size_t i = lo; i < hi; i++) {
        fe k;
        jac r;
        f_from_mont(&FRF, &k, &c->scalars[i]);
        jac_mul_canon(&r, c->base, k.l);
        jac_to_aff(&c->out[i], &r);
    }
}
void orc_g1_mul_many(const u64* base_aff, const u64* scalars_mont, size_t n, u64* out_aff, int threads) {
    mulmany_ctx c = {(const aff*)base_aff, (const fe*)scalars_mont, (aff*)out_aff};
    parallel_for(n, threads, mulmany_job, &c);
}

/* ---------------------------------------------------------------- vector helpers of the prover (halo2_proofs::arithmetic /
 * poly): used by oracle/pyprover.py's array-based create_proof so that the restated prover runs at the reference's cost
 * profile (C inner loops, threads) instead of Python integer loops */
/* ff::BatchInvert over a slice, in place; zeros stay zero.  halo2 calls it per `parallelize` chunk. */
typedef struct { fe* a; } binv_ctx;
static void binv_job(void* v, size_t lo, size_t hi, int tid) {
    (void)tid;
    fe* a = ((binv_ctx*)v)->a;
    if (hi <= lo) return;
    fe* pre = (fe*)malloc(sizeof(fe) * (hi - lo));
    fe acc = FRF.r;
    for (size_t i = lo; i < hi; i++) {
        pre[i - lo] = acc;
        if (!f_is_zero(&a[i])) f_mul(&FRF, &acc, &acc, &a[i]);
    }
    fe inv;
    f_inv(&FRF, &inv, &acc);
    for (size_t i = hi; i-- > lo;) {
        if (f_is_zero(&a[i])) continue;
        fe t;
        f_mul(&FRF, &t, &inv, &pre[i - lo]);
        f_mul(&FRF, &inv, &inv, &a[i]);
        a[i] = t;
    }
    free(pre);
}
void orc_fr_batch_invert(u64* a, size_t n, int threads) {
    binv_ctx c = {(fe*)a};
    parallel_for(n, threads, binv_job, &c);
}
/* z[0] = start; z[i] = z[i-1] * f[i-1] for i < n (the serial running product of permutation::prover / lookup::prover) */
void orc_fr_running_product(const u64* f, size_t n, const u64* start, u64* z) {
    const fe* ff = (const fe*)f;
    fe* zz = (fe*)z;
    if (n == 0) return;
    zz[0] = *(const fe*)start;
    for (size_t i = 1; i < n; i++) f_mul(&FRF, &zz[i], &zz[i - 1], &ff[i - 1]);
}
/* arithmetic::eval_polynomial: per-thread Horner over contiguous chunks, each scaled by point^start, then summed */
typedef struct { const fe* poly; size_t n, chunk; const fe* x; fe* parts; } evp_ctx;
static void evp_job(void* v, size_t lo, size_t hi, int tid) {
    (void)tid;
    evp_ctx* c = (evp_ctx*)v;
    for (size_t part = lo; part < hi; part++) {
        size_t s0 = part * c->chunk, s1 = s0 + c->chunk;
        if (s1 > c->n) s1 = c->n;
        fe acc;
        memset(&acc, 0, sizeof(acc));
        for (size_t i = s1; i-- > s0;) {
            f_mul(&FRF, &acc, &acc, c->x);
            f_add(&FRF, &acc, &acc, &c->poly[i]);
        }
        u64 e[4] = {(u64)s0, 0, 0, 0};
        fe pw;
        f_pow(&FRF, &pw, c->x, e);
        f_mul(&FRF, &c->parts[part], &acc, &pw);
    }
}
void orc_eval_polynomial(const u64* poly, size_t n, const u64* x, u64* out, int threads) {
    fe* o = (fe*)out;
    memset(o, 0, sizeof(fe));
    if (n == 0) return;
    if (threads < 1) threads = 1;
    size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
    size_t nparts = (n + chunk - 1) / chunk;
    fe* parts = (fe*)calloc(nparts, sizeof(fe));
    evp_ctx c = {(const fe*)poly, n, chunk, (const fe*)x, parts};
    parallel_for(nparts, threads, evp_job, &c);
    for (size_t i = 0; i < nparts; i++) f_add(&FRF, o, o, &parts[i]);
    free(parts);
}
/* arithmetic::kate_division: q = a / (X - b), n - 1 coefficients */
void orc_kate_division(const u64* a, size_t n, const u64* b, u64* q) {
    const fe* aa = (const fe*)a;
    fe* qq = (fe*)q;
    fe tmp;
    memset(&tmp, 0, sizeof(tmp));
    for (size_t i = n - 1; i >= 1; i--) {
        fe t;
        f_mul(&FRF, &t, &tmp, (const fe*)b);
        f_add(&FRF, &tmp, &aa[i], &t);
        qq[i - 1] = tmp;
    }
}
/* acc[i] = acc[i] * s + p[i]  (Polynomial * scalar + &Polynomial, the fold of h pieces / theta compression) */
typedef struct { fe* acc; const fe* p; const fe* s; } sadd_ctx;
static void sadd_job(void* v, size_t lo, size_t hi, int tid) {
    (void)tid;
    sadd_ctx* c = (sadd_ctx*)v;
    for (size_t i = lo; i < hi; i++) {
        f_mul(&FRF, &c->acc[i], &c->acc[i], c->s);
        f_add(&FRF, &c->acc[i], &c->acc[i], &c->p[i]);
    }
}
void orc_fr_scale_add(u64* acc, const u64* p, const u64* s, size_t n, int threads) {
    sadd_ctx c = {(fe*)acc, (const fe*)p, (const fe*)s};
    parallel_for(n, threads, sadd_job, &c);
}
/* acc[i] += s * p[i]  (the GWC batch: poly_acc + poly * power_of_v) */
static void axpy_job(void* v, size_t lo, size_t hi, int tid) {
    (void)tid;
    sadd_ctx* c = (sadd_ctx*)v;
    for (size_t i = lo; i < hi; i++) {
        fe t;
        f_mul(&FRF, &t, &c->p[i], c->s);
        f_add(&FRF, &c->acc[i], &c->acc[i], &t);
    }
}
void orc_fr_axpy(u64* acc, const u64* p, const u64* s, size_t n, int threads) {
    sadd_ctx c = {(fe*)acc, (const fe*)p, (const fe*)s};
    parallel_for(n, threads, axpy_job, &c);
}
/* out[i] = omega^i * scale for i < n (the delta^j omega^i beta column of the permutation numerator), serial like the reference's
 * per-chunk running multiplication */
void orc_fr_geometric(const u64* start, const u64* ratio, size_t n, u64* out) {
    fe* o = (fe*)out;
    if (n == 0) return;
    o[0] = *(const fe*)start;
    for (size_t i = 1; i < n; i++) f_mul(&FRF, &o[i], &o[i - 1], (const fe*)ratio);
}

/* ---------------------------------------------------------------- best_multiexp (a3; Appendix B.1) */
enum { B_NONE = 0, B_AFFINE = 1, B_PROJ = 2 };
typedef struct { int tag; jac p; } bucket_t; /* Affine keeps x,y in p.x,p.y */

static void bucket_add_assign(bucket_t* b, const aff* q) {
    if (b->tag == B_NONE) {
        b->tag = B_AFFINE; b->p.x = q->x; b->p.y = q->y;
    } else if (b->tag == B_AFFINE) {
        aff a = {b->p.x, b->p.y};
        jac t;
        jac_from_aff(&t, &a);
        jac_add_aff(&b->p, &t, q);
        b->tag = B_PROJ;
    } else {
        jac_add_aff(&b->p, &b->p, q);
    }
}
static void bucket_add_to(const bucket_t* b, jac* other) {
    if (b->tag == B_NONE) return;
    if (b->tag == B_AFFINE) {
        aff a = {b->p.x, b->p.y};
        jac_add_aff(other, other, &a);
    } else {
        jac_add(other, other, &b->p);
    }
}
static void multiexp_serial(const fe* coeffs_mont, const aff* bases, size_t n, jac* acc) {
    fe* reprs = (fe*)malloc(sizeof(fe) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) f_from_mont(&FRF, &reprs[i], &coeffs_mont[i]);
    unsigned c;
    if (n < 4) c = 1;
    else if (n < 32) c = 3;
    else c = (unsigned)ceil(log((double)n));
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    bucket_t* buckets = (bucket_t*)malloc(sizeof(bucket_t) * nb);
    for (int seg = (int)segments - 1; seg >= 0; seg--) {
        for (unsigned d = 0; d < c; d++) jac_double(acc, acc);
        for (size_t b = 0; b < nb; b++) buckets[b].tag = B_NONE;
        size_t skip_bits = (size_t)seg * c, skip_bytes = skip_bits / 8;
        for (size_t i = 0; i < n; i++) {
            u64 digit = 0;
            if (skip_bytes < 32) {
                unsigned char v[8] = {0};
                const unsigned char* bytes = (const unsigned char*)reprs[i].l;
                size_t len = 32 - skip_bytes;
                if (len > 8) len = 8;
                memcpy(v, bytes + skip_bytes, len);
                u64 tmp;
                memcpy(&tmp, v, 8);
                tmp >>= skip_bits - skip_bytes * 8;
                digit = tmp % ((u64)1 << c);
            }
            if (digit != 0) bucket_add_assign(&buckets[digit - 1], &bases[i]);
        }
        jac running;
        jac_set_id(&running);
        for (size_t b = nb; b-- > 0;) {
            bucket_add_to(&buckets[b], &running);
            jac_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(reprs);
}
typedef struct { const fe* s; const aff* b; size_t n, chunk; jac* partial; } msm_ctx;
static void msm_job(void* vctx, size_t lo, size_t hi, int tid) {
    msm_ctx* c = (msm_ctx*)vctx;
    (void)tid;
    for (size_t k = lo; k < hi; k++) {
        size_t a = k * c->chunk, e = a + c->chunk;
        if (e > c->n) e = c->n;
        jac_set_id(&c->partial[k]);
        multiexp_serial(c->s + a, c->b + a, e - a, &c->partial[k]);
    }
}
/* threads = rayon::current_num_threads() in the reference */
int orc_best_multiexp(const u64* scalars, const u64* bases, size_t n, u64* out_jac, int threads) {
    jac acc;
    jac_set_id(&acc);
    if (threads < 1) threads = 1;
    if (n > (size_t)threads) {
        size_t chunk = n / threads;
        size_t nchunks = (n + chunk - 1) / chunk;
        jac* partial = (jac*)malloc(sizeof(jac) * nchunks);
        msm_ctx c = {(const fe*)scalars, (const aff*)bases, n, chunk, partial};
        parallel_for(nchunks, (int)nchunks, msm_job, &c);
        for (size_t k = 0; k < nchunks; k++) jac_add(&acc, &acc, &partial[k]);
        free(partial);
    } else {
        multiexp_serial((const fe*)scalars, (const aff*)bases, n, &acc);
    }
    memcpy(out_jac, &acc, sizeof(acc));
    return 0;
}
/* definition-level MSM (double-and-add per term), for cross-checking the Pippenger restatement */
int orc_msm_naive(const u64* scalars, const u64* bases, size_t n, u64* out_jac) {
    jac acc, t;
    jac_set_id(&acc);
    for (size_t i = 0; i < n; i++) {
        fe k;
        f_from_mont(&FRF, &k, (const fe*)(scalars + 4 * i));
        jac_mul_canon(&t, (const aff*)(bases + 8 * i), k.l);
        jac_add(&acc, &acc, &t);
    }
    memcpy(out_jac, &acc, sizeof(acc));
    return 0;
}

/* ---------------------------------------------------------------- best_fft (a4; Appendix B.2) */
static inline uint32_t bitrev32(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static void butterfly_rec(fe* a, size_t n, size_t twiddle_chunk, const fe* tw) {
    const field_t* F = &FRF;
    if (n == 2) {
        fe t = a[1];
        f_sub(F, &a[1], &a[0], &t);
        f_add(F, &a[0], &a[0], &t);
        return;
    }
    size_t half = n / 2;
    butterfly_rec(a, half, twiddle_chunk * 2, tw);
    butterfly_rec(a + half, half, twiddle_chunk * 2, tw);
    fe t = a[half];
    f_sub(F, &a[half], &a[0], &t);
    f_add(F, &a[0], &a[0], &t);
    for (size_t i = 1; i < half; i++) {
        f_mul(F, &t, &a[half + i], &tw[i * twiddle_chunk]);
        f_sub(F, &a[half + i], &a[i], &t);
        f_add(F, &a[i], &a[i], &t);
    }
}
typedef struct { fe* a; size_t n; const fe* tw; size_t sub; size_t tchunk; size_t half; } fft_ctx;
static void fft_sub_job(void* v, size_t lo, size_t hi, int tid) {
    fft_ctx* c = (fft_ctx*)v;
    (void)tid;
    for (size_t s = lo; s < hi; s++) {
        if (c->sub >= 2) butterfly_rec(c->a + s * c->sub, c->sub, c->tchunk, c->tw);
    }
}
static void fft_stage_job(void* v, size_t lo, size_t hi, int tid) {
    /* butterflies lo..hi of one upper stage: block size 2*half, twiddle stride tchunk */
    fft_ctx* c = (fft_ctx*)v;
    const field_t* F = &FRF;
    (void)tid;
    for (size_t k = lo; k < hi; k++) {
        size_t blk = k / c->half, i = k % c->half;
        fe* x = c->a + blk * 2 * c->half + i;
        fe* y = x + c->half;
        fe t;
        if (i == 0) t = *y; else f_mul(F, &t, y, &c->tw[i * c->tchunk]);
        f_sub(F, y, x, &t);
        f_add(F, x, x, &t);
    }
}
typedef struct { fe* tw; fe omega; size_t n; } tw_ctx;
static void tw_job(void* v, size_t lo, size_t hi, int tid) {
    tw_ctx* c = (tw_ctx*)v;
    (void)tid;
    if (lo >= hi) return;
    u64 e[4] = {lo, 0, 0, 0};
    fe w;
    f_pow(&FRF, &w, &c->omega, e);
    for (size_t i = lo; i < hi; i++) { c->tw[i] = w; f_mul(&FRF, &w, &w, &c->omega); }
}
void orc_best_fft(u64* av, const u64* omega, uint32_t log_n, int threads) {
    fe* a = (fe*)av;
    size_t n = (size_t)1 << log_n;
    if (threads < 1) threads = 1;
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitrev32((uint32_t)k, log_n);
        if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    if (n < 2) return;
    fe* tw = (fe*)malloc(sizeof(fe) * (n / 2));
    tw_ctx tc = {tw, *(const fe*)omega, n};
    parallel_for(n / 2, threads, tw_job, &tc);
    /* The reference joins recursively; here the bottom sub-transforms run one per task and the upper
     * log2(tasks) stages are parallelised over butterflies: same butterflies, same twiddles. */
    uint32_t lt = 0;
    while (((size_t)2 << lt) <= (size_t)threads && lt + 1 < log_n) lt++;
    size_t nsub = (size_t)1 << lt, sub = n >> lt;
    fft_ctx c = {a, n, tw, sub, nsub, 0};
    parallel_for(nsub, threads, fft_sub_job, &c);
    for (uint32_t lvl = lt; lvl-- > 0;) {
        c.half = n >> (lvl + 1);
        c.tchunk = (size_t)1 << lvl;
        parallel_for(n / 2, threads, fft_stage_job, &c);
    }
    free(tw);
}

/* ---------------------------------------------------------------- EvaluationDomain (a5-a7; Appendix B.3) */
typedef struct {
    uint32_t j, k, extended_k;
    size_t n, ext_n, qdeg;
    fe omega, omega_inv, ext_omega, ext_omega_inv, zeta, zeta_inv, ifft_div, ext_ifft_div;
    fe* t_evals;
    size_t t_len;
    int threads;
} orc_domain;

static const fe ZETA_CANON = {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};
static const fe ROOT_CANON = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};

orc_domain* orc_domain_new(uint32_t j, uint32_t k, int threads) {
    const field_t* F = &FRF;
    orc_domain* d = (orc_domain*)calloc(1, sizeof(*d));
    d->j = j; d->k = k; d->n = (size_t)1 << k; d->qdeg = j - 1; d->threads = threads;
    uint32_t ek = k;
    while (((size_t)1 << ek) < d->n * d->qdeg) ek++;
    d->extended_k = ek; d->ext_n = (size_t)1 << ek;
    fe w;
    f_to_mont(F, &w, &ROOT_CANON);
    for (uint32_t i = ek; i < 28; i++) f_sqr(F, &w, &w);
    d->ext_omega = w;
    f_inv(F, &d->ext_omega_inv, &w);
    for (uint32_t i = k; i < ek; i++) f_sqr(F, &w, &w);
    d->omega = w;
    f_inv(F, &d->omega_inv, &w);
    f_to_mont(F, &d->zeta, &ZETA_CANON);
    f_sqr(F, &d->zeta_inv, &d->zeta);
    fe t;
    f_from_u64(F, &t, (u64)d->n);
    f_inv(F, &d->ifft_div, &t);
    f_from_u64(F, &t, (u64)d->ext_n);
    f_inv(F, &d->ext_ifft_div, &t);
    d->t_len = (size_t)1 << (ek - k);
    d->t_evals = (fe*)malloc(sizeof(fe) * d->t_len);
    fe cur = d->zeta;
    u64 e[4] = {d->n, 0, 0, 0};
    for (size_t i = 0; i < d->t_len; i++) {
        fe v;
        f_pow(F, &v, &cur, e);
        f_sub(F, &v, &v, &F->r);
        f_inv(F, &d->t_evals[i], &v);
        f_mul(F, &cur, &cur, &d->ext_omega);
    }
    return d;
}
void orc_domain_free(orc_domain* d) { if (d) { free(d->t_evals); free(d); } }
uint32_t orc_domain_extended_k(const orc_domain* d) { return d->extended_k; }
void orc_domain_constants(const orc_domain* d, u64* out /* omega, omega_inv, ext_omega, ext_omega_inv: 16 u64 */) {
    memcpy(out, &d->omega, 32); memcpy(out + 4, &d->omega_inv, 32);
    memcpy(out + 8, &d->ext_omega, 32); memcpy(out + 12, &d->ext_omega_inv, 32);
}
void orc_domain_t_evaluations(const orc_domain* d, u64* out) { memcpy(out, d->t_evals, 32 * d->t_len); }

typedef struct { fe* a; const fe* c; size_t m; } scale_ctx;
static void scale_mod_job(void* v, size_t lo, size_t hi, int tid) {
    scale_ctx* c = (scale_ctx*)v;
    (void)tid;
    for (size_t i = lo; i < hi; i++) f_mul(&FRF, &c->a[i], &c->a[i], &c->c[i % c->m]);
}
void orc_lagrange_to_coeff(const orc_domain* d, u64* a) {
    orc_best_fft(a, d->omega_inv.l, d->k, d->threads);
    scale_ctx c = {(fe*)a, &d->ifft_div, 1};
    parallel_for(d->n, d->threads, scale_mod_job, &c);
}
void orc_coeff_to_lagrange(const orc_domain* d, u64* a) { orc_best_fft(a, d->omega.l, d->k, d->threads); }
/* in: n coefficients; out: ext_n evaluations over the zeta-coset */
void orc_coeff_to_extended(const orc_domain* d, const u64* in, u64* out) {
    fe z[3] = {FRF.r, d->zeta, d->zeta_inv};
    memcpy(out, in, 32 * d->n);
    memset(out + 4 * d->n, 0, 32 * (d->ext_n - d->n));
    scale_ctx c = {(fe*)out, z, 3};
    parallel_for(d->n, d->threads, scale_mod_job, &c);
    orc_best_fft(out, d->ext_omega.l, d->extended_k, d->threads);
}
/* in place over ext_n; the first n*(j-1) entries are the result; returns that length */
size_t orc_extended_to_coeff(const orc_domain* d, u64* a) {
    orc_best_fft(a, d->ext_omega_inv.l, d->extended_k, d->threads);
    fe z[3];
    z[0] = d->ext_ifft_div;
    f_mul(&FRF, &z[1], &d->ext_ifft_div, &d->zeta_inv);
    f_mul(&FRF, &z[2], &d->ext_ifft_div, &d->zeta);
    scale_ctx c = {(fe*)a, z, 3};
    parallel_for(d->ext_n, d->threads, scale_mod_job, &c);
    return d->n * d->qdeg;
}
void orc_divide_by_vanishing(const orc_domain* d, u64* a) {
    scale_ctx c = {(fe*)a, d->t_evals, d->t_len};
    parallel_for(d->ext_n, d->threads, scale_mod_job, &c);
}

/* ---------------------------------------------------------------- synthetic inputs */
typedef struct { u64 s[4]; } xo_t;
static inline u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
static void xo_seed(xo_t* x, u64 seed) {
    u64 s = seed;
    for (int i = 0; i < 4; i++) {
        s += 0x9E3779B97F4A7C15ULL;
        u64 z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        x->s[i] = z ^ (z >> 31);
    }
}
static inline u64 xo_next(xo_t* x) {
    u64* s = x->s;
    u64 result = rotl(s[1] * 5, 7) * 9;
    u64 t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
}
/* 512 random bits mod r, returned in Montgomery form: lo*R + hi*R^2 mont-reduced (from_uniform_bytes) */
static void xo_uniform_fr(xo_t* x, fe* out) {
    fe lo, hi, r3, a, b;
    for (int i = 0; i < 4; i++) lo.l[i] = xo_next(x);
    for (int i = 0; i < 4; i++) hi.l[i] = xo_next(x);
    /* value = lo + hi*2^256; Montgomery form = lo*R + hi*R^2 (mod r) = mont(lo,R2) + mont(hi,R3) */
    f_mul(&FRF, &r3, &FRF.r2, &FRF.r2);
    f_mul(&FRF, &a, &lo, &FRF.r2);
    f_mul(&FRF, &b, &hi, &r3);
    f_add(&FRF, out, &a, &b);
}
void orc_fill_uniform_fr(u64 seed, size_t n, u64* out) {
    xo_t x;
    xo_seed(&x, seed);
    for (size_t i = 0; i < n; i++) xo_uniform_fr(&x, (fe*)(out + 4 * i));
}
/* witness-like column (SURVEY.md 8d "W"): rows >= used are zero except the last 6 (blinding) which are uniform;
 * of the used rows 45% < 2^8, 35% < 2^64, 10% < 2^134, 10% uniform. */
void orc_fill_witness_fr(u64 seed, size_t n, size_t used, u64* out) {
    xo_t x;
    xo_seed(&x, seed);
    for (size_t i = 0; i < n; i++) {
        fe* o = (fe*)(out + 4 * i);
        fe v = {{0, 0, 0, 0}};
        if (i + 6 >= n) { xo_uniform_fr(&x, o); continue; }
        if (i >= used) { *o = v; continue; }
        u64 sel = xo_next(&x) % 100;
        if (sel < 45) { v.l[0] = xo_next(&x) & 0xff; f_to_mont(&FRF, o, &v); }
        else if (sel < 80) { v.l[0] = xo_next(&x); f_to_mont(&FRF, o, &v); }
        else if (sel < 90) { v.l[0] = xo_next(&x); v.l[1] = xo_next(&x); v.l[2] = xo_next(&x) & 0x3f; f_to_mont(&FRF, o, &v); }
        else xo_uniform_fr(&x, o);
    }
}
/* bases P_i = (start+i+1) * G, G = (1, 2), written affine; per-thread scalar-mul start then a chain of adds,
 * normalised with a batched inversion (Montgomery's trick) */
typedef struct { aff* out; size_t n; size_t start; } gen_ctx;
static void gen_job(void* v, size_t lo, size_t hi, int tid) {
    gen_ctx* c = (gen_ctx*)v;
    (void)tid;
    if (lo >= hi) return;
    const field_t* F = &FQF;
    aff g;
    f_from_u64(F, &g.x, 1);
    f_from_u64(F, &g.y, 2);
    u64 k[4] = {c->start + lo + 1, 0, 0, 0};
    jac cur;
    jac_mul_canon(&cur, &g, k);
    size_t m = hi - lo;
    jac* pts = (jac*)malloc(sizeof(jac) * m);
    fe* pre = (fe*)malloc(sizeof(fe) * m);
    for (size_t i = 0; i < m; i++) { pts[i] = cur; jac_add_aff(&cur, &cur, &g); }
    fe acc = F->r;
    for (size_t i = 0; i < m; i++) { pre[i] = acc; f_mul(F, &acc, &acc, &pts[i].z); }
    fe ainv;
    f_inv(F, &ainv, &acc);
    for (size_t i = m; i-- > 0;) {
        fe zi, zi2, zi3;
        f_mul(F, &zi, &ainv, &pre[i]);
        f_mul(F, &ainv, &ainv, &pts[i].z);
        f_sqr(F, &zi2, &zi);
        f_mul(F, &zi3, &zi2, &zi);
        f_mul(F, &c->out[lo + i].x, &pts[i].x, &zi2);
        f_mul(F, &c->out[lo + i].y, &pts[i].y, &zi3);
    }
    free(pts);
    free(pre);
}
void orc_gen_bases(size_t start, size_t n, u64* out, int threads) {
    gen_ctx c = {(aff*)out, n, start};
    parallel_for(n, threads, gen_job, &c);
}

/* ---------------------------------------------------------------- Evaluator::evaluate_h (a9; Appendix B.5)
 * Restates halo2_proofs::plonk::evaluation::{GraphEvaluator::evaluate, Evaluator::evaluate_h} and the l0 / l_last /
 * l_active_row construction of plonk::keygen.  Uses only the STRUCT LAYOUTS of include/de_b200.h (the descriptor a
 * ProvingKey is marshalled into); no product code is linked. */
#include "../include/de_b200.h"

typedef struct {
    const orc_domain* d;
    const de_pk_desc* desc;
    fe **fixed, **sigma, **advice, **instance, **permz, **lookup;
    fe *l0, *l_last, *l_active;
    fe y, beta, gamma, theta, delta;
    const fe* challenges;
    fe* out;
    size_t n_sets;
} evalh_ctx;

static inline size_t rot_idx(size_t idx, int rot, size_t rot_scale, size_t size) {
    long long v = (long long)idx + (long long)rot * (long long)rot_scale;
    long long m = (long long)size;
    v %= m;
    if (v < 0) v += m;
    return (size_t)v;
}
static fe graph_fetch(const evalh_ctx* c, const de_graph* g, const de_value_source* s, const fe* inter, const size_t* rots, const fe* prev) {
    switch (s->kind) {
        case DE_VAL_CONSTANT: return *(const fe*)&g->constants[s->index];
        case DE_VAL_INTERMEDIATE: return inter[s->index];
        case DE_VAL_FIXED: return c->fixed[s->index][rots[s->rotation]];
        case DE_VAL_ADVICE: return c->advice[s->index][rots[s->rotation]];
        case DE_VAL_INSTANCE: return c->instance[s->index][rots[s->rotation]];
        case DE_VAL_CHALLENGE: return c->challenges[s->index];
        case DE_VAL_BETA: return c->beta;
        case DE_VAL_GAMMA: return c->gamma;
        case DE_VAL_THETA: return c->theta;
        case DE_VAL_Y: return c->y;
        default: return *prev;
    }
}
static fe graph_eval(const evalh_ctx* c, const de_graph* g, fe* inter, size_t idx, size_t rot_scale, size_t size, const fe* prev) {
    const field_t* F = &FRF;
    size_t rots[64];
    for (uint32_t r = 0; r < g->n_rotations; r++) rots[r] = rot_idx(idx, g->rotations[r], rot_scale, size);
    fe last;
    memset(&last, 0, sizeof(last));
    for (uint32_t i = 0; i < g->n_calcs; i++) {
        const de_calculation* k = &g->calcs[i];
        fe a = graph_fetch(c, g, &k->a, inter, rots, prev), b, r;
        switch (k->op) {
            case DE_CALC_ADD: b = graph_fetch(c, g, &k->b, inter, rots, prev); f_add(F, &r, &a, &b); break;
            case DE_CALC_SUB: b = graph_fetch(c, g, &k->b, inter, rots, prev); f_sub(F, &r, &a, &b); break;
            case DE_CALC_MUL: b = graph_fetch(c, g, &k->b, inter, rots, prev); f_mul(F, &r, &a, &b); break;
            case DE_CALC_SQUARE: f_sqr(F, &r, &a); break;
            case DE_CALC_DOUBLE: f_dbl(F, &r, &a); break;
            case DE_CALC_NEGATE: f_neg(F, &r, &a); break;
            case DE_CALC_HORNER:
                b = graph_fetch(c, g, &k->b, inter, rots, prev);
                r = a;
                for (uint32_t h = 0; h < k->horner_len; h++) {
                    fe part = graph_fetch(c, g, &g->horner_parts[k->horner_first + h], inter, rots, prev);
                    f_mul(F, &r, &r, &b);
                    f_add(F, &r, &r, &part);
                }
                break;
            default: r = a; break;
        }
        inter[k->target] = r;
        last = r;
    }
    return last;
}
#define FOLD(expr_fe)                              \
    do {                                           \
        f_mul(F, &value, &value, &c->y);           \
        fe t__ = (expr_fe);                        \
        f_add(F, &value, &value, &t__);            \
    } while (0)
static inline fe fe_mul2(const fe* a, const fe* b) { fe r; f_mul(&FRF, &r, a, b); return r; }
static inline fe fe_sub2(const fe* a, const fe* b) { fe r; f_sub(&FRF, &r, a, b); return r; }
static inline fe fe_add2(const fe* a, const fe* b) { fe r; f_add(&FRF, &r, a, b); return r; }

static void evalh_job(void* v, size_t lo, size_t hi, int tid) {
    const evalh_ctx* c = (const evalh_ctx*)v;
    const field_t* F = &FRF;
    const de_pk_desc* desc = c->desc;
    const orc_domain* d = c->d;
    (void)tid;
    if (lo >= hi) return;
    const size_t size = d->ext_n, rot_scale = (size_t)1 << (d->extended_k - d->k);
    uint32_t max_inter = desc->gates.n_intermediates;
    for (uint32_t i = 0; i < desc->n_lookups; i++)
        if (desc->lookups[i].n_intermediates > max_inter) max_inter = desc->lookups[i].n_intermediates;
    fe* inter = (fe*)malloc(sizeof(fe) * (max_inter + 1));
    const fe one = F->r;
    fe zero;
    memset(&zero, 0, sizeof(zero));
    const int last_rotation = -((int)desc->blinding_factors + 1);
    fe delta_start = fe_mul2(&c->beta, &d->zeta);
    u64 e[4] = {lo, 0, 0, 0};
    fe beta_term;
    f_pow(F, &beta_term, &d->ext_omega, e);
    for (size_t idx = lo; idx < hi; idx++) {
        /* custom gates */
        fe value = graph_eval(c, &desc->gates, inter, idx, rot_scale, size, &zero);
        const size_t r_next = rot_idx(idx, 1, rot_scale, size), r_prev = rot_idx(idx, -1, rot_scale, size);
        const size_t r_last = rot_idx(idx, last_rotation, rot_scale, size);
        /* permutation */
        if (c->n_sets) {
            fe t = fe_sub2(&one, &c->permz[0][idx]);
            FOLD(fe_mul2(&t, &c->l0[idx]));
            const fe* zl = &c->permz[c->n_sets - 1][idx];
            t = fe_mul2(zl, zl);
            t = fe_sub2(&t, zl);
            FOLD(fe_mul2(&t, &c->l_last[idx]));
            for (size_t s = 1; s < c->n_sets; s++) {
                t = fe_sub2(&c->permz[s][idx], &c->permz[s - 1][r_last]);
                FOLD(fe_mul2(&t, &c->l0[idx]));
            }
            fe current_delta = fe_mul2(&delta_start, &beta_term);
            for (size_t s = 0; s < c->n_sets; s++) {
                size_t c0 = s * desc->chunk_len, c1 = c0 + desc->chunk_len;
                if (c1 > desc->n_perm_columns) c1 = desc->n_perm_columns;
                fe left = c->permz[s][r_next], right = c->permz[s][idx];
                for (size_t col = c0; col < c1; col++) {
                    uint32_t kind = desc->perm_column_kind[col], index = desc->perm_column_index[col];
                    fe** src = kind == DE_VAL_ADVICE ? c->advice : (kind == DE_VAL_FIXED ? c->fixed : c->instance);
                    const fe* val = &src[index][idx];
                    fe u = fe_mul2(&c->beta, &c->sigma[col][idx]);
                    u = fe_add2(val, &u);
                    u = fe_add2(&u, &c->gamma);
                    left = fe_mul2(&left, &u);
                    u = fe_add2(val, &current_delta);
                    u = fe_add2(&u, &c->gamma);
                    right = fe_mul2(&right, &u);
                    current_delta = fe_mul2(&current_delta, &c->delta);
                }
                t = fe_sub2(&left, &right);
                FOLD(fe_mul2(&t, &c->l_active[idx]));
            }
        }
        f_mul(F, &beta_term, &beta_term, &d->ext_omega);
        /* lookups */
        for (uint32_t n = 0; n < desc->n_lookups; n++) {
            fe table_value = graph_eval(c, &desc->lookups[n], inter, idx, rot_scale, size, &zero);
            fe *zc = c->lookup[n], *ac = c->lookup[desc->n_lookups + n], *sc = c->lookup[2 * desc->n_lookups + n];
            fe a_minus_s = fe_sub2(&ac[idx], &sc[idx]);
            fe t = fe_sub2(&one, &zc[idx]);
            FOLD(fe_mul2(&t, &c->l0[idx]));
            t = fe_mul2(&zc[idx], &zc[idx]);
            t = fe_sub2(&t, &zc[idx]);
            FOLD(fe_mul2(&t, &c->l_last[idx]));
            fe u = fe_add2(&ac[idx], &c->beta), w = fe_add2(&sc[idx], &c->gamma);
            t = fe_mul2(&zc[r_next], &u);
            t = fe_mul2(&t, &w);
            u = fe_mul2(&zc[idx], &table_value);
            t = fe_sub2(&t, &u);
            FOLD(fe_mul2(&t, &c->l_active[idx]));
            FOLD(fe_mul2(&a_minus_s, &c->l0[idx]));
            u = fe_sub2(&ac[idx], &ac[r_prev]);
            t = fe_mul2(&a_minus_s, &u);
            FOLD(fe_mul2(&t, &c->l_active[idx]));
        }
        c->out[idx] = value;
    }
    free(inter);
}
static fe** cosets_of(const orc_domain* d, const de_fr* const* polys, size_t count) {
    fe** out = (fe**)calloc(count ? count : 1, sizeof(fe*));
    for (size_t i = 0; i < count; i++) {
        out[i] = (fe*)malloc(sizeof(fe) * d->ext_n);
        orc_coeff_to_extended(d, (const u64*)polys[i], (u64*)out[i]);
    }
    return out;
}
static void free_cosets(fe** c, size_t count) {
    for (size_t i = 0; i < count; i++) free(c[i]);
    free(c);
}
static fe* indicator_coset(const orc_domain* d, size_t lo, size_t hi) {
    fe* lag = (fe*)calloc(d->n, sizeof(fe));
    for (size_t i = lo; i < hi; i++) lag[i] = FRF.r;
    orc_lagrange_to_coeff(d, (u64*)lag);
    fe* ext = (fe*)malloc(sizeof(fe) * d->ext_n);
    orc_coeff_to_extended(d, (const u64*)lag, (u64*)ext);
    free(lag);
    return ext;
}
typedef struct {
    const orc_domain* d;
    const de_pk_desc* desc; /* borrowed: must outlive the pk */
    fe **fixed, **sigma;
    fe *l0, *l_last, *l_active;
} orc_pk;
/* keygen_pk's resident part: fixed / sigma / l0 / l_last / l_active_row cosets, computed once */
orc_pk* orc_pk_new(const orc_domain* d, const de_pk_desc* desc) {
    orc_pk* pk = (orc_pk*)calloc(1, sizeof(*pk));
    pk->d = d;
    pk->desc = desc;
    pk->fixed = cosets_of(d, desc->fixed_coeff, desc->n_fixed);
    pk->sigma = cosets_of(d, desc->sigma_coeff, desc->n_perm_columns);
    const size_t bf = desc->blinding_factors;
    pk->l0 = indicator_coset(d, 0, 1);
    pk->l_last = indicator_coset(d, d->n - bf - 1, d->n - bf);
    fe* l_blind = indicator_coset(d, d->n - bf, d->n);
    pk->l_active = (fe*)malloc(sizeof(fe) * d->ext_n);
    for (size_t i = 0; i < d->ext_n; i++) {
        fe t;
        f_add(&FRF, &t, &pk->l_last[i], &l_blind[i]);
        f_sub(&FRF, &pk->l_active[i], &FRF.r, &t);
    }
    free(l_blind);
    return pk;
}
void orc_pk_free(orc_pk* pk) {
    if (!pk) return;
    free_cosets(pk->fixed, pk->desc->n_fixed);
    free_cosets(pk->sigma, pk->desc->n_perm_columns);
    free(pk->l0); free(pk->l_last); free(pk->l_active);
    free(pk);
}
int orc_evaluate_h(const orc_pk* pk, const de_fr* const* advice_coeff, const de_fr* const* instance_coeff,
                   const de_challenges* ch, const de_fr* const* perm_z_coeff, const de_fr* const* lookup_coeff, u64* h_ext_out) {
    const orc_domain* d = pk->d;
    const de_pk_desc* desc = pk->desc;
    evalh_ctx c;
    memset(&c, 0, sizeof(c));
    c.d = d;
    c.desc = desc;
    c.n_sets = desc->n_perm_columns ? (desc->n_perm_columns + desc->chunk_len - 1) / desc->chunk_len : 0;
    c.fixed = pk->fixed;
    c.sigma = pk->sigma;
    c.advice = cosets_of(d, advice_coeff, desc->n_advice);
    c.instance = cosets_of(d, instance_coeff, desc->n_instance);
    c.permz = cosets_of(d, perm_z_coeff, c.n_sets);
    c.lookup = cosets_of(d, lookup_coeff, 3 * (size_t)desc->n_lookups);
    c.l0 = pk->l0; c.l_last = pk->l_last; c.l_active = pk->l_active;
    memcpy(&c.y, &ch->y, 32); memcpy(&c.beta, &ch->beta, 32); memcpy(&c.gamma, &ch->gamma, 32); memcpy(&c.theta, &ch->theta, 32);
    memcpy(&c.delta, &desc->delta, 32);
    c.challenges = (const fe*)ch->challenges;
    c.out = (fe*)h_ext_out;
    parallel_for(d->ext_n, d->threads, evalh_job, &c);
    free_cosets(c.advice, desc->n_advice);
    free_cosets(c.instance, desc->n_instance); free_cosets(c.permz, c.n_sets); free_cosets(c.lookup, 3 * (size_t)desc->n_lookups);
    return 0;
}
