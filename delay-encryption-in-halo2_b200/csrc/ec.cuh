// ec.cuh — BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels.
//
// Replaces halo2curves::bn256::{G1Affine, G1} addition/doubling as used by best_multiexp's bucket loop
// (reference call chain benches/delay_enc.rs:123 -> create_proof -> commit_lagrange).
// Buckets are kept in extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): a mixed addition
// costs 8M + 2S and never needs an inversion.  All exceptional cases (identity, P + P, P - P) are handled exactly,
// because synthetic bases such as (i+1)*G make them reachable.
#pragma once
#include "field.cuh"

namespace de {

struct alignas(16) Affine {
    Fq x, y;  // identity = (0, 0), as halo2curves encodes it
};
struct alignas(16) Jac {
    Fq x, y, z;  // identity: z = 0
};
struct alignas(16) XYZZ {
    Fq x, y, zz, zzz;  // identity: zz = 0
};

DE_D bool is_identity(const Affine& p) { return p.x.is_zero() && p.y.is_zero(); }
DE_D bool is_identity(const XYZZ& p) { return p.zz.is_zero(); }

DE_D XYZZ xyzz_identity() {
    XYZZ r;
    r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
    return r;
}
DE_D XYZZ xyzz_from_affine(const Affine& p) {
    XYZZ r;
    if (is_identity(p)) return xyzz_identity();
    r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one();
    return r;
}
// 2 * (affine point), "mdbl-2008-s-1" with a = 0
DE_D XYZZ xyzz_dbl_affine(const Affine& p) {
    XYZZ r;
    Fq u = dbl(p.y);
    Fq v = sqr(u);
    Fq w = mul(u, v);
    Fq s = mul(p.x, v);
    Fq x2 = sqr(p.x);
    Fq m = add(dbl(x2), x2);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}
// 2 * P, "dbl-2008-s-1" with a = 0
DE_D XYZZ xyzz_dbl(const XYZZ& p) {
    if (is_identity(p)) return p;
    XYZZ r;
    Fq u = dbl(p.y);
    Fq v = sqr(u);
    Fq w = mul(u, v);
    Fq s = mul(p.x, v);
    Fq x2 = sqr(p.x);
    Fq m = add(dbl(x2), x2);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = mul(v, p.zz);
    r.zzz = mul(w, p.zzz);
    return r;
}
// acc += q (affine, already sign-adjusted), "madd-2008-s"
DE_D void xyzz_madd(XYZZ& acc, const Affine& q) {
    if (is_identity(q)) return;
    if (is_identity(acc)) {
        acc = xyzz_from_affine(q);
        return;
    }
    Fq u2 = mul(q.x, acc.zz);
    Fq s2 = mul(q.y, acc.zzz);
    Fq p = sub(u2, acc.x);
    Fq r = sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl_affine(q);
        else acc = xyzz_identity();
        return;
    }
    Fq pp = sqr(p);
    Fq ppp = mul(p, pp);
    Fq qq = mul(acc.x, pp);
    Fq x3 = sub(sub(sqr(r), ppp), dbl(qq));
    Fq y3 = sub(mul(r, sub(qq, x3)), mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = mul(acc.zz, pp);
    acc.zzz = mul(acc.zzz, ppp);
}
// acc += o, "add-2008-s"
DE_D void xyzz_add(XYZZ& acc, const XYZZ& o) {
    if (is_identity(o)) return;
    if (is_identity(acc)) {
        acc = o;
        return;
    }
    Fq u1 = mul(acc.x, o.zz);
    Fq u2 = mul(o.x, acc.zz);
    Fq s1 = mul(acc.y, o.zzz);
    Fq s2 = mul(o.y, acc.zzz);
    Fq p = sub(u2, u1);
    Fq r = sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_dbl(acc);
        else acc = xyzz_identity();
        return;
    }
    Fq pp = sqr(p);
    Fq ppp = mul(p, pp);
    Fq qq = mul(u1, pp);
    Fq x3 = sub(sub(sqr(r), ppp), dbl(qq));
    Fq y3 = sub(mul(r, sub(qq, x3)), mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = mul(mul(acc.zz, o.zz), pp);
    acc.zzz = mul(mul(acc.zzz, o.zzz), ppp);
}
// XYZZ -> Jacobian without inversion: (X*ZZ, Y*ZZZ, ZZ) since Z = ZZ gives Z^2 = ZZ^2, Z^3 = ZZZ^2
DE_D Jac xyzz_to_jac(const XYZZ& p) {
    Jac r;
    if (is_identity(p)) {
        r.x = Fq::zero(); r.y = Fq::zero(); r.z = Fq::zero();
        return r;
    }
    r.x = mul(p.x, p.zz);
    r.y = mul(p.y, p.zzz);
    r.z = p.zz;
    return r;
}

DE_D Affine load_affine(const Affine* p) {
    Affine r;
    r.x = load(&p->x);
    r.y = load(&p->y);
    return r;
}
DE_D void store_affine(Affine* p, const Affine& v) {
    store(&p->x, v.x);
    store(&p->y, v.y);
}
DE_D XYZZ load_xyzz(const XYZZ* p) {
    XYZZ r;
    r.x = load(&p->x); r.y = load(&p->y); r.zz = load(&p->zz); r.zzz = load(&p->zzz);
    return r;
}
DE_D void store_xyzz(XYZZ* p, const XYZZ& v) {
    store(&p->x, v.x); store(&p->y, v.y); store(&p->zz, v.zz); store(&p->zzz, v.zzz);
}

}  // namespace de
