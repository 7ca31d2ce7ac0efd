// ec_quad.cuh — XYZZ point addition and doubling by FOUR lanes per point, for the latency-bound ends of the bucket reduction.
//
// A general XYZZ addition is 14 field multiplications, a doubling 9; one thread runs them back to back (~0.4 us each: the
// carry chains of a 256-bit Montgomery product leave a lone warp waiting on itself), so a tree level costs ~6 us and the 15
// doublings of the final fold ~60 us whatever the occupancy.  The data flow is much shallower than that: the multiplications of
// "add-2008-s" form 6 rounds of independent products, those of "dbl-2008-s-1" 4.  Here the four lanes of a quad hold one
// coordinate each (role = lane & 3: 0 X, 1 Y, 2 ZZ, 3 ZZZ), every round is ONE mul() executed by all four lanes on operands
// fetched from their neighbours with shuffles, and a point operation costs 6 (4) multiplication latencies instead of 14 (9).
// The formulas, and therefore the points, are those of ec.cuh; exceptional cases (identity operands, P + P, P - P) are
// resolved by per-quad selects, the doubling inside an addition by a warp-uniform branch.
//
// All 32 lanes of the warp must call these functions together (the shuffles use the full mask); lanes of a quad that has
// nothing to do pass identities (zz = 0) and ignore the result.
#pragma once
#include "ec.cuh"

namespace de {

// the Fq held by the lane with role `src` of this lane's quad
DE_D Fq quad_fetch(const Fq& v, unsigned int src) {
    const unsigned int from = (threadIdx.x & 28u) | src;  // lane index within the warp: quad base + role
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], from);
    return r;
}
DE_D bool quad_flag(bool f, unsigned int src) { return __shfl_sync(0xffffffffu, f ? 1 : 0, (threadIdx.x & 28u) | src) != 0; }

// one coordinate of an XYZZ point in shared / global memory: coordinate `role` of *p
DE_D Fq quad_load(const XYZZ* p, unsigned int role) { return load(&p->x + role); }
DE_D void quad_store(XYZZ* p, unsigned int role, const Fq& c) { store(&p->x + role, c); }

// c <- coordinate `role` of 2 * P, P given by its coordinates across the quad.  4 multiplication rounds.
DE_D Fq quad_dbl(const Fq& c, unsigned int role) {
    const Fq t = role == 1 ? dbl(c) : c;                 // role 1: U = 2 Y
    const Fq s1 = sqr(t);                                 // role 0: X^2, role 1: V = U^2
    const Fq v = quad_fetch(s1, 1);
    const Fq r2 = mul(t, v);                              // role 0: S = X V, role 1: W = U V, role 2: ZZ3 = ZZ V
    const Fq w = quad_fetch(r2, 1);
    const Fq m = add(dbl(s1), s1);                        // role 0: M = 3 X^2
    const Fq r3 = mul(role == 0 ? m : w, role == 0 ? m : c);  // role 0: M^2, role 1: W Y, role 3: ZZZ3 = W ZZZ
    const Fq x3 = sub(r3, dbl(r2));                       // role 0: X3 = M^2 - 2 S
    const Fq m0 = quad_fetch(m, 0), s0 = quad_fetch(r2, 0), x30 = quad_fetch(x3, 0);
    const Fq y3 = sub(mul(m0, sub(s0, x30)), r3);         // role 1: Y3 = M (S - X3) - W Y
    // an identity (ZZ = 0) stays one: ZZ3 = 0 * V
    return role == 0 ? x3 : role == 1 ? y3 : role == 2 ? r2 : r3;
}

// c <- coordinate `role` of A + B (a, b: this lane's coordinate of each).  6 multiplication rounds.
DE_D Fq quad_add(const Fq& a, const Fq& b, unsigned int role) {
    const bool a_inf = quad_flag(a.is_zero(), 2), b_inf = quad_flag(b.is_zero(), 2);
    const unsigned int hi = role | 2u;                    // roles 0, 2 -> 2 (ZZ); roles 1, 3 -> 3 (ZZZ)
    const Fq m1 = mul(a, quad_fetch(b, hi));              // U1 = X1 ZZ2 | S1 = Y1 ZZZ2 | ZZ1 ZZ2 | ZZZ1 ZZZ2
    const Fq m2 = mul(b, quad_fetch(a, hi));              // U2 = X2 ZZ1 | S2 = Y2 ZZZ1 | -       | -
    const Fq d = sub(m2, m1);                             // P            | R
    const bool p_zero = quad_flag(d.is_zero(), 0), r_zero = quad_flag(d.is_zero(), 1);
    const Fq sq = sqr(d);                                 // PP           | R^2
    const Fq pp = quad_fetch(sq, 0), p = quad_fetch(d, 0);
    const Fq r4 = mul((role & 1u) ? p : m1, pp);          // Q = U1 PP    | PPP = P PP   | ZZ3 = ZZ1 ZZ2 PP | PPP
    const Fq r5 = mul(m1, r4);                            // -            | S1 PPP       | -                | ZZZ3 = ZZZ1 ZZZ2 PPP
    const Fq rr = quad_fetch(sq, 1), ppp = quad_fetch(r4, 1);
    const Fq x3 = sub(sub(rr, ppp), dbl(r4));             // role 0: X3 = R^2 - PPP - 2 Q
    const Fq q0 = quad_fetch(r4, 0), x30 = quad_fetch(x3, 0);
    const Fq y3 = sub(mul(d, sub(q0, x30)), r5);          // role 1: Y3 = R (Q - X3) - S1 PPP
    Fq out = role == 0 ? x3 : role == 1 ? y3 : role == 2 ? r4 : r5;
    const bool regular = !a_inf && !b_inf;
    // P + P: rare (equal partial sums need equal multisets of bases), so the doubling sits behind a warp-uniform branch
    if (__any_sync(0xffffffffu, regular && p_zero && r_zero)) {
        const Fq twice = quad_dbl(a, role);
        if (regular && p_zero && r_zero) out = twice;
    }
    if (regular && p_zero && !r_zero) out = Fq::zero();  // P - P
    if (b_inf) out = a;
    else if (a_inf) out = b;
    return out;
}

}  // namespace de
