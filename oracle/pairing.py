"""BN254 optimal-ate pairing in plain Python integers.  TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py).

Needed by the restated verifier (oracle/pyprover.py: verify_proof ends with the KZG check
e(left, [s]_2) == e(right, [1]_2), halo2_proofs::poly::kzg::strategy / DualMSM::check, reached from the reference at
/root/reference/benches/delay_enc.rs:153-160).  The arithmetic lives in halo2curves::bn256 (un-vendored); the check only
depends on the pairing being bilinear and non-degenerate, so the textbook construction below (Fq12 as Fq[w]/(w^12 - 18 w^6 + 82),
Miller loop over 6x+2, plain final exponentiation) decides accept / reject exactly like the reference's.
Self-checked for bilinearity in tests/test_pyprover.py.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from pyoracle import FQ, FR

ATE_LOOP_COUNT = 29793968203157093288  # 6x + 2, x = 4965661367192848881
LOG_ATE = 63
# Fq12 modulus: w^12 = 18 w^6 - 82
_MC6, _MC0 = 18, -82


class Fq12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % FQ for x in c]

    @staticmethod
    def one():
        return Fq12([1] + [0] * 11)

    @staticmethod
    def zero():
        return Fq12([0] * 12)

    def __add__(self, o):
        return Fq12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return Fq12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return Fq12([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k: int):
        return Fq12([a * k for a in self.c])

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        a, b = self.c, o.c
        t = [0] * 23
        for i, ai in enumerate(a):
            if ai:
                for j, bj in enumerate(b):
                    t[i + j] += ai * bj
        for k in range(22, 11, -1):  # w^k = 18 w^(k-6) - 82 w^(k-12)
            v = t[k]
            if v:
                t[k - 6] += _MC6 * v
                t[k - 12] += _MC0 * v
        return Fq12(t[:12])

    def __pow__(self, e: int):
        r, b = Fq12.one(), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def inv(self):
        # a^(q^12 - 2) would be very slow; use the extended Euclid on polynomials instead
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [82, 0, 0, 0, 0, 0, FQ - 18, 0, 0, 0, 0, 0, 1]

        def deg(p):
            d = len(p) - 1
            while d and p[d] == 0:
                d -= 1
            return d

        def poly_div(a, b):
            da, db = deg(a), deg(b)
            temp = list(a)
            o = [0] * len(a)
            for i in range(da - db, -1, -1):
                o[i] = (o[i] + temp[db + i] * pow(b[db], -1, FQ)) % FQ
                for c in range(db + 1):
                    temp[c + i] = (temp[c + i] - o[c]) % FQ
            return o[: deg(o) + 1]

        while deg(low):
            r = poly_div(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % FQ for x in nm]
            new = [x % FQ for x in new]
            lm, low, hm, high = nm, new, lm, low
        li = pow(low[0], -1, FQ)
        return Fq12([x * li for x in lm[:12]])

    def __truediv__(self, o):
        return self * o.inv()


# ---- Fq2 = Fq[u] / (u^2 + 1) for G2 point arithmetic ------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % FQ, (a[1] + b[1]) % FQ)
def f2_sub(a, b): return ((a[0] - b[0]) % FQ, (a[1] - b[1]) % FQ)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % FQ, (a[0] * b[1] + a[1] * b[0]) % FQ)
def f2_neg(a): return ((-a[0]) % FQ, (-a[1]) % FQ)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, FQ)
    return (a[0] * d % FQ, (-a[1]) * d % FQ)


G2_B = f2_mul((3, 0), f2_inv((9, 1)))  # twist curve y^2 = x^3 + 3 / (9 + u)
# halo2curves::bn256::G2Affine::generator() (the alt_bn128 / EIP-197 generator)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))

G2Point = Optional[Tuple[Tuple[int, int], Tuple[int, int]]]


def g2_is_on_curve(p: G2Point) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), G2_B)


def g2_double(p: G2Point) -> G2Point:
    if p is None:
        return None
    x, y = p
    if y == (0, 0):
        return None
    lam = f2_mul(f2_mul((3, 0), f2_mul(x, x)), f2_inv(f2_mul((2, 0), y)))
    x3 = f2_sub(f2_mul(lam, lam), f2_add(x, x))
    return (x3, f2_sub(f2_mul(lam, f2_sub(x, x3)), y))


def g2_add(p: G2Point, q: G2Point) -> G2Point:
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        return g2_double(p) if p[1] == q[1] else None
    lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), p[0]), q[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(p[0], x3)), p[1]))


def g2_mul(p: G2Point, k: int) -> G2Point:
    k %= FR
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_double(acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


# ---- pairing ----------------------------------------------------------------------------------------------------------
def _w_pow(k: int) -> Fq12:
    c = [0] * 12
    c[k] = 1
    return Fq12(c)


_W2, _W3 = _w_pow(2), _w_pow(3)


def _twist(p):
    """G2 point over Fq2 -> point on y^2 = x^3 + 3 over Fq12 (u = w^6 - 9)."""
    (x0, x1), (y0, y1) = p
    nx = Fq12([x0 - 9 * x1, 0, 0, 0, 0, 0, x1, 0, 0, 0, 0, 0])
    ny = Fq12([y0 - 9 * y1, 0, 0, 0, 0, 0, y1, 0, 0, 0, 0, 0])
    return (nx * _W2, ny * _W3)


def _embed(p):
    return (Fq12([p[0]] + [0] * 11), Fq12([p[1]] + [0] * 11))


def _pt_double(p):
    x, y = p
    lam = (x * x).scale(3) / y.scale(2)
    nx = lam * lam - x.scale(2)
    return (nx, lam * (x - nx) - y)


def _pt_add(p, q):
    if p[0] == q[0]:
        return _pt_double(p) if p[1] == q[1] else None
    lam = (q[1] - p[1]) / (q[0] - p[0])
    nx = lam * lam - p[0] - q[0]
    return (nx, lam * (p[0] - nx) - p[1])


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) / y1.scale(2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q_g2: G2Point, p_g1) -> Fq12:
    """f_{6x+2,Q}(P) with the two Frobenius line corrections, WITHOUT the final exponentiation."""
    if q_g2 is None or p_g1 is None:
        return Fq12.one()
    Q, P = _twist(q_g2), _embed(p_g1)
    R, f = Q, Fq12.one()
    for i in range(LOG_ATE, -1, -1):
        f = f * f * _linefunc(R, R, P)
        R = _pt_double(R)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(R, Q, P)
            R = _pt_add(R, Q)
    Q1 = (Q[0] ** FQ, Q[1] ** FQ)
    nQ2 = (Q1[0] ** FQ, -(Q1[1] ** FQ))
    f = f * _linefunc(R, Q1, P)
    R = _pt_add(R, Q1)
    f = f * _linefunc(R, nQ2, P)
    return f


def final_exponentiation(f: Fq12) -> Fq12:
    return f ** ((FQ ** 12 - 1) // FR)


def pairing(q_g2: G2Point, p_g1) -> Fq12:
    return final_exponentiation(miller_loop(q_g2, p_g1))


def pairing_check(pairs: List[Tuple[G2Point, Optional[Tuple[int, int]]]]) -> bool:
    """prod e(P_i, Q_i) == 1 (one shared final exponentiation, as multi_miller_loop + final_exponentiation do)."""
    f = Fq12.one()
    for q, p in pairs:
        f = f * miller_loop(q, p)
    return final_exponentiation(f) == Fq12.one()
