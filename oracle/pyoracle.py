"""CPU oracle (Python big-int restatement) for the delay-encryption-in-halo2 proving hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

PARITY UNPINNED at the MSM / NTT boundary: the arithmetic of this path lives in the un-vendored
dependency halo2_proofs (git tag v2023_04_20, /root/reference/Cargo.toml:17) and, transitively,
halo2curves 0.3.x.  Neither source nor a Rust toolchain exists here, and no reference test holds a
golden commitment / NTT vector.  What IS pinned:
  * Fr arithmetic, canonical encoding and inversion, by the reference's own Poseidon known-answer
    vectors (/root/reference/src/poseidon/permutation.rs:154-158,190-196) reproduced in
    `poseidon_kat()` from a restatement of its Grain generator (/root/reference/src/poseidon/grain.rs:12-69).
  * MSM and NTT results are mathematically unique (exact integers, canonical representatives), so any
    correct implementation equals the reference's result; the convention-bearing constants
    (ROOT_OF_UNITY, ZETA, DELTA, Montgomery R) are checked numerically in tests/test_oracle.py.

Everything here is written for clarity, not speed; oracle/oracle.c is the fast restatement and is itself
cross-checked against this file.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------------------------
# BN254 constants (halo2curves::bn256)
# ----------------------------------------------------------------------------------------------
FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R256 = 1 << 256
FR_S = 28                       # 2-adicity of r-1
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (FR - 1) >> FR_S, FR)
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, FR)
G1_B = 3
G1_GEN = (1, 2)


def inv(a: int, p: int) -> int:
    return pow(a, -1, p)


def to_mont(a: int, p: int) -> int:
    return (a * R256) % p


def from_mont(a: int, p: int) -> int:
    return (a * inv(R256, p)) % p


def limbs64(a: int) -> List[int]:
    return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs64(l: Sequence[int]) -> int:
    return sum(int(x) << (64 * i) for i, x in enumerate(l))


def from_uniform_bytes(b: bytes, p: int = FR) -> int:
    """ff::FromUniformBytes<64>: 512-bit little-endian integer reduced mod p."""
    assert len(b) == 64
    return int.from_bytes(b, "little") % p


# ----------------------------------------------------------------------------------------------
# PRNG shared with the C oracle and the tests: SplitMix64-seeded xoshiro256** 
# ----------------------------------------------------------------------------------------------
M64 = (1 << 64) - 1


class Xoshiro:
    def __init__(self, seed: int):
        s = seed & M64
        self.s = []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & M64
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            self.s.append(z ^ (z >> 31))

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & M64

    def next_u64(self) -> int:
        s = self.s
        result = (self._rotl((s[1] * 5) & M64, 7) * 9) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return result

    def uniform_fr(self) -> int:
        """'uniform Fr' = 512 random bits mod r (eight u64 draws, little-endian)."""
        v = 0
        for i in range(8):
            v |= self.next_u64() << (64 * i)
        return v % FR


# ----------------------------------------------------------------------------------------------
# Poseidon known-answer pin (restates /root/reference/src/poseidon/grain.rs and the naive permutation
# SpecRef::permute at /root/reference/src/poseidon/permutation.rs:59-80)
# ----------------------------------------------------------------------------------------------
class _Grain:
    def __init__(self, t: int, r_f: int, r_p: int):
        bits: List[int] = []

        def app(n, v):
            for i in range(n - 1, -1, -1):
                bits.append((v >> i) & 1)

        app(2, 1)
        app(4, 0)
        app(12, 254)
        app(12, t)
        app(10, r_f)
        app(10, r_p)
        app(30, (1 << 30) - 1)
        assert len(bits) == 80
        self.b = bits
        for _ in range(160):
            self._new_bit()

    def _new_bit(self) -> int:
        b = self.b
        nb = b[0] ^ b[62] ^ b[51] ^ b[38] ^ b[23] ^ b[13]
        b.pop(0)
        b.append(nb)
        return nb

    def _next(self) -> int:
        while not self._new_bit():
            self._new_bit()
        return self._new_bit()

    def _take254(self) -> int:
        v = 0
        for i in range(254):
            v |= self._next() << (253 - i)
        return v

    def next_field_element(self) -> int:
        while True:
            v = self._take254()
            if v < FR:
                return v

    def next_field_element_without_rejection(self) -> int:
        return self._take254() % FR


def poseidon_params(t: int, r_f: int, r_p: int):
    g = _Grain(t, r_f, r_p)
    constants = [[g.next_field_element() for _ in range(t)] for _ in range(r_f + r_p)]
    xs = [g.next_field_element_without_rejection() for _ in range(t)]
    ys = [g.next_field_element_without_rejection() for _ in range(t)]
    mds = [[inv((x + y) % FR, FR) for y in ys] for x in xs]
    return constants, mds


def poseidon_permute_ref(state: List[int], t: int, r_f: int, r_p: int) -> List[int]:
    constants, mds = poseidon_params(t, r_f, r_p)
    half = r_f // 2

    def apply_mds(s):
        return [sum(mds[i][j] * s[j] for j in range(t)) % FR for i in range(t)]

    s = list(state)
    for rnd in range(r_f + r_p):
        s = [(a + c) % FR for a, c in zip(s, constants[rnd])]
        if rnd < half or rnd >= half + r_p:
            s = [pow(a, 5, FR) for a in s]
        else:
            s[0] = pow(s[0], 5, FR)
        s = apply_mds(s)
    return s


POSEIDON_KAT_T3 = [  # /root/reference/src/poseidon/permutation.rs:154-158
    7853200120776062878684798364095072458815029376092732009249414926327459813530,
    7142104613055408817911962100316808866448378443474503659992478482890339429929,
    6549537674122432311777789598043107870002137484850126429160507761192163713804,
]
POSEIDON_KAT_T5 = [  # /root/reference/src/poseidon/permutation.rs:190-196
    18821383157269793795438455681495246036402687001665670618754263018637548127333,
    7817711165059374331357136443537800893307845083525445872661165200086166013245,
    16733335996448830230979566039396561240864200624113062088822991822580465420551,
    6644334865470350789317807668685953492649391266180911382577082600917830417726,
    3372108894677221197912083238087960099443657816445944159266857514496320565191,
]


def poseidon_kat() -> bool:
    ok3 = poseidon_permute_ref([0, 1, 2], 3, 8, 57) == POSEIDON_KAT_T3
    ok5 = poseidon_permute_ref([0, 1, 2, 3, 4], 5, 8, 60) == POSEIDON_KAT_T5
    return ok3 and ok5


# ----------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 3 over Fq.  Affine points are (x, y) or None (identity).
# ----------------------------------------------------------------------------------------------
Affine = Optional[Tuple[int, int]]


def g1_is_on_curve(p: Affine) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - G1_B) % FQ == 0


def g1_neg(p: Affine) -> Affine:
    return None if p is None else (p[0], (-p[1]) % FQ)


def g1_add(p: Affine, q: Affine) -> Affine:
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % FQ == 0:
            return None
        lam = 3 * x1 * x1 * inv(2 * y1, FQ) % FQ
    else:
        lam = (y2 - y1) * inv((x2 - x1) % FQ, FQ) % FQ
    x3 = (lam * lam - x1 - x2) % FQ
    y3 = (lam * (x1 - x3) - y1) % FQ
    return (x3, y3)


# Jacobian arithmetic for speed inside the Python oracle (X, Y, Z); identity has Z == 0.
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % FQ
    B = Y * Y % FQ
    C = B * B % FQ
    D = 2 * ((X + B) * (X + B) - A - C) % FQ
    E = 3 * A % FQ
    F = E * E % FQ
    X3 = (F - 2 * D) % FQ
    Y3 = (E * (D - X3) - 8 * C) % FQ
    Z3 = 2 * Y * Z % FQ
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % FQ
    Z2Z2 = Z2 * Z2 % FQ
    U1 = X1 * Z2Z2 % FQ
    U2 = X2 * Z1Z1 % FQ
    S1 = Y1 * Z2 * Z2Z2 % FQ
    S2 = Y2 * Z1 * Z1Z1 % FQ
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % FQ
    Rr = (S2 - S1) % FQ
    HH = H * H % FQ
    HHH = H * HH % FQ
    V = U1 * HH % FQ
    X3 = (Rr * Rr - HHH - 2 * V) % FQ
    Y3 = (Rr * (V - X3) - S1 * HHH) % FQ
    Z3 = Z1 * Z2 * H % FQ
    return (X3, Y3, Z3)


def jac_to_affine(P) -> Affine:
    X, Y, Z = P
    if Z % FQ == 0:
        return None
    zi = inv(Z, FQ)
    zi2 = zi * zi % FQ
    return (X * zi2 % FQ, Y * zi2 * zi % FQ)


def _to_jac(p: Affine):
    return (1, 1, 0) if p is None else (p[0], p[1], 1)


def g1_mul(p: Affine, k: int) -> Affine:
    k %= FR
    acc = (1, 1, 0)
    base = _to_jac(p)
    while k:
        if k & 1:
            acc = _jac_add(acc, base)
        base = _jac_double(base)
        k >>= 1
    return jac_to_affine(acc)


def msm_naive(scalars: Sequence[int], bases: Sequence[Affine]) -> Affine:
    """Definition of best_multiexp: sum_i scalars[i] * bases[i]."""
    assert len(scalars) == len(bases)
    acc = (1, 1, 0)
    for s, b in zip(scalars, bases):
        if s % FR == 0 or b is None:
            continue
        acc = _jac_add(acc, _to_jac(g1_mul(b, s)))
    return jac_to_affine(acc)


def msm_pippenger(scalars: Sequence[int], bases: Sequence[Affine]) -> Affine:
    """multiexp_serial restated: unsigned c-bit windows, c = ceil(ln n),
    segments = 256/c + 1 processed high to low with c doublings between, running-sum bucket reduction."""
    n = len(scalars)
    assert n == len(bases)
    if n == 0:
        return None
    c = 1 if n < 4 else (3 if n < 32 else int(math.ceil(math.log(n))))
    segments = 256 // c + 1
    acc = (1, 1, 0)
    for seg in range(segments - 1, -1, -1):
        for _ in range(c):
            acc = _jac_double(acc)
        buckets = [(1, 1, 0)] * ((1 << c) - 1)
        for s, b in zip(scalars, bases):
            d = ((s % FR) >> (seg * c)) & ((1 << c) - 1)
            if d and b is not None:
                buckets[d - 1] = _jac_add(buckets[d - 1], _to_jac(b))
        running = (1, 1, 0)
        for bk in reversed(buckets):
            running = _jac_add(running, bk)
            acc = _jac_add(acc, running)
    return jac_to_affine(acc)


# ----------------------------------------------------------------------------------------------
# NTT (best_fft): natural order in and out, A[j] = sum_i a[i] * omega^(i*j)
# ----------------------------------------------------------------------------------------------
def dft_naive(a: Sequence[int], omega: int) -> List[int]:
    n = len(a)
    pw = [pow(omega, e, FR) for e in range(n)]
    return [sum(a[i] * pw[(i * j) % n] for i in range(n)) % FR for j in range(n)]


def _bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def best_fft(a: Sequence[int], omega: int, log_n: int) -> List[int]:
    """Radix-2 DIT with bit-reversal permutation first, as best_fft does."""
    n = 1 << log_n
    assert len(a) == n
    a = list(a)
    for k in range(n):
        rk = _bitrev(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % FR
    m = 1
    while m < n:
        step = n // (2 * m)
        for start in range(0, n, 2 * m):
            for j in range(m):
                t = a[start + j + m] * tw[j * step] % FR
                u = a[start + j]
                a[start + j] = (u + t) % FR
                a[start + j + m] = (u - t) % FR
        m *= 2
    return a


# ----------------------------------------------------------------------------------------------
# EvaluationDomain
# ----------------------------------------------------------------------------------------------
class EvaluationDomain:
    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek
        self.extended_n = 1 << ek
        self.extended_omega = pow(FR_ROOT_OF_UNITY, 1 << (FR_S - ek), FR)
        self.extended_omega_inv = inv(self.extended_omega, FR)
        self.omega = pow(self.extended_omega, 1 << (ek - k), FR)
        self.omega_inv = inv(self.omega, FR)
        self.g_coset = FR_ZETA
        self.g_coset_inv = FR_ZETA * FR_ZETA % FR
        self.ifft_divisor = inv(self.n, FR)
        self.extended_ifft_divisor = inv(self.extended_n, FR)
        self.t_evaluations = [
            inv((pow(FR_ZETA * pow(self.extended_omega, i, FR), self.n, FR) - 1) % FR, FR)
            for i in range(1 << (ek - k))
        ]

    def lagrange_to_coeff(self, a):
        return [x * self.ifft_divisor % FR for x in best_fft(a, self.omega_inv, self.k)]

    def coeff_to_lagrange(self, a):
        return best_fft(a, self.omega, self.k)

    def coeff_to_extended(self, a):
        assert len(a) == self.n
        z = [1, self.g_coset, self.g_coset_inv]
        b = [x * z[i % 3] % FR for i, x in enumerate(a)] + [0] * (self.extended_n - self.n)
        return best_fft(b, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a):
        assert len(a) == self.extended_n
        b = best_fft(a, self.extended_omega_inv, self.extended_k)
        z = [1, self.g_coset_inv, self.g_coset]
        b = [x * self.extended_ifft_divisor % FR * z[i % 3] % FR for i, x in enumerate(b)]
        return b[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a):
        assert len(a) == self.extended_n
        m = len(self.t_evaluations)
        return [x * self.t_evaluations[i % m] % FR for i, x in enumerate(a)]

    def rotate_extended(self, a, rot: int):
        sh = (rot * (1 << (self.extended_k - self.k))) % self.extended_n
        return list(a[sh:]) + list(a[:sh])


def eval_poly(coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % FR
    return acc


# ----------------------------------------------------------------------------------------------
# evaluate_h by the DEFINITION: every polynomial is evaluated with Horner at the coset point
# x = zeta * extended_omega^idx (rotations multiply x by omega^rot); expressions are walked as trees, not as the compiled
# GraphEvaluator program.  Independent of the NTT, of rotation index arithmetic and of the graph compiler.
# Expressions are the plain tuples ("const", v) / ("fixed"|"advice"|"instance", col, rot) / ("challenge", i) /
# ("neg", e) / ("sum", a, b) / ("prod", a, b) / ("scaled", e, v).
# ----------------------------------------------------------------------------------------------
def evaluate_h_definition(dom: "EvaluationDomain", gates, lookups, perm_columns, chunk_len, blinding_factors, fixed, sigma,
                          advice, instance, permz, lookup_polys, y, beta, gamma, theta, challenges=(), rows=None):
    """All polynomials are coefficient lists of canonical ints.  perm_columns: [(kind, index)] with kind 2=fixed,
    3=advice, 4=instance.  lookup_polys: [all z | all a' | all s'].  Returns {idx: h(idx)} for idx in rows (default all)."""
    n = dom.n

    def lag_poly(lo, hi):
        return dom.lagrange_to_coeff([1 if lo <= i < hi else 0 for i in range(n)])

    l0p = lag_poly(0, 1)
    l_lastp = lag_poly(n - blinding_factors - 1, n - blinding_factors)
    l_blindp = lag_poly(n - blinding_factors, n)
    cols = {2: fixed, 3: advice, 4: instance}
    n_sets = -(-len(perm_columns) // chunk_len) if perm_columns else 0
    out = {}
    for idx in (rows if rows is not None else range(dom.extended_n)):
        x = FR_ZETA * pow(dom.extended_omega, idx, FR) % FR

        def at(poly, rot=0):
            return eval_poly(poly, x * pow(dom.omega, rot, FR) % FR)

        def ev(e):
            t = e[0]
            if t == "const":
                return e[1] % FR
            if t == "fixed":
                return at(fixed[e[1]], e[2])
            if t == "advice":
                return at(advice[e[1]], e[2])
            if t == "instance":
                return at(instance[e[1]], e[2])
            if t == "challenge":
                return challenges[e[1]]
            if t == "neg":
                return (-ev(e[1])) % FR
            if t == "sum":
                return (ev(e[1]) + ev(e[2])) % FR
            if t == "prod":
                return ev(e[1]) * ev(e[2]) % FR
            if t == "scaled":
                return ev(e[1]) * e[2] % FR
            raise ValueError(t)

        l0, l_last = at(l0p), at(l_lastp)
        l_active = (1 - (l_last + at(l_blindp))) % FR
        value = 0
        for g in gates:
            value = (value * y + ev(g)) % FR
        if n_sets:
            last_rot = -(blinding_factors + 1)
            value = (value * y + (1 - at(permz[0])) * l0) % FR
            zl = at(permz[-1])
            value = (value * y + (zl * zl - zl) * l_last) % FR
            for s in range(1, n_sets):
                value = (value * y + (at(permz[s]) - at(permz[s - 1], last_rot)) * l0) % FR
            col_no = 0
            for s in range(n_sets):
                left, right = at(permz[s], 1), at(permz[s])
                for kind, index in perm_columns[s * chunk_len:(s + 1) * chunk_len]:
                    v = at(cols[kind][index])
                    left = left * (v + beta * at(sigma[col_no]) + gamma) % FR
                    right = right * (v + pow(FR_DELTA, col_no, FR) * beta % FR * x + gamma) % FR
                    col_no += 1
                value = (value * y + (left - right) * l_active) % FR
        for li, (inp, tab) in enumerate(lookups):
            nl = len(lookups)
            z, a, s = lookup_polys[li], lookup_polys[nl + li], lookup_polys[2 * nl + li]
            ci = 0
            for e in inp:
                ci = (ci * theta + ev(e)) % FR
            ct = 0
            for e in tab:
                ct = (ct * theta + ev(e)) % FR
            zv, av, sv = at(z), at(a), at(s)
            value = (value * y + (1 - zv) * l0) % FR
            value = (value * y + (zv * zv - zv) * l_last) % FR
            value = (value * y + (at(z, 1) * (av + beta) * (sv + gamma) - zv * (ci + beta) * (ct + gamma)) * l_active) % FR
            value = (value * y + (av - sv) * l0) % FR
            value = (value * y + (av - sv) * (av - at(a, -1)) * l_active) % FR
        out[idx] = value % FR
    return out
