"""Synthetic inputs for the BASELINE.json configurations: seeded, reproducible, generated without any
GPU or checker code so that bench.py's timed arm does not depend on test infrastructure.

  * uniform Fr columns: any 4-limb value below r is the Montgomery image of a uniformly distributed field element, so uniform
    columns are drawn directly as limbs (rejection on the top limb) with numpy's PCG64.
  * witness-like columns ("W"): rows >= used are zero except the last 6 (blinding, uniform); of the used rows 45 % < 2^8,
    35 % < 2^64, 10 % < 2^134, 10 % uniform — the value mix of the big-integer / range-check rows of the RSA circuit.
  * bases: P_i = (i + 1) * G on BN254 G1 (G = (1, 2)), affine Montgomery — a chain of Jacobian additions normalised with one
    batched inversion, in Python integers.
"""
from __future__ import annotations

import numpy as np

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_R = 1 << 256
_TOP_FR = FR >> 192


def uniform_fr(seed: int, n: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    out = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    # top limb: uniform below floor(r / 2^192) keeps every value < r (the excluded sliver has measure < 2^-60)
    out[:, 3] = rng.integers(0, _TOP_FR, size=n, dtype=np.uint64)
    return out


def _mont_limbs(v: int, p: int):
    m = v * _R % p
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def witness_fr(seed: int, n: int, used: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros((n, 4), dtype=np.uint64)
    sel = rng.integers(0, 100, size=used)
    small = rng.integers(0, 256, size=used, dtype=np.uint64)
    limb = rng.integers(0, 1 << 63, size=(used, 3), dtype=np.uint64)
    uni = uniform_fr(seed ^ 0x5555, used)
    for i in range(used):
        s = sel[i]
        if s < 45:
            v = int(small[i])
        elif s < 80:
            v = int(limb[i, 0]) * 2 + 1
        elif s < 90:
            v = int(limb[i, 0]) | (int(limb[i, 1]) << 64) | ((int(limb[i, 2]) & 0x3F) << 128)
        else:
            out[i] = uni[i]
            continue
        out[i] = _mont_limbs(v, FR)
    out[n - 6:] = uniform_fr(seed ^ 0xAAAA, 6)
    return out


def gen_bases(n: int, start: int = 0) -> np.ndarray:
    """P_i = (start + i + 1) * G, affine, Montgomery limbs, shape (n, 8)."""
    p = FQ

    def jdbl(P):
        X, Y, Z = P
        A = X * X % p; B = Y * Y % p; Cc = B * B % p
        D = 2 * ((X + B) * (X + B) - A - Cc) % p
        E = 3 * A % p
        X3 = (E * E - 2 * D) % p
        return (X3, (E * (D - X3) - 8 * Cc) % p, 2 * Y * Z % p)

    def jadd_affine(P, q):
        X1, Y1, Z1 = P
        x2, y2 = q
        if Z1 == 0:
            return (x2, y2, 1)
        Z1Z1 = Z1 * Z1 % p
        U2 = x2 * Z1Z1 % p
        S2 = y2 * Z1 * Z1Z1 % p
        if U2 == X1:
            return jdbl(P) if S2 == Y1 else (1, 1, 0)
        H = (U2 - X1) % p
        Rr = (S2 - Y1) % p
        HH = H * H % p
        HHH = H * HH % p
        V = X1 * HH % p
        X3 = (Rr * Rr - HHH - 2 * V) % p
        return (X3, (Rr * (V - X3) - Y1 * HHH) % p, Z1 * H % p)

    G = (1, 2)
    # (start + 1) * G by double-and-add
    acc = (1, 1, 0)
    for bit in bin(start + 1)[2:]:
        acc = jdbl(acc) if acc[2] else acc
        if bit == "1":
            acc = jadd_affine(acc, G)
    pts = []
    for _ in range(n):
        pts.append(acc)
        acc = jadd_affine(acc, G)
    # batched inversion of the Z coordinates
    pre, run = [], 1
    for (_, _, z) in pts:
        pre.append(run)
        run = run * z % p
    inv = pow(run, -1, p)
    out = np.empty((n, 8), dtype=np.uint64)
    for i in range(n - 1, -1, -1):
        X, Y, Z = pts[i]
        zi = inv * pre[i] % p
        inv = inv * Z % p
        zi2 = zi * zi % p
        out[i, :4] = _mont_limbs(X * zi2 % p, p)
        out[i, 4:] = _mont_limbs(Y * zi2 * zi % p, p)
    return out
