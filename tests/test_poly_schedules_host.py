"""Opening-phase schedules of csrc/poly.cu on the CPU (plain integers mod r): the segmented polynomial evaluation with its
power-tree combine (k_eval_polynomial) and the two-level carry scan of kate_division (k_kate_carries) are restated thread by
thread - same chunking, same folds, same order - and must give halo2's eval_polynomial / kate_division (oracle/pyoracle.py,
oracle/pyprover.py; reference call sites: halo2_proofs::arithmetic::{eval_polynomial, kate_division} under
/root/reference/benches/delay_enc.rs:123).  The CUDA kernels are checked on the GPU by tests/test_gpu_prover.py; this file pins
the index arithmetic (ragged last chunks, threads without work, one segment vs several) where no GPU is needed."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pyoracle as po  # noqa: E402
import pyprover as pp  # noqa: E402

R = po.FR


def eval_polynomial_schedule(poly, x, segs, threads):
    """k_eval_polynomial: grid (evaluation, segs) x `threads` threads (a power of two).  Global thread g Horner-evaluates the chunk
    [g * chunk, (g + 1) * chunk); with X = x^chunk a CTA folds slot t + d into slot t with ONE multiplication by X^d per level;
    the CTA that finishes last folds the segments by Horner in Y = X^threads."""
    n = len(poly)
    chunk = (n + segs * threads - 1) // (segs * threads)
    X = pow(x, chunk, R)
    pw = [X]
    while (1 << (len(pw) - 1)) < threads:
        pw.append(pw[-1] * pw[-1] % R)  # pw[k] = X^(2^k); the last one is Y
    partial = []
    for seg in range(segs):
        sm = []
        for tid in range(threads):
            lo = (seg * threads + tid) * chunk
            hi = min(lo + chunk, n)
            acc = 0
            for i in range(hi, lo, -1):  # empty when lo >= n: that thread contributes zero
                acc = (acc * x + poly[i - 1]) % R
            sm.append(acc)
        d, k = threads >> 1, threads.bit_length() - 2
        while d >= 1:
            for tid in range(d):
                sm[tid] = (sm[tid] + sm[tid + d] * pw[k]) % R
            d >>= 1
            k -= 1
        partial.append(sm[0])
    Y = pw[threads.bit_length() - 1]
    r = partial[-1]
    for s in range(segs - 2, -1, -1):
        r = (r * Y + partial[s]) % R
    return r


@pytest.mark.parametrize("n", [0, 1, 2, 7, 8, 9, 63, 64, 65, 255, 1000, 1025])
@pytest.mark.parametrize("segs,threads", [(1, 8), (4, 8), (4, 16), (3, 4)])
def test_segmented_evaluation_matches_eval_polynomial(n, segs, threads):
    rng = po.Xoshiro(0xE7A1 + 131 * n + segs)
    poly = [rng.uniform_fr() for _ in range(n)]
    for x in (rng.uniform_fr(), 0, 1, R - 1):
        assert eval_polynomial_schedule(poly, x, segs, threads) == po.eval_poly(poly, x)


def kate_division_schedule(a, b, chunk, scan_threads):
    """k_kate_chunk_values / k_kate_carries / k_kate_write: t_i = q[i - 1] satisfies t_i = a[i] + b * t_(i + 1), t_n = 0.
    Per chunk c = [lo, hi): V_c = sum a[i] b^(i - lo), so t_lo = V_c + B * t_hi with B = b^chunk (the last chunk may be shorter,
    and it is the first element of the descending order: its carry is zero whatever its multiplier).  carry[c] = t_(hi_c): thread t
    composes G consecutive maps y -> V + B y by Horner, one Hillis-Steele sweep multiplies by (B^G)^d at step d, the thread then
    replays its chunks from the carry that enters them."""
    n = len(a)
    nchunks = (n + chunk - 1) // chunk
    vals = []
    for c in range(nchunks):
        lo, hi = c * chunk, min((c + 1) * chunk, n)
        acc = 0
        for i in range(hi, lo, -1):
            acc = (acc * b + a[i - 1]) % R
        vals.append(acc)
    B = pow(b, chunk, R)
    G = (nchunks + scan_threads - 1) // scan_threads
    addend = lambda c: vals[c + 1] if c + 1 < nchunks else 0
    s_add = []
    for tid in range(scan_threads):
        A = 0
        for g in range(G):
            e = tid * G + g
            if e >= nchunks:
                break
            A = (addend(nchunks - 1 - e) + B * A) % R
        s_add.append(A)
    Md = pow(B, G, R)
    d = 1
    while d < scan_threads:
        prev = list(s_add)
        for tid in range(d, scan_threads):
            s_add[tid] = (prev[tid] + Md * prev[tid - d]) % R
        Md = Md * Md % R
        d <<= 1
    carries = [None] * nchunks
    for tid in range(scan_threads):
        cur = s_add[tid - 1] if tid else 0
        for g in range(G):
            e = tid * G + g
            if e >= nchunks:
                break
            c = nchunks - 1 - e
            cur = (addend(c) + B * cur) % R
            carries[c] = cur
    q = [0] * n
    for c in range(nchunks):
        lo, hi = c * chunk, min((c + 1) * chunk, n)
        cur = carries[c]
        for i in range(hi, lo, -1):
            q[i - 1] = cur
            cur = (a[i - 1] + cur * b) % R
    return q[: n - 1]


@pytest.mark.parametrize("n", [2, 3, 4, 5, 31, 32, 33, 64, 100, 257, 700])
@pytest.mark.parametrize("chunk,scan_threads", [(4, 4), (4, 8), (8, 2), (32, 512), (3, 16)])
def test_two_level_carry_scan_matches_kate_division(n, chunk, scan_threads):
    rng = po.Xoshiro(0xCA7E + 17 * n + chunk)
    a = [rng.uniform_fr() for _ in range(n)]
    for b in (rng.uniform_fr(), 0, 1, R - 1):
        assert kate_division_schedule(a, b, chunk, scan_threads) == pp.kate_division(a, b)
