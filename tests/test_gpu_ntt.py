"""a4-a7 parity (GPU): best_fft and the EvaluationDomain operations through the C ABI == oracle, bit-exact, plus
size-independent properties at the benchmark sizes."""
import numpy as np
import pytest

import de_b200
import orc
import pyoracle as po
from util import fr, fr_ints, golden

pytestmark = pytest.mark.gpu


def omega_for(log_n):
    return orc.fr_mont_from_ints([pow(po.FR_ROOT_OF_UNITY, 1 << (po.FR_S - log_n), po.FR)])[0]


def test_golden_ntt(ctx):
    for case in golden()["ntt"]:
        got = ctx.best_fft(fr(case["in"]), fr([case["omega"]])[0], case["log_n"])
        assert fr_ints(got) == [int(x, 16) for x in case["out"]]


@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 18, 19, 20])
def test_best_fft_matches_oracle(ctx, log_n):
    a = orc.uniform_fr(0xDE06 + log_n, 1 << log_n)
    w = omega_for(log_n)
    got = ctx.best_fft(a, w, log_n)
    want = orc.best_fft(a, w, log_n)
    assert (got == want).all()


@pytest.mark.parametrize("log_n", [21, 22, 23])
def test_best_fft_three_pass(ctx, log_n):
    a = orc.uniform_fr(0xDE06 + log_n, 1 << log_n)
    w = omega_for(log_n)
    assert (ctx.best_fft(a, w, log_n) == orc.best_fft(a, w, log_n)).all()


def test_best_fft_inverse_round_trip_large(ctx):
    # size-independent property at a sweep size: iNTT(NTT(a)) * 1/N == a  (two-level twiddle path, log_n > 22)
    log_n = 24
    a = orc.uniform_fr(77, 1 << log_n)
    w = omega_for(log_n)
    w_inv = orc.fr_inv(w.reshape(1, 4))[0]
    f = ctx.best_fft(a, w, log_n)
    b = ctx.best_fft(f, w_inv, log_n)
    n_inv = orc.fr_inv(orc.fr_mont_from_ints([1 << log_n]))
    idx = np.random.default_rng(0).integers(0, 1 << log_n, 4096)
    assert (orc.fr_mul(b[idx], np.repeat(n_inv, len(idx), axis=0)) == a[idx]).all()
    # spot check of the forward transform against the definition on a few outputs is covered by linearity:
    a2 = orc.uniform_fr(78, 1 << log_n)
    f2 = ctx.best_fft(a2, w, log_n)
    fs = ctx.best_fft(orc.fr_add(a, a2), w, log_n)
    assert (fs[idx] == orc.fr_add(f[idx], f2[idx])).all()


def test_best_fft_rejects_bad_length(ctx):
    with pytest.raises(ValueError):
        ctx.best_fft(orc.uniform_fr(1, 6), omega_for(3), 3)


def test_ntt_batched_device(ctx):
    import torch
    log_n, batch = 12, 5
    n = 1 << log_n
    a = orc.uniform_fr(5, n * batch).reshape(batch, n, 4)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ctx.best_fft_dev(d, omega_for(log_n), log_n, batch=batch, stride=n)
    ctx.sync()
    got = d.cpu().numpy().view(np.uint64)
    for b in range(batch):
        assert (got[b] == orc.best_fft(a[b], omega_for(log_n), log_n)).all()


def test_golden_domain(ctx):
    for case in golden()["domain"]:
        d = de_b200.EvaluationDomain(case["j"], case["k"], ctx)
        assert d.extended_k == case["extended_k"]
        assert fr_ints(d.omega) == [int(case["omega"], 16)]
        assert fr_ints(d.extended_omega) == [int(case["extended_omega"], 16)]
        coeff = fr(case["coeff"])
        assert fr_ints(d.coeff_to_extended(coeff)) == [int(x, 16) for x in case["extended"]]
        assert fr_ints(d.coeff_to_lagrange(coeff)) == [int(x, 16) for x in case["lagrange"]]
        assert (d.lagrange_to_coeff(fr(case["lagrange"])) == coeff).all()


# (j, k): pose_enc shape (degree 3, k=11), delay_enc (degree 5, k=16), mod_pow (degree 5, k=17), small and odd cases
@pytest.mark.parametrize("j,k", [(3, 11), (5, 16), (5, 17), (5, 4), (4, 9), (9, 6), (2, 5)])
def test_domain_ops_match_oracle(ctx, j, k):
    d = de_b200.EvaluationDomain(j, k, ctx)
    o = orc.Domain(j, k)
    assert d.extended_k == o.extended_k
    assert (d.omega == o.omega).all() and (d.omega_inv == o.omega_inv).all()
    assert (d.extended_omega == o.ext_omega).all() and (d.extended_omega_inv == o.ext_omega_inv).all()
    a = orc.uniform_fr(100 * j + k, 1 << k)
    ext = d.coeff_to_extended(a)
    assert (ext == o.coeff_to_extended(a)).all()
    assert (d.divide_by_vanishing_poly(ext) == o.divide_by_vanishing(ext)).all()
    back = d.extended_to_coeff(ext)
    assert back.shape[0] == (j - 1) << k
    assert (back == o.extended_to_coeff(ext)).all()
    assert (back[: 1 << k] == a).all() and not back[1 << k:].any()
    assert (d.lagrange_to_coeff(a) == o.lagrange_to_coeff(a)).all()
    assert (d.coeff_to_lagrange(a) == o.coeff_to_lagrange(a)).all()


def test_domain_batched_device(ctx):
    import torch
    j, k, batch = 5, 10, 7
    d = de_b200.EvaluationDomain(j, k, ctx)
    o = orc.Domain(j, k)
    a = orc.uniform_fr(9, d.n * batch).reshape(batch, d.n, 4)
    d_in = torch.from_numpy(a.view(np.int64)).cuda()
    d_out = torch.empty((batch, d.extended_n, 4), dtype=torch.int64, device="cuda")
    d.coeff_to_extended_dev(d_in, d_out, batch=batch)
    ctx.sync()
    got = d_out.cpu().numpy().view(np.uint64)
    for b in range(batch):
        assert (got[b] == o.coeff_to_extended(a[b])).all()
    d.lagrange_to_coeff_dev(d_in, batch=batch)
    ctx.sync()
    got = d_in.cpu().numpy().view(np.uint64)
    for b in range(batch):
        assert (got[b] == o.lagrange_to_coeff(a[b])).all()
