"""a3 / a8 parity (GPU): best_multiexp and ParamsKZG::commit / commit_lagrange through the C ABI == oracle (compared as
affine points, bit-exact), including the exceptional cases of the group law and skewed (witness-like) scalars."""
import numpy as np
import pytest

import de_b200
import orc
import pyoracle as po
from util import affine_ints, fr, golden, points

pytestmark = pytest.mark.gpu


def same_point(a, b):
    return (orc.g1_to_affine(a.reshape(1, 12)) == orc.g1_to_affine(b.reshape(1, 12))).all()


def test_golden_msm(ctx):
    for case in golden()["msm"]:
        got = ctx.best_multiexp(fr(case["scalars"]), points(case["bases"]))
        want = None if case["result"] is None else (int(case["result"][0], 16), int(case["result"][1], 16))
        assert affine_ints(got) == want


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 1000, 4096, 1 << 14, (1 << 16) + 17])
def test_best_multiexp_matches_oracle(ctx, n):
    s = orc.uniform_fr(0xDE04 + n, n)
    b = orc.gen_bases(n)
    assert same_point(ctx.best_multiexp(s, b), orc.best_multiexp(s, b))


def test_best_multiexp_length_mismatch(ctx):
    with pytest.raises(ValueError):
        ctx.best_multiexp(orc.uniform_fr(1, 4), orc.gen_bases(5))


def test_exceptional_cases(ctx):
    n = 256
    bases = orc.gen_bases(n)
    ones = orc.fr_mont_from_ints([1] * n)
    # sum of (i+1)G hits acc == next base (doubling) inside a bucket
    assert affine_ints(ctx.best_multiexp(ones, bases)) == po.g1_mul(po.G1_GEN, n * (n + 1) // 2)
    # all zero scalars -> identity
    assert affine_ints(ctx.best_multiexp(np.zeros((n, 4), dtype=np.uint64), bases)) is None
    # the same base n times -> repeated doubling
    same = np.repeat(bases[:1], n, axis=0)
    assert affine_ints(ctx.best_multiexp(ones, same)) == po.g1_mul(po.G1_GEN, n)
    # P + (-P) -> identity; identity bases are skipped
    pair = bases[:4].copy()
    pair[1, :4] = pair[0, :4]
    y = orc.fq_ints_from_mont(pair[0, 4:].reshape(1, 4))[0]
    pair[1, 4:] = orc.fq_mont_from_ints([po.FQ - y])[0]
    pair[2] = 0
    s = orc.fr_mont_from_ints([5, 5, 9, 0])
    assert affine_ints(ctx.best_multiexp(s, pair)) is None
    # scalars r-1, r-2 (maximal signed-digit carries)
    big = orc.fr_mont_from_ints([po.FR - 1, po.FR - 2, 1 << 253, (1 << 128) - 1])
    assert same_point(ctx.best_multiexp(big, bases[:4]), orc.best_multiexp(big, bases[:4]))


def test_witness_like_scalars(ctx):
    # SURVEY.md 8d distribution "W": many zeros and small values -> heavy buckets, split tasks and the warp merge
    n = 1 << 15
    s = orc.witness_fr(0xDE05, n, used=int(n * 0.77))
    b = orc.gen_bases(n)
    assert same_point(ctx.best_multiexp(s, b), orc.best_multiexp(s, b))
    # one value repeated everywhere: a single bucket per window holds every point
    s2 = np.repeat(orc.fr_mont_from_ints([0x1234567]), n, axis=0)
    assert same_point(ctx.best_multiexp(s2, b), orc.best_multiexp(s2, b))


@pytest.mark.parametrize("k", [4, 10, 13])
def test_params_commit_matches_oracle(ctx, k):
    n = 1 << k
    g = orc.gen_bases(n)
    gl = orc.gen_bases(n, start=n)
    params = de_b200.ParamsKZG(k, g, gl, ctx)
    s = orc.uniform_fr(k, n)
    assert same_point(params.commit(s), orc.best_multiexp(s, g))
    assert same_point(params.commit_lagrange(s), orc.best_multiexp(s, gl))
    # shorter polynomial: uses g[..len]
    assert same_point(params.commit(s[: n // 2 + 1]), orc.best_multiexp(s[: n // 2 + 1], g[: n // 2 + 1]))
    w = orc.witness_fr(k + 1, n, used=n // 2)
    assert same_point(params.commit_lagrange(w), orc.best_multiexp(w, gl))


def test_commit_batch_and_ranges(ctx):
    k = 12
    n = 1 << k
    gl = orc.gen_bases(n)
    params = de_b200.ParamsKZG(k, None, gl, ctx)
    polys = [orc.uniform_fr(50 + i, n) for i in range(4)] + [orc.witness_fr(60, n, n // 3)]
    got = params.commit_batch(1, polys)
    for i, p in enumerate(polys):
        assert same_point(got[i], orc.best_multiexp(p, gl))
    # base-range sharding: partial commitments over disjoint ranges sum to the commitment (multi-GPU row 8e)
    parts = np.stack([params.commit_range(1, polys[0], lo, hi) for lo, hi in ((0, 1000), (1000, 1000), (1000, 3000), (3000, n))])
    assert same_point(ctx.g1_sum(parts), orc.best_multiexp(polys[0], gl))


def test_commit_graph_replay(ctx):
    """the commitment's launch sequence is captured as a CUDA graph on the second call with the same buffers (one proof in
    flight) and replayed afterwards: new scalars written in place, a witness-like column (another task length chosen on the
    device, heavy buckets) and a return to uniform ones must all give the oracle's points"""
    import torch
    k, count = 11, 3
    n = 1 << k
    gl = orc.gen_bases(n)
    params = de_b200.ParamsKZG(k, None, gl, ctx)
    d = torch.empty((count, n, 4), dtype=torch.int64, device="cuda")
    launches = []
    for it in range(6):
        polys = [orc.witness_fr(300 + it, n, n // 2) if it == 3 else orc.uniform_fr(200 + 7 * it + j, n) for j in range(count)]
        d.copy_(torch.from_numpy(np.stack(polys).view(np.int64)))
        torch.cuda.synchronize()
        before = ctx.launches
        got = params.commit_batch_dev(1, d, n, count)
        launches.append(ctx.launches - before)
        for j in range(count):
            assert same_point(got[j], orc.best_multiexp(polys[j], gl)), (it, j)
    assert len(set(launches)) == 1, launches  # replays keep counting the kernels inside the graph


def test_commit_delay_enc_size(ctx):
    # the bench configuration: k = 16 commit_lagrange, uniform and witness-like columns
    k = 16
    n = 1 << k
    gl = orc.gen_bases(n)
    params = de_b200.ParamsKZG(k, None, gl, ctx)
    for s in (orc.uniform_fr(0xDE03, n), orc.witness_fr(0xDE03, n, 50400)):
        assert same_point(params.commit_lagrange(s), orc.best_multiexp(s, gl))


def test_batch_normalize(ctx):
    n = 300
    s = orc.uniform_fr(5, n)
    b = orc.gen_bases(n)
    pts = np.stack([ctx.best_multiexp(s[: i + 1], b[: i + 1]) for i in (0, 7, 299)] + [np.zeros(12, dtype=np.uint64)])
    assert (ctx.batch_normalize(pts) == orc.g1_to_affine(pts)).all()


def test_g1_mul_base_matches_oracle(ctx):
    """de_g1_mul_base_dev (fixed-base multiplications of ParamsKZG::setup / synthetic SRS generation) vs the oracle"""
    import torch
    n = 300
    s = orc.uniform_fr(0xBA5E, n)
    s[0] = 0                                   # -> identity
    s[1] = orc.fr_mont_from_ints([1])[0]       # -> the base itself
    s[2] = orc.fr_mont_from_ints([po.FR - 1])[0]  # -> minus the base
    s[3:40] = orc.fr_mont_from_ints(list(range(2, 39)))
    base = orc.gen_bases(1)[0]
    d_s = torch.from_numpy(s.view(np.int64)).cuda()
    d_out = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    ctx.g1_mul_base_dev(base, d_s, n, d_out)
    ctx.sync()
    got = d_out.cpu().numpy().view(np.uint64)
    assert (got == orc.g1_mul_many(base, s)).all()
    assert not got[0].any() and (got[1] == base).all()


def test_commit_sharded_single_process():
    """de_commit_sharded: base ranges on several contexts (all the visible GPUs; two contexts on one GPU when only one is
    visible), one host thread each, partial points summed: equals the oracle's best_multiexp over the whole vector"""
    import torch
    from de_b200 import sharding
    k = 12
    n = 1 << k
    bases, s = orc.gen_bases(n), orc.uniform_fr(0x5A4D, n)
    ndev = torch.cuda.device_count()
    for devices in ([0, 0], [0, 0, 0], list(range(ndev)) if ndev > 1 else [0]):
        sp = sharding.ShardedParams(k, bases, devices, basis=1)
        got = sp.commit(s)
        want = orc.best_multiexp(s, bases)
        assert (orc.g1_to_affine(got) == orc.g1_to_affine(want)).all(), devices
        sp.close()


def test_params_setup_matches_restated_setup(ctx, tmp_path):
    """ParamsKZG::setup on the device (fixed-base multiplications) equals the restated setup: g, g_lagrange, s_g2; and the
    RawBytes file written from it reads back to the same commitments"""
    import pairing
    import pyprover as pp
    k, s = 7, 0x5EC2E7123456789
    want = pp.setup(k, s)
    params = de_b200.ParamsKZG.setup(k, s, ctx)
    assert (params.g_host == want.g_mont).all() and (params.g_lagrange_host == want.g_lagrange_mont).all()
    raw = params.s_g2
    coords = [int.from_bytes(raw[32 * i:32 * i + 32], "little") * pow(1 << 256, -1, po.FQ) % po.FQ for i in range(4)]
    assert ((coords[0], coords[1]), (coords[2], coords[3])) == pairing.g2_mul(pairing.G2_GEN, s)
    path = tmp_path / "params_7"
    params.write(path)
    again = de_b200.ParamsKZG.read(path, ctx)
    poly = orc.uniform_fr(77, 1 << k)
    assert (orc.g1_to_affine(again.commit_lagrange(poly)) == orc.g1_to_affine(orc.best_multiexp(poly, want.g_lagrange_mont))).all()
    assert (orc.g1_to_affine(params.commit(poly)) == orc.g1_to_affine(orc.best_multiexp(poly, want.g_mont))).all()
    params.close(); again.close()
