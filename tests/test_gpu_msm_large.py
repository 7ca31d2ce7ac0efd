"""BASELINE config 4 under oracle parity (GPU): best_multiexp / ParamsKZG::commit_lagrange at the sweep sizes.

  2^18, 2^20   full compare with the oracle's best_multiexp run here (uniform and witness-like scalars)
  2^22, 2^24   compare with the committed oracle results of tests/golden/msm_large.json (generator:
               tests/golden/make_golden_msm_large.py); the inputs are regenerated from their seeds and their SHA-256 is
               checked against the fixture first, so a mismatch of the point cannot be an input mismatch
  sharded      the base-range split of SURVEY.md 8e on the product's exchange path: NCCL all_gather of the 96-byte partials
               (one process per visible GPU) and the in-process de_commit_sharded, both against the oracle at 2^18
"""
import hashlib
import json
import os
import socket
import sys

import numpy as np
import pytest

import de_b200
import orc

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def inputs(log_n, kind):
    n = 1 << log_n
    if kind == "U":
        return orc.uniform_fr(0xDE04 + log_n, n)
    return orc.witness_fr(0xDE05 + log_n, n, int(n * 0.77))


def affine_xy(ctx, jac):
    a = ctx.batch_normalize(np.ascontiguousarray(jac).reshape(1, 12))[0]
    v = orc.fq_ints_from_mont(a.reshape(2, 4))
    return hex(v[0]), hex(v[1])


@pytest.mark.parametrize("log_n", [18, 20])
def test_msm_matches_oracle_at_sweep_sizes(ctx, log_n):
    n = 1 << log_n
    bases = orc.gen_bases(n)
    params = de_b200.ParamsKZG(log_n, None, bases, ctx)
    try:
        for kind in ("U", "W"):
            s = inputs(log_n, kind)
            want = orc.g1_to_affine(orc.best_multiexp(s, bases).reshape(1, 12))
            got_raw = orc.g1_to_affine(ctx.best_multiexp(s, bases).reshape(1, 12))       # best_multiexp(coeffs, bases)
            got_tab = orc.g1_to_affine(params.commit_lagrange(s).reshape(1, 12))         # ParamsKZG::commit_lagrange
            assert (got_raw == want).all(), (log_n, kind, "best_multiexp")
            assert (got_tab == want).all(), (log_n, kind, "commit_lagrange")
    finally:
        params.close()


def _golden_cases():
    with open(os.path.join(HERE, "golden", "msm_large.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("log_n", [22, 24])
def test_msm_matches_committed_oracle_result(ctx, log_n):
    cases = [c for c in _golden_cases() if c["log_n"] == log_n]
    assert {c["scalars"] for c in cases} == {"U", "W"}
    n = 1 << log_n
    bases = orc.gen_bases(n)
    bases_sha = hashlib.sha256(bases.tobytes()).hexdigest()
    params = de_b200.ParamsKZG(log_n, None, bases, ctx)
    try:
        for c in cases:
            s = inputs(log_n, c["scalars"])
            assert hashlib.sha256(s.tobytes()).hexdigest() == c["scalars_sha256"] and bases_sha == c["bases_sha256"], "inputs differ from the fixture's"
            assert affine_xy(ctx, params.commit_lagrange(s)) == (c["x"], c["y"]), (log_n, c["scalars"], "commit_lagrange")
            if c["scalars"] == "U":
                assert affine_xy(ctx, ctx.best_multiexp(s, bases)) == (c["x"], c["y"]), (log_n, "best_multiexp")
    finally:
        params.close()


# ---- base-range sharding on the product's multi-process path ------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, log_n, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
    import torch
    import torch.distributed as dist
    import de_b200 as d
    import orc as o
    from de_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n = 1 << log_n
    bases = o.gen_bases(n)
    lo, hi = sharding.base_range(n, rank, world)
    m = hi - lo
    ks = max(1, (m - 1).bit_length())
    pad = np.zeros((1 << ks, 8), dtype=np.uint64)
    pad[:m] = bases[lo:hi]
    ctx = d.Context(rank)
    params = d.ParamsKZG(ks, None, pad, ctx)
    res = {}
    for kind in ("U", "W"):
        s = o.uniform_fr(0xDE04 + log_n, n) if kind == "U" else o.witness_fr(0xDE05 + log_n, n, int(n * 0.77))
        d_s = torch.from_numpy(np.ascontiguousarray(s[lo:hi]).view(np.int64)).cuda()
        got = sharding.sharded_commit(params, 1, d_s, m)          # local MSM -> NCCL all_gather -> de_g1_sum, on every rank
        want = o.best_multiexp(s, bases)
        res[kind] = bool((o.g1_to_affine(got.reshape(1, 12)) == o.g1_to_affine(want.reshape(1, 12))).all())
    params.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump(res, f)


def test_sharded_commit_nccl_matches_oracle(tmp_path):
    """sharding.sharded_commit with one process per visible GPU (NCCL; a single rank when one GPU is visible - NCCL refuses two
    ranks on one device)"""
    import torch
    import torch.multiprocessing as mp
    world = max(1, min(torch.cuda.device_count(), 8))
    while world & (world - 1):
        world -= 1
    mp.spawn(_nccl_worker, args=(world, _free_port(), 18, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        with open(tmp_path / f"rank{r}.json") as f:
            assert json.load(f) == {"U": True, "W": True}, f"rank {r} of {world}"


def test_commit_sharded_in_process_at_sweep_size():
    """de_commit_sharded (one process, one host thread per shard) at 2^18 over every visible GPU - or 2 and 4 contexts on
    one GPU - against the oracle"""
    import torch
    from de_b200 import sharding
    log_n = 18
    n = 1 << log_n
    bases = orc.gen_bases(n)
    ndev = torch.cuda.device_count()
    layouts = [[0, 0], [0, 0, 0, 0]] + ([list(range(ndev))] if ndev > 1 else [])
    s = {k: inputs(log_n, k) for k in ("U", "W")}
    want = {k: orc.g1_to_affine(orc.best_multiexp(s[k], bases).reshape(1, 12)) for k in s}
    for devices in layouts:
        sp = sharding.ShardedParams(log_n, bases, devices, basis=1)
        try:
            for k in s:
                assert (orc.g1_to_affine(sp.commit(s[k]).reshape(1, 12)) == want[k]).all(), (devices, k)
        finally:
            sp.close()


def test_int_peak_is_plausible(ctx):
    """de_int_peak (the roofline denominator bench.py measures in the run): B200 sustains ~66 G Fr multiplications / s"""
    v = ctx.int_peak()
    assert 20.0 < v < 200.0, v
    s = ctx.int_peak_sqr()  # the dedicated squaring: fewer wide products, never slower than a product
    assert v * 0.98 < s < 2.0 * v, (v, s)
