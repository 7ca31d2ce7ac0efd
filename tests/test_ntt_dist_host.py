"""Multi-GPU NTT ("single huge vector" row) on the CPU: the four-step schedule and the layouts of
de_b200/sharding.py (cyclic input slices, exchange slots, contiguous output blocks) are run with the oracle's field arithmetic
- all ranks simulated in one process for W = 1, 2, 4, 8, and as two real gloo processes that exchange the columns - and must
reproduce best_fft.  The CUDA kernels (de_ntt_dist_stage1 / 2) implement exactly these two stages; tests/test_gpu_ntt_dist.py
checks them on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _omega(log_n):
    import orc
    import pyoracle as po
    return orc.fr_mont_from_ints([pow(po.FR_ROOT_OF_UNITY, 1 << (po.FR_S - log_n), po.FR)])[0]


def _pow(w, e):
    import orc
    import pyoracle as po
    return orc.fr_mont_from_ints([pow(orc.fr_ints_from_mont(w.reshape(1, 4))[0], e, po.FR)])[0]


def stage1(x, omega, log_n, world, rank):
    """local M-point transform with root omega^W, then column j2 times omega^(rank * j2)"""
    import orc
    import pyoracle as po
    m = log_n - (world.bit_length() - 1)
    b = orc.best_fft(np.ascontiguousarray(x), _pow(omega, world), m)
    w = orc.fr_ints_from_mont(omega.reshape(1, 4))[0]
    tw = orc.fr_mont_from_ints([pow(w, rank * j, po.FR) for j in range(1 << m)])
    return orc.fr_mul(b, tw)


def stage2(z, omega, log_n, world):
    """W-point transform down the columns of z[W][C]; returns out[j1][c]"""
    import orc
    c = z.shape[0] // world
    z = z.reshape(world, c, 4)
    ww = _pow(omega, (1 << log_n) // world)
    out = np.zeros_like(z)
    for j1 in range(world):
        acc = np.zeros((c, 4), dtype=np.uint64)
        for i1 in range(world):
            f = np.repeat(_pow(ww, i1 * j1).reshape(1, 4), c, axis=0)
            acc = orc.fr_add(acc, orc.fr_mul(np.ascontiguousarray(z[i1]), f))
        out[j1] = acc
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_schedule_simulated_ranks(world):
    import orc
    from de_b200 import sharding
    log_n = 11 + world.bit_length() - 1
    n = 1 << log_n
    a = orc.uniform_fr(0xD157 + world, n)
    omega = _omega(log_n)
    m, c = sharding.ntt_layout(log_n, world)
    z = [np.zeros((m, 4), dtype=np.uint64) for _ in range(world)]
    for r in range(world):
        t = stage1(sharding.ntt_input_slice(a, r, world), omega, log_n, world, r)
        for j2 in range(m):
            q, pos = sharding.ntt_exchange_slot(log_n, world, r, j2)
            z[q][pos] = t[j2]
    out = [np.zeros((m, 4), dtype=np.uint64) for _ in range(world)]
    for q in range(world):
        o = stage2(z[q], omega, log_n, world)
        for j1 in range(world):
            out[j1][q * c:(q + 1) * c] = o[j1]
    want = orc.best_fft(a, omega, log_n)
    for r in range(world):
        lo, hi = sharding.ntt_output_range(log_n, r, world)
        assert (out[r] == want[lo:hi]).all()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_layouts_are_partitions(world):
    """the cyclic input slices cover every index once, the exchange slots of all (source rank, column) pairs fill every rank's buffer
    exactly once, and the output blocks tile [0, N)"""
    from de_b200 import sharding
    log_n = 11 + world.bit_length() - 1
    n = 1 << log_n
    m, c = sharding.ntt_layout(log_n, world)
    assert m * world == n and c * world == m
    idx = np.arange(n)
    seen = np.concatenate([sharding.ntt_input_slice(idx, r, world) for r in range(world)])
    assert sorted(seen.tolist()) == list(range(n))
    filled = np.zeros((world, m), dtype=np.int64)
    for src in range(world):
        for col in range(m):
            q, pos = sharding.ntt_exchange_slot(log_n, world, src, col)
            assert 0 <= q < world and pos // c == src      # row `src` of rank q's buffer
            filled[q, pos] += 1
    assert (filled == 1).all()
    blocks = [sharding.ntt_output_range(log_n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_layout_rejects_bad_arguments():
    from de_b200 import sharding
    for log_n, world in [(11, 2), (29, 2), (16, 3), (16, 16), (16, 0)]:
        with pytest.raises(ValueError):
            sharding.ntt_layout(log_n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import orc
    from de_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log_n = 12
    a = orc.uniform_fr(0xD150, 1 << log_n)  # every rank derives the same vector, keeps only its cyclic slice
    omega = _omega(log_n)
    m, c = sharding.ntt_layout(log_n, world)
    t = stage1(sharding.ntt_input_slice(a, rank, world), omega, log_n, world, rank)
    # exchange 1: column j2 goes to rank j2 / C (the peer stores of stage 1)
    cols = [None] * world
    dist.all_gather_object(cols, t)
    z = np.zeros((m, 4), dtype=np.uint64)
    for src in range(world):
        for j2 in range(rank * c, (rank + 1) * c):
            q, pos = sharding.ntt_exchange_slot(log_n, world, src, j2)
            assert q == rank
            z[pos] = cols[src][j2]
    o = stage2(z, omega, log_n, world)
    # exchange 2: output j1 of every column goes to rank j1 (the peer stores of stage 2)
    outs = [None] * world
    dist.all_gather_object(outs, o)
    mine = np.concatenate([outs[q][rank] for q in range(world)], axis=0)
    lo, hi = sharding.ntt_output_range(log_n, rank, world)
    assert (mine == orc.best_fft(a, omega, log_n)[lo:hi]).all()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_two_rank_ntt(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
