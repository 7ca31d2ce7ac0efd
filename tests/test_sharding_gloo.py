"""Multi-GPU row on the CPU: world_size-2 gloo processes exercise the host logic of proof-batch sharding
(rank -> proofs, no data-path collective, max-over-ranks timing) and of MSM base-range sharding (disjoint ranges covering
[0, n), partial points combined by addition).  The group arithmetic here is the oracle's: no GPU is involved."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


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
    # 1. batch-of-proofs sharding: 7 proofs over 2 ranks
    mine = sharding.proofs_for_rank(7, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(7))
    # 2. max-over-ranks timing
    t = sharding.max_over_ranks(10.0 + rank)
    assert t == 10.0 + world - 1
    # 3. MSM base-range sharding: each rank commits its slice, the partial points are exchanged and added
    n = 512
    scalars = orc.uniform_fr(900, n)
    bases = orc.gen_bases(n)
    lo, hi = sharding.base_range(n, rank, world)
    partial = orc.best_multiexp(np.ascontiguousarray(scalars[lo:hi]), np.ascontiguousarray(bases[lo:hi]), threads=1)
    parts = sharding.gather_partials(partial)  # the product's exchange step (all_gather of 96 bytes per rank)
    assert parts.shape == (world, 12) and (parts[rank] == partial).all()
    acc = np.zeros(12, dtype=np.uint64)
    for p in parts:
        acc = orc.g1_add(acc, np.ascontiguousarray(p))
    want = orc.best_multiexp(scalars, bases, threads=1)
    assert (orc.g1_to_affine(acc) == orc.g1_to_affine(want)).all()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_two_rank_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
