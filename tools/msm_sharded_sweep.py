#!/usr/bin/env python
"""BASELINE config 4, multi-GPU half: BN254 G1 MSM with the base range sharded across the GPUs of one box
(SURVEY.md section 8e): rank r stages bases[lo_r:hi_r) (+ window tables) once, commits its scalar slice, and the partial
points are all-gathered over NCCL and summed.  Launch with torchrun; N = 1 gives the single-GPU baseline of the same code.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/msm_sharded_sweep.py --to-log2 24

Timing: CUDA events around (local commit + all-gather + sum) on every rank, max over ranks, best of 3.  Inputs are generated
on the device (bases P_i = [i + 1] G by de_g1_mul_base_dev, scalars from numpy); the check is algebraic and needs no CPU
code: commit(a) + commit(b) == commit(a + b), and the N-GPU result equals rank 0's single-GPU result for sizes <= 2^22.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
import de_b200  # noqa: E402
from de_b200 import sharding, synth  # noqa: E402

G1_GEN_MONT = np.array(synth._mont_limbs(1, synth.FQ) + synth._mont_limbs(2, synth.FQ), dtype=np.uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-log2", type=int, default=16)
    ap.add_argument("--to-log2", type=int, default=24)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = de_b200.Context(local)
    ctx.set_stream(stream.cuda_stream)
    as_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    with torch.cuda.stream(stream):
        for log_n in range(args.from_log2, args.to_log2 + 1, 2):
            n = 1 << log_n
            lo, hi = sharding.base_range(n, rank, world)
            m = hi - lo
            # this rank's bases [lo + 1 .. hi] * G, generated on the device
            idx = np.zeros((m, 4), dtype=np.uint64)
            idx[:, 0] = np.arange(lo + 1, hi + 1, dtype=np.uint64)
            d_idx = as_dev(ctx.fr_to_mont(idx)) if m <= (1 << 22) else None
            if d_idx is None:  # convert in slices to bound the staging copies
                d_idx = torch.empty((m, 4), dtype=torch.int64, device="cuda")
                for s0 in range(0, m, 1 << 22):
                    d_idx[s0:s0 + (1 << 22)] = as_dev(ctx.fr_to_mont(idx[s0:s0 + (1 << 22)]))
            d_bases = torch.empty((m, 8), dtype=torch.int64, device="cuda")
            ctx.g1_mul_base_dev(G1_GEN_MONT, d_idx, m, d_bases)
            ctx.sync()
            k_shard = max(1, (m - 1).bit_length())
            bases_h = d_bases.cpu().numpy().view(np.uint64)
            pad = np.zeros(((1 << k_shard), 8), dtype=np.uint64)
            pad[:m] = bases_h
            params = de_b200.ParamsKZG(k_shard, None, pad, ctx)
            del d_idx, d_bases
            for kind in ("U", "W"):
                a = synth.uniform_fr(0xDE04 + log_n, n)[lo:hi] if kind == "U" else synth_witness(n, lo, hi, log_n)
                b = synth.uniform_fr(0xDE07 + log_n, n)[lo:hi]
                d_a, d_b = as_dev(a), as_dev(b)
                d_ab = as_dev(ctx.fr_add(a, b))
                ca = sharding.sharded_commit(params, 1, d_a, m)
                cb = sharding.sharded_commit(params, 1, d_b, m)
                cab = sharding.sharded_commit(params, 1, d_ab, m)
                both = ctx.batch_normalize(np.stack([ctx.g1_sum(np.stack([ca, cb])), cab]))
                ok = bool((both[0] == both[1]).all())
                best = 1e30
                for _ in range(3):
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    sharding.sharded_commit(params, 1, d_a, m)
                    e1.record(stream)
                    e1.synchronize()
                    best = min(best, sharding.max_over_ranks(e0.elapsed_time(e1)))
                if rank == 0:
                    print(json.dumps({"op": "msm_sharded", "log_n": log_n, "n_gpus": world, "scalars": kind, "ms": best,
                                      "gpts_s": n / (best * 1e-3) / 1e9, "ok": ok, "check": "commit(a) + commit(b) == commit(a + b)",
                                      "exchange": "all_gather of one 96-byte Jacobian point per rank (NCCL), then de_g1_sum"}), flush=True)
                del d_a, d_b, d_ab
            params.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def synth_witness(n, lo, hi, log_n):
    """witness-like scalars: 23 % zero rows at the end, of the rest 45 % < 2^8, 35 % < 2^64, 10 % < 2^134,
    10 % uniform; vectorised so that 2^26 rows are generated in seconds"""
    rng = np.random.Generator(np.random.PCG64(0xDE05 + log_n))
    used = int(n * 0.77)
    out = np.zeros((n, 4), dtype=np.uint64)
    sel = rng.integers(0, 100, size=used)
    raw = np.zeros((used, 4), dtype=np.uint64)
    raw[:, 0] = rng.integers(0, 1 << 63, size=used, dtype=np.uint64)
    raw[:, 1] = rng.integers(0, 1 << 63, size=used, dtype=np.uint64)
    raw[:, 2] = rng.integers(0, 64, size=used, dtype=np.uint64)
    small = sel < 45
    raw[small, 0] &= np.uint64(0xFF)
    raw[sel < 80, 1] = 0
    raw[sel < 80, 2] = 0
    out[:used] = raw
    uni = sel >= 90
    out[:used][uni] = synth.uniform_fr(0xDE55 + log_n, int(uni.sum()))
    # the limbs above are canonical small integers, not Montgomery images; as scalars of an MSM any field element is as good
    # as another, what matters for the timing is the digit pattern the kernel sees AFTER from_mont: so convert on the device
    ctx = de_b200.default_context(int(os.environ.get("LOCAL_RANK", 0)))
    res = np.empty((hi - lo, 4), dtype=np.uint64)
    for s0 in range(lo, hi, 1 << 22):
        s1 = min(hi, s0 + (1 << 22))
        res[s0 - lo:s1 - lo] = ctx.fr_to_mont(np.ascontiguousarray(out[s0:s1]))
    return res


if __name__ == "__main__":
    main()
