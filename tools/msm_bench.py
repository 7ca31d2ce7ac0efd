#!/usr/bin/env python
"""MSM measurements for bench.py (BASELINE config 4, `metric`'s "MSM Gpts/s"): WHOLE commitments, not one kernel.

  measure_single   one GPU: ParamsKZG::commit_lagrange (bases + window tables resident) at 2^16 (alone and as the batch of 8 a
                   proof round looks like), 2^20 and 2^24, uniform (U) and witness-like (W) scalars; every size is checked on the
                   device (raw best_multiexp over the plain bases == table-based commit, two different window layouts)
  measure_sharded  N ranks (torchrun): the base range split of SURVEY.md 8e - rank r stages bases[lo_r:hi_r) once, commits its
                   scalar slice, the 96-byte partial points are all-gathered over NCCL and summed on every rank; timed as
                   local commit + exchange + sum, CUDA events, max over ranks; rank 0 also runs the same commitment on ONE GPU
                   and the two results must be the same point

Inputs are generated on the device (scalars with torch's generator, bases P_i = [i + 1] G by de_g1_mul_base_dev).  Bit-exact
parity of these operations with the CPU restatement is what tests/test_gpu_msm.py / test_gpu_msm_large.py establish (to 2^24);
this module measures.  Run alone: python tools/msm_bench.py [--gpus N under torchrun]."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import de_b200  # noqa: E402
from de_b200 import sharding, synth  # noqa: E402
from sweep import G1_GEN_MONT, ev_time, uniform_fr_dev, witness_fr_dev  # noqa: E402


def device_bases(ctx, lo: int, hi: int):
    """bases [lo + 1 .. hi] * G as an (hi - lo, 8) int64 CUDA tensor (affine, Montgomery)"""
    m = hi - lo
    idx = np.zeros((m, 4), dtype=np.uint64)
    idx[:, 0] = np.arange(lo + 1, hi + 1, dtype=np.uint64)
    d_idx = torch.empty((m, 4), dtype=torch.int64, device="cuda")
    for s0 in range(0, m, 1 << 22):
        d_idx[s0:s0 + (1 << 22)] = torch.from_numpy(ctx.fr_to_mont(idx[s0:s0 + (1 << 22)]).view(np.int64)).cuda()
    d_bases = torch.empty((m, 8), dtype=torch.int64, device="cuda")
    ctx.g1_mul_base_dev(G1_GEN_MONT, d_idx, m, d_bases)
    ctx.sync()
    return d_bases


def same_point(ctx, a, b) -> bool:
    aff = ctx.batch_normalize(np.stack([a, b]))
    return bool((aff[0] == aff[1]).all() and aff[0].any())


def measure_single(ctx, stream, log_ns=(16, 20, 24), keep_host=()):
    """-> list of dicts; for log_n in keep_host the (scalars, bases, result) host arrays are kept under "_host" so that the
    caller's CPU leg can time and compare the oracle on the same input"""
    out = []
    with torch.cuda.stream(stream):
        for log_n in log_ns:
            n = 1 << log_n
            d_bases = device_bases(ctx, 0, n)
            bases_h = d_bases.cpu().numpy().view(np.uint64)
            params = de_b200.ParamsKZG(log_n, None, bases_h, ctx)
            for kind in ("U", "W"):
                d_a = uniform_fr_dev(n, 0xDE04 + log_n) if kind == "U" else witness_fr_dev(ctx, n, 0xDE05 + log_n)
                got = params.commit_batch_dev(1, d_a, n, 1)[0]
                raw = ctx.best_multiexp_dev(d_a, d_bases, n)
                ok = same_point(ctx, got, raw)
                ms = ev_time(stream, lambda: params.commit_batch_dev(1, d_a, n, 1), 3)
                rec = {"log_n": log_n, "scalars": kind, "polys_per_call": 1, "commit_ms": ms, "gpts_s": n / (ms * 1e-3) / 1e9, "ok": ok,
                       "check": "table-based commit == raw best_multiexp over the plain bases (device)"}
                if log_n in keep_host and kind == "U":
                    rec["_host"] = (d_a.cpu().numpy().view(np.uint64), bases_h, got)
                out.append(rec)
                if log_n <= 17 and kind == "U":
                    # the shape a proof round has: several same-size polynomials in one launch sequence
                    d_b = torch.stack([uniform_fr_dev(n, 0xDE40 + i) for i in range(8)])
                    ms8 = ev_time(stream, lambda: params.commit_batch_dev(1, d_b, n, 8), 3)
                    out.append({"log_n": log_n, "scalars": kind, "polys_per_call": 8, "commit_ms": ms8, "gpts_s": 8 * n / (ms8 * 1e-3) / 1e9,
                                "ok": ok, "check": "same kernels as the single commit above"})
                    del d_b
                del d_a
            params.close()
            del d_bases, bases_h
    return out


def measure_sharded(ctx, stream, rank: int, world: int, log_n: int, reps: int = 5, single_gpu: bool = True):
    """one commitment of 2^log_n uniform scalars with the base range split over `world` ranks; returns the record on rank 0"""
    import torch.distributed as dist
    n = 1 << log_n
    lo, hi = sharding.base_range(n, rank, world)
    m = hi - lo
    with torch.cuda.stream(stream):
        d_bases = device_bases(ctx, lo, hi)
        k_shard = max(1, (m - 1).bit_length())
        pad = np.zeros(((1 << k_shard), 8), dtype=np.uint64)
        pad[:m] = d_bases.cpu().numpy().view(np.uint64)
        params = de_b200.ParamsKZG(k_shard, None, pad, ctx)
        del pad, d_bases
        # every rank draws the same full vector (same generator, same seed, same GPU model) and keeps its slice
        d_full = uniform_fr_dev(n, 0xDE04 + log_n)
        d_a = d_full[lo:hi].clone()
        if not (single_gpu and rank == 0):
            del d_full
        torch.cuda.synchronize()
        got = sharding.sharded_commit(params, 1, d_a, m)
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sharding.sharded_commit(params, 1, d_a, m)
            e1.record(stream)
            e1.synchronize()
            best = min(best, sharding.max_over_ranks(e0.elapsed_time(e1)))
        params.close()
        del d_a
        rec = {"op": "msm_multi_gpu", "n_gpus": world, "log_n": log_n, "scalars": "U", "ms": best, "gpts_s": n / (best * 1e-3) / 1e9,
               "what": "one commitment: local MSM over the rank's base range (tables resident) + NCCL all_gather of the 96-byte partial "
                       "points + de_g1_sum, max over ranks, best of %d" % reps}
        if single_gpu and rank == 0:
            # the same commitment on ONE GPU: with resident tables when they fit the per-basis budget, else raw best_multiexp
            d_all = device_bases(ctx, 0, n)
            if log_n <= 24:
                p1 = de_b200.ParamsKZG(log_n, None, d_all.cpu().numpy().view(np.uint64), ctx)
                del d_all
                one = p1.commit_batch_dev(1, d_full, n, 1)[0]
                ms1 = ev_time(stream, lambda: p1.commit_batch_dev(1, d_full, n, 1), 2)
                p1.close()
                rec["single_gpu_path"] = "ParamsKZG::commit_lagrange, tables resident"
            else:
                one = ctx.best_multiexp_dev(d_full, d_all, n)
                ms1 = ev_time(stream, lambda: ctx.best_multiexp_dev(d_full, d_all, n), 1)
                del d_all
                rec["single_gpu_path"] = "best_multiexp over the plain bases (window tables of 2^%d points exceed the per-basis budget)" % log_n
            rec.update({"single_gpu_ms": ms1, "speedup_vs_single_gpu": ms1 / best, "ok": same_point(ctx, got, one),
                        "check": "the %d-GPU commitment and the single-GPU commitment are the same point" % world})
            del d_full
        if world > 1:
            dist.barrier()
    torch.cuda.empty_cache()
    return rec if rank == 0 else None


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, nargs="*", default=[16, 20, 24])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    ctx = de_b200.Context(local)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        for log_n in args.log_n:
            rec = measure_sharded(ctx, stream, rank, world, log_n)
            if rank == 0:
                print(json.dumps(rec), flush=True)
        dist.destroy_process_group()
    else:
        for rec in measure_single(ctx, stream, args.log_n):
            print(json.dumps(rec), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
