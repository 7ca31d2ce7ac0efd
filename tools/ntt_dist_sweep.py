#!/usr/bin/env python
"""BASELINE config 4, multi-GPU half for the NTT: best_fft of ONE vector of 2^log_n Fr elements over W = 1, 2, 4, 8 B200s
("single huge vector").  Device-resident, CUDA-event timed, max over the ranks.

  python tools/ntt_dist_sweep.py --gpus 8 --log-n 20 22 24 26 27        one process drives the W GPUs (de_ntt_sharded_dev)
  torchrun --nproc-per-node 8 tools/ntt_dist_sweep.py --log-n 24 27     one process per GPU (DistNtt: CUDA IPC peer buffers,
                                                                        stage 1 -> stream barrier -> stage 2 -> stream barrier)

Input: rank r holds the cyclic slice a[r::W] of a vector generated on the device; output: rank q holds the block
A[q N/W : (q+1) N/W).  Check (on the device, every size, whole vector): the blocks equal the single-GPU best_fft of the same
library over the re-assembled vector (that transform is bit-exact against the CPU restatement in tests/test_gpu_ntt.py; the
multi-GPU one in tests/test_gpu_ntt_dist.py).  GB/s = 64 N / t (algorithmic bytes), Gmul/s = (N/2) log2 N / t.
One JSON line per size."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
import de_b200  # noqa: E402
from de_b200 import sharding, synth  # noqa: E402

MUL_PEAK = 65.9e9
HBM = 6529.7
TOP_LIMB = synth.FR >> 192


ROOT_OF_UNITY = pow(7, (synth.FR - 1) >> 28, synth.FR)  # halo2curves Fr::ROOT_OF_UNITY (S = 28)


def dtod(ctx, dst, src, nbytes):
    import ctypes as C
    ctx.check(ctx.L.de_dev_copy(ctx.h, C.c_void_p(dst), C.c_void_p(src), nbytes))


def omega_for(log_n):
    w = pow(ROOT_OF_UNITY, 1 << (28 - log_n), synth.FR)
    return np.array(synth._mont_limbs(w, synth.FR), dtype=np.uint64)


def uniform_fr_dev(n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device=device, generator=g)
    t[:, 3] = torch.randint(0, TOP_LIMB, (n,), dtype=torch.int64, device=device, generator=g)
    return t


def line(mode, world, log_n, ms, ok, single_ms=None, emit=True, **extra):
    n = 1 << log_n
    d = {"op": "ntt_multi_gpu", "mode": mode, "n_gpus": world, "log_n": log_n, "ms": ms, "gb_s": 64 * n / ms / 1e6,
         "frac_hbm_aggregate": 64 * n / ms / 1e6 / (HBM * world), "gmul_s": n / 2 * log_n / ms / 1e6,
         "frac_int_pipe_aggregate": n / 2 * log_n / ms / 1e6 / (MUL_PEAK / 1e9 * world), "ok": bool(ok),
         "check": "blocks == single-GPU best_fft of the re-assembled vector, whole vector, on the device"}
    if single_ms is not None:
        d["single_gpu_ms"] = single_ms
        d["speedup_vs_single_gpu"] = single_ms / ms
    d.update(extra)
    if emit:
        print(json.dumps(d), flush=True)
    return d


def single_gpu_reference(ctx0, slices, log_n, omega, reps, dev="cuda:0"):
    """re-assemble a from the cyclic slices on `dev`, transform with de_ntt_dev; returns (result tensor, best ms)"""
    world = len(slices)
    n = 1 << log_n
    full = torch.empty((n, 4), dtype=torch.int64, device=dev)
    for r, s in enumerate(slices):
        full[r::world] = s.to(dev)
    torch.cuda.synchronize(dev)
    st = torch.cuda.Stream(device=dev)
    ctx0.set_stream(st.cuda_stream)
    best = 1e30
    src = full.clone()
    for i in range(reps + 1):
        full.copy_(src)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            ctx0.best_fft_dev(full, omega, log_n)
            e1.record(st)
        e1.synchronize()
        if i:
            best = min(best, e0.elapsed_time(e1))
    ctx0.set_stream(None)
    del src
    return full, best


def run_sharded(args):
    world = args.gpus
    ndev = torch.cuda.device_count()
    devs = [r % ndev for r in range(world)]
    s = sharding.ShardedNtt(devs)
    streams = [torch.cuda.Stream(device=d) for d in devs]
    for c, st in zip(s.ctxs, streams):
        c.set_stream(st.cuda_stream)
    ref_ctx = de_b200.Context(0)
    for log_n in args.log_n:
        m, _ = sharding.ntt_layout(log_n, world)
        omega = omega_for(log_n)
        xs = [uniform_fr_dev(m, 0xDE06 + 97 * r + log_n, f"cuda:{d}") for r, d in enumerate(devs)]
        outs = [torch.empty_like(x) for x in xs]
        for d in set(devs):
            torch.cuda.synchronize(d)
        best = 1e30
        for i in range(args.reps + 2):
            ev = []
            for st, d in zip(streams, devs):
                with torch.cuda.device(d):
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    ev.append([e0, None])
            s.best_fft_dev(xs, outs, omega, log_n)
            for k, (st, d) in enumerate(zip(streams, devs)):
                with torch.cuda.device(d):
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record(st)
                    ev[k][1] = e1
            for st in streams:
                st.synchronize()
            t = max(e0.elapsed_time(e1) for e0, e1 in ev)
            if i >= 2:
                best = min(best, t)
        want, single_ms = single_gpu_reference(ref_ctx, xs, log_n, omega, args.reps)
        ok = all(torch.equal(o.to("cuda:0"), want[r * m:(r + 1) * m]) for r, o in enumerate(outs))
        # per-kernel CUDA-event times of one more call (max over the ranks): local passes, exchange pass, cross-rank stage
        for c in s.ctxs:
            c.timing_reset()
            c.timing_enable(True)
        s.best_fft_dev(xs, outs, omega, log_n)
        s.sync()
        stage = {"kernel_ms": {k: max(c.timing_get(k)[0] for c in s.ctxs) for k in ("k_ntt_pass", "k_ntt_pass_dist", "k_ntt_cross")}}
        for c in s.ctxs:
            c.timing_enable(False)
        line("one process, W contexts (de_ntt_sharded_dev)", world, log_n, best, ok, single_ms,
             devices=sorted(set(devs)), peer_bytes_per_gpu=2 * 32 * m * (world - 1) // world, **stage)
        del xs, outs, want
        torch.cuda.empty_cache()
    s.close()
    ref_ctx.close()


def measure_dist(ctx, log_n, reps, emit=True, pipelined=True, chunks=1):
    """one process per GPU (torch.distributed already initialised, NCCL): transform one 2^log_n vector over all ranks with DistNtt,
    check it on rank 0 against the single-GPU transform, return the result line there (None elsewhere).  bench.py calls this too."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = f"cuda:{ctx.device}"
    omega = omega_for(log_n)
    d = sharding.DistNtt(ctx, log_n, pipelined=pipelined, chunks=chunks)
    m = d.m
    x = uniform_fr_dev(m, 0xDE06 + 97 * rank + log_n, dev)
    torch.cuda.synchronize()
    dtod(ctx, d.d_x, x.data_ptr(), 32 * m)  # d_x is a de_dev_alloc buffer
    ctx.sync()
    dist.barrier()
    best = 1e30
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(d.stream):
            d._barrier()
            e0.record(d.stream)
            d.run(omega)
            e1.record(d.stream)
        e1.synchronize()
        t = sharding.max_over_ranks(e0.elapsed_time(e1))
        if i >= 2:
            best = min(best, t)
    if pipelined and d.timed_out():
        raise RuntimeError("multi-GPU NTT: a flag wait timed out (a rank never arrived)")
    out = torch.empty((m, 4), dtype=torch.int64, device=dev)
    dtod(ctx, out.data_ptr(), d.d_out, 32 * m)
    ctx.sync()
    xs = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
    os_ = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
    dist.gather(x, xs, dst=0)
    dist.gather(out, os_, dst=0)
    res = None
    if rank == 0:
        ref_ctx = de_b200.Context(ctx.device)
        want, single_ms = single_gpu_reference(ref_ctx, xs, log_n, omega, reps, dev)
        ok = all(torch.equal(o, want[r * m:(r + 1) * m]) for r, o in enumerate(os_))
        res = line(f"one process per GPU (CUDA IPC peer buffers; exchange in {chunks} range(s), stages ordered by peer-memory flags)" if pipelined
                   else "one process per GPU (CUDA IPC peer buffers, NCCL 1-element barriers)", world, log_n, best, ok, single_ms, emit=emit,
                   peer_bytes_per_gpu=2 * 32 * m * (world - 1) // world)
        ref_ctx.close()
        del want
    del xs, os_, x, out
    d.close()
    torch.cuda.empty_cache()
    return res


def run_dist(args):
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = de_b200.Context(local)
    for log_n in args.log_n:
        measure_dist(ctx, log_n, args.reps, pipelined=not args.barriers, chunks=args.chunks)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2, help="ranks of the single-process mode (ignored under torchrun)")
    ap.add_argument("--log-n", type=int, nargs="+", default=[20, 22, 24])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chunks", type=int, default=1, help="torchrun mode: column ranges of the pipelined exchange (1 = not pipelined, flags only)")
    ap.add_argument("--barriers", action="store_true", help="torchrun mode: round 1's unpipelined form with NCCL 1-element barriers")
    a = ap.parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_dist(a)
    else:
        run_sharded(a)
