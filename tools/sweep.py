#!/usr/bin/env python
"""BASELINE config 4, single-GPU half: BN254 G1 MSM and Fr NTT size sweeps on one B200, device-resident, CUDA-event timed.
(The multi-GPU half is tools/msm_sharded_sweep.py.)

  NTT: best_fft in place (de_ntt_dev); GB/s = 64 * N / t (algorithmic bytes) and the fraction of the
       integer-pipe ceiling (N/2 * log2 N multiplies at the measured 65.9 Gmul/s).
  MSM: ParamsKZG.commit_lagrange (bases + window tables resident) and raw best_multiexp on uniform ("U") and witness-like
       ("W") scalars; Gpts/s = N / t.

Everything is generated on the device (scalars with torch's generator, bases P_i = [i + 1] G by de_g1_mul_base_dev) and the
checks are algebraic, run on the device: lagrange_to_coeff(coeff_to_lagrange(a)) == a over the whole vector, and
commit(a) + commit(b) == commit(a + b) with raw best_multiexp agreeing with the table-based commit.  Bit-exact parity with
the CPU restatement at these operations is what tests/test_gpu_ntt.py and tests/test_gpu_msm.py establish (up to 2^23 / 2^17);
this tool only measures.  One JSON line per measurement (profiles/r01_sweep*.jsonl are committed runs).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
import de_b200  # noqa: E402
from de_b200 import synth  # noqa: E402

MUL_PEAK = 65.9e9
HBM = 6529.7
TOP_LIMB = synth.FR >> 192
G1_GEN_MONT = np.array(synth._mont_limbs(1, synth.FQ) + synth._mont_limbs(2, synth.FQ), dtype=np.uint64)


def uniform_fr_dev(n, seed):
    """n Montgomery-form field elements, uniform over [0, r) up to a 2^-60 sliver: any limb pattern below r is the image of
    some field element"""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] = torch.randint(0, TOP_LIMB, (n,), dtype=torch.int64, device="cuda", generator=g)
    return t


def witness_fr_dev(ctx, n, seed):
    """witness-like scalars: the last 23 % rows zero; of the rest 45 % < 2^8, 35 % < 2^64, 10 % < 2^134,
    10 % uniform.  Built as canonical integers, converted to Montgomery form by the library in slices."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    used = int(n * 0.77)
    sel = torch.randint(0, 100, (used,), device="cuda", generator=g)
    raw = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    lo = torch.randint(0, (1 << 63) - 1, (used,), dtype=torch.int64, device="cuda", generator=g)
    raw[:used, 0] = torch.where(sel < 45, lo & 0xFF, lo)
    mid = torch.randint(0, (1 << 63) - 1, (used,), dtype=torch.int64, device="cuda", generator=g)
    raw[:used, 1] = torch.where(sel >= 80, mid, torch.zeros_like(mid))
    raw[:used, 2] = torch.where(sel >= 80, mid & 0x3F, torch.zeros_like(mid))
    host = raw.cpu().numpy().view(np.uint64)
    out = np.empty_like(host)
    for s0 in range(0, n, 1 << 22):
        out[s0:s0 + (1 << 22)] = ctx.fr_to_mont(np.ascontiguousarray(host[s0:s0 + (1 << 22)]))
    t = torch.from_numpy(out.view(np.int64)).cuda()
    uni = uniform_fr_dev(used, seed + 1)
    t[:used] = torch.where((sel >= 90).unsqueeze(1), uni, t[:used])
    return t


def ev_time(stream, fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main(cpu=None):
    """cpu: an object with best_multiexp / best_fft / g1_to_affine / ncpu (the CPU restatement of the reference's algorithms).  This
    tool never imports the checker itself: tests/sweep_vs_cpu.py passes it in for the 'vs reference' column of BASELINE config 4."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--msm-to", type=int, default=24)
    ap.add_argument("--msm-from", type=int, default=16, help="2^26 on one GPU wants DE_MSM_TABLE_MB=80000 (64 GiB of window tables)")
    ap.add_argument("--ntt-to", type=int, default=27)
    ap.add_argument("--ntt-from", type=int, default=16)
    ap.add_argument("--cpu-to", type=int, default=0,
                    help="with a CPU implementation passed to main() (tests/sweep_vs_cpu.py): time it on the same inputs up to 2^CPU_TO "
                         "and compare the results - BASELINE config 4's 'vs reference' column")
    args = ap.parse_args()
    import time
    orc = cpu if args.cpu_to else None
    if args.cpu_to and cpu is None:
        raise SystemExit("--cpu-to needs the CPU implementation: run tests/sweep_vs_cpu.py instead")
    stream = torch.cuda.Stream()
    ctx = de_b200.Context(0)
    ctx.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        # ---------------- NTT
        for log_n in range(args.ntt_from, args.ntt_to + 1):
            n = 1 << log_n
            dom = de_b200.EvaluationDomain(2, log_n, ctx)  # j = 2: extended_k = k, omega = ROOT_OF_UNITY^(2^(28 - k))
            a = uniform_fr_dev(n, 0xDE06 + log_n)
            d = a.clone()
            dom.coeff_to_lagrange_dev(d)
            moved = not torch.equal(d, a)
            dom.lagrange_to_coeff_dev(d)
            ctx.sync()
            ok = bool(torch.equal(d, a)) and moved
            ms = ev_time(stream, lambda: ctx.best_fft_dev(d, dom.omega, log_n), 5)
            gbs = 64.0 * n / (ms * 1e-3) / 1e9
            muls = n / 2 * log_n
            rec = {"op": "ntt", "log_n": log_n, "ms": ms, "gb_s": gbs, "frac_hbm": gbs / HBM, "gmul_s": muls / (ms * 1e-3) / 1e9,
                   "frac_int_pipe": muls / (ms * 1e-3) / MUL_PEAK, "ok": ok,
                   "check": "lagrange_to_coeff(coeff_to_lagrange(a)) == a over the whole vector (device compare)"}
            if orc is not None and log_n <= args.cpu_to:
                host = a.cpu().numpy().view(np.uint64)
                t0 = time.perf_counter()
                want = orc.best_fft(host, dom.omega, log_n)
                rec["cpu_best_fft_ms"] = (time.perf_counter() - t0) * 1e3
                rec["cpu_cores"] = orc.ncpu()
                rec["cpu_gb_s"] = 64.0 * n / (rec["cpu_best_fft_ms"] * 1e-3) / 1e9
                g = a.clone()
                ctx.best_fft_dev(g, dom.omega, log_n)
                ctx.sync()
                rec["equals_cpu_best_fft"] = bool((g.cpu().numpy().view(np.uint64) == want).all())
                rec["speedup_vs_cpu"] = rec["cpu_best_fft_ms"] / ms
                del g
            print(json.dumps(rec), flush=True)
            dom.close()
            del d, a
        # ---------------- MSM
        for log_n in [16, 17, 18, 20, 22, 24, 26]:
            if log_n > args.msm_to:
                break
            if log_n < args.msm_from:
                continue
            n = 1 << log_n
            idx = np.zeros((n, 4), dtype=np.uint64)
            idx[:, 0] = np.arange(1, n + 1, dtype=np.uint64)
            d_idx = torch.empty((n, 4), dtype=torch.int64, device="cuda")
            for s0 in range(0, n, 1 << 22):
                d_idx[s0:s0 + (1 << 22)] = torch.from_numpy(ctx.fr_to_mont(idx[s0:s0 + (1 << 22)]).view(np.int64)).cuda()
            d_bases = torch.empty((n, 8), dtype=torch.int64, device="cuda")
            ctx.g1_mul_base_dev(G1_GEN_MONT, d_idx, n, d_bases)
            ctx.sync()
            del d_idx
            params = de_b200.ParamsKZG(log_n, None, d_bases.cpu().numpy().view(np.uint64), ctx)
            for kind in ("U", "W"):
                d_a = uniform_fr_dev(n, 0xDE04 + log_n) if kind == "U" else witness_fr_dev(ctx, n, 0xDE05 + log_n)
                d_b = uniform_fr_dev(n, 0xDE07 + log_n)
                # a + b in the field: element-wise through the library, in slices
                ab = np.empty((n, 4), dtype=np.uint64)
                ha, hb = d_a.cpu().numpy().view(np.uint64), d_b.cpu().numpy().view(np.uint64)
                for s0 in range(0, n, 1 << 22):
                    ab[s0:s0 + (1 << 22)] = ctx.fr_add(np.ascontiguousarray(ha[s0:s0 + (1 << 22)]), np.ascontiguousarray(hb[s0:s0 + (1 << 22)]))
                d_ab = torch.from_numpy(ab.view(np.int64)).cuda()
                ca = params.commit_batch_dev(1, d_a, n, 1)[0]
                cb = params.commit_batch_dev(1, d_b, n, 1)[0]
                cab = params.commit_batch_dev(1, d_ab, n, 1)[0]
                raw = ctx.best_multiexp_dev(d_a, d_bases, n)
                aff = ctx.batch_normalize(np.stack([ctx.g1_sum(np.stack([ca, cb])), cab, ca, raw]))
                ok = bool((aff[0] == aff[1]).all() and (aff[2] == aff[3]).all() and aff[2].any())
                ms_c = ev_time(stream, lambda: params.commit_batch_dev(1, d_a, n, 1), 3)
                ms_r = ev_time(stream, lambda: ctx.best_multiexp_dev(d_a, d_bases, n), 3)
                rec = {"op": "msm", "log_n": log_n, "scalars": kind, "commit_ms": ms_c, "commit_gpts_s": n / (ms_c * 1e-3) / 1e9,
                       "raw_ms": ms_r, "raw_gpts_s": n / (ms_r * 1e-3) / 1e9, "ok": ok,
                       "check": "commit(a) + commit(b) == commit(a + b); raw best_multiexp == table-based commit"}
                if orc is not None and log_n <= args.cpu_to:
                    bases_h = d_bases.cpu().numpy().view(np.uint64)
                    t0 = time.perf_counter()
                    want = orc.best_multiexp(ha, bases_h)
                    rec["cpu_best_multiexp_ms"] = (time.perf_counter() - t0) * 1e3
                    rec["cpu_cores"] = orc.ncpu()
                    rec["cpu_gpts_s"] = n / (rec["cpu_best_multiexp_ms"] * 1e-3) / 1e9
                    rec["equals_cpu_best_multiexp"] = bool((orc.g1_to_affine(want.reshape(1, 12)) == ctx.batch_normalize(ca.reshape(1, 12))).all())
                    rec["speedup_vs_cpu"] = rec["cpu_best_multiexp_ms"] / ms_c
                print(json.dumps(rec), flush=True)
                del d_a, d_b, d_ab
            params.close()
            del d_bases
    ctx.close()


if __name__ == "__main__":
    main()
