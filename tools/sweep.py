#!/usr/bin/env python
"""BASELINE config 4: BN254 G1 MSM and Fr NTT size sweeps on one B200, device-resident, CUDA-event timed.

  MSM: ParamsKZG.commit_lagrange (bases + window tables resident) and raw best_multiexp, uniform ("U") and witness-like
       ("W") scalars; Gpts/s = N / t.  Checked against the CPU oracle up to 2^20 and by linearity
       commit(a) + commit(b) == commit(a + b) above.
  NTT: best_fft in place; GB/s = 64 * N / t (algorithmic bytes) and the fraction of the integer-pipe ceiling
       (N/2 * log2 N multiplies at the measured 65.9 Gmul/s).  Checked against the oracle up to 2^20 and by
       inverse(forward(a)) / N == a on a sample above.

Writes one JSON line per measurement (profiles/r01_sweep.jsonl is a committed run).  The oracle is used ONLY as the checker.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import de_b200  # noqa: E402
import orc  # noqa: E402
import pyoracle as po  # noqa: E402

MUL_PEAK = 65.9e9
HBM = 6529.7


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--msm-max", type=int, default=22)
    ap.add_argument("--ntt-max", type=int, default=26)
    ap.add_argument("--check-max", type=int, default=20)
    args = ap.parse_args()
    stream = torch.cuda.Stream()
    ctx = de_b200.Context(0)
    ctx.set_stream(stream.cuda_stream)
    as_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    with torch.cuda.stream(stream):
        # ---------------- NTT
        for log_n in range(16, args.ntt_max + 1):
            n = 1 << log_n
            a = orc.uniform_fr(0xDE06, n)
            w = orc.fr_mont_from_ints([pow(po.FR_ROOT_OF_UNITY, 1 << (po.FR_S - log_n), po.FR)])[0]
            d = as_dev(a)
            ctx.best_fft_dev(d, w, log_n)
            ctx.sync()
            got = d.cpu().numpy().view(np.uint64)
            if log_n <= args.check_max:
                ok = bool((got == orc.best_fft(a, w, log_n)).all())
                how = "full compare with the oracle"
            else:
                w_inv = orc.fr_inv(w.reshape(1, 4))[0]
                ctx.best_fft_dev(d, w_inv, log_n)
                ctx.sync()
                back = d.cpu().numpy().view(np.uint64)
                idx = np.random.default_rng(1).integers(0, n, 4096)
                n_inv = orc.fr_inv(orc.fr_mont_from_ints([n]))
                ok = bool((orc.fr_mul(back[idx], np.repeat(n_inv, len(idx), axis=0)) == a[idx]).all())
                how = "inverse(forward(a)) / N == a on 4096 sampled positions"
            ms = ev_time(stream, lambda: ctx.best_fft_dev(d, w, log_n), 5)
            gbs = 64.0 * n / (ms * 1e-3) / 1e9
            muls = n / 2 * log_n
            print(json.dumps({"op": "ntt", "log_n": log_n, "ms": ms, "gb_s": gbs, "frac_hbm": gbs / HBM,
                              "gmul_s": muls / (ms * 1e-3) / 1e9, "frac_int_pipe": muls / (ms * 1e-3) / MUL_PEAK, "ok": ok, "check": how}),
                  flush=True)
            del d
        # ---------------- MSM
        for log_n in [16, 17, 18, 20, 22, 24, 26]:
            if log_n > args.msm_max:
                break
            n = 1 << log_n
            t0 = time.time()
            bases = orc.gen_bases(n)
            t_gen = time.time() - t0
            params = de_b200.ParamsKZG(log_n, None, bases, ctx)
            d_bases = as_dev(bases)
            for dist in ("U", "W"):
                s = orc.uniform_fr(0xDE04, n) if dist == "U" else orc.witness_fr(0xDE05, n, int(n * 0.77))
                d_s = as_dev(s)
                got = params.commit_batch_dev(1, d_s, n, 1)[0]
                raw = ctx.best_multiexp_dev(d_s, d_bases, n)
                same_paths = bool((ctx.batch_normalize(np.stack([got, raw]))[0] == ctx.batch_normalize(np.stack([got, raw]))[1]).all())
                if log_n <= args.check_max:
                    ok = bool((ctx.batch_normalize(got.reshape(1, 12)) == orc.g1_to_affine(orc.best_multiexp(s, bases))).all())
                    how = "commit == oracle best_multiexp (affine)"
                else:
                    s2 = orc.uniform_fr(0xDE07, n)
                    c2 = params.commit_batch_dev(1, as_dev(s2), n, 1)[0]
                    c3 = params.commit_batch_dev(1, as_dev(orc.fr_add(s, s2)), n, 1)[0]
                    lhs = ctx.g1_sum(np.stack([got, c2]))
                    ok = bool((ctx.batch_normalize(np.stack([lhs, c3]))[0] == ctx.batch_normalize(np.stack([lhs, c3]))[1]).all())
                    how = "commit(a) + commit(b) == commit(a + b)"
                ms_c = ev_time(stream, lambda: params.commit_batch_dev(1, d_s, n, 1), 3)
                ms_r = ev_time(stream, lambda: ctx.best_multiexp_dev(d_s, d_bases, n), 3)
                print(json.dumps({"op": "msm", "log_n": log_n, "scalars": dist, "commit_ms": ms_c, "commit_gpts_s": n / (ms_c * 1e-3) / 1e9,
                                  "raw_ms": ms_r, "raw_gpts_s": n / (ms_r * 1e-3) / 1e9, "ok": ok and same_paths, "check": how,
                                  "bases_gen_s": t_gen}), flush=True)
                del d_s
            params.close()
            del d_bases
    ctx.close()


if __name__ == "__main__":
    main()
