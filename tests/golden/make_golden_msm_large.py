#!/usr/bin/env python
"""Generates tests/golden/msm_large.json: the oracle's best_multiexp (oracle/oracle.c, the restated reference algorithm of
SURVEY.md Appendix B.1) at the BASELINE config-4 sizes too large to recompute inside the GPU test run (2^22, 2^24), for the
two scalar sets of SURVEY.md section 8d (U uniform, W witness-like) over the bases P_i = [i + 1] G.

    python tests/golden/make_golden_msm_large.py          # ~3 min on 8 cores

Each entry holds the canonical affine coordinates of the result and a SHA-256 over the inputs' raw bytes, so that the GPU
test (tests/test_gpu_msm_large.py) can prove it regenerated the same inputs before comparing the point."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import orc  # noqa: E402

SIZES = (22, 24)


def inputs(log_n: int, kind: str):
    """the config-4 inputs (shared with the GPU test): seeds 0xDE04 / 0xDE05 (+ log_n)"""
    n = 1 << log_n
    if kind == "U":
        return orc.uniform_fr(0xDE04 + log_n, n)
    return orc.witness_fr(0xDE05 + log_n, n, int(n * 0.77))


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {"generator": "tests/golden/make_golden_msm_large.py", "oracle": "oracle/oracle.c orc_best_multiexp", "cases": []}
    for log_n in SIZES:
        n = 1 << log_n
        bases = orc.gen_bases(n)
        for kind in ("U", "W"):
            s = inputs(log_n, kind)
            t0 = time.time()
            r = orc.best_multiexp(s, bases)
            dt = time.time() - t0
            aff = orc.g1_to_affine(r.reshape(1, 12))[0]
            xy = orc.fq_ints_from_mont(aff.reshape(2, 4))
            out["cases"].append({"log_n": log_n, "scalars": kind, "x": hex(xy[0]), "y": hex(xy[1]),
                                 "scalars_sha256": digest(s), "bases_sha256": digest(bases),
                                 "oracle_s": round(dt, 2), "oracle_threads": orc.ncpu()})
            print(out["cases"][-1], flush=True)
    with open(os.path.join(HERE, "msm_large.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
