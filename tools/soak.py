#!/usr/bin/env python
"""Stability check: N proofs back to back on one prover (delay_enc shape, k = 14 so that a proof takes ~4 ms); every proof must be
byte-identical to the first, host RSS and device memory must not grow."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))
import de_b200  # noqa: E402
from de_b200 import circuits, keygen, synth  # noqa: E402


def rss_mb():
    with open("/proc/self/status") as f:
        for line in f:
            if line.startswith("VmRSS"):
                return int(line.split()[1]) / 1024.0
    return 0.0


def main():
    n_proofs = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    k = 14
    asg = circuits.satisfied_assignment(True, k, 0x50A4, 12000)
    n = 1 << k
    ctx = de_b200.Context(0)
    keys = keygen.keygen(ctx, asg.shape, k, synth.gen_bases(n, 0), synth.gen_bases(n, n), asg.fixed, asg.copies, 0x50A4)
    adv = torch.from_numpy(np.stack([ctx.fr_to_mont(keygen.canonical_limbs(c)) for c in asg.advice]).view(np.int64)).cuda()
    rnd = torch.from_numpy(synth.uniform_fr(1, keys.prover.random_count).view(np.int64)).cuda()
    first = keys.prover.create_proof_dev(adv, rnd)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    r0 = rss_mb()
    t0 = time.time()
    for i in range(n_proofs):
        p = keys.prover.create_proof_dev(adv, rnd)
        assert p == first, f"proof {i} differs"
    dt = time.time() - t0
    free1, _ = torch.cuda.mem_get_info()
    print(f"{n_proofs} proofs, {1e3 * dt / n_proofs:.2f} ms each, all identical; host RSS {r0:.0f} -> {rss_mb():.0f} MB; "
          f"device free {free0 >> 20} -> {free1 >> 20} MiB; launches {ctx.launches}")
    keys.close()
    ctx.close()


if __name__ == "__main__":
    main()
