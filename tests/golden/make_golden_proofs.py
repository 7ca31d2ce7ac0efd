"""Generates tests/golden/golden_proofs_v1.json: whole proofs from the CPU restatement of create_proof (oracle/pyprover.py) on
seeded, satisfied circuits of the two constraint-system shapes of the reference (MainGate only: pose_enc; MainGate +
RangeChip: mod_pow / delay_enc).  Every input is derived from the recorded seeds: the circuit (de_b200/circuits.py), the SRS
secret, vk.transcript_repr and the stream of Fr::random draws (xoshiro256**, oracle/pyoracle.py).  The reference itself cannot
run here (no Rust toolchain), so these pin the CUDA prover and the restatement to each other across code changes; every
golden proof is also checked by the restated verifier at generation time.

Run:  python tests/golden/make_golden_proofs.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, "..", "..", "delay-encryption-in-halo2_b200"))
import pyoracle as po  # noqa: E402
import pyprover as pp  # noqa: E402
from de_b200 import circuits, plonk  # noqa: E402

CASES = [
    dict(name="maingate_k5", with_lookups=False, k=5, used=20, seed=0x601D05, n_public=0),
    dict(name="range_k6", with_lookups=True, k=6, used=40, seed=0x601D06, n_public=0),
    dict(name="range_k7_public", with_lookups=True, k=7, used=100, seed=0x601D07, n_public=4),
]


def build(case):
    asg = circuits.satisfied_assignment(case["with_lookups"], case["k"], case["seed"], case["used"], n_public=case["n_public"])
    params = pp.setup(case["k"], case["seed"] ^ 0x5EC2E7)
    q = pp.Queries(*plonk.collect_queries(asg.shape))
    pk = pp.keygen(params, asg.shape, q, asg.fixed, asg.copies, case["seed"] + 1)
    return asg, params, q, pk


def draws_for(case, count=None):
    rng = po.Xoshiro(case["seed"] ^ 0x99)
    if count is None:
        return rng.uniform_fr
    return [rng.uniform_fr() for _ in range(count)]


def main():
    out = {"version": 1, "note": "proofs as hex; see make_golden_proofs.py for how every input derives from the seeds", "cases": []}
    for case in CASES:
        asg, params, q, pk = build(case)
        proof = pp.create_proof(params, pk, asg.advice, asg.instances, draws_for(case))
        assert pp.verify_proof(params, pk.vk, asg.instances, proof)
        out["cases"].append(dict(case, proof=proof.hex(), sha256=hashlib.sha256(proof).hexdigest(), bytes=len(proof)))
    with open(os.path.join(HERE, "golden_proofs_v1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "proofs")


if __name__ == "__main__":
    main()
