"""Regenerates tests/golden/frontend_v1.json: digests of what the circuit front-end emits for seeded inputs (fixed columns,
advice columns, copy constraints, results, used rows) for the reference's three bench circuits.  A regression pin of THIS
repository's row layout (keys made from one build must fit witnesses made by another), not a pin against halo2wrong's layout,
which is not recoverable here.  Usage: python tests/golden/make_golden_frontend.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))


def digest(a) -> str:
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cases():
    from de_b200 import frontend as fe
    n, e, x = fe.sample_rsa_inputs(0xDE03)
    yield "delay_enc_k16_seed_de03", fe.delay_enc(n, e, x, [0, 0], k=16)
    n, e, x = fe.sample_rsa_inputs(0xDE01)
    yield "mod_pow_k17_seed_de01", fe.mod_pow(n, e, x, k=17)
    yield "pose_enc_k11_key_5eed", fe.pose_enc((0x5EED1, 0x5EED2), [0, 0], k=11)


def describe(syn) -> dict:
    return {"k": syn.k, "used_rows": syn.used_rows, "n_fixed": int(syn.fixed.shape[0]), "n_copies": int(len(syn.copies)),
            "outputs": [hex(v) for v in syn.outputs], "fixed_sha256": digest(syn.fixed), "advice_sha256": digest(syn.advice),
            "copies_sha256": digest(syn.copies)}


if __name__ == "__main__":
    out = {name: describe(syn) for name, syn in cases()}
    with open(os.path.join(ROOT, "tests", "golden", "frontend_v1.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print({k: v["used_rows"] for k, v in out.items()})
