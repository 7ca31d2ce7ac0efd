"""Generates tests/golden/golden_v1.json from the definition-level Python oracle (oracle/pyoracle.py): naive DFT,
double-and-add MSM, EvaluationDomain by direct polynomial evaluation.  The reference itself cannot run here (no Rust
toolchain, un-vendored halo2_proofs), so these vectors pin the C oracle and the CUDA path to the mathematical definition
of each operation; the Poseidon vectors inside pyoracle.py are the reference's own (permutation.rs:154-158,190-196).

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyoracle as po  # noqa: E402


def hx(v):
    return hex(v)


def main():
    out = {"version": 1, "note": "canonical (non-Montgomery) integers as hex"}
    # NTT by the definition A[j] = sum a[i] w^(ij)
    ntt = []
    for log_n, seed in ((1, 0xA1), (3, 0xA3), (5, 0xA5), (7, 0xA7)):
        rng = po.Xoshiro(seed)
        n = 1 << log_n
        a = [rng.uniform_fr() for _ in range(n)]
        omega = pow(po.FR_ROOT_OF_UNITY, 1 << (po.FR_S - log_n), po.FR)
        ntt.append({"log_n": log_n, "omega": hx(omega), "in": [hx(x) for x in a], "out": [hx(x) for x in po.dft_naive(a, omega)]})
    out["ntt"] = ntt
    # MSM by double-and-add per term
    msm = []
    for n, seed in ((1, 0xB1), (5, 0xB5), (33, 0xB33), (100, 0xB100)):
        rng = po.Xoshiro(seed)
        s = [rng.uniform_fr() for _ in range(n)]
        if n >= 5:
            s[1] = 0
            s[2] = 1
            s[3] = po.FR - 1
        bases = [po.g1_mul(po.G1_GEN, 7 * i + 3) for i in range(n)]
        r = po.msm_naive(s, bases)
        msm.append({"n": n, "scalars": [hx(x) for x in s], "bases": [[hx(b[0]), hx(b[1])] for b in bases],
                    "result": None if r is None else [hx(r[0]), hx(r[1])]})
    out["msm"] = msm
    # EvaluationDomain: coeff_to_extended by evaluating the polynomial on the zeta-coset
    dom = []
    for j, k, seed in ((3, 3, 0xC3), (5, 4, 0xC4)):
        d = po.EvaluationDomain(j, k)
        rng = po.Xoshiro(seed)
        a = [rng.uniform_fr() for _ in range(d.n)]
        ext = [po.eval_poly(a, po.FR_ZETA * pow(d.extended_omega, i, po.FR) % po.FR) for i in range(d.extended_n)]
        lag = [po.eval_poly(a, pow(d.omega, i, po.FR)) for i in range(d.n)]
        dom.append({"j": j, "k": k, "extended_k": d.extended_k, "omega": hx(d.omega), "extended_omega": hx(d.extended_omega),
                    "coeff": [hx(x) for x in a], "extended": [hx(x) for x in ext], "lagrange": [hx(x) for x in lag],
                    "t_evaluations": [hx(x) for x in d.t_evaluations]})
    out["domain"] = dom
    with open(os.path.join(HERE, "golden_v1.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote golden_v1.json")


if __name__ == "__main__":
    main()
