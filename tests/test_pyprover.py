"""The restated create_proof / verify_proof (oracle/pyprover.py) and the pairing under it (oracle/pairing.py).

Pins available without the Rust crate: a single proof of the MainGate + RangeChip shape is
31 points + 58 scalars = 2848 bytes and one of the Poseidon-only shape 17 + 39 = 1792 bytes — the sizes the reference's
README |pi| figures decompose into (/root/reference/benches/README.md:56-63,89-99) — and the proofs
must be accepted by the verifier's final pairing check while any corrupted proof is rejected.
"""
import pytest

import pairing
import pyoracle as po
import pyprover as pp
from de_b200 import circuits, plonk


def test_pairing_bilinear_and_nondegenerate():
    assert pairing.g2_is_on_curve(pairing.G2_GEN)
    assert pairing.g2_mul(pairing.G2_GEN, po.FR - 1) == (pairing.G2_GEN[0], pairing.f2_neg(pairing.G2_GEN[1]))
    e1 = pairing.pairing(pairing.G2_GEN, po.G1_GEN)
    assert e1 != pairing.Fq12.one() and e1 ** po.FR == pairing.Fq12.one()
    a, b = 0x1234567, 0x7654321
    e2 = pairing.pairing(pairing.g2_mul(pairing.G2_GEN, b), po.g1_mul(po.G1_GEN, a))
    assert e2 == e1 ** (a * b)
    assert pairing.pairing_check([(pairing.G2_GEN, po.g1_mul(po.G1_GEN, a)),
                                  (pairing.g2_mul(pairing.G2_GEN, a), po.g1_neg(po.G1_GEN))])


def test_point_encoding_round_trip():
    for i in (1, 2, 3, 0xDEADBEEF):
        p = po.g1_mul(po.G1_GEN, i)
        assert pp.g1_from_bytes(pp.g1_to_bytes(p)) == p
        assert pp.g1_from_bytes(pp.g1_to_bytes(po.g1_neg(p))) == po.g1_neg(p)
    assert pp.g1_to_bytes(None) == bytes(32) and pp.g1_from_bytes(bytes(32)) is None


def test_permute_expression_pair_rule():
    inp = [5, 3, 5, 0, 0, 3, 5, 9]
    tab = [0, 3, 5, 9, 7, 7, 1, 2]
    a, s = pp.permute_expression_pair(inp, tab, 8)
    assert a == sorted(inp)
    assert sorted(s) == sorted(tab)
    for i in range(8):
        assert a[i] == s[i] or (i > 0 and a[i] == a[i - 1])
    # leftover table values (ascending: 1, 2, 7, 7) fill the repeated rows from the LAST repeated row backwards
    assert s == [0, 7, 3, 7, 5, 2, 1, 9]
    with pytest.raises(ValueError):
        pp.permute_expression_pair([4], [0], 1)


def test_kate_division_is_exact_quotient():
    rng = po.Xoshiro(5)
    a = [rng.uniform_fr() for _ in range(33)]
    b = rng.uniform_fr()
    a[0] = (a[0] - po.eval_poly(a, b)) % po.FR  # make b a root
    q = pp.kate_division(a, b)
    z = rng.uniform_fr()
    assert po.eval_poly(q, z) * (z - b) % po.FR == po.eval_poly(a, z)


@pytest.mark.parametrize("with_lookups,k,used,size", [(False, 5, 20, 1792), (True, 6, 40, 2848)])
def test_create_proof_verifies(with_lookups, k, used, size):
    asg = circuits.satisfied_assignment(with_lookups, k, 0xDE00 + k, used)
    circuits.check_assignment(asg)
    params = pp.setup(k, 0x1234567)
    q = pp.Queries(*plonk.collect_queries(asg.shape))
    assert len(q.advice) == 6 and len(q.fixed) == asg.shape.n_fixed and q.instance == [(0, 0)]
    pk = pp.keygen(params, asg.shape, q, asg.fixed, asg.copies, 0xABCDEF)
    proof = pp.create_proof(params, pk, asg.advice, asg.instances, po.Xoshiro(99).uniform_fr)
    assert len(proof) == size
    assert pp.verify_proof(params, pk.vk, asg.instances, proof)
    bad = bytearray(proof)
    bad[-40] ^= 1  # one of the opening evaluations... any bit flip must be rejected
    assert not pp.verify_proof(params, pk.vk, asg.instances, bytes(bad))
    # an unsatisfied witness gives a proof the verifier rejects
    adv = [list(c) for c in asg.advice]
    adv[0][1] = (adv[0][1] + 1) % po.FR
    try:
        proof2 = pp.create_proof(params, pk, adv, asg.instances, po.Xoshiro(99).uniform_fr)
    except ValueError:
        return  # a broken lookup input is caught by permute_expression_pair already
    assert not pp.verify_proof(params, pk.vk, asg.instances, proof2)


def test_public_inputs_are_bound():
    """instance column values enter the transcript and the permutation argument: the right inputs verify, others do not"""
    k = 6
    asg = circuits.satisfied_assignment(True, k, 0xDE16, 40, n_public=3)
    circuits.check_assignment(asg)
    assert len(asg.instances[0]) == 3
    params = pp.setup(k, 0x1234567)
    q = pp.Queries(*plonk.collect_queries(asg.shape))
    pk = pp.keygen(params, asg.shape, q, asg.fixed, asg.copies, 0xABCDEF)
    proof = pp.create_proof(params, pk, asg.advice, asg.instances, po.Xoshiro(5).uniform_fr)
    assert pp.verify_proof(params, pk.vk, asg.instances, proof)
    wrong = [list(asg.instances[0])]
    wrong[0][1] = (wrong[0][1] + 1) % po.FR
    assert not pp.verify_proof(params, pk.vk, wrong, proof)


def test_golden_proofs_reproduce():
    """tests/golden/golden_proofs_v1.json (make_golden_proofs.py): the restated prover reproduces the committed proof bytes"""
    import json
    import os
    sys_path = os.path.join(os.path.dirname(__file__), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_proofs", os.path.join(sys_path, "make_golden_proofs.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = json.load(open(os.path.join(sys_path, "golden_proofs_v1.json")))
    assert [c["name"] for c in gold["cases"]] == [c["name"] for c in mg.CASES]
    for c in gold["cases"]:
        asg, params, q, pk = mg.build(c)
        proof = pp.create_proof(params, pk, asg.advice, asg.instances, mg.draws_for(c))
        assert proof.hex() == c["proof"], c["name"]
        assert len(proof) == c["bytes"] == (2848 if c["with_lookups"] else 1792)


@pytest.mark.parametrize("with_lookups,k,used,n_public", [(False, 5, 20, 0), (True, 6, 40, 2), (True, 9, 300, 3)])
def test_array_based_prover_equals_integer_prover(with_lookups, k, used, n_public):
    """create_proof_fast (Montgomery arrays, C vector loops) is the same algorithm as create_proof (Python integers)"""
    asg = circuits.satisfied_assignment(with_lookups, k, 0xFA57 + k, used, n_public=n_public)
    params = pp.setup(k, 0x1234567)
    q = pp.Queries(*plonk.collect_queries(asg.shape))
    pk = pp.keygen(params, asg.shape, q, asg.fixed, asg.copies, 0xABCDEF)
    rng = po.Xoshiro(99)
    draws = [rng.uniform_fr() for _ in range(pp.random_count(asg.shape, 1 << k))]
    it = iter(draws)
    a = pp.create_proof(params, pk, asg.advice, asg.instances, lambda: next(it))
    b = pp.create_proof_fast(params, pk, asg.advice, asg.instances, pp.to_mont(draws))
    assert a == b


def test_second_shape_degree4_fixed_in_permutation():
    """no instance column, one degree-3 gate, one single-expression lookup (cs.degree() = 4), a fixed column in the permutation"""
    k = 6
    asg = circuits.mul_table_assignment(k, 0xB2, 40)
    circuits.check_assignment(asg)
    assert asg.shape.degree() == 4 and asg.shape.n_perm_sets == 2 and asg.copies
    params = pp.setup(k, 0x1234567)
    q = pp.Queries(*plonk.collect_queries(asg.shape))
    pk = pp.keygen(params, asg.shape, q, asg.fixed, asg.copies, 7)
    draws = [po.Xoshiro(3).uniform_fr() for _ in range(pp.random_count(asg.shape, 1 << k))]
    it = iter(draws)
    proof = pp.create_proof(params, pk, asg.advice, asg.instances, lambda: next(it))
    assert proof == pp.create_proof_fast(params, pk, asg.advice, asg.instances, pp.to_mont(draws))
    # points: 3 advice + 2 lookup + 2 perm z + 1 lookup z + 1 random + 3 h + openings; evals: see Appendix C's formula
    assert pp.verify_proof(params, pk.vk, asg.instances, proof)
    bad = bytearray(proof)
    bad[100] ^= 4
    assert not pp.verify_proof(params, pk.vk, asg.instances, bytes(bad))
