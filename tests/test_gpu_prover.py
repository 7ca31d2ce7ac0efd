"""create_proof on the GPU (csrc/prover.cu through the C ABI: de_prover_create / de_create_proof) against the CPU
restatement (oracle/pyprover.py) on satisfied circuits of the reference's bench shapes
(/root/reference/benches/pose_enc.rs:184 k = 11, delay_enc.rs:181 k = 16, mod_pow.rs:258 k = 17):

  * the proof BYTES are identical for the same witness, the same vk.transcript_repr and the same stream of Fr::random draws;
  * the proof passes the restated verify_proof (Blake2bRead, VerifierGWC, pairing check);
  * the GPU prover consumes exactly the number of random draws the CPU restatement consumes.
Also the opening-phase kernels on their own: eval_polynomial, kate_division, transcript-form commitments.
"""
import numpy as np
import pytest

import de_b200
import orc
import pyoracle as po
import pyprover as pp
from de_b200 import circuits, plonk

pytestmark = pytest.mark.gpu


def mont(vals):
    return pp.to_mont(vals)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 256, 257, 1000, 8191, 8192, (1 << 14) + 3, (1 << 16) + 1])
def test_eval_polynomial_and_kate_division(ctx, n):
    rng = po.Xoshiro(0xE0 + n)
    a = [rng.uniform_fr() for _ in range(n)]
    b = rng.uniform_fr()
    got = pp.from_mont(ctx.eval_polynomial(mont(a), mont([b])[0]).reshape(1, 4))[0]
    assert got == po.eval_poly(a, b)
    if n >= 2:
        q = pp.from_mont(ctx.kate_division(mont(a), mont([b])[0]))
        assert q == pp.kate_division(a, b)


def test_kate_division_special_points(ctx):
    rng = po.Xoshiro(77)
    a = [rng.uniform_fr() for _ in range(100)]
    for b in (0, 1, po.FR - 1):
        q = pp.from_mont(ctx.kate_division(mont(a), mont([b])[0]))
        assert q == pp.kate_division(a, b)
    with pytest.raises(de_b200.DeError):
        ctx.kate_division(mont(a[:1]), mont([5])[0])


def test_commit_canonical_matches_oracle(ctx):
    import torch
    k = 8
    n = 1 << k
    bases = orc.gen_bases(n)
    params = de_b200.ParamsKZG(k, g=bases, g_lagrange=bases, ctx=ctx)
    cols = np.stack([orc.uniform_fr(900 + i, n) for i in range(3)] + [np.zeros((n, 4), dtype=np.uint64)])
    d = torch.from_numpy(cols.view(np.int64)).cuda()
    got = params.commit_batch_canonical_dev(1, d, n, 4)
    for i in range(4):
        aff = orc.g1_to_affine(orc.best_multiexp(cols[i], bases))
        x, y = orc.fq_ints_from_mont(aff.reshape(-1, 4))
        assert got[i].tobytes() == pp.fq_to_repr(x) + pp.fq_to_repr(y)
    params.close()


def run_both(with_lookups, k, used, seed, check_bytes=True, n_public=0, asg=None):
    asg = asg or circuits.satisfied_assignment(with_lookups, k, seed, used, n_public=n_public)
    shape = asg.shape
    n = 1 << k
    oparams = pp.setup(k, 0x5EC2E7 + k)
    q = pp.Queries(*plonk.collect_queries(shape))
    repr_ = 0xABCDEF0123456789 + k
    opk = pp.keygen(oparams, shape, q, asg.fixed, asg.copies, repr_)
    ctx = de_b200.Context(0)
    params = de_b200.ParamsKZG(k, oparams.g_mont, oparams.g_lagrange_mont, ctx)
    dom = de_b200.EvaluationDomain(shape.degree(), k, ctx)
    pk = plonk.ProvingKey(dom, shape, [mont(p) for p in opk.fixed_polys], [mont(p) for p in opk.sigma_polys])
    prover = plonk.Prover(params, pk, q.advice, q.fixed, repr_)
    rng = po.Xoshiro(seed ^ 0x99)
    draws = [rng.uniform_fr() for _ in range(prover.random_count)]
    proof = prover.create_proof([mont(c) for c in asg.advice], [mont(v) if v else np.zeros((0, 4), dtype=np.uint64) for v in asg.instances],
                                mont(draws))
    assert len(proof) == prover.proof_size
    it = iter(draws)
    used_draws = [0]

    def next_random():
        used_draws[0] += 1
        return next(it)

    want = None
    if check_bytes and k >= 14:
        # the array-based restatement (byte-identical to the integer one: tests/test_pyprover.py) keeps the large cases short
        assert pp.random_count(shape, n) == prover.random_count
        want = pp.create_proof_fast(oparams, opk, asg.advice, asg.instances, mont(draws))
    elif check_bytes:
        want = pp.create_proof(oparams, opk, asg.advice, asg.instances, next_random)
        assert used_draws[0] == prover.random_count, "the GPU prover and the restated prover disagree on the number of RNG draws"
    ok = pp.verify_proof(oparams, opk.vk, asg.instances, proof)
    # a second proof with the same inputs is byte-identical (no state leaks between proofs)
    proof2 = prover.create_proof([mont(c) for c in asg.advice], [mont(v) if v else np.zeros((0, 4), dtype=np.uint64) for v in asg.instances],
                                 mont(draws))
    prover.close(); pk.close(); dom.close(); params.close(); ctx.close()
    return proof, want, ok, proof2


def first_diff(a, b):
    for i in range(0, min(len(a), len(b)), 32):
        if a[i:i + 32] != b[i:i + 32]:
            return i // 32
    return None


@pytest.mark.parametrize("name,with_lookups,k,used,size", [("tiny-maingate", False, 5, 20, 1792), ("tiny-range", True, 6, 40, 2848),
                                                           ("small-range", True, 10, 900, 2848), ("pose_enc", False, 11, 1450, 1792)])
def test_create_proof_bytes_match_restated_prover(name, with_lookups, k, used, size):
    proof, want, ok, proof2 = run_both(with_lookups, k, used, 0xDE00 + k)
    assert len(proof) == size
    assert proof == want, f"{name}: first differing 32-byte proof element: #{first_diff(proof, want)}"
    assert ok, "the restated verifier rejects the GPU proof"
    assert proof2 == proof


def test_create_proof_with_public_inputs():
    """non-empty instance column: values hashed into the transcript, instance polynomial in the permutation argument"""
    proof, want, ok, proof2 = run_both(True, 8, 200, 0xDE18, n_public=5)
    assert proof == want, f"first differing 32-byte proof element: #{first_diff(proof, want)}"
    assert ok and proof2 == proof


@pytest.mark.parametrize("k,used", [(6, 40), (10, 800)])
def test_create_proof_second_shape(k, used):
    """circuits.mul_table_assignment: no instance column, cs.degree() = 4 (three h pieces, chunks of 2), one lookup with single
    expressions, a fixed column inside the permutation"""
    asg = circuits.mul_table_assignment(k, 0xB200 + k, used)
    proof, want, ok, proof2 = run_both(None, k, used, 0xB200 + k, asg=asg)
    assert proof == want, f"first differing 32-byte proof element: #{first_diff(proof, want)}"
    assert ok and proof2 == proof


def test_create_proof_delay_enc_shape_k16():
    proof, want, ok, proof2 = run_both(True, 16, 50400, 0xDE03)
    assert len(proof) == 2848
    assert proof == want, f"first differing 32-byte proof element: #{first_diff(proof, want)}"
    assert ok
    assert proof2 == proof


def test_create_proof_mod_pow_shape_k17():
    """/root/reference/benches/mod_pow.rs:258: K = 17 although the circuit's 41 766 rows would fit k = 16"""
    proof, want, ok, proof2 = run_both(True, 17, 41766, 0xDE01)
    assert len(proof) == 2848
    assert proof == want, f"first differing 32-byte proof element: #{first_diff(proof, want)}"
    assert ok
    assert proof2 == proof


def test_unsatisfied_lookup_is_reported():
    k = 6
    asg = circuits.satisfied_assignment(True, k, 0xDE06, 40)
    shape = asg.shape
    # break a range row: a sub-limb outside the table
    row = next(r for r in range(40) if asg.fixed[13][r] == 1)
    asg.advice[0][row] = 12345
    oparams = pp.setup(k, 0x5EC2E7)
    q = pp.Queries(*plonk.collect_queries(shape))
    opk = pp.keygen(oparams, shape, q, asg.fixed, asg.copies, 1)
    ctx = de_b200.Context(0)
    params = de_b200.ParamsKZG(k, oparams.g_mont, oparams.g_lagrange_mont, ctx)
    dom = de_b200.EvaluationDomain(shape.degree(), k, ctx)
    pk = plonk.ProvingKey(dom, shape, [mont(p) for p in opk.fixed_polys], [mont(p) for p in opk.sigma_polys])
    prover = plonk.Prover(params, pk, q.advice, q.fixed, 1)
    draws = mont([po.Xoshiro(1).uniform_fr() for _ in range(prover.random_count)])
    with pytest.raises(de_b200.DeError, match="ConstraintSystemFailure"):
        prover.create_proof([mont(c) for c in asg.advice], [np.zeros((0, 4), dtype=np.uint64)], draws)
    with pytest.raises(ValueError):
        pp.create_proof(oparams, opk, asg.advice, asg.instances, po.Xoshiro(1).uniform_fr)
    prover.close(); pk.close(); dom.close(); params.close(); ctx.close()


def test_golden_proofs_on_gpu():
    """the committed golden proofs (tests/golden/golden_proofs_v1.json) are reproduced byte for byte by the CUDA prover"""
    import importlib.util
    import json
    import os
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_proofs", os.path.join(gdir, "make_golden_proofs.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = json.load(open(os.path.join(gdir, "golden_proofs_v1.json")))
    for c in gold["cases"]:
        asg, oparams, q, opk = mg.build(c)
        ctx = de_b200.Context(0)
        params = de_b200.ParamsKZG(c["k"], oparams.g_mont, oparams.g_lagrange_mont, ctx)
        dom = de_b200.EvaluationDomain(asg.shape.degree(), c["k"], ctx)
        pk = plonk.ProvingKey(dom, asg.shape, [mont(p) for p in opk.fixed_polys], [mont(p) for p in opk.sigma_polys])
        prover = plonk.Prover(params, pk, q.advice, q.fixed, c["seed"] + 1)
        draws = mg.draws_for(c, prover.random_count)
        proof = prover.create_proof([mont(col) for col in asg.advice], [mont(v) if v else np.zeros((0, 4), dtype=np.uint64) for v in asg.instances],
                                    mont(draws))
        assert proof.hex() == c["proof"], c["name"]
        prover.close(); pk.close(); dom.close(); params.close(); ctx.close()


def test_prover_argument_errors():
    """the asserts / Err results of the Rust API map to DE_ERR_ARG: wrong context pairing, short random stream, oversized public
    inputs (Error::InstanceTooLarge)"""
    k = 5
    asg = circuits.satisfied_assignment(False, k, 0xDE25, 20)
    oparams = pp.setup(k, 0x5EC2E7)
    q = pp.Queries(*plonk.collect_queries(asg.shape))
    opk = pp.keygen(oparams, asg.shape, q, asg.fixed, asg.copies, 1)
    ctx, ctx2 = de_b200.Context(0), de_b200.Context(0)
    params = de_b200.ParamsKZG(k, oparams.g_mont, oparams.g_lagrange_mont, ctx)
    params2 = de_b200.ParamsKZG(k, oparams.g_mont, oparams.g_lagrange_mont, ctx2)
    dom = de_b200.EvaluationDomain(asg.shape.degree(), k, ctx)
    pk = plonk.ProvingKey(dom, asg.shape, [mont(p) for p in opk.fixed_polys], [mont(p) for p in opk.sigma_polys])
    with pytest.raises(de_b200.DeError, match="same context"):
        plonk.Prover(params2, pk, q.advice, q.fixed, 1)
    prover = plonk.Prover(params, pk, q.advice, q.fixed, 1)
    draws = mont([po.Xoshiro(2).uniform_fr() for _ in range(prover.random_count)])
    adv = [mont(c) for c in asg.advice]
    with pytest.raises(de_b200.DeError, match="random"):
        prover.create_proof(adv, [np.zeros((0, 4), dtype=np.uint64)], draws[:-1])
    with pytest.raises(de_b200.DeError, match="InstanceTooLarge"):
        prover.create_proof(adv, [np.zeros((1 << k, 4), dtype=np.uint64)], draws)
    # still usable after the errors
    proof = prover.create_proof(adv, [np.zeros((0, 4), dtype=np.uint64)], draws)
    assert len(proof) == 1792 and pp.verify_proof(oparams, opk.vk, asg.instances, proof)
    prover.close(); pk.close(); dom.close(); params.close(); params2.close(); ctx.close(); ctx2.close()


@pytest.mark.parametrize("seed", range(0x5EED00, 0x5EED0C))
def test_create_proof_many_seeds(seed):
    """different witnesses, copy-constraint graphs, table hit patterns and random streams: shapes alternate between MainGate,
    MainGate + RangeChip (with public inputs) and the degree-4 shape; k and the number of used rows vary with the seed"""
    import random
    rng = random.Random(seed)
    kind = seed % 3
    k = rng.choice([6, 7, 8, 9])
    used = rng.randrange(20, (1 << k) - 6)
    if kind == 2:
        asg = circuits.mul_table_assignment(k, seed, used)
    else:
        asg = circuits.satisfied_assignment(kind == 1, k, seed, used, uniform_values=bool(seed & 4), n_public=rng.randrange(0, 4))
    circuits.check_assignment(asg)
    proof, want, ok, proof2 = run_both(None, k, used, seed, asg=asg)
    assert proof == want, f"first differing 32-byte proof element: #{first_diff(proof, want)}"
    assert ok and proof2 == proof
