"""a9 parity (GPU): Evaluator::evaluate_h through the C ABI == the C oracle, bit-exact over the whole extended domain, for
the MainGate (pose_enc) and MainGate + RangeChip (mod_pow / delay_enc) constraint-system shapes."""
import numpy as np
import pytest

import de_b200
import orc
from de_b200 import plonk
from test_evaluator_oracle import random_instance, run_oracle

pytestmark = pytest.mark.gpu


def gpu_eval(ctx, shape, k, inst, ch):
    dom = de_b200.EvaluationDomain(shape.degree(), k, ctx)
    pk = plonk.ProvingKey(dom, shape, inst["fixed"], inst["sigma"])
    return dom, pk, pk.evaluate_h(inst["advice"], inst["instance"], *ch, inst["permz"], inst["lookup"])


@pytest.mark.parametrize("with_lookups,k", [(False, 4), (True, 4), (False, 11), (True, 9), (True, 12)])
def test_evaluate_h_matches_oracle(ctx, with_lookups, k):
    shape = plonk.main_gate_shape(with_lookups)
    inst = random_instance(shape, k, 31 + k)
    ch = (0xABCDEF + k, 0x1234567, 0x89ABCD, 0x55AA55)
    _, want = run_oracle(shape, k, inst, ch)
    _, _, got = gpu_eval(ctx, shape, k, inst, ch)
    assert (got == want).all()


def test_evaluate_h_delay_enc_size(ctx):
    # bench configuration: k = 16, degree 5 (extended domain 2^18), 15 fixed, 5 lookups
    shape = plonk.main_gate_shape(True)
    k = 16
    inst = random_instance(shape, k, 99)
    ch = (3, 5, 7, 11)
    odom, want = run_oracle(shape, k, inst, ch)
    dom, pk, got = gpu_eval(ctx, shape, k, inst, ch)
    assert dom.extended_k == 18
    assert (got == want).all()
    # the pipeline that follows in create_proof: divide by the vanishing polynomial and go back to coefficients
    h = dom.divide_by_vanishing_poly(got)
    assert (h == odom.divide_by_vanishing(want)).all()
    assert (dom.extended_to_coeff(h) == odom.extended_to_coeff(h)).all()


def test_pk_upload_rejects_wrong_degree(ctx):
    shape = plonk.main_gate_shape(True)
    dom = de_b200.EvaluationDomain(3, 6, ctx)
    inst = random_instance(shape, 6, 1)
    with pytest.raises(ValueError):
        plonk.ProvingKey(dom, shape, inst["fixed"], inst["sigma"])
