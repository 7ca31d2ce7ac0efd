"""a9 on the CPU: the C restatement of Evaluator::evaluate_h (compiled GraphEvaluator programs, rotation indices, coset
NTTs) against the definition-level Python evaluation (expression trees, Horner at coset points)."""
import numpy as np
import pytest

import orc
import pyoracle as po
from de_b200 import plonk


def random_instance(shape, k, seed):
    n = 1 << k
    def polys(count, s):
        return [orc.uniform_fr(seed * 1000 + s * 50 + i, n) for i in range(count)]
    return dict(fixed=polys(shape.n_fixed, 1), sigma=polys(len(shape.perm_columns), 2), advice=polys(shape.n_advice, 3),
                instance=polys(shape.n_instance, 4), permz=polys(shape.n_perm_sets, 5), lookup=polys(3 * len(shape.lookups), 6))


def run_oracle(shape, k, inst, ch):
    dom = orc.Domain(shape.degree(), k, threads=4)
    desc, keep = plonk.marshal_pk_desc(shape, inst["fixed"], inst["sigma"])
    chs, keep2 = plonk.marshal_challenges(*ch)
    return dom, orc.evaluate_h(dom, desc, inst["advice"], inst["instance"], chs, inst["permz"], inst["lookup"])


@pytest.mark.parametrize("with_lookups,k", [(False, 4), (True, 4), (True, 5)])
def test_oracle_evaluate_h_matches_definition(with_lookups, k):
    shape = plonk.main_gate_shape(with_lookups)
    assert shape.degree() == (5 if with_lookups else 3)
    assert shape.chunk_len == (3 if with_lookups else 1) and shape.n_perm_sets == (2 if with_lookups else 6)
    inst = random_instance(shape, k, 7 + k)
    ch = (0x1111 + k, 0x2222, 0x3333, 0x4444)
    dom, got = run_oracle(shape, k, inst, ch)
    pd = po.EvaluationDomain(shape.degree(), k)
    assert pd.extended_k == dom.extended_k
    ints = {name: [orc.fr_ints_from_mont(p) for p in v] for name, v in inst.items()}
    rows = list(range(0, pd.extended_n, 7)) + [pd.extended_n - 1, 1, 2]
    want = po.evaluate_h_definition(pd, shape.gates, shape.lookups, shape.perm_columns, shape.chunk_len, shape.blinding_factors,
                                    ints["fixed"], ints["sigma"], ints["advice"], ints["instance"], ints["permz"], ints["lookup"],
                                    *ch, rows=rows)
    got_ints = orc.fr_ints_from_mont(got)
    for idx in rows:
        assert got_ints[idx] == want[idx], idx


def test_graph_compiler_shares_subexpressions():
    g = plonk.compile_gates(plonk.main_gate_shape(False).gates)
    # 14 column reads, 9 products for the 7 multiplicative terms (a*b and c*d shared once each), 8 additions, 1 horner
    ops = [c[0] for c in g.calculations]
    assert ops.count(plonk.STORE) == 15 and ops.count(plonk.HORNER) == 1
    assert g.rotations == [0, 1] and g.constants[:3] == [0, 1, 2]
    assert g.num_intermediates == len(g.calculations)
