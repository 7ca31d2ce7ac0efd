"""a1 parity (GPU): device Fr/Fq arithmetic through the C ABI == oracle, bit-exact."""
import numpy as np
import pytest

import orc
import pyoracle as po

pytestmark = pytest.mark.gpu


def edge(p):
    return orc.ints_to_limbs([0, 1, p - 1, p - 2, (1 << 256) % p, 2, (p - 1) // 2, 0xFFFFFFFF, 1 << 224, (1 << 253) + 5])


@pytest.mark.parametrize("field,p", [("fr", po.FR), ("fq", po.FQ)])
def test_field_ops_bit_exact(ctx, field, p):
    n = 1 << 16
    a, b = orc.uniform_fr(21, n), orc.uniform_fr(22, n)
    e = edge(p)
    a[: len(e)] = e
    b[: len(e)] = e[::-1]
    b[len(e): 2 * len(e)] = e
    for op in ("mul", "add", "sub"):
        got = getattr(ctx, f"{field}_{op}")(a, b)
        want = getattr(orc, f"{field}_{op}")(a, b)
        assert (got == want).all(), op


def test_mont_conversions(ctx):
    a = orc.uniform_fr(23, 5000)
    c = ctx.fr_from_mont(a)
    assert (c == orc.fr_from_mont(a)).all()
    assert (ctx.fr_to_mont(c) == a).all()


def test_empty_vector(ctx):
    z = np.zeros((0, 4), dtype=np.uint64)
    assert ctx.fr_mul(z, z).shape == (0, 4)
