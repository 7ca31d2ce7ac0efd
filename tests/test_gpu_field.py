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


@pytest.mark.parametrize("field,p", [("fr", po.FR), ("fq", po.FQ)])
def test_dedicated_squaring_bit_exact(ctx, field, p):
    """DE_OP_SQR (ff::Field::square): field.cuh's 36-product squaring rows == the oracle's a * a"""
    n = 1 << 16
    a = orc.uniform_fr(25, n)
    e = edge(p)
    a[: len(e)] = e
    more = orc.ints_to_limbs([p - 1 - (1 << k) for k in range(0, 250, 7)] + [(1 << k) - 1 for k in range(1, 254, 5)])
    a[len(e): len(e) + len(more)] = more
    got = getattr(ctx, f"{field}_square")(a)
    assert (got == getattr(orc, f"{field}_mul")(a, a)).all()


def test_mont_conversions(ctx):
    a = orc.uniform_fr(23, 5000)
    c = ctx.fr_from_mont(a)
    assert (c == orc.fr_from_mont(a)).all()
    assert (ctx.fr_to_mont(c) == a).all()


def test_empty_vector(ctx):
    z = np.zeros((0, 4), dtype=np.uint64)
    assert ctx.fr_mul(z, z).shape == (0, 4)


@pytest.mark.parametrize("field,p", [("fr", po.FR), ("fq", po.FQ)])
def test_field_inversion(ctx, field, p):
    """DE_OP_INV: ff::Field::invert on the device (binary extended Euclid), 0 -> 0"""
    import ctypes as C
    n = 5000
    a = orc.uniform_fr(24, n)
    e = edge(p)
    a[: len(e)] = e
    fn = ctx.L.de_fr_vec_op if field == "fr" else ctx.L.de_fq_vec_op
    out = np.empty_like(a)
    ctx.check(fn(ctx.h, 5, a.ctypes.data_as(C.c_void_p), None, out.ctypes.data_as(C.c_void_p), n))
    vals = getattr(orc, field + "_ints_from_mont")(a)
    want = getattr(orc, field + "_mont_from_ints")([pow(v, -1, p) if v % p else 0 for v in vals])
    assert (out == want).all()
    back = getattr(ctx, field + "_mul")(out, a)
    one = getattr(orc, field + "_mont_from_ints")([1])[0]
    nz = np.array([v % p != 0 for v in vals])
    assert (back[nz] == one).all()
