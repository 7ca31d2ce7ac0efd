"""The device Montgomery arithmetic (csrc/field.cuh) compiled for the HOST with every PTX carry-chain instruction emulated,
checked against the oracle: proves the limb schedule (even/odd IMAD chains, final subtraction) without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "delay-encryption-in-halo2_b200", "csrc")


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(HERE, "hostshim", "libfield_host.so")
    src = os.path.join(HERE, "hostshim", "field_host.cpp")
    hdr = os.path.join(CSRC, "field.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", so, src])
    return C.CDLL(so)


def call(L, fn, a, b):
    o = np.empty_like(a)
    getattr(L, fn)(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), C.c_size_t(a.shape[0]))
    return o


def edge(p):
    return orc.ints_to_limbs([0, 1, p - 1, p - 2, (1 << 256) % p, 2, (p - 1) // 2, 0xFFFFFFFF, 1 << 224, (1 << 253) + 5])


@pytest.mark.parametrize("field,p", [("fr", po.FR), ("fq", po.FQ)])
def test_device_field_schedule_on_host(shim, field, p):
    n = 20000
    a, b = orc.uniform_fr(11, n), orc.uniform_fr(12, n)  # values < r < q: valid in both fields
    e = edge(p)
    a[: len(e)] = e
    b[: len(e)] = e[::-1]
    b[len(e): 2 * len(e)] = e
    ref = {"mul": getattr(orc, field + "_mul"), "add": getattr(orc, field + "_add"), "sub": getattr(orc, field + "_sub")}
    for op, f in ref.items():
        assert (call(shim, f"h_{field}_{op}", a, b) == f(a, b)).all(), op


@pytest.mark.parametrize("field,p", [("fr", po.FR), ("fq", po.FQ)])
def test_dedicated_squaring_on_host(shim, field, p):
    """field.cuh sqr(): the 36-product CIOS rows (a_i times the doubled upper part) against the oracle's a * a, edge values
    (0, 1, p - 1, p - 2, all-ones limbs, the largest inputs of the 9-limb window) included"""
    n = 50000
    a = orc.uniform_fr(21, n)
    e = edge(p)
    a[: len(e)] = e
    more = orc.ints_to_limbs([p - 1 - (1 << k) for k in range(0, 250, 7)] + [(1 << k) - 1 for k in range(1, 254, 5)] + [((1 << 253) | (0x7FFFFFFF << s)) % p for s in range(0, 220, 32)])
    a[len(e): len(e) + len(more)] = more
    mul = getattr(orc, field + "_mul")
    assert (call(shim, f"h_{field}_sqr", a, a) == mul(a, a)).all()


def test_from_mont_on_host(shim):
    a = orc.uniform_fr(13, 1000)
    assert (call(shim, "h_fr_from_mont", a, a) == orc.fr_from_mont(a)).all()


@pytest.mark.parametrize("field,p", [("fr", po.FR), ("fq", po.FQ)])
def test_binary_gcd_inversion_on_host(shim, field, p):
    """field.cuh inv(): binary extended Euclid on 8 x 32-bit limbs, Montgomery in / out, against modular inverses in Python"""
    n = 3000
    a = orc.uniform_fr(14, n)
    e = edge(p)
    a[: len(e)] = e
    got = call(shim, f"h_{field}_inv", a, a)
    from_m = getattr(orc, field + "_ints_from_mont")
    to_m = getattr(orc, field + "_mont_from_ints")
    vals = from_m(a)
    want = to_m([pow(v, -1, p) if v % p else 0 for v in vals])
    assert (got == want).all()

