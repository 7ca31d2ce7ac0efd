"""csrc/transcript.hpp (host Blake2b transcript + scalar Fr helpers of the C++ prover) against hashlib and the Python
restatement of Blake2bWrite / Challenge255 in oracle/pyprover.py."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import pyoracle as po
import pyprover as pp

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "delay-encryption-in-halo2_b200", "csrc")


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(HERE, "hostshim", "libtranscript_host.so")
    src = os.path.join(HERE, "hostshim", "transcript_host.cpp")
    hdr = os.path.join(CSRC, "transcript.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC, "-o", so, src])
    return C.CDLL(so)


def limbs(v):
    return np.array(po.limbs64(v), dtype=np.uint64)


def test_blake2b_matches_hashlib(shim):
    rng = np.random.default_rng(1)
    for n in (0, 1, 63, 64, 127, 128, 129, 255, 256, 257, 1000):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        for split in {0, n // 3, n}:
            out = (C.c_uint8 * 64)()
            shim.h_blake2b_personal(b"Halo2-Transcript", data, C.c_size_t(n), C.c_size_t(split), out)
            assert bytes(out) == hashlib.blake2b(data, digest_size=64, person=b"Halo2-Transcript").digest(), (n, split)


def test_host_fr_ops(shim):
    rng = po.Xoshiro(11)
    p = C.c_void_p
    for _ in range(200):
        a, b = rng.uniform_fr(), rng.uniform_fr()
        am, bm = limbs(po.to_mont(a, po.FR)), limbs(po.to_mont(b, po.FR))
        o = np.zeros(4, dtype=np.uint64)
        shim.h_fr_mul4(am.ctypes.data_as(p), bm.ctypes.data_as(p), o.ctypes.data_as(p))
        assert po.from_limbs64(o) == po.to_mont(a * b % po.FR, po.FR)
        shim.h_fr_add4(am.ctypes.data_as(p), bm.ctypes.data_as(p), o.ctypes.data_as(p))
        assert po.from_limbs64(o) == po.to_mont((a + b) % po.FR, po.FR)
        e = rng.next_u64()
        shim.h_fr_pow4(am.ctypes.data_as(p), C.c_uint64(e), o.ctypes.data_as(p))
        assert po.from_limbs64(o) == po.to_mont(pow(a, e, po.FR), po.FR)
        wide = bytes(rng.next_u64() & 0xFF for _ in range(64))
        shim.h_fr_from_wide(wide, o.ctypes.data_as(p))
        assert po.from_limbs64(o) == po.to_mont(po.from_uniform_bytes(wide), po.FR)
    wide = b"\xff" * 64
    o = np.zeros(4, dtype=np.uint64)
    shim.h_fr_from_wide(wide, o.ctypes.data_as(p))
    assert po.from_limbs64(o) == po.to_mont(po.from_uniform_bytes(wide), po.FR)


def test_transcript_script_matches_python(shim):
    rng = po.Xoshiro(12)
    t = pp.Transcript()
    script = bytearray()
    want_ch = []
    for i in range(60):
        op = i % 5
        if op in (0, 3):
            pt = po.g1_mul(po.G1_GEN, rng.uniform_fr())
            script += b"p" + pp.fq_to_repr(pt[0]) + pp.fq_to_repr(pt[1])
            t.write_point(pt)
        elif op == 1:
            s = rng.uniform_fr()
            script += b"s" + pp.fr_to_repr(s)
            t.write_scalar(s)
        elif op == 2:
            s = rng.uniform_fr()
            script += b"c" + pp.fr_to_repr(s)
            t.common_scalar(s)
        else:
            script += b"q"
            want_ch.append(t.squeeze_challenge())
    proof = (C.c_uint8 * 8192)()
    ch = (C.c_uint8 * (32 * 64))()
    nc = C.c_size_t()
    shim.h_transcript_script.restype = C.c_size_t
    m = shim.h_transcript_script(bytes(script), C.c_size_t(len(script)), proof, ch, C.byref(nc))
    assert bytes(proof[:m]) == t.finalize()
    got = [int.from_bytes(bytes(ch[32 * i:32 * i + 32]), "little") for i in range(nc.value)]
    assert got == want_ch


def test_host_batch_normalize_to_canonical(shim):
    import orc
    rng = po.Xoshiro(13)
    pts = [po.g1_mul(po.G1_GEN, rng.uniform_fr()) for _ in range(6)] + [None]
    jac = []
    for p in pts:
        if p is None:
            jac += [5, 7, 0]
            continue
        z = rng.uniform_fr() % po.FQ or 1
        jac += [p[0] * z * z % po.FQ, p[1] * z * z * z % po.FQ, z]
    jm = orc.fq_mont_from_ints(jac).reshape(-1, 12)
    out = (C.c_uint8 * (64 * len(pts)))()
    shim.h_g1_jacobian_to_canonical(jm.ctypes.data_as(C.c_void_p), C.c_size_t(len(pts)), out)
    for i, p in enumerate(pts):
        want = bytes(64) if p is None else pp.fq_to_repr(p[0]) + pp.fq_to_repr(p[1])
        assert bytes(out[64 * i:64 * i + 64]) == want
