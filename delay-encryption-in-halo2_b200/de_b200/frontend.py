"""Circuit front-end: the reference's bench circuits as witness generators, behind the C ABI
(`de_circuit_synthesize`, include/de_b200.h; C++ in frontend/).  Host only - no GPU, no Context.

    asg = frontend.delay_enc(n, e, x, message, k=16)       DelayEncryptCircuit   /root/reference/src/lib.rs:103-318
    asg = frontend.mod_pow(n, e, x, k=17)                  RSACircuit            /root/reference/benches/mod_pow.rs:36-140
    asg = frontend.pose_enc(key, message, k=11)            PoseidonEncCircuit    /root/reference/src/encryption/chip.rs:114-198

`asg` carries the fixed columns, advice columns (Montgomery limbs, (n, 4) uint64) and copy constraints of the shape
plonk.main_gate_shape() describes; keygen.keygen_from_assignment() turns it into keys, Prover.create_proof takes asg.advice.
Integers n, e, x are Python ints; field elements are ints (canonical) on the way in."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib, plonk
from ._lib import DeError

MOD_POW, POSE_ENC, DELAY_ENC, RSA_PKCS1, BIGINT_SQUARE, BIGINT_OPS, POSEIDON_HASH = 0, 1, 2, 3, 4, 5, 6
BITS_LEN, EXP_LIMB_BITS = 2048, 5       # src/lib.rs:122-124
MESSAGE_CAPACITY = 2                    # src/encryption/poseidon_enc.rs:10


class _Fr(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


class _Desc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("k", C.c_uint32), ("bits_len", C.c_uint32), ("exp_bits", C.c_uint32),
                ("n", C.c_void_p), ("n_len", C.c_size_t), ("e", C.c_void_p), ("e_len", C.c_size_t), ("x", C.c_void_p), ("x_len", C.c_size_t),
                ("message", C.c_void_p), ("message_len", C.c_uint32), ("key", _Fr * 2), ("witness_only", C.c_uint32), ("threads", C.c_uint32), ("reuse_buffer", C.c_uint32)]


class _Info(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_fixed", C.c_uint32), ("n_advice", C.c_uint32), ("n_outputs", C.c_uint32),
                ("used_rows", C.c_uint64), ("n_copies", C.c_uint64), ("synthesis_ms", C.c_double)]


def _mont(vals) -> np.ndarray:
    return np.array([plonk.mont_limbs(int(v)) for v in vals], dtype=np.uint64).reshape(-1, 4)


def from_mont(a: np.ndarray) -> List[int]:
    """(m, 4) Montgomery limbs -> canonical ints (host integers; for tests and small outputs)"""
    rinv = pow(1 << 256, -1, plonk.FR)
    raw = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") * rinv % plonk.FR for i in range(0, len(raw), 32)]


@dataclass
class SynthesizedCircuit:
    """what one Circuit::synthesize pass leaves behind"""
    shape: plonk.ConstraintSystemShape
    k: int
    fixed: np.ndarray          # (n_fixed, n, 4) Montgomery
    advice: np.ndarray         # (5, n, 4) Montgomery
    copies: np.ndarray         # (n_copies, 4) uint32: (left column, left row, right column, right row), permutation positions
    outputs: List[int]         # canonical ints
    used_rows: int
    synthesis_ms: float
    instances: list = field(default_factory=lambda: [[]])
    _handle: object = None

    def sigma(self, omega_mont: np.ndarray) -> np.ndarray:
        """permutation::keygen::Assembly::build_pk: (6, n, 4) sigma columns in lagrange form"""
        L = _lib.load()
        n_cols = len(self.shape.perm_columns)
        out = np.empty((n_cols, 1 << self.k, 4), dtype=np.uint64)
        om = np.ascontiguousarray(omega_mont, dtype=np.uint64).reshape(4)
        dl = _mont([plonk.FR_DELTA]).reshape(4)
        rc = L.de_assignment_sigma(self._handle.h, om.ctypes.data_as(C.c_void_p), dl.ctypes.data_as(C.c_void_p), n_cols, out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise DeError(rc, "de_assignment_sigma failed")
        return out


class _Handle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            _lib.load().de_assignment_free(self.h)
        except Exception:
            pass


def synthesize(kind: int, k: int, n: int = 0, e: int = 0, x: int = 0, message=(), key=(0, 0), bits_len: int = BITS_LEN,
               exp_bits: int = EXP_LIMB_BITS, witness_only: bool = False, threads: int = 1) -> SynthesizedCircuit:
    """witness_only: the pass create_proof makes (advice columns only: `fixed` is empty, `copies` too); threads: host threads
    a witness-only pass may use for the RSA region (same rows for every value)"""
    L = _lib.load()
    d = _Desc()
    d.kind, d.k, d.bits_len, d.exp_bits = kind, k, bits_len, exp_bits
    d.witness_only = 1 if witness_only else 0
    d.threads = threads
    nb = max(1, bits_len // 8)
    bufs = [np.frombuffer(int(v).to_bytes(max(nb, (int(v).bit_length() + 7) // 8), "little"), dtype=np.uint8).copy() for v in (n, e, x)]
    d.n, d.e, d.x = (b.ctypes.data for b in bufs)
    d.n_len, d.e_len, d.x_len = (len(b) for b in bufs)
    msg = _mont(message)
    d.message, d.message_len = (msg.ctypes.data if len(message) else None), len(message)
    kk = _mont(key)
    for i in range(2):
        for j in range(4):
            d.key[i].l[j] = int(kk[i, j])
    h = C.c_void_p()
    rc = L.de_circuit_synthesize(C.byref(d), C.byref(h))
    if rc != 0:
        raise DeError(rc, L.de_frontend_last_error().decode())
    handle = _Handle(h)
    info = _Info()
    L.de_assignment_info(h, C.byref(info))
    nrows = 1 << k
    fixed = np.empty((0 if witness_only else info.n_fixed, nrows, 4), dtype=np.uint64)
    advice = np.empty((info.n_advice, nrows, 4), dtype=np.uint64)
    for c in range(fixed.shape[0]):
        L.de_assignment_fixed(h, c, fixed[c].ctypes.data_as(C.c_void_p))
    for c in range(info.n_advice):
        L.de_assignment_advice(h, c, advice[c].ctypes.data_as(C.c_void_p))
    copies = np.empty((info.n_copies, 4), dtype=np.uint32)
    if info.n_copies:
        L.de_assignment_copies(h, copies.ctypes.data_as(C.c_void_p))
    outs = np.empty((max(info.n_outputs, 1), 4), dtype=np.uint64)
    L.de_assignment_outputs(h, outs.ctypes.data_as(C.c_void_p))
    shape = plonk.main_gate_shape(info.n_fixed == 15)
    return SynthesizedCircuit(shape, k, fixed, advice, copies, from_mont(outs[: info.n_outputs]), int(info.used_rows), float(info.synthesis_ms),
                              [[] for _ in range(shape.n_instance)], handle)


class WitnessPass:
    """The synthesis pass create_proof makes (advice columns only) for fixed circuit inputs, written into a caller's buffer
    of 5 * 2^k field elements (numpy array or pinned torch tensor): de_circuit_witness.  Thread-safe across instances; the call
    releases the GIL."""

    def __init__(self, kind: int, k: int, n: int = 0, e: int = 0, x: int = 0, message=(), key=(0, 0), bits_len: int = BITS_LEN,
                 exp_bits: int = EXP_LIMB_BITS, threads: int = 1):
        self.L = _lib.load()
        d = _Desc()
        d.kind, d.k, d.bits_len, d.exp_bits = kind, k, bits_len, exp_bits
        d.threads = threads
        nb = max(1, bits_len // 8)
        self._bufs = [np.frombuffer(int(v).to_bytes(nb, "little"), dtype=np.uint8).copy() for v in (n, e, x)]
        d.n, d.e, d.x = (b.ctypes.data for b in self._bufs)
        d.n_len = d.e_len = d.x_len = nb
        self._msg = _mont(message)
        d.message, d.message_len = (self._msg.ctypes.data if len(message) else None), len(message)
        kk = _mont(key)
        for i in range(2):
            for j in range(4):
                d.key[i].l[j] = int(kk[i, j])
        self.desc, self.k = d, k
        self.info = _Info()

    def run(self, out, reuse: bool = False) -> float:
        """fills `out` (5 * 2^k * 4 uint64); returns the synthesis time in ms.  reuse: `out` holds the result of an earlier run of
        this circuit (any inputs) and need not be zeroed again"""
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        self.desc.reuse_buffer = 1 if reuse else 0
        rc = self.L.de_circuit_witness(C.byref(self.desc), C.c_void_p(ptr), C.byref(self.info))
        if rc != 0:
            raise DeError(rc, self.L.de_frontend_last_error().decode())
        return float(self.info.synthesis_ms)


def delay_enc(n: int, e: int, x: int, message=(0,) * MESSAGE_CAPACITY, k: int = 16) -> SynthesizedCircuit:
    return synthesize(DELAY_ENC, k, n, e, x, message)


def mod_pow(n: int, e: int, x: int, k: int = 17) -> SynthesizedCircuit:
    return synthesize(MOD_POW, k, n, e, x)


def pose_enc(key, message=(0,) * MESSAGE_CAPACITY, k: int = 11) -> SynthesizedCircuit:
    return synthesize(POSE_ENC, k, message=message, key=key)


def rsa_pkcs1(n: int, e: int, signature: int, digest_limbs, k: int = 17) -> SynthesizedCircuit:
    return synthesize(RSA_PKCS1, k, n, e, signature, digest_limbs)


def bigint_square(a: int, expected: int, bits_len: int = BITS_LEN, k: int = 14) -> SynthesizedCircuit:
    """the big-integer chip's square test circuit (/root/reference/src/big_integer/chip.rs:2918-3030): outputs[0] is the
    is_equal_muled bit of a * a against `expected`, outputs[1:] the 2 * num_limbs - 1 uncarried ("Muled") limbs of a * a"""
    return synthesize(BIGINT_SQUARE, k, a, 0, expected, bits_len=bits_len)


def poseidon_hash(inputs, k: int = 12) -> SynthesizedCircuit:
    """PoseidonHashCircuit of the reference's hasher test (/root/reference/src/hash/chip.rs:113-236): outputs = the five state
    words after HasherChip::hash (words 1..4 are constrained equal to the native sponge's update + squeeze(1))"""
    return synthesize(POSEIDON_HASH, k, message=list(inputs))


BIGINT_OPS_RECORDS = ("add", "sub", "sub_overflow", "mul_mod", "pow_mod", "pow_mod_fixed_exp", "is_equal_fresh", "is_less_than",
                      "is_less_than_or_equal", "in_field")


def bigint_ops(a: int, b: int, n: int, bits_len: int = 256, exp_bits: int = EXP_LIMB_BITS, k: int = 14):
    """the big-integer chip's operator tests in one circuit (/root/reference/src/big_integer/chip.rs:1479-2806); returns
    (SynthesizedCircuit, {operator: integer value of its result}) - limbs recomposed with 64-bit weights"""
    syn = synthesize(BIGINT_OPS, k, n, b, a, bits_len=bits_len, exp_bits=exp_bits)
    out, at = {}, 0
    for name in BIGINT_OPS_RECORDS:
        cnt = syn.outputs[at]
        limbs = syn.outputs[at + 1: at + 1 + cnt]
        out[name] = sum(v << (64 * i) for i, v in enumerate(limbs))
        at += 1 + cnt
    assert at == len(syn.outputs)
    return syn, out


def poseidon_permute(state, t: int = 5, r_f: int = 8, r_p: int = 57) -> List[int]:
    L = _lib.load()
    s = _mont(state)
    rc = L.de_poseidon_permute(t, r_f, r_p, s.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise DeError(rc, L.de_frontend_last_error().decode())
    return from_mont(s)


def poseidon_encrypt(key, message) -> List[int]:
    L = _lib.load()
    kk, m = _mont(key), _mont(message)
    out = np.zeros((3, 4), dtype=np.uint64)
    rc = L.de_poseidon_cipher(0, kk.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), len(message), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise DeError(rc, L.de_frontend_last_error().decode())
    return from_mont(out)


def poseidon_decrypt(key, cipher):
    """-> message (list of ints) or None when the tag does not authenticate"""
    L = _lib.load()
    kk, c = _mont(key), _mont(cipher)
    out = np.zeros((MESSAGE_CAPACITY, 4), dtype=np.uint64)
    rc = L.de_poseidon_cipher(1, kk.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), len(cipher), out.ctypes.data_as(C.c_void_p))
    if rc == _lib.DE_ERR_UNSUPPORTED:
        return None
    if rc != 0:
        raise DeError(rc, L.de_frontend_last_error().decode())
    return from_mont(out)


def sample_rsa_inputs(seed: int, bits_len: int = BITS_LEN, exp_bits: int = EXP_LIMB_BITS):
    """n of exactly bits_len bits, e < 2^exp_bits, x < n, drawn as the reference's benches draw them
    (/root/reference/benches/delay_enc.rs:57-66) but from a seeded generator"""
    import random
    rng = random.Random(seed)
    n = 0
    while n.bit_length() != bits_len:
        n = rng.getrandbits(bits_len)
    e = rng.getrandbits(exp_bits) % n
    x = rng.getrandbits(bits_len) % n
    return n, e, x
