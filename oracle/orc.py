"""ctypes loader for oracle/liboracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY — see oracle/oracle.c.

All field elements cross this boundary as numpy uint64 arrays of shape (..., 4): little-endian limbs in
Montgomery form, the same bytes the C ABI of the product library takes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        P = C.c_void_p
        for name in ("orc_fr_mul", "orc_fr_add", "orc_fr_sub", "orc_fq_mul", "orc_fq_add", "orc_fq_sub"):
            getattr(L, name).argtypes = [P, P, P, C.c_size_t]
        for name in ("orc_fr_from_mont", "orc_fr_to_mont", "orc_fq_from_mont", "orc_fq_to_mont", "orc_fr_inv",
                     "orc_g1_to_affine"):
            getattr(L, name).argtypes = [P, P, C.c_size_t]
        L.orc_best_multiexp.argtypes = [P, P, C.c_size_t, P, C.c_int]
        L.orc_msm_naive.argtypes = [P, P, C.c_size_t, P]
        L.orc_best_fft.argtypes = [P, P, C.c_uint32, C.c_int]
        L.orc_domain_new.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_domain_new.restype = P
        L.orc_domain_free.argtypes = [P]
        L.orc_domain_extended_k.argtypes = [P]
        L.orc_domain_extended_k.restype = C.c_uint32
        L.orc_domain_constants.argtypes = [P, P]
        L.orc_domain_t_evaluations.argtypes = [P, P]
        for name in ("orc_lagrange_to_coeff", "orc_coeff_to_lagrange", "orc_divide_by_vanishing"):
            getattr(L, name).argtypes = [P, P]
        L.orc_coeff_to_extended.argtypes = [P, P, P]
        L.orc_extended_to_coeff.argtypes = [P, P]
        L.orc_extended_to_coeff.restype = C.c_size_t
        L.orc_fill_uniform_fr.argtypes = [C.c_uint64, C.c_size_t, P]
        L.orc_fill_witness_fr.argtypes = [C.c_uint64, C.c_size_t, C.c_size_t, P]
        L.orc_gen_bases.argtypes = [C.c_size_t, C.c_size_t, P, C.c_int]
        L.orc_g1_mul.argtypes = [P, P, P]
        L.orc_g1_add.argtypes = [P, P, P]
        L.orc_g1_mul_many.argtypes = [P, P, C.c_size_t, P, C.c_int]
        L.orc_fr_batch_invert.argtypes = [P, C.c_size_t, C.c_int]
        L.orc_fr_running_product.argtypes = [P, C.c_size_t, P, P]
        L.orc_eval_polynomial.argtypes = [P, C.c_size_t, P, P, C.c_int]
        L.orc_kate_division.argtypes = [P, C.c_size_t, P, P]
        L.orc_fr_scale_add.argtypes = [P, P, P, C.c_size_t, C.c_int]
        L.orc_fr_axpy.argtypes = [P, P, P, C.c_size_t, C.c_int]
        L.orc_fr_geometric.argtypes = [P, P, C.c_size_t, P]
        L.orc_g1_on_curve.argtypes = [P]
        L.orc_g1_on_curve.restype = C.c_int
        L.orc_pk_new.argtypes = [P, P]
        L.orc_pk_new.restype = P
        L.orc_pk_free.argtypes = [P]
        L.orc_evaluate_h.argtypes = [P, P, P, P, P, P, P]
        L.orc_evaluate_h.restype = C.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def ncpu() -> int:
    return len(os.sched_getaffinity(0))


# ---- int <-> limb helpers (python big ints, canonical values) -------------------------------------
def ints_to_limbs(vals) -> np.ndarray:
    raw = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4).copy()


def limbs_to_ints(a: np.ndarray):
    raw = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def _un(fn, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    fn(_p(a), _p(out), a.size // 4)
    return out


def _bin(fn, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    fn(_p(a), _p(b), _p(out), a.size // 4)
    return out


def fr_mul(a, b): return _bin(lib().orc_fr_mul, a, b)
def fr_add(a, b): return _bin(lib().orc_fr_add, a, b)
def fr_sub(a, b): return _bin(lib().orc_fr_sub, a, b)
def fq_mul(a, b): return _bin(lib().orc_fq_mul, a, b)
def fq_add(a, b): return _bin(lib().orc_fq_add, a, b)
def fq_sub(a, b): return _bin(lib().orc_fq_sub, a, b)
def fr_to_mont(a): return _un(lib().orc_fr_to_mont, a)
def fr_from_mont(a): return _un(lib().orc_fr_from_mont, a)
def fq_to_mont(a): return _un(lib().orc_fq_to_mont, a)
def fq_from_mont(a): return _un(lib().orc_fq_from_mont, a)
def fr_inv(a): return _un(lib().orc_fr_inv, a)


def fr_mont_from_ints(vals): return fr_to_mont(ints_to_limbs(vals))
def fr_ints_from_mont(a): return limbs_to_ints(fr_from_mont(a))
def fq_mont_from_ints(vals): return fq_to_mont(ints_to_limbs(vals))
def fq_ints_from_mont(a): return limbs_to_ints(fq_from_mont(a))


def uniform_fr(seed: int, n: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_fill_uniform_fr(seed, n, _p(out))
    return out


def witness_fr(seed: int, n: int, used: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_fill_witness_fr(seed, n, used, _p(out))
    return out


def gen_bases(n: int, start: int = 0, threads: int | None = None) -> np.ndarray:
    """P_i = (start + i + 1) * G as affine Montgomery (n, 8) u64."""
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_gen_bases(start, n, _p(out), threads or ncpu())
    return out


def best_multiexp(scalars: np.ndarray, bases: np.ndarray, threads: int | None = None) -> np.ndarray:
    n = scalars.size // 4
    assert bases.size // 8 == n
    out = np.empty(12, dtype=np.uint64)
    lib().orc_best_multiexp(_p(np.ascontiguousarray(scalars)), _p(np.ascontiguousarray(bases)), n, _p(out),
                            threads or ncpu())
    return out


def msm_naive(scalars: np.ndarray, bases: np.ndarray) -> np.ndarray:
    n = scalars.size // 4
    out = np.empty(12, dtype=np.uint64)
    lib().orc_msm_naive(_p(np.ascontiguousarray(scalars)), _p(np.ascontiguousarray(bases)), n, _p(out))
    return out


def g1_to_affine(j: np.ndarray) -> np.ndarray:
    j = np.ascontiguousarray(j, dtype=np.uint64).reshape(-1, 12)
    out = np.empty((j.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_to_affine(_p(j), _p(out), j.shape[0])
    return out


def g1_mul_many(base_aff: np.ndarray, scalars: np.ndarray, threads: int | None = None) -> np.ndarray:
    """[scalars[i]] base for every i: affine Montgomery (n, 8)."""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = scalars.size // 4
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_g1_mul_many(_p(np.ascontiguousarray(base_aff, dtype=np.uint64)), _p(scalars), n, _p(out), threads or ncpu())
    return out


def g1_add(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_add(_p(np.ascontiguousarray(p)), _p(np.ascontiguousarray(q)), _p(out))
    return out


def g1_on_curve(a: np.ndarray) -> bool:
    return bool(lib().orc_g1_on_curve(_p(np.ascontiguousarray(a))))


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int | None = None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.uint64).copy()
    assert out.size == 4 << log_n
    lib().orc_best_fft(_p(out), _p(np.ascontiguousarray(omega)), log_n, threads or ncpu())
    return out


class Domain:
    """EvaluationDomain::new(j, k) restated (oracle.c: orc_domain_new)."""

    def __init__(self, j: int, k: int, threads: int | None = None):
        self.h = lib().orc_domain_new(j, k, threads or ncpu())
        self.j, self.k = j, k
        self.n = 1 << k
        self.extended_k = lib().orc_domain_extended_k(self.h)
        self.ext_n = 1 << self.extended_k
        c = np.empty((4, 4), dtype=np.uint64)
        lib().orc_domain_constants(self.h, _p(c))
        self.omega, self.omega_inv, self.ext_omega, self.ext_omega_inv = c[0], c[1], c[2], c[3]

    def __del__(self):
        try:
            lib().orc_domain_free(self.h)
        except Exception:
            pass

    def t_evaluations(self):
        out = np.empty((1 << (self.extended_k - self.k), 4), dtype=np.uint64)
        lib().orc_domain_t_evaluations(self.h, _p(out))
        return out

    def lagrange_to_coeff(self, a):
        out = np.ascontiguousarray(a).copy()
        lib().orc_lagrange_to_coeff(self.h, _p(out))
        return out

    def coeff_to_lagrange(self, a):
        out = np.ascontiguousarray(a).copy()
        lib().orc_coeff_to_lagrange(self.h, _p(out))
        return out

    def coeff_to_extended(self, a):
        a = np.ascontiguousarray(a)
        out = np.empty((self.ext_n, 4), dtype=np.uint64)
        lib().orc_coeff_to_extended(self.h, _p(a), _p(out))
        return out

    def extended_to_coeff(self, a):
        out = np.ascontiguousarray(a).copy()
        m = lib().orc_extended_to_coeff(self.h, _p(out))
        return out.reshape(-1, 4)[:m].copy()

    def divide_by_vanishing(self, a):
        out = np.ascontiguousarray(a).copy()
        lib().orc_divide_by_vanishing(self.h, _p(out))
        return out


class Pk:
    """keygen_pk's resident cosets (oracle.c: orc_pk_new).  desc_struct is the ctypes image of de_pk_desc."""

    def __init__(self, domain: "Domain", desc_struct, keepalive=None):
        self.domain, self.desc, self.keep = domain, desc_struct, keepalive
        self.h = lib().orc_pk_new(domain.h, C.byref(desc_struct))

    def __del__(self):
        try:
            lib().orc_pk_free(self.h)
        except Exception:
            pass

    def evaluate_h(self, advice, instance, challenges_struct, permz, lookup):
        """lookup: [all z | all a' | all s']; polynomial lists are (n, 4) uint64 Montgomery arrays in coefficient form."""
        def arr(polys):
            polys = [np.ascontiguousarray(p, dtype=np.uint64) for p in polys]
            a = (C.c_void_p * max(len(polys), 1))(*[p.ctypes.data for p in polys])
            return a, polys
        a, k1 = arr(advice)
        i, k2 = arr(instance)
        z, k3 = arr(permz)
        l, k4 = arr(lookup)
        out = np.empty((self.domain.ext_n, 4), dtype=np.uint64)
        rc = lib().orc_evaluate_h(self.h, a, i, C.byref(challenges_struct), z, l, _p(out))
        assert rc == 0
        return out


def evaluate_h(domain: "Domain", desc_struct, advice, instance, challenges_struct, permz, lookup):
    return Pk(domain, desc_struct).evaluate_h(advice, instance, challenges_struct, permz, lookup)


# ---- vector helpers of the restated prover (oracle.c: "vector helpers of the prover") ---------------------------------------
def fr_batch_invert(a: np.ndarray, threads: int | None = None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().orc_fr_batch_invert(_p(out), out.size // 4, threads or ncpu())
    return out


def fr_running_product(f: np.ndarray, start: np.ndarray) -> np.ndarray:
    f = np.ascontiguousarray(f, dtype=np.uint64)
    z = np.empty_like(f)
    lib().orc_fr_running_product(_p(f), f.size // 4, _p(np.ascontiguousarray(start, dtype=np.uint64)), _p(z))
    return z


def eval_polynomial(poly: np.ndarray, x: np.ndarray, threads: int | None = None) -> np.ndarray:
    poly = np.ascontiguousarray(poly, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_eval_polynomial(_p(poly), poly.size // 4, _p(np.ascontiguousarray(x, dtype=np.uint64)), _p(out), threads or ncpu())
    return out


def kate_division(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    n = a.size // 4
    q = np.empty((n - 1, 4), dtype=np.uint64)
    lib().orc_kate_division(_p(a), n, _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(q))
    return q


def fr_scale_add(acc: np.ndarray, p: np.ndarray, s: np.ndarray, threads: int | None = None) -> None:
    """acc = acc * s + p, in place"""
    lib().orc_fr_scale_add(_p(acc), _p(np.ascontiguousarray(p, dtype=np.uint64)), _p(np.ascontiguousarray(s, dtype=np.uint64)), acc.size // 4,
                           threads or ncpu())


def fr_axpy(acc: np.ndarray, p: np.ndarray, s: np.ndarray, threads: int | None = None) -> None:
    """acc += s * p, in place"""
    lib().orc_fr_axpy(_p(acc), _p(np.ascontiguousarray(p, dtype=np.uint64)), _p(np.ascontiguousarray(s, dtype=np.uint64)), acc.size // 4,
                      threads or ncpu())


def fr_geometric(start: np.ndarray, ratio: np.ndarray, n: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_fr_geometric(_p(np.ascontiguousarray(start, dtype=np.uint64)), _p(np.ascontiguousarray(ratio, dtype=np.uint64)), n, _p(out))
    return out
