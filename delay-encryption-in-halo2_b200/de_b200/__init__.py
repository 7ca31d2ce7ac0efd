"""de_b200 — host-side mirror (Python) of the halo2_proofs arithmetic / poly surface that the delay-encryption circuits'
prover reaches, backed by libde_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/de_b200.h).  Names and argument meaning follow halo2_proofs @ v2023_04_20:

    best_multiexp(coeffs, bases)                    arithmetic::best_multiexp
    best_fft(a, omega, log_n)                       arithmetic::best_fft (in place in Rust; returns the result here)
    EvaluationDomain(j, k)                          poly::EvaluationDomain::new
        .coeff_to_extended / .extended_to_coeff / .lagrange_to_coeff / .coeff_to_lagrange / .divide_by_vanishing_poly
    ParamsKZG(k, g, g_lagrange)                     poly::kzg::commitment::ParamsKZG (bases staged in HBM once)
        .commit(poly) / .commit_lagrange(poly)

Field elements are numpy uint64 arrays of shape (n, 4) (Montgomery limbs, halo2curves' memory layout); points are
(n, 8) affine / (12,) Jacobian.  `*_dev` methods take torch CUDA tensors (uint64/int64 storage) and stay on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeError

__all__ = ["Context", "EvaluationDomain", "ParamsKZG", "best_multiexp", "best_fft", "DeError", "default_context", "read_params_raw",
           "write_params_raw"]


def _np(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None:
        assert a.size % cols == 0
    return a


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor


class Context:
    """de_ctx: one device, one stream.  Not thread-safe (one per host thread / GPU)."""

    def __init__(self, device: int = 0):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.de_ctx_create(device, C.byref(h))
        if rc != 0:
            raise DeError(rc, self.L.de_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc != 0:
            raise DeError(rc, self.L.de_last_error(self.h).decode())

    def set_stream(self, cuda_stream: int | None):
        self.check(self.L.de_ctx_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self.check(self.L.de_ctx_sync(self.h))

    def set_mode(self, throughput: bool):
        """de_ctx_set_mode: latency (default, one proof at a time) or throughput (several contexts share the GPU)"""
        self.check(self.L.de_ctx_set_mode(self.h, 1 if throughput else 0))

    @property
    def launches(self) -> int:
        return int(self.L.de_launch_count(self.h))

    def timing_enable(self, on: bool = True):
        self.check(self.L.de_timing_enable(self.h, 1 if on else 0))

    def timing_reset(self):
        self.check(self.L.de_timing_reset(self.h))

    def timing_get(self, kernel: str):
        """(total_ms, total_units, launches) of a named kernel since the last reset"""
        ms, units, n = C.c_double(), C.c_double(), C.c_uint64()
        self.check(self.L.de_timing_get(self.h, kernel.encode(), C.byref(ms), C.byref(units), C.byref(n)))
        return ms.value, units.value, int(n.value)

    def int_peak(self) -> float:
        """de_int_peak: Fr Montgomery multiplications per second (in G/s) this device sustains, measured now"""
        v = C.c_double()
        self.check(self.L.de_int_peak(self.h, C.byref(v)))
        return v.value

    def int_peak_sqr(self) -> float:
        """de_int_peak_sqr: Fr squarings per second (in G/s), the dedicated squaring of field.cuh"""
        v = C.c_double()
        self.check(self.L.de_int_peak_sqr(self.h, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "h", None):
            self.L.de_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- a1 ----
    def _vec(self, fn, op, a, b):
        a = _np(a)
        out = np.empty_like(a)
        bb = _np(b) if b is not None else None
        self.check(fn(self.h, op, _ptr(a), _ptr(bb), _ptr(out), a.size // 4))
        return out

    def fr_mul(self, a, b): return self._vec(self.L.de_fr_vec_op, _lib.OP_MUL, a, b)
    def fr_add(self, a, b): return self._vec(self.L.de_fr_vec_op, _lib.OP_ADD, a, b)
    def fr_sub(self, a, b): return self._vec(self.L.de_fr_vec_op, _lib.OP_SUB, a, b)
    def fr_from_mont(self, a): return self._vec(self.L.de_fr_vec_op, _lib.OP_FROM_MONT, a, None)
    def fr_to_mont(self, a): return self._vec(self.L.de_fr_vec_op, _lib.OP_TO_MONT, a, None)
    def fq_from_mont(self, a): return self._vec(self.L.de_fq_vec_op, _lib.OP_FROM_MONT, a, None)
    def fq_to_mont(self, a): return self._vec(self.L.de_fq_vec_op, _lib.OP_TO_MONT, a, None)
    def fq_mul(self, a, b): return self._vec(self.L.de_fq_vec_op, _lib.OP_MUL, a, b)
    def fq_add(self, a, b): return self._vec(self.L.de_fq_vec_op, _lib.OP_ADD, a, b)
    def fq_sub(self, a, b): return self._vec(self.L.de_fq_vec_op, _lib.OP_SUB, a, b)
    def fr_square(self, a): return self._vec(self.L.de_fr_vec_op, _lib.OP_SQR, a, None)
    def fq_square(self, a): return self._vec(self.L.de_fq_vec_op, _lib.OP_SQR, a, None)

    # ---- a3 / a4 ----
    def best_multiexp(self, coeffs, bases):
        coeffs, bases = _np(coeffs), _np(bases)
        n = coeffs.size // 4
        if bases.size // 8 != n:
            raise ValueError("best_multiexp: coeffs.len() != bases.len()")  # assert_eq! in the reference
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.L.de_msm(self.h, _ptr(coeffs), _ptr(bases), n, _ptr(out)))
        return out

    def best_multiexp_dev(self, d_coeffs, d_bases, n: int):
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.L.de_msm_dev(self.h, _ptr(d_coeffs), _ptr(d_bases), n, _ptr(out)))
        return out

    def best_fft(self, a, omega, log_n: int):
        a = _np(a).copy()
        if a.size // 4 != (1 << log_n):
            raise ValueError("best_fft: a.len() != 1 << log_n")  # assert_eq! in the reference
        omega = _np(omega)
        self.check(self.L.de_ntt(self.h, _ptr(a), _ptr(omega), log_n))
        return a

    def best_fft_dev(self, d_a, omega, log_n: int, batch: int = 1, stride: int | None = None):
        omega = _np(omega)
        self.check(self.L.de_ntt_dev(self.h, _ptr(d_a), _ptr(omega), log_n, batch, stride or (1 << log_n)))

    def eval_polynomial(self, poly, point):
        """arithmetic::eval_polynomial: poly (n, 4) coefficients, point (4,) -> (4,) (all Montgomery)"""
        poly, point = _np(poly), _np(point)
        out = np.zeros(4, dtype=np.uint64)
        self.check(self.L.de_eval_polynomial(self.h, _ptr(poly), poly.size // 4, _ptr(point), _ptr(out)))
        return out

    def kate_division(self, a, b):
        """arithmetic::kate_division: quotient of a(X) by (X - b), len(a) - 1 coefficients"""
        a, b = _np(a), _np(b)
        n = a.size // 4
        out = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
        self.check(self.L.de_kate_division(self.h, _ptr(a), n, _ptr(b), _ptr(out)))
        return out

    def g1_mul_base_dev(self, base, d_scalars, n: int, d_out):
        """d_out[i] = [d_scalars[i]] base: base (8,) affine Montgomery host array, scalars / out CUDA tensors"""
        base = _np(base)
        self.check(self.L.de_g1_mul_base_dev(self.h, _ptr(base), _ptr(d_scalars), n, _ptr(d_out)))

    def batch_normalize(self, points):
        """group::Curve::batch_normalize: (count, 12) Jacobian -> (count, 8) affine"""
        points = _np(points).reshape(-1, 12)
        out = np.zeros((points.shape[0], 8), dtype=np.uint64)
        self.check(self.L.de_g1_batch_normalize(self.h, _ptr(points), points.shape[0], _ptr(out)))
        return out

    def g1_sum(self, points):
        points = _np(points)
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.L.de_g1_sum(self.h, _ptr(points), points.size // 12, _ptr(out)))
        return out


_default = {}


def default_context(device: int = 0) -> Context:
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def best_multiexp(coeffs, bases, ctx: Context | None = None):
    return (ctx or default_context()).best_multiexp(coeffs, bases)


def best_fft(a, omega, log_n: int, ctx: Context | None = None):
    return (ctx or default_context()).best_fft(a, omega, log_n)


class EvaluationDomain:
    """poly::EvaluationDomain::new(j, k) with j = cs.degree()."""

    def __init__(self, j: int, k: int, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        L = self.ctx.L
        h = C.c_void_p()
        self.ctx.check(L.de_domain_create(self.ctx.h, j, k, C.byref(h)))
        self.h = h
        ek = C.c_uint32()
        consts = np.zeros((4, 4), dtype=np.uint64)
        self.ctx.check(L.de_domain_info(h, C.byref(ek), _ptr(consts)))
        self.j, self.k, self.extended_k = j, k, int(ek.value)
        self.n, self.extended_n = 1 << k, 1 << self.extended_k
        self.omega, self.omega_inv, self.extended_omega, self.extended_omega_inv = consts
        self.quotient_poly_degree = j - 1

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.de_domain_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk_len(self, a, n, what):
        a = _np(a)
        if a.size // 4 != n:
            raise ValueError(f"{what}: wrong length {a.size // 4}, expected {n}")
        return a

    def coeff_to_extended(self, a):
        a = self._chk_len(a, self.n, "coeff_to_extended")
        out = np.empty((self.extended_n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.de_coeff_to_extended(self.h, _ptr(a), _ptr(out)))
        return out

    def extended_to_coeff(self, a):
        a = self._chk_len(a, self.extended_n, "extended_to_coeff").copy()
        m = C.c_size_t()
        self.ctx.check(self.ctx.L.de_extended_to_coeff(self.h, _ptr(a), C.byref(m)))
        return a.reshape(-1, 4)[: m.value].copy()

    def lagrange_to_coeff(self, a):
        a = self._chk_len(a, self.n, "lagrange_to_coeff").copy()
        self.ctx.check(self.ctx.L.de_lagrange_to_coeff(self.h, _ptr(a)))
        return a

    def coeff_to_lagrange(self, a):
        a = self._chk_len(a, self.n, "coeff_to_lagrange").copy()
        self.ctx.check(self.ctx.L.de_coeff_to_lagrange(self.h, _ptr(a)))
        return a

    def divide_by_vanishing_poly(self, a):
        a = self._chk_len(a, self.extended_n, "divide_by_vanishing_poly").copy()
        self.ctx.check(self.ctx.L.de_divide_by_vanishing(self.h, _ptr(a)))
        return a

    # device-resident, batched
    def coeff_to_extended_dev(self, d_coeff, d_ext, batch=1, in_stride=None, out_stride=None):
        self.ctx.check(self.ctx.L.de_coeff_to_extended_dev(self.h, _ptr(d_coeff), in_stride or self.n, _ptr(d_ext),
                                                           out_stride or self.extended_n, batch))

    def extended_to_coeff_dev(self, d_ext, batch=1, stride=None):
        m = C.c_size_t()
        self.ctx.check(self.ctx.L.de_extended_to_coeff_dev(self.h, _ptr(d_ext), stride or self.extended_n, batch, C.byref(m)))
        return m.value

    def lagrange_to_coeff_dev(self, d_a, batch=1, stride=None):
        self.ctx.check(self.ctx.L.de_lagrange_to_coeff_dev(self.h, _ptr(d_a), stride or self.n, batch))

    def coeff_to_lagrange_dev(self, d_a, batch=1, stride=None):
        self.ctx.check(self.ctx.L.de_coeff_to_lagrange_dev(self.h, _ptr(d_a), stride or self.n, batch))

    def divide_by_vanishing_poly_dev(self, d_ext, batch=1, stride=None):
        self.ctx.check(self.ctx.L.de_divide_by_vanishing_dev(self.h, _ptr(d_ext), stride or self.extended_n, batch))


def read_params_raw(path):
    """ParamsKZG::read(reader, SerdeFormat::RawBytes), the format the reference's benches cache their SRS in
    (/root/reference/benches/delay_enc.rs:43-54): k as u32 LE, g[2^k] and g_lagrange[2^k] as raw
    64-byte affine points (Montgomery limbs, the in-memory layout the C ABI takes), then g2 and s_g2 (128 bytes each).
    Returns dict(k, g, g_lagrange, g2, s_g2) with the point arrays memory-mapped (no copy until upload)."""
    import os
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        k = int.from_bytes(f.read(4), "little")
    n = 1 << k
    if k > 28 or size != 4 + 2 * n * 64 + 256:
        raise ValueError(f"{path}: not a RawBytes ParamsKZG file for k = {k} (size {size})")
    g = np.memmap(path, dtype=np.uint64, mode="r", offset=4, shape=(n, 8))
    gl = np.memmap(path, dtype=np.uint64, mode="r", offset=4 + n * 64, shape=(n, 8))
    tail = np.fromfile(path, dtype=np.uint8, offset=4 + 2 * n * 64)
    return dict(k=k, g=g, g_lagrange=gl, g2=tail[:128].tobytes(), s_g2=tail[128:].tobytes())


def write_params_raw(path, k, g, g_lagrange, g2: bytes, s_g2: bytes):
    """ParamsKZG::write(writer, SerdeFormat::RawBytes)"""
    g, gl = _np(g), _np(g_lagrange)
    if g.size != 8 << k or gl.size != 8 << k or len(g2) != 128 or len(s_g2) != 128:
        raise ValueError("write_params_raw: wrong sizes")
    with open(path, "wb") as f:
        f.write(int(k).to_bytes(4, "little"))
        f.write(g.tobytes())
        f.write(gl.tobytes())
        f.write(g2)
        f.write(s_g2)


_FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
# halo2curves::bn256::G2Affine::generator(): (x.c0, x.c1), (y.c0, y.c1)
_G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634),
           (8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531))


def _g2_mul(point, k: int):
    """[k] point on the twist y^2 = x^3 + 3 / (9 + u) over Fq2 = Fq[u] / (u^2 + 1): ONE scalar multiplication per setup
    (s_g2 = [s] G2), done in Python integers; the product has no other G2 arithmetic."""
    q = _FQ

    def mul2(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)

    def inv2(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, q)
        return (a[0] * d % q, -a[1] * d % q)

    def sub2(a, b):
        return ((a[0] - b[0]) % q, (a[1] - b[1]) % q)

    def add_pts(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if P[1] != Q[1] or P[1] == (0, 0):
                return None
            lam = mul2(mul2((3, 0), mul2(P[0], P[0])), inv2(mul2((2, 0), P[1])))
        else:
            lam = mul2(sub2(Q[1], P[1]), inv2(sub2(Q[0], P[0])))
        x3 = sub2(sub2(mul2(lam, lam), P[0]), Q[0])
        return (x3, sub2(mul2(lam, sub2(P[0], x3)), P[1]))

    acc = None
    for bit in bin(k % _FR)[2:] if k % _FR else "":
        acc = add_pts(acc, acc)
        if bit == "1":
            acc = add_pts(acc, point)
    return acc


def _g2_raw_bytes(pt) -> bytes:
    """SerdeFormat::RawBytes of a G2Affine: x.c0, x.c1, y.c0, y.c1 as Montgomery limbs (identity: zeros)"""
    if pt is None:
        return bytes(128)
    out = b""
    for coord in pt:
        for c in coord:
            out += (c * (1 << 256) % _FQ).to_bytes(32, "little")
    return out


class ParamsKZG:
    """poly::kzg::commitment::ParamsKZG: g / g_lagrange staged (with window tables) in HBM once."""

    @classmethod
    def setup(cls, k: int, s: int, ctx: Context | None = None, keep_host: bool = True):
        """ParamsKZG::setup with a caller-supplied secret s (the reference draws it from OsRng, benches/delay_enc.rs:43):
        g[i] = [s^i] G, g_lagrange[i] = [l_i(s)] G with l_i(s) = omega^i (s^n - 1) / (n (s - omega^i)), g2 = G2, s_g2 = [s] G2.
        The 2 * 2^k fixed-base multiplications run on the device (de_g1_mul_base_dev); the scalars (powers of s, one batched
        inversion) are one-time host integers."""
        import torch
        ctx = ctx or default_context()
        n = 1 << k
        s %= _FR
        root = pow(7, (_FR - 1) >> 28, _FR)
        omega = pow(root, 1 << (28 - k), _FR)
        pows, w = [1] * n, [1] * n
        for i in range(1, n):
            pows[i] = pows[i - 1] * s % _FR
            w[i] = w[i - 1] * omega % _FR
        den = [(s - w[i]) * n % _FR for i in range(n)]
        pre, acc = [], 1
        for d in den:
            pre.append(acc)
            acc = acc * d % _FR
        inv = pow(acc, -1, _FR)  # s is not an n-th root of unity (probability n / r)
        sn1 = (pow(s, n, _FR) - 1) % _FR
        lag = [0] * n
        for i in range(n - 1, -1, -1):
            lag[i] = w[i] * sn1 % _FR * (inv * pre[i] % _FR) % _FR
            inv = inv * den[i] % _FR
        gen = np.array([(1 << 256) % _FQ >> (64 * j) & 0xFFFFFFFFFFFFFFFF for j in range(4)] +
                       [2 * (1 << 256) % _FQ >> (64 * j) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)
        outs = []
        for scal in (pows, lag):
            raw = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in scal), dtype=np.uint64).reshape(n, 4)
            d_s = torch.from_numpy(ctx.fr_to_mont(raw).view(np.int64)).cuda(ctx.device)
            d_o = torch.empty((n, 8), dtype=torch.int64, device=d_s.device)
            ctx.g1_mul_base_dev(gen, d_s, n, d_o)
            ctx.sync()
            outs.append(d_o.cpu().numpy().view(np.uint64))
        p = cls(k, outs[0], outs[1], ctx)
        p.g2, p.s_g2 = _g2_raw_bytes(_G2_GEN), _g2_raw_bytes(_g2_mul(_G2_GEN, s))
        if keep_host:
            p.g_host, p.g_lagrange_host = outs
        return p

    def write(self, path):
        """ParamsKZG::write(.., SerdeFormat::RawBytes); needs the host copies kept by setup() / read()"""
        write_params_raw(path, self.k, self.g_host, self.g_lagrange_host, self.g2, self.s_g2)

    @classmethod
    def read(cls, path, ctx: Context | None = None):
        """ParamsKZG::read(.., SerdeFormat::RawBytes): the file's point arrays are uploaded straight from the mapping"""
        d = read_params_raw(path)
        p = cls(d["k"], np.ascontiguousarray(d["g"]), np.ascontiguousarray(d["g_lagrange"]), ctx)
        p.g2, p.s_g2 = d["g2"], d["s_g2"]
        p.g_host, p.g_lagrange_host = d["g"], d["g_lagrange"]
        return p

    def __init__(self, k: int, g=None, g_lagrange=None, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.k, self.n = k, 1 << k
        g = _np(g) if g is not None else None
        gl = _np(g_lagrange) if g_lagrange is not None else None
        for b in (g, gl):
            if b is not None and b.size // 8 != self.n:
                raise ValueError("ParamsKZG: basis length != 2^k")
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.de_params_upload(self.ctx.h, k, _ptr(g), _ptr(gl), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.de_params_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _commit(self, basis, poly):
        poly = _np(poly)
        out = np.zeros(12, dtype=np.uint64)
        self.ctx.check(self.ctx.L.de_commit(self.h, basis, _ptr(poly), poly.size // 4, _ptr(out)))
        return out

    def commit(self, poly): return self._commit(0, poly)
    def commit_lagrange(self, poly): return self._commit(1, poly)

    def commit_batch(self, basis: int, polys):
        polys = [_np(p) for p in polys]
        n = polys[0].size // 4
        arr = (C.c_void_p * len(polys))(*[p.ctypes.data for p in polys])
        out = np.zeros((len(polys), 12), dtype=np.uint64)
        self.ctx.check(self.ctx.L.de_commit_batch(self.h, basis, arr, n, len(polys), _ptr(out)))
        return out

    def commit_batch_dev(self, basis: int, d_scalars, n: int, count: int, stride: int | None = None):
        out = np.zeros((count, 12), dtype=np.uint64)
        self.ctx.check(self.ctx.L.de_commit_batch_dev(self.h, basis, _ptr(d_scalars), stride or n, n, count, _ptr(out)))
        return out

    def commit_batch_canonical_dev(self, basis: int, d_scalars, n: int, count: int, stride: int | None = None):
        """commitments as (count, 64) bytes: canonical little-endian x || y (the transcript's encoding)"""
        out = np.zeros((count, 64), dtype=np.uint8)
        self.ctx.check(self.ctx.L.de_commit_batch_canonical_dev(self.h, basis, _ptr(d_scalars), stride or n, n, count,
                                                                out.ctypes.data_as(C.c_void_p)))
        return out

    def commit_range(self, basis: int, poly, lo: int, hi: int):
        poly = _np(poly)
        out = np.zeros(12, dtype=np.uint64)
        self.ctx.check(self.ctx.L.de_commit_range(self.h, basis, _ptr(poly), lo, hi, _ptr(out)))
        return out
