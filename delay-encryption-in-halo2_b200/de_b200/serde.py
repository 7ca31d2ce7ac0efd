"""`SerdeFormat::RawBytes` readers / writers for the key files the reference's benches cache next to the SRS
(/root/reference/benches/delay_enc.rs:84-115: `vk.write(.., RawBytes)` -> `VerifyingKey::read::<_, Circuit>(.., RawBytes)`,
`pk.write` -> `ProvingKey::read`; the params file is `de_b200.ParamsKZG.read / write`).

Layout as halo2_proofs v2023_04_20 writes it (plonk.rs `VerifyingKey::write`, `ProvingKey::write`, poly.rs `Polynomial::write`,
helpers.rs; restated from the published source - the crate is not vendored in the reference):

    vk :=  k: u32 BE | n_fixed: u32 BE | fixed_commitments[n_fixed] | permutation commitments[P] | selectors
           commitment = G1Affine raw: x, y as 4 x u64 LE Montgomery limbs (64 B); selectors: n_sel vectors of n bools, 8 per byte,
           least significant bit first
    pk :=  vk | l0 | l_last | l_active_row | fixed_values | fixed_polys | fixed_cosets | permutations | polys | cosets
           polynomial = len: u32 BE | len field elements raw (32 B Montgomery limbs); a slice = count: u32 BE | polynomials
           l0 / l_last / l_active_row and the cosets live on the extended domain (2^extended_k values), the rest on 2^k

The counts a reader cannot infer from the bytes (P permutation columns, selector count) come from the constraint system, which
halo2 rebuilds by running `Circuit::configure` inside `read`; here they come from the `ConstraintSystemShape`.  RawBytes holds
exactly the in-memory representation the C ABI takes, so arrays are handed to the device without conversion.  The reader
auto-detects little-endian headers as well (one reading of the format would be wrong by endianness only; sizes pin the rest:
8 + 64 (F + P) + n_sel n / 8 bytes for a vk - 968 B and 17 736 B for the reference's two shapes, /root/reference/benches/README.md:56-99).
"""
from __future__ import annotations

import io
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class VerifyingKeyRaw:
    k: int
    fixed_commitments: np.ndarray          # (F, 8) uint64: affine x | y, Montgomery
    permutation_commitments: np.ndarray    # (P, 8)
    selectors: List[np.ndarray] = field(default_factory=list)   # bool arrays of n entries


@dataclass
class ProvingKeyRaw:
    vk: VerifyingKeyRaw
    l0: np.ndarray                         # (ext_n, 4) extended-domain evaluations
    l_last: np.ndarray
    l_active_row: np.ndarray
    fixed_values: np.ndarray               # (F, n, 4) lagrange
    fixed_polys: np.ndarray                # (F, n, 4) coefficients
    fixed_cosets: np.ndarray               # (F, ext_n, 4)
    permutations: np.ndarray               # (P, n, 4) sigma columns, lagrange
    polys: np.ndarray                      # (P, n, 4) sigma polynomials
    cosets: np.ndarray                     # (P, ext_n, 4)


def _u32(v: int, order: str) -> bytes:
    return int(v).to_bytes(4, order)


class _Reader:
    def __init__(self, data, order="big"):
        self.b = memoryview(data)
        self.pos = 0
        self.order = order

    def u32(self) -> int:
        v = int.from_bytes(self.b[self.pos:self.pos + 4], self.order)
        self.pos += 4
        return v

    def take(self, nbytes: int) -> memoryview:
        if self.pos + nbytes > len(self.b):
            raise ValueError("RawBytes key file is truncated")
        v = self.b[self.pos:self.pos + nbytes]
        self.pos += nbytes
        return v

    def polynomial(self) -> np.ndarray:
        m = self.u32()
        return np.frombuffer(self.take(32 * m), dtype=np.uint64).reshape(m, 4)

    def slice(self) -> np.ndarray:
        count = self.u32()
        polys = [self.polynomial() for _ in range(count)]
        if not polys:
            return np.zeros((0, 0, 4), dtype=np.uint64)
        if any(p.shape != polys[0].shape for p in polys):
            raise ValueError("polynomials of one slice differ in length")
        return np.stack(polys)


def _detect_order(data) -> str:
    k_be, k_le = int.from_bytes(data[:4], "big"), int.from_bytes(data[:4], "little")
    if 1 <= k_be <= 28:
        return "big"
    if 1 <= k_le <= 28:
        return "little"
    raise ValueError("not a RawBytes key file: k out of range in either byte order")


def n_selectors_of(shape) -> int:
    """selector vectors halo2 keeps in the vk for this constraint system: the RangeChip's two complex selectors for the RSA
    shape (vk size 17 736 B at k = 16), none for the Poseidon-only shape (968 B) - SURVEY.md Appendix C"""
    return 2 if shape.n_fixed == 15 else 0


def _read_vk(r: _Reader, n_perm: int, n_sel: int) -> VerifyingKeyRaw:
    k = r.u32()
    nf = r.u32()
    if not (1 <= k <= 28) or nf > 4096:
        raise ValueError("not a RawBytes verifying key")
    fixed = np.frombuffer(r.take(64 * nf), dtype=np.uint64).reshape(nf, 8)
    perm = np.frombuffer(r.take(64 * n_perm), dtype=np.uint64).reshape(n_perm, 8)
    n = 1 << k
    sels = []
    for _ in range(n_sel):
        packed = np.frombuffer(r.take((n + 7) // 8), dtype=np.uint8)
        sels.append(np.unpackbits(packed, bitorder="little")[:n].astype(bool))
    return VerifyingKeyRaw(k, fixed, perm, sels)


def read_vk(path_or_bytes, shape) -> VerifyingKeyRaw:
    """VerifyingKey::read(reader, SerdeFormat::RawBytes)"""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    r = _Reader(data, _detect_order(data))
    vk = _read_vk(r, len(shape.perm_columns), n_selectors_of(shape))
    if vk.fixed_commitments.shape[0] != shape.n_fixed:
        raise ValueError(f"vk holds {vk.fixed_commitments.shape[0]} fixed commitments, the constraint system has {shape.n_fixed}")
    if r.pos != len(data):
        raise ValueError(f"{len(data) - r.pos} trailing bytes after the verifying key")
    return vk


def read_pk(path_or_bytes, shape, n_selectors: int | None = None) -> ProvingKeyRaw:
    """ProvingKey::read(reader, SerdeFormat::RawBytes).  Accepts a path (memory-mapped: a k = 16 key is 276 MiB) or bytes.
    n_selectors: selector vectors in the embedded vk (default: what this repository's shapes have)"""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = path_or_bytes
    else:
        data = np.memmap(path_or_bytes, dtype=np.uint8, mode="r")
    r = _Reader(data, _detect_order(bytes(data[:4])))
    vk = _read_vk(r, len(shape.perm_columns), n_selectors_of(shape) if n_selectors is None else n_selectors)
    l0, l_last, l_active = r.polynomial(), r.polynomial(), r.polynomial()
    fixed_values, fixed_polys, fixed_cosets = r.slice(), r.slice(), r.slice()
    permutations, polys, cosets = r.slice(), r.slice(), r.slice()
    if r.pos != len(data):
        raise ValueError(f"{len(data) - r.pos} trailing bytes after the proving key")
    n = 1 << vk.k
    if fixed_values.shape[:2] != (shape.n_fixed, n) or polys.shape[:2] != (len(shape.perm_columns), n):
        raise ValueError("proving key does not match the constraint system (column counts / 2^k)")
    if l0.shape[0] % n or fixed_cosets.shape[1] != l0.shape[0] or cosets.shape[1] != l0.shape[0]:
        raise ValueError("proving key: inconsistent extended-domain lengths")
    return ProvingKeyRaw(vk, l0, l_last, l_active, fixed_values, fixed_polys, fixed_cosets, permutations, polys, cosets)


def _write_vk(f, vk: VerifyingKeyRaw, order: str):
    f.write(_u32(vk.k, order))
    f.write(_u32(vk.fixed_commitments.shape[0], order))
    f.write(np.ascontiguousarray(vk.fixed_commitments, dtype=np.uint64).tobytes())
    f.write(np.ascontiguousarray(vk.permutation_commitments, dtype=np.uint64).tobytes())
    for s in vk.selectors:
        f.write(np.packbits(np.asarray(s, dtype=bool), bitorder="little").tobytes())


def _write_poly(f, p, order):
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    f.write(_u32(p.shape[0], order))
    f.write(p.tobytes())


def _write_slice(f, ps, order):
    f.write(_u32(len(ps), order))
    for p in ps:
        _write_poly(f, p, order)


def write_vk(path, vk: VerifyingKeyRaw, order: str = "big"):
    """VerifyingKey::write(writer, SerdeFormat::RawBytes)"""
    with (open(path, "wb") if not isinstance(path, io.IOBase) else path) as f:
        _write_vk(f, vk, order)


def write_pk(path, pk: ProvingKeyRaw, order: str = "big"):
    """ProvingKey::write(writer, SerdeFormat::RawBytes)"""
    with (open(path, "wb") if not isinstance(path, io.IOBase) else path) as f:
        _write_vk(f, pk.vk, order)
        for p in (pk.l0, pk.l_last, pk.l_active_row):
            _write_poly(f, p, order)
        for ps in (pk.fixed_values, pk.fixed_polys, pk.fixed_cosets, pk.permutations, pk.polys, pk.cosets):
            _write_slice(f, ps, order)


def commitments_to_affine_mont(ctx, xy_canonical: np.ndarray) -> np.ndarray:
    """(m, 64) canonical little-endian x || y (the transcript form keygen produces) -> (m, 8) Montgomery limbs (the raw form)"""
    c = np.ascontiguousarray(xy_canonical, dtype=np.uint8).reshape(-1, 64).view(np.uint64).reshape(-1, 4)
    return ctx.fq_mul(c, np.tile(_FQ_R2, (c.shape[0], 1))).reshape(-1, 8)


_FQ_R2 = np.array([0xF32CFC5B538AFA89, 0xB5E71911D44501FB, 0x47AB1EFF0A417FF6, 0x06D89F71CAB8351F], dtype=np.uint64)


def proving_key_raw_from_keys(keys, fixed_values: np.ndarray, sigma_values: np.ndarray, selectors=()) -> ProvingKeyRaw:
    """What keygen_pk holds, assembled from this library's keys (keygen.Keys): the polynomials are the ones uploaded to the
    device; l0 / l_last / l_active_row and every coset are produced by the device's own lagrange_to_coeff / coeff_to_extended,
    exactly the operations keygen_pk runs (plonk/keygen.rs: l0 = e_0, l_blind = sum of the last blinding_factors rows' basis
    polynomials, l_last = e_(n - blinding_factors - 1), l_active_row = 1 - l_last - l_blind on the extended domain)."""
    dom, shape, ctx = keys.domain, keys.pk.shape, keys.domain.ctx
    n, k = dom.n, dom.k
    one = np.array(_mont_one(), dtype=np.uint64)
    bf = shape.blinding_factors

    def basis(rows):
        v = np.zeros((n, 4), dtype=np.uint64)
        for r in rows:
            v[r] = one
        return dom.coeff_to_extended(dom.lagrange_to_coeff(v))

    l0 = basis([0])
    l_blind = basis(range(n - bf, n))
    l_last = basis([n - bf - 1])
    ones = np.tile(one, (dom.extended_n, 1))
    l_active = ctx.fr_sub(ctx.fr_sub(ones, l_last), l_blind)
    h = keys.host
    fixed_polys, sigma_polys = np.stack(h["fixed_polys"]) if h["fixed_polys"] else np.zeros((0, n, 4), np.uint64), np.stack(h["sigma_polys"])
    fixed_cosets = np.stack([dom.coeff_to_extended(p) for p in fixed_polys]) if len(fixed_polys) else np.zeros((0, dom.extended_n, 4), np.uint64)
    cosets = np.stack([dom.coeff_to_extended(p) for p in sigma_polys])
    vk = VerifyingKeyRaw(k, commitments_to_affine_mont(ctx, keys.fixed_commitments), commitments_to_affine_mont(ctx, keys.permutation_commitments),
                         [np.asarray(s, dtype=bool) for s in selectors])
    return ProvingKeyRaw(vk, l0, l_last, l_active, np.ascontiguousarray(fixed_values), fixed_polys, fixed_cosets, np.ascontiguousarray(sigma_values),
                         sigma_polys, cosets)


def _mont_one():
    r = (1 << 256) % 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    return [(r >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def keys_from_proving_key_raw(ctx, pkraw: ProvingKeyRaw, shape, g, g_lagrange, transcript_repr: int, queries=None):
    """A prover from a key FILE: the polynomials go to the device as they are (de_pk_upload recomputes the cosets there);
    returns keygen.Keys"""
    from . import EvaluationDomain, ParamsKZG
    from .keygen import Keys
    from .plonk import Prover, ProvingKey, collect_queries
    k = pkraw.vk.k
    params = ParamsKZG(k, g, g_lagrange, ctx)
    domain = EvaluationDomain(shape.degree(), k, ctx)
    if domain.extended_n != pkraw.l0.shape[0]:
        raise ValueError("the key file's extended domain does not match cs.degree()")
    fixed_polys = [np.ascontiguousarray(p) for p in pkraw.fixed_polys]
    sigma_polys = [np.ascontiguousarray(p) for p in pkraw.polys]
    pk = ProvingKey(domain, shape, fixed_polys, sigma_polys)
    aq, fq, _ = queries if queries is not None else collect_queries(shape)
    prover = Prover(params, pk, aq, fq, transcript_repr)
    host = dict(k=k, g=g, g_lagrange=g_lagrange, shape=shape, fixed_polys=fixed_polys, sigma_polys=sigma_polys, advice_queries=aq, fixed_queries=fq,
                transcript_repr=transcript_repr)
    to_xy = lambda m: ctx.fq_from_mont(np.ascontiguousarray(m).reshape(-1, 4)).reshape(-1, 8).view(np.uint8).reshape(-1, 64)
    return Keys(params, domain, pk, prover, to_xy(pkraw.vk.fixed_commitments), to_xy(pkraw.vk.permutation_commitments), host)
