"""CPU restatement of halo2_proofs::plonk::{keygen_vk, keygen_pk, create_proof, verify_proof} with the KZG / GWC
multi-open and the Blake2b transcript, as the reference's benches drive them
(/root/reference/benches/delay_enc.rs:86,103,123-131,153-160; mod_pow.rs:163,180,201,230; pose_enc.rs:89,106,127,156).

TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py): tests/ compare the CUDA prover's proof bytes with create_proof() below
on the same witness and the same stream of random field elements, and check that both proofs pass verify_proof() below.

PARITY UNPINNED against the Rust binary: halo2_proofs (tag v2023_04_20, /root/reference/Cargo.toml:17) is not vendored
and cannot be built here, and the reference holds no golden proof.  This file follows the
upstream algorithm as SURVEY.md Appendix B / E / F records it: order of commitments, challenges (theta, beta, gamma, y, x, v),
evaluations, RNG draws, lookup permutation rule, grand products, GWC grouping by point, byte encodings.  Two inputs the Rust
prover derives itself are PARAMETERS here because their derivation cannot be restated without the crate sources:
  * vk.transcript_repr (a Blake2b hash of the Debug string of the pinned verifying key), and
  * the order of cs.advice_queries / cs.fixed_queries / cs.instance_queries (set by the circuit's configure()).

Field elements are canonical Python ints; points are affine (x, y) tuples or None.  Heavy steps (MSM, NTT, evaluate_h)
go through the C oracle (orc), which tests/test_oracle.py and tests/test_evaluator_oracle.py pin against the
definition-level code of pyoracle.py.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

import orc
import pairing
import pyoracle as po
from pyoracle import FQ, FR, FR_DELTA

FIXED, ADVICE, INSTANCE = 2, 3, 4  # de_value_kind codes used for "any column" descriptors


# ---------------------------------------------------------------------------------------------------------------------
# encodings
# ---------------------------------------------------------------------------------------------------------------------
def fr_to_repr(v: int) -> bytes:
    return (v % FR).to_bytes(32, "little")


def fq_to_repr(v: int) -> bytes:
    return (v % FQ).to_bytes(32, "little")


def g1_to_bytes(p) -> bytes:
    """G1Affine::to_bytes: x little-endian, sign(y) = y_repr[0] & 1 in bit 7 of byte 31; identity = zeros."""
    if p is None:
        return bytes(32)
    b = bytearray(fq_to_repr(p[0]))
    b[31] |= (p[1] & 1) << 7
    return bytes(b)


def g1_from_bytes(b: bytes):
    if b == bytes(32):
        return None
    raw = bytearray(b)
    sign = raw[31] >> 7
    raw[31] &= 0x7F
    x = int.from_bytes(raw, "little")
    if x >= FQ:
        raise ValueError("invalid point encoding")
    rhs = (x * x * x + 3) % FQ
    y = pow(rhs, (FQ + 1) // 4, FQ)
    if y * y % FQ != rhs:
        raise ValueError("point not on curve")
    if (y & 1) != sign:
        y = FQ - y
    return (x, y)


# ---------------------------------------------------------------------------------------------------------------------
# transcript::{Blake2bWrite, Blake2bRead} with Challenge255
# ---------------------------------------------------------------------------------------------------------------------
class Transcript:
    PREFIX_CHALLENGE, PREFIX_POINT, PREFIX_SCALAR = b"\x00", b"\x01", b"\x02"

    def __init__(self, proof: Optional[bytes] = None):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.out = bytearray()
        self.inp = proof
        self.pos = 0

    def squeeze_challenge(self) -> int:
        self.state.update(self.PREFIX_CHALLENGE)
        digest = self.state.copy().digest()
        return int.from_bytes(digest, "little") % FR

    def common_point(self, p):
        if p is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.state.update(self.PREFIX_POINT)
        self.state.update(fq_to_repr(p[0]))
        self.state.update(fq_to_repr(p[1]))

    def common_scalar(self, s: int):
        self.state.update(self.PREFIX_SCALAR)
        self.state.update(fr_to_repr(s))

    def write_point(self, p):
        self.common_point(p)
        self.out += g1_to_bytes(p)

    def write_scalar(self, s: int):
        self.common_scalar(s)
        self.out += fr_to_repr(s)

    def read_point(self):
        b = self.inp[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        p = g1_from_bytes(bytes(b))
        self.common_point(p)
        return p

    def read_scalar(self) -> int:
        b = self.inp[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof too short")
        self.pos += 32
        v = int.from_bytes(b, "little")
        if v >= FR:
            raise ValueError("invalid scalar encoding")
        self.common_scalar(v)
        return v

    def finalize(self) -> bytes:
        return bytes(self.out)


# ---------------------------------------------------------------------------------------------------------------------
# glue to the C oracle: ints <-> Montgomery limb arrays
# ---------------------------------------------------------------------------------------------------------------------
def to_mont(vals: Sequence[int]) -> np.ndarray:
    return orc.fr_mont_from_ints([v % FR for v in vals])


def from_mont(a: np.ndarray) -> List[int]:
    return orc.fr_ints_from_mont(np.ascontiguousarray(a))


def points_to_mont(pts) -> np.ndarray:
    flat = []
    for p in pts:
        flat += [0, 0] if p is None else [p[0], p[1]]
    return orc.fq_mont_from_ints(flat).reshape(-1, 8)


def jac_mont_to_affine(j: np.ndarray):
    a = orc.g1_to_affine(j)
    x, y = orc.fq_ints_from_mont(a.reshape(-1, 4))
    return None if (x == 0 and y == 0) else (x, y)


def batch_invert(vals: List[int]) -> List[int]:
    """ff::BatchInvert: zero entries stay zero."""
    out = list(vals)
    acc, pre = 1, []
    for v in vals:
        pre.append(acc)
        if v % FR:
            acc = acc * v % FR
    inv = pow(acc, -1, FR)
    for i in range(len(vals) - 1, -1, -1):
        v = vals[i] % FR
        if v:
            out[i] = inv * pre[i] % FR
            inv = inv * v % FR
        else:
            out[i] = 0
    return out


# ---------------------------------------------------------------------------------------------------------------------
# ParamsKZG::setup with a caller-supplied secret (the reference draws s from OsRng, benches/delay_enc.rs:43)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class Params:
    k: int
    n: int
    g: list                 # affine int points [s^i] G
    g_lagrange: list        # [l_i(s)] G
    g2: tuple
    s_g2: tuple
    g_mont: np.ndarray = None
    g_lagrange_mont: np.ndarray = None

    def commit(self, coeffs: Sequence[int]):
        return jac_mont_to_affine(orc.best_multiexp(to_mont(coeffs), self.g_mont[:len(coeffs)]))

    def commit_lagrange(self, values: Sequence[int]):
        return jac_mont_to_affine(orc.best_multiexp(to_mont(values), self.g_lagrange_mont[:len(values)]))


def fixed_base_mul_many(scalars: Sequence[int]) -> np.ndarray:
    """[scalars[i]] G for the generator G = (1, 2): affine Montgomery (n, 8)."""
    return orc.g1_mul_many(points_to_mont([po.G1_GEN])[0], to_mont(scalars))


def setup(k: int, s: int) -> Params:
    n = 1 << k
    s %= FR
    pows = [1] * n
    for i in range(1, n):
        pows[i] = pows[i - 1] * s % FR
    # l_i(s) = omega^i (s^n - 1) / (n (s - omega^i))
    dom = po.EvaluationDomain(3, k)
    sn_minus_1 = (pow(s, n, FR) - 1) % FR
    w = [1] * n
    for i in range(1, n):
        w[i] = w[i - 1] * dom.omega % FR
    den = batch_invert([(s - w[i]) * n % FR for i in range(n)])
    lag = [w[i] * sn_minus_1 % FR * den[i] % FR for i in range(n)]
    g_mont = fixed_base_mul_many(pows)
    gl_mont = fixed_base_mul_many(lag)

    def to_pts(a):
        ints = orc.fq_ints_from_mont(a.reshape(-1, 4))
        return [None if (ints[2 * i] == 0 and ints[2 * i + 1] == 0) else (ints[2 * i], ints[2 * i + 1]) for i in range(len(a))]

    return Params(k, n, to_pts(g_mont), to_pts(gl_mont), pairing.G2_GEN, pairing.g2_mul(pairing.G2_GEN, s), g_mont, gl_mont)


# ---------------------------------------------------------------------------------------------------------------------
# constraint system view + keygen
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class Queries:
    """cs.advice_queries / fixed_queries / instance_queries: (column index, rotation) in the circuit's query order."""
    advice: List[Tuple[int, int]]
    fixed: List[Tuple[int, int]]
    instance: List[Tuple[int, int]]

    def index(self, kind: int, col: int, rot: int) -> int:
        lst = {ADVICE: self.advice, FIXED: self.fixed, INSTANCE: self.instance}[kind]
        return lst.index((col, rot))


class PermutationAssembly:
    """permutation::keygen::Assembly: cycles over the permutation columns, merged by copy()."""

    def __init__(self, n_cols: int, n: int):
        self.mapping = [[(c, r) for r in range(n)] for c in range(n_cols)]
        self.aux = [[(c, r) for r in range(n)] for c in range(n_cols)]
        self.sizes = [[1] * n for _ in range(n_cols)]

    def copy(self, lc: int, lr: int, rc: int, rr: int):
        if self.aux[lc][lr] == self.aux[rc][rr]:
            return
        left, right = self.aux[lc][lr], self.aux[rc][rr]
        if self.sizes[left[0]][left[1]] < self.sizes[right[0]][right[1]]:
            left, right = right, left
        self.sizes[left[0]][left[1]] += self.sizes[right[0]][right[1]]
        i = right
        while True:
            self.aux[i[0]][i[1]] = left
            i = self.mapping[i[0]][i[1]]
            if i == right:
                break
        self.mapping[lc][lr], self.mapping[rc][rr] = self.mapping[rc][rr], self.mapping[lc][lr]


@dataclass
class VerifyingKey:
    k: int
    shape: object
    queries: Queries
    fixed_commitments: list
    permutation_commitments: list
    transcript_repr: int


@dataclass
class ProvingKey:
    vk: VerifyingKey
    fixed_values: List[List[int]]
    fixed_polys: List[List[int]]
    sigma_values: List[List[int]]
    sigma_polys: List[List[int]]


def _cdomain(shape, k: int) -> "orc.Domain":
    return orc.Domain(shape.degree(), k)


def _l2c(dom, values: Sequence[int]) -> List[int]:
    return from_mont(dom.lagrange_to_coeff(to_mont(values)))


def keygen(params: Params, shape, queries: Queries, fixed_values: List[List[int]],
           copies: Sequence[Tuple[int, int, int, int]], transcript_repr: int) -> ProvingKey:
    """keygen_vk + keygen_pk.  fixed_values: one list of n ints per fixed column (selectors already compressed into fixed
    columns); copies: (perm column position, row, perm column position, row) equality constraints."""
    n, k = params.n, params.k
    dom = _cdomain(shape, k)
    asm = PermutationAssembly(len(shape.perm_columns), n)
    for lc, lr, rc, rr in copies:
        asm.copy(lc, lr, rc, rr)
    omega = po.EvaluationDomain(shape.degree(), k).omega
    omega_pows = [1] * n
    for i in range(1, n):
        omega_pows[i] = omega_pows[i - 1] * omega % FR
    sigma_values = []
    for c in range(len(shape.perm_columns)):
        col = []
        for r in range(n):
            mc, mr = asm.mapping[c][r]
            col.append(pow(FR_DELTA, mc, FR) * omega_pows[mr] % FR)
        sigma_values.append(col)
    fixed_polys = [_l2c(dom, f) for f in fixed_values]
    sigma_polys = [_l2c(dom, s) for s in sigma_values]
    vk = VerifyingKey(k, shape, queries, [params.commit_lagrange(f) for f in fixed_values],
                      [params.commit_lagrange(s) for s in sigma_values], transcript_repr % FR)
    return ProvingKey(vk, [list(f) for f in fixed_values], fixed_polys, sigma_values, sigma_polys)


# ---------------------------------------------------------------------------------------------------------------------
# expression evaluation
# ---------------------------------------------------------------------------------------------------------------------
def eval_expr_rows(e, n: int, fixed, advice, instance, challenges) -> List[int]:
    """plonk::evaluation::evaluate(expression, size, rot_scale = 1, ...) over the n lagrange rows."""
    t = e[0]
    if t == "const":
        return [e[1] % FR] * n
    if t in ("fixed", "advice", "instance"):
        col = {"fixed": fixed, "advice": advice, "instance": instance}[t][e[1]]
        rot = e[2]
        return [col[(i + rot) % n] for i in range(n)]
    if t == "challenge":
        return [challenges[e[1]]] * n
    if t == "neg":
        return [(-v) % FR for v in eval_expr_rows(e[1], n, fixed, advice, instance, challenges)]
    if t == "sum":
        a = eval_expr_rows(e[1], n, fixed, advice, instance, challenges)
        b = eval_expr_rows(e[2], n, fixed, advice, instance, challenges)
        return [(x + y) % FR for x, y in zip(a, b)]
    if t == "prod":
        a = eval_expr_rows(e[1], n, fixed, advice, instance, challenges)
        b = eval_expr_rows(e[2], n, fixed, advice, instance, challenges)
        return [x * y % FR for x, y in zip(a, b)]
    if t == "scaled":
        return [v * e[2] % FR for v in eval_expr_rows(e[1], n, fixed, advice, instance, challenges)]
    raise ValueError(t)


def eval_expr_point(e, q: Queries, fixed_evals, advice_evals, instance_evals, challenges) -> int:
    """Expression::evaluate with the query evaluations read from the proof (verifier side)."""
    t = e[0]
    if t == "const":
        return e[1] % FR
    if t == "fixed":
        return fixed_evals[q.index(FIXED, e[1], e[2])]
    if t == "advice":
        return advice_evals[q.index(ADVICE, e[1], e[2])]
    if t == "instance":
        return instance_evals[q.index(INSTANCE, e[1], e[2])]
    if t == "challenge":
        return challenges[e[1]]
    if t == "neg":
        return (-eval_expr_point(e[1], q, fixed_evals, advice_evals, instance_evals, challenges)) % FR
    if t == "sum":
        return (eval_expr_point(e[1], q, fixed_evals, advice_evals, instance_evals, challenges)
                + eval_expr_point(e[2], q, fixed_evals, advice_evals, instance_evals, challenges)) % FR
    if t == "prod":
        return (eval_expr_point(e[1], q, fixed_evals, advice_evals, instance_evals, challenges)
                * eval_expr_point(e[2], q, fixed_evals, advice_evals, instance_evals, challenges)) % FR
    if t == "scaled":
        return eval_expr_point(e[1], q, fixed_evals, advice_evals, instance_evals, challenges) * e[2] % FR
    raise ValueError(t)


# ---------------------------------------------------------------------------------------------------------------------
# lookup::prover::permute_expression_pair
# ---------------------------------------------------------------------------------------------------------------------
def permute_expression_pair(input_values: List[int], table_values: List[int], usable_rows: int):
    """Returns (permuted_input[:usable], permuted_table[:usable]) before the random tail is appended.  Raises on an input
    value that is missing from the table (Error::ConstraintSystemFailure)."""
    permuted_input = sorted(input_values[:usable_rows])  # Fr's Ord = canonical integer order
    leftover: Dict[int, int] = {}
    for v in table_values[:usable_rows]:
        leftover[v] = leftover.get(v, 0) + 1
    permuted_table = [0] * usable_rows
    repeated_rows = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            if leftover.get(v, 0) <= 0:
                raise ValueError("ConstraintSystemFailure: lookup input value not in table")
            leftover[v] -= 1
        else:
            repeated_rows.append(row)
    for v in sorted(leftover):  # BTreeMap iteration order
        for _ in range(leftover[v]):
            permuted_table[repeated_rows.pop()] = v
    assert not repeated_rows
    return permuted_input, permuted_table


# ---------------------------------------------------------------------------------------------------------------------
# create_proof
# ---------------------------------------------------------------------------------------------------------------------
def rotate_omega(omega: int, x: int, rot: int) -> int:
    return x * pow(omega, rot, FR) % FR


def kate_division(a: Sequence[int], b: int) -> List[int]:
    """arithmetic::kate_division: quotient of a(X) by (X - b), len(a) - 1 coefficients."""
    q = [0] * (len(a) - 1)
    tmp = 0
    for i in range(len(a) - 1, 0, -1):
        tmp = (a[i] + tmp * b) % FR
        q[i - 1] = tmp
    return q


@dataclass
class ProofTrace:
    """Intermediate values of one create_proof run, for kernel-level parity tests."""
    challenges: Dict[str, int] = field(default_factory=dict)
    commitments: List = field(default_factory=list)
    evals: List[int] = field(default_factory=list)
    columns: Dict[str, list] = field(default_factory=dict)


def create_proof(params: Params, pk: ProvingKey, advice_values: List[List[int]], instances: List[List[int]],
                 next_random: Callable[[], int], trace: Optional[ProofTrace] = None) -> bytes:
    """plonk::create_proof for ONE circuit instance, KZG + ProverGWC + Blake2bWrite/Challenge255, single phase, no
    user challenges (the shape of every circuit in the reference).  advice_values: the synthesized advice columns (n each;
    the last blinding_factors + 1 rows are overwritten with random values, as the prover does).  next_random() returns the
    next Fr::random(rng) draw; draws happen in the reference's order."""
    vk, shape, q = pk.vk, pk.vk.shape, pk.vk.queries
    n, k = params.n, params.k
    dom = _cdomain(shape, k)
    pdom = po.EvaluationDomain(shape.degree(), k)
    omega = pdom.omega
    bf = shape.blinding_factors
    usable = n - (bf + 1)
    tr = Transcript()
    T = trace if trace is not None else ProofTrace()

    # vk.hash_into
    tr.common_scalar(vk.transcript_repr)
    # instance columns: values are hashed (QUERY_INSTANCE = false for KZG), polys by iFFT
    assert len(instances) == shape.n_instance
    instance_values = []
    for vals in instances:
        assert len(vals) <= usable, "InstanceTooLarge"
        col = [0] * n
        for i, v in enumerate(vals):
            tr.common_scalar(v)
            col[i] = v % FR
        instance_values.append(col)
    instance_polys = [_l2c(dom, c) for c in instance_values]

    # advice: blinding rows, blinds, commitments
    advice = [list(c) for c in advice_values]
    assert len(advice) == shape.n_advice and all(len(c) == n for c in advice)
    for col in advice:
        for r in range(usable, n):
            col[r] = next_random()
    for _ in advice:
        next_random()  # Blind(Scalar::random): unused by KZG but drawn
    for col in advice:
        c = params.commit_lagrange(col)
        tr.write_point(c)
        T.commitments.append(c)
    challenges: List[int] = []
    theta = tr.squeeze_challenge()

    # lookups: compress, permute, commit
    lookups = []
    for inp, tab in shape.lookups:
        def compress(exprs):
            acc = [0] * n
            for e in exprs:
                vals = eval_expr_rows(e, n, pk.fixed_values, advice, instance_values, challenges)
                acc = [(a * theta + v) % FR for a, v in zip(acc, vals)]
            return acc
        ci, ct = compress(inp), compress(tab)
        pi, pt = permute_expression_pair(ci, ct, usable)
        pi += [next_random() for _ in range(bf + 1)]
        pt += [next_random() for _ in range(bf + 1)]
        next_random()  # permuted input blind
        cpi = params.commit_lagrange(pi)
        next_random()  # permuted table blind
        cpt = params.commit_lagrange(pt)
        tr.write_point(cpi)
        tr.write_point(cpt)
        T.commitments += [cpi, cpt]
        lookups.append(dict(ci=ci, ct=ct, pi=pi, pt=pt))
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()

    # permutation argument: grand products per chunk of columns
    cols_any = {ADVICE: advice, FIXED: pk.fixed_values, INSTANCE: instance_values}
    chunk = shape.chunk_len
    perm_z = []
    deltaomega = 1
    last_z = 1
    for s0 in range(0, len(shape.perm_columns), chunk):
        columns = shape.perm_columns[s0:s0 + chunk]
        modified = [1] * n
        for ci_, (kind, index) in enumerate(columns):
            vals, sig = cols_any[kind][index], pk.sigma_values[s0 + ci_]
            modified = [m * ((beta * sg + gamma + v) % FR) % FR for m, sg, v in zip(modified, sig, vals)]
        modified = batch_invert(modified)
        for (kind, index) in columns:
            vals = cols_any[kind][index]
            dw = deltaomega
            for i in range(n):
                modified[i] = modified[i] * ((dw * beta + gamma + vals[i]) % FR) % FR
                dw = dw * omega % FR
            deltaomega = deltaomega * FR_DELTA % FR
        z = [last_z]
        for row in range(1, n):
            z.append(z[row - 1] * modified[row - 1] % FR)
        for r in range(n - bf, n):
            z[r] = next_random()
        last_z = z[n - (bf + 1)]
        next_random()  # blind
        cz = params.commit_lagrange(z)
        tr.write_point(cz)
        T.commitments.append(cz)
        perm_z.append(z)

    # lookup grand products
    for L in lookups:
        prod = batch_invert([(beta + a) * (gamma + s) % FR for a, s in zip(L["pi"], L["pt"])])
        prod = [p * ((ci + beta) % FR) % FR * ((ct + gamma) % FR) % FR for p, ci, ct in zip(prod, L["ci"], L["ct"])]
        z, state = [], 1
        for cur in [1] + prod:
            state = state * cur % FR
            z.append(state)
        z = z[:n - bf] + [next_random() for _ in range(bf)]
        next_random()  # blind
        cz = params.commit_lagrange(z)
        tr.write_point(cz)
        T.commitments.append(cz)
        L["z"] = z

    # vanishing argument: random polynomial
    random_poly = [next_random() for _ in range(n)]
    next_random()  # blind
    c_random = params.commit(random_poly)
    tr.write_point(c_random)
    T.commitments.append(c_random)
    y = tr.squeeze_challenge()

    # to coefficient form
    advice_polys = [_l2c(dom, c) for c in advice]
    perm_polys = [_l2c(dom, z) for z in perm_z]
    for L in lookups:
        L["z_poly"], L["pi_poly"], L["pt_poly"] = _l2c(dom, L["z"]), _l2c(dom, L["pi"]), _l2c(dom, L["pt"])

    # h(X)
    from de_b200 import plonk  # the constraint-system descriptor is marshalled by the product's host layer
    desc, keep = plonk.marshal_pk_desc(shape, [to_mont(p) for p in pk.fixed_polys], [to_mont(p) for p in pk.sigma_polys])
    opk = orc.Pk(dom, desc, keep)
    chs, keep2 = plonk.marshal_challenges(y, beta, gamma, theta, challenges)
    lookup_block = [to_mont(L["z_poly"]) for L in lookups] + [to_mont(L["pi_poly"]) for L in lookups] + \
                   [to_mont(L["pt_poly"]) for L in lookups]
    h_ext = opk.evaluate_h([to_mont(p) for p in advice_polys], [to_mont(p) for p in instance_polys], chs,
                           [to_mont(p) for p in perm_polys], lookup_block)
    h_coeff = from_mont(dom.extended_to_coeff(dom.divide_by_vanishing(h_ext)))
    n_pieces = shape.degree() - 1
    h_pieces = [h_coeff[i * n:(i + 1) * n] for i in range(n_pieces)]
    for _ in h_pieces:
        next_random()  # h blinds
    for piece in h_pieces:
        c = params.commit(piece)
        tr.write_point(c)
        T.commitments.append(c)
    x = tr.squeeze_challenge()
    xn = pow(x, n, FR)

    # evaluations
    def ev(poly, rot=0):
        v = po.eval_poly(poly, rotate_omega(omega, x, rot))
        tr.write_scalar(v)
        T.evals.append(v)
        return v

    for col, rot in q.advice:
        ev(advice_polys[col], rot)
    for col, rot in q.fixed:
        ev(pk.fixed_polys[col], rot)
    h_poly = [0] * n
    for piece in reversed(h_pieces):
        h_poly = [(a * xn + b) % FR for a, b in zip(h_poly, piece)]
    ev(random_poly)
    for sp in pk.sigma_polys:
        ev(sp)
    last_rot = -(bf + 1)
    for si, zp in enumerate(perm_polys):
        ev(zp)
        ev(zp, 1)
        if si + 1 < len(perm_polys):
            ev(zp, last_rot)
    for L in lookups:
        ev(L["z_poly"])
        ev(L["z_poly"], 1)
        ev(L["pi_poly"])
        ev(L["pi_poly"], -1)
        ev(L["pt_poly"])

    # multi-open queries in create_proof's order: (rotation, polynomial)
    queries = [(rot, advice_polys[col]) for col, rot in q.advice]
    for zp in perm_polys:
        queries += [(0, zp), (1, zp)]
    for zp in list(reversed(perm_polys))[1:]:
        queries.append((last_rot, zp))
    for L in lookups:
        queries += [(0, L["z_poly"]), (0, L["pi_poly"]), (0, L["pt_poly"]), (-1, L["pi_poly"]), (1, L["z_poly"])]
    queries += [(rot, pk.fixed_polys[col]) for col, rot in q.fixed]
    queries += [(0, sp) for sp in pk.sigma_polys]
    queries += [(0, h_poly), (0, random_poly)]

    # ProverGWC::create_proof
    v = tr.squeeze_challenge()
    point_sets: List[Tuple[int, list]] = []  # construct_intermediate_sets: grouped by point, first-appearance order
    for rot, poly in queries:
        for entry in point_sets:
            if entry[0] == rot:
                entry[1].append(poly)
                break
        else:
            point_sets.append((rot, [poly]))
    for rot, polys in point_sets:
        z = rotate_omega(omega, x, rot)
        acc = [0] * n
        eval_acc = 0
        pv = 1
        for poly in polys:
            acc = [(a + pv * c) % FR for a, c in zip(acc, poly)]
            eval_acc = (eval_acc + pv * po.eval_poly(poly, z)) % FR
            pv = pv * v % FR
        acc[0] = (acc[0] - eval_acc) % FR
        w = params.commit(kate_division(acc, z))
        tr.write_point(w)
        T.commitments.append(w)
    T.challenges.update(theta=theta, beta=beta, gamma=gamma, y=y, x=x, v=v)
    T.columns.update(advice=advice, perm_z=perm_z, lookups=lookups, random_poly=random_poly, h_pieces=h_pieces)
    return tr.finalize()


# ---------------------------------------------------------------------------------------------------------------------
# create_proof once more, on Montgomery limb arrays with the C oracle's vector loops: the SAME algorithm, step for step, as
# create_proof() above (tests/test_pyprover.py checks byte equality), but at the reference's cost profile (C inner loops on
# all cores) instead of Python integer loops.  Used where k >= 16 and as the complete-prover CPU timing of bench.py.
# ---------------------------------------------------------------------------------------------------------------------
def _m(v: int) -> np.ndarray:
    return to_mont([v])[0]


def _rows_expr(e, n, fixed, advice, instance, challenges):
    """plonk::evaluation::evaluate over the n rows; columns are (n, 4) Montgomery arrays"""
    t = e[0]
    if t == "const":
        return np.tile(_m(e[1]), (n, 1))
    if t in ("fixed", "advice", "instance"):
        col = {"fixed": fixed, "advice": advice, "instance": instance}[t][e[1]]
        return np.ascontiguousarray(np.roll(col, -e[2], axis=0)) if e[2] else col
    if t == "challenge":
        return np.tile(_m(challenges[e[1]]), (n, 1))
    if t == "neg":
        return orc.fr_sub(np.zeros((n, 4), dtype=np.uint64), _rows_expr(e[1], n, fixed, advice, instance, challenges))
    if t == "sum":
        return orc.fr_add(_rows_expr(e[1], n, fixed, advice, instance, challenges), _rows_expr(e[2], n, fixed, advice, instance, challenges))
    if t == "prod":
        return orc.fr_mul(_rows_expr(e[1], n, fixed, advice, instance, challenges), _rows_expr(e[2], n, fixed, advice, instance, challenges))
    if t == "scaled":
        return orc.fr_mul(_rows_expr(e[1], n, fixed, advice, instance, challenges), np.tile(_m(e[2]), (n, 1)))
    raise ValueError(t)


def _pk_arrays(params: Params, pk: ProvingKey):
    """Montgomery images of the proving key's columns and the evaluator's resident cosets (built once per pk)"""
    if getattr(pk, "_arrays", None) is None:
        from de_b200 import plonk
        shape = pk.vk.shape
        dom = _cdomain(shape, params.k)
        fp = [to_mont(p) for p in pk.fixed_polys]
        sp = [to_mont(p) for p in pk.sigma_polys]
        desc, keep = plonk.marshal_pk_desc(shape, fp, sp)
        pk._arrays = dict(dom=dom, fixed_values=[to_mont(c) for c in pk.fixed_values], sigma_values=[to_mont(c) for c in pk.sigma_values],
                          fixed_polys=fp, sigma_polys=sp, opk=orc.Pk(dom, desc, keep),
                          omega_pows=orc.fr_geometric(_m(1), dom.omega, params.n))
    return pk._arrays


def create_proof_fast(params: Params, pk: ProvingKey, advice_values, instances, randoms: np.ndarray) -> bytes:
    """advice_values: (n, 4) Montgomery arrays (or int lists); randoms: (count, 4) Montgomery array of the Fr::random draws."""
    from de_b200 import plonk
    vk, shape, q = pk.vk, pk.vk.shape, pk.vk.queries
    n, k = params.n, params.k
    A = _pk_arrays(params, pk)
    dom = A["dom"]
    pdom = po.EvaluationDomain(shape.degree(), k)
    omega = pdom.omega
    bf = shape.blinding_factors
    usable = n - (bf + 1)
    tr = Transcript()
    rpos = [0]

    def draw(count):
        out = randoms[rpos[0]:rpos[0] + count]
        assert out.shape[0] == count, "random stream exhausted"
        rpos[0] += count
        return out

    def commit_point(basis_mont, scalars):
        return jac_mont_to_affine(orc.best_multiexp(np.ascontiguousarray(scalars), basis_mont[:scalars.shape[0]]))

    tr.common_scalar(vk.transcript_repr)
    instance_values = []
    for vals in instances:
        assert len(vals) <= usable, "InstanceTooLarge"
        col = np.zeros((n, 4), dtype=np.uint64)
        if len(vals):
            col[:len(vals)] = to_mont(vals)
        for v in vals:
            tr.common_scalar(v)
        instance_values.append(col)
    instance_polys = [dom.lagrange_to_coeff(c) for c in instance_values]
    advice = [np.ascontiguousarray(c, dtype=np.uint64).copy() if isinstance(c, np.ndarray) else to_mont(c) for c in advice_values]
    for col in advice:
        col[usable:] = draw(bf + 1)
    draw(len(advice))
    for col in advice:
        tr.write_point(commit_point(params.g_lagrange_mont, col))
    challenges: List[int] = []
    theta = tr.squeeze_challenge()
    theta_m = _m(theta)
    lookups = []
    for inp, tab in shape.lookups:
        def compress(exprs):
            acc = np.zeros((n, 4), dtype=np.uint64)
            for e in exprs:
                orc.fr_scale_add(acc, _rows_expr(e, n, A["fixed_values"], advice, instance_values, challenges), theta_m)
            return acc
        ci, ct = compress(inp), compress(tab)
        pi_i, pt_i = permute_expression_pair(from_mont(ci), from_mont(ct), usable)
        pi = np.concatenate([to_mont(pi_i), draw(bf + 1)])
        pt = np.concatenate([to_mont(pt_i), draw(bf + 1)])
        draw(2)
        tr.write_point(commit_point(params.g_lagrange_mont, pi))
        tr.write_point(commit_point(params.g_lagrange_mont, pt))
        lookups.append(dict(ci=ci, ct=ct, pi=pi, pt=pt))
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    beta_m, gamma_m = np.tile(_m(beta), (n, 1)), np.tile(_m(gamma), (n, 1))
    cols_any = {ADVICE: advice, FIXED: A["fixed_values"], INSTANCE: instance_values}
    chunk = shape.chunk_len
    perm_z = []
    deltaomega = 1
    last_z = _m(1)
    for s0 in range(0, len(shape.perm_columns), chunk):
        columns = shape.perm_columns[s0:s0 + chunk]
        modified = np.tile(_m(1), (n, 1))
        for ci_, (kind, index) in enumerate(columns):
            t = orc.fr_add(orc.fr_add(orc.fr_mul(beta_m, A["sigma_values"][s0 + ci_]), gamma_m), cols_any[kind][index])
            modified = orc.fr_mul(modified, t)
        modified = orc.fr_batch_invert(modified)
        for (kind, index) in columns:
            dw = orc.fr_mul(A["omega_pows"], np.tile(_m(deltaomega * beta % FR), (n, 1)))  # delta^j omega^i beta
            modified = orc.fr_mul(modified, orc.fr_add(orc.fr_add(dw, gamma_m), cols_any[kind][index]))
            deltaomega = deltaomega * FR_DELTA % FR
        z = orc.fr_running_product(modified, last_z)
        z[n - bf:] = draw(bf)
        last_z = z[n - (bf + 1)].copy()
        draw(1)
        tr.write_point(commit_point(params.g_lagrange_mont, z))
        perm_z.append(z)
    for L in lookups:
        den = orc.fr_batch_invert(orc.fr_mul(orc.fr_add(beta_m, L["pi"]), orc.fr_add(gamma_m, L["pt"])))
        prod = orc.fr_mul(orc.fr_mul(den, orc.fr_add(L["ci"], beta_m)), orc.fr_add(L["ct"], gamma_m))
        z = orc.fr_running_product(prod, _m(1))
        z[n - bf:] = draw(bf)
        draw(1)
        tr.write_point(commit_point(params.g_lagrange_mont, z))
        L["z"] = z
    random_poly = np.ascontiguousarray(draw(n))
    draw(1)
    tr.write_point(commit_point(params.g_mont, random_poly))
    y = tr.squeeze_challenge()

    advice_polys = [dom.lagrange_to_coeff(c) for c in advice]
    perm_polys = [dom.lagrange_to_coeff(z) for z in perm_z]
    for L in lookups:
        L["z_poly"], L["pi_poly"], L["pt_poly"] = dom.lagrange_to_coeff(L["z"]), dom.lagrange_to_coeff(L["pi"]), dom.lagrange_to_coeff(L["pt"])
    chs, keep2 = plonk.marshal_challenges(y, beta, gamma, theta, challenges)
    lookup_block = [L["z_poly"] for L in lookups] + [L["pi_poly"] for L in lookups] + [L["pt_poly"] for L in lookups]
    h_ext = A["opk"].evaluate_h(advice_polys, instance_polys, chs, perm_polys, lookup_block)
    h_coeff = dom.extended_to_coeff(dom.divide_by_vanishing(h_ext))
    n_pieces = shape.degree() - 1
    h_pieces = [np.ascontiguousarray(h_coeff[i * n:(i + 1) * n]) for i in range(n_pieces)]
    draw(n_pieces)
    for piece in h_pieces:
        tr.write_point(commit_point(params.g_mont, piece))
    x = tr.squeeze_challenge()
    xn = pow(x, n, FR)

    def point(rot):
        return _m(rotate_omega(omega, x, rot))

    def ev(poly, rot=0):
        tr.write_scalar(from_mont(orc.eval_polynomial(poly, point(rot)).reshape(1, 4))[0])

    for col, rot in q.advice:
        ev(advice_polys[col], rot)
    for col, rot in q.fixed:
        ev(A["fixed_polys"][col], rot)
    h_poly = np.zeros((n, 4), dtype=np.uint64)
    xn_m = _m(xn)
    for piece in reversed(h_pieces):
        orc.fr_scale_add(h_poly, piece, xn_m)
    ev(random_poly)
    for sp in A["sigma_polys"]:
        ev(sp)
    last_rot = -(bf + 1)
    for si, zp in enumerate(perm_polys):
        ev(zp)
        ev(zp, 1)
        if si + 1 < len(perm_polys):
            ev(zp, last_rot)
    for L in lookups:
        ev(L["z_poly"])
        ev(L["z_poly"], 1)
        ev(L["pi_poly"])
        ev(L["pi_poly"], -1)
        ev(L["pt_poly"])
    queries = [(rot, advice_polys[col]) for col, rot in q.advice]
    for zp in perm_polys:
        queries += [(0, zp), (1, zp)]
    for zp in list(reversed(perm_polys))[1:]:
        queries.append((last_rot, zp))
    for L in lookups:
        queries += [(0, L["z_poly"]), (0, L["pi_poly"]), (0, L["pt_poly"]), (-1, L["pi_poly"]), (1, L["z_poly"])]
    queries += [(rot, A["fixed_polys"][col]) for col, rot in q.fixed]
    queries += [(0, sp) for sp in A["sigma_polys"]]
    queries += [(0, h_poly), (0, random_poly)]
    v = tr.squeeze_challenge()
    point_sets: List[Tuple[int, list]] = []
    for rot, poly in queries:
        for entry in point_sets:
            if entry[0] == rot:
                entry[1].append(poly)
                break
        else:
            point_sets.append((rot, [poly]))
    for rot, polys in point_sets:
        acc = np.zeros((n, 4), dtype=np.uint64)
        pv = 1
        for poly in polys:
            orc.fr_axpy(acc, poly, _m(pv))
            pv = pv * v % FR
        # kate_division does not read the constant coefficient: subtracting the batched evaluation is a no-op for the quotient
        w = orc.kate_division(acc, point(rot))
        tr.write_point(commit_point(params.g_mont, w))
    assert rpos[0] == random_count(shape, n), "random draw count differs from the int-based prover"
    return tr.finalize()


def random_count(shape, n: int) -> int:
    """number of Fr::random draws of one create_proof"""
    bf, L = shape.blinding_factors, len(shape.lookups)
    return shape.n_advice * (bf + 2) + L * (2 * (bf + 1) + 2) + shape.n_perm_sets * (bf + 1) + L * (bf + 1) + n + 1 + (shape.degree() - 1)


# ---------------------------------------------------------------------------------------------------------------------
# verify_proof (VerifierGWC, single strategy)
# ---------------------------------------------------------------------------------------------------------------------
def l_i_range(omega: int, n: int, x: int, xn: int, rotations: Sequence[int]) -> List[int]:
    inv = batch_invert([(x - pow(omega, r, FR)) % FR for r in rotations])
    common = (xn - 1) * pow(n, -1, FR) % FR
    return [i * common % FR * pow(omega, r, FR) % FR for i, r in zip(inv, rotations)]


def verify_proof(params: Params, vk: VerifyingKey, instances: List[List[int]], proof: bytes) -> bool:
    shape, q = vk.shape, vk.queries
    n, k = params.n, params.k
    pdom = po.EvaluationDomain(shape.degree(), k)
    omega = pdom.omega
    bf = shape.blinding_factors
    chunk = shape.chunk_len
    tr = Transcript(proof)
    try:
        tr.common_scalar(vk.transcript_repr)
        for vals in instances:
            for v_ in vals:
                tr.common_scalar(v_)
        advice_commitments = [tr.read_point() for _ in range(shape.n_advice)]
        challenges: List[int] = []
        theta = tr.squeeze_challenge()
        lookups = [dict(pi_c=tr.read_point(), pt_c=tr.read_point()) for _ in shape.lookups]
        beta = tr.squeeze_challenge()
        gamma = tr.squeeze_challenge()
        n_sets = -(-len(shape.perm_columns) // chunk) if shape.perm_columns else 0
        perm_c = [tr.read_point() for _ in range(n_sets)]
        for L in lookups:
            L["z_c"] = tr.read_point()
        random_c = tr.read_point()
        y = tr.squeeze_challenge()
        h_c = [tr.read_point() for _ in range(shape.degree() - 1)]
        x = tr.squeeze_challenge()
        advice_evals = [tr.read_scalar() for _ in q.advice]
        fixed_evals = [tr.read_scalar() for _ in q.fixed]
        random_eval = tr.read_scalar()
        sigma_evals = [tr.read_scalar() for _ in shape.perm_columns]
        perm_evals = []
        for si in range(n_sets):
            e = dict(z=tr.read_scalar(), z_next=tr.read_scalar())
            if si + 1 < n_sets:
                e["z_last"] = tr.read_scalar()
            perm_evals.append(e)
        for L in lookups:
            L.update(z=tr.read_scalar(), z_next=tr.read_scalar(), pi=tr.read_scalar(), pi_inv=tr.read_scalar(), pt=tr.read_scalar())
    except ValueError:
        return False

    xn = pow(x, n, FR)
    l_evals = l_i_range(omega, n, x, xn, list(range(-(bf + 1), 1)))
    assert len(l_evals) == 2 + bf
    l_last, l_blind, l_0 = l_evals[0], sum(l_evals[1:1 + bf]) % FR, l_evals[1 + bf]
    # instance evaluations by Lagrange interpolation of the public inputs
    max_rot = max([r for _, r in q.instance] + [0])
    min_rot = min([r for _, r in q.instance] + [0])
    max_len = max([len(v_) for v_ in instances] + [0])
    l_i_s = l_i_range(omega, n, x, xn, list(range(-max_rot, max_len + abs(min_rot))))
    instance_evals = []
    for col, rot in q.instance:
        vals = instances[col]
        off = max_rot - rot
        instance_evals.append(sum(a * b for a, b in zip(vals, l_i_s[off:off + len(vals)])) % FR)

    def any_eval(kind, index):
        lst = {ADVICE: advice_evals, FIXED: fixed_evals, INSTANCE: instance_evals}[kind]
        return lst[q.index(kind, index, 0)]

    exprs = [eval_expr_point(g, q, fixed_evals, advice_evals, instance_evals, challenges) for g in shape.gates]
    active = (1 - (l_last + l_blind)) % FR
    if n_sets:
        exprs.append(l_0 * (1 - perm_evals[0]["z"]) % FR)
        zl = perm_evals[-1]["z"]
        exprs.append((zl * zl - zl) * l_last % FR)
        for si in range(1, n_sets):
            exprs.append((perm_evals[si]["z"] - perm_evals[si - 1]["z_last"]) * l_0 % FR)
        for si in range(n_sets):
            columns = shape.perm_columns[si * chunk:(si + 1) * chunk]
            left = perm_evals[si]["z_next"]
            for ci_, (kind, index) in enumerate(columns):
                left = left * ((any_eval(kind, index) + beta * sigma_evals[si * chunk + ci_] + gamma) % FR) % FR
            right = perm_evals[si]["z"]
            cur_delta = beta * x % FR * pow(FR_DELTA, si * chunk, FR) % FR
            for kind, index in columns:
                right = right * ((any_eval(kind, index) + cur_delta + gamma) % FR) % FR
                cur_delta = cur_delta * FR_DELTA % FR
            exprs.append((left - right) * active % FR)
    for L, (inp, tab) in zip(lookups, shape.lookups):
        def compress(es):
            acc = 0
            for e in es:
                acc = (acc * theta + eval_expr_point(e, q, fixed_evals, advice_evals, instance_evals, challenges)) % FR
            return acc
        left = L["z_next"] * ((L["pi"] + beta) % FR) % FR * ((L["pt"] + gamma) % FR) % FR
        right = L["z"] * ((compress(inp) + beta) % FR) % FR * ((compress(tab) + gamma) % FR) % FR
        exprs.append(l_0 * (1 - L["z"]) % FR)
        exprs.append(l_last * (L["z"] * L["z"] - L["z"]) % FR)
        exprs.append((left - right) * active % FR)
        exprs.append(l_0 * (L["pi"] - L["pt"]) % FR)
        exprs.append((L["pi"] - L["pt"]) * (L["pi"] - L["pi_inv"]) % FR * active % FR)
    expected_h = 0
    for e in exprs:
        expected_h = (expected_h * y + e) % FR
    expected_h = expected_h * pow((xn - 1) % FR, -1, FR) % FR
    # h commitment = sum_i xn^i * h_c[i] as a list of (scalar, point) terms
    h_msm = []
    for c in reversed(h_c):
        h_msm = [(s * xn % FR, p) for s, p in h_msm]
        h_msm.append((1, c))

    # queries: (rotation, commitment-as-MSM terms, eval)
    last_rot = -(bf + 1)
    queries = [(rot, [(1, advice_commitments[col])], advice_evals[i]) for i, (col, rot) in enumerate(q.advice)]
    for si in range(n_sets):
        queries += [(0, [(1, perm_c[si])], perm_evals[si]["z"]), (1, [(1, perm_c[si])], perm_evals[si]["z_next"])]
    for si in reversed(range(n_sets - 1)):
        queries.append((last_rot, [(1, perm_c[si])], perm_evals[si]["z_last"]))
    for L in lookups:
        queries += [(0, [(1, L["z_c"])], L["z"]), (0, [(1, L["pi_c"])], L["pi"]), (0, [(1, L["pt_c"])], L["pt"]),
                    (-1, [(1, L["pi_c"])], L["pi_inv"]), (1, [(1, L["z_c"])], L["z_next"])]
    queries += [(rot, [(1, vk.fixed_commitments[col])], fixed_evals[i]) for i, (col, rot) in enumerate(q.fixed)]
    queries += [(0, [(1, c)], e) for c, e in zip(vk.permutation_commitments, sigma_evals)]
    queries += [(0, h_msm, expected_h), (0, [(1, random_c)], random_eval)]

    v = tr.squeeze_challenge()
    point_sets: List[Tuple[int, list]] = []
    for rot, msm, e in queries:
        for entry in point_sets:
            if entry[0] == rot:
                entry[1].append((msm, e))
                break
        else:
            point_sets.append((rot, [(msm, e)]))
    try:
        ws = [tr.read_point() for _ in point_sets]
    except ValueError:
        return False
    if tr.pos != len(proof):
        pass  # trailing bytes are ignored by the reference's reader as well (it reads from a stream)
    u = tr.squeeze_challenge()
    commitment_multi: List[Tuple[int, object]] = []
    eval_multi = 0
    witness, witness_aux = [], []
    pu = 1
    for (rot, items), wi in zip(point_sets, ws):
        z = rotate_omega(omega, x, rot)
        pv = 1
        eval_batch = 0
        for msm, e in items:
            commitment_multi += [(s * pv % FR * pu % FR, p) for s, p in msm]
            eval_batch = (eval_batch + pv * e) % FR
            pv = pv * v % FR
        eval_multi = (eval_multi + pu * eval_batch) % FR
        witness_aux.append((pu * z % FR, wi))
        witness.append((pu, wi))
        pu = pu * u % FR

    def msm_eval(terms):
        if not terms:
            return None
        return po.msm_naive([s for s, _ in terms], [p for _, p in terms])

    left = msm_eval(witness)
    right = msm_eval(witness_aux + commitment_multi + [(eval_multi, po.g1_neg(params.g[0]))])
    # e(left, [s]_2) == e(right, [1]_2)
    return pairing.pairing_check([(params.s_g2, left), (params.g2, po.g1_neg(right))])
