"""Host-side mirror of the pieces of halo2_proofs::plonk the quotient evaluator needs:

  * Expression trees and GraphEvaluator — a restatement of plonk::evaluation::GraphEvaluator::add_expression, which compiles
    gate / lookup expressions into the (constants, rotations, calculations) program Evaluator::evaluate_h interprets.  In a
    Rust integration the already-compiled `pk.ev` is serialised instead; this compiler exists so that the
    Python tests and bench.py can describe constraint systems the same way halo2 does.
  * ConstraintSystemShape / ProvingKey — what de_pk_upload needs from a ProvingKey: fixed and sigma polynomials, the
    permutation column list, chunk_len = cs.degree() - 2, blinding_factors, the compiled programs.
  * main_gate_shape() — the constraint-system shape of the delay-encryption circuits: halo2wrong's MainGate (5 advice,
    9 fixed, one degree-3 gate) with or without RangeChip lookups.  The halo2wrong source is not
    available here, so the lookup expressions are shape-equivalent stand-ins (same column counts, degrees, rotations).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _lib

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FR_DELTA = 0x09226B6E22C6F0CA64EC26AAD4C86E715B5F898E5E963F25870E56BBE533E9A2
_R = 1 << 256

CONSTANT, INTERMEDIATE, FIXED, ADVICE, INSTANCE, CHALLENGE, BETA, GAMMA, THETA, Y, PREVIOUS = range(11)
ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE = range(8)


# ---- expressions (halo2_proofs::plonk::Expression) --------------------------------------------------------------
def Const(v): return ("const", v % FR)
def Fixed(col, rot=0): return ("fixed", col, rot)
def Advice(col, rot=0): return ("advice", col, rot)
def Instance(col, rot=0): return ("instance", col, rot)
def Challenge(i): return ("challenge", i)
def Neg(e): return ("neg", e)
def Sum(a, b): return ("sum", a, b)
def Prod(a, b): return ("prod", a, b)
def Scaled(e, v): return ("scaled", e, v % FR)


def sum_all(terms):
    acc = terms[0]
    for t in terms[1:]:
        acc = Sum(acc, t)
    return acc


def expr_degree(e) -> int:
    t = e[0]
    if t in ("const", "challenge"):
        return 0
    if t in ("fixed", "advice", "instance"):
        return 1
    if t in ("neg", "scaled"):
        return expr_degree(e[1])
    if t == "sum":
        return max(expr_degree(e[1]), expr_degree(e[2]))
    return expr_degree(e[1]) + expr_degree(e[2])


class GraphEvaluator:
    """plonk::evaluation::GraphEvaluator: constants start as [0, 1, 2]; identical calculations are shared."""

    def __init__(self):
        self.constants: List[int] = [0, 1, 2]
        self.rotations: List[int] = []
        self.calculations: List[tuple] = []  # (op, a, b, parts, target)
        self.num_intermediates = 0

    def add_rotation(self, rot: int) -> int:
        if rot in self.rotations:
            return self.rotations.index(rot)
        self.rotations.append(rot)
        return len(self.rotations) - 1

    def add_constant(self, c: int):
        c %= FR
        if c in self.constants:
            return (CONSTANT, self.constants.index(c), 0)
        self.constants.append(c)
        return (CONSTANT, len(self.constants) - 1, 0)

    def add_calculation(self, op, a, b=(CONSTANT, 0, 0), parts=()):
        key = (op, a, b, tuple(parts))
        for c in self.calculations:
            if c[:4] == key:
                return (INTERMEDIATE, c[4], 0)
        target = self.num_intermediates
        self.num_intermediates += 1
        self.calculations.append(key + (target,))
        return (INTERMEDIATE, target, 0)

    def add_expression(self, e):
        zero, one, two = (CONSTANT, 0, 0), (CONSTANT, 1, 0), (CONSTANT, 2, 0)
        t = e[0]
        if t == "const":
            return self.add_constant(e[1])
        if t in ("fixed", "advice", "instance"):
            kind = {"fixed": FIXED, "advice": ADVICE, "instance": INSTANCE}[t]
            return self.add_calculation(STORE, (kind, e[1], self.add_rotation(e[2])))
        if t == "challenge":
            return self.add_calculation(STORE, (CHALLENGE, e[1], 0))
        if t == "neg":
            if e[1][0] == "const":
                return self.add_constant(-e[1][1])
            ra = self.add_expression(e[1])
            return ra if ra == zero else self.add_calculation(NEGATE, ra)
        if t == "sum":
            a, b = e[1], e[2]
            if b[0] == "neg":
                ra, rb = self.add_expression(a), self.add_expression(b[1])
                if ra == zero:
                    return self.add_calculation(NEGATE, rb)
                if rb == zero:
                    return ra
                return self.add_calculation(SUB, ra, rb)
            if a[0] == "neg":
                ra, rb = self.add_expression(a[1]), self.add_expression(b)
                if ra == zero:
                    return rb
                if rb == zero:
                    return self.add_calculation(NEGATE, ra)
                return self.add_calculation(SUB, rb, ra)
            ra, rb = self.add_expression(a), self.add_expression(b)
            if ra == zero:
                return rb
            if rb == zero:
                return ra
            return self.add_calculation(ADD, ra, rb) if ra <= rb else self.add_calculation(ADD, rb, ra)
        if t == "prod":
            ra, rb = self.add_expression(e[1]), self.add_expression(e[2])
            if ra == zero or rb == zero:
                return zero
            if ra == one:
                return rb
            if rb == one:
                return ra
            if ra == two:
                return self.add_calculation(DOUBLE, rb)
            if rb == two:
                return self.add_calculation(DOUBLE, ra)
            if ra == rb:
                return self.add_calculation(SQUARE, ra)
            return self.add_calculation(MUL, ra, rb) if ra <= rb else self.add_calculation(MUL, rb, ra)
        if t == "scaled":
            if e[2] == 0:
                return zero
            if e[2] == 1:
                return self.add_expression(e[1])
            cst = self.add_constant(e[2])
            ra = self.add_expression(e[1])
            return self.add_calculation(MUL, ra, cst)
        raise ValueError(f"unknown expression {t}")


def compile_gates(polys: Sequence) -> GraphEvaluator:
    """Evaluator::new, custom gates: value = Horner(PreviousValue, gate polynomials, Y)."""
    g = GraphEvaluator()
    parts = [g.add_expression(p) for p in polys]
    g.add_calculation(HORNER, (PREVIOUS, 0, 0), (Y, 0, 0), parts)
    return g


def compile_lookup(input_exprs: Sequence, table_exprs: Sequence) -> GraphEvaluator:
    """Evaluator::new, one lookup: (theta-compressed input + beta) * (theta-compressed table + gamma)."""
    g = GraphEvaluator()

    def lc(exprs):
        parts = [g.add_expression(e) for e in exprs]
        return g.add_calculation(HORNER, (CONSTANT, 0, 0), (THETA, 0, 0), parts)

    ci = lc(input_exprs)
    ct = lc(table_exprs)
    right_gamma = g.add_calculation(ADD, ct, (GAMMA, 0, 0))
    left_beta = g.add_calculation(ADD, ci, (BETA, 0, 0))
    g.add_calculation(MUL, left_beta, right_gamma)
    return g


def compile_compress(exprs: Sequence) -> GraphEvaluator:
    """lookup::prover::commit_permuted's compress_expressions: fold(0, |acc, e| acc * theta + e) over the argument's
    expressions, as a graph whose result is the compressed value of a row."""
    g = GraphEvaluator()
    parts = [g.add_expression(e) for e in exprs]
    g.add_calculation(HORNER, (CONSTANT, 0, 0), (THETA, 0, 0), parts)
    return g


@dataclass
class ConstraintSystemShape:
    n_fixed: int
    n_advice: int
    n_instance: int
    gates: List                    # gate polynomials (expressions)
    lookups: List                  # [(input_exprs, table_exprs)]
    perm_columns: List             # [(kind, index)] in permutation order
    blinding_factors: int = 5

    def degree(self) -> int:
        """cs.degree(): max over the permutation argument (3 when non-empty), lookups (input + table + 1, at least 4... as halo2
        computes it) and gates; at least 3."""
        deg = 3 if self.perm_columns else 1
        for inp, tab in self.lookups:
            di = max(expr_degree(e) for e in inp)
            dt = max(expr_degree(e) for e in tab)
            deg = max(deg, max(4, 2 + di + dt))  # lookup::Argument::required_degree
        for g in self.gates:
            deg = max(deg, expr_degree(g))
        return deg

    @property
    def chunk_len(self) -> int:
        return self.degree() - 2

    @property
    def n_perm_sets(self) -> int:
        return -(-len(self.perm_columns) // self.chunk_len) if self.perm_columns else 0


def main_gate_shape(with_range_lookups: bool) -> ConstraintSystemShape:
    """MainGate (+ RangeChip) shape of the reference circuits: 5 advice, 1 instance, 6 permutation columns, 9 or 15 fixed."""
    a, b, c, d, e = (Advice(i) for i in range(5))
    e_next = Advice(4, 1)
    sa, sb, sc, sd, se, se_next, s_mul_ab, s_mul_cd, s_constant = (Fixed(i) for i in range(9))
    gate = sum_all([Prod(a, sa), Prod(b, sb), Prod(c, sc), Prod(d, sd), Prod(e, se), Prod(Prod(a, b), s_mul_ab),
                    Prod(Prod(c, d), s_mul_cd), Prod(se_next, e_next), s_constant])
    lookups = []
    n_fixed = 9
    if with_range_lookups:
        n_fixed = 15
        t_tag, t_value, tag_comp, tag_over, s_comp, s_over = (Fixed(i) for i in range(9, 15))
        for col in range(4):  # composition lookups on a, b, c, d
            lookups.append(([Prod(s_comp, tag_comp), Prod(s_comp, Advice(col))], [t_tag, t_value]))
        lookups.append(([Prod(s_over, tag_over), Prod(s_over, Advice(4))], [t_tag, t_value]))  # overflow lookup
    perm = [(ADVICE, i) for i in range(5)] + [(INSTANCE, 0)]
    return ConstraintSystemShape(n_fixed, 5, 1, [gate], lookups, perm)


def collect_queries(shape: "ConstraintSystemShape"):
    """cs.advice_queries / cs.fixed_queries / cs.instance_queries as (column, rotation) lists in first-query order:
    enable_equality queries every permutation column at Rotation::cur() first (ConstraintSystem::query_any_index), then the
    gates' and the lookups' expressions are walked left to right.  A Rust integration passes the real lists of its
    ConstraintSystem instead (their order is fixed by the circuit's configure(), which is not restated here)."""
    q = {"advice": [], "fixed": [], "instance": []}

    def add(kind, col, rot):
        if (col, rot) not in q[kind]:
            q[kind].append((col, rot))

    for kind, index in shape.perm_columns:
        add({ADVICE: "advice", FIXED: "fixed", INSTANCE: "instance"}[kind], index, 0)

    def walk(e):
        t = e[0]
        if t in q:
            add(t, e[1], e[2])
        elif t in ("neg", "scaled"):
            walk(e[1])
        elif t in ("sum", "prod"):
            walk(e[1])
            walk(e[2])

    for g in shape.gates:
        walk(g)
    for inp, tab in shape.lookups:
        for e in list(inp) + list(tab):
            walk(e)
    return q["advice"], q["fixed"], q["instance"]


# ---- marshalling to the C ABI (include/de_b200.h) -----------------------------------------------------------------
class _ValueSource(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("index", C.c_uint32), ("rotation", C.c_uint32)]


class _Calculation(C.Structure):
    _fields_ = [("op", C.c_uint32), ("a", _ValueSource), ("b", _ValueSource), ("horner_first", C.c_uint32),
                ("horner_len", C.c_uint32), ("target", C.c_uint32)]


class _Graph(C.Structure):
    _fields_ = [("constants", C.c_void_p), ("n_constants", C.c_uint32), ("rotations", C.c_void_p), ("n_rotations", C.c_uint32),
                ("calcs", C.c_void_p), ("n_calcs", C.c_uint32), ("horner_parts", C.c_void_p), ("n_horner_parts", C.c_uint32),
                ("n_intermediates", C.c_uint32)]


class _Fr(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


class _PkDesc(C.Structure):
    _fields_ = [("n_fixed", C.c_uint32), ("n_advice", C.c_uint32), ("n_instance", C.c_uint32), ("fixed_coeff", C.c_void_p),
                ("n_perm_columns", C.c_uint32), ("perm_column_kind", C.c_void_p), ("perm_column_index", C.c_void_p),
                ("sigma_coeff", C.c_void_p), ("chunk_len", C.c_uint32), ("blinding_factors", C.c_uint32), ("delta", _Fr),
                ("gates", _Graph), ("n_lookups", C.c_uint32), ("lookups", C.c_void_p)]


class _Challenges(C.Structure):
    _fields_ = [("y", _Fr), ("beta", _Fr), ("gamma", _Fr), ("theta", _Fr), ("challenges", C.c_void_p), ("n_challenges", C.c_uint32)]


def mont_limbs(v: int):
    m = (v % FR) * _R % FR
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def _fr_struct(v: int) -> _Fr:
    f = _Fr()
    for i, l in enumerate(mont_limbs(v)):
        f.l[i] = l
    return f


def _marshal_graph(g: GraphEvaluator, keep: list) -> _Graph:
    consts = np.array([mont_limbs(c) for c in g.constants], dtype=np.uint64).reshape(-1, 4)
    rots = np.array(g.rotations, dtype=np.int32)
    calcs = (_Calculation * max(len(g.calculations), 1))()
    parts_flat = []
    for i, (op, a, b, parts, target) in enumerate(g.calculations):
        calcs[i].op = op
        calcs[i].a = _ValueSource(*a)
        calcs[i].b = _ValueSource(*b)
        calcs[i].horner_first = len(parts_flat)
        calcs[i].horner_len = len(parts)
        calcs[i].target = target
        parts_flat.extend(parts)
    hp = (_ValueSource * max(len(parts_flat), 1))()
    for i, p in enumerate(parts_flat):
        hp[i] = _ValueSource(*p)
    keep += [consts, rots, calcs, hp]
    out = _Graph()
    out.constants = consts.ctypes.data
    out.n_constants = len(g.constants)
    out.rotations = rots.ctypes.data if len(g.rotations) else None
    out.n_rotations = len(g.rotations)
    out.calcs = C.addressof(calcs)
    out.n_calcs = len(g.calculations)
    out.horner_parts = C.addressof(hp)
    out.n_horner_parts = len(parts_flat)
    out.n_intermediates = g.num_intermediates
    return out


def marshal_pk_desc(shape: ConstraintSystemShape, fixed_coeff: Sequence[np.ndarray], sigma_coeff: Sequence[np.ndarray]):
    """-> (_PkDesc, keepalive list).  Polynomials are (n, 4) uint64 Montgomery arrays in coefficient form."""
    keep: list = []
    assert len(fixed_coeff) == shape.n_fixed and len(sigma_coeff) == len(shape.perm_columns)
    fixed = [np.ascontiguousarray(f, dtype=np.uint64) for f in fixed_coeff]
    sigma = [np.ascontiguousarray(s, dtype=np.uint64) for s in sigma_coeff]
    fptr = (C.c_void_p * max(len(fixed), 1))(*[f.ctypes.data for f in fixed])
    sptr = (C.c_void_p * max(len(sigma), 1))(*[s.ctypes.data for s in sigma])
    kinds = np.array([k for k, _ in shape.perm_columns], dtype=np.uint32)
    idxs = np.array([i for _, i in shape.perm_columns], dtype=np.uint32)
    gates = compile_gates(shape.gates)
    lookups = [compile_lookup(i, t) for i, t in shape.lookups]
    lg = (_Graph * max(len(lookups), 1))()
    for i, g in enumerate(lookups):
        lg[i] = _marshal_graph(g, keep)
    d = _PkDesc()
    d.n_fixed, d.n_advice, d.n_instance = shape.n_fixed, shape.n_advice, shape.n_instance
    d.fixed_coeff = C.addressof(fptr)
    d.n_perm_columns = len(shape.perm_columns)
    d.perm_column_kind = kinds.ctypes.data if len(kinds) else None
    d.perm_column_index = idxs.ctypes.data if len(idxs) else None
    d.sigma_coeff = C.addressof(sptr)
    d.chunk_len = shape.chunk_len
    d.blinding_factors = shape.blinding_factors
    d.delta = _fr_struct(FR_DELTA)
    d.gates = _marshal_graph(gates, keep)
    d.n_lookups = len(lookups)
    d.lookups = C.addressof(lg)
    keep += [fixed, sigma, fptr, sptr, kinds, idxs, lg]
    return d, keep


class _ProverDesc(C.Structure):
    _fields_ = [("n_advice_queries", C.c_uint32), ("advice_query_column", C.c_void_p), ("advice_query_rotation", C.c_void_p),
                ("n_fixed_queries", C.c_uint32), ("fixed_query_column", C.c_void_p), ("fixed_query_rotation", C.c_void_p),
                ("lookup_input_graphs", C.c_void_p), ("lookup_table_graphs", C.c_void_p), ("transcript_repr", _Fr)]


def marshal_challenges(y: int, beta: int, gamma: int, theta: int, challenges: Sequence[int] = ()):
    ch = _Challenges()
    ch.y, ch.beta, ch.gamma, ch.theta = _fr_struct(y), _fr_struct(beta), _fr_struct(gamma), _fr_struct(theta)
    arr = np.array([mont_limbs(c) for c in challenges], dtype=np.uint64).reshape(-1, 4)
    ch.challenges = arr.ctypes.data if len(challenges) else None
    ch.n_challenges = len(challenges)
    return ch, [arr]


def _ptr_array(polys):
    polys = [np.ascontiguousarray(p, dtype=np.uint64) for p in polys]
    arr = (C.c_void_p * max(len(polys), 1))(*[p.ctypes.data for p in polys])
    return arr, polys


class ProvingKey:
    """The evaluator's view of plonk::ProvingKey, staged in HBM (de_pk_upload)."""

    def __init__(self, domain, shape: ConstraintSystemShape, fixed_coeff, sigma_coeff):
        if domain.j != shape.degree():
            raise ValueError(f"EvaluationDomain was built for degree {domain.j}, constraint system has degree {shape.degree()}")
        self.domain, self.shape, self.ctx = domain, shape, domain.ctx
        desc, keep = marshal_pk_desc(shape, fixed_coeff, sigma_coeff)
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.de_pk_upload(domain.h, C.byref(desc), C.byref(h)))
        self.h = h
        del keep

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.de_pk_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def evaluate_h(self, advice_coeff, instance_coeff, y, beta, gamma, theta, perm_z_coeff, lookup_coeff, challenges=()):
        """Evaluator::evaluate_h for one proof.  lookup_coeff: [all product z | all permuted inputs | all permuted tables]."""
        s = self.shape
        assert len(advice_coeff) == s.n_advice and len(instance_coeff) == s.n_instance
        assert len(perm_z_coeff) == s.n_perm_sets and len(lookup_coeff) == 3 * len(s.lookups)
        ch, k1 = marshal_challenges(y, beta, gamma, theta, challenges)
        a, k2 = _ptr_array(advice_coeff)
        i, k3 = _ptr_array(instance_coeff)
        z, k4 = _ptr_array(perm_z_coeff)
        l, k5 = _ptr_array(lookup_coeff)
        out = np.empty((self.domain.extended_n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.L.de_evaluate_h(self.h, a, i, C.byref(ch), z, l, out.ctypes.data_as(C.c_void_p)))
        return out

    def evaluate_h_dev(self, d_advice, d_instance, y, beta, gamma, theta, d_perm_z, d_lookup, d_h_ext, stride=None, challenges=()):
        ch, k1 = marshal_challenges(y, beta, gamma, theta, challenges)

        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None else None

        self.ctx.check(self.ctx.L.de_evaluate_h_dev(self.h, p(d_advice), p(d_instance), C.byref(ch), p(d_perm_z), p(d_lookup),
                                                    stride or self.domain.n, p(d_h_ext)))

    def extend_dev(self, d_advice, d_instance, d_perm_z, d_lookup, stride=None):
        """first half of evaluate_h: coeff_to_extended of every per-proof polynomial into the pk's HBM workspace"""
        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None else None

        self.ctx.check(self.ctx.L.de_pk_extend_dev(self.h, p(d_advice), p(d_instance), p(d_perm_z), p(d_lookup), stride or self.domain.n))

    def evaluate_h_rows_dev(self, y, beta, gamma, theta, d_h_ext, challenges=()):
        """second half: the fused row kernel over the extended domain"""
        ch, k1 = marshal_challenges(y, beta, gamma, theta, challenges)
        self.ctx.check(self.ctx.L.de_evaluate_h_rows_dev(self.h, C.byref(ch), C.c_void_p(d_h_ext.data_ptr())))


class Prover:
    """plonk::create_proof behind the C ABI (de_prover_create / de_create_proof): KZG + ProverGWC + Blake2bWrite.

    params: de_b200.ParamsKZG, pk: ProvingKey (same Context).  advice_queries / fixed_queries: the ConstraintSystem's query
    lists (collect_queries(shape) for the synthetic shapes).  transcript_repr: vk.transcript_repr as an integer."""

    def __init__(self, params, pk: ProvingKey, advice_queries, fixed_queries, transcript_repr: int):
        self.params, self.pk, self.ctx = params, pk, pk.ctx
        shape = pk.shape
        keep: list = []
        aq_c = np.array([c for c, _ in advice_queries], dtype=np.uint32)
        aq_r = np.array([r for _, r in advice_queries], dtype=np.int32)
        fq_c = np.array([c for c, _ in fixed_queries], dtype=np.uint32)
        fq_r = np.array([r for _, r in fixed_queries], dtype=np.int32)
        nl = len(shape.lookups)
        gi = (_Graph * max(nl, 1))()
        gt = (_Graph * max(nl, 1))()
        for i, (inp, tab) in enumerate(shape.lookups):
            gi[i] = _marshal_graph(compile_compress(inp), keep)
            gt[i] = _marshal_graph(compile_compress(tab), keep)
        d = _ProverDesc()
        d.n_advice_queries, d.n_fixed_queries = len(aq_c), len(fq_c)
        d.advice_query_column = aq_c.ctypes.data if len(aq_c) else None
        d.advice_query_rotation = aq_r.ctypes.data if len(aq_r) else None
        d.fixed_query_column = fq_c.ctypes.data if len(fq_c) else None
        d.fixed_query_rotation = fq_r.ctypes.data if len(fq_r) else None
        d.lookup_input_graphs = C.addressof(gi)
        d.lookup_table_graphs = C.addressof(gt)
        d.transcript_repr = _fr_struct(transcript_repr)
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.de_prover_create(params.h, pk.h, C.byref(d), C.byref(h)))
        self.h = h
        self.random_count = int(self.ctx.L.de_prover_random_count(h))
        self.proof_size = int(self.ctx.L.de_prover_proof_size(h))
        self._proof = np.zeros(self.proof_size, dtype=np.uint8)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.de_prover_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create_proof(self, advice, instances, randoms) -> bytes:
        """advice: list of (n, 4) uint64 Montgomery columns (numpy, or pinned torch tensors); instances: list of (len, 4)
        arrays (may be empty); randoms: (>= random_count, 4) Montgomery field elements = the Fr::random draws in order."""
        def addr(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

        adv = [a if hasattr(a, "data_ptr") else np.ascontiguousarray(a, dtype=np.uint64) for a in advice]
        ins = [np.ascontiguousarray(np.asarray(i, dtype=np.uint64).reshape(-1, 4)) for i in instances]
        while len(ins) < self.pk.shape.n_instance:
            ins.append(np.zeros((0, 4), dtype=np.uint64))
        rnd = randoms if hasattr(randoms, "data_ptr") else np.ascontiguousarray(randoms, dtype=np.uint64)
        n_rnd = (rnd.numel() if hasattr(rnd, "numel") else rnd.size) // 4
        ap = (C.c_void_p * max(len(adv), 1))(*[addr(a) for a in adv])
        ip = (C.c_void_p * max(len(ins), 1))(*[i.ctypes.data for i in ins])
        il = (C.c_size_t * max(len(ins), 1))(*[i.shape[0] for i in ins])
        m = C.c_size_t()
        self.ctx.check(self.ctx.L.de_create_proof(self.h, ap, ip, il, C.c_void_p(addr(rnd)), n_rnd,
                                                  self._proof.ctypes.data_as(C.c_void_p), self._proof.size, C.byref(m)))
        return self._proof[: m.value].tobytes()

    def create_proof_dev(self, d_advice, d_randoms, instances=(), advice_stride=None) -> bytes:
        """advice columns (n_advice, n, 4) and random draws (>= random_count, 4) as CUDA tensors already in HBM"""
        ins = [np.ascontiguousarray(np.asarray(i, dtype=np.uint64).reshape(-1, 4)) for i in instances]
        while len(ins) < self.pk.shape.n_instance:
            ins.append(np.zeros((0, 4), dtype=np.uint64))
        ip = (C.c_void_p * max(len(ins), 1))(*[i.ctypes.data for i in ins])
        il = (C.c_size_t * max(len(ins), 1))(*[i.shape[0] for i in ins])
        m = C.c_size_t()
        self.ctx.check(self.ctx.L.de_create_proof_dev(self.h, C.c_void_p(d_advice.data_ptr()), advice_stride or self.pk.domain.n, ip, il,
                                                      C.c_void_p(d_randoms.data_ptr()), d_randoms.numel() // 4,
                                                      self._proof.ctypes.data_as(C.c_void_p), self._proof.size, C.byref(m)))
        return self._proof[: m.value].tobytes()
