"""Synthetic SATISFIED circuits with the constraint-system shape of the reference's benches, so
that create_proof can be run end to end — real proofs that a verifier accepts — without the halo2wrong front-end, which
is out of scope.

The reference's circuits (/root/reference/src/lib.rs:127-163 DelayEncryptCircuit, benches/mod_pow.rs:41-140 RSACircuit,
src/encryption/chip.rs:128-198 PoseidonEncCircuit) all compile to halo2wrong's MainGate (5 advice columns, one degree-3
gate over 9 fixed columns) plus, for the RSA ones, RangeChip's tagged lookup tables.  What the prover's cost depends on is
that shape, the number of used rows and the value distribution of the witness; this module builds an assignment with those
properties whose every constraint holds:

  * arithmetic rows: witness-like values and random selector constants, with s_constant solved so
    that the MainGate polynomial vanishes; e(omega X) couples consecutive rows through se_next;
  * range rows: a, b, c, d are sub-limbs below 2^bits of the row's tag, e is their composition (the gate checks it) and the
    four composition lookups find (tag, sub-limb) in the fixed table; overflow rows look e up under the overflow tag;
  * copy constraints between cells holding equal values (cycles of length 2 and 3) across all five advice columns.

Everything is seeded and pure Python integers (canonical field values); nothing here touches the GPU or any checker code.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List, Tuple

from . import plonk
from .plonk import ADVICE, FIXED, FR, INSTANCE


@dataclass
class Assignment:
    shape: plonk.ConstraintSystemShape
    k: int
    fixed: List[List[int]]                       # n_fixed columns of n canonical values (unusable rows are zero)
    advice: List[List[int]]                      # n_advice columns of n values (last blinding_factors + 1 rows unused)
    instances: List[List[int]]                   # public inputs per instance column (the benches pass one empty column)
    copies: List[Tuple[int, int, int, int]]      # (perm column position, row, perm column position, row)
    used_rows: int


def _witness_value(rng: random.Random) -> int:
    s = rng.randrange(100)
    if s < 45:
        return rng.randrange(1 << 8)
    if s < 80:
        return rng.randrange(1 << 64)
    if s < 90:
        return rng.randrange(1 << 134)
    return rng.randrange(FR)


def range_table(bit_lens=(8, 4, 1), overflow_bits=(6,)):
    """(tag, value) rows of the RangeChip stand-in table: tag 0 holds only 0; tag t >= 1 holds 0 .. 2^bits - 1."""
    rows = [(0, 0)]
    tags = {}
    t = 1
    for b in list(bit_lens) + list(overflow_bits):
        tags[t] = b
        rows += [(t, v) for v in range(1 << b)]
        t += 1
    return rows, tags


def satisfied_assignment(with_range_lookups: bool, k: int, seed: int, used_rows: int, uniform_values: bool = False,
                         copy_fraction: float = 0.25, n_public: int = 0) -> Assignment:
    """n_public > 0 exposes that many witness cells as public inputs: instance row i is copy-constrained to an advice cell
    (the reference's benches pass an empty instance column; MainGate's `expose_public` does exactly this)."""
    shape = plonk.main_gate_shape(with_range_lookups)
    n = 1 << k
    usable = n - (shape.blinding_factors + 1)
    used_rows = min(used_rows, usable)
    rng = random.Random(seed)
    F = [[0] * n for _ in range(shape.n_fixed)]
    A = [[0] * n for _ in range(shape.n_advice)]
    SA, SB, SC, SD, SE, SE_NEXT, S_MUL_AB, S_MUL_CD, S_CONST = range(9)
    kinds = ["arith"] * used_rows
    n_comp = 0
    if with_range_lookups:
        T_TAG, T_VALUE, TAG_COMP, TAG_OVER, S_COMP, S_OVER = range(9, 15)
        bit_lens, over = ((8, 4, 1), (6,)) if usable >= 400 else ((3, 2, 1), (2,))
        table, tags = range_table(bit_lens, over)
        assert len(table) <= usable, "table does not fit the usable rows"
        for r, (t, v) in enumerate(table):
            F[T_TAG][r], F[T_VALUE][r] = t, v
        n_comp = len(bit_lens)
        for r in range(used_rows):
            s = rng.randrange(100)
            kinds[r] = "range" if s < 30 else ("overflow" if s < 40 else "arith")
    value = (lambda: rng.randrange(FR)) if uniform_values else (lambda: _witness_value(rng))
    for r in range(used_rows):
        if kinds[r] == "arith":
            for c in range(5):
                A[c][r] = value()
        elif kinds[r] == "range":
            t = 1 + rng.randrange(n_comp)
            b = tags[t]
            limbs = [rng.randrange(1 << b) for _ in range(4)]
            for c in range(4):
                A[c][r] = limbs[c]
            A[4][r] = sum(l << (b * i) for i, l in enumerate(limbs))
            F[TAG_COMP][r], F[S_COMP][r] = t, 1
        else:
            t = n_comp + 1
            for c in range(4):
                A[c][r] = value()
            A[4][r] = rng.randrange(1 << tags[t])
            F[TAG_OVER][r], F[S_OVER][r] = t, 1
    # copy constraints: make the target cell equal to the source, then record the equality.  Targets are cells of
    # arithmetic rows (no lookup constrains them); sources are any used cell.
    copies = []
    arith_rows = [r for r in range(used_rows) if kinds[r] == "arith"]
    touched = set()
    if arith_rows and used_rows > 1:
        for _ in range(int(copy_fraction * used_rows)):
            src = (rng.randrange(5), rng.randrange(used_rows))
            cycle = [src]
            for _ in range(1 + (rng.randrange(4) == 0)):
                dst = (rng.randrange(5), arith_rows[rng.randrange(len(arith_rows))])
                if dst in touched or dst == src:
                    continue
                cycle.append(dst)
            for dst in cycle[1:]:
                A[dst[0]][dst[1]] = A[src[0]][src[1]]
                touched.add(dst)
                copies.append((src[0], src[1], dst[0], dst[1]))
            touched.add(src)
    # a range row's composition must survive the copies: only arithmetic-row cells were overwritten, so it does.
    # selectors of arithmetic rows, then s_constant so that the gate vanishes on every usable row
    for r in range(used_rows):
        if kinds[r] == "arith":
            for s in (SA, SB, SC, SD, SE):
                F[s][r] = rng.randrange(1, 1 << 16) if rng.randrange(4) else 0
            F[S_MUL_AB][r] = rng.randrange(2)
            F[S_MUL_CD][r] = rng.randrange(2)
            F[SE_NEXT][r] = rng.randrange(2) if r + 1 < used_rows else 0
        elif kinds[r] == "range":
            b = tags[F[TAG_COMP][r]]
            F[SA][r], F[SB][r], F[SC][r], F[SD][r] = 1, 1 << b, 1 << (2 * b), 1 << (3 * b)
            F[SE][r] = FR - 1
        # overflow rows: the gate is switched off (all selectors zero)
    for r in range(usable):
        a, b, c, d, e = (A[i][r] for i in range(5))
        e_next = A[4][(r + 1) % n]
        acc = (a * F[SA][r] + b * F[SB][r] + c * F[SC][r] + d * F[SD][r] + e * F[SE][r] + a * b % FR * F[S_MUL_AB][r]
               + c * d % FR * F[S_MUL_CD][r] + F[SE_NEXT][r] * e_next) % FR
        F[S_CONST][r] = (-acc) % FR
    instances = [[] for _ in range(shape.n_instance)]
    if n_public:
        inst_pos = next(i for i, (kind, _) in enumerate(shape.perm_columns) if kind == INSTANCE)
        for i in range(min(n_public, used_rows)):
            col, row = rng.randrange(5), rng.randrange(used_rows)
            instances[0].append(A[col][row])
            copies.append((inst_pos, i, col, row))
    return Assignment(shape, k, F, A, instances, copies, used_rows)


def check_assignment(asg: Assignment) -> None:
    """MockProver-style check of gates, lookups and copy constraints on the usable rows (raises AssertionError)."""
    shape, n = asg.shape, 1 << asg.k
    usable = n - (shape.blinding_factors + 1)
    cols = {"fixed": asg.fixed, "advice": asg.advice,
            "instance": [list(v) + [0] * (n - len(v)) for v in asg.instances] + [[0] * n for _ in range(shape.n_instance - len(asg.instances))]}

    def ev(e, r):
        t = e[0]
        if t == "const":
            return e[1]
        if t in cols:
            return cols[t][e[1]][(r + e[2]) % n]
        if t == "neg":
            return (-ev(e[1], r)) % FR
        if t == "sum":
            return (ev(e[1], r) + ev(e[2], r)) % FR
        if t == "prod":
            return ev(e[1], r) * ev(e[2], r) % FR
        if t == "scaled":
            return ev(e[1], r) * e[2] % FR
        raise ValueError(t)

    for g in shape.gates:
        for r in range(usable):
            assert ev(g, r) == 0, f"gate not satisfied on row {r}"
    for inp, tab in shape.lookups:
        table = {tuple(ev(e, r) for e in tab) for r in range(usable)}
        for r in range(usable):
            assert tuple(ev(e, r) for e in inp) in table, f"lookup input of row {r} not in table"
    anyc = {ADVICE: asg.advice, FIXED: asg.fixed, INSTANCE: cols["instance"]}
    for lc, lr, rc, rr in asg.copies:
        (k1, i1), (k2, i2) = shape.perm_columns[lc], shape.perm_columns[rc]
        assert anyc[k1][i1][lr] == anyc[k2][i2][rr], "copy constraint between unequal cells"


def mul_table_assignment(k: int, seed: int, used_rows: int) -> Assignment:
    """A second, differently shaped satisfied circuit for the generic prover paths the MainGate shapes do not reach:
    3 advice columns, NO instance column, 2 fixed columns, one gate q * (a * b - c) of degree 3, one single-expression lookup
    a in t (so cs.degree() = 4: three quotient pieces, permutation chunks of 2), and a permutation over the three advice
    columns AND the fixed table column (a copy constraint may tie an advice cell to a fixed cell)."""
    from .plonk import Advice, ConstraintSystemShape, Fixed, Neg, Prod, Sum
    n = 1 << k
    rng = random.Random(seed)
    q, t = Fixed(0), Fixed(1)
    a, b, c = Advice(0), Advice(1), Advice(2)
    shape = ConstraintSystemShape(n_fixed=2, n_advice=3, n_instance=0, gates=[Prod(q, Sum(Prod(a, b), Neg(c)))], lookups=[([a], [t])],
                                  perm_columns=[(ADVICE, 0), (ADVICE, 1), (ADVICE, 2), (FIXED, 1)])
    usable = n - (shape.blinding_factors + 1)
    used_rows = min(used_rows, usable)
    assert usable >= 16
    F = [[0] * n for _ in range(2)]
    A = [[0] * n for _ in range(3)]
    for r in range(16):
        F[1][r] = r  # table: 0 .. 15 (then zeros)
    copies = []
    for r in range(used_rows):
        F[0][r] = 1
        A[0][r], A[1][r] = rng.randrange(4), rng.randrange(4)
        A[2][r] = A[0][r] * A[1][r]
    for _ in range(used_rows // 3):
        i, j = rng.randrange(used_rows), rng.randrange(used_rows)
        if A[2][i] >= 16:
            continue  # the copied value becomes a lookup input: it must stay inside the table
        # a[j] := c[i]; row j's product is recomputed
        A[0][j] = A[2][i]
        A[2][j] = A[0][j] * A[1][j] % FR
        copies.append((2, i, 0, j))
    # copies recorded before a later overwrite of their source would be stale: keep only those that still hold
    copies = [cp for cp in copies if A[cp[0]][cp[1]] == A[cp[2]][cp[3]]]
    for _ in range(used_rows // 4):
        i = rng.randrange(used_rows)
        if A[0][i] < 16:
            copies.append((0, i, 3, A[0][i]))  # advice cell = fixed table cell holding the same value
    return Assignment(shape, k, F, A, [], copies, used_rows)
