"""MockProver-style checker (TEST INFRASTRUCTURE ONLY, like everything under oracle/): does an assignment satisfy the
constraint system?  What halo2_proofs::dev::MockProver::verify checks for the circuits of the reference
(/root/reference/src/lib.rs:353 mock_prover_verify, src/big_integer/chip.rs:1460-1466 MockProver::run) restated over integer
columns: every gate polynomial vanishes on every usable row, every lookup input tuple occurs in the table, every copy
constraint joins equal cells.  Gate / lookup expressions (de_b200.plonk's tuples) are compiled once to Python source, so a
2^16-row circuit is checked in about a second."""
from __future__ import annotations

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _src(e) -> str:
    t = e[0]
    if t == "const":
        return str(e[1])
    if t in ("fixed", "advice", "instance"):
        return f"{t}[{e[1]}][(r + {e[2]}) % n]"
    if t == "neg":
        return f"(-({_src(e[1])}))"
    if t == "sum":
        return f"({_src(e[1])} + {_src(e[2])})"
    if t == "prod":
        return f"({_src(e[1])} * {_src(e[2])})"
    if t == "scaled":
        return f"({_src(e[1])} * {e[2]})"
    raise ValueError(t)


def _compile(exprs):
    body = ", ".join(f"({_src(e)}) % FR" for e in exprs)
    return eval(f"lambda fixed, advice, instance, r, n: ({body},)", {"FR": FR})


def check(shape, k: int, fixed, advice, instances, copies, blinding_factors: int = 5):
    """fixed / advice: lists of columns (n canonical ints each); instances: list of (short) columns; copies: (lcol, lrow, rcol,
    rrow) over shape.perm_columns.  Raises AssertionError naming the first violated constraint."""
    n = 1 << k
    usable = n - (blinding_factors + 1)
    inst = [list(v) + [0] * (n - len(v)) for v in instances]
    while len(inst) < shape.n_instance:
        inst.append([0] * n)
    gates = _compile(shape.gates)
    for r in range(usable):
        vals = gates(fixed, advice, inst, r, n)
        if any(vals):
            raise AssertionError(f"gate {[i for i, v in enumerate(vals) if v][0]} is not satisfied on row {r}")
    for li, (inp, tab) in enumerate(shape.lookups):
        fin, ftab = _compile(inp), _compile(tab)
        table = {ftab(fixed, advice, inst, r, n) for r in range(usable)}
        for r in range(usable):
            if fin(fixed, advice, inst, r, n) not in table:
                raise AssertionError(f"lookup {li}: the input of row {r} is not in the table")
    cols = {3: advice, 2: fixed, 4: inst}  # plonk.ADVICE / FIXED / INSTANCE
    for lc, lr, rc, rr in copies:
        (k1, i1), (k2, i2) = shape.perm_columns[lc], shape.perm_columns[rc]
        if cols[k1][i1][lr] != cols[k2][i2][rr]:
            raise AssertionError(f"copy constraint between unequal cells ({lc}, {lr}) and ({rc}, {rr})")
