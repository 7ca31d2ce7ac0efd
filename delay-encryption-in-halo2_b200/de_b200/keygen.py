"""Host-side mirror of halo2_proofs::plonk::{keygen_vk, keygen_pk} for the pieces that feed the GPU prover
(/root/reference/benches/delay_enc.rs:86,103): the permutation assembly (copy constraints -> sigma columns), the fixed /
sigma polynomials in coefficient form (lagrange_to_coeff on the device) and the verifying key's commitments
(commit_lagrange on the device).  Field values enter as canonical Python integers and are converted to Montgomery limbs
on the GPU (de_fr_vec_op TO_MONT); nothing here computes field arithmetic on the CPU beyond building the sigma labels
delta^col * omega^row, which keygen does once per circuit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import EvaluationDomain, ParamsKZG
from .plonk import FR, FR_DELTA, ConstraintSystemShape, ProvingKey, Prover, collect_queries


def canonical_limbs(vals: Sequence[int]) -> np.ndarray:
    """(len, 4) uint64 little-endian limbs of canonical integers (NOT Montgomery)"""
    raw = b"".join(int(v % FR).to_bytes(32, "little") for v in vals)
    return np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4).copy()


class PermutationAssembly:
    """permutation::keygen::Assembly: every cell starts as its own cycle; copy() merges the cycles of two cells."""

    def __init__(self, n_columns: int, n: int):
        self.n_columns, self.n = n_columns, n
        self.mapping = np.arange(n_columns * n, dtype=np.int64)  # cell id = column * n + row -> next cell of its cycle
        self.aux = np.arange(n_columns * n, dtype=np.int64)      # representative of the cell's cycle
        self.sizes = np.ones(n_columns * n, dtype=np.int64)

    def copy(self, left_column: int, left_row: int, right_column: int, right_row: int):
        l, r = left_column * self.n + left_row, right_column * self.n + right_row
        if self.aux[l] == self.aux[r]:
            return
        left_cycle, right_cycle = int(self.aux[l]), int(self.aux[r])
        if self.sizes[left_cycle] < self.sizes[right_cycle]:
            left_cycle, right_cycle = right_cycle, left_cycle
        self.sizes[left_cycle] += self.sizes[right_cycle]
        i = right_cycle
        while True:
            self.aux[i] = left_cycle
            i = int(self.mapping[i])
            if i == right_cycle:
                break
        self.mapping[l], self.mapping[r] = self.mapping[r], self.mapping[l]

    def sigma_values(self, omega: int) -> List[List[int]]:
        """permutations[col][row] = delta^(mapped column) * omega^(mapped row)"""
        n = self.n
        omega_pows = [1] * n
        for i in range(1, n):
            omega_pows[i] = omega_pows[i - 1] * omega % FR
        delta_pows = [pow(FR_DELTA, c, FR) for c in range(self.n_columns)]
        out = []
        for c in range(self.n_columns):
            col = []
            for r in range(n):
                m = int(self.mapping[c * n + r])
                mc, mr = divmod(m, n)
                col.append(omega_pows[mr] if mc == 0 else delta_pows[mc] * omega_pows[mr] % FR)
            out.append(col)
        return out


@dataclass
class Keys:
    params: ParamsKZG
    domain: EvaluationDomain
    pk: ProvingKey
    prover: Prover
    fixed_commitments: np.ndarray        # (n_fixed, 64) canonical x || y (vk.fixed_commitments)
    permutation_commitments: np.ndarray  # (n_perm_columns, 64)      (vk.permutation.commitments)
    host: dict = None                    # what another context needs to stage the same keys (see clone_on)

    def clone_on(self, ctx) -> "Keys":
        """the same ParamsKZG / ProvingKey staged on another context (one per in-flight proof or per GPU)"""
        h = self.host
        params = ParamsKZG(h["k"], h["g"], h["g_lagrange"], ctx)
        domain = EvaluationDomain(h["shape"].degree(), h["k"], ctx)
        pk = ProvingKey(domain, h["shape"], h["fixed_polys"], h["sigma_polys"])
        prover = Prover(params, pk, h["advice_queries"], h["fixed_queries"], h["transcript_repr"])
        return Keys(params, domain, pk, prover, self.fixed_commitments, self.permutation_commitments, h)

    def close(self):
        self.prover.close()
        self.pk.close()
        self.domain.close()
        self.params.close()


def keygen(ctx, shape: ConstraintSystemShape, k: int, g, g_lagrange, fixed_values: Sequence[Sequence[int]],
           copies: Sequence[Tuple[int, int, int, int]], transcript_repr: int, queries=None) -> Keys:
    """keygen_vk + keygen_pk + the prover object.  fixed_values: canonical integers per fixed column (n each)."""
    n = 1 << k
    domain = EvaluationDomain(shape.degree(), k, ctx)
    omega = int.from_bytes(ctx.fr_from_mont(domain.omega.reshape(1, 4)).tobytes(), "little")
    asm = PermutationAssembly(len(shape.perm_columns), n)
    for lc, lr, rc, rr in copies:
        asm.copy(lc, lr, rc, rr)
    sigma_values = asm.sigma_values(omega)
    cols = [ctx.fr_to_mont(canonical_limbs(c)) for c in list(fixed_values) + sigma_values]
    block = np.stack(cols) if cols else np.zeros((0, n, 4), dtype=np.uint64)
    return _keygen_from_columns(ctx, shape, k, g, g_lagrange, domain, block, transcript_repr, queries)


def keygen_from_synthesized(ctx, syn, g, g_lagrange, transcript_repr: int, queries=None) -> Keys:
    """keygen_vk + keygen_pk for a circuit the front-end synthesised (de_b200.frontend: fixed columns and copy constraints of
    one Circuit::synthesize pass, /root/reference/benches/delay_enc.rs:86,103).  The sigma columns come from the library's
    permutation assembly (de_assignment_sigma); everything is already in Montgomery form."""
    domain = EvaluationDomain(syn.shape.degree(), syn.k, ctx)
    block = np.concatenate([syn.fixed, syn.sigma(domain.omega)], axis=0)
    return _keygen_from_columns(ctx, syn.shape, syn.k, g, g_lagrange, domain, block, transcript_repr, queries)


def _keygen_from_columns(ctx, shape, k, g, g_lagrange, domain, block, transcript_repr, queries) -> Keys:
    """block: (n_fixed + n_perm_columns, n, 4) Montgomery lagrange values, fixed columns first"""
    import torch
    n = 1 << k
    params = ParamsKZG(k, g, g_lagrange, ctx)
    ncols = block.shape[0]
    d = torch.from_numpy(np.ascontiguousarray(block).view(np.int64)).cuda(ctx.device)
    commitments = params.commit_batch_canonical_dev(1, d, n, ncols) if ncols else np.zeros((0, 64), dtype=np.uint8)
    if ncols:
        domain.lagrange_to_coeff_dev(d, batch=ncols)
    ctx.sync()
    polys = d.cpu().numpy().view(np.uint64)
    nf = shape.n_fixed
    nsig = ncols - nf
    pk = ProvingKey(domain, shape, [polys[i] for i in range(nf)], [polys[nf + i] for i in range(nsig)])
    aq, fq, _ = queries if queries is not None else collect_queries(shape)
    prover = Prover(params, pk, aq, fq, transcript_repr)
    host = dict(k=k, g=g, g_lagrange=g_lagrange, shape=shape, fixed_polys=[polys[i] for i in range(nf)],
                sigma_polys=[polys[nf + i] for i in range(nsig)], advice_queries=aq, fixed_queries=fq,
                transcript_repr=transcript_repr)
    return Keys(params, domain, pk, prover, commitments[:nf], commitments[nf:], host)
