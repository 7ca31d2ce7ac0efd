"""The polynomial pipeline of halo2_proofs::plonk::create_proof for one proof, in the order create_proof issues it
(SURVEY.md section 3.2), driven through the C ABI.  This is the hot-path schedule the benches of the reference exercise
(/root/reference/benches/delay_enc.rs:123, mod_pow.rs:201, pose_enc.rs:127):

    commit_lagrange x A (advice)  ->  commit_lagrange x 2L (permuted input / table)  ->  commit_lagrange x (Z + L) (grand products)
    ->  commit x 1 (random poly)  ->  lagrange_to_coeff x (A + I + Z + 3L)  ->  evaluate_h (coeff_to_extended x (A + I + Z + 3L) + the
    fused row kernel)  ->  divide_by_vanishing_poly  ->  extended_to_coeff  ->  commit x (d - 1) (h pieces)  ->  commit x R (openings)

For the RSA / delay-encryption shape that is 31 MSMs, 23 + 23 + 1 transforms.  What create_proof does
BETWEEN these calls on the host (witness synthesis, transcript hashing, lookup sorting, grand products, evaluations at x,
Kate division) is outside SURVEY.md section 8's hot path; its outputs are inputs here.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import EvaluationDomain, ParamsKZG
from .plonk import ConstraintSystemShape, ProvingKey

N_OPENING_POINTS = 4  # upper bound; see Workload.n_openings


@dataclass
class Workload:
    """Column counts of one proof for a constraint-system shape."""
    shape: ConstraintSystemShape
    k: int

    @property
    def n(self):
        return 1 << self.k

    @property
    def n_lookups(self):
        return len(self.shape.lookups)

    @property
    def n_cols(self):  # polynomials that go lagrange -> coeff -> extended
        s = self.shape
        return s.n_advice + s.n_instance + s.n_perm_sets + 3 * self.n_lookups

    @property
    def n_openings(self):
        """distinct evaluation points of the GWC multi-open: x, omega x, omega^last x, plus omega^-1 x when lookups exist
        (31 = 27 + 4 commitments with RangeChip, 17 = 14 + 3 without)"""
        return 4 if self.n_lookups else 3

    @property
    def n_msm(self):
        s = self.shape
        return s.n_advice + 3 * self.n_lookups + s.n_perm_sets + 1 + (s.degree() - 1) + self.n_openings

    def offsets(self):
        """row offsets inside the column block: advice | instance | perm z | lookup z | lookup a' | lookup s'"""
        s, L = self.shape, self.n_lookups
        o = {}
        o["advice"] = 0
        o["instance"] = s.n_advice
        o["permz"] = o["instance"] + s.n_instance
        o["lookup_z"] = o["permz"] + s.n_perm_sets
        o["lookup_a"] = o["lookup_z"] + L
        o["lookup_s"] = o["lookup_a"] + L
        return o


class HotPathProver:
    """Holds the per-(ParamsKZG, ProvingKey) state resident in HBM and runs the per-proof schedule.

    Two CUDA streams per prover: the commitments run on the context's stream, the transforms and the quotient evaluator on
    a second context / stream.  lagrange_to_coeff and coeff_to_extended of a column do not depend on any transcript
    challenge, so they overlap the same rounds' MSMs; the row kernel of evaluate_h is issued only after the last
    pre-`y` commitment has been read back (create_proof's order), and the h-piece commitments wait for the quotient."""

    def __init__(self, ctx, workload: Workload, g, g_lagrange, fixed_coeff, sigma_coeff):
        import os
        import torch
        from . import Context
        self.ctx, self.w = ctx, workload
        s = workload.shape
        self.params = ParamsKZG(workload.k, g, g_lagrange, ctx)
        # DE_PROVER_OVERLAP=0 puts the transforms on the commitments' stream (A/B switch for measurements)
        self.overlap = os.environ.get("DE_PROVER_OVERLAP", "1") != "0"
        self.stream_b = torch.cuda.Stream() if self.overlap else torch.cuda.current_stream()
        self.ctx_b = Context(ctx.device)
        self.ctx_b.set_stream(self.stream_b.cuda_stream)
        with torch.cuda.stream(self.stream_b):
            self.domain = EvaluationDomain(s.degree(), workload.k, self.ctx_b)
            self.pk = ProvingKey(self.domain, s, fixed_coeff, sigma_coeff)
        self.stream_b.synchronize()
        self._work = None
        self._h = None

    # -- device-resident: `cols` (n_cols, n, 4) int64 CUDA tensor in lagrange form, `random_poly` (1, n, 4) and `openings`
    #    (N_OPENING_POINTS, n, 4) in coefficient form.  Returns the (n_msm, 12) commitments (host).
    def prove_dev(self, cols, random_poly, openings, challenges):
        import torch
        w, s, n = self.w, self.w.shape, self.w.n
        o, L = w.offsets(), w.n_lookups
        y, beta, gamma, theta = challenges
        out = []
        P = self.params
        stream_a = torch.cuda.current_stream()
        if self._work is None or self._work.shape != cols.shape:
            self._work = torch.empty_like(cols)
            self._h = torch.empty((self.domain.extended_n, 4), dtype=cols.dtype, device=cols.device)
        work = self._work
        # stream B: every column to coefficient form and onto the extended coset (challenge-independent)
        self.stream_b.wait_stream(stream_a)
        with torch.cuda.stream(self.stream_b):
            work.copy_(cols)
            self.domain.lagrange_to_coeff_dev(work, batch=w.n_cols)
            self.pk.extend_dev(work[o["advice"]:], work[o["instance"]:] if s.n_instance else None,
                               work[o["permz"]:] if s.n_perm_sets else None, work[o["lookup_z"]:] if L else None)
        # stream A: the commitment rounds, each read back before the next (the transcript needs them)
        out.append(P.commit_batch_dev(1, cols[o["advice"]:], n, s.n_advice))
        if L:
            out.append(P.commit_batch_dev(1, cols[o["lookup_a"]:], n, 2 * L))
        if s.n_perm_sets + L:
            out.append(P.commit_batch_dev(1, cols[o["permz"]:], n, s.n_perm_sets + L))
        out.append(P.commit_batch_dev(0, random_poly, n, 1))
        # y is known from here on: row kernel, division by the vanishing polynomial, back to coefficients
        with torch.cuda.stream(self.stream_b):
            self.pk.evaluate_h_rows_dev(y, beta, gamma, theta, self._h)
            self.domain.divide_by_vanishing_poly_dev(self._h)
            self.domain.extended_to_coeff_dev(self._h)
        stream_a.wait_stream(self.stream_b)
        out.append(P.commit_batch_dev(0, self._h, n, s.degree() - 1))
        out.append(P.commit_batch_dev(0, openings, n, w.n_openings))
        return np.concatenate(out, axis=0)

    # -- host inputs (pinned CPU tensors or numpy arrays of the same shapes): every step copies them to the device
    def prove_host(self, cols_h, random_h, openings_h, challenges, staging):
        """staging: dict of preallocated CUDA tensors {"cols", "random", "openings"}; copies are issued on the current stream"""
        staging["cols"].copy_(cols_h, non_blocking=True)
        staging["random"].copy_(random_h, non_blocking=True)
        staging["openings"].copy_(openings_h, non_blocking=True)
        return self.prove_dev(staging["cols"], staging["random"], staging["openings"], challenges)

    def close(self):
        self.pk.close()
        self.domain.close()
        self.params.close()
        self.ctx_b.close()

    @property
    def launches(self) -> int:
        return self.ctx.launches + self.ctx_b.launches
