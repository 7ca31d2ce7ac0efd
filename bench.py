#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 proving backend on the reference's delay_enc configuration.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is ONE create_proof per in-flight prover for the delay_enc configuration (BASELINE.json configs[2]: the
DelayEncryptCircuit at its bench k = 16, /root/reference/benches/delay_enc.rs:123-131,181): the whole of
halo2_proofs::plonk::create_proof (KZG, ProverGWC, Blake2b transcript) through de_create_proof[_dev] — 31 MSMs of 2^16,
23 iNTTs of 2^16, 23 coset NTTs + 1 iNTT of 2^18, the quotient evaluator over 2^18 rows, 10 lookup sorts, 7 grand products,
58 opening evaluations, 4 Kate divisions — producing the 2848-byte proof.  The circuit is a satisfied synthetic assignment
of the MainGate + RangeChip shape (de_b200/circuits.py, seed 0xDE03, 50 400 used rows): tests/test_gpu_prover.py checks
that exactly this prover's bytes equal the CPU restatement's and pass the restated verifier.

  value      proofs/s with the advice columns and the random draws already resident in HBM (device-timed, CUDA events)
  e2e        the same through the host-buffer entry point: pinned host advice columns + random draws are copied to the
             device inside the timed region; commitments / evaluations are read back as the transcript needs them
  roofline   dominant kernel (k_msm_accumulate), integer-pipe bound: field multiplications actually executed (10 per bucket
             addition, additions counted on the device) / measured launch time, against the Montgomery-multiply peak measured
             IN THIS RUN (de_int_peak); `hbm` inside it keeps the algorithmic-bytes fraction of the measured HBM copy bandwidth
  msm        whole commitments (not one kernel) at 2^16 (alone / batch of 8), 2^20, 2^24 with the CPU restatement's
             best_multiexp on the same inputs beside them; msm_multi_gpu (N > 1): one commitment with the base range sharded
             over the ranks, partial points exchanged over NCCL, against the same commitment on one GPU
  other_configs  short runs of the other two bench circuits (pose_enc k = 11, mod_pow k = 17)
  cpu_baseline   the whole create_proof by the restated reference algorithms (oracle/: best_multiexp / best_fft / evaluate_h /
             lookup permutation / grand products / openings as halo2_proofs v2023_04_20 implements them, C + pthreads on all
             host cores, Python sequencing + transcript), one proof, byte-compared with the GPU's

--impl reference times that CPU restatement alone (the Rust prover cannot be built: no cargo/rustc, un-vendored crates).
N > 1 (torchrun): independent proofs are sharded one stream of proofs per GPU, no collective on the data path (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "delay-encryption-in-halo2_b200"))

# BASELINE.json configs: [2] delay_enc is the headline (the metric's name); [1] pose_enc and [0] mod_pow are selectable with
# --config for the parity / context runs.  (k, range lookups, used rows, seed, reference bench file:line)
CONFIGS = {
    "delay_enc": (16, True, 50400, 0xDE03, "benches/delay_enc.rs:181"),
    "mod_pow": (17, True, 41766, 0xDE01, "benches/mod_pow.rs:258"),
    "pose_enc": (11, False, 1450, 0xDE02, "benches/pose_enc.rs:184"),
}
K = 16
USED_ROWS = 50400
SEED = 0xDE03
WITH_LOOKUPS = True
CONFIG_NAME = "delay_enc"
DEFAULT_WORKLOAD = WORKLOAD = ("delay_enc k=16 create_proof of the reference's DelayEncryptCircuit (RSA-2048 x^e mod n with a 5-bit e -> Poseidon hash -> "
            "Poseidon encryption of the 2-word zero message; witness by the C++ front-end, 44658 used rows; MainGate + RangeChip shape: 5 "
            "advice, 15 fixed, 5 lookups, 6 permutation columns): 31 MSM 2^16 (KZG bases resident), 23 iNTT 2^16, 23 coset-NTT 2^18, "
            "evaluate_h over 2^18 rows, 1 iNTT 2^18, 10 lookup sorts, 7 grand products, 58 evaluations, 4 Kate divisions, "
            "Blake2b transcript -> 2848-byte proof")
DEFAULT_METRIC = METRIC = "delay_enc_create_proof_proofs_per_s"
TRANSCRIPT_REPR = 0xDE1A7E9C0DE
UNIT = "proofs/s"
LATENCY_WITNESS_THREADS = max(1, min(12, (os.cpu_count() or 1) - 2))
MUL_PEAK_FALLBACK_GMULS = 65.9  # only if de_int_peak fails: this pool's B200 by tools/int_peak (profiles/r01_int_peak.jsonl)
MULS_PER_ADD = 10       # XYZZ mixed addition: 8M + 2S (the convention of round 1; add_rooflines weighs the 2 squarings by their measured cost)
# DRAM bytes of a k_msm_accumulate launch per executed bucket addition, from ONE `ncu --set full` capture (dram__bytes_read.sum +
# dram__bytes_write.sum of the launch / its bucket additions); the file names the capture it came from.  Absent -> traffic: null.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "msm_accumulate_traffic.json")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


WITNESS = "real"   # --witness: "real" = the reference's circuit through the C++ front-end; "synthetic" = round 1's stand-in


class Circuit:
    """The proved circuit of the current configuration.  "real": Circuit::synthesize of the reference's bench circuit by the C++
    front-end (de_b200.frontend; DelayEncryptCircuit src/lib.rs:103-318, RSACircuit benches/mod_pow.rs:36-140, PoseidonEncCircuit
    src/encryption/chip.rs:114-198) on inputs drawn as the benches draw them, from a seeded generator.  "synthetic": a satisfied
    random assignment of the same constraint-system shape (de_b200/circuits.py)."""

    def __init__(self):
        from de_b200 import circuits, frontend as fe
        self.kind = WITNESS
        if WITNESS == "synthetic":
            self.asg = circuits.satisfied_assignment(WITH_LOOKUPS, K, SEED, USED_ROWS, uniform_values=not WITH_LOOKUPS)
            self.shape, self.instances, self.used_rows, self.syn = self.asg.shape, self.asg.instances, self.asg.used_rows, None
            return
        kind = {"delay_enc": fe.DELAY_ENC, "mod_pow": fe.MOD_POW, "pose_enc": fe.POSE_ENC}[CONFIG_NAME]
        n, e, x = fe.sample_rsa_inputs(SEED) if kind != fe.POSE_ENC else (0, 0, 0)
        self.inputs = dict(kind=kind, k=K, n=n, e=e, x=x, message=() if kind == fe.MOD_POW else (0,) * fe.MESSAGE_CAPACITY,
                           key=(SEED * 0x9E3779B97F4A7C15 % fe.plonk.FR, (SEED + 1) * 0xC2B2AE3D27D4EB4F % fe.plonk.FR))
        self.syn = fe.synthesize(**self.inputs)
        self.shape, self.instances, self.used_rows = self.syn.shape, self.syn.instances, self.syn.used_rows

    def keygen(self, ctx, g, g_lagrange):
        from de_b200 import keygen
        if self.syn is None:
            return keygen.keygen(ctx, self.shape, K, g, g_lagrange, self.asg.fixed, self.asg.copies, TRANSCRIPT_REPR)
        return keygen.keygen_from_synthesized(ctx, self.syn, g, g_lagrange, TRANSCRIPT_REPR)

    def advice_mont(self, ctx):
        import numpy as np
        from de_b200 import keygen
        if self.syn is None:
            return np.stack([ctx.fr_to_mont(keygen.canonical_limbs(c)) for c in self.asg.advice])
        return self.syn.advice

    def witness_pass(self, threads=1):
        """a fresh de_circuit_witness runner (None for the synthetic stand-in, which has no front-end)"""
        from de_b200 import frontend as fe
        return fe.WitnessPass(threads=threads, **self.inputs) if self.syn is not None else None

    def cpu_view(self):
        """(fixed columns as integer lists, copy constraints as tuples, advice columns in Montgomery form) for the CPU arm's own
        keygen; conversions use the checker library: only the CPU legs call this"""
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import numpy as np
        import orc
        if self.syn is None:
            return self.asg.fixed, self.asg.copies, [orc.fr_mont_from_ints(c) for c in self.asg.advice]
        fixed = [orc.fr_ints_from_mont(np.ascontiguousarray(c)) for c in self.syn.fixed]
        return fixed, [tuple(int(v) for v in c) for c in self.syn.copies], [np.ascontiguousarray(c) for c in self.syn.advice]

    def witness_stats(self, ctx):
        """value distribution of the advice cells of the used rows (SURVEY.md 8d guessed 45 % < 2^8, 35 % < 2^64, 10 % < 2^134,
        10 % wider): what the MSM's zero-digit skipping sees"""
        import numpy as np
        adv = self.advice_mont(ctx)[:, :self.used_rows]
        can = np.stack([ctx.fr_from_mont(np.ascontiguousarray(c)) for c in adv]).reshape(-1, 4)
        hi3, hi2, hi1 = can[:, 3] != 0, can[:, 2] != 0, can[:, 1] != 0
        zero = ~(hi3 | hi2 | hi1) & (can[:, 0] == 0)
        lt8 = ~(hi3 | hi2 | hi1) & (can[:, 0] < 256) & ~zero
        lt64 = ~(hi3 | hi2 | hi1) & ~lt8 & ~zero
        lt134 = ~hi3 & (can[:, 2] < 64) & (hi2 | hi1)
        wide = ~(zero | lt8 | lt64 | lt134)
        t = float(can.shape[0])
        return {"cells": int(t), "zero": float(zero.sum() / t), "lt_2^8": float(lt8.sum() / t), "lt_2^64": float(lt64.sum() / t),
                "lt_2^134": float(lt134.sum() / t), "wider": float(wide.sum() / t),
                "survey_guess": {"lt_2^8": 0.45, "lt_2^64": 0.35, "lt_2^134": 0.10, "wider": 0.10}}


def build_circuit():
    """The circuit of the current configuration and the SRS bases."""
    from de_b200 import synth
    n = 1 << K
    return Circuit(), synth.gen_bases(n, start=0), synth.gen_bases(n, start=n)


def random_draws(count):
    from de_b200 import synth
    return synth.uniform_fr(SEED + 777, count)


def build_cpu_inputs(circ, advice_mont, g, g_lagrange):
    """Columns for the CPU arm's hot-path schedule: the assignment's advice columns; every other column (permuted lookup
    columns, grand products, random / opening polynomials, fixed and sigma polynomials) is a uniform column of the same
    size — the reference algorithms' cost does not depend on their values."""
    import numpy as np
    from de_b200 import prover, synth
    shape = circ.shape
    w = prover.Workload(shape, K)
    n, o = w.n, w.offsets()
    cols = np.empty((w.n_cols, n, 4), dtype=np.uint64)
    for i in range(shape.n_advice):
        cols[o["advice"] + i] = advice_mont[i]
    cols[o["instance"]] = 0
    for i in range(o["permz"], w.n_cols):
        cols[i] = synth.uniform_fr(SEED + 100 + i, n)
    return dict(shape=shape, w=w, cols=cols, random=synth.uniform_fr(SEED + 200, n).reshape(1, n, 4),
                openings=synth.uniform_fr(SEED + 201, n * w.n_openings).reshape(w.n_openings, n, 4),
                fixed=[synth.uniform_fr(SEED + 300 + i, n) for i in range(shape.n_fixed)],
                sigma=[synth.uniform_fr(SEED + 400 + i, n) for i in range(len(shape.perm_columns))], g=g, g_lagrange=g_lagrange,
                challenges=(0x1D2C3B4A59687766, 0x0F1E2D3C4B5A6978, 0x1122334455667788, 0x99AABBCCDDEEFF00))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx, self.proc, self.lines = device_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_setup(inp, state):
    """CPU 'ProvingKey' (fixed / sigma / l0 / l_last / l_active cosets), as keygen_pk builds it once in the reference."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    from de_b200 import plonk
    shape = inp["shape"]
    state["dom"] = orc.Domain(shape.degree(), K)
    desc, keep = plonk.marshal_pk_desc(shape, inp["fixed"], inp["sigma"])
    state["pk"] = orc.Pk(state["dom"], desc, keep)
    state["ch"], state["chkeep"] = plonk.marshal_challenges(*inp["challenges"])


def cpu_reference_step(inp, state):
    """One proof's hot path on the host cores with the restated reference algorithms (the only place the oracle is executed
    by bench.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import orc
    from de_b200 import prover
    shape, w = inp["shape"], inp["w"]
    n, o, L = w.n, w.offsets(), w.n_lookups
    if "pk" not in state:
        cpu_reference_setup(inp, state)
    dom, pk = state["dom"], state["pk"]
    cols = inp["cols"]
    pts = []
    for i in range(shape.n_advice):
        pts.append(orc.best_multiexp(cols[o["advice"] + i], inp["g_lagrange"]))
    for i in range(2 * L):
        pts.append(orc.best_multiexp(cols[o["lookup_a"] + i], inp["g_lagrange"]))
    for i in range(shape.n_perm_sets + L):
        pts.append(orc.best_multiexp(cols[o["permz"] + i], inp["g_lagrange"]))
    pts.append(orc.best_multiexp(inp["random"][0], inp["g"]))
    coeff = [dom.lagrange_to_coeff(cols[i]) for i in range(w.n_cols)]
    h = pk.evaluate_h(coeff[o["advice"]:o["advice"] + shape.n_advice], coeff[o["instance"]:o["instance"] + shape.n_instance],
                      state["ch"], coeff[o["permz"]:o["permz"] + shape.n_perm_sets], coeff[o["lookup_z"]:])
    h = dom.divide_by_vanishing(h)
    hc = dom.extended_to_coeff(h)
    for i in range(shape.degree() - 1):
        pts.append(orc.best_multiexp(np.ascontiguousarray(hc[i * n:(i + 1) * n]), inp["g"]))
    for i in range(w.n_openings):
        pts.append(orc.best_multiexp(inp["openings"][i], inp["g"]))
    return np.stack(pts)


def cpu_prover_setup(circ, g, g_lagrange):
    """keys of the restated CPU prover for the same circuit / SRS / transcript_repr (keygen: one-time in the reference too)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pyprover as pp
    from de_b200 import plonk
    n = 1 << K
    oparams = pp.Params(K, n, None, None, None, None, np.ascontiguousarray(g), np.ascontiguousarray(g_lagrange))
    fixed, copies, _ = circ.cpu_view()
    opk = pp.keygen(oparams, circ.shape, pp.Queries(*plonk.collect_queries(circ.shape)), fixed, copies, TRANSCRIPT_REPR)
    pp._pk_arrays(oparams, opk)
    return pp, oparams, opk


DEFAULT_CPU_SAMPLE = CPU_SAMPLE = ("one whole delay_enc k=16 create_proof per step (Circuit::synthesize by this repository's C++ front-end, then the prover), the same "
              "circuit, keys and random draws as the GPU arm: the restated "
              "reference algorithms (oracle/: best_multiexp, best_fft, evaluate_h, lookup permutation, grand products, eval_polynomial, "
              "kate_division in C with pthreads on all host cores; Python only sequences the calls and hashes the transcript), not the "
              "Rust binary; proof bytes equal the GPU's")
CPU_HOT_SAMPLE = ("the same proof's 31 best_multiexp 2^16, 23 + 23 + 1 best_fft and evaluate_h over 2^18 rows alone (SURVEY.md section 8a "
                  "rows a3-a9), leaving out the reference's host-side work between them")


def run_reference(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import orc
    orc.build()
    circ, g, g_lagrange = build_circuit()
    _, _, advice_mont = circ.cpu_view()
    pp, oparams, opk = cpu_prover_setup(circ, g, g_lagrange)
    randoms = random_draws(pp.random_count(circ.shape, 1 << K))
    wp = circ.witness_pass()
    buf = np.empty((circ.shape.n_advice, 1 << K, 4), dtype=np.uint64)
    proof = None
    synth_ms = []

    def step():
        # the reference's create_proof starts with Circuit::synthesize (benches/delay_enc.rs:123-131): the witness pass is part of
        # the step.  The Rust synthesize cannot run here; this repository's C++ front-end stands in for it on BOTH arms.
        cols = advice_mont
        if wp is not None:
            synth_ms.append(wp.run(buf))
            cols = [buf[i] for i in range(buf.shape[0])]
        return pp.create_proof_fast(oparams, opk, cols, circ.instances, randoms)

    for _ in range(max(args.warmup, 0)):
        proof = step()
    del synth_ms[:]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof = step()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    cores = orc.ncpu()
    import hashlib
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": K, "seed": hex(SEED), "proof_bytes": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": CPU_SAMPLE},
        "synthesis_ms": (sum(synth_ms) / len(synth_ms)) if synth_ms else None,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def measure_int_peak(ctx):
    """Fr Montgomery multiplications / s of this GPU, measured now (de_int_peak): the denominator of every int-pipe fraction"""
    try:
        return ctx.int_peak(), "measured_in_run (de_int_peak: dependent Fr Montgomery product chains, 16 warps/SM x 2 chains, best of 4)"
    except Exception as e:  # noqa: BLE001 - the line must still be printed
        return MUL_PEAK_FALLBACK_GMULS, f"fallback (de_int_peak failed: {e}); tools/int_peak on this pool's B200, profiles/r01_int_peak.jsonl"


def proof_bench(args, local_rank, world, steps, warmup, B, detail):
    """The proof arms for the CURRENT configuration globals (K, WITH_LOOKUPS, ...): throughput (value), end-to-end (host buffers),
    latency (one proof in flight; with `detail` also the per-kernel CUDA-event timings).  Returns (line dict, artefacts dict)."""
    import hashlib
    import numpy as np
    import torch
    import de_b200
    from de_b200 import keygen, sharding
    import torch.distributed as dist

    circ, g, g_lagrange = build_circuit()
    shape = circ.shape
    n = 1 << K
    main_stream = torch.cuda.current_stream()
    as_i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64))

    class Worker:
        """one in-flight proof: its own context, stream, ParamsKZG / ProvingKey / prover buffers"""

        def __init__(self, first=None):
            self.stream = torch.cuda.Stream()
            self.ctx = de_b200.Context(local_rank)
            self.ctx.set_stream(self.stream.cuda_stream)
            with torch.cuda.stream(self.stream):
                if first is None:
                    self.keys = circ.keygen(self.ctx, g, g_lagrange)
                else:
                    self.keys = first.keys.clone_on(self.ctx)
            self.proof = None
            # end-to-end arm: this worker's own witness pass and TWO pinned staging buffers for the advice columns, so that
            # Circuit::synthesize of the next proof (a host thread; the C call releases the GIL) runs while the GPU proves the
            # current one - what a prover service with a queue of statements does
            self.wp = circ.witness_pass()
            # the single-statement arm lets the pass use idle host cores (the RSA region's mul_mod row ranges are independent);
            # the throughput arm keeps one thread per pass, its cores belong to the other provers in flight
            self.wp_single = circ.witness_pass(threads=LATENCY_WITNESS_THREADS) if first is None else None
            self.synth_single_ms = []
            self.staged = [False, False]
            self.stage = [torch.empty((shape.n_advice, n, 4), dtype=torch.int64).pin_memory() for _ in range(2)] if self.wp is not None else None
            self.synth_ms = []

        def _synth(self, slot, wp=None, log=None):
            # a staging buffer that already holds a pass of this circuit is not zeroed again (de_circuit_desc.reuse_buffer)
            (self.synth_ms if log is None else log).append((wp or self.wp).run(self.stage[slot], reuse=self.staged[slot]))
            self.staged[slot] = True

        def run(self, nsteps, host, serial=False):
            with torch.cuda.stream(self.stream):
                if host and self.wp is not None and serial:
                    # one statement at a time: synthesize, then prove (the latency a single caller of create_proof sees)
                    for _ in range(nsteps):
                        self._synth(0, self.wp_single, self.synth_single_ms)
                        self.proof = self.keys.prover.create_proof([self.stage[0][c] for c in range(shape.n_advice)], [], randoms_h)
                    return
                if host and self.wp is not None:
                    # every proof gets its own witness pass: step i's runs during step i - 1's create_proof (the first one up front)
                    self._synth(0)
                    for i in range(nsteps):
                        nxt = None
                        if i + 1 < nsteps:
                            nxt = threading.Thread(target=self._synth, args=((i + 1) & 1,))
                            nxt.start()
                        buf = self.stage[i & 1]
                        self.proof = self.keys.prover.create_proof([buf[c] for c in range(shape.n_advice)], [], randoms_h)
                        if nxt is not None:
                            nxt.join()
                    return
                for _ in range(nsteps):
                    if host:
                        self.proof = self.keys.prover.create_proof([advice_h[i] for i in range(shape.n_advice)], [], randoms_h)
                    else:
                        self.proof = self.keys.prover.create_proof_dev(advice_d, randoms_d)

    workers = [Worker()]
    ctx = workers[0].ctx
    prover0 = workers[0].keys.prover
    advice_mont = np.ascontiguousarray(circ.advice_mont(ctx))
    randoms = random_draws(prover0.random_count)
    advice_h, randoms_h = as_i64(advice_mont).pin_memory(), as_i64(randoms).pin_memory()
    advice_d, randoms_d = advice_h.cuda(), randoms_h.cuda()
    workers += [Worker(workers[0]) for _ in range(B - 1)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(active, nsteps, host, serial=False):
        """nsteps on every active worker concurrently; device time between two events on the main stream that fence all
        worker streams; max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main_stream)
        for wk in active:
            wk.stream.wait_event(e0)
        threads = [threading.Thread(target=wk.run, args=(nsteps, host, serial)) for wk in active]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for wk in active:
            main_stream.wait_stream(wk.stream)
        e1.record(main_stream)
        barrier()
        return sharding.max_over_ranks(e0.elapsed_time(e1))

    throughput_mode = (B > 1) if args.mode == "auto" else args.mode == "throughput"
    for wk in workers:
        wk.ctx.set_mode(throughput=throughput_mode)  # several provers share the GPU
    timed(workers, warmup, False)
    first_proof = workers[0].proof
    assert len(first_proof) == prover0.proof_size == (2848 if WITH_LOOKUPS else 1792)
    # ---- throughput arm (value): B proofs in flight per GPU, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sum(wk.ctx.launches for wk in workers)
    ms_dev = timed(workers, steps, False)
    launches = sum(wk.ctx.launches for wk in workers) - launches0
    clocks = sampler.stop()
    for wk in workers:
        assert wk.proof == first_proof, "proof bytes differ between workers / steps"
    # ---- end-to-end arm: pinned host advice + random draws in, proof bytes out, every proof
    timed(workers, 2, True)
    for wk in workers:
        del wk.synth_ms[:]
    ms_e2e = timed(workers, steps, True)
    for wk in workers:
        assert wk.proof == first_proof, "host-buffer path disagrees with the device-resident path"
    synth_all = [v for wk in workers for v in wk.synth_ms]
    # ---- latency arm: ONE proof in flight; per-kernel CUDA-event timing is taken here (no overlapping streams)
    ctx.set_mode(throughput=False)
    timed(workers[:1], 2, False)
    lat_steps = max(5, min(steps, 20))
    ms_lat = timed(workers[:1], lat_steps, False)  # no per-kernel events in here: their creation / recording costs ~0.1 ms per proof
    assert workers[0].proof == first_proof, "latency-mode proof differs"
    ms_lat_timed = ms_lat
    if detail:
        # the same arm once more with a CUDA-event pair around every timed launch: the per-kernel figures of `roofline`
        ctx.timing_reset()
        ctx.timing_enable(True)
        ms_lat_timed = timed(workers[:1], lat_steps, False)
        ctx.timing_enable(False)
    ms_lat_e2e = None
    if workers[0].wp is not None:
        timed(workers[:1], 2, True, serial=True)
        ms_lat_e2e = timed(workers[:1], lat_steps, True, serial=True)

    h2d = advice_h.numel() * 8 + prover0.random_count * 32
    n_evals = 58 if WITH_LOOKUPS else 39
    n_points = (len(first_proof) - 32 * n_evals) // 32
    d2h = n_points * 64 + n_evals * 32
    value = world * steps * B / (ms_dev / 1000.0)
    e2e_value = world * steps * B / (ms_e2e / 1000.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (8x32-bit Montgomery limbs, IMAD.WIDE carry chains)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": K, "seed": hex(SEED), "proofs_per_step_per_gpu": B, "proof_bytes": len(first_proof),
                   "proof_sha256": hashlib.sha256(first_proof).hexdigest(),
                   "mode": "DE_MODE_THROUGHPUT" if throughput_mode else "DE_MODE_LATENCY",
                   "in_flight": f"{B} independent proofs per GPU, one host thread + CUDA stream each (BASELINE config 5: 64 proofs over "
                                "8 GPUs = 8 per GPU)",
                   "l2": f"per-step working set ~{0.8 * B * 2.0 ** (K - 16):.2f} GB (columns, cosets, pk cosets, base tables) > 126 MB L2; no explicit flush",
                   "sharding": "independent proofs across GPUs, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * B, "d2h_bytes_per_step": d2h * B,
                "ms_per_step": ms_e2e / steps,
                "includes": ("Circuit::synthesize for EVERY proof (witness pass of the C++ front-end on a host thread per in-flight prover, "
                             "double-buffered: proof i + 1 is synthesised while proof i is on the GPU) + H2D of the advice columns and random "
                             "draws + create_proof + D2H of commitments / evaluations") if synth_all else
                            "H2D of the advice columns and random draws + create_proof + D2H of commitments / evaluations",
                "synthesis_ms": (sum(synth_all) / len(synth_all)) if synth_all else None},
        "synthesis_ms": (sum(synth_all) / len(synth_all)) if synth_all else None,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "latency": {"create_proof_s": ms_lat / lat_steps / 1000.0, "proofs_in_flight": 1, "steps": lat_steps, "mode": "DE_MODE_LATENCY",
                    "create_proof_s_with_kernel_events": ms_lat_timed / lat_steps / 1000.0,
                    "e2e_create_proof_s": (ms_lat_e2e / lat_steps / 1000.0) if ms_lat_e2e else None,
                    "e2e_witness_threads": LATENCY_WITNESS_THREADS if ms_lat_e2e else None,
                    "e2e_synthesis_ms": (sum(workers[0].synth_single_ms[-lat_steps:]) / lat_steps) if ms_lat_e2e else None,
                    "note": "create_proof_s: witness resident in HBM; e2e_create_proof_s: Circuit::synthesize on the host, then H2D and the "
                            "proof, strictly one after the other (what one caller of the reference's create_proof waits for); that arm's "
                            "witness pass emits the independent row ranges of the circuit (the mul_mod calls of pow_mod, x < n, the Poseidon "
                            "regions) on e2e_witness_threads host threads"},
    }
    line["config"]["witness"] = circ.kind
    line["config"]["used_rows"] = int(circ.used_rows)
    if detail:
        line["witness_stats"] = circ.witness_stats(ctx)
    art = dict(circ=circ, g=g, g_lagrange=g_lagrange, advice_mont=advice_mont, randoms=randoms, first_proof=first_proof, workers=workers,
               ctx=ctx, advice_d=advice_d, ms_lat=ms_lat_timed, lat_steps=lat_steps, shape=shape)
    return line, art


def add_rooflines(line, art, rank):
    """roofline (dominant kernel k_msm_accumulate, integer-pipe bound: executed field multiplications against the multiply peak
    measured in this run), kernel shares, and the NTT GB/s block"""
    import torch
    ctx, ms_ref, lat_steps, shape = art["ctx"], art["ms_lat"], art["lat_steps"], art["shape"]
    acc_ms, acc_pts, acc_n = ctx.timing_get("k_msm_accumulate")
    ntt_ms, ntt_el, ntt_n = ctx.timing_get("k_ntt_pass")
    ev_ms, ev_rows, ev_n = ctx.timing_get("k_eval_h")
    red_ms, _, red_n = ctx.timing_get("k_msm_digit_sums")
    _, bucket_adds, _ = ctx.timing_get("msm_bucket_adds")
    ctx.timing_enable(False)
    peaks, peak_kind = _peaks()
    mul_peak, mul_peak_src = measure_int_peak(ctx)
    line["int_peak_gmul_s"] = mul_peak
    try:
        sqr_peak = ctx.int_peak_sqr()
    except Exception:  # noqa: BLE001
        sqr_peak = mul_peak
    line["int_peak_gsqr_s"] = sqr_peak
    # one mixed bucket addition in multiplication-equivalents: 8 products + 2 squarings at their measured relative cost
    muls_per_add = 8.0 + 2.0 * mul_peak / sqr_peak
    if acc_n:
        pts_per_launch = acc_pts / acc_n
        adds_per_launch = bucket_adds / acc_n
        avg_ms = acc_ms / acc_n
        gmuls = muls_per_add * bucket_adds / (acc_ms * 1e-3) / 1e9
        hbm_gbs = 96.0 * pts_per_launch / (avg_ms * 1e-3) / 1e9
        traffic, traffic_src = None, "no ncu capture on file"
        try:
            with open(TRAFFIC_FILE) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_bucket_add"] * adds_per_launch
            traffic_src = tj["source"]
        except Exception:
            pass
        line["roofline"] = {
            "kernel": "k_msm_accumulate", "bound": "int_pipe", "achieved": gmuls, "peak": mul_peak, "unit": "Gmul/s (Fr/Fq Montgomery multiplications)",
            "frac": gmuls / mul_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": mul_peak_src,
            "work": {"bucket_adds_per_launch": adds_per_launch, "muls_per_bucket_add": muls_per_add, "points_per_launch": pts_per_launch,
                     "muls_per_bucket_add_note": "8 products + 2 squarings x (int_peak_gmul_s / int_peak_gsqr_s): the dedicated squaring "
                                                 "is cheaper than a product, and counting it as one would overstate the fraction",
                     "bucket_adds_per_proof": bucket_adds / lat_steps,
                     "note": "achieved = muls_per_bucket_add x bucket additions actually executed (counted on the device: non-zero signed 16-bit digits) / "
                             "kernel time; SURVEY.md 8d's fixed 160 mul/point convention would also count the zero digits of the "
                             "witness columns, which are legitimately skipped"},
            "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"], "peak_source": peak_kind,
                    "note": "algorithmic 96 B/point (32 B scalar + 64 B base) / kernel time: the kernel is nowhere near the memory roof"},
            "measured_in": "latency arm run once more with CUDA events around every timed launch on its stream (one proof in flight; "
                           "share_of_step is against that pass, latency.create_proof_s_with_kernel_events)",
            "share_of_step": acc_ms / ms_ref, "launches": acc_n, "avg_launch_ms": avg_ms}
        line["msm_fill_gpts_s"] = acc_pts / (acc_ms * 1e-3) / 1e9
    if ntt_n:
        line["kernel_share"] = {"k_msm_accumulate": acc_ms / ms_ref, "k_msm_digit_sums": red_ms / ms_ref,
                                "k_ntt_pass": ntt_ms / ms_ref, "k_eval_h": ev_ms / ms_ref,
                                "note": "CUDA-event time of each kernel family / single-proof latency; the transforms run on the prover's second "
                                        "stream concurrently with the bucket kernels, so the shares overlap and do not add up"}
    # NTT GB/s on its own (inside the prover the transforms overlap the commitments): the proof's batch of coset transforms,
    # coeff_to_extended of all per-proof columns, 64 B of algorithmic traffic per output element
    if rank == 0:
        wk0 = art["workers"][0]
        kz = wk0.keys
        n = 1 << K
        n_cols = shape.n_advice + shape.n_instance + shape.n_perm_sets + 3 * len(shape.lookups)
        ext_n = kz.domain.extended_n
        with torch.cuda.stream(wk0.stream):
            src = torch.zeros((n_cols, n, 4), dtype=torch.int64, device="cuda")
            src[:shape.n_advice] = art["advice_d"]
            dst = torch.empty((n_cols, ext_n, 4), dtype=torch.int64, device="cuda")
            for _ in range(3):
                kz.domain.coeff_to_extended_dev(src, dst, batch=n_cols)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(wk0.stream)
            for _ in range(10):
                kz.domain.coeff_to_extended_dev(src, dst, batch=n_cols)
            e1.record(wk0.stream)
            e1.synchronize()
            ms_ntt = e0.elapsed_time(e1) / 10
        gbs = 64.0 * ext_n * n_cols / (ms_ntt * 1e-3) / 1e9
        ek = ext_n.bit_length() - 1
        line["ntt_gb_s"] = gbs
        line["ntt"] = {"what": f"coeff_to_extended of {n_cols} columns 2^{K} -> 2^{ek} alone on the GPU", "ms": ms_ntt, "gb_s": gbs,
                       "frac_hbm": gbs / peaks["hbm_gbs"], "gmul_s": n_cols * ext_n / 2 * ek / (ms_ntt * 1e-3) / 1e9,
                       "frac_int_pipe": n_cols * ext_n / 2 * ek / (ms_ntt * 1e-3) / 1e9 / mul_peak,
                       "int_pipe_peak_gmul_s": mul_peak}
        del src, dst
    return mul_peak


def close_workers(art):
    for wk in art["workers"]:
        wk.keys.close()
        wk.ctx.close()
    art["workers"] = []


def set_config(name, k=0):
    global K, USED_ROWS, SEED, WITH_LOOKUPS, CONFIG_NAME, WORKLOAD, METRIC, CPU_SAMPLE
    K, WITH_LOOKUPS, USED_ROWS, SEED, where = CONFIGS[name]
    K = k or K
    CONFIG_NAME = name
    if name == "delay_enc" and not k and WITNESS == "real":
        WORKLOAD, METRIC, CPU_SAMPLE = DEFAULT_WORKLOAD, DEFAULT_METRIC, DEFAULT_CPU_SAMPLE
        return
    METRIC = f"{name}_create_proof_proofs_per_s" if not k else f"{name}_k{K}_create_proof_proofs_per_s"
    WORKLOAD = (f"{name} k={K} create_proof (/root/reference/{where}; MainGate{' + RangeChip' if WITH_LOOKUPS else ''} shape, "
                + ("witness of the reference's circuit by the C++ front-end)" if WITNESS == "real" else f"satisfied synthetic witness, {USED_ROWS} used rows)"))
    CPU_SAMPLE = DEFAULT_CPU_SAMPLE.replace("delay_enc k=16", f"{name} k={K}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ntt-multi-gpu", action="store_true", help="skip the one-vector-over-all-ranks NTT line of a multi-GPU run")
    ap.add_argument("--no-msm", action="store_true", help="skip the whole-commitment MSM block (1 GPU) / msm_multi_gpu (N GPUs)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short pose_enc / mod_pow runs of the 1-GPU line")
    ap.add_argument("--inflight", type=int, default=8, help="independent proofs in flight per GPU (one prover + stream each)")
    ap.add_argument("--config", default="delay_enc", choices=sorted(CONFIGS), help="circuit shape / size (default: the headline delay_enc)")
    ap.add_argument("--mode", default="auto", choices=["auto", "latency", "throughput"],
                    help="de_ctx_set_mode of the in-flight arms (auto: throughput when more than one proof is in flight)")
    ap.add_argument("--k", type=int, default=0, help="override the config's k (the reference's README also times k = 15 ... 19)")
    ap.add_argument("--witness", default="real", choices=["real", "synthetic"],
                    help="real: the reference's circuit synthesised by the C++ front-end (default); synthetic: round 1's satisfied random assignment")
    args = ap.parse_args()
    global WITNESS
    WITNESS = args.witness
    headline = args.config == "delay_enc" and not args.k
    set_config(args.config, args.k)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing hygiene: at least three untimed steps

    import numpy as np
    import torch
    import torch.distributed as dist
    import de_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 backend has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = max(1, args.inflight)
    line, art = proof_bench(args, local_rank, world, args.steps, args.warmup, B, detail=True)
    mul_peak = add_rooflines(line, art, rank)
    circ, g, g_lagrange, advice_mont, randoms = art["circ"], art["g"], art["g_lagrange"], art["advice_mont"], art["randoms"]
    first_proof, shape = art["first_proof"], art["shape"]
    close_workers(art)
    del art
    torch.cuda.empty_cache()
    sys.path.insert(0, os.path.join(ROOT, "tools"))

    if world > 1 and not args.no_ntt_multi_gpu:
        # BASELINE's "NTT GB/s at 1/2/4/8 B200": ONE 2^24 vector over all ranks (four-step transform, both exchanges as peer-memory
        # stores from inside the kernels), checked on rank 0 against the single-GPU transform of the same vector
        import ntt_dist_sweep
        nctx = de_b200.Context(local_rank)
        try:
            res = ntt_dist_sweep.measure_dist(nctx, 24, 5, emit=False)
        except (RuntimeError, ValueError) as e:  # raised on every rank together (layout check; DistNtt's collective setup)
            res = {"error": str(e)}
        nctx.close()
        if rank == 0:
            line["ntt_multi_gpu"] = res
    msm_single = None
    if not args.no_msm and headline:
        import msm_bench
        mstream = torch.cuda.Stream()
        mctx = de_b200.Context(local_rank)
        mctx.set_stream(mstream.cuda_stream)
        if world > 1:
            # north_star's first split: "MSM base ranges are sharded per commitment ... combined over NVLink" 
            res = []
            for log_n in [24] + ([26] if world >= 4 else []):
                try:
                    r = msm_bench.measure_sharded(mctx, mstream, rank, world, log_n, reps=5)
                except (RuntimeError, ValueError) as e:
                    r = {"op": "msm_multi_gpu", "log_n": log_n, "error": str(e)}
                res.append(r)
            if rank == 0:
                line["msm_multi_gpu"] = res
        else:
            try:
                msm_single = msm_bench.measure_single(mctx, mstream, (16, 20, 24), keep_host=() if args.no_cpu_baseline else (16, 20))
            except (RuntimeError, ValueError) as e:
                msm_single = [{"error": str(e)}]
        mctx.close()
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc
        n = 1 << K
        inp = build_cpu_inputs(circ, list(advice_mont), g, g_lagrange)
        # the prover's blinding rows, so that the CPU's advice commitments can be compared with the proof's
        usable = n - (shape.blinding_factors + 1)
        for i in range(shape.n_advice):
            inp["cols"][i, usable:] = randoms[i * (n - usable):(i + 1) * (n - usable)]
        state = {}
        cpu_reference_setup(inp, state)  # keygen_pk's cosets: one-time in the reference too, not timed
        t0 = time.perf_counter()
        cpu_out = cpu_reference_step(inp, state)
        dt = time.perf_counter() - t0
        aff = orc.g1_to_affine(cpu_out[:shape.n_advice])
        xy = orc.fq_from_mont(aff.reshape(-1, 4)).reshape(-1, 8)
        same = True
        for i in range(shape.n_advice):
            c = bytearray(xy[i, :4].tobytes())
            c[31] |= (int(xy[i, 4]) & 1) << 7
            same = same and bytes(c) == first_proof[32 * i:32 * i + 32]
        # the COMPLETE restated create_proof on the CPU, same circuit, keys, transcript_repr and random draws: its bytes must equal
        # the GPU's; this is the like-for-like CPU baseline of the step
        pp, oparams, opk = cpu_prover_setup(circ, g, g_lagrange)
        wp = circ.witness_pass()
        t0 = time.perf_counter()
        cols = list(advice_mont)
        if wp is not None:  # Circuit::synthesize is the first thing the reference's create_proof does
            buf = np.empty_like(advice_mont)
            wp.run(buf)
            cols = [buf[i] for i in range(buf.shape[0])]
        cpu_proof = pp.create_proof_fast(oparams, opk, cols, circ.instances, randoms)
        dt_full = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dt_full, "unit": UNIT, "cores": orc.ncpu(), "kind": "port", "sample": CPU_SAMPLE,
                                "create_proof_s": dt_full, "proof_bytes_match_gpu": bool(cpu_proof == first_proof),
                                "synthesis_ms": wp.info.synthesis_ms if wp is not None else None,
                                "hot_path_only": {"s": dt, "sample": CPU_HOT_SAMPLE, "advice_commitments_match_gpu_proof": bool(same)}}
        if "ntt" in line:
            # the reference's coeff_to_extended (restated best_fft, all host cores) on one of the proof's columns
            col = np.ascontiguousarray(advice_mont[0])
            t0 = time.perf_counter()
            ext_cpu = state["dom"].coeff_to_extended(col)
            dt_ntt = time.perf_counter() - t0
            line["ntt"]["cpu"] = {"what": f"coeff_to_extended of ONE column 2^{K} -> 2^{int(ext_cpu.shape[0]).bit_length() - 1} (restated best_fft)",
                                  "ms": dt_ntt * 1e3, "gb_s": 64.0 * ext_cpu.shape[0] / dt_ntt / 1e9, "cores": orc.ncpu()}
        if msm_single:
            # the reference's best_multiexp (restated, all host cores) on the very inputs the GPU committed, results compared
            for rec in msm_single:
                host = rec.pop("_host", None)
                if host is None:
                    continue
                scal, bases, got = host
                t0 = time.perf_counter()
                want = orc.best_multiexp(scal, bases)
                rec["cpu_best_multiexp_s"] = time.perf_counter() - t0
                rec["cpu_gpts_s"] = scal.shape[0] / rec["cpu_best_multiexp_s"] / 1e9
                rec["cpu_cores"] = orc.ncpu()
                rec["equals_cpu_best_multiexp"] = bool((orc.g1_to_affine(got.reshape(1, 12)) == orc.g1_to_affine(want.reshape(1, 12))).all())
    if msm_single is not None and rank == 0:
        for rec in msm_single:
            rec.pop("_host", None)
        line["msm"] = {"what": "WHOLE ParamsKZG::commit_lagrange calls (digits, sort, bucket fill, reduction, readback), bases + window tables "
                               "resident, scalars resident, one GPU, CUDA events, best of 3; Gpts/s = polynomials x 2^log_n / time",
                       "runs": msm_single}
        u = {(r.get("log_n"), r.get("polys_per_call")): r.get("gpts_s") for r in msm_single if r.get("scalars") == "U"}
        line["msm_gpts_s"] = u.get((24, 1))
    if rank == 0 and world == 1 and headline and not args.no_other_configs:
        # BASELINE configs[0] / [1] under the driver's eyes: short runs of the other two bench circuits (their byte parity with the
        # CPU restatement is in tests/test_gpu_prover.py)
        others = {}
        for name in ("pose_enc", "mod_pow"):
            try:
                set_config(name)
                l2, a2 = proof_bench(args, local_rank, world, 5, 3, B, detail=False)
                others[name] = {"metric": l2["metric"], "value": l2["value"], "unit": UNIT, "k": K, "proofs_per_step_per_gpu": B, "steps": 5,
                                "ms_per_step": l2["ms_per_step"], "e2e": l2["e2e"]["value"], "latency_create_proof_s": l2["latency"]["create_proof_s"],
                                "proof_bytes": l2["config"]["proof_bytes"], "proof_sha256": l2["config"]["proof_sha256"],
                                "workload": l2["config"]["workload"]}
                close_workers(a2)
                del a2
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001 - the headline line must still be printed
                others[name] = {"error": f"{type(e).__name__}: {e}"}
        set_config(args.config, args.k)
        line["other_configs"] = others
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
