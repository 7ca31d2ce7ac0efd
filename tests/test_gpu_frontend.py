"""The reference's real circuits through the whole stack on the GPU: the front-end's witness
(de_circuit_synthesize / de_circuit_witness) -> keygen (keygen_from_synthesized: the library's permutation assembly, device
commit / lagrange_to_coeff) -> de_create_proof.  The proof bytes must equal the CPU restatement's (oracle/pyprover.py, its own
keygen from the same fixed columns and copy constraints) and pass the restated verifier.
  PoseidonEncCircuit k = 11 (/root/reference/benches/pose_enc.rs:184), DelayEncryptCircuit k = 16 (benches/delay_enc.rs:181)."""
import ctypes as C

import numpy as np
import pytest

import de_b200
import orc
import pyoracle as po
import pyprover as pp
from de_b200 import _lib, frontend as fe, keygen, plonk

pytestmark = pytest.mark.gpu


def ints(cols):
    return [orc.fr_ints_from_mont(np.ascontiguousarray(c)) for c in cols]


def prove_both(syn, seed, fast):
    k, shape = syn.k, syn.shape
    oparams = pp.setup(k, 0x5EC2E7 + k)
    q = pp.Queries(*plonk.collect_queries(shape))
    repr_ = 0xDE1A7E9C0DE + k
    copies = [tuple(int(v) for v in c) for c in syn.copies]
    opk = pp.keygen(oparams, shape, q, ints(syn.fixed), copies, repr_)
    ctx = de_b200.Context(0)
    keys = keygen.keygen_from_synthesized(ctx, syn, oparams.g_mont, oparams.g_lagrange_mont, repr_)
    rng = po.Xoshiro(seed)
    draws = pp.to_mont([rng.uniform_fr() for _ in range(keys.prover.random_count)])
    proof = keys.prover.create_proof([syn.advice[i] for i in range(5)], [np.zeros((0, 4), dtype=np.uint64)], draws)
    adv_ints = ints(syn.advice)
    if fast:
        want = pp.create_proof_fast(oparams, opk, adv_ints, syn.instances, draws)
    else:
        it = iter(pp.from_mont(draws))
        want = pp.create_proof(oparams, opk, adv_ints, syn.instances, lambda: next(it))
    ok = pp.verify_proof(oparams, opk.vk, syn.instances, proof)
    return proof, want, ok, keys, ctx, draws


def test_pose_enc_real_circuit_proof():
    syn = fe.pose_enc((0xC0FFEE, 0xBEEF), [0, 0], k=11)
    proof, want, ok, keys, ctx, draws = prove_both(syn, 0xE11, fast=False)
    assert len(proof) == 1792 and proof == want and ok
    keys.close(); ctx.close()


def test_delay_enc_real_circuit_proof_and_witness_pass():
    n, e, x = fe.sample_rsa_inputs(0xDE03)
    syn = fe.delay_enc(n, e, x, [0, 0], k=16)
    proof, want, ok, keys, ctx, draws = prove_both(syn, 0xE16, fast=True)
    assert len(proof) == 2848 and proof == want and ok
    # create_proof's own synthesis pass (advice only), written straight into a pinned staging buffer
    import torch
    buf = torch.empty((5, 1 << 16, 4), dtype=torch.int64).pin_memory()
    d = fe._Desc()
    d.kind, d.k, d.bits_len, d.exp_bits = fe.DELAY_ENC, 16, 2048, 5
    nb = 256
    bufs = [np.frombuffer(int(v).to_bytes(nb, "little"), dtype=np.uint8).copy() for v in (n, e, x)]
    d.n, d.e, d.x = (b.ctypes.data for b in bufs)
    d.n_len = d.e_len = d.x_len = nb
    msg = np.zeros((2, 4), dtype=np.uint64)
    d.message, d.message_len = msg.ctypes.data, 2
    info = fe._Info()
    rc = _lib.load().de_circuit_witness(C.byref(d), C.c_void_p(buf.data_ptr()), C.byref(info))
    assert rc == 0 and info.used_rows == syn.used_rows
    assert (buf.numpy().view(np.uint64) == syn.advice).all()
    proof2 = keys.prover.create_proof([buf[i] for i in range(5)], [np.zeros((0, 4), dtype=np.uint64)], draws)
    assert proof2 == proof
    # the same pass on several host threads, into the buffer it already filled (threads / reuse_buffer): the same proof
    d.threads, d.reuse_buffer = 6, 1
    for _ in range(2):
        assert _lib.load().de_circuit_witness(C.byref(d), C.c_void_p(buf.data_ptr()), C.byref(info)) == 0
        assert keys.prover.create_proof([buf[i] for i in range(5)], [np.zeros((0, 4), dtype=np.uint64)], draws) == proof
    keys.close(); ctx.close()


def test_mod_pow_real_circuit_proof_k17():
    """RSACircuit at the bench's K = 17 (/root/reference/benches/mod_pow.rs:258)"""
    n, e, x = fe.sample_rsa_inputs(0xDE01)
    syn = fe.mod_pow(n, e, x, k=17)
    proof, want, ok, keys, ctx, draws = prove_both(syn, 0xE17, fast=True)
    assert len(proof) == 2848 and proof == want and ok
    keys.close(); ctx.close()
