"""The C++ host-side mirror (host/halo2_b200.hpp): compiles everywhere (CPU test), runs bit-exact against the oracle on a GPU."""
import os
import subprocess

import numpy as np
import pytest

import orc
import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "delay-encryption-in-halo2_b200")
EXE = os.path.join(ROOT, "tests", "hostcpp", "host_mirror_test")


def build_exe():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(PKG, "libde_b200.so")):
        g.build()
    src = os.path.join(ROOT, "tests", "hostcpp", "host_mirror_test.cpp")
    hdr = os.path.join(PKG, "host", "halo2_b200.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(PKG, "host"), src, "-o", EXE, "-L", PKG, "-lde_b200",
                               f"-Wl,-rpath,{PKG}"])
    return EXE


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(build_exe())


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path):
    exe = build_exe()
    k = 9
    n = 1 << k
    s = orc.uniform_fr(404, n)
    b = orc.gen_bases(n)
    d = orc.Domain(5, k)
    s.tofile(tmp_path / "scalars.bin")
    b.tofile(tmp_path / "bases.bin")
    d.omega.tofile(tmp_path / "omega.bin")
    orc.best_fft(s, d.omega, k).tofile(tmp_path / "fft.bin")
    d.coeff_to_extended(s).tofile(tmp_path / "ext.bin")
    orc.g1_to_affine(orc.best_multiexp(s, b)).tofile(tmp_path / "msm_affine.bin")
    # best_fft_sharded (two contexts): a 2^13 vector
    d13 = orc.Domain(2, 13)
    s13 = orc.uniform_fr(407, 1 << 13)
    s13.tofile(tmp_path / "sharded_in.bin")
    d13.omega.tofile(tmp_path / "sharded_omega.bin")
    orc.best_fft(s13, d13.omega, 13).tofile(tmp_path / "sharded_fft.bin")
    # eval_polynomial / kate_division
    point = orc.uniform_fr(405, 1)
    point.tofile(tmp_path / "point.bin")
    orc.eval_polynomial(s, point[0]).tofile(tmp_path / "eval.bin")
    orc.kate_division(s, point[0]).tofile(tmp_path / "kate.bin")
    # a whole proof from a serialised proving key (MainGate-only shape, k = 5), expected bytes from the restated prover
    import pyprover as pp
    from de_b200 import circuits, plonk
    pdir = tmp_path / "proof"
    pdir.mkdir()
    kk = 5
    asg = circuits.satisfied_assignment(False, kk, 0xC9905, 20)
    shape = asg.shape
    oparams = pp.setup(kk, 0x5EC2E7)
    q = pp.Queries(*plonk.collect_queries(shape))
    trepr = 0xC0FFEE
    opk = pp.keygen(oparams, shape, q, asg.fixed, asg.copies, trepr)
    draws = orc.uniform_fr(406, pp.random_count(shape, 1 << kk))
    want = pp.create_proof_fast(oparams, opk, asg.advice, asg.instances, draws)
    keep = []
    graph = plonk._marshal_graph(plonk.compile_gates(shape.gates), keep)
    consts, rots, calcs, parts = keep
    np.array([kk, shape.n_fixed, shape.n_advice, shape.n_instance, len(shape.perm_columns), shape.chunk_len, shape.blinding_factors,
              graph.n_intermediates, shape.degree()], dtype=np.uint32).tofile(pdir / "meta.bin")
    np.stack([pp.to_mont(p) for p in opk.fixed_polys]).tofile(pdir / "fixed_coeff.bin")
    np.stack([pp.to_mont(p) for p in opk.sigma_polys]).tofile(pdir / "sigma_coeff.bin")
    # the same key as a SerdeFormat::RawBytes ProvingKey file (cosets / l-polynomials are not consumed by the reader's caller)
    from de_b200 import serde
    F_, P_, n_, ext_ = shape.n_fixed, len(shape.perm_columns), 1 << kk, 1 << orc.Domain(shape.degree(), kk).extended_k
    zz = lambda *sh: np.zeros(sh, dtype=np.uint64)
    mm = lambda cols: np.stack([pp.to_mont(c) for c in cols])
    serde.write_pk(pdir / "pk.bin", serde.ProvingKeyRaw(serde.VerifyingKeyRaw(kk, zz(F_, 8), zz(P_, 8), []), zz(ext_, 4), zz(ext_, 4), zz(ext_, 4),
                                                        mm(opk.fixed_values), mm(opk.fixed_polys), zz(F_, ext_, 4), mm(opk.sigma_values),
                                                        mm(opk.sigma_polys), zz(P_, ext_, 4)))
    np.stack([pp.to_mont(c) for c in asg.advice]).tofile(pdir / "advice.bin")
    draws.tofile(pdir / "randoms.bin")
    np.array(plonk.mont_limbs(plonk.FR_DELTA), dtype=np.uint64).tofile(pdir / "delta.bin")
    np.array(plonk.mont_limbs(trepr), dtype=np.uint64).tofile(pdir / "transcript_repr.bin")
    oparams.g_mont.tofile(pdir / "g.bin")
    oparams.g_lagrange_mont.tofile(pdir / "g_lagrange.bin")
    np.array([kind for kind, _ in shape.perm_columns], dtype=np.uint32).tofile(pdir / "perm_kind.bin")
    np.array([i for _, i in shape.perm_columns], dtype=np.uint32).tofile(pdir / "perm_index.bin")
    consts.tofile(pdir / "g_consts.bin")
    rots.tofile(pdir / "g_rots.bin")
    open(pdir / "g_calcs.bin", "wb").write(bytes(calcs)[: 40 * len(plonk.compile_gates(shape.gates).calculations)])
    n_parts = sum(len(c[3]) for c in plonk.compile_gates(shape.gates).calculations)
    open(pdir / "g_parts.bin", "wb").write(bytes(parts)[: 12 * n_parts])
    np.array([c for c, _ in q.advice], dtype=np.uint32).tofile(pdir / "aq_col.bin")
    np.array([r_ for _, r_ in q.advice], dtype=np.int32).tofile(pdir / "aq_rot.bin")
    np.array([c for c, _ in q.fixed], dtype=np.uint32).tofile(pdir / "fq_col.bin")
    np.array([r_ for _, r_ in q.fixed], dtype=np.int32).tofile(pdir / "fq_rot.bin")
    open(pdir / "want_proof.bin", "wb").write(want)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


CIRCUITS_EXE = os.path.join(ROOT, "tests", "hostcpp", "host_circuits_test")


def test_cpp_mirror_circuits_match_python_frontend(tmp_path):
    """DelayEncryptCircuit / RSACircuit / PoseidonEncCircuit of host/halo2_b200.hpp (the reference's circuit types: src/lib.rs:103,
    benches/mod_pow.rs:36, src/encryption/chip.rs:114) against the Python mirror of the same front-end: results, advice, copy
    constraints, used rows; threaded witness pass equals the keygen pass; too few rows throws.  No GPU involved."""
    from de_b200 import frontend as fe
    src = os.path.join(ROOT, "tests", "hostcpp", "host_circuits_test.cpp")
    hdr = os.path.join(PKG, "host", "halo2_b200.hpp")
    lib = os.path.join(PKG, "libde_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    if not os.path.exists(CIRCUITS_EXE) or os.path.getmtime(CIRCUITS_EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(PKG, "host"), src, "-o", CIRCUITS_EXE, "-L", PKG, "-lde_b200",
                               f"-Wl,-rpath,{PKG}"])
    n, e, x = fe.sample_rsa_inputs(0xC99)
    for name, v in (("n", n), ("e", e), ("x", x)):
        (tmp_path / f"{name}.bin").write_bytes(int(v).to_bytes(256, "little"))
    key = (0x5EED1, 0x5EED2)
    orc.fr_mont_from_ints(list(key)).tofile(tmp_path / "key.bin")
    r = subprocess.run([CIRCUITS_EXE, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr

    def ints(name):
        return orc.fr_ints_from_mont(np.fromfile(tmp_path / name, dtype=np.uint64).reshape(-1, 4))

    d = fe.delay_enc(n, e, x, [0, 0], k=16)
    assert ints("delay_outputs.bin") == d.outputs
    assert sum(l << (64 * i) for i, l in enumerate(d.outputs[:32])) == pow(x, e, n)
    assert (np.fromfile(tmp_path / "delay_advice0.bin", dtype=np.uint64).reshape(-1, 4) == d.advice[0]).all()
    assert (np.fromfile(tmp_path / "delay_copies.bin", dtype=np.uint32).reshape(-1, 4) == d.copies).all()
    assert int((tmp_path / "delay_rows.txt").read_text()) == d.used_rows
    assert ints("rsa_outputs.bin") == fe.mod_pow(n, e, x, k=17).outputs
    p = fe.pose_enc(key, [0, 0], k=11)
    assert ints("pose_outputs.bin") == p.outputs == fe.poseidon_encrypt(key, [0, 0])
    assert (np.fromfile(tmp_path / "pose_fixed3.bin", dtype=np.uint64).reshape(-1, 4) == p.fixed[3]).all()
