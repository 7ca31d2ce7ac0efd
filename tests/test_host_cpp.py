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
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
