"""The C-ABI boundary: every function include/de_b200.h declares is exported by libde_b200.so and bound by the Python
host layer; without a GPU the library must refuse to create a context (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "de_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(de_[a-z0-9_]+)\s*\(", src)))


def test_header_and_library_agree():
    import __graft_entry__ as g
    from de_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    L = C.CDLL(_lib.LIB_PATH)
    fns = header_functions()
    assert len(fns) >= 30
    assert sorted(_lib.SYMBOLS) == fns
    for f in fns:
        assert hasattr(L, f), f"libde_b200.so does not export {f}"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import de_b200
    with pytest.raises(de_b200.DeError) as e:
        de_b200.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "delay-encryption-in-halo2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "NOTES", f"{f} mentions the oracle: the product path must not use it"


def test_params_raw_bytes_file_round_trip(tmp_path):
    """SerdeFormat::RawBytes params file: k | g | g_lagrange | g2 | s_g2, Montgomery limbs verbatim"""
    import numpy as np
    import de_b200
    import orc
    k = 4
    g, gl = orc.gen_bases(16), orc.gen_bases(16, start=16)
    g2, s_g2 = bytes(range(128)), bytes(range(128, 256))
    path = tmp_path / "params_4"
    de_b200.write_params_raw(path, k, g, gl, g2, s_g2)
    assert path.stat().st_size == 4 + 2 * 16 * 64 + 256
    d = de_b200.read_params_raw(path)
    assert d["k"] == k and (np.asarray(d["g"]) == g).all() and (np.asarray(d["g_lagrange"]) == gl).all()
    assert d["g2"] == g2 and d["s_g2"] == s_g2
    with open(path, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        de_b200.read_params_raw(path)


def test_rust_sys_crate_declares_every_header_symbol():
    """integration/de-b200-sys/src/lib.rs (compile-unverified: no Rust toolchain here) stays in sync with include/de_b200.h"""
    import re
    from de_b200 import _lib
    src = open(os.path.join(ROOT, "integration", "de-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (de_[a-z0-9_]+)\(", src))
    assert declared == set(_lib.SYMBOLS)


def test_circuit_kinds_agree_across_header_python_and_rust():
    """enum de_circuit_kind of include/de_b200.h, the constants of de_b200/frontend.py and of the Rust -sys crate"""
    import re
    from de_b200 import frontend as fe
    hdr = open(os.path.join(ROOT, "include", "de_b200.h")).read()
    body = re.search(r"enum de_circuit_kind \{(.*?)\};", hdr, re.S).group(1)
    kinds = {m.group(1): int(m.group(2)) for m in re.finditer(r"DE_CIRCUIT_([A-Z0-9_]+) = (\d+)", body)}
    assert sorted(kinds.values()) == list(range(len(kinds))) and len(kinds) == 7
    for name, value in kinds.items():
        assert getattr(fe, name) == value, name
    rs = open(os.path.join(ROOT, "integration", "de-b200-sys", "src", "lib.rs")).read()
    rust = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const DE_CIRCUIT_([A-Z0-9_]+): u32 = (\d+);", rs)}
    assert rust == kinds
