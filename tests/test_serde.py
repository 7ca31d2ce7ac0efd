"""SerdeFormat::RawBytes key files (/root/reference/benches/delay_enc.rs:84-115): layout, sizes
against the reference README's pk / vk sizes, round trips, error paths.  CPU only; the GPU round trip through the prover is
tests/test_gpu_serde.py."""
import io

import numpy as np
import pytest

from de_b200 import plonk, serde


def rnd(shape, seed):
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 1 << 62, size=shape, dtype=np.uint64)


def make_pk(shape, k, ext_factor):
    n, ext = 1 << k, (1 << k) * ext_factor
    F, P = shape.n_fixed, len(shape.perm_columns)
    sels = [rnd((n,), 90 + i) % 2 == 1 for i in range(serde.n_selectors_of(shape))]
    vk = serde.VerifyingKeyRaw(k, rnd((F, 8), 1), rnd((P, 8), 2), sels)
    return serde.ProvingKeyRaw(vk, rnd((ext, 4), 3), rnd((ext, 4), 4), rnd((ext, 4), 5), rnd((F, n, 4), 6), rnd((F, n, 4), 7),
                               rnd((F, ext, 4), 8), rnd((P, n, 4), 9), rnd((P, n, 4), 10), rnd((P, ext, 4), 11))


@pytest.mark.parametrize("with_range,ext_factor", [(False, 2), (True, 4)])
@pytest.mark.parametrize("order", ["big", "little"])
def test_pk_vk_round_trip(tmp_path, with_range, ext_factor, order):
    shape = plonk.main_gate_shape(with_range)
    k = 6
    pk = make_pk(shape, k, ext_factor)
    serde.write_pk(tmp_path / "pk", pk, order)
    serde.write_vk(tmp_path / "vk", pk.vk, order)
    back = serde.read_pk(tmp_path / "pk", shape)
    vk = serde.read_vk(tmp_path / "vk", shape)
    for name in ("l0", "l_last", "l_active_row", "fixed_values", "fixed_polys", "fixed_cosets", "permutations", "polys", "cosets"):
        assert (getattr(back, name) == getattr(pk, name)).all(), name
    for v in (back.vk, vk):
        assert v.k == k and (v.fixed_commitments == pk.vk.fixed_commitments).all() and (v.permutation_commitments == pk.vk.permutation_commitments).all()
        assert len(v.selectors) == len(pk.vk.selectors) and all((a == b).all() for a, b in zip(v.selectors, pk.vk.selectors))


def test_sizes_match_the_reference_readme():
    """vk: 968 B (Poseidon-only shape) and 17.32 KiB at k = 16 (RSA shape); pk: 32 (3 ext + (F + P)(2 n + ext)) + headers =
    276.0 MiB at k = 16 (/root/reference/benches/README.md:56-63, 89-99)"""
    def vk_size(shape, k):
        f = io.BytesIO()
        n = 1 << k
        serde._write_vk(f, serde.VerifyingKeyRaw(k, np.zeros((shape.n_fixed, 8), np.uint64), np.zeros((len(shape.perm_columns), 8), np.uint64),
                                                 [np.zeros(n, bool)] * serde.n_selectors_of(shape)), "big")
        return len(f.getvalue())
    assert vk_size(plonk.main_gate_shape(False), 11) == 968
    assert vk_size(plonk.main_gate_shape(True), 16) == 8 + 64 * 21 + 2 * 8192 == 17736
    # pk size from the layout, without materialising 276 MiB
    F, P, n, ext = 15, 6, 1 << 16, 1 << 18
    body = 3 * (4 + 32 * ext) + 2 * (4 + F * (4 + 32 * n)) + (4 + F * (4 + 32 * ext)) + 2 * (4 + P * (4 + 32 * n)) + (4 + P * (4 + 32 * ext))
    assert abs((17736 + body) / 2 ** 20 - 276.0) < 0.1


def test_read_errors(tmp_path):
    shape = plonk.main_gate_shape(False)
    pk = make_pk(shape, 5, 2)
    buf = io.BytesIO()
    serde._write_vk(buf, pk.vk, "big")
    good = buf.getvalue()
    with pytest.raises(ValueError):
        serde.read_vk(good + b"x", shape)                         # trailing bytes
    with pytest.raises(ValueError):
        serde.read_vk(good[:-3], shape)                           # truncated
    with pytest.raises(ValueError):
        serde.read_vk(b"\xff\xff\xff\xff" + good[4:], shape)      # k out of range in both byte orders
    with pytest.raises(ValueError):
        serde.read_vk(good, plonk.main_gate_shape(True))          # another constraint system
