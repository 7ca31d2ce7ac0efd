"""A proving key through its RawBytes file and back into the prover (SURVEY.md section 8f row 4;
/root/reference/benches/delay_enc.rs:101-115 pk.write -> ProvingKey::read): the prover staged from the FILE produces the same
proof bytes as the prover staged by keygen, and the file's cosets / l0 / l_last / l_active_row equal the CPU restatement's."""
import numpy as np
import pytest

import de_b200
import orc
import pyoracle as po
import pyprover as pp
from de_b200 import frontend as fe, keygen, plonk, serde

pytestmark = pytest.mark.gpu


def test_pk_file_round_trip_gives_identical_proof(tmp_path):
    k = 11
    syn = fe.pose_enc((0xAB, 0xCD), [0, 0], k=k)
    oparams = pp.setup(k, 0x5EC2E7 + k)
    ctx = de_b200.Context(0)
    repr_ = 0x1234567
    keys = keygen.keygen_from_synthesized(ctx, syn, oparams.g_mont, oparams.g_lagrange_mont, repr_)
    sigma = syn.sigma(keys.domain.omega)
    raw = serde.proving_key_raw_from_keys(keys, syn.fixed, sigma)
    serde.write_pk(tmp_path / "pk_pose_enc_11", raw)
    serde.write_vk(tmp_path / "vk_pose_enc_11", raw.vk)
    assert (tmp_path / "vk_pose_enc_11").stat().st_size == 968     # /root/reference/benches/README.md:90
    n, ext = 1 << k, keys.domain.extended_n
    F, P = 9, 6
    assert (tmp_path / "pk_pose_enc_11").stat().st_size == 968 + 3 * (4 + 32 * ext) + 2 * (4 + F * (4 + 32 * n)) + (4 + F * (4 + 32 * ext)) \
        + 2 * (4 + P * (4 + 32 * n)) + (4 + P * (4 + 32 * ext))
    back = serde.read_pk(tmp_path / "pk_pose_enc_11", syn.shape)
    # the file's derived parts against the CPU restatement of keygen_pk
    odom = orc.Domain(syn.shape.degree(), k)
    assert (back.fixed_cosets[3] == odom.coeff_to_extended(np.ascontiguousarray(back.fixed_polys[3]))).all()
    assert (back.cosets[5] == odom.coeff_to_extended(np.ascontiguousarray(back.polys[5]))).all()
    e0 = np.zeros((n, 4), dtype=np.uint64)
    e0[0] = orc.fr_mont_from_ints([1])[0]
    assert (back.l0 == odom.coeff_to_extended(odom.lagrange_to_coeff(e0))).all()
    one_minus = orc.fr_sub(np.tile(orc.fr_mont_from_ints([1]), (ext, 1)), orc.fr_add(np.ascontiguousarray(back.l_last), np.ascontiguousarray(back.l_active_row)))
    # 1 - l_last - l_active_row = l_blind: the coset of the polynomial that is 1 on the last 5 rows
    blind = np.zeros((n, 4), dtype=np.uint64)
    blind[n - 5:] = orc.fr_mont_from_ints([1])[0]
    assert (one_minus == odom.coeff_to_extended(odom.lagrange_to_coeff(blind))).all()
    # a prover staged from the file, on a second context
    ctx2 = de_b200.Context(0)
    keys2 = serde.keys_from_proving_key_raw(ctx2, back, syn.shape, oparams.g_mont, oparams.g_lagrange_mont, repr_)
    assert (keys2.fixed_commitments == keys.fixed_commitments).all() and (keys2.permutation_commitments == keys.permutation_commitments).all()
    rng = po.Xoshiro(0x5E4DE)
    draws = pp.to_mont([rng.uniform_fr() for _ in range(keys.prover.random_count)])
    adv = [syn.advice[i] for i in range(5)]
    p1 = keys.prover.create_proof(adv, [np.zeros((0, 4), dtype=np.uint64)], draws)
    p2 = keys2.prover.create_proof(adv, [np.zeros((0, 4), dtype=np.uint64)], draws)
    assert p1 == p2 and len(p1) == 1792
    keys.close(); keys2.close(); ctx.close(); ctx2.close()
