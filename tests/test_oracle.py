"""CPU tests of the oracle itself: the reference's own known-answer vectors, the golden fixtures, and the C restatement
against the Python restatement.  (The oracle is the checker for the CUDA path; see oracle/oracle.c header.)"""
import numpy as np
import pytest

import orc
import pyoracle as po
from util import affine_ints, fr, fr_ints, golden, points


def test_poseidon_known_answers_from_reference():
    # /root/reference/src/poseidon/permutation.rs:131-206 (hadeshash poseidonperm_x5_254_3 / _5)
    assert po.poseidon_permute_ref([0, 1, 2], 3, 8, 57) == po.POSEIDON_KAT_T3
    assert po.poseidon_permute_ref([0, 1, 2, 3, 4], 5, 8, 60) == po.POSEIDON_KAT_T5


def test_bn254_constants():
    assert po.FR.bit_length() == 254 and po.FQ.bit_length() == 254
    assert (po.FR - 1) % (1 << 28) == 0 and (po.FR - 1) % (1 << 29) != 0
    w = po.FR_ROOT_OF_UNITY
    assert pow(w, 1 << 28, po.FR) == 1 and pow(w, 1 << 27, po.FR) != 1
    assert w == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert pow(po.FR_ZETA, 3, po.FR) == 1 and po.FR_ZETA != 1
    assert po.FR_DELTA == 0x09226B6E22C6F0CA64EC26AAD4C86E715B5F898E5E963F25870E56BBE533E9A2
    assert po.g1_is_on_curve(po.G1_GEN) and po.g1_mul(po.G1_GEN, po.FR) is None


def test_debug_format_pin():
    # /root/reference/src/big_integer/mod.rs:506: Fp::one() prints as canonical big-endian hex 0x00..01:
    # Montgomery one must decode to the integer 1
    one_m = orc.fr_mont_from_ints([1])
    assert orc.fr_ints_from_mont(one_m) == [1]
    assert [int(x) for x in one_m[0]] == [0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F]


def test_rsa_known_answer_from_reference():
    # /root/reference/src/rsa/chip.rs:706-716: a valid (n, signature, SHA-256 digest) triple.  sig^65537 mod n must carry
    # the PKCS#1 v1.5 limb pattern the chip compares against (:148-199): pins 64-bit limb decomposition and mod-pow.
    n = int("27333278531038650284292446400685983964543820405055158402397263907659995327446166369388984969315774410223081038389734916442552953312548988147687296936649645550823280957757266695625382122565413076484125874545818286099364801140117875853249691189224238587206753225612046406534868213180954324992542640955526040556053150097561640564120642863954208763490114707326811013163227280580130702236406906684353048490731840275232065153721031968704703853746667518350717957685569289022049487955447803273805415754478723962939325870164033644600353029240991739641247820015852898600430315191986948597672794286676575642204004244219381500407")
    sig = int("27166015521685750287064830171899789431519297967327068200526003963687696216659347317736779094212876326032375924944649760206771585778103092909024744594654706678288864890801000499430246054971129440518072676833029702477408973737931913964693831642228421821166326489172152903376352031367604507095742732994611253344812562891520292463788291973539285729019102238815435155266782647328690908245946607690372534644849495733662205697837732960032720813567898672483741410294744324300408404611458008868294953357660121510817012895745326996024006347446775298357303082471522757091056219893320485806442481065207020262668955919408138704593")
    digest = int("83814198383102558219731078260892729932246618004265700685467928187377105751529")
    powed = pow(sig, 65537, n)
    limbs = [(powed >> (64 * i)) & ((1 << 64) - 1) for i in range(32)]
    assert sum(l << (64 * i) for i, l in enumerate(limbs[:4])) == digest           # 1. hash in the low 4 limbs
    assert limbs[4] == 217300885422736416 and limbs[5] == 938447882527703397      # 2. DigestInfo prefix (:151-154)
    assert limbs[6] & 0xFFFFFFFF == 3158320 and limbs[6] >> 32 == 0xFFFFFFFF      #    remaining 24 bits + PS (:177-182)
    assert all(l == 0xFFFFFFFFFFFFFFFF for l in limbs[7:31])                      # 3. PS
    assert limbs[31] == 562949953421311                                           #    0x00 || 0x01 || ff.. (:190-191)
    assert pow((sig + 1) % n, 65537, n) != powed


def test_c_oracle_matches_python_oracle_fields():
    a = orc.uniform_fr(1, 300)
    b = orc.uniform_fr(2, 300)
    ai, bi = orc.fr_ints_from_mont(a), orc.fr_ints_from_mont(b)
    assert orc.fr_ints_from_mont(orc.fr_mul(a, b)) == [x * y % po.FR for x, y in zip(ai, bi)]
    assert orc.fr_ints_from_mont(orc.fr_add(a, b)) == [(x + y) % po.FR for x, y in zip(ai, bi)]
    assert orc.fr_ints_from_mont(orc.fr_sub(a, b)) == [(x - y) % po.FR for x, y in zip(ai, bi)]
    assert orc.fr_ints_from_mont(orc.fr_inv(a[:20])) == [pow(x, -1, po.FR) for x in ai[:20]]
    aq, bq = orc.fq_ints_from_mont(a), orc.fq_ints_from_mont(b)
    assert orc.fq_ints_from_mont(orc.fq_mul(a, b)) == [x * y % po.FQ for x, y in zip(aq, bq)]


def test_prng_shared_between_oracles():
    r = po.Xoshiro(0xDE04)
    assert orc.fr_ints_from_mont(orc.uniform_fr(0xDE04, 50)) == [r.uniform_fr() for _ in range(50)]


def test_golden_ntt():
    for case in golden()["ntt"]:
        a = fr(case["in"])
        omega = fr([case["omega"]])[0]
        want = [int(x, 16) for x in case["out"]]
        assert fr_ints(orc.best_fft(a, omega, case["log_n"], threads=1)) == want
        assert fr_ints(orc.best_fft(a, omega, case["log_n"], threads=4)) == want
        assert po.best_fft([int(x, 16) for x in case["in"]], int(case["omega"], 16), case["log_n"]) == want


def test_golden_msm():
    for case in golden()["msm"]:
        s = fr(case["scalars"])
        b = points(case["bases"])
        want = None if case["result"] is None else (int(case["result"][0], 16), int(case["result"][1], 16))
        for th in (1, 3):
            assert affine_ints(orc.best_multiexp(s, b, threads=th)) == want
        assert affine_ints(orc.msm_naive(s, b)) == want


def test_golden_domain():
    for case in golden()["domain"]:
        d = orc.Domain(case["j"], case["k"], threads=2)
        assert d.extended_k == case["extended_k"]
        assert fr_ints(d.omega) == [int(case["omega"], 16)]
        assert fr_ints(d.ext_omega) == [int(case["extended_omega"], 16)]
        coeff = fr(case["coeff"])
        ext = d.coeff_to_extended(coeff)
        assert fr_ints(ext) == [int(x, 16) for x in case["extended"]]
        assert fr_ints(d.coeff_to_lagrange(coeff)) == [int(x, 16) for x in case["lagrange"]]
        assert (d.lagrange_to_coeff(fr(case["lagrange"])) == coeff).all()
        assert fr_ints(d.t_evaluations()) == [int(x, 16) for x in case["t_evaluations"]]
        back = d.extended_to_coeff(ext)
        assert back.shape[0] == (case["j"] - 1) << case["k"]
        assert (back[: 1 << case["k"]] == coeff).all() and not back[1 << case["k"]:].any()


@pytest.mark.parametrize("k", [1, 2, 5, 8])
def test_c_oracle_domain_vs_python(k):
    d, pd = orc.Domain(5, k), po.EvaluationDomain(5, k)
    x = orc.uniform_fr(10 + k, 1 << k)
    xi = orc.fr_ints_from_mont(x)
    assert fr_ints(orc.best_fft(x, d.omega, k)) == po.best_fft(xi, pd.omega, k)
    e = d.coeff_to_extended(x)
    assert fr_ints(e) == pd.coeff_to_extended(xi)
    assert fr_ints(d.divide_by_vanishing(e)) == pd.divide_by_vanishing_poly(fr_ints(e))
    assert fr_ints(d.extended_to_coeff(e)) == pd.extended_to_coeff(fr_ints(e))


def test_c_oracle_msm_edge_cases():
    n = 64
    bases = orc.gen_bases(n)
    # all-zero scalars -> identity; all-ones -> sum of bases; duplicates of one base -> doubling paths
    z = np.zeros((n, 4), dtype=np.uint64)
    assert affine_ints(orc.best_multiexp(z, bases)) is None
    ones = orc.fr_mont_from_ints([1] * n)
    assert affine_ints(orc.best_multiexp(ones, bases)) == po.g1_mul(po.G1_GEN, n * (n + 1) // 2)
    same = np.repeat(bases[:1], n, axis=0)
    assert affine_ints(orc.best_multiexp(ones, same)) == po.g1_mul(po.G1_GEN, n)
    # P + (-P) = identity
    neg = bases[:2].copy()
    neg[1, :4] = neg[0, :4]
    y = orc.fq_ints_from_mont(neg[0, 4:].reshape(1, 4))[0]
    neg[1, 4:] = orc.fq_mont_from_ints([po.FQ - y])[0]
    assert affine_ints(orc.best_multiexp(ones[:2], neg)) is None
    assert affine_ints(orc.best_multiexp(ones[:0], bases[:0])) is None  # empty input
