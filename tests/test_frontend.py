"""Circuit front-end (C++ in delay-encryption-in-halo2_b200/frontend/, C ABI de_circuit_*): the
reference's chips and bench circuits as witness generators.  CPU tests (the front-end needs no GPU):

  * the reference's own known-answer vectors: Poseidon permutations (src/poseidon/permutation.rs:154-158,190-196), the
    PKCS#1 v1.5 RSA triples (src/rsa/chip.rs:706-716 valid, :796-806 corrupted), encrypt -> decrypt (poseidon_enc.rs:167-177);
  * every synthesised circuit satisfies its constraint system (oracle/mockprover.py, the MockProver check the reference's
    tests run: src/lib.rs:353, src/encryption/chip.rs:237, src/rsa/chip.rs:341);
  * used rows against the reference README's row counts; the witness-only pass equals the full pass; the library's permutation
    assembly equals the Python one.
"""
import numpy as np
import pytest

import mockprover
import orc
import pyoracle as po
from de_b200 import frontend as fe
from de_b200 import keygen, plonk


def ints(cols):
    return [orc.fr_ints_from_mont(np.ascontiguousarray(c)) for c in cols]


def mock_check(syn):
    mockprover.check(syn.shape, syn.k, ints(syn.fixed), ints(syn.advice), syn.instances, [tuple(int(v) for v in c) for c in syn.copies])


# the reference's PKCS#1 v1.5 known-answer triples (n, signature, SHA-256 digest as an integer):
# /root/reference/src/rsa/chip.rs:706-716 and :751-761 verify, :796-806 (one digit of the second signature changed) does not
RSA_N1 = int("27333278531038650284292446400685983964543820405055158402397263907659995327446166369388984969315774410223081038389734916442552953312548988147687296936649645550823280957757266695625382122565413076484125874545818286099364801140117875853249691189224238587206753225612046406534868213180954324992542640955526040556053150097561640564120642863954208763490114707326811013163227280580130702236406906684353048490731840275232065153721031968704703853746667518350717957685569289022049487955447803273805415754478723962939325870164033644600353029240991739641247820015852898600430315191986948597672794286676575642204004244219381500407")
RSA_SIG1 = int("27166015521685750287064830171899789431519297967327068200526003963687696216659347317736779094212876326032375924944649760206771585778103092909024744594654706678288864890801000499430246054971129440518072676833029702477408973737931913964693831642228421821166326489172152903376352031367604507095742732994611253344812562891520292463788291973539285729019102238815435155266782647328690908245946607690372534644849495733662205697837732960032720813567898672483741410294744324300408404611458008868294953357660121510817012895745326996024006347446775298357303082471522757091056219893320485806442481065207020262668955919408138704593")
RSA_N2 = int("24226501697440012621102249466312043787685293040734225606346036389705515508545746221669035424138747582133889500686654172873671086178893587422987328751464627501601101326475761646014534358699943642495332701081302954020983110372109611581202820849485662540890985814355975252780310958088652613376767040069489530039075302709233494829280591680666351811024913107949144932224439129715181798714328219977771472462901856297952813239115577652450722815852332547886777292613005505949100406231716599634852632308325816916535875123863510650526931916871614411907700873376659841257216885666098127478325534982891697988739616416855214839339")
RSA_SIG2 = int("18928545496959757512579438348223103860103247450097569223971486743312798156950374943336714741350742176674694049986481729075548718599712271054643150030165230392897481507710187505775911256946250999396358633095137650326818007610162375520522758780751710735664264200260854016867498935206556916247099180950775474524799944404833222133011134000549939512938205188018503377612813102061504146765520561811620128786062447005833886367575841545493555268747671930923697279690399480501746857825917608323993022396398648205737336204493624060285359455268389160802763426461171262704764369336704988874821898000892148693988241020931055723252")
RSA_SIG2_BAD = int("18928545496959756512579438348223103860103247450097569223971486743312798156950374943336714741350742176674694049986481729075548718599712271054643150030165230392897481507710187505775911256946250999396358633095137650326818007610162375520522758780751710735664264200260854016867498935206556916247099180950775474524799944404833222133011134000549939512938205188018503377612813102061504146765520561811620128786062447005833886367575841545493555268747671930923697279690399480501746857825917608323993022396398648205737336204493624060285359455268389160802763426461171262704764369336704988874821898000892148693988241020931055723252")
RSA_DIGEST = int("83814198383102558219731078260892729932246618004265700685467928187377105751529")


def test_poseidon_permutation_known_answers():
    want3 = [7853200120776062878684798364095072458815029376092732009249414926327459813530,
             7142104613055408817911962100316808866448378443474503659992478482890339429929,
             6549537674122432311777789598043107870002137484850126429160507761192163713804]
    assert fe.poseidon_permute([0, 1, 2], 3, 8, 57) == want3
    want5 = [18821383157269793795438455681495246036402687001665670618754263018637548127333,
             7817711165059374331357136443537800893307845083525445872661165200086166013245,
             16733335996448830230979566039396561240864200624113062088822991822580465420551,
             6644334865470350789317807668685953492649391266180911382577082600917830417726,
             3372108894677221197912083238087960099443657816445944159266857514496320565191]
    assert fe.poseidon_permute([0, 1, 2, 3, 4], 5, 8, 60) == want5
    # the optimised parameter set agrees with the plain round function of the checker's own restatement
    assert fe.poseidon_permute([5, 6, 7, 8, 9], 5, 8, 57) == po.poseidon_permute_ref([5, 6, 7, 8, 9], 5, 8, 57)


@pytest.mark.parametrize("t", range(3, 11))
def test_poseidon_optimised_against_plain_rounds(t):
    """/root/reference/src/poseidon/permutation.rs:84-129 cross_test: the optimised parameter set (sparse MDS factorisation,
    pre-merged constants) against the plain round function, R_F = 8, R_P = 57, every width the reference runs (3 ... 10)"""
    import random
    rng = random.Random(0xC705 + t)
    for _ in range(2):
        state = [rng.randrange(po.FR) for _ in range(t)]
        assert fe.poseidon_permute(state, t, 8, 57) == po.poseidon_permute_ref(state, t, 8, 57)


def _sponge_hash(inputs):
    """poseidon.rs: Poseidon::new_hash(8, 57), update(inputs), squeeze(1), on the plain round function of the checker"""
    state = [1 << 64, 0, 0, 0, 0]
    full, rest = len(inputs) // 4 * 4, len(inputs) % 4
    for off in range(0, full, 4):
        for i in range(4):
            state[1 + i] = (state[1 + i] + inputs[off + i]) % po.FR
        state = po.poseidon_permute_ref(state, 5, 8, 57)
    for i, v in enumerate(list(inputs[full:]) + [1]):
        state[1 + i] = (state[1 + i] + v) % po.FR
    return po.poseidon_permute_ref(state, 5, 8, 57)


def test_poseidon_hash_circuit():
    """/root/reference/src/hash/chip.rs:203-236 test_example_hash (four zero inputs, T = 5, RATE = 4) and the same circuit for
    every input count around the rate: HasherChip equals the native sponge (constrained inside the circuit) and the checker's
    plain-round restatement; the circuit is satisfied"""
    import random
    rng = random.Random(0x4A54)
    syn = fe.poseidon_hash([0, 0, 0, 0])
    assert syn.outputs == _sponge_hash([0, 0, 0, 0])
    mock_check(syn)
    for count in (0, 1, 3, 4, 5, 8, 11):
        inputs = [rng.randrange(po.FR) for _ in range(count)]
        syn = fe.poseidon_hash(inputs)
        assert syn.outputs == _sponge_hash(inputs), count
        mock_check(syn)
    # the hash region of DelayEncryptCircuit is this circuit over the 11 packed limbs (test_delay_enc_circuit checks the key)


def test_encrypt_decrypt_round_trip():
    for key in ((0, 0), (0x1234, po.FR - 5)):
        c = fe.poseidon_encrypt(key, [0, 0])   # the message of the reference's test and benches
        assert len(c) == 3 and fe.poseidon_decrypt(key, c) == [0, 0]
        bad = list(c)
        bad[2] = (bad[2] + 1) % po.FR
        assert fe.poseidon_decrypt(key, bad) is None
        assert fe.poseidon_decrypt((key[0] + 1, key[1]), c) is None
    # ciphertext words are state + message; the tag is state[1] after the second permutation
    s = fe.poseidon_permute([0, 0, 7, 9, 1], 5, 8, 57)
    c = fe.poseidon_encrypt((7, 9), [0, 0])
    assert c[0] == s[1] and c[1] == s[2] and c[2] == fe.poseidon_permute(s, 5, 8, 57)[1]


def test_pose_enc_circuit_satisfied_and_row_count():
    key = (0x5EED1, 0x5EED2)
    syn = fe.pose_enc(key, [0, 0], k=11)
    assert syn.outputs == fe.poseidon_encrypt(key, [0, 0])
    assert abs(syn.used_rows - 1450) <= 5          # /root/reference/benches/README.md:90 "1450" used rows
    assert syn.fixed.shape == (9, 2048, 4) and syn.advice.shape == (5, 2048, 4)
    mock_check(syn)
    wit = fe.synthesize(fe.POSE_ENC, 11, message=[0, 0], key=key, witness_only=True)
    assert (wit.advice == syn.advice).all() and wit.fixed.shape[0] == 0 and len(wit.copies) == 0


def test_pose_enc_nonzero_message_is_rejected_like_the_reference():
    # the circuit adds the message twice (encryption/chip.rs:93-103 then chip.rs:237-249), the native cipher never absorbs a
    # short message (poseidon_enc.rs:107-124): they agree on the all-zero message only, and assert_equal fails otherwise
    with pytest.raises(fe.DeError):
        fe.pose_enc((1, 2), [3, 4], k=11)


def test_tampered_witness_is_caught_by_the_checker():
    syn = fe.pose_enc((11, 22), [0, 0], k=11)
    fixed, copies = ints(syn.fixed), [tuple(int(v) for v in c) for c in syn.copies]
    # a cell the gate reads (s_mul_ab = 1 on that row): the gate fails; a cell only a copy constraint reads: the copy fails
    row = next(r for r in range(syn.used_rows) if fixed[6][r] == 1)
    adv = ints(syn.advice)
    adv[0][row] = (adv[0][row] + 1) % po.FR
    with pytest.raises(AssertionError, match="gate|copy"):
        mockprover.check(syn.shape, syn.k, fixed, adv, syn.instances, copies)
    adv = ints(syn.advice)
    lc, lr, rc, rr = copies[len(copies) // 2]
    adv[rc][rr] = (adv[rc][rr] + 1) % po.FR
    with pytest.raises(AssertionError):
        mockprover.check(syn.shape, syn.k, fixed, adv, syn.instances, copies)


def test_sigma_columns_match_python_assembly():
    syn = fe.pose_enc((3, 4), [0, 0], k=11)
    n = 1 << syn.k
    omega = pow(po.FR_ROOT_OF_UNITY, 1 << (po.FR_S - syn.k), po.FR)
    got = ints(syn.sigma(orc.fr_mont_from_ints([omega])[0]))
    asm = keygen.PermutationAssembly(len(syn.shape.perm_columns), n)
    for lc, lr, rc, rr in syn.copies:
        asm.copy(int(lc), int(lr), int(rc), int(rr))
    assert got == asm.sigma_values(omega)


def test_mod_pow_circuit():
    n, e, x = fe.sample_rsa_inputs(0xDE01)
    syn = fe.mod_pow(n, e, x, k=16)    # 41 766 rows fit k = 16 (the bench uses K = 17)
    assert sum(l << (64 * i) for i, l in enumerate(syn.outputs)) == pow(x, e, n)
    assert abs(syn.used_rows - 41766) / 41766 < 0.05    # /root/reference/benches/README.md:73
    assert syn.fixed.shape[0] == 15
    mock_check(syn)


def test_mod_pow_circuit_1024_bits():
    """/root/reference/src/rsa/chip.rs:462-513 TestRSAModPow1024Circuit: the same chip at BITS_LEN = 1024 (16 limbs)"""
    for seed in (7, 8):
        n, e, x = fe.sample_rsa_inputs(seed, bits_len=1024)
        syn = fe.synthesize(fe.MOD_POW, 15, n, e, x, bits_len=1024)
        assert len(syn.outputs) <= 16 and sum(l << (64 * i) for i, l in enumerate(syn.outputs)) == pow(x, e, n)
        mock_check(syn)
        wit = fe.synthesize(fe.MOD_POW, 15, n, e, x, bits_len=1024, witness_only=True, threads=3)
        assert (wit.advice == syn.advice).all()


def test_delay_enc_circuit():
    n, e, x = fe.sample_rsa_inputs(0xDE03)
    syn = fe.delay_enc(n, e, x, [0, 0], k=16)
    limbs, key, cipher = syn.outputs[:32], syn.outputs[32:34], syn.outputs[34:]
    assert sum(l << (64 * i) for i, l in enumerate(limbs)) == pow(x, e, n)
    # hash input: three 64-bit limbs per element, the last from limbs 30 and 31 (src/lib.rs:229-255); key = words 1, 2
    packed = [limbs[3 * i] + (limbs[3 * i + 1] << 64) + (limbs[3 * i + 2] << 128) for i in range(10)] + [limbs[30] + (limbs[31] << 64)]
    state = [1 << 64, 0, 0, 0, 0]
    for off in (0, 4, 8):
        chunk = packed[off:off + 4]
        for i, v in enumerate(chunk):
            state[1 + i] = (state[1 + i] + v) % po.FR
        if len(chunk) < 4:
            state[1 + len(chunk)] = (state[1 + len(chunk)] + 1) % po.FR
        state = fe.poseidon_permute(state, 5, 8, 57)
    assert key == state[1:3]
    assert cipher == fe.poseidon_encrypt(key, [0, 0]) and fe.poseidon_decrypt(key, cipher) == [0, 0]
    mock_check(syn)
    wit = fe.synthesize(fe.DELAY_ENC, 16, n, e, x, [0, 0], witness_only=True)
    assert (wit.advice == syn.advice).all()
    # rows per exponent bit: README 7 981 ((58 417 - 34 473) / 3, benches/README.md:57-58)
    rows = {b: fe.synthesize(fe.DELAY_ENC, 17, n, 1, x, [0, 0], exp_bits=b, witness_only=True).used_rows for b in (3, 6)}
    assert abs((rows[6] - rows[3]) / 3 - 7981) / 7981 < 0.03


def test_witness_pass_on_several_threads_writes_the_same_rows():
    """de_circuit_desc.threads: the mul_mod row ranges of pow_mod are emitted by worker threads into reserved rows; every cell
    must equal the one-thread pass (and the full synthesis) for any thread count, exponent and a dirty destination buffer"""
    for seed, e_override in ((11, None), (12, 0), (13, 1), (14, 31)):
        n, e, x = fe.sample_rsa_inputs(seed)
        e = e if e_override is None else e_override
        for kind, k in ((fe.DELAY_ENC, 16), (fe.MOD_POW, 17)):
            if kind == fe.MOD_POW and e_override is not None:
                continue
            want = np.empty((5, 1 << k, 4), dtype=np.uint64)
            fe.WitnessPass(kind, k, n=n, e=e, x=x, message=(0, 0)).run(want)
            for threads in (2, 7):
                got = np.full((5, 1 << k, 4), 0x5A5A5A5A5A5A5A5A, dtype=np.uint64)
                wp = fe.WitnessPass(kind, k, n=n, e=e, x=x, message=(0, 0), threads=threads)
                wp.run(got)
                wp.run(got)  # second run: the row count of a range is known, all ranges go to the workers
                assert np.array_equal(got, want), (seed, kind, threads)
    # ... and the keygen-time synthesis (fixed columns and copy constraints as well) fills the advice columns identically
    n, e, x = fe.sample_rsa_inputs(11)
    got = np.empty((5, 1 << 16, 4), dtype=np.uint64)
    fe.WitnessPass(fe.DELAY_ENC, 16, n=n, e=e, x=x, message=(0, 0), threads=5).run(got)
    assert np.array_equal(got, fe.synthesize(fe.DELAY_ENC, 16, n, e, x, [0, 0]).advice)
    # reuse_buffer: a destination that holds an earlier pass of the same circuit (other inputs) is not zeroed and still ends
    # up identical - every pass writes the same cells
    buf = np.full((5, 1 << 16, 4), 0xA5A5, dtype=np.uint64)
    for seed, ex in ((21, 30), (22, 31), (23, 1), (11, None)):
        n, e, x = fe.sample_rsa_inputs(seed)
        e = e if ex is None else ex
        for threads in (1, 6):
            fe.WitnessPass(fe.DELAY_ENC, 16, n=n, e=e, x=x, message=(0, 0), threads=threads).run(buf, reuse=seed != 21 or threads != 1)
    assert np.array_equal(buf, got)
    # the one value the layout depends on: x^0 mod n = 1 is a one-limb constant, 31 rows fewer (its own keys, its own buffers)
    short = [fe.synthesize(fe.DELAY_ENC, 16, n, ee, x, [0, 0], witness_only=True, threads=t).used_rows for ee in (0, e) for t in (1, 4, 4)]
    assert short[0] == short[1] == short[2] == short[3] - 31 and short[3] == short[4] == short[5]
    # not enough rows: reported, not written past the reservation
    with pytest.raises(Exception, match="not enough rows"):
        fe.WitnessPass(fe.DELAY_ENC, 15, n=n, e=3, x=x, message=(0, 0), threads=4).run(np.empty((5, 1 << 15, 4), dtype=np.uint64))


def test_witness_passes_run_concurrently():
    """the bench keeps eight provers per GPU, each with its own witness pass on its own host thread (and, in the
    single-statement arm, worker threads inside the pass): passes share nothing but the process-wide table of region lengths"""
    import threading
    stmts = [fe.sample_rsa_inputs(40 + i) for i in range(6)]
    want = []
    for n, e, x in stmts:
        buf = np.empty((5, 1 << 16, 4), dtype=np.uint64)
        fe.WitnessPass(fe.DELAY_ENC, 16, n=n, e=e, x=x, message=(0, 0)).run(buf)
        want.append(buf)
    got = [np.full((5, 1 << 16, 4), 7, dtype=np.uint64) for _ in stmts]
    errors = []

    def work(i):
        try:
            n, e, x = stmts[i]
            wp = fe.WitnessPass(fe.DELAY_ENC, 16, n=n, e=e, x=x, message=(0, 0), threads=1 + i % 4)
            for r in range(3):
                wp.run(got[i], reuse=r > 0)
        except Exception as ex:  # pragma: no cover
            errors.append(ex)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(stmts))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_rsa_pkcs1_known_answers():
    digest = [(RSA_DIGEST >> (64 * i)) & (2 ** 64 - 1) for i in range(4)]
    ok = fe.rsa_pkcs1(RSA_N1, 65537, RSA_SIG1, digest, k=17)
    assert ok.outputs == [1]
    mock_check(ok)
    assert fe.rsa_pkcs1(RSA_N2, 65537, RSA_SIG2, digest, k=17).outputs == [1]
    assert fe.rsa_pkcs1(RSA_N2, 65537, RSA_SIG2_BAD, digest, k=17).outputs == [0]


def test_errors():
    n, e, x = fe.sample_rsa_inputs(1)
    with pytest.raises(fe.DeError):
        fe.delay_enc(n, e, x, [0, 0], k=12)       # not enough rows
    with pytest.raises(fe.DeError):
        fe.mod_pow(n, 1 << 6, x, k=17)            # e wider than EXP_LIMB_BITS
    with pytest.raises(fe.DeError):
        fe.mod_pow(n, e, n + 5, k=17)             # x >= n: assert_in_field


# /root/reference/src/big_integer/chip.rs:2940-3020 (test_square_circuit): a 1000-bit integer in 64-bit limbs and the 31
# uncarried ("Muled") limb products of its square, as the reference's own test writes them
BIG_SQUARE_A_LIMBS = [4819187580044832333, 9183764011217009606, 11426964127496009747, 17898263845095661790, 12102522037140783322,
                      4029304176671511763, 11339410859987005436, 12120243430436644729, 2888435820322958146, 7612614626488966390,
                      3872170484348249672, 9589147526444685354, 16391157694429928307, 12256166884204507566, 4257963982333550934,
                      916988490704]
BIG_SQUARE_MULED = [int(v) for v in """
    23224568931658367244754058218082222889 88516562921839445888640380379840781596 194478888615417946406783868151393774738
    382395265476432217957523230769986571504 575971019676008360859069855433378813941 670174995752918677131397897218932582682
    780239872348808029089572423614905198300 850410093737715640261630122959874522628 800314959349304909735238452892956199392
    906862855407309870283714027678210238070 967727310654811444144097720329196927129 825671020037461535758117365587238596380
    991281789723902700168027417052185830252 1259367815833216292413970809061165585320 1351495628781923848799708082622582598675
    1451028634949220760698564802414695011932 1290756126635958771067082204577975256756 936482288980049848345464202850902738826
    886330568585033438612679243731110283692 823948310509772835433730556487356331346 649341353489205691855914543942648985328
    497838205323760437611385487609464464168 430091148520710550273018448938020664564 474098876922017329965321439330710234148
    536697574159375092388958994084813127393 483446024935732188792400155524449880972 289799562463011227421662267162524920264
    104372664369829937912234314161010649544 18130279752377737976455635841349605284 7809007931264072381739139035072
    840867892083599894415616""".split()]


def test_bigint_square_known_answer():
    """the big-integer chip's own square test: the chip's limb convolution reproduces the reference's 31 limb products, the
    carry-chain comparison with the constant product accepts, the circuit is satisfied; a product off by one is rejected"""
    a = sum(v << (64 * i) for i, v in enumerate(BIG_SQUARE_A_LIMBS))
    want = sum(v << (64 * i) for i, v in enumerate(BIG_SQUARE_MULED))
    assert len(BIG_SQUARE_MULED) == 31 and want == a * a
    syn = fe.bigint_square(a, want, bits_len=2048, k=14)       # LIMB_WIDTH 64, BITS_LEN 2048 as in the reference's test macro
    assert syn.outputs[0] == 1
    muled = syn.outputs[1:]
    assert len(muled) == 2 * 32 - 1
    assert muled[:31] == BIG_SQUARE_MULED and not any(muled[31:])
    mock_check(syn)
    # a wrong product: the circuit stays satisfiable (is_equal_muled returns a bit), the bit is zero
    bad = fe.bigint_square(a, want + 1, bits_len=2048, k=14)
    assert bad.outputs[0] == 0 and bad.outputs[1:32] == BIG_SQUARE_MULED
    mock_check(bad)
    # a 1024-bit chip has 31 Muled limbs, the 2000-bit product needs 32 carried ones: the reference's assign_constant
    # asserts num_limbs <= max_num_limbs (chip.rs:1270), the front-end reports the same condition as an error
    from de_b200._lib import DeError
    with pytest.raises(DeError, match="too many limbs"):
        fe.bigint_square(a, want, bits_len=1024, k=13)


def _bigint_ops_expected(a, b, n, exp_bits):
    e = b & ((1 << exp_bits) - 1)
    return {"add": a + b, "sub": abs(a - b), "sub_overflow": int(a <= b), "mul_mod": a * b % n, "pow_mod": pow(a, e, n),
            "pow_mod_fixed_exp": pow(a, e, n), "is_equal_fresh": int(a == b), "is_less_than": int(a < b),
            "is_less_than_or_equal": int(a <= b), "in_field": int(a < n)}


@pytest.mark.parametrize("bits_len", [128, 256, 512])
def test_bigint_chip_operators(bits_len):
    """the operator tests of the reference's big-integer chip (src/big_integer/chip.rs:1479-2806: add, sub with and without
    overflow, mul_mod, pow_mod with a variable and a fixed exponent, equality and order comparisons, in-field) against Python
    integers, each circuit checked by the MockProver restatement as the reference's tests do"""
    import random
    rng = random.Random(0xB16 + bits_len)
    n = rng.getrandbits(bits_len) | (1 << (bits_len - 1)) | 1
    cases = []
    for _ in range(2):
        a, b = rng.randrange(n), rng.randrange(n)
        cases += [(a, b), (b, a)]
    a = rng.randrange(n)
    cases += [(a, a), (0, a), (a, 0), (n - 1, n - 1), (1, (1 << 64) - 1), ((1 << 64) - 1, 1 << 64)]
    for i, (a, b) in enumerate(cases):
        syn, got = fe.bigint_ops(a, b, n, bits_len=bits_len, k=14 if bits_len < 512 else 15)
        assert got == _bigint_ops_expected(a, b, n, fe.EXP_LIMB_BITS), (bits_len, i)
        if i < 3 or a == b:
            mock_check(syn)


def test_bigint_chip_operator_errors():
    """mul_mod needs the quotient to fit the limbs of b (the reference's decompose panics); operands wider than the chip are refused"""
    from de_b200._lib import DeError
    with pytest.raises(DeError):
        fe.bigint_ops(1 << 130, 1, (1 << 127) | 1, bits_len=128)
    with pytest.raises(DeError, match="does not fit"):
        fe.bigint_ops((1 << 128) - 1, (1 << 128) - 1, 3, bits_len=128)


def test_bigint_chip_operators_edge_limb_fuzz():
    """the same circuit on operands whose limbs are drawn from {0, 1, 2, 2^32 +- 1, 2^63 +- 1, 2^64 - 1, 2^64 - 2} and random words,
    moduli that are not normalised (top limb small or zero bits below): the carries of add / sub, the quotient estimate of the
    front-end's long division (biguint.hpp, Knuth D) and the carry chain of is_equal_muled at their corner values"""
    import random
    rng = random.Random(12345)
    alpha = [0, 1, 2, 1 << 63, (1 << 64) - 1, (1 << 63) - 1, (1 << 63) + 1, 1 << 32, (1 << 32) - 1, (1 << 64) - 2]

    def pick(nl):
        return sum((rng.choice(alpha) if rng.random() < 0.7 else rng.getrandbits(64)) << (64 * i) for i in range(nl))

    done = 0
    for it in range(160):
        nl = rng.choice([2, 3, 4])
        n = pick(nl)
        if n < 2:
            continue
        a, b = pick(nl) % n, pick(nl) % n
        syn, got = fe.bigint_ops(a, b, n, bits_len=64 * nl, k=14)
        assert got == _bigint_ops_expected(a, b, n, fe.EXP_LIMB_BITS), (hex(a), hex(b), hex(n))
        if it % 40 == 0:
            mock_check(syn)
        done += 1
    assert done > 140


def test_frontend_golden_digests():
    """tests/golden/frontend_v1.json (generator: tests/golden/make_golden_frontend.py): the rows the front-end emits for seeded inputs
    do not move - a proving key made by one build fits the witness of another"""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_frontend", os.path.join(here, "golden", "make_golden_frontend.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    want = json.load(open(os.path.join(here, "golden", "frontend_v1.json")))
    got = {name: gen.describe(syn) for name, syn in gen.cases()}
    assert got == want
