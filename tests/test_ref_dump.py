"""Consumer of the golden vectors a Rust machine produces with integration/dump_golden (tests/golden/REF_DUMP_FORMAT.md): the
staged half of pinned parity (last bullet).  With tests/golden/ref_* present the CPU restatement (and,
under -m gpu, the CUDA path) must reproduce the reference's best_multiexp / best_fft / domain results and the proof BYTES of a
seeded create_proof; without them those tests skip.  The loaders themselves are always exercised: `self_dump` writes a dump of
the same format from the restatement, and the same checks run against it."""
import glob
import json
import os

import numpy as np
import pytest

import orc
import pyoracle as po
import pyprover as pp
from de_b200 import circuits, plonk, serde

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FR = po.FR


# ---- loaders ------------------------------------------------------------------------------------------------------------
def load_msm(path):
    raw = np.fromfile(path, dtype=np.uint8)
    n = int(raw[:8].view(np.uint64)[0])
    body = raw[8:].view(np.uint64)
    scalars, bases, res = body[:4 * n].reshape(n, 4), body[4 * n:12 * n].reshape(n, 8), body[12 * n:12 * n + 8]
    assert body.size == 12 * n + 8
    return scalars, bases, res


def load_fft(path):
    raw = np.fromfile(path, dtype=np.uint8)
    log_n = int(raw[:4].view(np.uint32)[0])
    n = 1 << log_n
    body = raw[4:].view(np.uint64)
    assert body.size == 4 + 8 * n
    return log_n, body[:4], body[4:4 + 4 * n].reshape(n, 4), body[4 + 4 * n:].reshape(n, 4)


def load_ext(path):
    raw = np.fromfile(path, dtype=np.uint8)
    j, k, ek = (int(v) for v in raw[:12].view(np.uint32))
    n, ext = 1 << k, 1 << ek
    pos = 12
    coeffs = raw[pos:pos + 32 * n].view(np.uint64).reshape(n, 4); pos += 32 * n
    extv = raw[pos:pos + 32 * ext].view(np.uint64).reshape(ext, 4); pos += 32 * ext
    m = int(raw[pos:pos + 8].view(np.uint64)[0]); pos += 8
    back = raw[pos:pos + 32 * m].view(np.uint64).reshape(m, 4)
    assert pos + 32 * m == raw.size
    return j, k, ek, coeffs, extv, back


def parse_expr(e):
    t = e[0]
    if t == "const":
        return ("const", int(e[1], 16) % FR)
    if t in ("fixed", "advice", "instance"):
        return (t, int(e[1]), int(e[2]))
    if t == "challenge":
        return ("challenge", int(e[1]))
    if t == "neg":
        return ("neg", parse_expr(e[1]))
    if t in ("sum", "prod"):
        return (t, parse_expr(e[1]), parse_expr(e[2]))
    if t == "scaled":
        return ("scaled", parse_expr(e[1]), int(e[2], 16) % FR)
    raise ValueError(t)


def dump_expr(e):
    t = e[0]
    if t == "const":
        return ["const", hex(e[1])]
    if t in ("fixed", "advice", "instance", "challenge"):
        return list(e)
    if t == "neg":
        return ["neg", dump_expr(e[1])]
    if t in ("sum", "prod"):
        return [t, dump_expr(e[1]), dump_expr(e[2])]
    return ["scaled", dump_expr(e[1]), hex(e[2])]


class RefProof:
    """everything ref_proof/ holds, as the provers take it"""

    def __init__(self, d):
        cs = json.load(open(os.path.join(d, "cs.json")))
        kinds = {"advice": plonk.ADVICE, "fixed": plonk.FIXED, "instance": plonk.INSTANCE}
        self.k = cs["k"]
        self.shape = plonk.ConstraintSystemShape(cs["n_fixed"], cs["n_advice"], cs["n_instance"], [parse_expr(g) for g in cs["gates"]],
                                                 [([parse_expr(e) for e in i], [parse_expr(e) for e in t]) for i, t in cs["lookups"]],
                                                 [(kinds[kd], int(ix)) for kd, ix in cs["permutation"]], cs["blinding_factors"])
        assert self.shape.degree() == cs["degree"], "cs.degree() restated differently from the reference"
        self.queries = pp.Queries(*[[(int(c), int(r)) for c, r in cs[q]] for q in ("advice_queries", "fixed_queries", "instance_queries")])
        self.transcript_repr = int(cs["transcript_repr"], 16)
        self.params = __import__("de_b200").read_params_raw(os.path.join(d, "params.bin"))
        n = 1 << self.k
        vk_size = os.path.getsize(os.path.join(d, "vk.bin"))
        n_sel = (vk_size - 8 - 64 * (self.shape.n_fixed + len(self.shape.perm_columns))) * 8 // n
        self.pk = serde.read_pk(os.path.join(d, "pk.bin"), self.shape, n_selectors=n_sel)
        self.advice = np.fromfile(os.path.join(d, "advice.bin"), dtype=np.uint64).reshape(self.shape.n_advice, n, 4)
        words = np.fromfile(os.path.join(d, "rng_u64.bin"), dtype="<u8")
        assert words.size % 8 == 0, "Fr::random consumes eight u64 per element"
        wide = words.tobytes()
        self.draws = pp.to_mont([int.from_bytes(wide[i:i + 64], "little") % FR for i in range(0, len(wide), 64)])
        self.proof = open(os.path.join(d, "proof.bin"), "rb").read()

    def cpu_keys(self):
        ints = lambda cols: [orc.fr_ints_from_mont(np.ascontiguousarray(c)) for c in cols]
        vk = pp.VerifyingKey(self.k, self.shape, self.queries, [], [], self.transcript_repr % FR)
        opk = pp.ProvingKey(vk, ints(self.pk.fixed_values), ints(self.pk.fixed_polys), ints(self.pk.permutations), ints(self.pk.polys))
        n = 1 << self.k
        oparams = pp.Params(self.k, n, None, None, None, None, np.ascontiguousarray(self.params["g"]), np.ascontiguousarray(self.params["g_lagrange"]))
        return oparams, opk


# ---- a dump of the same format from the restatement (keeps the consumer honest while no Rust machine exists) --------------
def self_dump(d):
    os.makedirs(os.path.join(d, "ref_proof"), exist_ok=True)
    n = 1 << 8
    s, b = orc.uniform_fr(1, n), orc.gen_bases(n)
    res = orc.g1_to_affine(orc.best_multiexp(s, b).reshape(1, 12))[0]
    with open(os.path.join(d, "ref_msm_8.bin"), "wb") as f:
        f.write(np.uint64(n).tobytes() + s.tobytes() + b.tobytes() + res.tobytes())
    dom = orc.Domain(5, 8)
    a = orc.uniform_fr(2, n)
    with open(os.path.join(d, "ref_fft_8.bin"), "wb") as f:
        f.write(np.uint32(8).tobytes() + dom.omega.tobytes() + a.tobytes() + orc.best_fft(a, dom.omega, 8).tobytes())
    ext = dom.coeff_to_extended(a)
    back = dom.extended_to_coeff(dom.divide_by_vanishing(ext))
    with open(os.path.join(d, "ref_ext_5_8.bin"), "wb") as f:
        f.write(np.array([5, 8, dom.extended_k], dtype=np.uint32).tobytes() + a.tobytes() + ext.tobytes() + np.uint64(back.shape[0]).tobytes() + back.tobytes())
    # proof: a satisfied MainGate + RangeChip-shaped circuit at k = 6
    k = 6
    asg = circuits.satisfied_assignment(True, k, 0xD0D0, 40)
    shape = asg.shape
    oparams = pp.setup(k, 0x5EC2E7)
    q = pp.Queries(*plonk.collect_queries(shape))
    repr_ = 0x123456789ABCDEF
    opk = pp.keygen(oparams, shape, q, asg.fixed, asg.copies, repr_)
    pd = os.path.join(d, "ref_proof")
    __import__("de_b200").write_params_raw(os.path.join(pd, "params.bin"), k, oparams.g_mont, oparams.g_lagrange_mont, bytes(128), bytes(128))
    kinds = {plonk.ADVICE: "advice", plonk.FIXED: "fixed", plonk.INSTANCE: "instance"}
    json.dump({"k": k, "n_fixed": shape.n_fixed, "n_advice": shape.n_advice, "n_instance": shape.n_instance, "degree": shape.degree(),
               "blinding_factors": shape.blinding_factors, "gates": [dump_expr(g) for g in shape.gates],
               "lookups": [[[dump_expr(e) for e in i], [dump_expr(e) for e in t]] for i, t in shape.lookups],
               "permutation": [[kinds[kd], ix] for kd, ix in shape.perm_columns],
               "advice_queries": q.advice, "fixed_queries": q.fixed, "instance_queries": q.instance, "transcript_repr": hex(repr_)},
              open(os.path.join(pd, "cs.json"), "w"))
    nn, extn = 1 << k, 1 << orc.Domain(shape.degree(), k).extended_k
    z = lambda *sh: np.zeros(sh, dtype=np.uint64)
    m = lambda cols: np.stack([pp.to_mont(c) for c in cols])
    F, P = shape.n_fixed, len(shape.perm_columns)
    vk = serde.VerifyingKeyRaw(k, z(F, 8), z(P, 8), [np.zeros(nn, bool)] * 2)
    raw = serde.ProvingKeyRaw(vk, z(extn, 4), z(extn, 4), z(extn, 4), m(opk.fixed_values), m(opk.fixed_polys), z(F, extn, 4), m(opk.sigma_values),
                              m(opk.sigma_polys), z(P, extn, 4))
    serde.write_pk(os.path.join(pd, "pk.bin"), raw)
    serde.write_vk(os.path.join(pd, "vk.bin"), vk)
    m(asg.advice).tofile(os.path.join(pd, "advice.bin"))
    # the random stream as 512-bit words whose reduction mod r gives the draws
    rng = po.Xoshiro(0xD1CE)
    count = pp.random_count(shape, nn)
    wide = [rng.next_u64() | (rng.next_u64() << 64) | (rng.next_u64() << 128) | (rng.next_u64() << 192) | (rng.next_u64() << 256) |
            (rng.next_u64() << 320) | (rng.next_u64() << 384) | (rng.next_u64() << 448) for _ in range(count)]
    open(os.path.join(pd, "rng_u64.bin"), "wb").write(b"".join(w.to_bytes(64, "little") for w in wide))
    it = iter([w % FR for w in wide])
    proof = pp.create_proof(oparams, opk, asg.advice, asg.instances, lambda: next(it))
    open(os.path.join(pd, "proof.bin"), "wb").write(proof)
    return d


# ---- the checks -----------------------------------------------------------------------------------------------------------
def check_cpu(d):
    checked = 0
    for path in sorted(glob.glob(os.path.join(d, "ref_msm_*.bin"))):
        s, b, res = load_msm(path)
        got = orc.g1_to_affine(orc.best_multiexp(np.ascontiguousarray(s), np.ascontiguousarray(b)).reshape(1, 12))[0]
        assert (got == res).all(), path
        checked += 1
    for path in sorted(glob.glob(os.path.join(d, "ref_fft_*.bin"))):
        log_n, omega, a, want = load_fft(path)
        assert (orc.best_fft(np.ascontiguousarray(a), np.ascontiguousarray(omega), log_n) == want).all(), path
        checked += 1
    for path in sorted(glob.glob(os.path.join(d, "ref_ext_*.bin"))):
        j, k, ek, coeffs, extv, back = load_ext(path)
        dom = orc.Domain(j, k)
        assert dom.extended_k == ek, path
        ext = dom.coeff_to_extended(np.ascontiguousarray(coeffs))
        assert (ext == extv).all(), path
        assert (dom.extended_to_coeff(dom.divide_by_vanishing(ext)) == back).all(), path
        checked += 1
    if os.path.isdir(os.path.join(d, "ref_proof")):
        rp = RefProof(os.path.join(d, "ref_proof"))
        oparams, opk = rp.cpu_keys()
        assert rp.draws.shape[0] == pp.random_count(rp.shape, 1 << rp.k), "create_proof drew a different number of field elements"
        got = pp.create_proof_fast(oparams, opk, [np.ascontiguousarray(c) for c in rp.advice], [[] for _ in range(rp.shape.n_instance)], rp.draws)
        assert got == rp.proof, "the restated create_proof does not reproduce the reference's proof bytes"
        checked += 1
    return checked


def check_gpu(d):
    import de_b200
    ctx = de_b200.Context(0)
    for path in sorted(glob.glob(os.path.join(d, "ref_msm_*.bin"))):
        s, b, res = load_msm(path)
        assert (ctx.batch_normalize(ctx.best_multiexp(s, b).reshape(1, 12))[0] == res).all(), path
    for path in sorted(glob.glob(os.path.join(d, "ref_fft_*.bin"))):
        log_n, omega, a, want = load_fft(path)
        assert (ctx.best_fft(a, omega, log_n) == want).all(), path
    for path in sorted(glob.glob(os.path.join(d, "ref_ext_*.bin"))):
        j, k, ek, coeffs, extv, back = load_ext(path)
        dom = de_b200.EvaluationDomain(j, k, ctx)
        ext = dom.coeff_to_extended(coeffs)
        assert dom.extended_k == ek and (ext == extv).all(), path
        assert (dom.extended_to_coeff(dom.divide_by_vanishing_poly(ext)) == back).all(), path
        dom.close()
    if os.path.isdir(os.path.join(d, "ref_proof")):
        rp = RefProof(os.path.join(d, "ref_proof"))
        keys = serde.keys_from_proving_key_raw(ctx, rp.pk, rp.shape, np.ascontiguousarray(rp.params["g"]), np.ascontiguousarray(rp.params["g_lagrange"]),
                                               rp.transcript_repr, queries=(rp.queries.advice, rp.queries.fixed, rp.queries.instance))
        proof = keys.prover.create_proof([np.ascontiguousarray(c) for c in rp.advice], [np.zeros((0, 4), dtype=np.uint64)] * rp.shape.n_instance, rp.draws)
        assert proof == rp.proof, "the CUDA prover does not reproduce the reference's proof bytes"
        keys.close()
    ctx.close()


HAVE_REF = bool(glob.glob(os.path.join(GOLD, "ref_*")))


def test_consumer_on_self_made_dump(tmp_path):
    assert check_cpu(self_dump(str(tmp_path))) == 4


@pytest.mark.skipif(not HAVE_REF, reason="no tests/golden/ref_* (produce them with integration/dump_golden on a Rust machine)")
def test_reference_dump_cpu():
    assert check_cpu(GOLD) > 0


@pytest.mark.gpu
def test_consumer_on_self_made_dump_gpu(tmp_path):
    check_gpu(self_dump(str(tmp_path)))


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="no tests/golden/ref_* (produce them with integration/dump_golden on a Rust machine)")
def test_reference_dump_gpu():
    check_gpu(GOLD)
