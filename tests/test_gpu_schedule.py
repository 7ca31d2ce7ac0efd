"""The create_proof hot-path schedule (de_b200/prover.py) at the three BASELINE circuit shapes, end to end through the C ABI:
every commitment of the schedule equals the oracle's best_multiexp and the quotient equals the oracle's evaluate_h pipeline.
  pose_enc : MainGate only, k = 11 (17 MSMs, 12 + 12 + 1 transforms)      /root/reference/benches/pose_enc.rs:184
  delay_enc: MainGate + RangeChip, k = 16 (31 MSMs, 23 + 23 + 1)          /root/reference/benches/delay_enc.rs:181
  mod_pow  : MainGate + RangeChip, k = 17                                  /root/reference/benches/mod_pow.rs:258
"""
import numpy as np
import pytest
import torch

import de_b200
import orc
from de_b200 import plonk, prover

pytestmark = pytest.mark.gpu


def make_inputs(shape, k, seed, used_rows):
    w = prover.Workload(shape, k)
    n, o = w.n, w.offsets()
    cols = np.empty((w.n_cols, n, 4), dtype=np.uint64)
    for i in range(shape.n_advice):
        cols[o["advice"] + i] = orc.witness_fr(seed + i, n, used_rows)
    cols[o["instance"]] = 0
    for i in range(o["permz"], w.n_cols):
        cols[i] = orc.uniform_fr(seed + 100 + i, n)
    return dict(w=w, cols=cols, random=orc.uniform_fr(seed + 200, n).reshape(1, n, 4),
                openings=orc.uniform_fr(seed + 201, n * w.n_openings).reshape(w.n_openings, n, 4),
                fixed=[orc.uniform_fr(seed + 300 + i, n) for i in range(shape.n_fixed)],
                sigma=[orc.uniform_fr(seed + 400 + i, n) for i in range(len(shape.perm_columns))],
                g=orc.gen_bases(n), g_lagrange=orc.gen_bases(n, start=n), ch=(0x1234 + k, 0x5678, 0x9ABC, 0xDEF0))


def oracle_schedule(shape, k, inp):
    w = inp["w"]
    n, o, L = w.n, w.offsets(), w.n_lookups
    dom = orc.Domain(shape.degree(), k)
    desc, keep = plonk.marshal_pk_desc(shape, inp["fixed"], inp["sigma"])
    pk = orc.Pk(dom, desc, keep)
    chs, keep2 = plonk.marshal_challenges(*inp["ch"])
    cols = inp["cols"]
    pts = [orc.best_multiexp(cols[o["advice"] + i], inp["g_lagrange"]) for i in range(shape.n_advice)]
    pts += [orc.best_multiexp(cols[o["lookup_a"] + i], inp["g_lagrange"]) for i in range(2 * L)]
    pts += [orc.best_multiexp(cols[o["permz"] + i], inp["g_lagrange"]) for i in range(shape.n_perm_sets + L)]
    pts.append(orc.best_multiexp(inp["random"][0], inp["g"]))
    coeff = [dom.lagrange_to_coeff(cols[i]) for i in range(w.n_cols)]
    h = pk.evaluate_h(coeff[:shape.n_advice], coeff[o["instance"]:o["instance"] + shape.n_instance], chs,
                      coeff[o["permz"]:o["permz"] + shape.n_perm_sets], coeff[o["lookup_z"]:])
    hc = dom.extended_to_coeff(dom.divide_by_vanishing(h))
    pts += [orc.best_multiexp(np.ascontiguousarray(hc[i * n:(i + 1) * n]), inp["g"]) for i in range(shape.degree() - 1)]
    pts += [orc.best_multiexp(inp["openings"][i], inp["g"]) for i in range(w.n_openings)]
    return np.stack(pts), hc


@pytest.mark.parametrize("name,with_lookups,k,used", [("pose_enc", False, 11, 1450), ("delay_enc", True, 16, 50400),
                                                      ("mod_pow", True, 17, 41766), ("tiny", True, 6, 40)])
def test_hot_path_schedule_matches_oracle(name, with_lookups, k, used):
    shape = plonk.main_gate_shape(with_lookups)
    inp = make_inputs(shape, k, 0xDE00 + k, used)
    w = inp["w"]
    assert w.n_msm == (31 if with_lookups else 17)
    assert w.n_cols == (23 if with_lookups else 12)
    want_pts, want_hc = oracle_schedule(shape, k, inp)
    stream = torch.cuda.Stream()
    ctx = de_b200.Context(0)
    ctx.set_stream(stream.cuda_stream)
    as_i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64))
    with torch.cuda.stream(stream):
        hp = prover.HotPathProver(ctx, w, inp["g"], inp["g_lagrange"], inp["fixed"], inp["sigma"])
        cols_d, random_d, open_d = as_i64(inp["cols"]).cuda(), as_i64(inp["random"]).cuda(), as_i64(inp["openings"]).cuda()
        got = hp.prove_dev(cols_d, random_d, open_d, inp["ch"])
        stream.synchronize()
        assert got.shape == (w.n_msm, 12)
        assert (ctx.batch_normalize(got) == orc.g1_to_affine(want_pts)).all(), name
        # the quotient polynomial left in the prover's buffer: (d - 1) * n coefficients
        m = (shape.degree() - 1) * w.n
        got_h = hp._h.cpu().numpy().view(np.uint64)[:m]
        assert (got_h == want_hc).all()
        # host-buffer path gives the same commitments
        staging = {"cols": torch.empty_like(cols_d), "random": torch.empty_like(random_d), "openings": torch.empty_like(open_d)}
        got2 = hp.prove_host(as_i64(inp["cols"]).pin_memory(), as_i64(inp["random"]).pin_memory(), as_i64(inp["openings"]).pin_memory(),
                             inp["ch"], staging)
        assert (ctx.batch_normalize(got2) == orc.g1_to_affine(want_pts)).all()
        hp.close()
    ctx.close()
