"""Multi-GPU best_fft parity (GPU): de_ntt_sharded_dev / de_ntt_dist_stage1 / 2 through the C ABI == the oracle's best_fft of the
whole vector, bit-exact.  The ranks are contexts of one process; with a single GPU they all live on device 0 (same kernels, same
peer-store addressing, the "peers" are buffers of the same device), with several GPUs one context per device."""
import ctypes as C

import numpy as np
import pytest

import orc
import pyoracle as po
from de_b200 import DeError, sharding

pytestmark = pytest.mark.gpu


def omega_for(log_n):
    return orc.fr_mont_from_ints([pow(po.FR_ROOT_OF_UNITY, 1 << (po.FR_S - log_n), po.FR)])[0]


def devices(world):
    import torch
    n = torch.cuda.device_count()
    return [r % n for r in range(world)]


@pytest.mark.parametrize("world,log_n", [(1, 11), (1, 16), (2, 12), (2, 15), (2, 19), (2, 21), (2, 22), (4, 13), (4, 18), (4, 22),
                                         (4, 23), (8, 14), (8, 17), (8, 20), (8, 23), (8, 24)])
def test_sharded_best_fft_matches_oracle(world, log_n):
    s = sharding.ShardedNtt(devices(world))
    try:
        a = orc.uniform_fr(0xD157 + 31 * world + log_n, 1 << log_n)
        w = omega_for(log_n)
        got = s.best_fft(a, w, log_n)
        assert (got == orc.best_fft(a, w, log_n)).all()
    finally:
        s.close()


@pytest.mark.parametrize("world,log_n", [(1, 12), (2, 13), (4, 16), (8, 18), (8, 22)])
def test_host_vector_entry_point(world, log_n):
    """de_ntt_sharded: natural-order host vector in and out (block upload, deal kernel, transform, block download)"""
    s = sharding.ShardedNtt(devices(world))
    try:
        a = orc.uniform_fr(0xD200 + world + log_n, 1 << log_n)
        w = omega_for(log_n)
        want = orc.best_fft(a, w, log_n)
        assert (s.best_fft_host(a, w, log_n) == want).all()
        assert (s.best_fft_host(a, w, log_n) == want).all()   # second call reuses every buffer
    finally:
        s.close()


def test_all_on_one_device_and_spread_agree():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    log_n = 18
    a = orc.uniform_fr(5, 1 << log_n)
    w = omega_for(log_n)
    for world in (2, 4, 8):
        if world > torch.cuda.device_count():
            continue
        s = sharding.ShardedNtt(list(range(world)))
        try:
            assert (s.best_fft(a, w, log_n) == orc.best_fft(a, w, log_n)).all()
        finally:
            s.close()


def test_in_place_and_back_to_back_calls():
    """d_out may alias d_x, and consecutive calls need no host synchronisation between them (the stream events order the exchange
    buffers' reuse): transform forward, then the block outputs re-dealt as the cyclic input of the inverse root."""
    import torch
    world, log_n = 4, 16
    n = 1 << log_n
    s = sharding.ShardedNtt(devices(world))
    try:
        a = orc.uniform_fr(9, n)
        b = orc.uniform_fr(10, n)
        w = omega_for(log_n)
        bufs_a = [torch.from_numpy(np.ascontiguousarray(sharding.ntt_input_slice(a, r, world)).view(np.int64)).to(f"cuda:{c.device}")
                  for r, c in enumerate(s.ctxs)]
        bufs_b = [torch.from_numpy(np.ascontiguousarray(sharding.ntt_input_slice(b, r, world)).view(np.int64)).to(f"cuda:{c.device}")
                  for r, c in enumerate(s.ctxs)]
        torch.cuda.synchronize()
        s.best_fft_dev(bufs_a, bufs_a, w, log_n)   # in place
        s.best_fft_dev(bufs_b, bufs_b, w, log_n)   # reuses the exchange buffers right behind it
        s.sync()
        got_a = np.concatenate([t.cpu().numpy().view(np.uint64) for t in bufs_a], axis=0)
        got_b = np.concatenate([t.cpu().numpy().view(np.uint64) for t in bufs_b], axis=0)
        assert (got_a == orc.best_fft(a, w, log_n)).all()
        assert (got_b == orc.best_fft(b, w, log_n)).all()
        # inverse: iNTT(NTT(a)) = N a
        w_inv = orc.fr_inv(w.reshape(1, 4))[0]
        back = s.best_fft(got_a, w_inv, log_n)
        n_f = np.repeat(orc.fr_mont_from_ints([n]), n, axis=0)
        assert (back == orc.fr_mul(a, n_f)).all()
    finally:
        s.close()


def test_linearity_at_sweep_size():
    """size-independent property at a size the oracle is not run on in the suite: NTT(a + b) == NTT(a) + NTT(b), and the result
    equals the single-GPU best_fft of the same library on sampled entries"""
    import de_b200
    world, log_n = 8, 25
    n = 1 << log_n
    s = sharding.ShardedNtt(devices(world))
    try:
        a = orc.uniform_fr(21, n)
        b = orc.uniform_fr(22, n)
        w = omega_for(log_n)
        fa = s.best_fft(a, w, log_n)
        fb = s.best_fft(b, w, log_n)
        fs = s.best_fft(orc.fr_add(a, b), w, log_n)
        idx = np.random.default_rng(1).integers(0, n, 1 << 16)
        assert (fs[idx] == orc.fr_add(np.ascontiguousarray(fa[idx]), np.ascontiguousarray(fb[idx]))).all()
        single = s.ctxs[0].best_fft(a, w, log_n)
        assert (single == fa).all()
    finally:
        s.close()


def test_rejects_bad_arguments():
    s = sharding.ShardedNtt(devices(2))
    try:
        import torch
        t = [torch.zeros(4 << 10, dtype=torch.int64, device="cuda:0") for _ in range(2)]
        w = omega_for(11)
        with pytest.raises(ValueError):
            s.best_fft_dev(t, t, w, 11)          # local transform below 2^11
        c0 = s.ctxs[0]
        ctxs = (C.c_void_p * 3)(*[c0.h] * 3)
        ptrs = (C.c_void_p * 3)(*[t[0].data_ptr()] * 3)
        rc = c0.L.de_ntt_sharded_dev(ctxs, 3, ptrs, ptrs, w.ctypes.data_as(C.c_void_p), 16)
        assert rc == -1 and b"1, 2, 4 or 8" in c0.L.de_last_error(c0.h)
        rc = c0.L.de_ntt_sharded_dev(ctxs, 2, ptrs, ptrs, w.ctypes.data_as(C.c_void_p), 11)
        assert rc == -1
        rc = c0.L.de_ntt_sharded_dev(ctxs, 2, ptrs, ptrs, w.ctypes.data_as(C.c_void_p), 16)   # the same context twice
        assert rc == -1 and b"distinct" in c0.L.de_last_error(c0.h)
        peers = (C.c_void_p * 2)(t[0].data_ptr(), None)
        rc = c0.L.de_ntt_dist_stage1(c0.h, C.c_void_p(t[0].data_ptr()), w.ctypes.data_as(C.c_void_p), 12, 2, 0, peers)
        assert rc == -1
        rc = c0.L.de_ntt_dist_stage2(c0.h, C.c_void_p(t[0].data_ptr()), w.ctypes.data_as(C.c_void_p), 12, 2, 2, peers)
        assert rc == -1
    finally:
        s.close()


def _flag_run(ctxs, xs, zs, outs, flags, omega, log_n, epoch, chunks, ranks):
    world = len(ctxs)
    om0 = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
    for r, c in enumerate(ctxs):  # several ranks of one device driven by one thread: nothing may allocate once a wait is queued
        c.check(c.L.de_ntt_dist_prepare(c.h, om0.ctypes.data_as(C.c_void_p), log_n, world, r))
    zp = (C.c_void_p * world)(*[t.data_ptr() for t in zs])
    op = (C.c_void_p * world)(*[t.data_ptr() for t in outs])
    fp = (C.c_void_p * world)(*[t.data_ptr() for t in flags])
    om = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
    for r in ranks:
        c = ctxs[r]
        c.check(c.L.de_ntt_dist_run(c.h, C.c_void_p(xs[r].data_ptr()), om.ctypes.data_as(C.c_void_p), log_n, world, r, zp, op, fp, epoch, chunks))


@pytest.mark.parametrize("world,log_n,chunks", [(2, 14, 1), (2, 21, 4), (4, 22, 2), (4, 23, 4)])
def test_flag_ordered_run_matches_oracle(world, log_n, chunks):
    """de_ntt_dist_run (the one-process-per-GPU form: stages ordered by flags in peer memory, exchange in `chunks` ranges) with the
    ranks as contexts of this process, ONE PER DEVICE: equals the oracle's best_fft; two calls in a row (epochs 1, 2) reuse
    buffers and flags.  (Several flag-ordered ranks on one device are not supported: a rank's stream parked on an event can hold
    the hardware queue another rank's kernels sit in, and the spinning wait then never sees their signal - that configuration is
    what the event-ordered de_ntt_sharded_dev is for.)"""
    import torch
    import de_b200
    if torch.cuda.device_count() < world:
        pytest.skip("needs one GPU per rank")
    devs = devices(world)
    for a_ in devs:
        for b_ in devs:
            if a_ != b_ and not torch.cuda.can_device_access_peer(a_, b_):
                pytest.skip("no peer access")
    ctxs = [de_b200.Context(d) for d in devs]
    try:
        n, m = 1 << log_n, (1 << log_n) // world
        w = omega_for(log_n)
        mk = lambda d, rows: torch.zeros((rows, 4), dtype=torch.int64, device=f"cuda:{d}")
        for d in devs:  # torch enables peer access between the devices on the first cross-device copy
            mk(d, 1).to(f"cuda:{devs[0]}")
        zs, outs = [mk(d, m) for d in devs], [mk(d, m) for d in devs]
        flags = [torch.zeros(72, dtype=torch.int32, device=f"cuda:{d}") for d in devs]
        for epoch in (1, 2):
            a = orc.uniform_fr(0xF1A6 + epoch + log_n, n)
            xs = [torch.from_numpy(np.ascontiguousarray(sharding.ntt_input_slice(a, r, world)).view(np.int64)).to(f"cuda:{d}") for r, d in enumerate(devs)]
            torch.cuda.synchronize()
            _flag_run(ctxs, xs, zs, outs, flags, w, log_n, epoch, chunks, range(world))
            for c in ctxs:
                c.sync()
            got = np.concatenate([o.cpu().numpy().view(np.uint64) for o in outs], axis=0)
            assert (got == orc.best_fft(a, w, log_n)).all(), epoch
            for c in ctxs:
                v = C.c_int(7)
                c.check(c.L.de_ntt_dist_error(c.h, C.byref(v)))
                assert v.value == 0
    finally:
        for c in ctxs:
            c.close()


def test_flag_wait_gives_up_instead_of_hanging():
    """a rank that never arrives: the waits time out (~2 s each) and de_ntt_dist_error reports it; the device stays usable"""
    import time
    import torch
    import de_b200
    ctxs = [de_b200.Context(0), de_b200.Context(0)]
    try:
        log_n, m = 13, 1 << 12
        w = omega_for(log_n)
        mk = lambda: torch.zeros((m, 4), dtype=torch.int64, device="cuda:0")
        xs, zs, outs = [mk(), mk()], [mk(), mk()], [mk(), mk()]
        flags = [torch.zeros(72, dtype=torch.int32, device="cuda:0") for _ in range(2)]
        t0 = time.time()
        _flag_run(ctxs, xs, zs, outs, flags, w, log_n, 1, 1, [0])      # rank 1 never runs
        v = C.c_int(0)
        ctxs[0].check(ctxs[0].L.de_ntt_dist_error(ctxs[0].h, C.byref(v)))
        assert v.value == 1 and time.time() - t0 < 30
        ctxs[0].check(ctxs[0].L.de_ntt_dist_error(ctxs[0].h, C.byref(v)))
        assert v.value == 0                                           # reading the flag clears it
        # the device is still usable: an ordinary transform on the same context
        a = orc.uniform_fr(0x71AE, 1 << log_n)
        assert (ctxs[0].best_fft(a, w, log_n) == orc.best_fft(a, w, log_n)).all()
    finally:
        for c in ctxs:
            c.close()
