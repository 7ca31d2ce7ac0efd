"""Host logic of the two multi-GPU partitionings of SURVEY.md section 8e (no data-path collective in either):

  * batch of independent proofs: proof i runs on rank i mod world_size; the only cross-rank traffic is the timing
    barrier / max reduction;
  * MSM base ranges: rank r commits scalars[lo:hi] against its slice of the resident base tables (de_commit_range) and the
    world_size partial points (96 bytes each) are added (de_g1_sum);
  * one NTT vector over W = 2, 4 or 8 GPUs (the one row of section 8e with a real exchange): four-step transform, cyclic input
    slices, contiguous natural-order output blocks, both transposes done as peer-memory stores from inside the kernels
    (de_ntt_dist_stage1 / 2) - ShardedNtt inside one process, DistNtt with one process per GPU (CUDA IPC + a stream barrier).
"""
from __future__ import annotations


def proofs_for_rank(n_proofs: int, rank: int, world: int):
    return list(range(rank, n_proofs, world))


def base_range(n: int, rank: int, world: int):
    """contiguous, disjoint, covering [0, n): the reference's own per-thread chunking (multiexp chunks of n / threads)"""
    chunk = -(-n // world)
    lo = min(n, rank * chunk)
    return lo, min(n, lo + chunk)


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_partials(partial, group=None):
    """All ranks contribute one Jacobian point (12 x u64, host numpy) and receive all of them: (world, 12).  96 bytes per
    rank over NCCL (NVLink) or gloo; this is the only exchange of the base-range sharded MSM."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(partial, dtype=np.uint64).reshape(1, 12)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64).reshape(12)).to(dev)
    out = [torch.empty(12, dtype=torch.int64, device=dev) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return torch.stack(out).cpu().numpy().view(np.uint64)


def sharded_commit(params_shard, basis: int, d_scalars_shard, n_shard: int, group=None):
    """ParamsKZG::commit with the bases sharded by contiguous ranges: `params_shard` holds this rank's base range
    (base_range(n, rank, world)) and `d_scalars_shard` the matching scalars.  Returns the full commitment (Jacobian, (12,))
    on every rank: local MSM -> all-gather of the partial points -> sum."""
    partial = params_shard.commit_batch_dev(basis, d_scalars_shard, n_shard, 1)[0]
    parts = gather_partials(partial, group)
    return params_shard.ctx.g1_sum(parts)


class ShardedParams:
    """One process, several GPUs: the SRS basis split into contiguous ranges, one ParamsKZG (bases + window tables) per device.
    commit() runs de_commit_sharded: one host thread per GPU, partial points summed on the first one."""

    def __init__(self, k: int, bases, devices, basis: int = 1):
        import numpy as np
        from . import Context, ParamsKZG
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        n = 1 << k
        assert bases.shape[0] == n
        self.basis, self.n = basis, n
        self.ctxs, self.shards, self.lo, self.len = [], [], [], []
        for r, dev in enumerate(devices):
            lo, hi = base_range(n, r, len(devices))
            m = hi - lo
            ks = max(1, (m - 1).bit_length())
            pad = np.zeros((1 << ks, 8), dtype=np.uint64)
            pad[:m] = bases[lo:hi]
            ctx = Context(dev)
            self.ctxs.append(ctx)
            self.shards.append(ParamsKZG(ks, pad if basis == 0 else None, pad if basis == 1 else None, ctx))
            self.lo.append(lo)
            self.len.append(m)

    def commit(self, scalars):
        import ctypes as C
        import numpy as np
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        assert scalars.size == 4 * self.n
        ns = len(self.shards)
        arr = (C.c_void_p * ns)(*[s.h for s in self.shards])
        lo = (C.c_size_t * ns)(*self.lo)
        ln = (C.c_size_t * ns)(*self.len)
        out = np.zeros(12, dtype=np.uint64)
        c0 = self.ctxs[0]
        c0.check(c0.L.de_commit_sharded(arr, lo, ln, ns, self.basis, scalars.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        for s in self.shards:
            s.close()
        for c in self.ctxs:
            c.close()


# ---- one NTT vector over several GPUs -----------------------------------------------------------------------------------
def ntt_layout(log_n: int, world: int):
    """(M, C): elements per rank and columns per (source rank, destination rank) pair of the exchange"""
    lw = world.bit_length() - 1
    if world < 1 or (1 << lw) != world or world > 8:
        raise ValueError("world must be 1, 2, 4 or 8")
    if log_n < 11 + lw or log_n > 28:
        raise ValueError("need 11 + log2(world) <= log_n <= 28")
    m = (1 << log_n) >> lw
    return m, m >> lw


def ntt_input_slice(a, rank: int, world: int):
    """rank's input of the multi-GPU transform: a[rank], a[rank + W], a[rank + 2 W], ...  (a: (N, 4) u64)"""
    return a[rank::world]


def ntt_output_range(log_n: int, rank: int, world: int):
    """rank's output block [lo, hi) of best_fft(a)"""
    m, _ = ntt_layout(log_n, world)
    return rank * m, (rank + 1) * m


def ntt_exchange_slot(log_n: int, world: int, src_rank: int, column: int):
    """where stage 1 on `src_rank` stores column j2 of its local transform: (destination rank, index in its exchange buffer)"""
    _, c = ntt_layout(log_n, world)
    return column // c, src_rank * c + column % c


class ShardedNtt:
    """best_fft of one vector over several GPUs driven by ONE process (de_ntt_sharded_dev): one Context per entry of `devices`
    (the same device may appear several times - used by the single-GPU parity tests)."""

    def __init__(self, devices):
        from . import Context
        self.ctxs = [Context(d) for d in devices]
        self.world = len(devices)

    def best_fft_dev(self, d_x, d_out, omega, log_n: int):
        """d_x[r] / d_out[r]: torch int64 tensors on rank r's device holding the cyclic input slice / receiving the output block.
        Asynchronous on the contexts' streams; sync() waits."""
        import ctypes as C
        import numpy as np
        w = len(self.ctxs)
        ntt_layout(log_n, w)
        om = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
        ctxs = (C.c_void_p * w)(*[c.h for c in self.ctxs])
        xs = (C.c_void_p * w)(*[t.data_ptr() for t in d_x])
        outs = (C.c_void_p * w)(*[t.data_ptr() for t in d_out])
        c0 = self.ctxs[0]
        c0.check(c0.L.de_ntt_sharded_dev(ctxs, w, xs, outs, om.ctypes.data_as(C.c_void_p), log_n))

    def sync(self):
        for c in self.ctxs:
            c.sync()

    def best_fft_host(self, a, omega, log_n: int):
        """arithmetic::best_fft on a host vector, natural order in and out, through de_ntt_sharded (the library deals the blocks to
        the cyclic slices on the devices; needs one context per rank, which this class always has)"""
        import ctypes as C
        import numpy as np
        a = np.array(a, dtype=np.uint64, order="C").reshape(-1, 4)
        if a.shape[0] != 1 << log_n:
            raise ValueError("best_fft: a.len() != 1 << log_n")
        w = len(self.ctxs)
        ntt_layout(log_n, w)
        om = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
        ctxs = (C.c_void_p * w)(*[c.h for c in self.ctxs])
        c0 = self.ctxs[0]
        c0.check(c0.L.de_ntt_sharded(ctxs, w, a.ctypes.data_as(C.c_void_p), om.ctypes.data_as(C.c_void_p), log_n))
        return a

    def best_fft(self, a, omega, log_n: int):
        """host vector in natural order in and out, dealing the cyclic slices on the host (exercises de_ntt_sharded_dev)"""
        import numpy as np
        import torch
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        if a.shape[0] != 1 << log_n:
            raise ValueError("best_fft: a.len() != 1 << log_n")
        w = self.world
        xs = [torch.from_numpy(np.ascontiguousarray(ntt_input_slice(a, r, w)).view(np.int64)).to(f"cuda:{c.device}")
              for r, c in enumerate(self.ctxs)]
        outs = [torch.empty_like(x) for x in xs]
        for c in self.ctxs:
            torch.cuda.synchronize(c.device)
        self.best_fft_dev(xs, outs, omega, log_n)
        self.sync()
        return np.concatenate([o.cpu().numpy().view(np.uint64) for o in outs], axis=0)

    def close(self):
        for c in self.ctxs:
            c.close()


class DistNtt:
    """The same transform with one process per GPU (torch.distributed for the one-time set-up only): every rank allocates its
    input, exchange and output buffers and a few flag words with de_dev_alloc, the 64-byte CUDA IPC handles are all-gathered
    once, and each call is ONE de_ntt_dist_run per rank: the peer-store pass in `chunks` ranges, each followed by a flag store
    to every rank; the cross stage of a range starts on the context's second stream as soon as every rank has signalled that
    range; a last flag round makes the call complete in stream order.  No library collective is on the data path or between the
    stages.  chunks = 1 (default) orders the two stages with flags only; 2 / 4 pipeline the cross stage under the later ranges
    of the pass, which measured slower on 8 B200s (2^27: 4.42 / 4.49 / 4.67 ms for 1 / 2 / 4 ranges: both stages share the
    NVLink egress and the pass' CTAs fill the register file).  pipelined=False keeps round 1's form (stage 1 -> 1-element NCCL
    all-reduce -> stage 2 -> all-reduce; 4.46 ms) for comparison."""

    FLAG_BYTES = 288  # DE_NTT_DIST_FLAG_BYTES

    def __init__(self, ctx, log_n: int, group=None, pipelined: bool = True, chunks: int = 1):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.ctx, self.log_n, self.group = ctx, log_n, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.m, self.c = ntt_layout(log_n, self.world)
        self.stream = torch.cuda.Stream(device=ctx.device)
        ctx.set_stream(self.stream.cuda_stream)
        self.token = torch.zeros(1, dtype=torch.int32, device=f"cuda:{ctx.device}")
        self.pipelined, self.chunks, self.epoch = pipelined, chunks, 0
        L = ctx.L
        self.own, self.mapped = [], []
        peers, err = [], None
        for which in range(4):  # input, exchange, output, flags
            p = C.c_void_p()
            h = (C.c_uint8 * 64)()
            try:
                ctx.check(L.de_dev_alloc(ctx.h, 32 * self.m if which < 3 else self.FLAG_BYTES, C.byref(p)))
                if which == 3:
                    zero = torch.zeros(self.FLAG_BYTES, dtype=torch.uint8, device=f"cuda:{ctx.device}")
                    ctx.check(L.de_dev_copy(ctx.h, C.c_void_p(p.value), C.c_void_p(zero.data_ptr()), self.FLAG_BYTES))
                    ctx.sync()
                self.own.append(p.value)
                ctx.check(L.de_ipc_export(ctx.h, C.c_void_p(p.value), h))
            except Exception as e:  # keep the collectives below matched on every rank, fail together afterwards
                err = err or e
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(h), group=group)
            ptrs = []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    ptrs.append(p.value)
                    continue
                q = C.c_void_p()
                try:
                    if err is None:
                        ctx.check(L.de_ipc_import(ctx.h, (C.c_uint8 * 64).from_buffer_copy(hb), C.byref(q)))
                        self.mapped.append(q.value)
                except Exception as e:
                    err = err or e
                ptrs.append(q.value)
            peers.append(ptrs)
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=f"cuda:{ctx.device}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self._release(lambda: dist.barrier(group=group))
            raise RuntimeError(f"DistNtt: mapping the ranks' buffers failed on at least one rank ({err or 'another rank'})")
        self.d_x, self.d_z, self.d_out, self.d_flags = self.own
        self.z_peers = (C.c_void_p * self.world)(*peers[1])
        self.out_peers = (C.c_void_p * self.world)(*peers[2])
        self.flag_peers = (C.c_void_p * self.world)(*peers[3])
        dist.barrier(group=group)  # every rank's flag words are zero before anybody signals

    def _barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self.token, group=self.group)

    def run(self, omega):
        """transforms d_x (this rank's cyclic slice) into d_out (this rank's block); asynchronous on self.stream"""
        import ctypes as C
        import numpy as np
        import torch
        om = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
        ctx, L = self.ctx, self.ctx.L
        if self.pipelined:
            self.epoch += 1
            ctx.check(L.de_ntt_dist_run(ctx.h, C.c_void_p(self.d_x), om.ctypes.data_as(C.c_void_p), self.log_n, self.world, self.rank,
                                        self.z_peers, self.out_peers, self.flag_peers, self.epoch, self.chunks))
            return
        with torch.cuda.stream(self.stream):
            ctx.check(L.de_ntt_dist_stage1(ctx.h, C.c_void_p(self.d_x), om.ctypes.data_as(C.c_void_p), self.log_n, self.world, self.rank,
                                           self.z_peers))
            self._barrier()
            ctx.check(L.de_ntt_dist_stage2(ctx.h, C.c_void_p(self.d_z), om.ctypes.data_as(C.c_void_p), self.log_n, self.world, self.rank,
                                           self.out_peers))
            self._barrier()

    def timed_out(self) -> bool:
        """True when a flag wait gave up (a rank never arrived) since the last call; synchronises the stream"""
        import ctypes as C
        v = C.c_int(0)
        self.ctx.check(self.ctx.L.de_ntt_dist_error(self.ctx.h, C.byref(v)))
        return bool(v.value)

    def _release(self, group_barrier=None):
        import ctypes as C
        for q in self.mapped:
            self.ctx.L.de_ipc_release(self.ctx.h, C.c_void_p(q))
        self.mapped = []
        if group_barrier:
            group_barrier()  # an exporter frees only after every importer has unmapped
        for p in self.own:
            self.ctx.L.de_dev_free(self.ctx.h, C.c_void_p(p))
        self.own = []
        self.ctx.set_stream(None)

    def close(self):
        import torch.distributed as dist
        self.stream.synchronize()
        dist.barrier(group=self.group)  # nobody unmaps or frees while a peer may still be storing into these buffers
        self._release(lambda: dist.barrier(group=self.group))
