"""Host logic of the two multi-GPU partitionings of SURVEY.md section 8e (no data-path collective in either):

  * batch of independent proofs: proof i runs on rank i mod world_size; the only cross-rank traffic is the timing
    barrier / max reduction;
  * MSM base ranges: rank r commits scalars[lo:hi] against its slice of the resident base tables (de_commit_range) and the
    world_size partial points (96 bytes each) are added (de_g1_sum).
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
