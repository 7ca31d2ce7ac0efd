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
