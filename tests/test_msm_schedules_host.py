"""Counting-sort front end of the MSM (csrc/msm.cuh: k_msm_digits, k_msm_scatter) restated lane by lane on the CPU: the signed
c-bit digit decomposition, the bucket histogram whose atomics are aggregated per warp where keys come in runs, and the rank
scatter that consumes the atomics' return values.  What must hold for ANY interleaving of the atomics:
  * sum_w digit_w * 2^(c w) == the scalar (digits in [-2^(c-1), 2^(c-1)], top carry absorbed by the last window);
  * inside every bucket the ranks handed out are a permutation of 0 .. count - 1, so the scatter writes every slot of the
    bucket's range exactly once;
  * aggregated and plain atomics produce the same histogram.
The kernels themselves are checked against the oracle's best_multiexp on the GPU (tests/test_gpu_msm.py); this file pins the
index logic (dead lanes of a ragged last warp, zero digits, runs of equal scalars) where no GPU is needed."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import pyoracle as po  # noqa: E402

INVALID = 0xFFFFFFFF


def signed_digits(s, c, W):
    """k_msm_digits: window w takes c bits at offset c*w plus the carry; v > 2^(c-1) becomes v - 2^c with a carry out"""
    half, out, carry = 1 << (c - 1), [], 0
    for w in range(W):
        v = ((s >> (c * w)) & ((1 << c) - 1)) + carry
        carry = 0
        if v > half:
            out.append(v - (1 << c))
            carry = 1
        else:
            out.append(v)
    assert carry == 0, "the top window must absorb the carry (c * W >= 255 for a 254-bit scalar)"
    return out


@pytest.mark.parametrize("c", [8, 11, 14, 15, 16, 17, 20])
def test_signed_digits_recompose(c):
    W = (255 + c - 1) // c
    rng = po.Xoshiro(0xD161 + c)
    vals = [rng.uniform_fr() for _ in range(200)] + [0, 1, po.FR - 1, (1 << (c - 1)), (1 << (c - 1)) + 1, (1 << c) - 1, (1 << 253) + 12345]
    for s in vals:
        d = signed_digits(s, c, W)
        assert all(-(1 << (c - 1)) <= x <= (1 << (c - 1)) for x in d)
        assert sum(x << (c * w) for w, x in enumerate(d)) == s


def histogram_with_ranks(keys_per_lane, aggregate, seed):
    """One window of k_msm_digits<AGG> over consecutive warps of 32 lanes (keys_per_lane: one key per scalar, INVALID for a zero
    digit; the last warp is padded with dead lanes).  Atomics of different warps interleave in a random order; inside a warp the
    plain form serialises its lanes in a random order too.  Returns (counts, ranks)."""
    rnd = random.Random(seed)
    n = len(keys_per_lane)
    warps = [list(range(w0, min(w0 + 32, n))) for w0 in range(0, n, 32)]
    counts, ranks = {}, [None] * n
    order = list(range(len(warps)))
    rnd.shuffle(order)
    for wi in order:
        lanes = warps[wi]
        keys = [keys_per_lane[i] for i in lanes] + [INVALID] * (32 - len(lanes))  # dead lanes carry no key
        runs = aggregate and any(keys[l] == keys[l + 1] and keys[l] != INVALID for l in range(31))
        if runs:
            for key in set(keys):
                if key == INVALID:
                    continue
                peers = [l for l in range(32) if keys[l] == key]
                first = counts.get(key, 0)  # the leader's single atomicAdd(popc(peers))
                counts[key] = first + len(peers)
                for pos, l in enumerate(peers):
                    if l < len(lanes):
                        ranks[lanes[l]] = first + pos  # first + popc(peers below this lane)
        else:
            lane_order = list(range(len(lanes)))
            rnd.shuffle(lane_order)
            for l in lane_order:
                key = keys[l]
                if key != INVALID:
                    ranks[lanes[l]] = counts.get(key, 0)
                    counts[key] = counts.get(key, 0) + 1
    return counts, ranks


def column(kind, n, rng):
    if kind == "uniform":
        return [rng.next_u64() % 32768 for _ in range(n)]
    if kind == "constant_tail":  # a permutation grand product behind the used rows
        head = [rng.next_u64() % 32768 for _ in range(n // 3)]
        return head + [4242] * (n - len(head))
    if kind == "bits":  # a witness column of 0 / 1: zero digits produce no entry
        return [INVALID if rng.next_u64() & 1 else 0 for _ in range(n)]
    return [INVALID] * n  # "empty"


@pytest.mark.parametrize("kind", ["uniform", "constant_tail", "bits", "empty"])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097])
def test_ranks_are_a_permutation_inside_every_bucket(kind, n):
    rng = po.Xoshiro(0xA66 + n)
    keys = column(kind, n, rng)
    plain_counts, _ = histogram_with_ranks(keys, False, 1)
    for seed in (2, 3):
        counts, ranks = histogram_with_ranks(keys, True, seed)
        assert counts == plain_counts
        per_bucket = {}
        for i, key in enumerate(keys):
            if key == INVALID:
                assert ranks[i] is None
            else:
                per_bucket.setdefault(key, []).append(ranks[i])
        for key, r in per_bucket.items():
            assert sorted(r) == list(range(counts[key])), key
        # the scatter: offsets = exclusive scan of the histogram; every slot of sorted[] is written exactly once
        offsets, run = {}, 0
        for key in sorted(counts):
            offsets[key] = run
            run += counts[key]
        slots = sorted(offsets[key] + ranks[i] for i, key in enumerate(keys) if key != INVALID)
        assert slots == list(range(run))
