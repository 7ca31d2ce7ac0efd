"""Shared helpers for the tests: golden-vector loading and conversions between canonical ints and Montgomery limbs."""
import json
import os

import numpy as np

import orc
import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with open(os.path.join(HERE, "golden", "golden_v1.json")) as f:
        return json.load(f)


def fr(vals):
    """canonical ints / hex strings -> Montgomery (n, 4) u64"""
    return orc.fr_mont_from_ints([int(v, 16) if isinstance(v, str) else int(v) for v in vals])


def fr_ints(a):
    return orc.fr_ints_from_mont(np.ascontiguousarray(a))


def points(pairs):
    """[(x, y) | None] canonical -> affine Montgomery (n, 8) u64; None -> identity (all zero)"""
    flat = []
    for p in pairs:
        if p is None:
            flat += [0, 0]
        else:
            flat += [int(p[0], 16) if isinstance(p[0], str) else p[0], int(p[1], 16) if isinstance(p[1], str) else p[1]]
    m = orc.fq_mont_from_ints(flat).reshape(-1, 8)
    for i, p in enumerate(pairs):
        if p is None:
            m[i] = 0
    return m


def affine_ints(jac):
    """(12,) Jacobian Montgomery -> (x, y) canonical ints or None"""
    a = orc.g1_to_affine(np.ascontiguousarray(jac).reshape(1, 12))[0]
    if not a.any():
        return None
    v = orc.fq_ints_from_mont(a.reshape(2, 4))
    return (v[0], v[1])
