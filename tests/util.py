"""Shared helpers for the test-suite (seeded inputs, limb conversions)."""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


def int_to_limbs(v: int, n: int = 4):
    return [(v >> (64 * i)) & MASK64 for i in range(n)]


def limbs_to_int(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def ints_to_array(vals, n: int = 4) -> np.ndarray:
    return np.array([int_to_limbs(v, n) for v in vals], dtype=np.uint64).reshape(len(vals), n)


def random_scalars(n: int, modulus: int, seed: int) -> np.ndarray:
    """Uniform in [0, modulus) by mask-and-reject (how ark's UniformRand samples), seeded."""
    rng = np.random.default_rng(seed)
    bits = modulus.bit_length()
    top_mask = np.uint64((1 << (bits - 192)) - 1)
    mod_limbs = np.array(int_to_limbs(modulus), dtype=np.uint64)
    out = np.zeros((n, 4), dtype=np.uint64)
    todo = np.arange(n)
    while todo.size:
        cand = rng.integers(0, 1 << 64, size=(todo.size, 4), dtype=np.uint64)
        cand[:, 3] &= top_mask
        # lexicographic compare from the top limb: cand < modulus
        lt = np.zeros(todo.size, dtype=bool)
        eq = np.ones(todo.size, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (cand[:, k] < mod_limbs[k])
            eq &= cand[:, k] == mod_limbs[k]
        out[todo[lt]] = cand[lt]
        todo = todo[~lt]
    return out


def scalars_to_ints(arr: np.ndarray):
    return [limbs_to_int(row) for row in arr]
