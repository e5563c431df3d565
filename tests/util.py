"""Shared helpers for the tests: deterministic inputs and KAT decoding."""
from __future__ import annotations

import hashlib

import numpy as np

MASK = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of SplitMix64 started at `seed` (vectorised; identical on any
    numpy version, unlike Generator streams)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform_below(seed: int, n: int, bound: int) -> np.ndarray:
    """n values in [0, bound) (bound <= 2^64)."""
    r = splitmix64(seed, n)
    if bound >= 1 << 64:
        return r
    return r % np.uint64(bound)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def kat_modulus(spec, gen_primes):
    """q given literally or as {"primes": [bits, ntt_size]}"""
    if isinstance(spec, dict):
        bits, ntt = spec["primes"]
        return gen_primes(1, bits, True, ntt)[0]
    return int(spec)


def kat_values(vals, q):
    """decode lists containing ints and "q-k" strings; a bare int/str stays scalar"""
    def one(v):
        if isinstance(v, str):
            assert v.startswith("q-")
            return q - int(v[2:])
        return int(v)
    if isinstance(vals, list):
        return np.array([one(v) for v in vals], dtype=np.uint64)
    return one(vals)
