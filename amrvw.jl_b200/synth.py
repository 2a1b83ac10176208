"""Reproducible synthetic inputs: counter-based, language independent.

u(seed, poly, k) = top 53 bits of splitmix64(seed ^ splitmix64(poly * GOLDEN + k)) / 2^53;
kind 0: N(0,1) by Box-Muller on (u_{2k}, u_{2k+1}); kind 1: U[0,1); kind 2: U[0,1) - 1/2.
"""
import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + _G
        x = (x ^ (x >> np.uint64(30))) * _M1
        x = (x ^ (x >> np.uint64(27))) * _M2
        return x ^ (x >> np.uint64(31))


def uniform01(seed, poly, k):
    poly = np.asarray(poly, dtype=np.uint64)
    k = np.asarray(k, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) ^ _splitmix64(poly * _G + k))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def fill_coeffs(seed, poly_ids, kind, count):
    """-> float64 array [len(poly_ids), count]."""
    poly_ids = np.atleast_1d(np.asarray(poly_ids, dtype=np.uint64))[:, None]
    k = np.arange(count, dtype=np.uint64)[None, :]
    if kind == 0:
        u1 = np.maximum(uniform01(seed, poly_ids, np.uint64(2) * k), 1e-300)
        u2 = uniform01(seed, poly_ids, np.uint64(2) * k + np.uint64(1))
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    u = uniform01(seed, poly_ids, k)
    return u if kind == 1 else u - 0.5


def shard_range(nbatch, rank, world):
    """Contiguous split of the batch over ranks: polynomials never interact, so the
    only multi-GPU step is this partition plus a host gather."""
    base, rem = divmod(int(nbatch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
