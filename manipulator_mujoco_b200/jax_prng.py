"""Host-side key handling of ``jax.random`` (threefry2x32), so that ``cem_planner.key = PRNGKey(0)``
and the ``key, subkey = jax.random.split(key)`` chain of the reference (``mjx_planner.py:80,314,388``)
produce the key the reference would sample with.  The draws themselves are generated on the GPU
(``cemk_jax_normal``); only the two-word key arithmetic happens here.

Restated from the published algorithm (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
Threefry-2x32 with 20 rounds) and jax/_src/prng.py of the pinned jax==0.5.3 (``requirements.txt:9``),
where ``jax_threefry_partitionable`` is on by default; ``partitionable=False`` gives the older layout.
jax itself cannot be imported here, so this is checked against the known-answer vectors of the
algorithm only (tests/test_jax_prng.py).
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(key, x0, x1):
    """Threefry-2x32-20 of counters (x0, x1) under ``key`` = (k0, k1); uint32 arrays in, two uint32 arrays out."""
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    x0 = np.array(x0, dtype=np.uint32, copy=True)
    x1 = np.array(x1, dtype=np.uint32, copy=True)
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 += x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 ^= x0
            x0 += ks[(g + 1) % 3]
            x1 += ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def PRNGKey(seed):
    """jax.random.PRNGKey(seed) raw key data: (high word, low word) of the 64-bit seed."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def as_key(key):
    """Accept a raw key (2 x uint32) or an integer seed."""
    a = np.asarray(key)
    if a.shape == (2,):
        return a.astype(np.uint32)
    if a.shape == ():
        return PRNGKey(int(a))
    raise ValueError(f"PRNG key must be two uint32 words or an integer seed, got shape {a.shape}")


def split(key, num=2, partitionable=True):
    """jax.random.split(key, num) -> [num, 2] raw keys."""
    key = as_key(key)
    if partitionable:                       # _threefry_split_foldlike: counters (0, i)
        b0, b1 = threefry2x32(key, np.zeros(num, np.uint32), np.arange(num, dtype=np.uint32))
        return np.stack([b0, b1], axis=1)
    # _threefry_split_original: bits of iota(2 num), halves as the two counter words
    b0, b1 = threefry2x32(key, np.arange(num, dtype=np.uint32), np.arange(num, 2 * num, dtype=np.uint32))
    return np.concatenate([b0, b1]).reshape(num, 2)
