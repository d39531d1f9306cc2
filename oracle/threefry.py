"""NumPy restatement of ``jax.random.uniform(jax.random.PRNGKey(seed), shape)`` (float32), the quantiser noise of
``decode`` (track_autoencoder_3d.py:254-257).  TEST INFRASTRUCTURE ONLY.

jax is an absent third-party dependency (``requirements.txt``: ``jax>=0.4.20``, unpinned), so this follows its published
algorithm (jax/_src/prng.py, jax/_src/random.py):

* ``PRNGKey(seed)`` for the default ``threefry2x32`` implementation is the uint32 pair (seed >> 32, seed & 0xffffffff).
* Threefry-2x32 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): 20 rounds, rotations
  (13, 15, 26, 6), (17, 29, 16, 24), key schedule (k0, k1, k0 ^ k1 ^ 0x1BD11BDA), one key injection every four rounds.
* 32 random bits for ``n`` values, ORIGINAL layout (``jax_threefry_partitionable=False``, the default up to jax 0.4.x):
  counters iota(n), padded to even length, split into halves (x0 = first half, x1 = second half), outputs concatenated.
  PARTITIONABLE layout (default from jax 0.5): the counter of element i is the 64-bit row-major index (hi, lo) and the
  element's bits are out0 ^ out1.
* ``uniform``: float32 in [0, 1) from the top 23 bits: bitcast((bits >> 9) | 0x3F800000) - 1.0.

Pinned by known answers: the Random123 / jax test vectors for the block function and jax's documented
``random.uniform(random.PRNGKey(0)) == 0.41845703`` (tests/test_oracle_model.py).  Which layout a given JAX installation
uses depends on its version and flags - that part cannot be pinned without JAX, so both are provided.
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(key, x0, x1):
    """The Threefry-2x32-20 block function on uint32 arrays x0, x1 with key = (k0, k1)."""
    with np.errstate(over="ignore"):
        k0, k1 = np.uint32(key[0]), np.uint32(key[1])
        ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
        x0 = np.asarray(x0, np.uint32) + ks[0]
        x1 = np.asarray(x1, np.uint32) + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r) ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def prng_key(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def random_bits(key, n, partitionable=False):
    """n uint32 values in the layout ``jax.random.bits(key, (n,))`` produces (flattened row-major for any shape)."""
    if n == 0:
        return np.zeros(0, np.uint32)
    if partitionable:
        idx = np.arange(n, dtype=np.uint64)
        o0, o1 = threefry2x32(key, (idx >> np.uint64(32)).astype(np.uint32), (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32))
        return o0 ^ o1
    if n >= 2 ** 32 - 1:
        raise ValueError("more than 2^32-2 values need jax's multi-block path, which is not restated")
    counts = np.arange(n, dtype=np.uint32)
    if n & 1:
        counts = np.concatenate([counts, np.zeros(1, np.uint32)])   # jax pads an odd count array with a zero
    half = counts.size // 2
    o0, o1 = threefry2x32(key, counts[:half], counts[half:])
    return np.concatenate([o0, o1])[:n]


def uniform(seed, shape, partitionable=False):
    """float32 U[0,1) with the bits of jax.random.uniform(PRNGKey(seed), shape)."""
    n = int(np.prod(shape)) if len(shape) else 1
    bits = random_bits(prng_key(seed), n, partitionable)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(np.float32(0.0), f).reshape(shape)
