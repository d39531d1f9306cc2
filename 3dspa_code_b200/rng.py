"""The quantiser noise of ``decode`` (track_autoencoder_3d.py:254-257): ``jax.random.uniform(jax.random.PRNGKey(0), shape)``.

The reference draws the SAME noise tensor on every call (the key is a constant), so reproducing it makes
``model.apply(variables, batch)`` deterministic and equal to the reference's default behaviour without passing ``noise``.
Host-side NumPy (12 288 values per clip); Threefry-2x32-20 counter-based generator as published (Salmon et al., SC'11) with
jax's two counter layouts:

* ``original``       - ``jax_threefry_partitionable=False`` (default up to jax 0.4.x): counters iota(n), zero-padded to even
                       length, split in halves, outputs concatenated;
* ``partitionable``  - default from jax 0.5: element i uses the 64-bit row-major index as its counter, bits = out0 ^ out1.

Known answers (tests/test_abi.py): Random123 block vectors, ``uniform(PRNGKey(0), ()) = 0.41845703`` and the first ten
``random.normal(PRNGKey(0), (10,))`` values of the jax quickstart for both layouts.
"""
from __future__ import annotations

import numpy as np

_R = (13, 15, 26, 6, 17, 29, 16, 24)
_PARITY = 0x1BD11BDA


def threefry2x32(k0, k1, c0, c1):
    """20-round Threefry on uint32 counter arrays (c0, c1) under key (k0, k1)."""
    m = np.uint64(0xFFFFFFFF)
    k = [np.uint64(k0), np.uint64(k1), np.uint64(int(k0) ^ int(k1) ^ _PARITY)]
    a = (np.asarray(c0, np.uint64) + k[0]) & m
    b = (np.asarray(c1, np.uint64) + k[1]) & m
    for rnd in range(20):
        r = np.uint64(_R[(rnd % 4) + 4 * ((rnd // 4) % 2)])
        a = (a + b) & m
        b = (((b << r) | (b >> (np.uint64(32) - r))) & m) ^ a
        if rnd % 4 == 3:
            inj = rnd // 4 + 1
            a = (a + k[inj % 3]) & m
            b = (b + k[(inj + 1) % 3] + np.uint64(inj)) & m
    return a.astype(np.uint32), b.astype(np.uint32)


def bits(seed, n, layout="original"):
    k0, k1 = (int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF
    if layout == "partitionable":
        i = np.arange(n, dtype=np.uint64)
        o0, o1 = threefry2x32(k0, k1, i >> np.uint64(32), i & np.uint64(0xFFFFFFFF))
        return o0 ^ o1
    if layout != "original":
        raise ValueError(f"unknown threefry layout {layout!r}")
    if n >= 2 ** 32 - 1:
        raise ValueError("too many values for one threefry block sequence")
    c = np.arange(n + (n & 1), dtype=np.uint64)
    if n & 1:
        c[-1] = 0
    h = c.size // 2
    o0, o1 = threefry2x32(k0, k1, c[:h], c[h:])
    return np.concatenate([o0, o1])[:n]


def jax_uniform(seed, shape, layout="original"):
    """float32 array equal to ``jax.random.uniform(jax.random.PRNGKey(seed), shape)``."""
    n = int(np.prod(shape)) if len(shape) else 1
    b = bits(seed, n, layout)
    f = ((b >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return np.maximum(np.float32(0.0), f).reshape(shape)
