"""Support/query split and query-point sampling.  TEST INFRASTRUCTURE.

Restates /root/reference/data_loader.py:56-110 (prepare_3d_batch) and
inference.py:559-590.  The reference draws from the *global legacy* NumPy stream
(``np.random.permutation`` then one ``np.random.randint(0, T)`` per query, in that order);
``RandomState(seed)`` with a vectorised ``randint(size=Q)`` yields the identical MT19937
stream.  Pinned bit-exactly by ``tests/golden/split_*.npz`` (made from the reference's own
``prepare_3d_batch`` after ``np.random.seed(seed)``).
"""
from __future__ import annotations

import numpy as np


def split_indices(num_total, num_support, num_query, num_frames, seed):
    rs = np.random.RandomState(seed)
    perm = rs.permutation(num_total)
    support = perm[:num_support]
    query = perm[num_support : num_support + num_query]
    frames = rs.randint(0, num_frames, size=num_query)
    return support, query, frames


def prepare_3d_batch(example, num_support_tracks=2048, num_query_tracks=2048, num_frames=150,
                     use_dino=True, use_depth=True, seed=0):
    """data_loader.py:56-110 with an explicit seed (reference: unseeded global RNG, D7)."""
    tracks, visible = example["tracks_3d"], example["visible"]
    sup, qry, frames = split_indices(tracks.shape[0], num_support_tracks, num_query_tracks, num_frames, seed)
    qt = tracks[qry]
    qpos = qt[np.arange(num_query_tracks), frames]  # [Q,3]
    # np.array([[t, x, y, z]]) of (int, f32, f32, f32) promotes to float64; jnp.array -> f32
    qp = np.concatenate([frames[:, None].astype(np.float64), qpos.astype(np.float64)], axis=1)
    batch = {
        "support_tracks": tracks[sup][None],
        "support_tracks_visible": visible[sup][None],
        "query_points": qp[None].astype(np.float32),
        "query_tracks": qt[None],
        "query_tracks_visible": visible[qry][None],
        "boundary_frame": np.array([num_frames]),
    }
    if use_dino and "dino_features" in example:
        batch["dino_features"] = example["dino_features"][sup][None]
    if use_depth and "depth_features" in example:
        batch["depth_features"] = example["depth_features"][sup][None]
    return batch
