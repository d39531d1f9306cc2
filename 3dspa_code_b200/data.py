"""Support/query split and query-point sampling (data_loader.py:56-110, inference.py:559-590).

Host-side index work, bit-exact with the reference: the reference draws ``np.random.permutation``
then one ``np.random.randint(0, T)`` per query from the global legacy MT19937 stream; a seeded
``RandomState`` with a vectorised ``randint(size=Q)`` yields the identical stream.  ``seed=None``
uses the global stream exactly like the reference (defect D7: unseeded).
"""
from __future__ import annotations

import numpy as np


def split_indices(num_total, num_support, num_query, num_frames, seed=None):
    rs = np.random.RandomState(seed) if seed is not None else np.random
    perm = rs.permutation(num_total)
    support = perm[:num_support]
    query = perm[num_support : num_support + num_query]
    frames = rs.randint(0, num_frames, size=num_query)
    return support, query, frames


def prepare_3d_batch(example, num_support_tracks=2048, num_query_tracks=2048, num_frames=150, use_dino=True,
                     use_depth=True, seed=None):
    """data_loader.py:56-110.  Returns NumPy arrays with a leading batch axis of 1."""
    tracks_3d = example["tracks_3d"]
    visible = example["visible"]
    sup, qry, frames = split_indices(tracks_3d.shape[0], num_support_tracks, num_query_tracks, num_frames, seed)
    query_tracks = tracks_3d[qry]
    nq = query_tracks.shape[0]
    pos = query_tracks[np.arange(nq), frames[:nq]]
    query_points = np.concatenate([frames[:nq, None].astype(np.float64), pos.astype(np.float64)], axis=1).astype(np.float32)
    batch = {
        "support_tracks": tracks_3d[sup][np.newaxis],
        "support_tracks_visible": visible[sup][np.newaxis],
        "query_points": query_points[np.newaxis],
        "query_tracks": query_tracks[np.newaxis],
        "query_tracks_visible": visible[qry][np.newaxis],
        "boundary_frame": np.array([num_frames]),
    }
    if use_dino and "dino_features" in example:
        batch["dino_features"] = example["dino_features"][sup][np.newaxis]
    if use_depth and "depth_features" in example:
        batch["depth_features"] = example["depth_features"][sup][np.newaxis]
    return batch


def prepare_2d_batch(example, num_support_tracks=2048, num_query_tracks=2048, num_frames=150, seed=None):
    """data_loader.py:13-53."""
    ex = {"tracks_3d": example["tracks"], "visible": example["visible"]}
    return prepare_3d_batch(ex, num_support_tracks, num_query_tracks, num_frames, False, False, seed)


def collate(batches):
    """Stack B single-clip batches along the batch axis."""
    return {k: np.concatenate([b[k] for b in batches], axis=0) for k in batches[0]}
