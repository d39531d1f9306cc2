"""Per-point restatement of the reference's feature lifting, the way the reference runs it: Python loops over
(track, frame) with NumPy scalar arithmetic.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

/root/reference/inference.py:287-336 (lift_2d_to_3d), :339-395 (sample_dino_features_for_tracks) and :398-447
(sample_depth_features_for_tracks) each walk ``for n in range(N): for t in range(T):`` and blend four neighbours per point.
``oracle/lifting.py`` is the vectorised form of the same arithmetic (pinned bit-exactly against the reference's functions
by tests/golden/lifting_*.npz); this file keeps the reference's COST MODEL - one interpreter iteration per point, NumPy
scalars, a 768-wide vector blend per DINO point - so ``bench.py`` can time "the reference's CPU lifting" on the GPU box,
where /root/reference does not exist.  tests/test_oracle_golden.py holds it bit-equal to ``oracle/lifting.py``.
"""
from __future__ import annotations

import numpy as np


def _corners(x, y, width, height):
    """floor / weights before clamping / clamped corner indices (inference.py:310-319)."""
    xf, yf = int(np.floor(x)), int(np.floor(y))
    wx, wy = x - xf, y - yf
    xa, xb = min(max(xf, 0), width - 1), min(max(xf + 1, 0), width - 1)
    ya, yb = min(max(yf, 0), height - 1), min(max(yf + 1, 0), height - 1)
    return xa, xb, ya, yb, wx, wy


def _bilinear(plane, x, y):
    """plane [H, W, ...]: the four-neighbour blend in the reference's association order (inference.py:326-329)."""
    xa, xb, ya, yb, wx, wy = _corners(x, y, plane.shape[1], plane.shape[0])
    return (plane[ya, xa] * (1 - wx) * (1 - wy) + plane[ya, xb] * wx * (1 - wy)
            + plane[yb, xa] * (1 - wx) * wy + plane[yb, xb] * wx * wy)


def lift_2d_to_3d(tracks_2d, depth, intrinsics=None):
    num, frames = tracks_2d.shape[:2]
    height, width = depth.shape[1:3]
    fx, fy, cx, cy = intrinsics if intrinsics is not None else (max(height, width), max(height, width), width / 2, height / 2)
    out = np.zeros((num, frames, 3))
    for n in range(num):
        for t in range(frames):
            x, y = tracks_2d[n, t]
            z = _bilinear(depth[t, :, :, 0], x, y)
            out[n, t] = ((x - cx) * z / fx, (y - cy) * z / fy, z)
    return out.astype(np.float32)


def sample_dino_features_for_tracks(dino_features, tracks_2d, video_shape):
    num, frames = tracks_2d.shape[:2]
    _, hp, wp, dim = dino_features.shape
    sy, sx = hp / video_shape[1], wp / video_shape[2]
    out = np.zeros((num, frames, dim))
    for n in range(num):
        for t in range(frames):
            x, y = tracks_2d[n, t]
            out[n, t] = _bilinear(dino_features[t], x * sx, y * sy)
    return out.astype(np.float32)


def sample_depth_features_for_tracks(depth, tracks_2d, feature_dim=256):
    num, frames = tracks_2d.shape[:2]
    out = np.zeros((num, frames, feature_dim))
    for n in range(num):
        prev = None
        for t in range(frames):
            x, y = tracks_2d[n, t]
            d = _bilinear(depth[t, :, :, 0], x, y)
            out[n, t, 0] = d
            out[n, t, 1] = d / 10.0
            if prev is not None:
                out[n, t, 2] = d - prev
            prev = d
    return out.astype(np.float32)
