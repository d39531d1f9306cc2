"""NumPy restatement of the reference's feature lifting / sampling.  TEST INFRASTRUCTURE.

Follows /root/reference/inference.py:287-336 (lift_2d_to_3d), :339-395
(sample_dino_features_for_tracks) and :398-447 (sample_depth_features_for_tracks).
Pinned bit-exactly against the reference's own functions by the golden vectors in
``tests/golden/lifting_*.npz`` (made by ``tests/golden/make_golden.py``, which imports the
reference through ``oracle/ref_import.py``).

Arithmetic note (the part that makes bit-exactness possible): the reference runs Python
double loops over NumPy *scalars*.  With float32 ``tracks_2d`` and float32 maps, NumPy-2 weak
scalar promotion keeps every operation in float32 (Python ints/floats adapt to the array
scalar), evaluated left to right with one rounding per operation; results are stored into a
float64 buffer and cast back to float32, which is exact.  This file reproduces that order
with vectorised float32 NumPy ops (NumPy never contracts a*b+c into an FMA).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _bilinear_setup(x, y, W, H):
    """x0=floor(x); wx=x-x0 BEFORE clamping; indices clamped to [0, W-1]/[0, H-1]
    (inference.py:310-319, 374-383, 415-424)."""
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    wx = (x - x0.astype(f32)).astype(f32)
    wy = (y - y0.astype(f32)).astype(f32)
    x1 = np.clip(x0 + 1, 0, W - 1)
    y1 = np.clip(y0 + 1, 0, H - 1)
    x0 = np.clip(x0, 0, W - 1)
    y0 = np.clip(y0, 0, H - 1)
    return x0, y0, x1, y1, wx, wy


def _blend(v00, v01, v10, v11, wx, wy):
    """z00*(1-wx)*(1-wy) + z01*wx*(1-wy) + z10*(1-wx)*wy + z11*wx*wy, left to right, f32
    (inference.py:326-329, 390-393, 431-434)."""
    one = f32(1.0)
    omx = (one - wx).astype(f32)
    omy = (one - wy).astype(f32)
    a = ((v00 * omx).astype(f32) * omy).astype(f32)
    b = ((v01 * wx).astype(f32) * omy).astype(f32)
    c = ((v10 * omx).astype(f32) * wy).astype(f32)
    d = ((v11 * wx).astype(f32) * wy).astype(f32)
    return (((a + b).astype(f32) + c).astype(f32) + d).astype(f32)


def _sample_depth(depth, tracks_2d):
    T, H, W = depth.shape[:3]
    x = tracks_2d[..., 0].astype(f32)
    y = tracks_2d[..., 1].astype(f32)
    x0, y0, x1, y1, wx, wy = _bilinear_setup(x, y, W, H)
    t = np.arange(T)[None, :]
    d = depth[..., 0].astype(f32)
    return _blend(d[t, y0, x0], d[t, y0, x1], d[t, y1, x0], d[t, y1, x1], wx, wy), x, y


def lift_2d_to_3d(tracks_2d, depth, intrinsics=None):
    """inference.py:287-336.  tracks_2d [N,T,2] f32 px, depth [T,H,W,1] f32 -> [N,T,3] f32."""
    H, W = depth.shape[1:3]
    if intrinsics is None:
        fx = fy = max(H, W)
        cx, cy = W / 2, H / 2
    else:
        fx, fy, cx, cy = intrinsics
    z, x, y = _sample_depth(depth, tracks_2d)
    X = (((x - f32(cx)).astype(f32) * z).astype(f32) / f32(fx)).astype(f32)
    Y = (((y - f32(cy)).astype(f32) * z).astype(f32) / f32(fy)).astype(f32)
    return np.stack([X, Y, z], axis=-1).astype(f32)


def sample_dino_features_for_tracks(dino_features, tracks_2d, video_shape):
    """inference.py:339-395.  dino [T,Hp,Wp,D] f32 -> [N,T,D] f32."""
    if dino_features is None:
        return None
    T, Hp, Wp, D = dino_features.shape
    _, H, W, _ = video_shape
    scale_h = Hp / H  # Python floats (weak) -> multiply happens in f32
    scale_w = Wp / W
    px = (tracks_2d[..., 0].astype(f32) * f32(scale_w)).astype(f32)
    py = (tracks_2d[..., 1].astype(f32) * f32(scale_h)).astype(f32)
    x0, y0, x1, y1, wx, wy = _bilinear_setup(px, py, Wp, Hp)
    t = np.arange(T)[None, :]
    f = dino_features.astype(f32)
    wx, wy = wx[..., None], wy[..., None]
    return _blend(f[t, y0, x0], f[t, y0, x1], f[t, y1, x0], f[t, y1, x1], wx, wy)


def sample_depth_features_for_tracks(depth, tracks_2d):
    """inference.py:398-447.  ch0=d, ch1=d/10, ch2=d_t-d_{t-1} (t>0), ch3..255=0."""
    if depth is None:
        return None
    d, _, _ = _sample_depth(depth, tracks_2d)
    N, T = d.shape
    out = np.zeros((N, T, 256), f32)
    out[..., 0] = d
    out[..., 1] = (d / f32(10.0)).astype(f32)
    out[:, 1:, 2] = (d[:, 1:] - d[:, :-1]).astype(f32)
    return out
