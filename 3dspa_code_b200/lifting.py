"""Host mirror of the reference's feature lifting (inference.py:287-447) on the K0 gather kernel.

Same function names, argument meaning and return shapes as the reference; NumPy in -> NumPy out
by default (``as_numpy=False`` keeps the results on the GPU for the fused inference pipeline).
float32 inputs give results bit-identical to the reference's Python loops.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _dev(a, device="cuda"):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def _ret(t, as_numpy):
    return t.cpu().numpy() if as_numpy and t is not None else t


def lift_2d_to_3d(tracks_2d, depth, intrinsics=None, as_numpy=True):
    """inference.py:287-336: [N,T,2] px + [T,H,W,1] depth -> [N,T,3] camera-space xyz (float32)."""
    xyz, _, _ = ops.lift_sample(_dev(tracks_2d), depth=_dev(depth), intrinsics=intrinsics, want_depth=False)
    return _ret(xyz, as_numpy)


def sample_dino_features_for_tracks(dino_features, tracks_2d, video_shape, as_numpy=True, out_dtype=torch.float32):
    """inference.py:339-395: bilinear sample of [T,Hp,Wp,D] patch features -> [N,T,D]."""
    if dino_features is None:
        return None
    _, H, W, _ = video_shape
    _, f, _ = ops.lift_sample(_dev(tracks_2d), dino=_dev(dino_features), video_hw=(H, W), out_dtype=out_dtype)
    return _ret(f, as_numpy)


def sample_depth_features_for_tracks(depth, tracks_2d, as_numpy=True, out_dtype=torch.float32):
    """inference.py:398-447: [N,T,256] = (d, d/10, d_t - d_{t-1}, 0...)."""
    if depth is None:
        return None
    _, _, z = ops.lift_sample(_dev(tracks_2d), depth=_dev(depth), want_xyz=False, out_dtype=out_dtype)
    return _ret(z, as_numpy)


def lift_and_sample(tracks_2d, depth, dino_features, video_shape, intrinsics=None, out_dtype=torch.float32):
    """All three in ONE pass over the track points (what run_inference needs, inference.py:543-557).
    Returns device tensors (xyz, dino_feat, depth_feat)."""
    _, H, W, _ = video_shape
    return ops.lift_sample(_dev(tracks_2d), depth=_dev(depth), dino=_dev(dino_features), video_hw=(H, W),
                           intrinsics=intrinsics, out_dtype=out_dtype)
