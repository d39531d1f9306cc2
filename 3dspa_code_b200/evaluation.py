"""Host mirror of the evaluation adapter (evaluate_tapvid3d.py:39-59) and the per-point score the visualiser consumes
(visualize.py:186, ``coords_score``), on the device.

``convert_predictions_to_tapvid3d_format`` keeps the reference's name, arguments and return types (NumPy ``[T,N,3]`` tracks,
``[T,N]`` bool occlusion, first clip of the batch).  ``reconstruction_score`` is the per-point reconstruction error
``|pred - target|_2`` in the same ``[T,N]`` order - the reference ships the consumer of that array but not its producer.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _dev(a):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def convert_predictions_to_tapvid3d_format(predictions, query_points=None, as_numpy=True):
    """predictions.tracks [B,Q,T,3], predictions.visible_logits [B,Q,T,1] -> (pred_tracks [T,Q,3], pred_occluded [T,Q])."""
    tr = _dev(predictions.tracks)[0]
    lg = _dev(predictions.visible_logits)[0, :, :, 0].contiguous()
    out_t, occ, _ = ops.to_tapvid3d(tr.contiguous(), lg)
    return (out_t.cpu().numpy(), occ.cpu().numpy()) if as_numpy else (out_t, occ)


def reconstruction_score(predictions, targets, as_numpy=True):
    """Per-point reconstruction error [T,Q,1] of clip 0 against ``targets['query_tracks']`` [B,Q,T,3]."""
    tr = _dev(predictions.tracks)[0].contiguous()
    lg = _dev(predictions.visible_logits)[0, :, :, 0].contiguous()
    tg = _dev(targets["query_tracks"])[0].contiguous()
    _, _, score = ops.to_tapvid3d(tr, lg, tg)
    score = score.unsqueeze(-1)
    return score.cpu().numpy() if as_numpy else score
