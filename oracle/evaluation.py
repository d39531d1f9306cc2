"""NumPy restatement of the evaluation adapter (evaluate_tapvid3d.py:39-59).  TEST INFRASTRUCTURE ONLY.
Pinned bit-exactly by tests/golden/tapvid3d_format.npz (made from the reference's own function)."""
import numpy as np


def convert_predictions_to_tapvid3d_format(tracks, visible_logits):
    """tracks [B,Q,T,3], visible_logits [B,Q,T,1] -> ([T,Q,3], [T,Q] bool occluded) of clip 0 (evaluate_tapvid3d.py:47-59)."""
    pred_tracks = np.transpose(np.asarray(tracks)[0], (1, 0, 2))
    pred_occluded = np.transpose(np.asarray(visible_logits)[0, :, :, 0] <= 0.0, (1, 0))
    return pred_tracks, pred_occluded


def reconstruction_score(tracks, target_tracks):
    """|pred - target|_2 per point, [T,Q,1] (the coords_score array of visualize.py:186; float32 accumulation in c order)."""
    d = (np.asarray(tracks, np.float32)[0] - np.asarray(target_tracks, np.float32)[0]).astype(np.float64)
    return np.transpose(np.sqrt((d * d).sum(-1)), (1, 0))[..., None]
