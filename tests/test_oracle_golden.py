"""Pin the oracle against fixtures produced by the REFERENCE'S OWN functions
(tests/golden/make_golden.py; inference.py:287-461, data_loader.py:56-110)."""
import os

import numpy as np
import pytest

from oracle import lifting, lifting_loops, split


def _check_lift(g, prefix=""):
    k = lambda n: g[prefix + n]
    depth, dino, tracks = k("depth"), k("dino"), k("tracks")
    intr = k("intrinsics")
    intr = None if np.isnan(intr).any() else tuple(float(v) for v in intr)
    T, H, W = depth.shape[:3]
    np.testing.assert_array_equal(lifting.lift_2d_to_3d(tracks, depth, intr), k("xyz"))
    np.testing.assert_array_equal(lifting.sample_dino_features_for_tracks(dino, tracks, (T, H, W, 3)), k("dino_feat"))
    np.testing.assert_array_equal(lifting.sample_depth_features_for_tracks(depth, tracks), k("depth_feat"))


def test_lifting_small_bit_exact(golden_dir):
    _check_lift(np.load(os.path.join(golden_dir, "lifting_small.npz")))


@pytest.mark.parametrize("case", ["", "integer/", "far_out/", "one_px/", "intrinsics/"])
def test_per_point_lifting_port_bit_exact(golden_dir, case):
    """oracle/lifting_loops.py (the per-point form bench.py times as the CPU lifting baseline) against the fixtures produced
    by the reference's own functions."""
    g = np.load(os.path.join(golden_dir, "lifting_edges.npz" if case else "lifting_small.npz"))
    depth, dino, tracks, intr = g[case + "depth"], g[case + "dino"], g[case + "tracks"][:6], g[case + "intrinsics"]
    intr = None if np.isnan(intr).any() else tuple(float(v) for v in intr)
    T, H, W = depth.shape[:3]
    np.testing.assert_array_equal(lifting_loops.lift_2d_to_3d(tracks, depth, intr), g[case + "xyz"][:6])
    np.testing.assert_array_equal(lifting_loops.sample_dino_features_for_tracks(dino, tracks, (T, H, W, 3)), g[case + "dino_feat"][:6])
    np.testing.assert_array_equal(lifting_loops.sample_depth_features_for_tracks(depth, tracks), g[case + "depth_feat"][:6])


@pytest.mark.parametrize("case", ["integer", "far_out", "one_px", "intrinsics"])
def test_lifting_edges_bit_exact(golden_dir, case):
    _check_lift(np.load(os.path.join(golden_dir, "lifting_edges.npz")), case + "/")


def test_lifting_integer_pixel_is_that_pixel(golden_dir):
    g = np.load(os.path.join(golden_dir, "lifting_edges.npz"))
    depth, tracks = g["integer/depth"], g["integer/tracks"]
    T, H, W = depth.shape[:3]
    z = lifting.lift_2d_to_3d(tracks, depth)[..., 2]
    xi = np.clip(tracks[..., 0].astype(int), 0, W - 1)
    yi = np.clip(tracks[..., 1].astype(int), 0, H - 1)
    inside = (tracks[..., 0] >= 0) & (tracks[..., 0] <= W - 1) & (tracks[..., 1] >= 0) & (tracks[..., 1] <= H - 1)
    ref = depth[np.arange(T)[None, :], yi, xi, 0]
    np.testing.assert_array_equal(z[inside], ref[inside])


def test_depth_feature_channels(golden_dir):
    g = np.load(os.path.join(golden_dir, "lifting_small.npz"))
    f = lifting.sample_depth_features_for_tracks(g["depth"], g["tracks"])
    assert f.shape[-1] == 256 and not f[..., 3:].any()
    assert not f[:, 0, 2].any()
    np.testing.assert_array_equal(f[..., 1], f[..., 0] / np.float32(10.0))


@pytest.mark.parametrize("seed", [0, 7])
def test_split_bit_exact(golden_dir, seed):
    g = np.load(os.path.join(golden_dir, "split_small.npz"))
    ntot, S, Q, T = (int(v) for v in g["meta"])
    ex = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("example/")}
    b = split.prepare_3d_batch(ex, S, Q, T, seed=seed)
    for k, v in b.items():
        ref = g[f"seed{seed}/{k}"]
        np.testing.assert_array_equal(np.asarray(v, dtype=ref.dtype) if k != "query_points" else v, ref.astype(np.float32) if k == "query_points" else ref, err_msg=k)


def test_split_is_a_partition():
    sup, qry, frames = split.split_indices(100, 60, 30, 17, seed=3)
    assert len(set(sup) | set(qry)) == 90 and not (set(sup) & set(qry))
    assert frames.min() >= 0 and frames.max() < 17


def test_tapvid3d_format_bit_exact(golden_dir):
    from oracle import evaluation as oe

    g = np.load(os.path.join(golden_dir, "tapvid3d_format.npz"))
    tr, occ = oe.convert_predictions_to_tapvid3d_format(g["tracks"], g["visible_logits"])
    np.testing.assert_array_equal(tr, g["pred_tracks"])
    np.testing.assert_array_equal(occ, g["pred_occluded"])
    assert occ.dtype == bool and occ[2, 1] and occ[0, 3]          # logits 0.0 and -0.0 are occluded
