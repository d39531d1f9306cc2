"""Generate the committed golden fixtures from the REFERENCE'S OWN code.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
It imports the reference's pure-NumPy functions unmodified through oracle/ref_import.py
(jax/flax/tf are stubbed; none of the called functions touch them) and stores small
input/output pairs:

  lifting_small.npz   inference.py:287-447  lift_2d_to_3d / sample_dino / sample_depth
  lifting_edges.npz   same, integer pixels, out-of-range points, 1x1 maps, custom intrinsics
  split_small.npz     data_loader.py:56-110 prepare_3d_batch after np.random.seed(seed)
  unflatten.npz       inference.py:450-461  _unflatten_params (stored as flat keys + expected paths)
  tapvid3d_format.npz evaluate_tapvid3d.py:39-59  convert_predictions_to_tapvid3d_format
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_import  # noqa: E402


def lifting_case(rs, N, T, H, W, Hp, Wp, D, spread=4.0, intrinsics=None, integer=False):
    depth = rs.uniform(0.5, 10.0, (T, H, W, 1)).astype(np.float32)
    dino = rs.standard_normal((T, Hp, Wp, D)).astype(np.float32)
    x = rs.uniform(-spread, W - 1 + spread, (N, T))
    y = rs.uniform(-spread, H - 1 + spread, (N, T))
    if integer:
        x, y = np.round(x), np.round(y)
    tracks = np.stack([x, y], -1).astype(np.float32)
    return depth, dino, tracks, intrinsics


def main():
    inf = ref_import.load("inference")
    dl = ref_import.load("data_loader")
    rs = np.random.RandomState(1234)

    def run(depth, dino, tracks, intr):
        T, H, W = depth.shape[:3]
        return dict(
            depth=depth, dino=dino, tracks=tracks,
            intrinsics=np.array(intr if intr is not None else [np.nan] * 4, np.float64),
            xyz=inf.lift_2d_to_3d(tracks, depth, intr),
            dino_feat=inf.sample_dino_features_for_tracks(dino, tracks, (T, H, W, 3)),
            depth_feat=inf.sample_depth_features_for_tracks(depth, tracks),
        )

    small = run(*lifting_case(rs, 24, 6, 28, 42, 2, 3, 16))
    np.savez_compressed(os.path.join(HERE, "lifting_small.npz"), **small)

    edges = {}
    for name, kw in {
        "integer": dict(N=12, T=3, H=9, W=7, Hp=3, Wp=2, D=8, integer=True),
        "far_out": dict(N=12, T=3, H=9, W=7, Hp=3, Wp=2, D=8, spread=40.0),
        "one_px": dict(N=6, T=2, H=1, W=1, Hp=1, Wp=1, D=4),
        "intrinsics": dict(N=10, T=4, H=14, W=14, Hp=1, Wp=1, D=8, intrinsics=(500.5, 480.25, 7.5, 6.25)),
    }.items():
        for k, v in run(*lifting_case(rs, **kw)).items():
            edges[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "lifting_edges.npz"), **edges)

    # split: the reference draws from the global legacy stream
    ntot, S, Q, T = 96, 48, 32, 11
    ex = {
        "tracks_3d": rs.standard_normal((ntot, T, 3)).astype(np.float32),
        "visible": (rs.uniform(size=(ntot, T, 1)) < 0.8).astype(np.float32),
        "dino_features": rs.standard_normal((ntot, T, 8)).astype(np.float32),
        "depth_features": rs.standard_normal((ntot, T, 4)).astype(np.float32),
    }
    out = {f"example/{k}": v for k, v in ex.items()}
    for seed in (0, 7):
        np.random.seed(seed)
        b = dl.prepare_3d_batch(ex, num_support_tracks=S, num_query_tracks=Q, num_frames=T)
        for k, v in b.items():
            out[f"seed{seed}/{k}"] = np.asarray(v)
    out["meta"] = np.array([ntot, S, Q, T])
    np.savez_compressed(os.path.join(HERE, "split_small.npz"), **out)

    # evaluate_tapvid3d.py:39-59 convert_predictions_to_tapvid3d_format (pure NumPy)
    ev = ref_import.load("evaluate_tapvid3d")

    class Pred:
        pass

    pr = Pred()
    pr.tracks = rs.standard_normal((2, 7, 5, 3)).astype(np.float32)
    pr.visible_logits = rs.standard_normal((2, 7, 5, 1)).astype(np.float32)
    pr.visible_logits[0, 1, 2, 0] = 0.0      # logit == 0 counts as occluded (<= 0)
    pr.visible_logits[0, 3, 0, 0] = -0.0
    qp = rs.standard_normal((2, 7, 4)).astype(np.float32)
    tr, occ = ev.convert_predictions_to_tapvid3d_format(pr, qp)
    np.savez_compressed(os.path.join(HERE, "tapvid3d_format.npz"), tracks=pr.tracks, visible_logits=pr.visible_logits, query_points=qp,
                        pred_tracks=tr, pred_occluded=occ)

    flat = {"a/b/kernel": np.arange(6.0).reshape(2, 3), "a/b/bias": np.zeros(3), "a/c/scale": np.ones(2), "top": np.array(3.0)}
    nested = inf._unflatten_params(flat)

    def paths(d, pre=""):
        for k, v in d.items():
            if isinstance(v, dict):
                yield from paths(v, pre + k + "|")
            else:
                yield pre + k

    np.savez(os.path.join(HERE, "unflatten.npz"), expected_paths=np.array(sorted(paths(nested))), **flat)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
