"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
public header declares (no compute calls), host logic (params, checkpoints, split) is exact."""
import ctypes
import os

import numpy as np
import pytest

from tests.helpers import product


@pytest.fixture(scope="module")
def spa():
    return product()


def test_library_exports_every_declared_symbol(spa):
    protos = spa._lib.PROTOTYPES
    assert len(protos) >= 28
    lib = ctypes.CDLL(spa._lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    assert spa.ops.version() >= 100
    assert isinstance(spa._lib.lib().spa3d_last_error(), bytes)


def test_header_has_no_torch_types():
    src = open(os.path.join(os.path.dirname(__file__), "..", "include", "spa3d_b200.h")).read()
    assert "torch" not in src.lower().replace("torch custom", "") and "at::" not in src and 'extern "C"' in src


def test_ops_refuse_cpu_tensors(spa):
    import torch
    with pytest.raises(ValueError):
        spa.ops.convert(torch.zeros(2, 2), torch.zeros(2, 2))


def test_pack_unpack_roundtrip_and_counts(spa):
    model = spa.TrackAutoEncoder3D()
    tree = model.init(0, {"dino_features": 1, "depth_features": 1})["params"]
    assert spa.params.count(tree) == 109_138_296
    meta = spa.params.tree_meta(tree)
    back = spa.params.unpack(spa.params.pack(tree), meta)
    f1, f2 = spa.params.flatten(tree), spa.params.flatten(back)
    assert f1.keys() == f2.keys() and all(np.array_equal(f1[k], f2[k]) for k in f1)
    assert "track_readout_attn/layer_3/self_att/dense_out/kernel" in f1
    t2 = spa.TrackAutoEncoder().init(0)["params"]
    assert spa.params.count(t2) == 68_333_080
    lazy = model.init(0, {})["params"]  # Flax creates projections only if the init batch has the feature
    assert "dino_projection" not in lazy and "depth_projection" not in lazy


def test_checkpoint_layouts(spa, tmp_path, golden_dir):
    c = dict(num_output_frames=4, num_latent_tokens=4, latent_token_dim=8, track_token_dim=16, encoder_latent_dim=16,
             decoder_num_channels=128 + 16, dino_feature_dim=8, depth_feature_dim=8)
    arch = {k: (16, 2, 16, 1) for k in ("itt", "t2l", "dec", "tra")}
    tree = spa.TrackAutoEncoder3D(**c).init(1, {"dino_features": 1, "depth_features": 1}, arch=arch)["params"]
    flat = spa.params.flatten(tree)
    p1 = str(tmp_path / "flat.npz")
    spa.save_checkpoint(p1, tree)
    p2 = str(tmp_path / "params.npz")
    np.savez(p2, params=np.array(tree, dtype=object))
    p3 = str(tmp_path / "opt.npz")
    np.savez(p3, optimizer=np.array({"target": tree}, dtype=object))
    for p in (p1, p2, p3):
        got = spa.params.flatten(spa.load_checkpoint(p))
        assert got.keys() == flat.keys() and all(np.array_equal(got[k], flat[k]) for k in flat)
    # _unflatten_params fixture produced by the reference's own function (inference.py:450-461)
    g = np.load(os.path.join(golden_dir, "unflatten.npz"))
    nested = spa.params.unflatten({k: g[k] for k in g.files if k != "expected_paths"})
    paths = sorted(k.replace("/", "|") for k in spa.params.flatten(nested))
    assert paths == list(g["expected_paths"])
    # warn-only structure check (inference.py:608-619)
    broken = spa.params.unflatten({k: v for k, v in flat.items() if not k.startswith("compressor")})
    with pytest.warns(UserWarning):
        probs = spa.check_structure(tree, broken)
    assert any("compressor" in p for p in probs)


def test_width_mismatch_is_reported(spa):
    tree = spa.TrackAutoEncoder3D().init(0, {"dino_features": 1})["params"]
    tree["dino_projection"]["kernel"] = np.zeros((768, 768), np.float32)  # the as-written width (R2)
    with pytest.raises(ValueError, match="ADDED"):
        spa.params.tree_meta(tree)


def test_head_divisibility_error(spa):
    with pytest.raises(ValueError, match="must divide"):  # attention.py:147-148
        spa.TrackAutoEncoder3D().init(0, {}, arch={k: (100, 8, 16, 1) for k in ("itt", "t2l", "dec", "tra")})


@pytest.mark.parametrize("seed", [0, 7])
def test_split_matches_reference_fixture(spa, golden_dir, seed):
    g = np.load(os.path.join(golden_dir, "split_small.npz"))
    ntot, S, Q, T = (int(v) for v in g["meta"])
    ex = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("example/")}
    b = spa.data.prepare_3d_batch(ex, S, Q, T, seed=seed)
    for k, v in b.items():
        ref = g[f"seed{seed}/{k}"]
        np.testing.assert_array_equal(v, ref.astype(v.dtype), err_msg=k)
    np.random.seed(seed)  # seed=None follows the global legacy stream exactly like the reference
    b2 = spa.data.prepare_3d_batch(ex, S, Q, T, seed=None)
    np.testing.assert_array_equal(b2["query_points"], b["query_points"])
