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


def test_host_pack_bf16_is_round_to_nearest_even(spa):
    """spa3d_host_pack_bf16 (host-only code of the library): bit-identical to torch's float32 -> bfloat16 conversion, for every
    thread count, ragged sizes, unaligned destinations, infinities, signed zeros, denormals and ties; NaN stays NaN."""
    import torch
    g = torch.Generator().manual_seed(3)
    for n in (0, 1, 31, 32, 33, 1000, (1 << 16) + 17, 3 * (1 << 16) + 5):
        src = torch.randn(n + 1, generator=g)[1:].contiguous() * 3
        if n >= 31:
            src[:7] = torch.tensor([float("inf"), -float("inf"), 0.0, -0.0, 1e-40, 3.3895314e38, float("nan")])
            bits = torch.tensor([0x3F808000, 0x3F818000, 0x3F807FFF, 0x3F808001, 0x7F7FFFFF], dtype=torch.int32)   # ties and neighbours
            src[7:12] = bits.view(torch.float32)
        want = src.to(torch.bfloat16)
        for threads in (1, 2, 5):
            for off in (0, 1):      # 1: destination not 64-byte aligned (plain stores instead of streaming stores)
                buf = torch.zeros(n + 8, dtype=torch.bfloat16)
                dst = buf[off:off + n]
                spa.ops.host_pack_bf16(src, dst, threads)
                same = (dst.view(torch.int16) == want.view(torch.int16)) | (dst.isnan() & want.isnan())
                assert bool(same.all()), (n, threads, off)
                assert float(buf[off + n:].abs().sum()) == 0 and float(buf[:off].abs().sum()) == 0
    with pytest.raises(ValueError):
        spa.ops.host_pack_bf16(torch.zeros(4), torch.zeros(4), 1)


def test_host_pack_policy(spa, monkeypatch):
    """apply_stream's host_pack="auto": on only for a single rank on a host with >= 12 cores, bf16 precision; explicit "bf16" always
    packs; None / the fp32 precision never do."""
    import os
    f = spa.TrackAutoEncoder3D._pack_threads
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(16)))
    assert f("auto", "bf16", None) == 10 and f("bf16", "bf16", 3) == 3 and f(None, "bf16", None) == 0 and f("auto", "fp32", None) == 0
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(8)))
    assert f("auto", "bf16", None) == 0 and f("bf16", "bf16", None) == 4
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(64)))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    assert f("auto", "bf16", None) == 0 and f("bf16", "bf16", None) == 10
    with pytest.raises(ValueError):
        f("fp8", "bf16", None)


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


def test_default_quantiser_noise_is_the_jax_stream(spa):
    """rng.jax_uniform == jax.random.uniform(PRNGKey(seed), shape): Random123 block vectors, the documented scalar draw and the
    jax quickstart's random.normal(key(0), (10,)) (a function of the same uniform bits) for both counter layouts; the product
    generator and the oracle restatement (written separately) agree bit for bit."""
    import importlib

    from scipy.special import erfinv

    from oracle import threefry as tf

    rng = importlib.import_module("3dspa_code_b200.rng")
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in kat:
        a, b = rng.threefry2x32(key[0], key[1], np.uint32([ctr[0]]), np.uint32([ctr[1]]))
        assert (int(a[0]), int(b[0])) == want
        a, b = tf.threefry2x32(key, np.uint32([ctr[0]]), np.uint32([ctr[1]]))
        assert (int(a[0]), int(b[0])) == want
    assert rng.jax_uniform(0, ()) == np.float32(0.41845703)
    quickstart = {
        "original": [-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377, -0.1521442, -0.67135346, -0.5908641, 0.73168886, 0.5673026],
        "partitionable": [1.6226422, 2.0252647, -0.43359444, -0.07861735, 0.1760909, -0.97208923, -0.49529874, 0.4943786, 0.6643493, -0.9501635],
    }
    lo = np.nextafter(np.float32(-1), np.float32(0))
    for layout, want in quickstart.items():
        b = rng.bits(0, 10, layout)
        f = ((b >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)
        u = np.maximum(lo, f * (np.float32(1) - lo) + lo)
        np.testing.assert_allclose(np.sqrt(2.0) * erfinv(u.astype(np.float64)), want, rtol=2e-6, atol=2e-7)
        for shape in ((), (3,), (2, 128, 96), (5, 7)):
            np.testing.assert_array_equal(rng.jax_uniform(0, shape, layout), tf.uniform(0, shape, layout == "partitionable"))
    u = rng.jax_uniform(0, (2, 128, 96))
    assert u.dtype == np.float32 and 0.0 <= u.min() and u.max() < 1.0 and abs(float(u.mean()) - 0.5) < 0.01
