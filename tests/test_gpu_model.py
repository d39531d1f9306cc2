"""End-to-end parity of the CUDA path (through the module mirror -> C ABI) against the oracle.

Tolerances are the north_star's: per-tensor max|a-b|/max|b| <= 1e-4 for the fp32 path and <= 2e-2
for the bf16 path, on reconstructed tracks and visibility logits.
"""
import numpy as np
import pytest
import torch

from oracle import model as om
from tests.helpers import SMALL_ARCH, make_inputs, product, quantiser_aware_reference, rel_err, small_cfg

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.fixture(scope="module")
def spa():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return product()


def oracle_fwd(cfg, tree, inp, noise, discretize=True, dtype=torch.float32):
    ocfg = om.Config3D(**{k: getattr(cfg, k) for k in om.Config3D.__dataclass_fields__})
    with torch.no_grad():
        return om.forward_3d(om.to_torch(tree, dtype), ocfg, om.cast_inputs(inp, dtype), torch.as_tensor(noise).to(dtype), discretize)


def randomize(tree, seed):
    rng = np.random.RandomState(seed)
    om._randomize(tree, rng)
    return tree


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("discretize", [False, True])
def test_small_arch_forward(spa, precision, discretize):
    c = small_cfg()
    model = spa.TrackAutoEncoder3D(**{k: getattr(c, k) for k in om.Config3D.__dataclass_fields__})
    inp, noise = make_inputs(c, B=2, N=10, Q=6)
    inp["boundary_frame"] = np.array([c.num_output_frames - 2, c.num_output_frames], np.int32)
    variables = model.init(3, inp, arch=SMALL_ARCH)
    randomize(variables["params"], 3)
    got = model.apply(variables, inp, noise=noise, discretize=discretize, precision=precision)
    tol = TOL[precision]
    if discretize:
        # round(z * 128) is discontinuous: the latents are held to the tolerance, and the decoder is compared on identical
        # rounding decisions (tests.helpers.quantiser_aware_reference) - the tolerance itself is NOT relaxed
        z_gpu = model.apply(variables, inp, method="encode", precision=precision)
        _, ref, z_ref, _ = quantiser_aware_reference(variables["params"], c, inp, noise, z_gpu)
        assert rel_err(z_gpu, z_ref) < tol, rel_err(z_gpu, z_ref)
    else:
        ref = oracle_fwd(model, variables["params"], inp, noise, discretize)
    assert rel_err(got.tracks, ref.tracks) < tol, rel_err(got.tracks, ref.tracks)
    assert rel_err(got.visible_logits, ref.visible_logits) < tol
    assert not got.certain_logits.any()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_reference_widths_forward(spa, precision):
    """The real architecture (109.14 M parameters, T=150, DINO 768 + depth 256) on a few tracks."""
    model = spa.TrackAutoEncoder3D()
    c = om.Config3D()
    inp, noise = make_inputs(c, B=1, N=24, Q=8, seed=4)
    variables = model.init(5, inp)
    randomize(variables["params"], 5)
    spa.ops.stats(reset=True)
    got = model.apply(variables, inp, noise=noise, discretize=False, precision=precision)
    st = spa.ops.stats()
    # the accurate mode runs its large contractions on the tensor cores too ("bf16 x 3"); the bf16 mode never leaves them
    assert (st["gemm_x3"] > 40 and st["gemm_tcgen05"] == 0) if precision == "fp32" else (st["gemm_tcgen05"] > 40 and st["gemm_bf16_fallback"] == 0), st
    ref = oracle_fwd(model, variables["params"], inp, noise, False)
    assert got.tracks.shape == (1, 8, 150, 3) and got.visible_logits.shape == (1, 8, 150, 1)
    assert rel_err(got.tracks, ref.tracks) < TOL[precision], rel_err(got.tracks, ref.tracks)
    assert rel_err(got.visible_logits, ref.visible_logits) < TOL[precision]


def test_encode_decode_methods_and_chunking(spa):
    c = small_cfg()
    fields = {k: getattr(c, k) for k in om.Config3D.__dataclass_fields__}
    model = spa.TrackAutoEncoder3D(**fields)
    inp, noise = make_inputs(c, B=2, N=8, Q=6)
    variables = model.init(7, inp, arch=SMALL_ARCH)
    full = model.apply(variables, inp, noise=noise, precision="fp32")
    lat = model.apply(variables, inp, method="encode", precision="fp32")
    assert lat.shape == (2, c.num_latent_tokens, c.latent_token_dim)
    ctx = model.apply(variables, inp, method="get_decoder_context", precision="fp32")
    assert ctx.query_frame.dtype == torch.int32
    dec = model.decode(variables, lat, ctx, noise=noise, precision="fp32")
    assert torch.equal(dec.tracks, full.tracks)
    fields["decoder_scan_chunk_size"] = 2
    chunked = spa.TrackAutoEncoder3D(**fields).apply(variables, inp, noise=noise, precision="fp32")
    assert rel_err(chunked.tracks, full.tracks) < 1e-6  # each query is independent of the others
    # no DINO / depth supplied although the tree has the projections (:140,145)
    inp2 = {k: v for k, v in inp.items() if k not in ("dino_features", "depth_features")}
    got = model.apply(variables, inp2, noise=noise, precision="fp32")
    ref = oracle_fwd(model, variables["params"], inp2, noise)
    assert rel_err(got.tracks, ref.tracks) < 1e-4
    # default query grid when query_points is absent (:214-226)
    inp3 = {k: v for k, v in inp.items() if k != "query_points"}
    got = model.apply(variables, inp3, noise=noise, precision="fp32")
    assert got.tracks.shape[1] == 1024
    ref = oracle_fwd(model, variables["params"], inp3, noise)
    assert rel_err(got.tracks, ref.tracks) < 1e-4


def test_properties_at_reference_scale(spa):
    """Size-independent properties at BASELINE config-2 sizes (oracle too slow there):
    support-permutation invariance and irrelevance of masked frames, bf16 path, N=2048, Q=512."""
    model = spa.TrackAutoEncoder3D()
    c = om.Config3D()
    inp, noise = make_inputs(c, B=1, N=2048, Q=512, seed=8)
    variables = model.init(9, inp)
    a = model.apply(variables, inp, noise=noise, precision="bf16")
    assert torch.isfinite(a.tracks).all() and torch.isfinite(a.visible_logits).all()
    perm = np.random.RandomState(0).permutation(2048)
    inp2 = dict(inp)
    for k in ("support_tracks", "support_tracks_visible", "dino_features", "depth_features"):
        inp2[k] = inp[k][:, perm]
    b = model.apply(variables, inp2, noise=noise, precision="bf16")
    assert rel_err(b.tracks, a.tracks) < 2e-2
    vis = inp["support_tracks_visible"][..., 0] > 0
    inp3 = dict(inp)
    junk = np.random.RandomState(1).standard_normal(inp["support_tracks"].shape).astype(np.float32)
    inp3["support_tracks"] = np.where(vis[..., None], inp["support_tracks"], junk)
    d = model.apply(variables, inp3, noise=noise, precision="bf16")
    assert rel_err(d.tracks, a.tracks) < 2e-2


def test_trajan_2d_forward(spa):
    """TRAJAN 2D (track_autoencoder.py) through the same kernels: masked-mean pooling, 4T head."""
    ocfg = om.Config2D(num_output_frames=10, num_latent_tokens=8, latent_token_dim=16, track_token_dim=32,
                       encoder_latent_dim=48, decoder_num_channels=128 + 48)
    model = spa.TrackAutoEncoder(**{k: getattr(ocfg, k) for k in om.Config2D.__dataclass_fields__})
    inp, noise = make_inputs(ocfg, B=2, N=7, Q=5, T=10, dino=False, depth=False, coords=2)
    inp["support_tracks_visible"][0, 2] = 0  # a fully invisible track: uniform attention + zero pooling weight
    variables = model.init(11, inp, arch=SMALL_ARCH)
    randomize(variables["params"], 11)
    got = model.apply(variables, inp, noise=noise, precision="fp32")
    with torch.no_grad():
        ref = om.forward_2d(om.to_torch(variables["params"]), ocfg, om.cast_inputs(inp, torch.float32), torch.as_tensor(noise))
    assert got.tracks.shape == (2, 5, 10, 2)
    assert rel_err(got.tracks, ref.tracks) < 1e-4
    assert rel_err(got.visible_logits, ref.visible_logits) < 1e-4
    assert rel_err(got.certain_logits, ref.certain_logits) < 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_streamed_host_inputs_equal_one_shot_upload(spa, precision):
    """Host-resident (pinned) inputs are uploaded chunk by chunk under the per-track transformer;
    the latents must equal the one-shot path bit for bit (tracks are independent until pooling)."""
    c = small_cfg()
    model = spa.TrackAutoEncoder3D(**{k: getattr(c, k) for k in om.Config3D.__dataclass_fields__})
    inp, noise = make_inputs(c, B=2, N=11, Q=6)
    inp["boundary_frame"] = np.array([c.num_output_frames - 3, c.num_output_frames], np.int32)
    variables = model.init(5, inp, arch=SMALL_ARCH)
    randomize(variables["params"], 5)
    eng = model.bind(variables, precision)
    dev_inp = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
    with torch.no_grad():
        ref = eng.encode(dev_inp)
        host_inp = {k: torch.as_tensor(v).pin_memory() for k, v in inp.items()}
        eng.stream_chunk = 4          # 11 tracks -> chunks of 4, 4, 3 per clip
        got = eng.encode(host_inp)
        eng.stream_chunk = 256
    assert torch.equal(got, ref)
    res_a = model.apply(variables, host_inp, noise=noise, precision=precision)
    res_b = model.apply(variables, dev_inp, noise=torch.as_tensor(noise).cuda(), precision=precision)
    assert torch.equal(res_a.tracks, res_b.tracks) and torch.equal(res_a.visible_logits, res_b.visible_logits)


def test_cuda_graph_replay_equals_eager(spa):
    """model.cuda_graph = True replays the whole forward as one CUDA graph when the same device buffers
    are passed again; results must equal the eager launches bit for bit, also after the buffers are
    refilled with another clip."""
    c = small_cfg()
    fields = {k: getattr(c, k) for k in om.Config3D.__dataclass_fields__}
    model = spa.TrackAutoEncoder3D(**fields)
    gmodel = spa.TrackAutoEncoder3D(**fields)
    gmodel.cuda_graph = True
    inp, noise = make_inputs(c, B=2, N=10, Q=6, seed=1)
    variables = model.init(6, inp, arch=SMALL_ARCH)
    randomize(variables["params"], 6)
    dev_inp = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
    dev_noise = torch.as_tensor(noise).cuda()
    ref = model.apply(variables, dev_inp, noise=dev_noise, precision="bf16")
    got = gmodel.apply(variables, dev_inp, noise=dev_noise, precision="bf16")      # capture + first replay
    assert torch.equal(got.tracks, ref.tracks) and torch.equal(got.visible_logits, ref.visible_logits)
    inp2, noise2 = make_inputs(c, B=2, N=10, Q=6, seed=2)
    for k, v in inp2.items():
        dev_inp[k].copy_(torch.as_tensor(v))                                        # refill the same buffers
    dev_noise.copy_(torch.as_tensor(noise2))
    ref2 = model.apply(variables, dev_inp, noise=dev_noise, precision="bf16")
    got2 = gmodel.apply(variables, dev_inp, noise=dev_noise, precision="bf16")     # pure replay
    assert len(gmodel._graphs) == 1
    assert torch.equal(got2.tracks, ref2.tracks) and torch.equal(got2.visible_logits, ref2.visible_logits)
    assert not torch.equal(ref2.tracks, ref.tracks)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_matches_reference_executed_golden(spa, precision, golden_dir):
    """The CUDA path against numbers produced by the reference's OWN source (tests/golden/make_golden_model.py:
    track_autoencoder_3d.py executed on NumPy stand-ins for flax, R2' widths, R1 key mask) - no oracle in between."""
    import os

    g = np.load(os.path.join(golden_dir, "model_3dspa.npz"))
    F = int(g["dec/num_output_frames"])
    model = spa.TrackAutoEncoder3D(num_output_frames=F, track_token_dim=768, depth_feature_dim=768)
    cfg = om.Config3D(num_output_frames=F, track_token_dim=768, depth_feature_dim=768)
    variables = {"params": om.init_params_3d(cfg, seed=int(g["full/seed"]), randomize_norms=True)}
    inp = {k[8:]: np.asarray(g[k], np.int32 if k.endswith("boundary_frame") else np.float32) for k in g.files if k.startswith("full/in/")}
    lat = model.apply(variables, inp, method="encode", precision=precision)
    assert rel_err(lat, torch.from_numpy(g["full/latents"])) < TOL[precision], rel_err(lat, torch.from_numpy(g["full/latents"]))
    # whole forward without the quantiser (a last-bit difference in a latent may flip a round(x*128) and move it by 1/128)
    got = model.apply(variables, inp, discretize=False, precision=precision)
    assert rel_err(got.tracks, torch.from_numpy(g["full/nodisc/tracks"])) < TOL[precision], rel_err(got.tracks, torch.from_numpy(g["full/nodisc/tracks"]))
    assert rel_err(got.visible_logits, torch.from_numpy(g["full/nodisc/visible_logits"])) < TOL[precision]
    # decoder as written (default widths) from the golden latents (float32-valued, so the quantiser rounds identically)
    model_d = spa.TrackAutoEncoder3D(num_output_frames=F)
    cfg_d = om.Config3D(num_output_frames=F)
    vars_d = {"params": om.init_params_3d(cfg_d, seed=int(g["dec/seed"]), randomize_norms=True)}
    inp_d = {k[7:]: np.asarray(g[k], np.int32 if k.endswith("boundary_frame") else np.float32) for k in g.files if k.startswith("dec/in/")}
    ctx = model_d.apply(vars_d, inp_d, method="get_decoder_context", precision=precision)
    lat_d = np.asarray(g["dec/latents"], np.float32)
    res = model_d.decode(vars_d, lat_d, ctx, discretize=False, precision=precision)
    assert rel_err(res.tracks, torch.from_numpy(g["dec/nodisc/tracks"])) < TOL[precision], rel_err(res.tracks, torch.from_numpy(g["dec/nodisc/tracks"]))
    assert rel_err(res.visible_logits, torch.from_numpy(g["dec/nodisc/visible_logits"])) < TOL[precision]
    res = model_d.decode(vars_d, lat_d, ctx, discretize=True, noise=np.asarray(g["dec/noise"], np.float32), precision=precision)
    assert rel_err(res.tracks, torch.from_numpy(g["dec/tracks"])) < TOL[precision], rel_err(res.tracks, torch.from_numpy(g["dec/tracks"]))
    assert rel_err(res.visible_logits, torch.from_numpy(g["dec/visible_logits"])) < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_trajan_matches_reference_executed_golden(spa, golden_dir, precision):
    """TRAJAN 2D end to end, as written (track_autoencoder.py executed on the flax stand-ins), fp32 and bf16 paths."""
    import os

    g = np.load(os.path.join(golden_dir, "model_trajan.npz"))
    F = int(g["num_output_frames"])
    model = spa.TrackAutoEncoder(num_output_frames=F)
    variables = {"params": om.init_params_2d(om.Config2D(num_output_frames=F), seed=int(g["seed"]), randomize_norms=True)}
    inp = {k[3:]: np.asarray(g[k], np.int32 if k.endswith("boundary_frame") else np.float32) for k in g.files if k.startswith("in/")}
    got = model.apply(variables, inp, discretize=False, precision=precision)   # quantiser off: no round(x*128) flips
    for name in ("tracks", "visible_logits", "certain_logits"):
        e = rel_err(getattr(got, name), torch.from_numpy(g["nodisc/" + name]))
        assert e < TOL[precision], (name, e)
    lat = model.apply(variables, inp, method="encode", precision=precision)
    assert rel_err(lat, torch.from_numpy(g["latents"])) < TOL[precision]


def test_forward_from_maps_matches_feature_path_and_oracle(spa):
    """lift -> sample -> embed fused ("project, then sample").  The per-track features are never built;
    tokens, and the final outputs, agree with (a) the unfused CUDA path fed with lifted features and (b) the oracle
    (NumPy lifting restatement + torch model restatement) within the bf16 tolerance."""
    from oracle import lifting as ol

    rs = np.random.RandomState(12)
    N, T, Q, H, W, Hp, Wp = 70, 150, 6, 20, 24, 5, 6
    model = spa.TrackAutoEncoder3D()
    c = om.Config3D()
    tracks_2d = np.stack([rs.uniform(-2, W + 1, (N, T)), rs.uniform(-2, H + 1, (N, T))], -1).astype(np.float32)
    depth = rs.uniform(0.5, 4.0, (T, H, W, 1)).astype(np.float32)
    dino_map = rs.standard_normal((T, Hp, Wp, 768)).astype(np.float32)
    visible = (rs.uniform(size=(N, T, 1)) < 0.85).astype(np.float32)
    video_shape = (T, H, W, 3)
    # the reference pipeline on the CPU oracle: lifted xyz + per-track features
    xyz = ol.lift_2d_to_3d(tracks_2d, depth)
    dfe = ol.sample_dino_features_for_tracks(dino_map, tracks_2d, video_shape)
    zfe = ol.sample_depth_features_for_tracks(depth, tracks_2d)
    qp = np.concatenate([rs.randint(0, T, (1, Q, 1)).astype(np.float32), rs.uniform(-1, 1, (1, Q, 3)).astype(np.float32)], -1)
    feat_inputs = {"support_tracks": xyz[None], "support_tracks_visible": visible[None], "query_points": qp,
                   "boundary_frame": np.array([T], np.int32), "dino_features": dfe[None], "depth_features": zfe[None]}
    variables = model.init(21, feat_inputs)
    randomize(variables["params"], 21)
    maps_inputs = {"support_tracks_2d": tracks_2d, "support_tracks_visible": visible, "depth": depth, "dino_map": dino_map,
                   "video_shape": video_shape, "query_points": qp}
    eng = model.bind(variables, "bf16")
    dev = "cuda"
    x_f, xyz_dev = eng.embed_tracks_from_maps(torch.from_numpy(tracks_2d).to(dev), torch.from_numpy(depth).to(dev),
                                              torch.from_numpy(dino_map).to(dev), (H, W))
    np.testing.assert_array_equal(xyz_dev.cpu().numpy(), xyz)   # lifting stays bit-exact
    x_u = eng.embed_tracks(torch.from_numpy(xyz[None]).to(dev), torch.from_numpy(dfe[None]).to(dev), torch.from_numpy(zfe[None]).to(dev), readout=True)
    assert rel_err(x_f, x_u) < 1e-2, rel_err(x_f, x_u)
    got = model.apply_from_maps(variables, maps_inputs, discretize=False)
    unf = model.apply(variables, feat_inputs, discretize=False, precision="bf16")
    ref = oracle_fwd(model, variables["params"], feat_inputs, np.zeros((1, 128, 96), np.float32), False)
    assert rel_err(got.tracks, unf.tracks) < 2e-2, rel_err(got.tracks, unf.tracks)
    assert rel_err(got.tracks, ref.tracks) < 2e-2, rel_err(got.tracks, ref.tracks)
    assert rel_err(got.visible_logits, ref.visible_logits) < 2e-2, rel_err(got.visible_logits, ref.visible_logits)


def test_default_noise_is_the_reference_draw(spa):
    """apply() without ``noise`` regenerates jax.random.uniform(PRNGKey(0), [B,128,96]) (track_autoencoder_3d.py:254-257)."""
    import importlib

    rng = importlib.import_module("3dspa_code_b200.rng")
    c = small_cfg()
    model = spa.TrackAutoEncoder3D(**{k: getattr(c, k) for k in om.Config3D.__dataclass_fields__})
    inp, _ = make_inputs(c, B=2, N=8, Q=5)
    variables = model.init(2, inp, arch=SMALL_ARCH)
    a = model.apply(variables, inp, precision="fp32")
    for layout in ("original", "partitionable"):
        model.noise_layout = layout
        b = model.apply(variables, inp, precision="fp32")
        n = rng.jax_uniform(0, (2, c.num_latent_tokens, c.latent_token_dim), layout)
        e = model.apply(variables, inp, noise=n, precision="fp32")
        assert torch.equal(b.tracks, e.tracks)
        if layout == "original":
            assert torch.equal(a.tracks, b.tracks)
        else:
            assert not torch.equal(a.tracks, b.tracks)


def test_tapvid3d_adapter_matches_reference_fixture(spa, golden_dir):
    """evaluate_tapvid3d.py:39-59 on the device, bit-exact against the reference's own function; reconstruction score vs oracle."""
    import importlib
    import os

    from oracle import evaluation as oe

    ev = importlib.import_module("3dspa_code_b200.evaluation")
    g = np.load(os.path.join(golden_dir, "tapvid3d_format.npz"))

    class Pred:
        tracks = g["tracks"]
        visible_logits = g["visible_logits"]

    tr, occ = ev.convert_predictions_to_tapvid3d_format(Pred, g["query_points"])
    np.testing.assert_array_equal(tr, g["pred_tracks"])
    np.testing.assert_array_equal(occ, g["pred_occluded"])
    assert occ.dtype == bool
    tgt = np.random.RandomState(0).standard_normal(g["tracks"].shape).astype(np.float32)
    sc = ev.reconstruction_score(Pred, {"query_tracks": tgt})
    np.testing.assert_allclose(sc, oe.reconstruction_score(g["tracks"], tgt), rtol=1e-6, atol=1e-7)


def test_sweep_clip_shape_properties(spa):
    """BASELINE config 5 clip shape (4096 support / 1024 query tracks, 518x518 video, 37x37 patch map) through the fused
    maps path: finite, invariant to the order of the support tracks, and the realism score has the visualiser's layout."""
    import importlib

    ev = importlib.import_module("3dspa_code_b200.evaluation")
    lifting = importlib.import_module("3dspa_code_b200.lifting")
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    S5, Q5, T, H, W = 4096, 1024, 150, 518, 518
    model = spa.TrackAutoEncoder3D()
    variables = model.init(4, {"dino_features": 1, "depth_features": 1})
    depth = torch.rand(T, H, W, 1, generator=g, device=dev) * 9.5 + 0.5
    dino = torch.randn(T, 37, 37, 768, generator=g, device=dev)
    tr2 = (torch.rand(S5 + Q5, T, 2, generator=g, device=dev) * (W - 1)).contiguous()
    vis = (torch.rand(S5, T, 1, generator=g, device=dev) < 0.9).float()
    q_xyz = lifting.lift_2d_to_3d(tr2[S5:], depth, as_numpy=False)
    qp = torch.cat([torch.zeros(Q5, 1, device=dev), q_xyz[:, 0]], -1)[None]
    inputs = {"support_tracks_2d": tr2[:S5], "support_tracks_visible": vis, "depth": depth, "dino_map": dino,
              "video_shape": (T, H, W, 3), "query_points": qp}
    a = model.apply_from_maps(variables, inputs)
    assert a.tracks.shape == (1, Q5, T, 3) and torch.isfinite(a.tracks).all() and torch.isfinite(a.visible_logits).all()
    perm = torch.randperm(S5, generator=g, device=dev)
    inputs2 = dict(inputs, support_tracks_2d=tr2[:S5][perm].contiguous(), support_tracks_visible=vis[perm].contiguous())
    b = model.apply_from_maps(variables, inputs2)
    assert rel_err(b.tracks, a.tracks) < 2e-2, rel_err(b.tracks, a.tracks)
    sc = ev.reconstruction_score(a, {"query_tracks": q_xyz[None]}, as_numpy=False)
    assert sc.shape == (T, Q5, 1) and torch.isfinite(sc).all()


def test_apply_stream_equals_per_clip_apply(spa):
    """model.apply_stream (uploads of clip i+1 under the forward of clip i, asynchronous read-back) yields, clip by clip, exactly
    what apply() returns for that clip - eagerly and with CUDA-graph replay; bfloat16 host features agree within the bf16 budget."""
    c = small_cfg()
    fields = {k: getattr(c, k) for k in om.Config3D.__dataclass_fields__}
    model = spa.TrackAutoEncoder3D(**fields)
    clips = [make_inputs(c, B=1, N=9, Q=5, seed=40 + i) for i in range(5)]
    variables = model.init(8, clips[0][0], arch=SMALL_ARCH)
    randomize(variables["params"], 8)
    pin = lambda d: {k: torch.as_tensor(v).pin_memory() for k, v in d.items()}
    batches = [pin(b) for b, _ in clips]
    noises = [torch.as_tensor(n).pin_memory() for _, n in clips]
    refs = [model.apply(variables, b, noise=n, precision="bf16") for b, n in clips]
    # host_pack="bf16": the float32 features are rounded to bfloat16 on host threads, one clip ahead, instead of on the device -
    # the same rounding, so not a bit of the result changes (7 clips: the ring of three staging buffers wraps twice)
    for graphed, pack in ((False, None), (True, None), (False, "bf16"), (True, "bf16")):
        m = spa.TrackAutoEncoder3D(**fields)
        m.cuda_graph = graphed
        rep = 1 if pack is None else 2
        got = list(m.apply_stream(variables, (batches * rep)[:7], noises=(noises * rep)[:7], precision="bf16", host_pack=pack, pack_threads=3))
        assert len(got) == min(7, len(refs) * rep)
        for g, r in zip(got, refs * rep):
            assert not g.tracks.is_cuda
            assert torch.equal(g.tracks, r.tracks.cpu()) and torch.equal(g.visible_logits, r.visible_logits.cpu())
    # the staging thread allocates pinned and device memory WHILE the consumer thread captures the forward (a fresh model captures inside the
    # stream): the capture must not be invalidated by another thread's CUDA calls (the capture is begun in relaxed mode)
    real_pack = spa.ops.host_pack_bf16
    hoard = []

    def noisy_pack(src, dst, threads):
        hoard.append(torch.empty(1 << 20).pin_memory())                       # cudaHostAlloc
        hoard.append(torch.empty((48 << 20) + len(hoard), device="cuda", dtype=torch.uint8))   # a fresh cudaMalloc
        return real_pack(src, dst, threads)

    spa.ops.host_pack_bf16 = noisy_pack
    try:
        for _ in range(3):
            m = spa.TrackAutoEncoder3D(**fields)
            m.cuda_graph = True
            got = list(m.apply_stream(variables, batches, noises=noises, precision="bf16", host_pack="bf16", pack_threads=2))
            for g, r in zip(got, refs):
                assert torch.equal(g.tracks, r.tracks.cpu())
    finally:
        spa.ops.host_pack_bf16 = real_pack
        del hoard
    # a consumer that stops early releases the staging thread
    gen = model.apply_stream(variables, batches, noises=noises, host_pack="bf16")
    next(gen)
    gen.close()
    import threading, time
    time.sleep(0.3)
    assert not any(t.name == "spa3d-host-pack" and t.is_alive() for t in threading.enumerate())
    # one clip, zero clips
    assert len(list(model.apply_stream(variables, batches[:1], noises=noises[:1]))) == 1
    assert list(model.apply_stream(variables, [], noises=[])) == []
    # bfloat16 features on the host (half the upload): same result up to the rounding the bf16 path applies anyway
    lowp = [dict(b, dino_features=b["dino_features"].to(torch.bfloat16).pin_memory(), depth_features=b["depth_features"].to(torch.bfloat16).pin_memory())
            for b in batches]
    got = list(model.apply_stream(variables, lowp, noises=noises, precision="bf16"))
    for g, r in zip(got, refs):
        assert rel_err(g.tracks, r.tracks) < 1e-2, rel_err(g.tracks, r.tracks)


def test_apply_stream_from_maps_equals_apply_from_maps(spa):
    rs = np.random.RandomState(13)
    N, T, Q, H, W, Hp, Wp = 40, 150, 5, 20, 24, 5, 6
    model = spa.TrackAutoEncoder3D()
    variables = model.init(22, {"dino_features": 1, "depth_features": 1})
    batches = []
    for i in range(3):
        batches.append({"support_tracks_2d": np.stack([rs.uniform(-2, W + 1, (N, T)), rs.uniform(-2, H + 1, (N, T))], -1).astype(np.float32),
                        "support_tracks_visible": (rs.uniform(size=(N, T, 1)) < 0.85).astype(np.float32),
                        "depth": rs.uniform(0.5, 4.0, (T, H, W, 1)).astype(np.float32), "dino_map": rs.standard_normal((T, Hp, Wp, 768)).astype(np.float32),
                        "video_shape": (T, H, W, 3),
                        "query_points": np.concatenate([rs.randint(0, T, (1, Q, 1)).astype(np.float32), rs.uniform(-1, 1, (1, Q, 3)).astype(np.float32)], -1)})
    noise = rs.uniform(size=(1, 128, 96)).astype(np.float32)
    refs = [model.apply_from_maps(variables, b, noise=noise) for b in batches]
    got = list(model.apply_stream(variables, batches, noises=[noise] * 3, from_maps=True, host_pack=None))
    for g, r in zip(got, refs):
        assert torch.equal(g.tracks, r.tracks.cpu()) and torch.equal(g.visible_logits, r.visible_logits.cpu())
    for pack in (None, "bf16"):   # the DINO patch map rounded to bf16 on the host instead of on the device: bit-identical
        gm = spa.TrackAutoEncoder3D()
        gm.cuda_graph = True          # the maps forward of each device buffer set replayed as one CUDA graph
        got = list(gm.apply_stream(variables, batches + batches, noises=[noise] * 6, from_maps=True, host_pack=pack, pack_threads=2))
        assert len(gm._graphs) == 2
        for g, r in zip(got, refs + refs):
            assert torch.equal(g.tracks, r.tracks.cpu()) and torch.equal(g.visible_logits, r.visible_logits.cpu())
