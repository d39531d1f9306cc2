"""End-to-end parity of the CUDA path (through the module mirror -> C ABI) against the oracle.

Tolerances are the north_star's: per-tensor max|a-b|/max|b| <= 1e-4 for the fp32 path and <= 2e-2
for the bf16 path, on reconstructed tracks and visibility logits.
"""
import numpy as np
import pytest
import torch

from oracle import model as om
from tests.helpers import SMALL_ARCH, make_inputs, product, rel_err, small_cfg

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
    ref = oracle_fwd(model, variables["params"], inp, noise, discretize)
    if precision == "bf16" and discretize:
        # a bf16-sized perturbation may flip a round(x*128): compare against the no-flip budget
        tol = 5e-2
    else:
        tol = TOL[precision]
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
    got = model.apply(variables, inp, noise=noise, discretize=False, precision=precision)
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
