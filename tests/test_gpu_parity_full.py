"""Parity at BASELINE.json sizes and widths, through the module mirror -> C ABI, against the oracle.

  cfg2   B=1, T=150, 2048 support / 512 query tracks, DINO 768 + depth 256, the real 109.14 M-parameter architecture,
         bf16 path with the quantiser ON (the configuration bench.py times) vs the fp32 oracle at the north_star's 2e-2;
  fp32   the accurate path at 256 / 64 tracks, quantiser on, at 1e-4;
  cfg5   one clip of the sweep's shape (518x518 video, 37x37x768 patch map) through apply_from_maps at 512 / 128 tracks;
  grads  bf16 backward at the real widths (Dh = 96, L = 151 / 129 / 128: the attention-backward and TN weight-gradient
         tile shapes the training bench runs) vs float64 autograd through the oracle.

Metric: per tensor max|a-b| / max|b|.  Quantiser-on comparisons use
tests.helpers.quantiser_aware_reference (latents held to the tolerance; decoder compared on identical round() decisions);
the plain, flip-unaware error is measured too and bounded.
"""
import importlib

import numpy as np
import pytest
import torch

from oracle import lifting as ol
from oracle import model as om
from tests.helpers import make_inputs, product, quantiser_aware_reference, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def spa():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return product()


def _real_model(spa, seed, inp):
    model = spa.TrackAutoEncoder3D()
    variables = model.init(seed, inp)
    om._randomize(variables["params"], np.random.RandomState(seed))   # norm scales / biases off their 1 / 0 defaults
    return model, variables


def _check_quantised(model, variables, inp, noise, precision, tol, plain_tol, apply=None):
    cfg = om.Config3D()
    got = (apply or (lambda: model.apply(variables, inp, noise=noise, discretize=True, precision=precision)))()
    z_gpu = model.apply(variables, inp, method="encode", precision=precision)
    ref_plain, ref_aware, z_ref, flipped = quantiser_aware_reference(variables["params"], cfg, inp, noise, z_gpu)
    ez = rel_err(z_gpu, z_ref)
    e_t, e_v = rel_err(got.tracks, ref_aware.tracks), rel_err(got.visible_logits, ref_aware.visible_logits)
    p_t, p_v = rel_err(got.tracks, ref_plain.tracks), rel_err(got.visible_logits, ref_plain.visible_logits)
    print(f"[{precision}] latents {ez:.2e}; flipped round() decisions {100 * flipped:.2f} %; same-rounding tracks {e_t:.2e} "
          f"logits {e_v:.2e}; plain tracks {p_t:.2e} logits {p_v:.2e}")
    assert ez < tol, ("latents", ez)
    assert e_t < tol and e_v < tol, ("same rounding decisions", e_t, e_v)
    assert p_t < plain_tol and p_v < plain_tol, ("plain", p_t, p_v)
    assert not got.certain_logits.any()
    return got


@pytest.mark.parametrize("residual_dtype", ["bf16", "fp32"])
def test_cfg2_full_size_bf16_quantiser_on(spa, residual_dtype):
    """BASELINE.json configs[1] exactly: the bf16 path bench.py times, quantiser on, vs the fp32 oracle - with the residual stream of
    the transformer stacks in bfloat16 (the default of the bf16 precision; measured 7.6e-3) and in float32 (6.1e-3)."""
    c = om.Config3D()
    inp, noise = make_inputs(c, B=1, N=2048, Q=512, seed=21, vis_p=0.9)
    model, variables = _real_model(spa, 22, inp)
    model.residual_dtype = residual_dtype
    assert model.bind(variables, "bf16").sdt == (torch.bfloat16 if residual_dtype == "bf16" else torch.float32)
    got = _check_quantised(model, variables, inp, noise, "bf16", 2e-2, 3e-2)
    assert got.tracks.shape == (1, 512, 150, 3) and got.visible_logits.shape == (1, 512, 150, 1)


def test_fp32_path_256_64_quantiser_on(spa):
    """The accurate path at a size the SIMT kernels finish quickly: 1e-4 on everything, quantiser on."""
    c = om.Config3D()
    inp, noise = make_inputs(c, B=1, N=256, Q=64, seed=23, vis_p=0.9)
    inp["boundary_frame"] = np.array([140], np.int32)
    model, variables = _real_model(spa, 24, inp)
    _check_quantised(model, variables, inp, noise, "fp32", 1e-4, 1e-2)


def test_cfg5_clip_through_maps_512_128(spa):
    """One clip of the realism sweep's shape through the fused maps path (lift -> sample -> embed never materialise the
    per-track features) vs the oracle pipeline: NumPy lifting restatement (bit-exact vs inference.py) + model restatement."""
    rs = np.random.RandomState(31)
    S5, Q5, T, H, W, Hp, Wp = 512, 128, 150, 518, 518, 37, 37
    tr2 = np.stack([rs.uniform(-4, W + 3, (S5 + Q5, T)), rs.uniform(-4, H + 3, (S5 + Q5, T))], -1).astype(np.float32)
    depth = rs.uniform(0.5, 4.0, (T, H, W, 1)).astype(np.float32)
    dino_map = rs.standard_normal((T, Hp, Wp, 768)).astype(np.float32)
    visible = (rs.uniform(size=(S5, T, 1)) < 0.9).astype(np.float32)
    video_shape = (T, H, W, 3)
    xyz = ol.lift_2d_to_3d(tr2, depth)
    # keep coordinates in the range the Fourier features are specified for (U(-1,1)^3): scale depth so |xyz| <~ 1
    s = 1.0 / max(1.0, float(np.abs(xyz).max()))
    depth = (depth * s).astype(np.float32)
    xyz = ol.lift_2d_to_3d(tr2, depth)
    dfe = ol.sample_dino_features_for_tracks(dino_map, tr2[:S5], video_shape)
    zfe = ol.sample_depth_features_for_tracks(depth, tr2[:S5])
    qt = rs.randint(0, T, Q5)
    qp = np.concatenate([qt[:, None].astype(np.float32), xyz[S5:][np.arange(Q5), qt]], -1)[None]
    feat_inputs = {"support_tracks": xyz[None, :S5], "support_tracks_visible": visible[None], "query_points": qp,
                   "boundary_frame": np.array([T], np.int32), "dino_features": dfe[None], "depth_features": zfe[None]}
    maps_inputs = {"support_tracks_2d": tr2[:S5], "support_tracks_visible": visible, "depth": depth, "dino_map": dino_map,
                   "video_shape": video_shape, "query_points": qp}
    noise = rs.uniform(size=(1, 128, 96)).astype(np.float32)
    model, variables = _real_model(spa, 32, feat_inputs)
    cfg = om.Config3D()
    got = model.apply_from_maps(variables, maps_inputs, noise=noise, discretize=True)
    eng = model.bind(variables, "bf16")
    with torch.no_grad():
        z_gpu, xyz_dev = eng.encode_from_maps(maps_inputs)
    np.testing.assert_array_equal(xyz_dev.cpu().numpy(), xyz[:S5])       # lifting is bit-exact
    ref_plain, ref_aware, z_ref, flipped = quantiser_aware_reference(variables["params"], cfg, feat_inputs, noise, z_gpu)
    ez = rel_err(z_gpu, z_ref)
    e_t, e_v = rel_err(got.tracks, ref_aware.tracks), rel_err(got.visible_logits, ref_aware.visible_logits)
    print(f"[maps] latents {ez:.2e}; flipped {100 * flipped:.2f} %; tracks {e_t:.2e} logits {e_v:.2e}; "
          f"plain tracks {rel_err(got.tracks, ref_plain.tracks):.2e}")
    assert ez < 2e-2 and e_t < 2e-2 and e_v < 2e-2, (ez, e_t, e_v)


def _cos(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = b.detach().double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def test_bf16_gradients_at_reference_widths(spa):
    """bf16 backward on the tile shapes the training bench runs - attn_bwd_tc_kernel at L = 151 (per-track transformer),
    129 (read-out) and 128 (latent transformers), Dh = 96, the TN weight-gradient GEMM at K = 384 / 768 / 1280 / 1536 -
    vs float64 autograd through the oracle: per parameter leaf, cosine >= 0.995 and max-error / max <= 5e-2.
    A smooth functional of the head output is differentiated (the L1 loss gradient is a sign: see test_gpu_train.py)."""
    te = importlib.import_module("3dspa_code_b200.train_engine")
    c = om.Config3D()
    B, N, Q, T = 1, 24, 8, 150
    inp, noise = make_inputs(c, B=B, N=N, Q=Q, seed=41, targets=True, vis_p=0.9)
    inp["boundary_frame"] = np.array([T - 5], np.int32)
    model, variables = _real_model(spa, 42, inp)
    tree = variables["params"]
    R = torch.randn(B * Q, 4 * T, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    p = om.to_torch(tree, torch.float64, requires_grad=True)
    ci = om.cast_inputs(inp, torch.float64)
    res = om.forward_3d(p, c, ci, torch.as_tensor(noise).double(), False)
    head = torch.cat([res.tracks[..., 0], res.tracks[..., 1], res.tracks[..., 2], res.visible_logits[..., 0]], dim=-1)
    (head.reshape(B * Q, 4 * T) * R).sum().backward()
    gref = {k: v.grad for k, v in om.flatten(p).items()}
    store = te.ParamStore(tree, "bf16")
    eng = te.TrainEngine(model, store)
    with torch.enable_grad():
        out = eng.forward_train(inp, noise, discretize=False)
    assert rel_err(out, head.reshape(B * Q, 4 * T)) < 2e-2
    out.backward(R.float().cuda())
    got = spa.params.flatten(store.grad_tree())
    assert got.keys() == gref.keys()
    stats = {k: (_cos(got[k], g), rel_err(got[k], g)) for k, g in gref.items()}
    worst_cos = min(stats.items(), key=lambda kv: kv[1][0])
    worst_rel = max(stats.items(), key=lambda kv: kv[1][1])
    print(f"[grads bf16, real widths] worst cosine {worst_cos[1][0]:.5f} ({worst_cos[0]}), worst rel err {worst_rel[1][1]:.3e} ({worst_rel[0]})")
    total = float(sum(float(g.double().pow(2).sum()) for g in gref.values()) ** 0.5)
    share = {k: float(g.double().norm()) / total for k, g in gref.items()}
    for k, v in sorted(stats.items(), key=lambda kv: kv[1][0])[:12]:
        print(f"    {k:70s} cos {v[0]:.5f} rel {v[1]:.3e} share-of-gradient-norm {share[k]:.2e}")

    # The logit path of the latents<-tracks cross-attention (its query / key projections and their RMSNorm scales, and the
    # latent initialiser through them) is ill-conditioned at random init: the 24 read-out tokens of the per-track transformer
    # are nearly identical, attention over them is nearly uniform, and the gradient with respect to q and k is a sum of
    # DIFFERENCES of nearly equal bf16 keys (measured: these leaves carry 1e-4 .. 5e-6 of the gradient norm).  Any bf16
    # evaluation loses digits there; the kernels themselves are held to 2e-2 on well-conditioned inputs in
    # tests/test_gpu_kernels.py::test_attention_bwd (same tile shapes).  They get a looser, still meaningful bound; every
    # other leaf is held to cosine >= 0.995 and max-error / max <= 5e-2.
    def ill_conditioned(k):
        return k == "initializer/state_init" or ("tracks_to_latents" in k and "cross_att" in k and
                                                 any(t in k for t in ("dense_query", "dense_key", "norm_query", "norm_key")))

    bad = {k: v for k, v in stats.items()
           if v[0] < (0.98 if ill_conditioned(k) else 0.995) or v[1] > (0.35 if ill_conditioned(k) else 5e-2)}
    assert not bad, sorted(bad.items(), key=lambda kv: kv[1][0])[:10]
    assert all(share[k] < 2e-3 for k in stats if ill_conditioned(k))
    # the gradient as ONE vector (what the optimiser's global-norm clip and update see)
    dot = sum(float((torch.as_tensor(np.asarray(got[k])).double().reshape(-1) * g.double().reshape(-1)).sum()) for k, g in gref.items())
    gn = float(sum(float(torch.as_tensor(np.asarray(got[k])).double().pow(2).sum()) for k in gref) ** 0.5)
    err = float(sum(float((torch.as_tensor(np.asarray(got[k])).double() - g.double()).pow(2).sum()) for k, g in gref.items()) ** 0.5)
    print(f"    whole gradient: cosine {dot / (gn * total):.6f}, relative L2 error {err / total:.3e}, norm ratio {gn / total:.5f}")
    assert dot / (gn * total) > 0.9995 and err / total < 3e-2
