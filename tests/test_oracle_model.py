"""Known-answer / property pins for the model oracle.

The reference ships no golden vectors for the model (PARITY UNPINNED), so these pins are the
ones derivable from the reference source itself.
"""
import math

import numpy as np
import pytest
import torch

from oracle import model as om
from tests.helpers import SMALL_ARCH, make_inputs, rel_err, small_cfg


@pytest.fixture(scope="module")
def small():
    cfg = small_cfg()
    p = om.init_params_3d(cfg, seed=1, randomize_norms=True, arch=SMALL_ARCH)
    return cfg, p


def fwd(cfg, p, inp, noise, dtype=torch.float32, discretize=True):
    with torch.no_grad():
        return om.forward_3d(om.to_torch(p, dtype), cfg, om.cast_inputs(inp, dtype), torch.as_tensor(noise).to(dtype), discretize)


def test_sinusoid_kat():
    # track_autoencoder.py:28-37: zeros -> 32 zeros then 32 ones per coordinate, coords outermost
    e = om.sinusoidal_embedding(torch.zeros(2, 3))
    assert e.shape == (2, 192)
    blk = e.reshape(2, 3, 64)
    assert torch.all(blk[..., :32] == 0) and torch.all(blk[..., 32:] == 1)
    x = torch.tensor([[0.3, -0.7]])
    e = om.sinusoidal_embedding(x).reshape(2, 64)
    s = torch.tensor([2 ** (i / 3) for i in range(32)], dtype=torch.float32)
    for c in range(2):
        arg = x[0, c] * s
        np.testing.assert_allclose(e[c, :32], torch.sin(arg.double()).float(), atol=1e-7)
        np.testing.assert_allclose(e[c, 32:], torch.sin((arg + torch.tensor(0.5 * math.pi)).double()).float(), atol=1e-7)


def test_append_time_feat_is_a_window():
    # track_autoencoder_3d.py:235-246: eye(128, C, 5t) einsum == lat[..., 5t:5t+128]
    lat = torch.randn(2, 3, 4, 1152)
    qf = torch.tensor([[0, 7, 149], [1, 50, 100]], dtype=torch.int32)
    out = om.append_time_feat(lat, qf)
    assert out.shape == (2, 3, 4, 1280)
    for b in range(2):
        for q in range(3):
            t = int(qf[b, q])
            assert torch.equal(out[b, q, :, 1152:], lat[b, q, :, 5 * t : 5 * t + 128])
    assert 5 * 149 + 127 < 1152


def test_query_time_feature_is_zero():
    # track_autoencoder_3d.py:268-269: int32 // 150.0 == 0 for frames 0..149 (defect D6, reproduced)
    qf = torch.arange(150, dtype=torch.int32)
    assert torch.all(torch.floor(qf.float() / 150.0) == 0)


def test_round_half_even():
    assert om.round_half_even(torch.tensor([0.5, 1.5, 2.5, -0.5])).tolist() == [0.0, 2.0, 2.0, -0.0]


def test_outputs_layout_and_certain_zero(small):
    cfg, p = small
    inp, noise = make_inputs(cfg)
    r = fwd(cfg, p, inp, noise)
    B, Q, T = 2, 6, cfg.num_output_frames
    assert r.tracks.shape == (B, Q, T, 3) and r.visible_logits.shape == (B, Q, T, 1)
    assert torch.all(r.certain_logits == 0)


def test_fp32_matches_fp64(small):
    cfg, p = small
    inp, noise = make_inputs(cfg)
    a, b = fwd(cfg, p, inp, noise, torch.float32, False), fwd(cfg, p, inp, noise, torch.float64, False)
    assert rel_err(a.tracks, b.tracks) < 2e-5 and rel_err(a.visible_logits, b.visible_logits) < 2e-5


def test_chunked_decode_equals_unchunked(small):
    cfg, p = small
    inp, noise = make_inputs(cfg)
    a = fwd(cfg, p, inp, noise)
    cfg2 = small_cfg(decoder_scan_chunk_size=2)
    b = fwd(cfg2, p, inp, noise)
    assert rel_err(a.tracks, b.tracks) < 1e-5


def test_queries_are_independent(small):
    cfg, p = small
    inp, noise = make_inputs(cfg)
    a = fwd(cfg, p, inp, noise)
    inp2 = dict(inp)
    inp2["query_points"] = inp["query_points"][:, :3]
    b = fwd(cfg, p, inp2, noise)
    assert rel_err(a.tracks[:, :3], b.tracks) < 1e-5


def test_support_permutation_invariance(small):
    cfg, p = small
    inp, noise = make_inputs(cfg)
    perm = np.random.RandomState(0).permutation(inp["support_tracks"].shape[1])
    inp2 = dict(inp)
    for k in ("support_tracks", "support_tracks_visible", "dino_features", "depth_features"):
        inp2[k] = inp[k][:, perm]
    a, b = fwd(cfg, p, inp, noise, torch.float64, False), fwd(cfg, p, inp2, noise, torch.float64, False)
    assert rel_err(a.tracks, b.tracks) < 1e-9


def test_masked_frames_do_not_matter(small):
    # R1: invisible frames and frames >= boundary_frame never reach the readout token
    cfg, p = small
    inp, noise = make_inputs(cfg)
    T = cfg.num_output_frames
    inp["boundary_frame"] = np.array([T - 3, T], np.int32)
    inp2 = {k: np.array(v) for k, v in inp.items()}
    vis = inp["support_tracks_visible"][..., 0] > 0
    dead = ~vis
    dead[0, :, T - 3 :] = True
    rs = np.random.RandomState(5)
    for k in ("support_tracks", "dino_features", "depth_features"):
        junk = rs.standard_normal(inp[k].shape).astype(np.float32)
        inp2[k] = np.where(dead[..., None], junk, inp[k])
    a, b = fwd(cfg, p, inp, noise, torch.float64, False), fwd(cfg, p, inp2, noise, torch.float64, False)
    assert rel_err(a.tracks, b.tracks) < 1e-9


def test_fully_masked_row_is_uniform():
    # attention.py:175: mask value is finfo.min, not -inf => all-masked rows give uniform weights
    rs = np.random.RandomState(0)
    p = om.to_torch(om._attn_init(rs, 16, 16, 2, 8), torch.float64)
    x = torch.randn(1, 5, 16, dtype=torch.float64)
    mask = torch.zeros(1, 1, 5, 5)
    out = om.mhdp_attention(p, x, x, mask)
    v = torch.einsum("...kd,dhc->...khc", x, p["dense_value"]["kernel"]).mean(dim=1, keepdim=True)
    expect = torch.einsum("...qhd,hdo->...qo", v.expand(1, 5, 2, 8), p["dense_out"]["kernel"]) + p["dense_out"]["bias"]
    assert rel_err(out, expect) < 1e-12


def test_loss_closed_forms():
    B, Q, T = 2, 3, 4
    pred = om.Results(torch.ones(B, Q, T, 3), torch.zeros(B, Q, T, 1), torch.zeros(B, Q, T, 1))
    tgt = {"query_tracks": torch.zeros(B, Q, T, 3), "query_tracks_visible": torch.ones(B, Q, T, 1)}
    l = om.compute_loss_3d(pred, tgt)
    assert abs(float(l["position_loss"]) - 3.0) < 1e-6  # sum over xyz / visible count
    assert abs(float(l["visible_loss"]) - math.log(2)) < 1e-6
    tgt["query_tracks_visible"] = torch.zeros(B, Q, T, 1)
    l = om.compute_loss_3d(pred, tgt)
    assert float(l["position_loss"]) == 0.0
    assert abs(float(l["visible_loss"]) - B * Q * T * math.log(2)) < 1e-5  # summed over ALL points / max(0,1)
    assert abs(float(l["total_loss"]) - 1e-8 * B * Q * T * math.log(2)) < 1e-12


def test_lr_schedule():
    assert om.learning_rate(0) == 0.0
    assert abs(om.learning_rate(5000) - 5e-5) < 1e-12
    assert abs(om.learning_rate(10000) - 1e-4) < 1e-12
    assert abs(om.learning_rate(1000000)) < 1e-12


def test_param_counts_match_survey():
    # 109.14 M (3DSPA, R2 widths) and 68.33 M (TRAJAN)
    assert om.count_params(om.init_params_3d(om.Config3D())) == 109_138_296
    assert om.count_params(om.init_params_2d(om.Config2D())) == 68_333_080


def test_trajan_2d_runs_and_masks():
    cfg = om.Config2D(num_output_frames=10, num_latent_tokens=8, latent_token_dim=16, track_token_dim=32,
                      encoder_latent_dim=48, decoder_num_channels=128 + 48)
    p = om.to_torch(om.init_params_2d(cfg, seed=2, randomize_norms=True, arch=SMALL_ARCH), torch.float64)
    inp, noise = make_inputs(cfg, B=1, N=5, Q=4, T=10, dino=False, depth=False, coords=2)
    with torch.no_grad():
        r = om.forward_2d(p, cfg, om.cast_inputs(inp, torch.float64), torch.as_tensor(noise).double())
    assert r.tracks.shape == (1, 4, 10, 2) and r.certain_logits.shape == (1, 4, 10, 1)
    assert r.certain_logits.abs().max() > 0  # TRAJAN does predict certainty (track_autoencoder.py:339)


def test_adamw_matches_torch():
    torch.manual_seed(0)
    w = [torch.randn(5, 3), torch.randn(7)]
    g = [torch.randn(5, 3) * 3, torch.randn(7) * 3]
    ref = [t.clone().requires_grad_(True) for t in w]
    opt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m = [torch.zeros_like(t) for t in w]
    v = [torch.zeros_like(t) for t in w]
    mine = [t.clone() for t in w]
    for step in range(1, 4):
        for r, gg in zip(ref, g):
            r.grad = gg.clone()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        om.adamw_step(mine, [gg.clone() for gg in g], m, v, step, 1e-3)
    for a, b in zip(mine, ref):
        assert rel_err(a, b) < 1e-5
