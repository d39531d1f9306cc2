"""Shared synthetic-input builders for the tests (SURVEY.md 8d recipes, scaled down)."""
import importlib

import numpy as np
import torch

from oracle import model as om

SMALL_ARCH = {"itt": (64, 2, 96, 2), "t2l": (64, 2, 128, 2), "dec": (64, 2, 128, 2), "tra": (64, 2, 96, 2)}


def small_cfg(**kw):
    base = dict(num_output_frames=12, num_latent_tokens=16, latent_token_dim=24, num_frequencies=32,
                track_token_dim=48, encoder_latent_dim=64, decoder_num_channels=128 + 80,
                dino_feature_dim=40, depth_feature_dim=16)
    base.update(kw)
    return om.Config3D(**base)


def make_inputs(cfg, B=2, N=10, Q=6, T=None, seed=0, dino=True, depth=True, vis_p=0.8, targets=False, coords=3):
    T = T or cfg.num_output_frames
    rs = np.random.RandomState(seed)
    inp = {
        "support_tracks": rs.uniform(-1, 1, (B, N, T, coords)).astype(np.float32),
        "support_tracks_visible": (rs.uniform(size=(B, N, T, 1)) < vis_p).astype(np.float32),
        "query_points": np.concatenate(
            [rs.randint(0, T, (B, Q, 1)).astype(np.float32), rs.uniform(-1, 1, (B, Q, coords)).astype(np.float32)], -1),
        "boundary_frame": np.full((B,), T, np.int32),
    }
    if dino:
        inp["dino_features"] = rs.standard_normal((B, N, T, cfg.dino_feature_dim)).astype(np.float32)
    if depth:
        inp["depth_features"] = rs.standard_normal((B, N, T, cfg.depth_feature_dim)).astype(np.float32)
    if targets:
        inp["query_tracks"] = rs.uniform(-1, 1, (B, Q, T, coords)).astype(np.float32)
        inp["query_tracks_visible"] = (rs.uniform(size=(B, Q, T, 1)) < vis_p).astype(np.float32)
    noise = rs.uniform(size=(B, cfg.num_latent_tokens, cfg.latent_token_dim)).astype(np.float32)
    return inp, noise


def rel_err(a, b):
    """Per-tensor max|a-b| / max|b| (the metric DESIGN.md fixes for the north_star tolerances)."""
    a = a.detach().cpu().double() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def product():
    return importlib.import_module("3dspa_code_b200")


def quantiser_aware_reference(params, cfg, inputs, noise, z_gpu):
    """Reference outputs for a forward with the quantiser ON (track_autoencoder_3d.py:251-260).

    round(z * 128) is discontinuous: a bf16-sized difference in a latent that sits next to a rounding boundary flips the
    decision and moves that latent by 1/128, which says nothing about kernel parity.  So the comparison is made in two parts:
      (1) the latents themselves (continuous) are held to the tolerance by the caller (``z_ref`` is returned);
      (2) the decoder is compared on identical rounding decisions: the oracle decodes ``where(flipped, z_gpu, z_ref)``,
          i.e. its own latents everywhere the two paths round the same way and the device value at the flipped positions
          (the forward value of the straight-through quantiser depends on the latent only through round()).
    Returns (ref_plain, ref_aware, z_ref, flipped_fraction): ``ref_plain`` is the oracle with its own rounding everywhere."""
    ci = om.cast_inputs(inputs, torch.float32)
    p = om.to_torch(params)
    nz = torch.as_tensor(np.asarray(noise)).float()
    with torch.no_grad():
        z_ref = om.encode_3d(p, cfg, ci)
        ctx = om.get_decoder_context(cfg, ci)
        ref_plain = om.decode_3d(p, cfg, z_ref, ctx, nz, True)
        zg = z_gpu.detach().cpu().float()
        flipped = torch.round(zg.clamp(-1, 1) * 128.0) != torch.round(z_ref.clamp(-1, 1) * 128.0)
        ref_aware = om.decode_3d(p, cfg, torch.where(flipped, zg, z_ref), ctx, nz, True)
    return ref_plain, ref_aware, z_ref, float(flipped.float().mean())
