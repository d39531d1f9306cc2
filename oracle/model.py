"""CPU restatement (torch, fp32 or fp64) of the reference model path.  TEST INFRASTRUCTURE.

Pinned (tests/test_oracle_golden_model.py, 1e-9 in float64) against golden vectors produced by
executing the reference's own model files on NumPy stand-ins for the absent jax / flax / optax
primitives (oracle/flax_shim.py, tests/golden/make_golden_model.py).  Not pinned: bit-level
agreement with XLA's kernels and the threefry noise stream; see ``oracle/__init__.py``.
Every function cites the reference lines it follows (paths relative to /root/reference).  Third-party semantics (Flax ``Dense``,
``LayerNorm``, ``RMSNorm``, ``dot_product_attention``, ``gelu``; flax>=0.7.5, jax>=0.4.20 per
``requirements.txt:2-6``, lower bounds only) are restated from their published definitions
(SURVEY.md Appendix B).

Parameters are nested dicts of torch tensors with the Flax tree naming.
All maths runs in the dtype of the parameters (float32 = "what the reference computes",
float64 = ground truth).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np
import torch

# ----------------------------------------------------------------------------------------
# Flax primitives
# ----------------------------------------------------------------------------------------

LN_EPS = 1e-6  # flax.linen.LayerNorm / RMSNorm default epsilon


def dense(x, p):
    """flax.linen.Dense: y = x @ kernel[in, out] + bias."""
    return x @ p["kernel"] + p["bias"]


def layer_norm(x, scale):
    """nn.LayerNorm(use_bias=False, use_scale=True) (attention.py:49,76,103).

    Flax computes the "fast" variance max(0, E[x^2] - E[x]^2), eps = 1e-6.
    """
    mu = x.mean(dim=-1, keepdim=True)
    ms = (x * x).mean(dim=-1, keepdim=True)
    var = torch.clamp(ms - mu * mu, min=0.0)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * scale


def rms_norm(x, scale):
    """nn.RMSNorm() over the last axis only, i.e. per head (attention.py:166-167)."""
    ms = (x * x).mean(dim=-1, keepdim=True)
    return x * torch.rsqrt(ms + LN_EPS) * scale


def gelu_tanh(x):
    """nn.gelu default (approximate=True), attention.py:106."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x)))


def sinusoidal_embedding(x, num_frequencies=32, scale_factor=1.0):
    """SinusoidalEmbedding.__call__ (track_autoencoder.py:18-38).

    scales = float32(2**(i/3)); args = [x*s, x*s + float32(0.5*pi)] evaluated in float32
    exactly as the reference does (one rounding per op); out[..., c, :] = sin(args);
    flattened "(coords d)".

    Conditioning note: arguments reach 2**(31/3) ~ 1290 x |coord|, so one float32 ulp of the
    argument moves the feature by ~1e-4, and in the decoder the embedding is applied to the
    *output* of a first embedding (track_autoencoder_3d.py:265-275), so even a 1-ulp
    difference between two libm ``sinf`` implementations is amplified to ~1e-4.  To make the
    feature well defined the oracle uses the correctly rounded float32 sine,
    float32(sin(float64(arg))), in BOTH the fp32 and the fp64 instance: the Fourier features
    are data (no gradient flows into coordinates) and are bit-identical between the two.
    """
    x32 = x.detach().to(torch.float32) / scale_factor  # "tracks / track_scale_factor" in float32
    scales = torch.tensor([2 ** (i / 3) for i in range(num_frequencies)], dtype=torch.float32)
    xs = x32[..., None] * scales  # [..., coords, F], float32 product
    half_pi = torch.tensor(0.5 * math.pi, dtype=torch.float32)
    arg = torch.cat([xs, xs + half_pi], dim=-1)  # [..., coords, 2F]
    out = torch.sin(arg.to(torch.float64)).to(torch.float32).to(x.dtype)
    return out.reshape(*x.shape[:-1], x.shape[-1] * 2 * num_frequencies)


# ----------------------------------------------------------------------------------------
# attention.py
# ----------------------------------------------------------------------------------------


def mhdp_attention(p, inputs_q, inputs_kv, mask=None, num_heads=8):
    """ImprovedMHDPAttention.__call__ (attention.py:124-185).

    mask: broadcastable to [..., H, Lq, Lk]; entries != 0 keep the logit, others are
    replaced by finfo(dtype).min (nn.dot_product_attention semantics).
    """
    wq, wk, wv = p["dense_query"]["kernel"], p["dense_key"]["kernel"], p["dense_value"]["kernel"]
    q = torch.einsum("...qd,dhk->...qhk", inputs_q, wq)
    k = torch.einsum("...kd,dhc->...khc", inputs_kv, wk)
    q = rms_norm(q, p["norm_query"]["scale"])
    k = rms_norm(k, p["norm_key"]["scale"])
    v = torch.einsum("...kd,dhc->...khc", inputs_kv, wv)
    depth = q.shape[-1]
    q = q / math.sqrt(depth)
    w = torch.einsum("...qhd,...khd->...hqk", q, k)
    if mask is not None:
        big_neg = torch.finfo(w.dtype).min
        w = torch.where(mask != 0, w, torch.full_like(w, big_neg))
    w = torch.softmax(w, dim=-1)
    x = torch.einsum("...hqk,...khd->...qhd", w, v)
    out = torch.einsum("...qhd,hdo->...qo", x, p["dense_out"]["kernel"]) + p["dense_out"]["bias"]
    return out


def transformer_block(p, queries, inputs_kv=None, qq_mask=None, qk_mask=None):
    """ImprovedTransformerBlock.__call__ (attention.py:66-108)."""
    normed = layer_norm(queries, p["norm_q"]["scale"])
    attn_out = queries
    attn_out = attn_out + mhdp_attention(p["self_att"], normed, normed, qq_mask)
    if inputs_kv is not None:
        attn_out = attn_out + mhdp_attention(p["cross_att"], normed, inputs_kv, qk_mask)
    normed2 = layer_norm(attn_out, p["norm_attn"]["scale"])
    h = gelu_tanh(dense(normed2, p["MLP_in"]))
    return attn_out + dense(h, p["MLP_out"])


def improved_transformer(p, queries, inputs_kv=None, qk_mask=None, qq_mask=None):
    """ImprovedTransformer.__call__ (attention.py:22-53).  Masks get a head axis."""
    num_layers = sum(1 for k in p if k.startswith("layer_"))
    if qk_mask is not None and qk_mask.dim() == inputs_kv.dim():
        qk_mask = qk_mask[..., None, :, :]
    if qq_mask is not None and qq_mask.dim() == queries.dim():
        qq_mask = qq_mask[..., None, :, :]
    for i in range(num_layers):
        queries = transformer_block(p[f"layer_{i}"], queries, inputs_kv, qq_mask, qk_mask)
    return layer_norm(queries, p["norm_encoder"]["scale"])


# ----------------------------------------------------------------------------------------
# track_autoencoder_3d.py  (with repairs R1 / R2)
# ----------------------------------------------------------------------------------------


@dataclass
class Results:
    """TrackAutoEncoderResults (track_autoencoder.py:72-105)."""

    tracks: torch.Tensor
    visible_logits: torch.Tensor
    certain_logits: torch.Tensor

    @property
    def visible(self):
        return (self.visible_logits > 0).to(torch.float32)

    @property
    def certain(self):
        return (self.certain_logits > 0).to(torch.float32)

    @property
    def visible_and_certain(self):
        return ((torch.sigmoid(self.visible_logits) * torch.sigmoid(self.certain_logits)) > 0.5).to(
            torch.float32
        )


@dataclass
class DecoderContext:
    """TrackAutoEncoderDecoderContext (track_autoencoder.py:108-114)."""

    decoder_query: torch.Tensor
    query_frame: torch.Tensor
    boundary_frame: Any


@dataclass
class Config3D:
    """Constructor fields of TrackAutoEncoder3D (track_autoencoder_3d.py:53-67)."""

    num_output_frames: int = 150
    num_latent_tokens: int = 128
    latent_token_dim: int = 96
    num_frequencies: int = 32
    track_scale_factor: float = 1.0
    time_scale_factor: float = 150.0
    track_token_dim: int = 384
    encoder_latent_dim: int = 512
    decoder_num_channels: int = 1280
    dino_feature_dim: int = 768
    depth_feature_dim: int = 256
    use_dino: bool = True
    use_depth: bool = True
    decoder_scan_chunk_size: Optional[int] = None


def embed_track_pos_visible_3d(p, cfg, tracks, visible, dino=None, depth=None):
    """embed_track_pos_visible (track_autoencoder_3d.py:123-149), repair R2 (widths)."""
    T = tracks.shape[-2]
    fr_id = (torch.arange(T, dtype=torch.float32) / T).to(tracks.dtype)  # jnp int/int -> f32
    fr_id = fr_id[None, None, :, None].expand(visible.shape)
    twt = torch.cat([tracks, fr_id], dim=-1)
    emb = sinusoidal_embedding(twt, cfg.num_frequencies, cfg.track_scale_factor)
    tok = dense(emb, p["track_token_projection"])
    if cfg.use_dino and dino is not None:
        tok = tok + dense(dino, p["dino_projection"])
    if cfg.use_depth and depth is not None:
        tok = tok + dense(depth, p["depth_projection"])
    return tok


def key_mask_3d(visible, boundary_frame):
    """Repair R1 of track_autoencoder_3d.py:167-184: key-only mask of length T+1.

    key 0 (readout) always on; key j>=1 on iff visible[b,n,j-1] and (j-1) < boundary[b].
    Returns bool [B, N, T+1].
    """
    B, N, T, _ = visible.shape
    time = torch.arange(T)
    partition = time[None, None, :] < boundary_frame.reshape(B, 1, 1)
    vis = visible[..., 0] != 0
    on = partition & vis
    return torch.cat([torch.ones(B, N, 1, dtype=torch.bool), on], dim=-1)


def encode_tracks_3d(p, cfg, tracks, visible, restart, dino=None, depth=None):
    """encode_tracks (track_autoencoder_3d.py:151-188)."""
    emb = embed_track_pos_visible_3d(p, cfg, tracks, visible, dino, depth)
    B, N, T, W = emb.shape
    readout = p["input_readout_token"]["state_init"].expand(B, N, 1, W)
    tokens = torch.cat([readout, emb], dim=-2)  # [B,N,T+1,W]
    km = key_mask_3d(visible, restart)  # [B,N,T+1]
    qq = km[:, :, None, None, :].expand(B, N, 1, T + 1, T + 1)  # head axis already present
    tokens = improved_transformer(p["input_track_transformer"], tokens, qq_mask=qq)
    return tokens[..., 0, :]


def encode_3d(p, cfg, inputs):
    """encode (track_autoencoder_3d.py:190-204)."""
    st = encode_tracks_3d(
        p,
        cfg,
        inputs["support_tracks"],
        inputs["support_tracks_visible"],
        inputs["boundary_frame"],
        inputs.get("dino_features"),
        inputs.get("depth_features"),
    )
    B = inputs["support_tracks"].shape[0]
    lat = p["initializer"]["state_init"].expand(B, *p["initializer"]["state_init"].shape)
    lat = improved_transformer(p["tracks_to_latents"], lat, st)
    return dense(lat, p["compressor"])


def round_half_even(x):
    return torch.round(x)  # torch.round is half-to-even, like jnp.round


def get_decoder_context(cfg, inputs, coords=3):
    """get_decoder_context (track_autoencoder_3d.py:206-233 / track_autoencoder.py:248-273)."""
    if "query_points" in inputs:
        qp = inputs["query_points"]
        decoder_query = qp[..., 1:]
        query_frame = round_half_even(qp[..., 0]).to(torch.int32)
    else:
        dt = inputs["support_tracks"].dtype
        gc = (torch.arange(32, dtype=torch.float32) / 32.0 + 1.0 / 64.0).to(dt)
        qx, qy = torch.meshgrid(gc, gc, indexing="xy")
        comps = [qx, qy] + ([torch.zeros_like(qx)] if coords == 3 else [])
        decoder_query = torch.stack(comps, dim=-1).reshape(-1, coords)
        decoder_query = decoder_query.expand(*inputs["support_tracks"].shape[:-3], *decoder_query.shape)
        query_frame = torch.zeros(decoder_query.shape[:-1], dtype=torch.int32)
    decoder_query = sinusoidal_embedding(decoder_query, cfg.num_frequencies, cfg.track_scale_factor)
    return DecoderContext(decoder_query, query_frame, inputs["boundary_frame"])


def append_time_feat(latents, query_frame):
    """append_time_feat (track_autoencoder_3d.py:235-246) as the literal one-hot einsum."""
    C = latents.shape[-1]
    idx = query_frame.to(torch.int64) * 5
    d = torch.arange(128)
    c = torch.arange(C)
    # jnp.eye(128, C, k)[d, c] = 1 iff c == d + k
    eye = (c[None, None, None, :] == (d[None, None, :, None] + idx[..., None, None])).to(latents.dtype)
    to_append = torch.einsum("...nc,...dc->...nd", latents, eye)
    return torch.cat([latents, to_append], dim=-1)


def quantize_latents(latents, noise, discretize=True):
    """decode() head (track_autoencoder_3d.py:251-260): clip, round to 1/128, add noise, STE.

    ``noise`` stands in for jax.random.uniform(PRNGKey(0), shape): the threefry bit stream
    cannot be generated here, so it is an explicit U[0,1) input.
    """
    latents = torch.clamp(latents, -1.0, 1.0)
    if discretize:
        disc = round_half_even(latents * 128.0) / 128.0
        disc = disc + noise / 128.0 - 1.0 / 256.0
        latents = latents - (latents - disc).detach()
    return latents


def decode_3d(p, cfg, latents, ctx, noise=None, discretize=True, out_coords=3):
    """decode (track_autoencoder_3d.py:248-307)."""
    latents = quantize_latents(latents, noise, discretize)
    latents = dense(latents, p["decompressor"])
    latents = improved_transformer(p["decompress_attn"], latents)
    tfeat = torch.floor(ctx.query_frame[..., None].to(latents.dtype) / cfg.time_scale_factor)
    queries = torch.cat([ctx.decoder_query, tfeat], dim=-1)
    pce = dense(sinusoidal_embedding(queries, cfg.num_frequencies, cfg.track_scale_factor), p["query_encoder"])
    Q = pce.shape[-2]
    lat = latents[:, None].expand(latents.shape[0], Q, *latents.shape[1:])
    lat = append_time_feat(lat, ctx.query_frame)
    tokens = torch.cat([pce[..., None, :], lat], dim=2)
    out = improved_transformer(p["track_readout_attn"], tokens)
    out = dense(out[..., 0, :], p["track_predictor"])
    nf = cfg.num_output_frames
    if out_coords == 3:
        tracks = torch.stack([out[..., :nf], out[..., nf : 2 * nf], out[..., 2 * nf : 3 * nf]], dim=-1)
        vis = out[..., 3 * nf :, None]
        cert = torch.zeros_like(vis)
    else:  # TRAJAN 2D (track_autoencoder.py:334-339)
        tracks = torch.stack([out[..., :nf], out[..., nf : 2 * nf]], dim=-1)
        vis = out[..., 2 * nf : 3 * nf, None]
        cert = out[..., 3 * nf :, None]
    return Results(tracks, vis, cert)


def forward_3d(p, cfg, inputs, noise=None, discretize=True):
    """TrackAutoEncoder3D.__call__ (track_autoencoder_3d.py:309-357)."""
    latents = encode_3d(p, cfg, inputs)
    if cfg.decoder_scan_chunk_size is None:
        ctx = get_decoder_context(cfg, inputs)
        return decode_3d(p, cfg, latents, ctx, noise, discretize)
    h = cfg.decoder_scan_chunk_size
    qp = inputs["query_points"]
    outs = []
    for s in range(0, qp.shape[-2], h):
        sub = dict(inputs)
        sub["query_points"] = qp[..., s : s + h, :]
        outs.append(decode_3d(p, cfg, latents, get_decoder_context(cfg, sub), noise, discretize))
    return Results(
        torch.cat([o.tracks for o in outs], dim=1),
        torch.cat([o.visible_logits for o in outs], dim=1),
        torch.cat([o.certain_logits for o in outs], dim=1),
    )


# ----------------------------------------------------------------------------------------
# track_autoencoder.py  (TRAJAN 2D, as written)
# ----------------------------------------------------------------------------------------


@dataclass
class Config2D:
    """Constructor fields of TrackAutoEncoder (track_autoencoder.py:120-135)."""

    num_output_frames: int = 150
    num_latent_tokens: int = 128
    latent_token_dim: int = 64
    num_frequencies: int = 32
    track_scale_factor: float = 1.0
    time_scale_factor: float = 150.0
    track_token_dim: int = 256
    encoder_latent_dim: int = 512
    decoder_num_channels: int = 1024
    decoder_scan_chunk_size: Optional[int] = None


def encode_tracks_2d(p, cfg, tracks, visible, restart):
    """encode_tracks (track_autoencoder.py:205-232): key mask, masked mean over time."""
    B, N, T, _ = tracks.shape
    fr_id = (torch.arange(T, dtype=torch.float32) / T).to(tracks.dtype)
    fr_id = fr_id[None, None, :, None].expand(visible.shape)
    emb = sinusoidal_embedding(torch.cat([tracks, fr_id], dim=-1), cfg.num_frequencies, cfg.track_scale_factor)
    tok = dense(emb, p["track_token_projection"])
    time = torch.arange(T)
    partition = time[None, None, :] < restart.reshape(B, 1, 1)
    vis = visible[..., 0] != 0
    km = partition & vis  # [B,N,T] keys
    qq = km[:, :, None, None, :].expand(B, N, 1, T, T)
    tok = improved_transformer(p["input_track_transformer"], tok, qq_mask=qq)
    v = vis[..., None].to(tok.dtype)
    return (tok * v).sum(dim=-2) / torch.clamp(v.sum(dim=-2), min=1.0)


def forward_2d(p, cfg, inputs, noise=None, discretize=True):
    """TrackAutoEncoder.__call__ (track_autoencoder.py:347-390), unchunked."""
    st = encode_tracks_2d(p, cfg, inputs["support_tracks"], inputs["support_tracks_visible"], inputs["boundary_frame"])
    B = st.shape[0]
    lat = p["initializer"]["state_init"].expand(B, *p["initializer"]["state_init"].shape)
    lat = improved_transformer(p["tracks_to_latents"], lat, st)
    lat = dense(lat, p["compressor"])
    ctx = get_decoder_context(cfg, inputs, coords=2)
    return decode_3d(p, cfg, lat, ctx, noise, discretize, out_coords=2)


# ----------------------------------------------------------------------------------------
# train.py: loss, LR schedule, optimiser
# ----------------------------------------------------------------------------------------


def bce_with_logits(logits, labels):
    """optax.sigmoid_binary_cross_entropy: -y*log_sigmoid(l) - (1-y)*log_sigmoid(-l)."""
    ls = torch.nn.functional.logsigmoid
    return -labels * ls(logits) - (1.0 - labels) * ls(-logits)


def compute_loss_3d(pred: Results, targets, l1_weight=5000.0, bce_weight=1e-8):
    """compute_loss_3d (train.py:96-129) (identical maths to compute_loss_2d :60-93)."""
    tt = targets["query_tracks"]
    tv = targets["query_tracks_visible"].to(pred.tracks.dtype)
    pos_err = (pred.tracks - tt).abs()
    denom = torch.clamp(tv.sum(), min=1.0)
    position_loss = (pos_err * tv).sum() / denom
    visible_loss = bce_with_logits(pred.visible_logits, tv).sum() / denom
    total = l1_weight * position_loss + bce_weight * visible_loss
    return {"total_loss": total, "position_loss": position_loss, "visible_loss": visible_loss}


def learning_rate(step, base_lr=1e-4, warmup_steps=10000, total_steps=1000000):
    """create_learning_rate_schedule (train.py:41-57): linear 0->base then cosine to 0."""
    if step < warmup_steps:
        return base_lr * step / warmup_steps
    t = min(step - warmup_steps, total_steps - warmup_steps) / (total_steps - warmup_steps)
    return base_lr * 0.5 * (1.0 + math.cos(math.pi * t))


def adamw_step(params, grads, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8, wd=0.01, clip=1.0):
    """optax.chain(clip_by_global_norm(1.0), adamw(lr, weight_decay=0.01)) (train.py:239-243).

    Flat lists of tensors, updated in place.  ``step`` is the 1-based count after this update.
    """
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    scale = torch.where(gn < clip, torch.ones_like(gn), clip / gn)  # optax: g * clip / max(gn, clip)
    for p_, g, m_, v_ in zip(params, grads, m, v):
        g = g * scale
        m_.mul_(b1).add_(g, alpha=1 - b1)
        v_.mul_(b2).addcmul_(g, g, value=1 - b2)
        mh = m_ / (1 - b1**step)
        vh = v_ / (1 - b2**step)
        p_.add_(-(lr) * (mh / (torch.sqrt(vh) + eps) + wd * p_))
    return gn


# ----------------------------------------------------------------------------------------
# Parameter trees (Flax naming / initialisers)
# ----------------------------------------------------------------------------------------


def _lecun_normal(rng, shape, fan_in):
    """jax.nn.initializers.lecun_normal: truncated normal (+-2 sigma), var 1/fan_in."""
    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def _dense_init(rng, din, dout):
    return {"kernel": _lecun_normal(rng, (din, dout), din), "bias": np.zeros(dout, np.float32)}


def _attn_init(rng, d_q, d_kv, heads, dh):
    return {
        "dense_query": {"kernel": _lecun_normal(rng, (d_q, heads, dh), d_q)},
        "dense_key": {"kernel": _lecun_normal(rng, (d_kv, heads, dh), d_kv)},
        "dense_value": {"kernel": _lecun_normal(rng, (d_kv, heads, dh), d_kv)},
        "norm_query": {"scale": np.ones(dh, np.float32)},
        "norm_key": {"scale": np.ones(dh, np.float32)},
        "dense_out": {"kernel": _lecun_normal(rng, (heads, dh, d_q), heads * dh), "bias": np.zeros(d_q, np.float32)},
    }


def _transformer_init(rng, d, qkv, heads, mlp, layers, d_kv=None):
    p = {}
    for i in range(layers):
        lp = {
            "norm_q": {"scale": np.ones(d, np.float32)},
            "norm_attn": {"scale": np.ones(d, np.float32)},
            "self_att": _attn_init(rng, d, d, heads, qkv // heads),
            "MLP_in": _dense_init(rng, d, mlp),
            "MLP_out": _dense_init(rng, mlp, d),
        }
        if d_kv is not None:
            lp["cross_att"] = _attn_init(rng, d, d_kv, heads, qkv // heads)
        p[f"layer_{i}"] = lp
    p["norm_encoder"] = {"scale": np.ones(d, np.float32)}
    return p


# (qkv_size, heads, mlp_size, layers) per transformer; reference values
# track_autoencoder_3d.py:89-112 and track_autoencoder.py:148-171.  Tests may shrink them.
ARCH_3D = {"itt": (768, 8, 1536, 3), "t2l": (768, 8, 2048, 4), "dec": (768, 8, 2048, 4), "tra": (768, 8, 1536, 4)}
ARCH_2D = {"itt": (512, 8, 1024, 2), "t2l": (512, 8, 2048, 6), "dec": (512, 8, 2048, 3), "tra": (512, 8, 1024, 4)}


def init_params_3d(cfg: Config3D, seed=0, has_dino=True, has_depth=True, randomize_norms=False, arch=None):
    """Parameter tree ``model.init(rng, batch)['params']`` would create (repair R2 widths).

    Distributions follow the Flax initialisers; the random stream is NumPy's, not JAX's.
    ``randomize_norms`` perturbs scales/biases so tests exercise them (Flax inits them to 1/0).
    """
    rng = np.random.RandomState(seed)
    a = arch or ARCH_3D
    W, E, D = cfg.track_token_dim, cfg.encoder_latent_dim, cfg.decoder_num_channels
    nf = cfg.num_frequencies
    p = {
        "initializer": {"state_init": rng.standard_normal((cfg.num_latent_tokens, E)).astype(np.float32)},
        "input_readout_token": {"state_init": rng.standard_normal((1, W)).astype(np.float32)},
        "track_token_projection": _dense_init(rng, 4 * 2 * nf, W),
        "input_track_transformer": _transformer_init(rng, W, *a["itt"]),
        "tracks_to_latents": _transformer_init(rng, E, *a["t2l"], d_kv=W),
        "compressor": _dense_init(rng, E, cfg.latent_token_dim),
        "decompressor": _dense_init(rng, cfg.latent_token_dim, D - 128),
        "decompress_attn": _transformer_init(rng, D - 128, *a["dec"]),
        "track_readout_attn": _transformer_init(rng, D, *a["tra"]),
        "query_encoder": _dense_init(rng, (3 * 2 * nf + 1) * 2 * nf, D),
        "track_predictor": _dense_init(rng, D, cfg.num_output_frames * 4),
    }
    if cfg.use_dino and has_dino:
        p["dino_projection"] = _dense_init(rng, cfg.dino_feature_dim, W)
    if cfg.use_depth and has_depth:
        p["depth_projection"] = _dense_init(rng, cfg.depth_feature_dim, W)
    if randomize_norms:
        _randomize(p, rng)
    return p


def init_params_2d(cfg: Config2D, seed=0, randomize_norms=False, arch=None):
    """TRAJAN tree (track_autoencoder.py:137-173)."""
    rng = np.random.RandomState(seed)
    a = arch or ARCH_2D
    W, E, D = cfg.track_token_dim, cfg.encoder_latent_dim, cfg.decoder_num_channels
    nf = cfg.num_frequencies
    p = {
        "initializer": {"state_init": rng.standard_normal((cfg.num_latent_tokens, E)).astype(np.float32)},
        "track_token_projection": _dense_init(rng, 3 * 2 * nf, W),
        "input_track_transformer": _transformer_init(rng, W, *a["itt"]),
        "tracks_to_latents": _transformer_init(rng, E, *a["t2l"], d_kv=W),
        "compressor": _dense_init(rng, E, cfg.latent_token_dim),
        "decompressor": _dense_init(rng, cfg.latent_token_dim, D - 128),
        "decompress_attn": _transformer_init(rng, D - 128, *a["dec"]),
        "track_readout_attn": _transformer_init(rng, D, *a["tra"]),
        "query_encoder": _dense_init(rng, (2 * 2 * nf + 1) * 2 * nf, D),
        "track_predictor": _dense_init(rng, D, cfg.num_output_frames * 4),
    }
    if randomize_norms:
        _randomize(p, rng)
    return p


def _randomize(p, rng):
    for k, v in p.items():
        if isinstance(v, dict):
            _randomize(v, rng)
        elif k == "scale":
            p[k] = (1.0 + 0.2 * rng.standard_normal(v.shape)).astype(np.float32)
        elif k == "bias":
            p[k] = (0.1 * rng.standard_normal(v.shape)).astype(np.float32)


def to_torch(tree, dtype=torch.float32, requires_grad=False):
    if isinstance(tree, dict):
        return {k: to_torch(v, dtype, requires_grad) for k, v in tree.items()}
    t = torch.as_tensor(np.asarray(tree)).to(dtype).clone()
    if requires_grad:
        t.requires_grad_(True)
    return t


def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        key = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(flatten(v, key))
        else:
            out[key] = v
    return out


def count_params(tree):
    return sum(int(np.prod(v.shape)) for v in flatten(tree).values())


def cast_inputs(inputs: Dict[str, Any], dtype):
    out = {}
    for k, v in inputs.items():
        t = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v.detach().cpu()
        out[k] = t.to(dtype) if t.is_floating_point() else t
    return out
