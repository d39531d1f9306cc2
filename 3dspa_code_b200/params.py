"""Parameter trees: Flax naming on the outside, kernel-friendly packing on the inside.

Outside (what ``init`` returns, what checkpoints hold): the nested ``params`` tree
``TrackAutoEncoder3D.init(rng, batch)['params']`` creates in the reference
(track_autoencoder_3d.py:69-115, attention.py:41-183), numpy float32.

Inside (what the kernels read): every Dense kernel transposed to [out, in] (K contiguous - the
operand layout of the tcgen05 GEMM), q/k/v kernels of a self-attention concatenated into one
[3*H*Dh, d] matrix, the three embedding projections concatenated along K.  ``pack`` / ``unpack``
are exact inverses, so the optimiser can work on the packed master copy and checkpoints can be
written back in the Flax layout.

Checkpoint IO follows inference.py:450-508: the three accepted ``.npz`` layouts
(``params`` pickled dict, ``optimizer -> target``, flat ``a/b/c`` keys) are read; the flat layout
is written.  Widths are inferred from the stored shapes, so both repair R2 (projections to
track_token_dim) and R2' (track_token_dim = 768) load.
"""
from __future__ import annotations

import math
import warnings
from typing import Dict

import numpy as np

ARCH_3D = {"itt": (768, 8, 1536, 3), "t2l": (768, 8, 2048, 4), "dec": (768, 8, 2048, 4), "tra": (768, 8, 1536, 4)}
ARCH_2D = {"itt": (512, 8, 1024, 2), "t2l": (512, 8, 2048, 6), "dec": (512, 8, 2048, 3), "tra": (512, 8, 1024, 4)}
TRANSFORMERS = {
    "itt": "input_track_transformer",
    "t2l": "tracks_to_latents",
    "dec": "decompress_attn",
    "tra": "track_readout_attn",
}


# ---- Flax initialisers ------------------------------------------------------------------------
def _lecun_normal(rng, shape, fan_in):
    """jax.nn.initializers.lecun_normal: normal truncated at +-2 sigma, variance 1/fan_in."""
    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def _dense(rng, din, dout):
    return {"kernel": _lecun_normal(rng, (din, dout), din), "bias": np.zeros(dout, np.float32)}


def _attn(rng, d_q, d_kv, heads, dh):
    return {
        "dense_query": {"kernel": _lecun_normal(rng, (d_q, heads, dh), d_q)},
        "dense_key": {"kernel": _lecun_normal(rng, (d_kv, heads, dh), d_kv)},
        "dense_value": {"kernel": _lecun_normal(rng, (d_kv, heads, dh), d_kv)},
        "norm_query": {"scale": np.ones(dh, np.float32)},
        "norm_key": {"scale": np.ones(dh, np.float32)},
        "dense_out": {"kernel": _lecun_normal(rng, (heads, dh, d_q), heads * dh), "bias": np.zeros(d_q, np.float32)},
    }


def _transformer(rng, d, qkv, heads, mlp, layers, d_kv=None):
    p = {}
    for i in range(layers):
        lp = {
            "norm_q": {"scale": np.ones(d, np.float32)},
            "norm_attn": {"scale": np.ones(d, np.float32)},
            "self_att": _attn(rng, d, d, heads, qkv // heads),
            "MLP_in": _dense(rng, d, mlp),
            "MLP_out": _dense(rng, mlp, d),
        }
        if d_kv is not None:
            lp["cross_att"] = _attn(rng, d, d_kv, heads, qkv // heads)
        p[f"layer_{i}"] = lp
    p["norm_encoder"] = {"scale": np.ones(d, np.float32)}
    return p


def init_tree(cfg, seed=0, has_dino=True, has_depth=True, arch=None, coords=3):
    """The tree ``model.init`` creates.  ``dino_projection`` / ``depth_projection`` exist only if
    the init batch carries the feature (Flax creates params lazily)."""
    for name, (qkv, heads, _, _) in (arch or (ARCH_3D if coords == 3 else ARCH_2D)).items():
        if qkv % heads:
            raise ValueError(f"num_heads={heads} must divide qk_size={qkv}.")  # attention.py:147-148
    a = arch or (ARCH_3D if coords == 3 else ARCH_2D)
    rng = np.random.RandomState(seed)
    W, E, D = cfg.track_token_dim, cfg.encoder_latent_dim, cfg.decoder_num_channels
    nf = cfg.num_frequencies
    p = {"initializer": {"state_init": rng.standard_normal((cfg.num_latent_tokens, E)).astype(np.float32)}}
    if coords == 3:
        p["input_readout_token"] = {"state_init": rng.standard_normal((1, W)).astype(np.float32)}
    p["track_token_projection"] = _dense(rng, (coords + 1) * 2 * nf, W)
    p["input_track_transformer"] = _transformer(rng, W, *a["itt"])
    p["tracks_to_latents"] = _transformer(rng, E, *a["t2l"], d_kv=W)
    p["compressor"] = _dense(rng, E, cfg.latent_token_dim)
    p["decompressor"] = _dense(rng, cfg.latent_token_dim, D - 128)
    p["decompress_attn"] = _transformer(rng, D - 128, *a["dec"])
    p["track_readout_attn"] = _transformer(rng, D, *a["tra"])
    p["query_encoder"] = _dense(rng, (coords * 2 * nf + 1) * 2 * nf, D)
    p["track_predictor"] = _dense(rng, D, cfg.num_output_frames * 4)
    if coords == 3 and getattr(cfg, "use_dino", False) and has_dino:
        p["dino_projection"] = _dense(rng, cfg.dino_feature_dim, W)
    if coords == 3 and getattr(cfg, "use_depth", False) and has_depth:
        p["depth_projection"] = _dense(rng, cfg.depth_feature_dim, W)
    return p


# ---- tree utilities ---------------------------------------------------------------------------
def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        key = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(flatten(v, key))
        else:
            out[key] = v
    return out


def unflatten(flat):
    """inference.py:450-461 (_unflatten_params)."""
    out = {}
    for key, value in flat.items():
        parts = key.split("/")
        d = out
        for part in parts[:-1]:
            d = d.setdefault(part, {})
        d[parts[-1]] = value
    return out


def count(tree):
    return sum(int(np.prod(np.shape(v))) for v in flatten(tree).values())


def _np(v):
    return np.asarray(v, dtype=np.float32)


# ---- packing ------------------------------------------------------------------------------------
def _pack_attn(p, prefix, out, is_cross):
    wq = _np(p["dense_query"]["kernel"])
    wk = _np(p["dense_key"]["kernel"])
    wv = _np(p["dense_value"]["kernel"])
    d_q, H, Dh = wq.shape
    A = H * Dh
    if is_cross:
        out[prefix + "Wq_t"] = np.ascontiguousarray(wq.reshape(d_q, A).T)
        out[prefix + "Wkv_t"] = np.ascontiguousarray(np.concatenate([wk.reshape(-1, A), wv.reshape(-1, A)], axis=1).T)
    else:
        out[prefix + "Wqkv_t"] = np.ascontiguousarray(
            np.concatenate([wq.reshape(d_q, A), wk.reshape(d_q, A), wv.reshape(d_q, A)], axis=1).T
        )
    out[prefix + "norm_query"] = _np(p["norm_query"]["scale"])
    out[prefix + "norm_key"] = _np(p["norm_key"]["scale"])
    wo = _np(p["dense_out"]["kernel"])  # [H, Dh, d]
    out[prefix + "Wo_t"] = np.ascontiguousarray(wo.reshape(A, -1).T)
    out[prefix + "bo"] = _np(p["dense_out"]["bias"])


def _unpack_attn(pk, prefix, heads, is_cross):
    Dh = pk[prefix + "norm_query"].shape[0]
    A = heads * Dh
    p = {}
    if is_cross:
        wq = pk[prefix + "Wq_t"].T  # [d, A]
        wkv = pk[prefix + "Wkv_t"].T  # [dkv, 2A]
        wk, wv = wkv[:, :A], wkv[:, A:]
    else:
        w = pk[prefix + "Wqkv_t"].T  # [d, 3A]
        wq, wk, wv = w[:, :A], w[:, A : 2 * A], w[:, 2 * A :]
    p["dense_query"] = {"kernel": np.ascontiguousarray(wq).reshape(-1, heads, Dh)}
    p["dense_key"] = {"kernel": np.ascontiguousarray(wk).reshape(-1, heads, Dh)}
    p["dense_value"] = {"kernel": np.ascontiguousarray(wv).reshape(-1, heads, Dh)}
    p["norm_query"] = {"scale": pk[prefix + "norm_query"]}
    p["norm_key"] = {"scale": pk[prefix + "norm_key"]}
    wo = pk[prefix + "Wo_t"].T  # [A, d]
    p["dense_out"] = {"kernel": np.ascontiguousarray(wo).reshape(heads, Dh, -1), "bias": pk[prefix + "bo"]}
    return p


def tree_meta(tree):
    """Architecture facts inferred from the tree shapes (widths, heads, layers, features)."""
    meta = {"coords": 3 if "input_readout_token" in tree else 2}
    meta["has_dino"] = "dino_projection" in tree
    meta["has_depth"] = "depth_projection" in tree
    for short, name in TRANSFORMERS.items():
        t = tree[name]
        layers = sum(1 for k in t if k.startswith("layer_"))
        wq = np.shape(t["layer_0"]["self_att"]["dense_query"]["kernel"])
        meta[short] = {
            "layers": layers,
            "d": wq[0],
            "heads": wq[1],
            "Dh": wq[2],
            "mlp": np.shape(t["layer_0"]["MLP_in"]["kernel"])[1],
            "cross": "cross_att" in t["layer_0"],
        }
    meta["W"] = np.shape(tree["track_token_projection"]["kernel"])[1]
    meta["fourier_in"] = np.shape(tree["track_token_projection"]["kernel"])[0]
    meta["dino_dim"] = np.shape(tree["dino_projection"]["kernel"])[0] if meta["has_dino"] else 0
    meta["depth_dim"] = np.shape(tree["depth_projection"]["kernel"])[0] if meta["has_depth"] else 0
    for nm, key in (("dino", "dino_projection"), ("depth", "depth_projection")):
        if meta[f"has_{nm}"] and np.shape(tree[key]["kernel"])[1] != meta["W"]:
            raise ValueError(
                f"{key} projects to {np.shape(tree[key]['kernel'])[1]} channels but track tokens are {meta['W']} wide: "
                "the features are ADDED to the track tokens (track_autoencoder_3d.py:142,147), see repair R2 in DESIGN.md"
            )
    meta["latent_tokens"], meta["E"] = np.shape(tree["initializer"]["state_init"])
    meta["latent_dim"] = np.shape(tree["compressor"]["kernel"])[1]
    meta["D"] = np.shape(tree["query_encoder"]["kernel"])[1]
    meta["query_in"] = np.shape(tree["query_encoder"]["kernel"])[0]
    meta["head_out"] = np.shape(tree["track_predictor"]["kernel"])[1]
    return meta


def pack(tree) -> Dict[str, np.ndarray]:
    """Flax tree -> flat dict of kernel-layout float32 arrays."""
    meta = tree_meta(tree)
    out = {}
    ws = [_np(tree["track_token_projection"]["kernel"])]
    out["embed.b_track"] = _np(tree["track_token_projection"]["bias"])
    if meta["has_dino"]:
        ws.append(_np(tree["dino_projection"]["kernel"]))
        out["embed.b_dino"] = _np(tree["dino_projection"]["bias"])
    if meta["has_depth"]:
        ws.append(_np(tree["depth_projection"]["kernel"]))
        out["embed.b_depth"] = _np(tree["depth_projection"]["bias"])
    out["embed.Wt"] = np.ascontiguousarray(np.concatenate(ws, axis=0).T)  # [W, 256 (+768) (+256)]
    out["latents_init"] = _np(tree["initializer"]["state_init"])
    if meta["coords"] == 3:
        out["readout_token"] = _np(tree["input_readout_token"]["state_init"])
    for short, name in TRANSFORMERS.items():
        t = tree[name]
        for i in range(meta[short]["layers"]):
            lp = t[f"layer_{i}"]
            pre = f"{short}.{i}."
            out[pre + "norm_q"] = _np(lp["norm_q"]["scale"])
            out[pre + "norm_attn"] = _np(lp["norm_attn"]["scale"])
            _pack_attn(lp["self_att"], pre + "self.", out, False)
            if "cross_att" in lp:
                _pack_attn(lp["cross_att"], pre + "cross.", out, True)
            out[pre + "W1_t"] = np.ascontiguousarray(_np(lp["MLP_in"]["kernel"]).T)
            out[pre + "b1"] = _np(lp["MLP_in"]["bias"])
            out[pre + "W2_t"] = np.ascontiguousarray(_np(lp["MLP_out"]["kernel"]).T)
            out[pre + "b2"] = _np(lp["MLP_out"]["bias"])
        out[f"{short}.norm_encoder"] = _np(t["norm_encoder"]["scale"])
    for key in ("compressor", "decompressor", "query_encoder", "track_predictor"):
        out[f"{key}.Wt"] = np.ascontiguousarray(_np(tree[key]["kernel"]).T)
        out[f"{key}.b"] = _np(tree[key]["bias"])
    return out


def unpack(pk, meta):
    """Inverse of ``pack`` (exact)."""
    pk = {k: np.asarray(v, dtype=np.float32) for k, v in pk.items()}
    tree = {}
    wcat = pk["embed.Wt"].T  # [K, W]
    k0 = meta["fourier_in"]
    tree["track_token_projection"] = {"kernel": np.ascontiguousarray(wcat[:k0]), "bias": pk["embed.b_track"]}
    if meta["has_dino"]:
        tree["dino_projection"] = {"kernel": np.ascontiguousarray(wcat[k0 : k0 + meta["dino_dim"]]), "bias": pk["embed.b_dino"]}
        k0 += meta["dino_dim"]
    if meta["has_depth"]:
        tree["depth_projection"] = {"kernel": np.ascontiguousarray(wcat[k0 : k0 + meta["depth_dim"]]), "bias": pk["embed.b_depth"]}
    tree["initializer"] = {"state_init": pk["latents_init"]}
    if meta["coords"] == 3:
        tree["input_readout_token"] = {"state_init": pk["readout_token"]}
    for short, name in TRANSFORMERS.items():
        m = meta[short]
        t = {}
        for i in range(m["layers"]):
            pre = f"{short}.{i}."
            lp = {
                "norm_q": {"scale": pk[pre + "norm_q"]},
                "norm_attn": {"scale": pk[pre + "norm_attn"]},
                "self_att": _unpack_attn(pk, pre + "self.", m["heads"], False),
                "MLP_in": {"kernel": np.ascontiguousarray(pk[pre + "W1_t"].T), "bias": pk[pre + "b1"]},
                "MLP_out": {"kernel": np.ascontiguousarray(pk[pre + "W2_t"].T), "bias": pk[pre + "b2"]},
            }
            if m["cross"]:
                lp["cross_att"] = _unpack_attn(pk, pre + "cross.", m["heads"], True)
            t[f"layer_{i}"] = lp
        t["norm_encoder"] = {"scale": pk[f"{short}.norm_encoder"]}
        tree[name] = t
    for key in ("compressor", "decompressor", "query_encoder", "track_predictor"):
        tree[key] = {"kernel": np.ascontiguousarray(pk[f"{key}.Wt"].T), "bias": pk[f"{key}.b"]}
    return tree


# ---- checkpoints (inference.py:464-508) ----------------------------------------------------------
def _unbox(entry):
    """A Python object stored through ``np.savez`` comes back as a 0-d object array (a pickled dict) or as a mapping."""
    return entry.item() if getattr(entry, "ndim", None) == 0 else dict(entry)


def _from_optimizer(state):
    # a saved optimiser state carries the parameters under "target" (flax.optim); a bare tree is taken as it is
    return state.get("target", state) if isinstance(state, dict) else state


# The three ``.npz`` layouts the reference's loader accepts (inference.py:464-508), tried in this order: the archive member that
# identifies the layout -> how the parameter tree is read from the archive.  Anything else is the flat ``a/b/c`` key layout.
_NPZ_LAYOUTS = (
    ("params", lambda z: _unbox(z["params"])),
    ("optimizer", lambda z: _from_optimizer(_unbox(z["optimizer"]))),
)


def load_checkpoint(path):
    """Read a ``.npz`` checkpoint in any of the three layouts the reference accepts."""
    if not str(path).endswith(".npz"):
        raise ValueError("only .npz checkpoints are supported (the Flax msgpack branch needs flax)")
    with np.load(path, allow_pickle=True) as z:
        for member, read in _NPZ_LAYOUTS:
            if member in z.files:
                return read(z)
        return unflatten({k: np.array(z[k]) for k in z.files})


def save_checkpoint(path, tree):
    """Write the flat ``a/b/c`` layout (no pickle)."""
    np.savez(path, **{k: np.asarray(v) for k, v in flatten(tree).items()})


def check_structure(expected, actual, path=""):
    """Warn-only structure/shape comparison, as inference.py:608-619."""
    problems = []
    if isinstance(expected, dict) and isinstance(actual, dict):
        for k in expected:
            if k not in actual:
                problems.append(f"Key {path}.{k} missing in checkpoint")
            else:
                problems += check_structure(expected[k], actual[k], f"{path}.{k}")
    elif hasattr(expected, "shape") and hasattr(actual, "shape") and tuple(expected.shape) != tuple(actual.shape):
        problems.append(f"Shape mismatch at {path}: {tuple(expected.shape)} vs {tuple(actual.shape)}")
    if path == "":
        for p in problems:
            warnings.warn(p)
    return problems
