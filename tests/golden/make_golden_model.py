"""Golden vectors for the MODEL path, produced by executing the reference's own source files.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_model.py

JAX / Flax / optax are not installable here, so the reference's ``attention.py``, ``track_autoencoder.py``,
``track_autoencoder_3d.py`` and the loss / schedule functions of ``train.py`` are executed UNMODIFIED, where they lie,
on ``oracle/flax_shim.py`` - NumPy stand-ins for the framework primitives only (Dense, DenseGeneral, LayerNorm, RMSNorm,
gelu, dot_product_attention, Module scoping, scan, vmap; see that file's header for the published semantics followed).
All model logic in the outputs below therefore comes from the reference's lines, in float64.

Parameters are NOT stored (too large): every case records the seed / constructor arguments and both sides rebuild the tree with
``oracle.model.init_params_*`` (NumPy ``RandomState`` stream, platform independent).  The reference code looks parameters up
by ITS OWN names and shapes, and the generator asserts that it consumed every leaf of the tree (naming parity).

  model_transformer.npz   attention.py:11-185       ImprovedTransformer: self-attention + key mask; cross-attention + mask
  model_trajan.npz        track_autoencoder.py:117-390   TrackAutoEncoder.__call__ as written (quantiser noise injected),
                                                    unchunked and decoder_scan_chunk_size=2; no query_points (32x32 grid)
  model_3dspa.npz         track_autoencoder_3d.py   (a) decode + get_decoder_context as written, default widths;
                                                    (b) embed_track_pos_visible as written with R2' widths (768) and
                                                        without features at default widths;
                                                    (c) encode_tracks AS WRITTEN raises (F3a) - the message is recorded;
                                                    (d) the whole forward under R2' with the R1 key mask: the embedding, the
                                                        three transformers, compressor, context and decode are the
                                                        reference's bound methods; only the six glue lines of
                                                        encode_tracks/encode (readout concat, mask, token 0) are re-composed here
  model_loss.npz          train.py:41-129           compute_loss_3d, compute_loss_2d, create_learning_rate_schedule
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import flax_shim as fs  # noqa: E402
from oracle import model as om  # noqa: E402

F64 = np.float64


def f32(a):
    return np.asarray(a, np.float32)


def to64(tree):
    return {k: to64(v) for k, v in tree.items()} if isinstance(tree, dict) else np.asarray(tree, F64)


def leaves(tree, prefix=""):
    out = []
    for k, v in tree.items():
        out += leaves(v, f"{prefix}{k}/") if isinstance(v, dict) else [f"{prefix}{k}"]
    return out


def assert_consumed(tree, allow_unused=()):
    unused = [p for p in leaves(tree) if p not in fs.accessed and not p.startswith(tuple(allow_unused))]
    assert not unused, f"reference code never read: {unused[:5]}"
    fs.accessed.clear()


def transformer_cases(ref):
    rs = np.random.RandomState(11)
    out = {}
    d, qkv, heads, mlp, layers = 24, 32, 4, 48, 2
    # (i) self-attention over [B, N, L, d] with a key mask given per (b, n)
    p = om._transformer_init(np.random.RandomState(5), d, qkv, heads, mlp, layers)
    om._randomize(p, np.random.RandomState(6))
    x = rs.standard_normal((2, 3, 7, d))
    keys_on = rs.uniform(size=(2, 3, 1, 7)) < 0.7
    keys_on[..., 0] = True
    keys_on[1, 2, 0, 1:] = False   # a sequence that sees one key only
    mask = np.broadcast_to(keys_on, (2, 3, 7, 7))
    tr = fs.bind(ref["attention"].ImprovedTransformer(qkv_size=qkv, num_heads=heads, mlp_size=mlp, num_layers=layers), to64(p))
    out.update({"self/x": x, "self/mask": mask, "self/y": tr(x, qq_mask=mask), "self/y_nomask": tr(x),
                "self/arch": np.array([d, qkv, heads, mlp, layers]), "self/seeds": np.array([5, 6])})
    assert_consumed(p)
    # (ii) parallel self + cross attention, keys of another width, a cross mask with a fully masked query row
    dkv = 20
    p = om._transformer_init(np.random.RandomState(7), d, qkv, heads, mlp, layers, d_kv=dkv)
    om._randomize(p, np.random.RandomState(8))
    q = rs.standard_normal((2, 5, d))
    kv = rs.standard_normal((2, 9, dkv))
    qk = rs.uniform(size=(2, 5, 9)) < 0.6
    qk[0, 1, :] = False            # fully masked row -> uniform weights (finfo.min fill, not -inf)
    tr = fs.bind(ref["attention"].ImprovedTransformer(qkv_size=qkv, num_heads=heads, mlp_size=mlp, num_layers=layers), to64(p))
    out.update({"cross/q": q, "cross/kv": kv, "cross/mask": qk, "cross/y": tr(q, kv, qk_mask=qk), "cross/y_nomask": tr(q, kv),
                "cross/arch": np.array([d, qkv, heads, mlp, layers, dkv]), "cross/seeds": np.array([7, 8])})
    assert_consumed(p)
    return out


def trajan_cases(ref):
    rs = np.random.RandomState(21)
    B, N, T, Q, F = 2, 5, 6, 4, 6
    cfg = om.Config2D(num_output_frames=F)
    p = om.init_params_2d(cfg, seed=3, randomize_norms=True)
    inputs = {
        "support_tracks": f32(rs.uniform(0, 1, (B, N, T, 2))),
        "support_tracks_visible": (rs.uniform(size=(B, N, T, 1)) < 0.75).astype(F64),
        "query_points": f32(np.concatenate([rs.randint(0, T, (B, Q, 1)).astype(F64), rs.uniform(0, 1, (B, Q, 2))], -1)),
        "boundary_frame": np.array([T, T - 2], np.int32),
    }
    noise = rs.uniform(size=(B, cfg.num_latent_tokens, cfg.latent_token_dim))
    fs.set_uniform(lambda shape: noise.reshape(shape))
    TA = ref["track_autoencoder"].TrackAutoEncoder
    m = fs.bind(TA(num_output_frames=F), to64(p))
    res = m(inputs)
    assert_consumed(p)
    out = {f"in/{k}": v for k, v in inputs.items()}
    out.update({"noise": noise, "seed": np.array(3), "num_output_frames": np.array(F), "latents": m.encode(inputs),
                "tracks": res.tracks, "visible_logits": res.visible_logits, "certain_logits": res.certain_logits})
    rn = m.decode(latents=m.encode(inputs), decoder_context=m.get_decoder_context(inputs), discretize=False)
    out.update({"nodisc/tracks": rn.tracks, "nodisc/visible_logits": rn.visible_logits, "nodisc/certain_logits": rn.certain_logits})
    mc = fs.bind(TA(num_output_frames=F, decoder_scan_chunk_size=2), to64(p))
    rc = mc(inputs)
    out.update({"chunked/tracks": rc.tracks, "chunked/visible_logits": rc.visible_logits, "chunked/certain_logits": rc.certain_logits})
    # default 32x32 query grid (no query_points): only the first 8 queries are stored
    grid_in = {k: v[:1] for k, v in inputs.items() if k != "query_points"}
    fs.set_uniform(lambda shape: noise[:1].reshape(shape))
    rg = m(grid_in)
    out.update({"grid/tracks8": rg.tracks[:, :8], "grid/visible_logits8": rg.visible_logits[:, :8], "grid/n_queries": np.array(rg.tracks.shape[1])})
    fs.accessed.clear()
    return out


def spa3d_cases(ref):
    rs = np.random.RandomState(31)
    B, N, T, Q, F = 2, 4, 5, 3, 5
    M3 = ref["track_autoencoder_3d"].TrackAutoEncoder3D
    out = {}

    def make_inputs(dino_dim, depth_dim):
        return {
            "support_tracks": f32(rs.uniform(-1, 1, (B, N, T, 3))),
            "support_tracks_visible": (rs.uniform(size=(B, N, T, 1)) < 0.75).astype(F64),
            "query_points": f32(np.concatenate([rs.randint(0, T, (B, Q, 1)).astype(F64), rs.uniform(-1, 1, (B, Q, 3))], -1)),
            "boundary_frame": np.array([T, T - 1], np.int32),
            "dino_features": f32(rs.standard_normal((B, N, T, dino_dim))).astype(F64),
            "depth_features": f32(rs.standard_normal((B, N, T, depth_dim))).astype(F64),
        }

    # (a) decoder as written, default widths
    cfg = om.Config3D(num_output_frames=F)
    p = om.init_params_3d(cfg, seed=4, randomize_norms=True)
    inputs = make_inputs(768, 256)
    latents = f32(rs.uniform(-1.3, 1.3, (B, cfg.num_latent_tokens, cfg.latent_token_dim))).astype(F64)   # float32-valued: x*128 rounds identically in fp32
    noise = rs.uniform(size=latents.shape)
    fs.set_uniform(lambda shape: noise.reshape(shape))
    m = fs.bind(M3(num_output_frames=F), to64(p))
    ctx = m.get_decoder_context(inputs)
    res = m.decode(latents=latents, decoder_context=ctx)
    res_nd = m.decode(latents=latents, decoder_context=ctx, discretize=False)
    out.update({f"dec/in/{k}": v for k, v in inputs.items()})
    out.update({"dec/latents": latents, "dec/noise": noise, "dec/seed": np.array(4), "dec/num_output_frames": np.array(F),
                "dec/ctx_query": ctx.decoder_query, "dec/ctx_frame": ctx.query_frame,
                "dec/tracks": res.tracks, "dec/visible_logits": res.visible_logits, "dec/certain_logits": res.certain_logits,
                "dec/nodisc/tracks": res_nd.tracks, "dec/nodisc/visible_logits": res_nd.visible_logits})
    # (b1) embedding without features, default widths
    out["emb/plain"] = m.embed_track_pos_visible(inputs["support_tracks"], inputs["support_tracks_visible"])
    # (c) encode_tracks as written cannot run: record what happens
    try:
        m.encode_tracks(inputs["support_tracks"], inputs["support_tracks_visible"], inputs["boundary_frame"])
        raise AssertionError("encode_tracks as written was expected to fail (F3a)")
    except ValueError as e:
        out["as_written_error"] = np.array(str(e))
    # ... and with features at the default widths the projections cannot be added to the tokens (F3b)
    p_aw = dict(p)
    p_aw["dino_projection"] = om._dense_init(np.random.RandomState(1), 768, 768)     # the widths the as-written setup() creates
    p_aw["depth_projection"] = om._dense_init(np.random.RandomState(2), 256, 256)
    m_aw = fs.bind(M3(num_output_frames=F), to64(p_aw))
    try:
        m_aw.embed_track_pos_visible(inputs["support_tracks"], inputs["support_tracks_visible"], inputs["dino_features"], inputs["depth_features"])
        raise AssertionError("default-width feature projections were expected to fail (F3b)")
    except ValueError as e:
        out["as_written_error_widths"] = np.array(str(e))
    fs.accessed.clear()

    # (b2, d) R2': the as-written constructor with consistent widths; R1 key mask
    cfg2 = om.Config3D(num_output_frames=F, track_token_dim=768, depth_feature_dim=768)
    p2 = om.init_params_3d(cfg2, seed=9, randomize_norms=True)
    in2 = make_inputs(768, 768)
    noise2 = rs.uniform(size=(B, cfg2.num_latent_tokens, cfg2.latent_token_dim))
    fs.set_uniform(lambda shape: noise2.reshape(shape))
    m2 = fs.bind(M3(num_output_frames=F, track_token_dim=768, depth_feature_dim=768), to64(p2))
    emb = m2.embed_track_pos_visible(in2["support_tracks"], in2["support_tracks_visible"], in2["dino_features"], in2["depth_features"])
    tokens = np.concatenate([m2.input_readout_token(emb.shape[:-2]), emb], axis=-2)
    on = (in2["support_tracks_visible"][..., 0] != 0) & (np.arange(T)[None, None, :] < in2["boundary_frame"][:, None, None])
    key_mask = np.concatenate([np.ones((B, N, 1), bool), on], -1)              # R1: readout key always on
    enc = m2.input_track_transformer(tokens, qq_mask=key_mask[:, :, None, :])  # [B, N, 1, T+1] -> broadcast over queries
    support = enc[..., 0, :]
    lat2 = m2.compressor(m2.tracks_to_latents(m2.initializer(batch_shape=(B,)), support))
    ctx2 = m2.get_decoder_context(in2)
    res2 = m2.decode(latents=lat2, decoder_context=ctx2)
    res2_nd = m2.decode(latents=lat2, decoder_context=ctx2, discretize=False)
    assert_consumed(p2)
    out.update({f"full/in/{k}": v for k, v in in2.items()})
    out.update({"full/noise": noise2, "full/seed": np.array(9), "full/emb": emb, "full/support_tokens": support, "full/latents": lat2,
                "full/tracks": res2.tracks, "full/visible_logits": res2.visible_logits, "full/certain_logits": res2.certain_logits,
                "full/nodisc/tracks": res2_nd.tracks, "full/nodisc/visible_logits": res2_nd.visible_logits})
    return out


def loss_cases():
    rs = np.random.RandomState(41)
    env = fs.load_loss()
    B, Q, T = 2, 5, 7

    class Pred:
        pass

    out = {}
    for tag, C in (("3d", 3), ("2d", 2)):
        pr = Pred()
        pr.tracks = rs.standard_normal((B, Q, T, C))
        pr.visible_logits = 4.0 * rs.standard_normal((B, Q, T, 1))
        pr.certain_logits = rs.standard_normal((B, Q, T, 1))
        tg = {"query_tracks": rs.standard_normal((B, Q, T, C)), "query_tracks_visible": (rs.uniform(size=(B, Q, T, 1)) < 0.7).astype(F64)}
        r = env[f"compute_loss_{tag}"](pr, tg)
        out.update({f"{tag}/pred_tracks": pr.tracks, f"{tag}/pred_visible_logits": pr.visible_logits, f"{tag}/tracks": tg["query_tracks"],
                    f"{tag}/visible": tg["query_tracks_visible"]})
        out.update({f"{tag}/{k}": np.asarray(v) for k, v in r.items()})
    # nothing visible: denominators clamp at 1
    pr = Pred()
    pr.tracks, pr.visible_logits = rs.standard_normal((1, 2, 3, 3)), rs.standard_normal((1, 2, 3, 1))
    r = env["compute_loss_3d"](pr, {"query_tracks": rs.standard_normal((1, 2, 3, 3)), "query_tracks_visible": np.zeros((1, 2, 3, 1))})
    out.update({"none/pred_tracks": pr.tracks, "none/pred_visible_logits": pr.visible_logits, "none/total_loss": np.asarray(r["total_loss"]),
                "none/visible_loss": np.asarray(r["visible_loss"])})
    steps = np.array([0, 1, 5000, 9999, 10000, 10001, 250000, 505000, 999999, 1000000, 1200000])
    sched = env["create_learning_rate_schedule"](1e-4, 10000, 1000000)
    out.update({"lr/steps": steps, "lr/values": np.array([float(sched(s)) for s in steps])})
    return out


def main():
    ref = fs.load_reference()
    for name, data in (("model_transformer", transformer_cases(ref)), ("model_trajan", trajan_cases(ref)),
                       ("model_3dspa", spa3d_cases(ref)), ("model_loss", loss_cases())):
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {len(data)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
