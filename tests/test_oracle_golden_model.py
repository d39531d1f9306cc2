"""Pin oracle/model.py against golden vectors produced by EXECUTING THE REFERENCE'S OWN MODEL FILES
(attention.py, track_autoencoder.py, track_autoencoder_3d.py, train.py:41-129) on NumPy stand-ins for the
absent jax / flax / optax primitives (oracle/flax_shim.py; generator tests/golden/make_golden_model.py).

Parameters are rebuilt from the recorded seeds on both sides; everything is float64, so agreement is to rounding."""
import os

import numpy as np
import pytest
import torch

from oracle import model as om

TOL = dict(rtol=1e-9, atol=1e-9)
F64 = torch.float64


def t(a, dtype=F64):
    a = np.asarray(a)
    if a.dtype == bool or np.issubdtype(a.dtype, np.integer):
        return torch.from_numpy(a.copy())
    return torch.from_numpy(a.copy()).to(dtype)


def close(actual, expected, **kw):
    tol = dict(TOL)
    tol.update(kw)
    np.testing.assert_allclose(actual.detach().numpy() if isinstance(actual, torch.Tensor) else actual, expected, **tol)


@pytest.fixture(scope="module")
def g_tr(golden_dir):
    return np.load(os.path.join(golden_dir, "model_transformer.npz"))


def test_transformer_self_attention_with_key_mask(g_tr):
    d, qkv, heads, mlp, layers = (int(v) for v in g_tr["self/arch"])
    s0, s1 = (int(v) for v in g_tr["self/seeds"])
    p = om._transformer_init(np.random.RandomState(s0), d, qkv, heads, mlp, layers)
    om._randomize(p, np.random.RandomState(s1))
    p = om.to_torch(p, F64)
    x = t(g_tr["self/x"])
    close(om.improved_transformer(p, x, qq_mask=t(g_tr["self/mask"])), g_tr["self/y"])
    close(om.improved_transformer(p, x), g_tr["self/y_nomask"])
    assert np.abs(g_tr["self/y"] - g_tr["self/y_nomask"]).max() > 1e-3   # the mask matters in this case


def test_transformer_parallel_cross_attention(g_tr):
    d, qkv, heads, mlp, layers, dkv = (int(v) for v in g_tr["cross/arch"])
    s0, s1 = (int(v) for v in g_tr["cross/seeds"])
    p = om._transformer_init(np.random.RandomState(s0), d, qkv, heads, mlp, layers, d_kv=dkv)
    om._randomize(p, np.random.RandomState(s1))
    p = om.to_torch(p, F64)
    q, kv = t(g_tr["cross/q"]), t(g_tr["cross/kv"])
    close(om.improved_transformer(p, q, kv, qk_mask=t(g_tr["cross/mask"])), g_tr["cross/y"])
    close(om.improved_transformer(p, q, kv), g_tr["cross/y_nomask"])


@pytest.fixture(scope="module")
def trajan(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_trajan.npz"))
    cfg = om.Config2D(num_output_frames=int(g["num_output_frames"]))
    p = om.to_torch(om.init_params_2d(cfg, seed=int(g["seed"]), randomize_norms=True), F64)
    inputs = {k[3:]: t(g[k]) for k in g.files if k.startswith("in/")}
    return g, cfg, p, inputs


def test_trajan_forward_as_written(trajan):
    g, cfg, p, inputs = trajan
    with torch.no_grad():
        res = om.forward_2d(p, cfg, inputs, noise=t(g["noise"]))
    close(res.tracks, g["tracks"])
    close(res.visible_logits, g["visible_logits"])
    close(res.certain_logits, g["certain_logits"])
    # the reference's scan-chunked decode gives the same numbers as its unchunked decode
    close(g["chunked/tracks"], g["tracks"])
    close(g["chunked/visible_logits"], g["visible_logits"])


def test_trajan_default_query_grid(trajan):
    g, cfg, p, inputs = trajan
    sub = {k: v[:1] for k, v in inputs.items() if k != "query_points"}
    with torch.no_grad():
        res = om.forward_2d(p, cfg, sub, noise=t(g["noise"][:1]))
    assert res.tracks.shape[1] == int(g["grid/n_queries"]) == 1024
    close(res.tracks[:, :8], g["grid/tracks8"])
    close(res.visible_logits[:, :8], g["grid/visible_logits8"])


@pytest.fixture(scope="module")
def g3(golden_dir):
    return np.load(os.path.join(golden_dir, "model_3dspa.npz"))


def test_3dspa_decoder_as_written(g3):
    cfg = om.Config3D(num_output_frames=int(g3["dec/num_output_frames"]))
    p = om.to_torch(om.init_params_3d(cfg, seed=int(g3["dec/seed"]), randomize_norms=True), F64)
    inputs = {k[7:]: t(g3[k]) for k in g3.files if k.startswith("dec/in/")}
    with torch.no_grad():
        ctx = om.get_decoder_context(cfg, inputs)
        close(ctx.decoder_query, g3["dec/ctx_query"])
        np.testing.assert_array_equal(ctx.query_frame.numpy(), g3["dec/ctx_frame"])
        res = om.decode_3d(p, cfg, t(g3["dec/latents"]), ctx, noise=t(g3["dec/noise"]))
        res_nd = om.decode_3d(p, cfg, t(g3["dec/latents"]), ctx, discretize=False)
        emb = om.embed_track_pos_visible_3d(p, cfg, inputs["support_tracks"], inputs["support_tracks_visible"])
    close(res.tracks, g3["dec/tracks"])
    close(res.visible_logits, g3["dec/visible_logits"])
    assert not g3["dec/certain_logits"].any() and not res.certain_logits.any()
    close(res_nd.tracks, g3["dec/nodisc/tracks"])
    close(res_nd.visible_logits, g3["dec/nodisc/visible_logits"])
    close(emb, g3["emb/plain"])


def test_3dspa_as_written_defects_are_real(g3):
    """the reference's own lines fail exactly where repairs R1 / R2 apply."""
    assert "dimension" in str(g3["as_written_error"]) or "shape" in str(g3["as_written_error"])
    assert "broadcast" in str(g3["as_written_error_widths"])


def test_3dspa_full_forward_r2prime_r1(g3):
    cfg = om.Config3D(num_output_frames=int(g3["dec/num_output_frames"]), track_token_dim=768, depth_feature_dim=768)
    p = om.to_torch(om.init_params_3d(cfg, seed=int(g3["full/seed"]), randomize_norms=True), F64)
    inputs = {k[8:]: t(g3[k]) for k in g3.files if k.startswith("full/in/")}
    with torch.no_grad():
        emb = om.embed_track_pos_visible_3d(p, cfg, inputs["support_tracks"], inputs["support_tracks_visible"],
                                            inputs["dino_features"], inputs["depth_features"])
        support = om.encode_tracks_3d(p, cfg, inputs["support_tracks"], inputs["support_tracks_visible"], inputs["boundary_frame"],
                                      inputs["dino_features"], inputs["depth_features"])
        lat = om.encode_3d(p, cfg, inputs)
        res = om.forward_3d(p, cfg, inputs, noise=t(g3["full/noise"]))
    close(emb, g3["full/emb"])
    close(support, g3["full/support_tokens"])
    close(lat, g3["full/latents"])
    close(res.tracks, g3["full/tracks"], rtol=1e-8, atol=1e-8)
    close(res.visible_logits, g3["full/visible_logits"], rtol=1e-8, atol=1e-8)


def test_loss_and_schedule_as_written(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_loss.npz"))
    pred = om.Results(t(g["3d/pred_tracks"]), t(g["3d/pred_visible_logits"]), torch.zeros(1))
    r = om.compute_loss_3d(pred, {"query_tracks": t(g["3d/tracks"]), "query_tracks_visible": t(g["3d/visible"])})
    for k in ("total_loss", "position_loss", "visible_loss"):
        close(r[k], g[f"3d/{k}"], rtol=1e-12, atol=0)
    pred = om.Results(t(g["none/pred_tracks"]), t(g["none/pred_visible_logits"]), torch.zeros(1))
    r = om.compute_loss_3d(pred, {"query_tracks": torch.zeros(1, 2, 3, 3, dtype=F64), "query_tracks_visible": torch.zeros(1, 2, 3, 1, dtype=F64)})
    close(r["visible_loss"], g["none/visible_loss"], rtol=1e-12, atol=0)
    for s, v in zip(g["lr/steps"], g["lr/values"]):
        assert om.learning_rate(int(s)) == pytest.approx(float(v), rel=1e-12, abs=1e-18)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources not present (GPU box)")
def test_generator_reproduces_committed_transformer_golden(g_tr):
    """Where the reference is present, re-execute its attention.py on the stand-ins and compare with the committed file."""
    from oracle import flax_shim as fs

    ref = fs.load_reference()
    d, qkv, heads, mlp, layers = (int(v) for v in g_tr["self/arch"])
    s0, s1 = (int(v) for v in g_tr["self/seeds"])
    p = om._transformer_init(np.random.RandomState(s0), d, qkv, heads, mlp, layers)
    om._randomize(p, np.random.RandomState(s1))
    p64 = om.flatten(p)
    tree = {}
    for path, v in p64.items():
        node = tree
        parts = path.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = np.asarray(v, np.float64)
    tr = fs.bind(ref["attention"].ImprovedTransformer(qkv_size=qkv, num_heads=heads, mlp_size=mlp, num_layers=layers), tree)
    x = g_tr["self/x"].copy()
    y = tr(x, qq_mask=g_tr["self/mask"])
    np.testing.assert_array_equal(x, g_tr["self/x"])           # inputs are not written through (jax immutability)
    np.testing.assert_allclose(np.asarray(y), g_tr["self/y"], rtol=1e-12, atol=1e-12)
