"""torch custom-op registration of the 3DSPA kernels: the ``spa3d::`` operator library.

SURVEY.md 8b, decision D-1: the reference is Flax/JAX, but neither jaxlib nor the XLA-FFI headers exist in this
environment, so the framework the kernels are REGISTERED with is torch (the XLA-FFI binder is ``csrc/spa3d_ffi.cc``,
unbuilt here).  Every operator below is defined with ``torch.library`` (schema + CUDA implementation + fake/meta
implementation + autograd formula where the op is differentiable), so it is visible to the dispatcher: it works under
``FakeTensorMode`` / ``torch.compile`` tracing, ``torch.library.opcheck`` and ``TorchDispatchMode``, and a caller composing the
ops directly gets gradients without the training engine.  The CUDA implementations are the C-ABI calls of ``ops.py`` (plain
device pointers, sizes and the current stream into lib3dspa_b200.so); there is no CPU implementation on purpose.

``ops.py`` routes its public functions through ``torch.ops.spa3d.*`` (``install`` below rebinds them), so the model,
the training engine, the tests and bench.py all reach the kernels through the dispatcher (about 2-7 us per call).

Replaces, per operator: the XLA-emitted ops behind ``nn.Dense`` / ``nn.DenseGeneral`` (attention.py:106-107,154-183),
``nn.LayerNorm`` (:49,76,103), ``nn.RMSNorm`` (:166-167), ``nn.dot_product_attention`` (:175), the embedding of
track_autoencoder_3d.py:123-149, the lifting of inference.py:287-447 and the loss of train.py:96-129.
"""
from __future__ import annotations

import math

import torch
from torch.library import Library, register_autograd, register_fake

_lib = Library("spa3d", "DEF")
_raw = {}          # name -> the ctypes-level implementation captured from ops.py by install()

F32, BF16 = torch.float32, torch.bfloat16


def _none(t):
    """Optional tensor results travel through the dispatcher as empty tensors."""
    return None if t is None or t.numel() == 0 else t


def _e(like):
    return like.new_empty(0)


# ---- schemas ---------------------------------------------------------------------------------------------------
_lib.define("gemm(Tensor a, Tensor wt, Tensor? bias, int act, Tensor? residual, ScalarType out_dtype, int impl) -> Tensor")
_lib.define("gemm_out(Tensor a, Tensor wt, Tensor? bias, int act, Tensor? residual, Tensor(a!) out, int impl) -> ()")
_lib.define("gemm_rmsnorm(Tensor a, Tensor wt, int Dh, int q_cols, int k_cols, Tensor scale_q, Tensor scale_k, bool save_rstd, int impl)"
            " -> (Tensor, Tensor)")
_lib.define("gemm_rmsnorm_out(Tensor a, Tensor wt, int Dh, int q_cols, int k_cols, Tensor scale_q, Tensor scale_k, Tensor(a!) out,"
            " Tensor(b!)? rstd, int impl) -> ()")
_lib.define("gemm_gelu(Tensor a, Tensor wt, Tensor bias, int impl, bool save_grad) -> (Tensor, Tensor)")
_lib.define("gemm_gelu_out(Tensor a, Tensor wt, Tensor bias, Tensor(a!) z, Tensor(b!) h, int impl, bool save_grad) -> ()")
_lib.define("layernorm_fwd_out(Tensor x, Tensor scale, Tensor(a!) out, Tensor(b!)? mean, Tensor(c!)? rstd, int rows, int ldx, int d) -> ()")
_lib.define("mlp_fused(Tensor a, Tensor w1t, Tensor b1, Tensor w2t, Tensor b2, Tensor residual) -> Tensor")
_lib.define("gemm_gelu_bwd(Tensor dy, Tensor wt, Tensor z, int impl, bool z_is_grad, Tensor(a!)? dz_colsum) -> Tensor")
_lib.define("gemm_dw(Tensor dy, Tensor x, Tensor(a!) dw, bool accumulate, int impl) -> ()")
_lib.define("attention_fwd(Tensor q, Tensor k, Tensor v, Tensor(a!) out, int batch, int heads, int Lq, int Lk, int Dh, Tensor? key_mask,"
            " bool save_stats) -> Tensor")
_lib.define("attention(Tensor q, Tensor k, Tensor v, Tensor? key_mask, int batch, int heads, int Lq, int Lk, int Dh) -> (Tensor, Tensor)")
_lib.define("attention_bwd(Tensor q, Tensor k, Tensor v, Tensor o, Tensor d_o, Tensor(a!) dq, Tensor(b!) dk, Tensor(c!) dv, Tensor stats,"
            " int batch, int heads, int Lq, int Lk, int Dh, Tensor? key_mask) -> ()")
_lib.define("layernorm_fwd(Tensor x, Tensor scale, ScalarType out_dtype, int rows, int ldx, int d, bool stats) -> (Tensor, Tensor, Tensor)")
_lib.define("layernorm_bwd(Tensor x, Tensor scale, Tensor mean, Tensor rstd, Tensor dy, Tensor(a!) dx, int rows, int ldx, int lddx, int d,"
            " bool accumulate, int num_partials, Tensor(b!)? dx_lowp, Tensor(c!)? dscale_accum, Tensor(d!)? dx_colsum) -> Tensor")
_lib.define("embed_fused_out(Tensor tracks, Tensor? dino, Tensor? depth, Tensor wt, Tensor bias, Tensor(a!) out, int T, int num_freq,"
            " float scale_factor, Tensor(b!)? a_cat) -> ()")
_lib.define("embed_fused(Tensor tracks, Tensor? dino, Tensor? depth, Tensor wt, Tensor bias, Tensor readout, int T, int num_freq,"
            " float scale_factor) -> (Tensor, Tensor)")
_lib.define("lift_sample(Tensor tracks_2d, Tensor? depth, Tensor? dino, int video_h, int video_w, float[]? intrinsics, ScalarType out_dtype,"
            " int depth_feature_dim, bool want_xyz, bool want_dino, bool want_depth) -> (Tensor, Tensor, Tensor)")
_lib.define("loss_fwd(Tensor head_out, Tensor target_tracks, Tensor target_vis, Tensor(a!) sums, int T) -> ()")
_lib.define("loss_bwd(Tensor head_out, Tensor target_tracks, Tensor target_vis, float l1_w, float bce_w, float inv_denom, int T) -> Tensor")
_lib.define("loss_sums(Tensor head_out, Tensor target_tracks, Tensor target_vis, int T) -> Tensor")

REGISTERED = ["mlp_fused", "gemm", "gemm_out", "gemm_rmsnorm", "gemm_rmsnorm_out", "gemm_gelu", "gemm_gelu_out", "layernorm_fwd_out", "gemm_gelu_bwd", "gemm_dw", "attention_fwd", "attention", "attention_bwd",
              "layernorm_fwd", "layernorm_bwd", "embed_fused_out", "embed_fused", "lift_sample", "loss_fwd", "loss_bwd", "loss_sums"]


# ---- CUDA implementations: the C-ABI calls of ops.py ---------------------------------------------------------------
def _gemm(a, wt, bias, act, residual, out_dtype, impl):
    return _raw["gemm"](a, wt, bias, act, residual, None, out_dtype, impl)


def _gemm_out(a, wt, bias, act, residual, out, impl):
    _raw["gemm"](a, wt, bias, act, residual, out, None, impl)


def _gemm_rmsnorm(a, wt, Dh, q_cols, k_cols, scale_q, scale_k, save_rstd, impl):
    r = _raw["gemm_rmsnorm"](a, wt, Dh, q_cols, k_cols, scale_q, scale_k, save_rstd, impl)
    return (r[0], r[1]) if save_rstd else (r, _e(scale_q))


def _gemm_gelu(a, wt, bias, impl, save_grad):
    return _raw["gemm_gelu"](a, wt, bias, impl, save_grad)


# out-variants (no autograd key, 2 us instead of 7 us per dispatch): what the engines call when no graph is being recorded
def _gemm_rmsnorm_out(a, wt, Dh, q_cols, k_cols, scale_q, scale_k, out, rstd, impl):
    _raw["gemm_rmsnorm"](a, wt, Dh, q_cols, k_cols, scale_q, scale_k, rstd is not None, impl, out, rstd)


def _gemm_gelu_out(a, wt, bias, z, h, impl, save_grad):
    _raw["gemm_gelu"](a, wt, bias, impl, save_grad, z, h)


def _layernorm_fwd_out(x, scale, out, mean, rstd, rows, ldx, d):
    _raw["layernorm_fwd"](x, scale, out.dtype, None if rows < 0 else rows, None if ldx < 0 else ldx, None if d < 0 else d, mean is not None, out,
                          mean, rstd)


def _mlp_fused(a, w1t, b1, w2t, b2, residual):
    return _raw["mlp_fused"](a, w1t, b1, w2t, b2, residual)


def _gemm_gelu_bwd(dy, wt, z, impl, z_is_grad, dz_colsum):
    return _raw["gemm_gelu_bwd"](dy, wt, z, impl, z_is_grad, dz_colsum)


def _gemm_dw(dy, x, dw, accumulate, impl):
    _raw["gemm_dw"](dy, x, dw, accumulate, impl)


def _attention_fwd(q, k, v, out, batch, heads, Lq, Lk, Dh, key_mask, save_stats):
    st = _raw["attention_fwd"](q, k, v, out, batch, heads, Lq, Lk, Dh, key_mask, save_stats)
    return st if st is not None else q.new_empty(0, dtype=F32)


def _attention(q, k, v, key_mask, batch, heads, Lq, Lk, Dh):
    o = torch.empty(batch * Lq, heads * Dh, device=q.device, dtype=q.dtype)
    st = _raw["attention_fwd"](q, k, v, o, batch, heads, Lq, Lk, Dh, key_mask, True)
    return o, st


def _attention_bwd(q, k, v, o, d_o, dq, dk, dv, stats, batch, heads, Lq, Lk, Dh, key_mask):
    _raw["attention_bwd"](q, k, v, o, d_o, dq, dk, dv, stats, batch, heads, Lq, Lk, Dh, key_mask)


def _layernorm_fwd(x, scale, out_dtype, rows, ldx, d, stats):
    r = _raw["layernorm_fwd"](x, scale, out_dtype, None if rows < 0 else rows, None if ldx < 0 else ldx, None if d < 0 else d, stats)
    if stats:
        return r
    return r, x.new_empty(0, dtype=F32), x.new_empty(0, dtype=F32)


def _layernorm_bwd(x, scale, mean, rstd, dy, dx, rows, ldx, lddx, d, accumulate, num_partials, dx_lowp, dscale_accum, dx_colsum):
    ds = _raw["layernorm_bwd"](x, scale, mean, rstd, dy, dx, None if rows < 0 else rows, None if ldx < 0 else ldx,
                               None if lddx < 0 else lddx, None if d < 0 else d, accumulate, num_partials, dx_lowp, dscale_accum, dx_colsum)
    return ds if ds is not None else scale.new_empty(0)


def _embed_fused_out(tracks, dino, depth, wt, bias, out, T, num_freq, scale_factor, a_cat):
    _raw["embed_fused"](tracks, dino, depth, wt, bias, out, T, num_freq, scale_factor, a_cat)


def _embed_fused(tracks, dino, depth, wt, bias, readout, T, num_freq, scale_factor):
    """tokens [seqs*(T+1), W] f32 with the read-out token in row 0 of every sequence, and the bf16 concatenated features."""
    rows = tracks.shape[0]
    seqs = rows // T
    out = torch.empty(rows + seqs, wt.shape[0], device=tracks.device, dtype=F32)
    a_cat = torch.empty(rows + seqs, wt.shape[1], device=tracks.device, dtype=wt.dtype)
    a_cat.view(seqs, T + 1, -1)[:, 0].zero_()
    _raw["embed_fused"](tracks, dino, depth, wt, bias, out, T, num_freq, scale_factor, a_cat)
    _raw["set_rows"](out, T + 1, readout.reshape(-1), seqs)
    return out, a_cat


def _lift_sample(tracks_2d, depth, dino, video_h, video_w, intrinsics, out_dtype, depth_feature_dim, want_xyz, want_dino, want_depth):
    vh = (video_h, video_w) if video_h > 0 else None
    xyz, df, zf = _raw["lift_sample"](tracks_2d, depth, dino, vh, intrinsics, out_dtype, depth_feature_dim, want_xyz, want_dino, want_depth)
    e = tracks_2d.new_empty(0)
    return (xyz if xyz is not None else e, df if df is not None else e, zf if zf is not None else e)


def _loss_fwd(head_out, target_tracks, target_vis, sums, T):
    _raw["loss_fwd"](head_out, target_tracks, target_vis, sums, T)


def _loss_bwd(head_out, target_tracks, target_vis, l1_w, bce_w, inv_denom, T):
    return _raw["loss_bwd"](head_out, target_tracks, target_vis, l1_w, bce_w, inv_denom, T)


def _loss_sums(head_out, target_tracks, target_vis, T):
    sums = torch.zeros(3, device=head_out.device, dtype=F32)
    _raw["loss_fwd"](head_out, target_tracks, target_vis, sums, T)
    return sums


for _n in REGISTERED:
    _lib.impl(_n, globals()["_" + _n], "CUDA")


# ---- fake (meta) implementations: shapes and dtypes only --------------------------------------------------------------
@register_fake("spa3d::gemm")
def _(a, wt, bias, act, residual, out_dtype, impl):
    return a.new_empty(a.shape[0], wt.shape[0], dtype=out_dtype)


@register_fake("spa3d::gemm_out")
def _(a, wt, bias, act, residual, out, impl):
    return None


@register_fake("spa3d::gemm_rmsnorm")
def _(a, wt, Dh, q_cols, k_cols, scale_q, scale_k, save_rstd, impl):
    M = a.shape[0]
    rstd = a.new_empty(M, (q_cols + k_cols) // Dh, dtype=F32) if save_rstd else scale_q.new_empty(0)
    return a.new_empty(M, wt.shape[0]), rstd


@register_fake("spa3d::gemm_rmsnorm_out")
def _(a, wt, Dh, q_cols, k_cols, scale_q, scale_k, out, rstd, impl):
    return None


@register_fake("spa3d::gemm_gelu_out")
def _(a, wt, bias, z, h, impl, save_grad):
    return None


@register_fake("spa3d::layernorm_fwd_out")
def _(x, scale, out, mean, rstd, rows, ldx, d):
    return None


@register_fake("spa3d::mlp_fused")
def _(a, w1t, b1, w2t, b2, residual):
    return torch.empty_like(residual)


@register_fake("spa3d::gemm_gelu")
def _(a, wt, bias, impl, save_grad):
    return a.new_empty(a.shape[0], wt.shape[0]), a.new_empty(a.shape[0], wt.shape[0])


@register_fake("spa3d::gemm_gelu_bwd")
def _(dy, wt, z, impl, z_is_grad, dz_colsum):
    return dy.new_empty(dy.shape[0], wt.shape[0])


@register_fake("spa3d::gemm_dw")
def _(dy, x, dw, accumulate, impl):
    return None


@register_fake("spa3d::attention_fwd")
def _(q, k, v, out, batch, heads, Lq, Lk, Dh, key_mask, save_stats):
    return q.new_empty((batch, heads, Lq, 2) if save_stats else (0,), dtype=F32)


@register_fake("spa3d::attention")
def _(q, k, v, key_mask, batch, heads, Lq, Lk, Dh):
    return q.new_empty(batch * Lq, heads * Dh), q.new_empty(batch, heads, Lq, 2, dtype=F32)


@register_fake("spa3d::attention_bwd")
def _(q, k, v, o, d_o, dq, dk, dv, stats, batch, heads, Lq, Lk, Dh, key_mask):
    return None


@register_fake("spa3d::layernorm_fwd")
def _(x, scale, out_dtype, rows, ldx, d, stats):
    r = x.shape[0] if rows < 0 else rows
    w = x.shape[-1] if d < 0 else d
    n = r if stats else 0
    return x.new_empty(r, w, dtype=out_dtype), x.new_empty(n, dtype=F32), x.new_empty(n, dtype=F32)


@register_fake("spa3d::layernorm_bwd")
def _(x, scale, mean, rstd, dy, dx, rows, ldx, lddx, d, accumulate, num_partials, dx_lowp, dscale_accum, dx_colsum):
    return scale.new_empty(0 if dscale_accum is not None else (x.shape[-1] if d < 0 else d))


@register_fake("spa3d::embed_fused_out")
def _(tracks, dino, depth, wt, bias, out, T, num_freq, scale_factor, a_cat):
    return None


@register_fake("spa3d::embed_fused")
def _(tracks, dino, depth, wt, bias, readout, T, num_freq, scale_factor):
    rows = tracks.shape[0]
    n = rows + rows // T
    return tracks.new_empty(n, wt.shape[0]), tracks.new_empty(n, wt.shape[1], dtype=wt.dtype)


@register_fake("spa3d::lift_sample")
def _(tracks_2d, depth, dino, video_h, video_w, intrinsics, out_dtype, depth_feature_dim, want_xyz, want_dino, want_depth):
    N, T = tracks_2d.shape[:2]
    e = tracks_2d.new_empty(0)
    xyz = tracks_2d.new_empty(N, T, 3) if (depth is not None and want_xyz) else e
    zf = tracks_2d.new_empty(N, T, depth_feature_dim, dtype=out_dtype) if (depth is not None and want_depth) else e
    df = tracks_2d.new_empty(N, T, dino.shape[3], dtype=out_dtype) if (dino is not None and want_dino) else e
    return xyz, df, zf


@register_fake("spa3d::loss_fwd")
def _(head_out, target_tracks, target_vis, sums, T):
    return None


@register_fake("spa3d::loss_bwd")
def _(head_out, target_tracks, target_vis, l1_w, bce_w, inv_denom, T):
    return torch.empty_like(head_out)


@register_fake("spa3d::loss_sums")
def _(head_out, target_tracks, target_vis, T):
    return head_out.new_empty(3)


# ---- autograd formulas (functional operators only) -----------------------------------------------------------------------
def _t(w):
    """[K,N] copy of a [N,K] weight for dX = dY . W (the training engine keeps these copies resident instead)."""
    out = torch.empty(w.shape[1], w.shape[0], device=w.device, dtype=w.dtype)
    if w.dtype == F32:
        _raw["shadow_weights"](w, None, out)
    else:
        out.copy_(w.t())
    return out


def _dw(dy, x, like):
    g = torch.empty(like.shape, device=like.device, dtype=F32)
    torch.ops.spa3d.gemm_dw(dy.contiguous(), x, g, False, 0)
    return g.to(like.dtype)


def _colsum(dy, like):
    out = torch.empty(dy.shape[1], device=dy.device, dtype=F32)
    _raw["colsum"](dy.contiguous(), out)
    return out.to(like.dtype)


def _gemm_setup(ctx, inputs, output):
    a, wt, bias, act, residual, out_dtype, impl = inputs
    ctx.save_for_backward(a, wt, bias if bias is not None else a.new_empty(0))
    ctx.act, ctx.has_bias, ctx.res = act, bias is not None, (residual.dtype if residual is not None else None)


def _gemm_backward(ctx, dy):
    a, wt, bias = ctx.saved_tensors
    g = dy.contiguous().to(a.dtype)
    if ctx.act:   # y = gelu(z) (+ residual): recompute the pre-activation (use spa3d::gemm_gelu to save it instead)
        z = torch.ops.spa3d.gemm(a, wt, bias if ctx.has_bias else None, 0, None, a.dtype, 0)
        gz = torch.empty_like(z)
        _raw["gelu_bwd"](z, g, gz)
        g = gz
    nd = ctx.needs_input_grad
    da = torch.ops.spa3d.gemm(g, _t(wt), None, 0, None, a.dtype, 0) if nd[0] else None
    dwt = _dw(g, a, wt) if nd[1] else None
    db = _colsum(g, bias) if (ctx.has_bias and nd[2]) else None
    dres = dy.to(ctx.res) if (ctx.res is not None and nd[4]) else None
    return da, dwt, db, None, dres, None, None


register_autograd("spa3d::gemm", _gemm_backward, setup_context=_gemm_setup)


def _rms_setup(ctx, inputs, output):
    a, wt, Dh, q_cols, k_cols, scale_q, scale_k, save_rstd, impl = inputs
    out, rstd = output
    if not save_rstd:
        ctx.ok = False
        return
    ctx.ok = True
    ctx.save_for_backward(a, wt, scale_q, scale_k, out, rstd)
    ctx.dims = (Dh, q_cols, k_cols)


def _rms_backward(ctx, dout, drstd):
    if not ctx.ok:
        raise RuntimeError("spa3d::gemm_rmsnorm: call with save_rstd=True to differentiate through it")
    a, wt, sq, sk, out, rstd = ctx.saved_tensors
    Dh, qc, kc = ctx.dims
    d = dout.contiguous().to(out.dtype).clone()
    hq, hk = qc // Dh, kc // Dh
    dsq = dsk = None
    if qc:
        dsq = _raw["head_rmsnorm_bwd"](out[:, :qc], sq, 1.0 / math.sqrt(Dh), rstd[:, :hq], d[:, :qc], hq, Dh)
    if kc:
        dsk = _raw["head_rmsnorm_bwd"](out[:, qc : qc + kc], sk, 1.0, rstd[:, hq:], d[:, qc : qc + kc], hk, Dh)
    nd = ctx.needs_input_grad
    da = torch.ops.spa3d.gemm(d, _t(wt), None, 0, None, a.dtype, 0) if nd[0] else None
    dwt = _dw(d, a, wt) if nd[1] else None
    return da, dwt, None, None, None, dsq, dsk, None, None


register_autograd("spa3d::gemm_rmsnorm", _rms_backward, setup_context=_rms_setup)


def _gelu_setup(ctx, inputs, output):
    a, wt, bias, impl, save_grad = inputs
    z, h = output
    ctx.save_for_backward(a, wt, z)
    ctx.zg = bool(save_grad)


def _gelu_backward(ctx, dz_in, dh):
    a, wt, z = ctx.saved_tensors
    g = dh.contiguous().to(z.dtype)
    dz = torch.empty_like(z)
    if ctx.zg:
        torch.mul(g, z, out=dz)      # z holds gelu'(z); (this generic formula is not the training engine's fused path)
    else:
        _raw["gelu_bwd"](z, g, dz)
    if dz_in is not None and not ctx.zg:
        dz = dz + dz_in.to(dz.dtype)
    nd = ctx.needs_input_grad
    da = torch.ops.spa3d.gemm(dz, _t(wt), None, 0, None, a.dtype, 0) if nd[0] else None
    dwt = _dw(dz, a, wt) if nd[1] else None
    db = _colsum(dz, dz.float()) if nd[2] else None
    return da, dwt, db, None, None


register_autograd("spa3d::gemm_gelu", _gelu_backward, setup_context=_gelu_setup)


def _attn_setup(ctx, inputs, output):
    q, k, v, key_mask, batch, heads, Lq, Lk, Dh = inputs
    o, stats = output
    ctx.save_for_backward(q, k, v, o, stats, key_mask if key_mask is not None else q.new_empty(0, dtype=torch.uint8))
    ctx.dims = (batch, heads, Lq, Lk, Dh, key_mask is not None)


def _attn_backward(ctx, d_o, dstats):
    q, k, v, o, stats, km = ctx.saved_tensors
    batch, heads, Lq, Lk, Dh, has_mask = ctx.dims
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    torch.ops.spa3d.attention_bwd(q, k, v, o, d_o.contiguous().to(q.dtype), dq, dk, dv, stats, batch, heads, Lq, Lk, Dh,
                                  km if has_mask else None)
    return dq, dk, dv, None, None, None, None, None, None


register_autograd("spa3d::attention", _attn_backward, setup_context=_attn_setup)


def _ln_setup(ctx, inputs, output):
    x, scale, out_dtype, rows, ldx, d, stats = inputs
    y, mean, rstd = output
    ctx.ok = bool(stats) and rows < 0 and ldx < 0 and d < 0
    if ctx.ok:
        ctx.save_for_backward(x, scale, mean, rstd)


def _ln_backward(ctx, dy, dmean, drstd):
    if not ctx.ok:
        raise RuntimeError("spa3d::layernorm_fwd: call with stats=True on a plain [rows, d] matrix to differentiate through it")
    x, scale, mean, rstd = ctx.saved_tensors
    dx = torch.empty_like(x)
    ds = torch.ops.spa3d.layernorm_bwd(x, scale, mean, rstd, dy.contiguous(), dx, -1, -1, -1, -1, False, 296, None, None, None)
    return dx, ds, None, None, None, None, None


register_autograd("spa3d::layernorm_fwd", _ln_backward, setup_context=_ln_setup)


def _embed_setup(ctx, inputs, output):
    tracks, dino, depth, wt, bias, readout, T, num_freq, scale_factor = inputs
    out, a_cat = output
    ctx.save_for_backward(a_cat, wt)
    ctx.T = T


def _embed_backward(ctx, dx, dacat):
    a_cat, wt = ctx.saved_tensors
    T = ctx.T
    dx = dx.contiguous().clone()
    W = dx.shape[1]
    seqs = dx.shape[0] // (T + 1)
    d_readout = torch.empty(W, device=dx.device, dtype=F32)
    _raw["colsum"](dx.view(seqs, (T + 1) * W)[:, :W], d_readout)
    _raw["set_rows"](dx, T + 1, torch.zeros(W, device=dx.device), seqs)     # the read-out rows do not reach the GEMM
    db = torch.empty(W, device=dx.device, dtype=F32)
    _raw["colsum"](dx, db)
    dwt = _dw(dx.to(a_cat.dtype), a_cat, wt)
    return None, None, None, dwt, db, d_readout.view(1, W), None, None, None


register_autograd("spa3d::embed_fused", _embed_backward, setup_context=_embed_setup)


def _loss_setup(ctx, inputs, output):
    head_out, target_tracks, target_vis, T = inputs
    ctx.save_for_backward(head_out, target_tracks, target_vis)
    ctx.T = T


def _loss_backward(ctx, g):
    head_out, tt, tv = ctx.saved_tensors
    w = g.detach().float().cpu().tolist()       # d total / d (position sum, bce sum, visible count): scalars of the C ABI
    return torch.ops.spa3d.loss_bwd(head_out, tt, tv, w[0], w[1], 1.0, ctx.T), None, None, None


register_autograd("spa3d::loss_sums", _loss_backward, setup_context=_loss_setup)


# ---- ops.py public names -> dispatcher -------------------------------------------------------------------------------------
def install(ns):
    """Capture the ctypes-level implementations of ops.py and rebind its public names to ``torch.ops.spa3d.*``."""
    for n in ("gemm", "gemm_rmsnorm", "gemm_gelu", "gemm_gelu_bwd", "gemm_dw", "attention_fwd", "attention_bwd", "layernorm_fwd",
              "layernorm_bwd", "embed_fused", "lift_sample", "loss_fwd", "loss_bwd", "colsum", "set_rows", "gelu_bwd", "head_rmsnorm_bwd",
              "shadow_weights", "mlp_fused"):
        _raw[n] = ns[n]
    o = torch.ops.spa3d

    grad_on = torch.is_grad_enabled
    empty = torch.empty

    def gemm(a, wt, bias=None, act=0, residual=None, out=None, out_dtype=None, impl=0):
        if out is None:
            if grad_on() or a.dtype != wt.dtype:      # a graph may be recorded (functional op with its autograd formula), or the bf16 x 3 form
                return o.gemm(a, wt, bias, int(act), residual, out_dtype or a.dtype, int(impl))
            out = empty(a.shape[0], wt.shape[0], device=a.device, dtype=out_dtype or a.dtype)
        o.gemm_out(a, wt, bias, int(act), residual, out, int(impl))
        return out

    def gemm_rmsnorm(a, wt, Dh, q_cols, k_cols, scale_q, scale_k, save_rstd=False, impl=0):
        if grad_on() or a.dtype != wt.dtype:
            out, rstd = o.gemm_rmsnorm(a, wt, int(Dh), int(q_cols), int(k_cols), scale_q, scale_k, bool(save_rstd), int(impl))
            return (out, rstd) if save_rstd else out
        out = empty(a.shape[0], wt.shape[0], device=a.device, dtype=a.dtype)
        rstd = empty(a.shape[0], (q_cols + k_cols) // Dh, device=a.device, dtype=F32) if save_rstd else None
        o.gemm_rmsnorm_out(a, wt, int(Dh), int(q_cols), int(k_cols), scale_q, scale_k, out, rstd, int(impl))
        return (out, rstd) if save_rstd else out

    def gemm_gelu(a, wt, bias, impl=0, save_grad=False):
        if grad_on():
            return o.gemm_gelu(a, wt, bias, int(impl), bool(save_grad))
        z = empty(a.shape[0], wt.shape[0], device=a.device, dtype=a.dtype)
        h = empty(a.shape[0], wt.shape[0], device=a.device, dtype=a.dtype)
        o.gemm_gelu_out(a, wt, bias, z, h, int(impl), bool(save_grad))
        return z, h

    def gemm_gelu_bwd(dy, wt, z, impl=0, z_is_grad=False, dz_colsum=None):
        return o.gemm_gelu_bwd(dy, wt, z, int(impl), bool(z_is_grad), dz_colsum)

    def mlp_fused(a, w1t, b1, w2t, b2, residual):
        return o.mlp_fused(a, w1t, b1, w2t, b2, residual)

    def gemm_dw(dy, x, dw, accumulate=True, impl=0):
        o.gemm_dw(dy, x, dw, bool(accumulate), int(impl))
        return dw

    def attention_fwd(q, k, v, out, batch, heads, Lq, Lk, Dh, key_mask=None, save_stats=False):
        return _none(o.attention_fwd(q, k, v, out, int(batch), int(heads), int(Lq), int(Lk), int(Dh), key_mask, bool(save_stats)))

    def attention_bwd(q, k, v, o_, d_o, dq, dk, dv, stats, batch, heads, Lq, Lk, Dh, key_mask=None):
        o.attention_bwd(q, k, v, o_, d_o, dq, dk, dv, stats, int(batch), int(heads), int(Lq), int(Lk), int(Dh), key_mask)

    def layernorm_fwd(x, scale, out_dtype, rows=None, ldx=None, d=None, stats=False, out=None):
        if out is not None or not grad_on():
            r_ = x.shape[0] if rows is None else int(rows)
            d_ = x.shape[-1] if d is None else int(d)
            if out is None:
                out = empty(r_, d_, device=x.device, dtype=out_dtype)
            mean = empty(r_, device=x.device, dtype=F32) if stats else None
            rstd = empty(r_, device=x.device, dtype=F32) if stats else None
            o.layernorm_fwd_out(x, scale, out, mean, rstd, -1 if rows is None else int(rows), -1 if ldx is None else int(ldx),
                                -1 if d is None else int(d))
            return (out, mean, rstd) if stats else out
        y, mean, rstd = o.layernorm_fwd(x, scale, out_dtype, -1 if rows is None else int(rows), -1 if ldx is None else int(ldx),
                                        -1 if d is None else int(d), bool(stats))
        return (y, mean, rstd) if stats else y

    def layernorm_bwd(x, scale, mean, rstd, dy, dx, rows=None, ldx=None, lddx=None, d=None, accumulate=False, num_partials=296,
                      dx_lowp=None, dscale_accum=None, dx_colsum=None):
        ds = o.layernorm_bwd(x, scale, mean, rstd, dy, dx, -1 if rows is None else int(rows), -1 if ldx is None else int(ldx),
                             -1 if lddx is None else int(lddx), -1 if d is None else int(d), bool(accumulate), int(num_partials),
                             dx_lowp, dscale_accum, dx_colsum)
        return None if dscale_accum is not None else ds

    def embed_fused(tracks, dino, depth, wt, bias, out, T, num_freq, scale_factor, a_cat=None):
        o.embed_fused_out(tracks, dino, depth, wt, bias, out, int(T), int(num_freq), float(scale_factor), a_cat)
        return out

    def lift_sample(tracks_2d, depth=None, dino=None, video_hw=None, intrinsics=None, out_dtype=F32, depth_feature_dim=256,
                    want_xyz=True, want_dino=True, want_depth=True):
        vh, vw = (int(video_hw[0]), int(video_hw[1])) if video_hw is not None else (0, 0)
        intr = [float(v) for v in intrinsics] if intrinsics is not None else None
        xyz, df, zf = o.lift_sample(tracks_2d, depth, dino, vh, vw, intr, out_dtype, int(depth_feature_dim), bool(want_xyz),
                                    bool(want_dino), bool(want_depth))
        return _none(xyz), _none(df), _none(zf)

    def loss_fwd(head_out, target_tracks, target_vis, sums, T):
        o.loss_fwd(head_out, target_tracks, target_vis, sums, int(T))
        return sums

    def loss_bwd(head_out, target_tracks, target_vis, l1_w, bce_w, inv_denom, T):
        return o.loss_bwd(head_out, target_tracks, target_vis, float(l1_w), float(bce_w), float(inv_denom), int(T))

    for fn in (mlp_fused, gemm, gemm_rmsnorm, gemm_gelu, gemm_gelu_bwd, gemm_dw, attention_fwd, attention_bwd, layernorm_fwd, layernorm_bwd,
               embed_fused, lift_sample, loss_fwd, loss_bwd):
        fn.__doc__ = (_raw[fn.__name__].__doc__ or "") + "\n    (dispatched through torch.ops.spa3d." + fn.__name__ + ")"
        ns[fn.__name__] = fn
