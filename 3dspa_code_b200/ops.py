"""Tensor-level wrappers over the C ABI (include/spa3d_b200.h).

PyTorch is used here only as plumbing: device memory (caching allocator), streams and dtypes.
Every function passes raw device pointers, sizes and the current CUDA stream to
lib3dspa_b200.so; nothing here computes on the host or falls back to torch kernels.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2

_DT = {torch.float32: F32, torch.bfloat16: BF16}
_TD = {F32: torch.float32, BF16: torch.bfloat16}

launch_count = 0  # number of kernel-launching C-ABI calls issued (bench.py reports it)


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} (float32 / bfloat16 only)") from None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("spa3d ops need CUDA tensors (there is no CPU path)")
    return ctypes.c_void_p(t.data_ptr())


def _ld(t):
    if t.dim() < 2:
        return t.shape[-1] if t.dim() == 1 else 1
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        raise ValueError("last dimension must be contiguous")
    return t.stride(-2)


def _call(name, *args):
    global launch_count
    launch_count += 1
    _lib.check(getattr(_lib.lib(), name)(*args), name)


def version():
    return _lib.lib().spa3d_version()


# ---- K0 ------------------------------------------------------------------------------------
def lift_sample(tracks_2d, depth=None, dino=None, video_hw=None, intrinsics=None, out_dtype=torch.float32,
                depth_feature_dim=256, want_xyz=True, want_dino=True, want_depth=True):
    """tracks_2d [N,T,2] f32, depth [T,H,W,(1)] f32, dino [T,Hp,Wp,D] f32 -> (xyz, dino_feat, depth_feat)."""
    N, T = tracks_2d.shape[:2]
    dev = tracks_2d.device
    H = W = Hp = Wp = D = 0
    xyz = dfeat = zfeat = None
    if depth is not None:
        H, W = depth.shape[1:3]
        if want_xyz:
            xyz = torch.empty(N, T, 3, device=dev, dtype=torch.float32)
        if want_depth:
            zfeat = torch.empty(N, T, depth_feature_dim, device=dev, dtype=out_dtype)
    if dino is not None and want_dino:
        Hp, Wp, D = dino.shape[1:4]
        dfeat = torch.empty(N, T, D, device=dev, dtype=out_dtype)
    vh, vw = video_hw if video_hw is not None else (H, W)
    intr = None
    if intrinsics is not None:
        intr = (ctypes.c_float * 4)(*[float(v) for v in intrinsics])
    ws, ws_bytes = None, 0
    if dfeat is not None and D % 128 == 0:   # cell-binned gather (corner rows read once per occupied patch cell): caller-owned scratch
        ws_bytes = int(_lib.lib().spa3d_lift_workspace_bytes(N, T, Hp, Wp))
        ws = torch.empty(ws_bytes // 4, device=dev, dtype=torch.int32)
    _call("spa3d_lift_sample_ws", _p(tracks_2d), _p(depth), _p(dino), _p(xyz), _p(dfeat), _p(zfeat), _DT[out_dtype],
          N, T, H, W, Hp, Wp, D, depth_feature_dim, vh, vw, ctypes.cast(intr, ctypes.c_void_p) if intr else None, _p(ws), ws_bytes,
          _stream())
    return xyz, dfeat, zfeat


# ---- elementwise ---------------------------------------------------------------------------
def fourier_features(x, out, num_freq=32, scale_factor=1.0, append_time=0, tail_zero=False, exact=True, out_row_group=0):
    rows, C = x.shape
    _call("spa3d_fourier_features", _p(x), _ld(x), _p(out), _ld(out), dt(out), rows, C, num_freq, float(scale_factor),
          int(append_time), int(tail_zero), int(exact), int(out_row_group), _stream())
    return out


def embed_fused_applicable(W, K_total, dino_dim, depth_dim, coords):
    return bool(_lib.lib().spa3d_embed_fused_applicable(int(W), int(K_total), int(dino_dim), int(depth_dim), int(coords)))


def embed_fused(tracks, dino, depth, wt, bias, out, T, num_freq, scale_factor, a_cat=None):
    """out[r + r//T + 1] = [Fourier(tracks[r], t/T) | dino[r] | depth[r]] @ wt^T + bias (fp32 out); a_cat (bf16,
    optional) receives the concatenated features at the same rows."""
    rows = tracks.shape[0]
    _call("spa3d_embed_fused", _p(tracks), _p(dino), _p(depth), _p(wt), _ld(wt), _p(bias), _p(out), _ld(out), _p(a_cat),
          _ld(a_cat) if a_cat is not None else 0, rows, int(T),
          dino.shape[1] if dino is not None else 0, depth.shape[1] if depth is not None else 0, wt.shape[0], int(num_freq),
          float(scale_factor), _stream())
    return out


def embed_sampled(xyz, tracks_2d, dfeat, proj, wt, wdep, bias, out, T, Hp, Wp, video_hw, num_freq, scale_factor):
    """Tokens from track points + a PROJECTED patch map (see spa3d_embed_sampled in include/spa3d_b200.h)."""
    rows = xyz.shape[0]
    _call("spa3d_embed_sampled", _p(xyz), _p(tracks_2d), _p(dfeat), _p(proj), _p(wt), _ld(wt), _p(wdep), _p(bias), _p(out), _ld(out),
          rows, int(T), int(Hp), int(Wp), int(video_hw[0]), int(video_hw[1]), wt.shape[0], int(num_freq), float(scale_factor), _stream())
    return out


def convert(src, dst, out_row_group=0):
    rows, cols = src.shape
    _call("spa3d_convert", _p(src), _ld(src), dt(src), _p(dst), _ld(dst), dt(dst), rows, cols, int(out_row_group), _stream())
    return dst


def set_rows(dst, row_stride, vec, rows):
    _call("spa3d_set_rows", _p(dst), _ld(dst), dt(dst), int(row_stride), _p(vec), int(rows), vec.numel(), _stream())
    return dst


def gemm_x3_applicable(M, N, K):
    return bool(_lib.lib().spa3d_gemm_x3_applicable(int(M), int(N), int(K)))


def split3(x, out=None):
    """[rows, K] fp32 -> [rows, 3K] bf16 (hi | mid | lo): the operand form of gemm_x3."""
    rows, K = x.shape
    if out is None:
        out = torch.empty(rows, 3 * K, device=x.device, dtype=torch.bfloat16)
    _call("spa3d_split3", _p(x), _ld(x), _p(out), _ld(out), int(rows), int(K), _stream())
    return out


def _is_split(a, wt):
    """fp32 activations against a three-term bf16 split weight [N, 3K]: the "bf16 x 3" tensor-core form of the accurate mode."""
    return a.dtype == torch.float32 and wt.dtype == torch.bfloat16 and wt.shape[1] == 3 * a.shape[1]


def _gemm_x3(a, w3, bias, act, residual, out):
    M, K = a.shape
    N = w3.shape[0]
    a3 = split3(a)
    z = out if (out is not None and act == ACT_NONE) else torch.empty(M, N, device=a.device, dtype=torch.float32)
    assert z.dtype == torch.float32, "the bf16 x 3 contraction writes float32"
    res = residual if act == ACT_NONE else None
    _call("spa3d_gemm_x3", _p(a3), _ld(a3), _p(w3), _ld(w3), _p(bias), _p(res), _ld(res) if res is not None else 0,
          dt(res) if res is not None else F32, _p(z), _ld(z), M, N, K, _stream())
    if act == ACT_NONE:
        return z
    h = out if out is not None else torch.empty_like(z)     # exact tanh-GELU on the fp32 pre-activation, then the residual
    gelu_fwd(z, h)
    if residual is not None:
        axpy(h, residual.contiguous())
    return h


def gemm(a, wt, bias=None, act=ACT_NONE, residual=None, out=None, out_dtype=None, impl=GEMM_AUTO):
    """out[M,N] = act(a[M,K] @ wt[N,K]^T + bias) + residual.  (fp32 ``a`` with a split3 weight: the bf16 x 3 contraction.)"""
    if _is_split(a, wt):
        return _gemm_x3(a, wt, bias, act, residual, out)
    M, K = a.shape
    N = wt.shape[0]
    assert wt.shape[1] == K and wt.dtype == a.dtype, (a.shape, wt.shape, a.dtype, wt.dtype)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype or a.dtype)
    _call("spa3d_gemm", _p(a), _ld(a), _p(wt), _ld(wt), dt(a), _p(bias), int(act), _p(residual),
          _ld(residual) if residual is not None else 0, dt(residual) if residual is not None else F32,
          _p(out), _ld(out), dt(out), M, N, K, int(impl), _stream())
    return out


def mlp_fused_applicable(D, Hd):
    return bool(_lib.lib().spa3d_mlp_fused_applicable(int(D), int(Hd)))


def mlp_fused(a, w1t, b1, w2t, b2, residual, out=None):
    """out = residual + gelu_tanh(a @ w1t^T + b1) @ w2t^T + b2 in one kernel (bf16 a / weights, fp32 residual and out)."""
    M, D = a.shape
    Hd = w1t.shape[0]
    if out is None:
        out = torch.empty(M, D, device=a.device, dtype=torch.float32)
    _call("spa3d_mlp_fused", _p(a), _ld(a), _p(w1t), _ld(w1t), _p(b1), _p(w2t), _ld(w2t), _p(b2), _p(residual), _ld(residual), _p(out), _ld(out),
          int(M), int(D), int(Hd), _stream())
    return out


def gemm_gelu(a, wt, bias, impl=GEMM_AUTO, save_grad=False, z=None, h=None):
    """(z, h): z = a @ wt^T + bias, h = gelu_tanh(z), both in a.dtype (one GEMM, two outputs).
    save_grad (bf16 tensor-core path): the first output is gelu_tanh'(z) instead of z (pair it with gemm_gelu_bwd(z_is_grad=True))."""
    M, K = a.shape
    N = wt.shape[0]
    if z is None:
        z = torch.empty(M, N, device=a.device, dtype=a.dtype)
        h = torch.empty(M, N, device=a.device, dtype=a.dtype)
    _call("spa3d_gemm_gelu", _p(a), _ld(a), _p(wt), _ld(wt), dt(a), _p(bias), _p(z), _ld(z), _p(h), _ld(h), M, N, K, int(bool(save_grad)),
          int(impl), _stream())
    return z, h


def gemm_gelu_bwd(dy, wt, z, impl=GEMM_AUTO, z_is_grad=False, dz_colsum=None):
    """dz = (dy @ wt^T) * gelu_tanh'(z)   (z_is_grad: ``z`` already holds gelu_tanh'(z)).
    dz_colsum ([N] fp32, tensor-core path): += column sums of dz (the gradient of MLP_in's bias) from the GEMM epilogue."""
    M, K = dy.shape
    N = wt.shape[0]
    dz = torch.empty(M, N, device=dy.device, dtype=dy.dtype)
    _call("spa3d_gemm_gelu_bwd", _p(dy), _ld(dy), _p(wt), _ld(wt), dt(dy), _p(z), _ld(z), _p(dz), _ld(dz), M, N, K,
          int(bool(z_is_grad)), _p(dz_colsum), int(impl), _stream())
    return dz


def gemm_dw(dy, x, dw, accumulate=True, impl=GEMM_AUTO):
    """dw[N,K] (+)= dy[M,N]^T @ x[M,K]  (fp32 dw; weight gradient in the packed [out,in] layout)."""
    M, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == M and dw.shape == (N, K) and dw.dtype == torch.float32 and dy.dtype == x.dtype
    _call("spa3d_gemm_dw", _p(dy), _ld(dy), _p(x), _ld(x), dt(dy), _p(dw), _ld(dw), M, N, K, int(accumulate), int(impl), _stream())
    return dw


def gemm_strided(a, sam, sak, b, sbk, sbn, out, M, N, K, accumulate=False):
    _call("spa3d_gemm_strided", _p(a), int(sam), int(sak), dt(a), _p(b), int(sbk), int(sbn), dt(b), _p(out), _ld(out),
          dt(out), int(M), int(N), int(K), int(accumulate), _stream())
    return out


def layernorm_fwd(x, scale, out_dtype, rows=None, ldx=None, d=None, stats=False, out=None, mean=None, rstd=None):
    rows = x.shape[0] if rows is None else rows
    d = x.shape[-1] if d is None else d
    ldx = _ld(x) if ldx is None else ldx
    if out is None:
        out = torch.empty(rows, d, device=x.device, dtype=out_dtype)
    if stats and mean is None:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("spa3d_layernorm_fwd", _p(x), int(ldx), dt(x), _p(scale), _p(out), _ld(out), dt(out), _p(mean), _p(rstd), int(rows), int(d), _stream())
    return (out, mean, rstd) if stats else out


def layernorm_bwd(x, scale, mean, rstd, dy, dx, rows=None, ldx=None, lddx=None, d=None, accumulate=False, num_partials=296,
                  dx_lowp=None, dscale_accum=None, dx_colsum=None):
    """dscale_accum: the [d] fp32 gradient of the scale; when given the kernel adds into it directly and
    nothing is returned (no partial buffer, no reduction pass).  dx_colsum ([d] fp32, d <= 512): += sum over rows of the final dx."""
    rows = x.shape[0] if rows is None else rows
    d = x.shape[-1] if d is None else d
    ldx = _ld(x) if ldx is None else ldx
    lddx = _ld(dx) if lddx is None else lddx
    partial = dscale_accum if dscale_accum is not None else torch.empty(num_partials, d, device=x.device, dtype=torch.float32)
    _call("spa3d_layernorm_bwd", _p(x), int(ldx), dt(x), _p(scale), _p(mean), _p(rstd), _p(dy), _ld(dy), dt(dy), _p(dx), int(lddx),
          dt(dx), int(accumulate), _p(dx_lowp), _ld(dx_lowp) if dx_lowp is not None else 0, _p(partial), num_partials,
          int(dscale_accum is not None), _p(dx_colsum), int(rows), int(d), _stream())
    if dscale_accum is not None:
        return None
    dscale = torch.empty(d, device=x.device, dtype=torch.float32)
    colsum(partial, dscale)
    return dscale


def head_rmsnorm_fwd(buf, scale, out_mul, heads, Dh, save_rstd=False):
    rows = buf.shape[0]
    rstd = torch.empty(rows, heads, device=buf.device, dtype=torch.float32) if save_rstd else None
    _call("spa3d_head_rmsnorm_fwd", _p(buf), _ld(buf), dt(buf), _p(scale), float(out_mul), _p(rstd), heads, rows, heads, Dh, _stream())
    return rstd


def gemm_rmsnorm(a, wt, Dh, q_cols, k_cols, scale_q, scale_k, save_rstd=False, impl=GEMM_AUTO, out=None, rstd=None):
    """QKV projection with fused per-head RMSNorm (q also multiplied by 1/sqrt(Dh))."""
    M, K = a.shape
    N = wt.shape[0]
    if _is_split(a, wt):     # accurate mode on the tensor cores: contraction, then the in-place normalisation of the q / k blocks
        out = _gemm_x3(a, wt, None, ACT_NONE, None, None)
        nh = (q_cols + k_cols) // Dh
        rstd = torch.empty(M, nh, device=a.device, dtype=torch.float32) if save_rstd else None
        lib_call = lambda view, scale, mul, rs, heads: _call(
            "spa3d_head_rmsnorm_fwd", _p(view), _ld(out), dt(out), _p(scale), float(mul), _p(rs), nh, M, heads, Dh, _stream())
        if q_cols:
            lib_call(out, scale_q, 1.0 / math.sqrt(Dh), rstd, q_cols // Dh)
        if k_cols:
            lib_call(out[:, q_cols:], scale_k, 1.0, rstd[:, q_cols // Dh :] if rstd is not None else None, k_cols // Dh)
        return (out, rstd) if save_rstd else out
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=a.dtype)
    nh = (q_cols + k_cols) // Dh
    if save_rstd and rstd is None:
        rstd = torch.empty(M, nh, device=a.device, dtype=torch.float32)
    _call("spa3d_gemm_rmsnorm", _p(a), _ld(a), _p(wt), _ld(wt), dt(a), _p(out), _ld(out), dt(out), M, N, K, Dh, q_cols, k_cols,
          _p(scale_q), _p(scale_k), 1.0 / math.sqrt(Dh), _p(rstd), int(impl), _stream())
    return (out, rstd) if save_rstd else out


def head_rmsnorm_bwd(y, scale, out_mul, rstd, d_io, heads, Dh, num_partials=296, dscale_accum=None):
    """rstd: a [rows, >=heads] view (row stride = rstd.stride(0)).  dscale_accum: see layernorm_bwd."""
    rows = y.shape[0]
    partial = dscale_accum if dscale_accum is not None else torch.empty(num_partials, Dh, device=y.device, dtype=torch.float32)
    _call("spa3d_head_rmsnorm_bwd", _p(y), _ld(y), dt(y), _p(scale), float(out_mul), _p(rstd), rstd.stride(0), _p(d_io), _ld(d_io), dt(d_io),
          _p(partial), num_partials, int(dscale_accum is not None), rows, heads, Dh, _stream())
    if dscale_accum is not None:
        return None
    dscale = torch.empty(Dh, device=y.device, dtype=torch.float32)
    colsum(partial, dscale)
    return dscale


def _cross_workspace(q, batch, heads, Lq, Lk, Dh):
    """Scratch of the tcgen05 cross-attention (per-chunk partial tiles), or None when that kernel does not cover the shape."""
    lib = _lib.lib()
    if Lq == Lk or not lib.spa3d_attention_cross_applicable(dt(q), int(Lq), int(Lk), int(Dh)):
        return None
    return torch.empty(lib.spa3d_attention_cross_workspace_bytes(int(batch), int(heads), int(Lq), int(Lk), int(Dh)) // 4,
                       device=q.device, dtype=torch.float32)


def attention_fwd(q, k, v, out, batch, heads, Lq, Lk, Dh, key_mask=None, save_stats=False):
    stats = torch.empty(batch, heads, Lq, 2, device=q.device, dtype=torch.float32) if save_stats else None
    ws = _cross_workspace(q, batch, heads, Lq, Lk, Dh)
    if ws is not None:   # latents <- tracks cross-attention: key chunks across CTAs + merge
        _call("spa3d_attention_cross_fwd", _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(out), _ld(out), dt(q), _p(key_mask), _p(stats),
              _p(ws), int(batch), heads, Lq, Lk, Dh, _stream())
        return stats
    _call("spa3d_attention_fwd", _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(out), _ld(out), dt(q), _p(key_mask), _p(stats),
          int(batch), heads, Lq, Lk, Dh, _stream())
    return stats


def attention_bwd(q, k, v, o, d_o, dq, dk, dv, stats, batch, heads, Lq, Lk, Dh, key_mask=None):
    delta = torch.empty(batch * heads * Lq, device=q.device, dtype=torch.float32)
    ws = _cross_workspace(q, batch, heads, Lq, Lk, Dh)
    if ws is not None:
        _call("spa3d_attention_cross_bwd", _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(o), _ld(o), _p(d_o), _ld(d_o), _p(dq), _ld(dq),
              _p(dk), _ld(dk), _p(dv), _ld(dv), dt(q), _p(key_mask), _p(stats), _p(delta), _p(ws), int(batch), heads, Lq, Lk, Dh, _stream())
        return
    _call("spa3d_attention_bwd", _p(q), _ld(q), _p(k), _ld(k), _p(v), _ld(v), _p(o), _ld(o), _p(d_o), _ld(d_o), _p(dq), _ld(dq),
          _p(dk), _ld(dk), _p(dv), _ld(dv), dt(q), _p(key_mask), _p(stats), _p(delta), int(batch), heads, Lq, Lk, Dh, _stream())


def masked_mean_fwd(tok, visible, S, T, out_dtype):
    W = tok.shape[1]
    out = torch.empty(S, W, device=tok.device, dtype=out_dtype)
    _call("spa3d_masked_mean_fwd", _p(tok), _ld(tok), dt(tok), _p(visible), _p(out), _ld(out), dt(out), int(S), int(T), W, _stream())
    return out


def gelu_fwd(x, y):
    rows, cols = x.shape
    _call("spa3d_gelu_fwd", _p(x), _ld(x), dt(x), _p(y), _ld(y), dt(y), rows, cols, _stream())
    return y


def build_key_mask(visible, boundary_frame, has_readout=True):
    B, N, T = visible.shape[:3]
    mask = torch.empty(B, N, T + (1 if has_readout else 0), device=visible.device, dtype=torch.uint8)
    _call("spa3d_build_key_mask", _p(visible), _p(boundary_frame), _p(mask), B, N, T, int(has_readout), _stream())
    return mask


def quantize_fwd(x, noise, discretize=True, save_mask=False):
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, device=x.device, dtype=torch.uint8) if save_mask else None
    _call("spa3d_quantize_fwd", _p(x), _p(noise), _p(y), _p(mask), x.numel(), int(discretize), _stream())
    return (y, mask) if save_mask else y


def quantize_bwd(dy, pass_mask):
    dx = torch.empty_like(dy)
    _call("spa3d_quantize_bwd", _p(dy), _p(pass_mask), _p(dx), dy.numel(), _stream())
    return dx


def decoder_tokens_fwd(lat, query_emb, query_frame, tokens, B, Q, L, C):
    _call("spa3d_decoder_tokens_fwd", _p(lat), dt(lat), _p(query_emb), dt(query_emb), _p(query_frame), _p(tokens), dt(tokens), B, Q, L, C, _stream())
    return tokens


def decoder_tokens_bwd(d_tokens, query_frame, d_lat, d_qe, B, Q, L, C):
    _call("spa3d_decoder_tokens_bwd", _p(d_tokens), dt(d_tokens), _p(query_frame), _p(d_lat), _p(d_qe), B, Q, L, C, _stream())


def split_outputs(head_out, T, coords=3):
    rows = head_out.shape[0]
    tracks = torch.empty(rows, T, coords, device=head_out.device, dtype=torch.float32)
    vis = torch.empty(rows, T, 1, device=head_out.device, dtype=torch.float32)
    cert = torch.empty(rows, T, 1, device=head_out.device, dtype=torch.float32)
    _call("spa3d_split_outputs", _p(head_out), _p(tracks), _p(vis), rows, T, coords, _p(cert), _stream())
    return tracks, vis, cert


def to_tapvid3d(tracks, visible_logits, target_tracks=None):
    """[Q,T,C], [Q,T(,1)] -> ([T,Q,C] f32, [T,Q] bool occluded, [T,Q] f32 reconstruction error or None)."""
    Q, T, C = tracks.shape
    out_t = torch.empty(T, Q, C, device=tracks.device, dtype=torch.float32)
    occ = torch.empty(T, Q, device=tracks.device, dtype=torch.uint8)
    score = torch.empty(T, Q, device=tracks.device, dtype=torch.float32) if target_tracks is not None else None
    _call("spa3d_to_tapvid3d", _p(tracks), _p(visible_logits), _p(target_tracks), _p(out_t), _p(occ), _p(score), int(Q), int(T), int(C), _stream())
    return out_t, occ.bool(), score


def loss_fwd(head_out, target_tracks, target_vis, sums, T):
    _call("spa3d_loss_fwd", _p(head_out), _p(target_tracks), _p(target_vis), _p(sums), head_out.shape[0], T, _stream())
    return sums


def loss_bwd(head_out, target_tracks, target_vis, l1_w, bce_w, inv_denom, T):
    d = torch.empty_like(head_out)
    _call("spa3d_loss_bwd", _p(head_out), _p(target_tracks), _p(target_vis), _p(d), float(l1_w), float(bce_w), float(inv_denom), head_out.shape[0], T, _stream())
    return d


def gelu_bwd(pre, dy, dx):
    rows, cols = pre.shape
    _call("spa3d_gelu_bwd", _p(pre), _ld(pre), dt(pre), _p(dy), _ld(dy), dt(dy), _p(dx), _ld(dx), dt(dx), rows, cols, _stream())
    return dx


def colsum(x, out, accumulate=False):
    rows, cols = x.shape
    _call("spa3d_colsum", _p(x), _ld(x), dt(x), _p(out), int(accumulate), rows, cols, _stream())
    return out


def shadow_weights(src, dst=None, dst_t=None):
    """dst [N,K] = cdt(src), dst_t [K,N] = cdt(src)^T from the fp32 master src [N,K], one pass."""
    N, K = src.shape
    ref = dst if dst is not None else dst_t
    _call("spa3d_shadow_weights", _p(src), _ld(src), _p(dst), _ld(dst) if dst is not None else 0, _p(dst_t),
          _ld(dst_t) if dst_t is not None else 0, dt(ref), N, K, _stream())


def fill_zero(t):
    assert t.is_contiguous()
    _call("spa3d_fill_zero", _p(t), t.numel() * t.element_size(), _stream())
    return t


def stats(reset=False):
    """{counter name: count} of the library's dispatch counters (spa3d_stats)."""
    lib = _lib.lib()
    n = lib.spa3d_stats(None, 0)
    buf = (ctypes.c_int64 * n)()
    lib.spa3d_stats(ctypes.cast(buf, ctypes.c_void_p), n)
    out = {lib.spa3d_stat_name(i).decode(): int(buf[i]) for i in range(n)}
    if reset:
        lib.spa3d_stats_reset()
    return out


def axpy(y, x, alpha=1.0):
    _call("spa3d_axpy", _p(y), _p(x), float(alpha), y.numel(), _stream())
    return y


def sumsq(g, out):
    """out += sum(g*g), summed in a fixed order (identical on every data-parallel replica)."""
    ws = torch.empty(1024, device=g.device, dtype=torch.float32)
    _call("spa3d_sumsq", _p(g), g.numel(), _p(out), _p(ws), _stream())


def adamw_step(p, g, m, v, sumsq_t, clip_norm, lr, b1, b2, eps, wd, step):
    _call("spa3d_adamw_step", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(sumsq_t), float(clip_norm), float(lr), float(b1), float(b2),
          float(eps), float(wd), int(step), _stream())


def host_pack_bf16(src: torch.Tensor, dst: torch.Tensor, threads: int = 1):
    """dst (host, bfloat16) = bf16(src) (host, float32), round to nearest even, on ``threads`` host threads (csrc/host_pack.cc).
    No device work and no launch: the staging step of model.apply_stream(host_pack="bf16").  Releases the GIL while it runs."""
    if src.is_cuda or dst.is_cuda or src.dtype != torch.float32 or dst.dtype != torch.bfloat16:
        raise ValueError("host_pack_bf16: float32 host source, bfloat16 host destination")
    if not (src.is_contiguous() and dst.is_contiguous()) or src.numel() != dst.numel():
        raise ValueError("host_pack_bf16: contiguous tensors of equal size")
    _lib.check(_lib.lib().spa3d_host_pack_bf16(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), src.numel(), int(threads)),
               "spa3d_host_pack_bf16")
    return dst


def inv_sqrt(d):
    return 1.0 / math.sqrt(d)


# ---- dispatcher registration ---------------------------------------------------------------------------------
# The functions above are the ctypes-level implementations.  torch_ops defines the ``spa3d::`` operator library
# (schema + CUDA impl + fake impl + autograd formula) on top of them and rebinds the public names of this module to
# ``torch.ops.spa3d.*``, so every caller - model, training engine, tests, bench.py - goes through the dispatcher.
from . import torch_ops as _torch_ops  # noqa: E402

_torch_ops.install(globals())
