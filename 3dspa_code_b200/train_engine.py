"""Training step of the 3DSPA hot path: forward + hand-written backward kernels + AdamW, data-parallel.

What the reference specifies (its script is a sketch):
  loss        compute_loss_3d               train.py:96-129
  optimiser   clip_by_global_norm(1.0) -> adamw(lr(t), weight_decay=0.01)   train.py:239-243
  schedule    linear warm-up 0 -> base over 10k steps, cosine to 0           train.py:41-57
The reference has no data parallelism; here clips shard across ranks and ONE gradient all-reduce
(sum) per step runs over NCCL, bucketed in backward-completion order and overlapped with the rest
of the backward pass.  Cross-sample couplings kept exact: the loss normaliser
max(sum visible, 1) is taken over the GLOBAL batch, the clip uses the all-reduced gradient, and
the quantiser noise is indexed by global clip position (the caller passes each rank its slice).

Autograd is used only as a scheduler: each Function below is one residual sub-block whose
forward and backward are sequences of C-ABI kernel launches.  Weight gradients are accumulated
in place (fp32) into one flat buffer that is also the NCCL buffer; the Functions return
gradients for activations only.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import dp, ops, params as P
from .engine import DecoderContext, DeviceWeights, Engine, _as_dev

GELU = ops.ACT_GELU


def _order_names(names):
    """Backward-completion order: head, read-out transformer (last layer first), query encoder,
    decompress transformer, (de)compressor, latents transformer, track transformer, embedding."""
    def key(n):
        top = n.split(".")[0]
        rank = {"track_predictor": 0, "tra": 1, "query_encoder": 2, "dec": 3, "decompressor": 4, "compressor": 5,
                "t2l": 6, "latents_init": 7, "itt": 8, "readout_token": 9, "embed": 10}[top]
        parts = n.split(".")
        layer = -int(parts[1]) if len(parts) > 2 and parts[1].isdigit() else 1  # norm_encoder first
        return (rank, layer, n)
    return sorted(names, key=key)


class ParamStore(DeviceWeights):
    """fp32 masters, gradients and Adam moments as single flat buffers (packed layout) plus the
    compute-dtype shadows [N,K] (forward, dX of the previous layer) and [K,N] (dX)."""

    def __init__(self, tree, precision="bf16", device="cuda"):
        meta = P.tree_meta(tree)
        packed = P.pack(tree)
        cdt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.names = _order_names(packed.keys())
        self.offsets, off = {}, 0
        for n in self.names:
            self.offsets[n] = off
            off += (packed[n].size + 63) // 64 * 64
        self.total = off
        self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32)
        self.m = torch.zeros(off, device=device, dtype=torch.float32)
        self.v = torch.zeros(off, device=device, dtype=torch.float32)
        f32, g = {}, {}
        for n in self.names:
            a = packed[n]
            o = self.offsets[n]
            f32[n] = self.flat[o : o + a.size].view(a.shape)
            f32[n].copy_(torch.from_numpy(a))
            g[n] = self.grad[o : o + a.size].view(a.shape)
        super().__init__(meta, f32, {}, cdt)
        self.g = g
        self.ct: Dict[str, torch.Tensor] = {}
        self.step_count = 0
        self.progress_cb = None   # called with the final-gradient frontier (elements) whenever a backward block completes (DP overlap)
        self._pending = None      # names whose gradient is not final yet in the backward pass being tracked (None: not tracking)
        self._by_prefix: Dict[str, List[str]] = {}
        self.lowp: Dict[int, torch.Tensor] = {}   # fp32 gradient data_ptr -> bf16 copy written by the producer kernel
        self.refresh()

    # -- which gradients are final (data-parallel overlap) --------------------------------------------------
    # The flat buffers are laid out in the order the backward pass USUALLY completes them, but nothing in autograd
    # guarantees that order (its ready queue is free to run independent branches - query_encoder vs the decompress
    # transformer, latents_init vs the per-track transformer - either way round).  So completion is tracked explicitly:
    # every backward block reports the parameters it has finished, and the all-reduce frontier is the lowest offset that
    # is still pending.  An out-of-order completion only delays a bucket; it can never send one early.
    def begin_tracking(self):
        self._pending = set(self.names)

    def end_tracking(self):
        left, self._pending = self._pending, None
        return left

    def with_prefix(self, prefix):
        hit = self._by_prefix.get(prefix)
        if hit is None:
            hit = self._by_prefix[prefix] = [n for n in self.names if n.startswith(prefix)]
        return hit

    def final_below(self):
        """Offset (elements) below which every gradient in the flat buffer is final."""
        if not self._pending:
            return self.total
        return min(self.offsets[n] for n in self._pending)

    def finish(self, names):
        """The backward block that just ran has enqueued the last kernel that writes the gradients of ``names``."""
        if self._pending is None:
            return
        self._pending.difference_update(names)
        if self.progress_cb is not None:
            self.progress_cb(self.final_below())

    def is_matrix(self, k):
        return self.f32[k].dim() == 2 and (k.endswith("_t") or k.endswith(".Wt"))

    def refresh(self):
        """Compute-dtype shadows of every matrix from its fp32 master, [N,K] and [K,N] (dX = dY . W) in ONE pass per matrix."""
        for k, v in self.f32.items():
            if self.is_matrix(k):
                dst, dst_t = self.c.get(k), self.ct.get(k)
                if dst_t is None:
                    dst_t = self.ct[k] = torch.empty(v.shape[1], v.shape[0], device=v.device, dtype=self.cdt)
                if self.cdt == torch.float32:
                    self.c[k] = v
                    ops.shadow_weights(v, None, dst_t)
                else:
                    if dst is None:
                        dst = self.c[k] = torch.empty_like(v, dtype=self.cdt)
                    ops.shadow_weights(v, dst, dst_t)
        self._refresh_embed_bias()

    def accum_dw(self, name, g, x):
        """grad[name] ([N,K], fp32) += g[M,N]^T x[M,K]  (reduction over tokens)."""
        ops.gemm_dw(g, x, self.g[name], accumulate=True)

    def accum_bias(self, name, dy):
        ops.colsum(dy, self.g[name], accumulate=True)

    def tree(self):
        """Current parameters as a Flax-layout numpy tree."""
        return P.unpack({k: v.detach().cpu().numpy() for k, v in self.f32.items()}, self.meta)

    def grad_tree(self):
        return P.unpack({k: v.detach().cpu().numpy() for k, v in self.g.items()}, self.meta)


def _c(st, t):
    """activation gradient in the compute dtype.  The LayerNorm backward that produced ``t`` has
    usually written the bf16 copy in the same pass (``_lowp_out``); otherwise one conversion pass."""
    if t.dtype == st.cdt:
        return t if t.is_contiguous() else t.contiguous()
    hit = st.lowp.pop(t.data_ptr(), None)
    if hit is not None and hit.shape == t.shape:
        return hit
    return ops.convert(t, torch.empty(t.shape, device=t.device, dtype=st.cdt))


def _gelu_grad_form(st, a, wt):
    """True when MLP_in can save gelu'(z) instead of z (tcgen05 epilogue: bf16, 16-byte rows) - the backward epilogue is then
    one multiply instead of a second tanh evaluation per element."""
    return (st.cdt == torch.bfloat16 and a.shape[1] % 8 == 0 and wt.shape[0] % 8 == 0 and a.stride(0) % 8 == 0
            and wt.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0 and wt.data_ptr() % 16 == 0)


def _colsum_fusable(st, d):
    """Bias gradients that are column sums of a LayerNorm-backward output can be reduced inside that kernel (bf16 mode, fp32
    rows of at most 512 columns, or the two-warps-per-row kernel for 1024 / 1280 / 1536): the per-track transformer (d = 384),
    the latent transformer (d = 512) and the read-out transformer (d = 1280)."""
    return st.cdt == torch.bfloat16 and ((d % 128 == 0 and d <= 512) or d in (1024, 1280, 1536))


def _lowp_out(st, t):
    """bf16 side output for the fp32 gradient tensor ``t`` (consumed by the next block's ``_c``)."""
    if st.cdt == torch.float32:
        return None
    st.lowp.clear()  # at most one pending copy: the gradient handed to the next backward
    buf = torch.empty(t.shape, device=t.device, dtype=st.cdt)
    st.lowp[t.data_ptr()] = buf
    return buf


# ---- residual sub-blocks ---------------------------------------------------------------------------
_MLP_LEAVES = ("norm_attn", "W1_t", "b1", "W2_t", "b2")   # parameters of the MLP sub-block of a layer (MlpBlockFn reports them)


class AttnBlockFn(torch.autograd.Function):
    """a = x + SelfAttn(LN(x)) [+ CrossAttn(LN(x), kv)]   (attention.py:75-100)."""

    @staticmethod
    def forward(ctx, x, kv, anchor, st, pre, m, batch, L, Lkv, key_mask, prev_b2=None, bo_done=False):
        f, w, cdt = st.f32, st.c, st.cdt
        H, Dh = m["heads"], m["Dh"]
        A = H * Dh
        xn, mean, rstd = ops.layernorm_fwd(x, f[pre + "norm_q"], cdt, stats=True)
        qkv, rqk = ops.gemm_rmsnorm(xn, w[pre + "self.Wqkv_t"], Dh, A, A, f[pre + "self.norm_query"], f[pre + "self.norm_key"], save_rstd=True)
        rq, rk = rqk[:, :H], rqk[:, H:]
        o = torch.empty(x.shape[0], A, device=x.device, dtype=cdt)
        stats = ops.attention_fwd(qkv[:, :A], qkv[:, A : 2 * A], qkv[:, 2 * A :], o, batch, H, L, L, Dh, key_mask, save_stats=True)
        a = ops.gemm(o, w[pre + "self.Wo_t"], f[pre + "self.bo"], residual=x, out_dtype=torch.float32)
        ctx.saved = dict(x=x, mean=mean, rstd=rstd, xn=xn, qkv=qkv, rq=rq, rk=rk, o=o, stats=stats)
        if kv is not None:
            cq, ck = f[pre + "cross.norm_query"], f[pre + "cross.norm_key"]
            qc, rqc = ops.gemm_rmsnorm(xn, w[pre + "cross.Wq_t"], Dh, A, 0, cq, ck, save_rstd=True)
            kvp, rkc = ops.gemm_rmsnorm(kv, w[pre + "cross.Wkv_t"], Dh, 0, A, cq, ck, save_rstd=True)
            oc = torch.empty(x.shape[0], A, device=x.device, dtype=cdt)
            stc = ops.attention_fwd(qc, kvp[:, :A], kvp[:, A:], oc, batch, H, L, Lkv, Dh, None, save_stats=True)
            a = ops.gemm(oc, w[pre + "cross.Wo_t"], f[pre + "cross.bo"], residual=a, out_dtype=torch.float32)
            ctx.saved.update(kv=kv, qc=qc, kvp=kvp, rqc=rqc, rkc=rkc, oc=oc, stc=stc)
        ctx.args = (st, pre, m, batch, L, Lkv, key_mask, kv is not None and kv.requires_grad)
        ctx.fuse = (prev_b2, bo_done)
        return a

    @staticmethod
    def backward(ctx, da):
        st, pre, m, batch, L, Lkv, key_mask, kv_grad = ctx.args
        prev_b2, bo_done = ctx.fuse
        s = ctx.saved
        f, cdt = st.f32, st.cdt
        H, Dh = m["heads"], m["Dh"]
        A = H * Dh
        da = da.contiguous()
        dac = _c(st, da)
        dkv = None
        dxn = None
        if "kv" in s:
            st.accum_bias(pre + "cross.bo", dac)   # the bf16 copy: half the bytes of the fp32 gradient
            st.accum_dw(pre + "cross.Wo_t", dac, s["oc"])
            d_oc = ops.gemm(dac, st.ct[pre + "cross.Wo_t"])
            dqc = torch.empty_like(s["qc"])
            dkvp = torch.empty_like(s["kvp"])
            ops.attention_bwd(s["qc"], s["kvp"][:, :A], s["kvp"][:, A:], s["oc"], d_oc, dqc, dkvp[:, :A], dkvp[:, A:], s["stc"],
                              batch, H, L, Lkv, Dh, None)
            ops.head_rmsnorm_bwd(s["qc"], f[pre + "cross.norm_query"], 1.0 / math.sqrt(Dh), s["rqc"], dqc, H, Dh, dscale_accum=st.g[pre + "cross.norm_query"])
            ops.head_rmsnorm_bwd(s["kvp"][:, :A], f[pre + "cross.norm_key"], 1.0, s["rkc"], dkvp[:, :A], H, Dh, dscale_accum=st.g[pre + "cross.norm_key"])
            st.accum_dw(pre + "cross.Wq_t", dqc, s["xn"])
            st.accum_dw(pre + "cross.Wkv_t", dkvp, s["kv"])
            dxn = ops.gemm(dqc, st.ct[pre + "cross.Wq_t"])
            if kv_grad:
                dkv = ops.gemm(dkvp, st.ct[pre + "cross.Wkv_t"])
        if not bo_done:   # otherwise already reduced by the norm_attn backward that produced da (MlpBlockFn)
            st.accum_bias(pre + "self.bo", dac)
        st.accum_dw(pre + "self.Wo_t", dac, s["o"])
        d_o = ops.gemm(dac, st.ct[pre + "self.Wo_t"])
        qkv = s["qkv"]
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :A], qkv[:, A : 2 * A], qkv[:, 2 * A :], s["o"], d_o, dqkv[:, :A], dqkv[:, A : 2 * A], dqkv[:, 2 * A :],
                          s["stats"], batch, H, L, L, Dh, key_mask)
        ops.head_rmsnorm_bwd(qkv[:, :A], f[pre + "self.norm_query"], 1.0 / math.sqrt(Dh), s["rq"], dqkv[:, :A], H, Dh, dscale_accum=st.g[pre + "self.norm_query"])
        ops.head_rmsnorm_bwd(qkv[:, A : 2 * A], f[pre + "self.norm_key"], 1.0, s["rk"], dqkv[:, A : 2 * A], H, Dh, dscale_accum=st.g[pre + "self.norm_key"])
        st.accum_dw(pre + "self.Wqkv_t", dqkv, s["xn"])
        dxn = ops.gemm(dqkv, st.ct[pre + "self.Wqkv_t"], residual=dxn)
        # dx = da + LN'(dxn): accumulate into da's buffer (da has no other consumer)
        # ... and its column sums are the gradient of the previous layer's MLP_out bias (da = d loss / d (that layer's output))
        ops.layernorm_bwd(s["x"], f[pre + "norm_q"], s["mean"], s["rstd"], dxn, da, accumulate=True, dx_lowp=_lowp_out(st, da),
                          dscale_accum=st.g[pre + "norm_q"], dx_colsum=st.g[prev_b2] if prev_b2 else None)
        ctx.saved = None
        # final now: this layer's attention parameters (self.bo unless the MLP block's norm backward reduced it - then that block
        # reported it) and, when fused here, MLP_out's bias of the previous layer (reported by that layer's MLP block, later: safe)
        st.finish([n for n in st.with_prefix(pre) if n[len(pre):] not in _MLP_LEAVES and not (bo_done and n == pre + "self.bo")])
        return da, dkv, None, None, None, None, None, None, None, None, None, None


class MlpBlockFn(torch.autograd.Function):
    """y = a + MLP_out(gelu(MLP_in(LN(a))))   (attention.py:102-108)."""

    @staticmethod
    def forward(ctx, a, anchor, st, pre, b2_done=False, fuse_bo=False):
        f, w, cdt = st.f32, st.c, st.cdt
        an, mean, rstd = ops.layernorm_fwd(a, f[pre + "norm_attn"], cdt, stats=True)
        zg = _gelu_grad_form(st, an, w[pre + "W1_t"])
        z, h = ops.gemm_gelu(an, w[pre + "W1_t"], f[pre + "b1"], save_grad=zg)   # z is gelu'(z) when zg
        y = ops.gemm(h, w[pre + "W2_t"], f[pre + "b2"], residual=a, out_dtype=torch.float32)
        ctx.saved = dict(a=a, mean=mean, rstd=rstd, an=an, z=z, h=h, zg=zg)
        ctx.args = (st, pre)
        ctx.fuse = (b2_done, fuse_bo)
        return y

    @staticmethod
    def backward(ctx, dy):
        st, pre = ctx.args
        b2_done, fuse_bo = ctx.fuse
        s = ctx.saved
        dy = dy.contiguous()
        dyc = _c(st, dy)
        if not b2_done:   # otherwise reduced by the next layer's norm_q backward, which produced dy
            st.accum_bias(pre + "b2", dyc)
        st.accum_dw(pre + "W2_t", dyc, s["h"])
        gb1 = st.g[pre + "b1"]
        b1_fused = s["zg"] and gb1.data_ptr() % 16 == 0   # column sums of dz from the GEMM epilogue
        dz = ops.gemm_gelu_bwd(dyc, st.ct[pre + "W2_t"], s["z"], z_is_grad=s["zg"], dz_colsum=gb1 if b1_fused else None)
        if not b1_fused:
            st.accum_bias(pre + "b1", dz)
        st.accum_dw(pre + "W1_t", dz, s["an"])
        dan = ops.gemm(dz, st.ct[pre + "W1_t"])
        ops.layernorm_bwd(s["a"], st.f32[pre + "norm_attn"], s["mean"], s["rstd"], dan, dy, accumulate=True, dx_lowp=_lowp_out(st, dy),
                          dscale_accum=st.g[pre + "norm_attn"], dx_colsum=st.g[pre + "self.bo"] if fuse_bo else None)
        ctx.saved = None
        # b2 is final either way: reduced above, or by the next layer's norm_q backward, which ran before this block
        st.finish([pre + n for n in _MLP_LEAVES] + ([pre + "self.bo"] if fuse_bo else []))
        return dy, None, None, None, None, None


class LastLayerFn(torch.autograd.Function):
    """Last layer of a transformer whose output is read at token 0 only (the read-out token,
    track_autoencoder_3d.py:187,286): keys and values are needed for every token, but the query,
    the output projection, the MLP and everything downstream only for token 0 of each sequence.
    Returns y0 [batch, d] = token 0 after both residual sub-blocks; the other rows are dead code in
    the reference's graph too (XLA eliminates them), so forward and gradients are unchanged."""

    @staticmethod
    def forward(ctx, x, anchor, st, pre, m, batch, L, key_mask, prev_b2=None):
        f, w, cdt = st.f32, st.c, st.cdt
        H, Dh = m["heads"], m["Dh"]
        A, d = H * Dh, m["d"]
        sq, sk = f[pre + "self.norm_query"], f[pre + "self.norm_key"]
        wqkv = w[pre + "self.Wqkv_t"]
        xn, mean, rstd = ops.layernorm_fwd(x, f[pre + "norm_q"], cdt, stats=True)
        kvp, rk = ops.gemm_rmsnorm(xn, wqkv[A:], Dh, 0, A, sq, sk, save_rstd=True)             # keys, values: all tokens
        xn0 = xn.view(batch, L * d)[:, :d]
        q0, rq = ops.gemm_rmsnorm(xn0, wqkv[:A], Dh, A, 0, sq, sk, save_rstd=True)             # queries: token 0
        o0 = torch.empty(batch, A, device=x.device, dtype=cdt)
        stats = ops.attention_fwd(q0, kvp[:, :A], kvp[:, A:], o0, batch, H, 1, L, Dh, key_mask, save_stats=True)
        a0 = ops.gemm(o0, w[pre + "self.Wo_t"], f[pre + "self.bo"], residual=x.view(batch, L * d)[:, :d], out_dtype=torch.float32)
        an0, mean2, rstd2 = ops.layernorm_fwd(a0, f[pre + "norm_attn"], cdt, stats=True)
        zg = _gelu_grad_form(st, an0, w[pre + "W1_t"])
        z0, h0 = ops.gemm_gelu(an0, w[pre + "W1_t"], f[pre + "b1"], save_grad=zg)
        y0 = ops.gemm(h0, w[pre + "W2_t"], f[pre + "b2"], residual=a0, out_dtype=torch.float32)
        ctx.saved = dict(x=x, xn=xn, mean=mean, rstd=rstd, kvp=kvp, rk=rk, q0=q0, rq=rq, o0=o0, stats=stats, a0=a0, an0=an0,
                         mean2=mean2, rstd2=rstd2, z0=z0, h0=h0, zg=zg)
        ctx.args = (st, pre, m, batch, L, key_mask)
        ctx.prev_b2 = prev_b2
        return y0

    @staticmethod
    def backward(ctx, dy0):
        st, pre, m, batch, L, key_mask = ctx.args
        s = ctx.saved
        f, cdt = st.f32, st.cdt
        H, Dh = m["heads"], m["Dh"]
        A, d = H * Dh, m["d"]
        dy0 = dy0.contiguous()
        # ---- MLP sub-block on the token-0 rows ----
        dyc = _c(st, dy0)
        st.accum_bias(pre + "b2", dyc)
        st.accum_dw(pre + "W2_t", dyc, s["h0"])
        dz0 = ops.gemm_gelu_bwd(dyc, st.ct[pre + "W2_t"], s["z0"], z_is_grad=s["zg"])
        st.accum_bias(pre + "b1", dz0)
        st.accum_dw(pre + "W1_t", dz0, s["an0"])
        dan0 = ops.gemm(dz0, st.ct[pre + "W1_t"])
        da0 = dy0.clone()
        ops.layernorm_bwd(s["a0"], f[pre + "norm_attn"], s["mean2"], s["rstd2"], dan0, da0, accumulate=True, dscale_accum=st.g[pre + "norm_attn"])
        # ---- attention sub-block: one query row per sequence ----
        dac = _c(st, da0)
        st.accum_bias(pre + "self.bo", dac)
        st.accum_dw(pre + "self.Wo_t", dac, s["o0"])
        d_o0 = ops.gemm(dac, st.ct[pre + "self.Wo_t"])
        kvp = s["kvp"]
        dq0 = torch.empty_like(s["q0"])
        dkv = torch.empty_like(kvp)
        ops.attention_bwd(s["q0"], kvp[:, :A], kvp[:, A:], s["o0"], d_o0, dq0, dkv[:, :A], dkv[:, A:], s["stats"], batch, H, 1, L, Dh, key_mask)
        ops.head_rmsnorm_bwd(s["q0"], f[pre + "self.norm_query"], 1.0 / math.sqrt(Dh), s["rq"], dq0, H, Dh, dscale_accum=st.g[pre + "self.norm_query"])
        ops.head_rmsnorm_bwd(kvp[:, :A], f[pre + "self.norm_key"], 1.0, s["rk"], dkv[:, :A], H, Dh, dscale_accum=st.g[pre + "self.norm_key"])
        gw = st.g[pre + "self.Wqkv_t"]                      # [3A, d]: query rows, then key and value rows
        xn = s["xn"]
        xn0 = xn.view(batch, L * d)[:, :d]
        ops.gemm_dw(dq0, xn0, gw[:A], accumulate=True)
        ops.gemm_dw(dkv, xn, gw[A:], accumulate=True)
        ct = st.ct[pre + "self.Wqkv_t"]                      # [d, 3A]
        dxn = ops.gemm(dkv, ct[:, A:])                       # every token, through keys and values
        dxn0 = dxn.view(batch, L * d)[:, :d]
        ops.gemm(dq0, ct[:, :A], residual=dxn0, out=dxn0)    # token 0 also through its query
        # dx = LN'(dxn) on every token, plus the residual path, which reaches token 0 only: the norm backward writes dx
        # (no zero-filled buffer to accumulate into), then the batch x d token-0 rows are patched in dx, in its bf16 copy and
        # in the column sums handed to the previous layer's MLP_out bias
        dx = torch.empty_like(s["x"])
        low = _lowp_out(st, dx)
        ops.layernorm_bwd(s["x"], f[pre + "norm_q"], s["mean"], s["rstd"], dxn, dx, accumulate=False, dx_lowp=low,
                          dscale_accum=st.g[pre + "norm_q"], dx_colsum=st.g[ctx.prev_b2] if ctx.prev_b2 else None)
        dx0 = dx.view(batch, L * d)[:, :d]
        dx0.add_(da0)
        if low is not None:
            low.view(batch, L * d)[:, :d].copy_(dx0)
        if ctx.prev_b2:
            ops.colsum(da0, st.g[ctx.prev_b2], accumulate=True)
        ctx.saved = None
        st.finish(st.with_prefix(pre))
        return dx, None, None, None, None, None, None, None, None


class FinalNormFn(torch.autograd.Function):
    """norm_encoder (attention.py:49) on every token or only on token 0 of each sequence."""

    @staticmethod
    def forward(ctx, x, anchor, st, name, batch, L, first):
        d = x.shape[1]
        if first:
            y, mean, rstd = ops.layernorm_fwd(x, st.f32[name], st.cdt, rows=batch, ldx=L * d, d=d, stats=True)
        else:
            y, mean, rstd = ops.layernorm_fwd(x, st.f32[name], st.cdt, stats=True)
        ctx.saved = (x, mean, rstd)
        ctx.args = (st, name, batch, L, first)
        return y

    @staticmethod
    def backward(ctx, dy):
        st, name, batch, L, first = ctx.args
        x, mean, rstd = ctx.saved
        d = x.shape[1]
        dy = dy.contiguous()
        if first:
            dx = torch.zeros_like(x)
            ds = ops.layernorm_bwd(x, st.f32[name], mean, rstd, dy, dx, rows=batch, ldx=L * d, lddx=L * d, d=d)
        else:
            dx = torch.empty_like(x)
            ds = ops.layernorm_bwd(x, st.f32[name], mean, rstd, dy, dx)
        ops.axpy(st.g[name], ds)
        ctx.saved = None
        st.finish([name])
        return dx, None, None, None, None, None, None


class EmbedFn(torch.autograd.Function):
    """tokens = [readout ; A_cat . W_embed + b]   (track_autoencoder_3d.py:123-165); A_cat is data."""

    @staticmethod
    def forward(ctx, anchor, st, a_cat, wt, bias, bias_names, seqs, T, ro, x=None):
        if x is None:   # a_cat was built by the unfused kernels; otherwise the fused K1 kernel produced x and a_cat together
            x = ops.gemm(a_cat, wt, bias, out_dtype=torch.float32)
        if ro:
            ops.set_rows(x, T + 1, st.f32["readout_token"].view(-1), seqs)
        ctx.saved = a_cat
        ctx.args = (st, bias_names, seqs, T, ro)
        return x

    @staticmethod
    def backward(ctx, dx):
        st, bias_names, seqs, T, ro = ctx.args
        a_cat = ctx.saved
        dx = dx.contiguous()
        W = dx.shape[1]
        if ro:
            # d readout_token = sum over sequences of row 0; those rows do not reach the GEMM
            tok = torch.empty(W, device=dx.device, dtype=torch.float32)
            ops.colsum(dx.view(seqs, (T + 1) * W)[:, :W], tok)
            ops.axpy(st.g["readout_token"].view(-1), tok)
            ops.set_rows(dx, T + 1, torch.zeros(W, device=dx.device), seqs)
        db = torch.empty(W, device=dx.device, dtype=torch.float32)
        ops.colsum(dx, db)
        for n in bias_names:
            ops.axpy(st.g[n], db)
        # a feature the tree has a projection for but the batch lacks (track_autoencoder_3d.py:140,145 skip it): its columns of
        # a_cat are zero, so its kernel receives a zero gradient from the same GEMM, and its bias is not in bias_names
        st.accum_dw("embed.Wt", _c(st, dx), a_cat)
        ctx.saved = None
        st.finish(st.with_prefix("embed.") + ["readout_token"])
        return (None,) * 10


class LatentInitFn(torch.autograd.Function):
    """ParamStateInit broadcast over the batch (track_autoencoder.py:41-53)."""

    @staticmethod
    def forward(ctx, anchor, st, B):
        li = st.f32["latents_init"]
        ctx.args = (st, B)
        return li.unsqueeze(0).expand(B, *li.shape).reshape(B * li.shape[0], li.shape[1]).contiguous()

    @staticmethod
    def backward(ctx, d):
        st, B = ctx.args
        g = st.g["latents_init"]
        ops.colsum(d.contiguous().view(B, g.numel()), g.view(-1), accumulate=True)
        st.finish(["latents_init"])
        return None, None, None


class DenseFn(torch.autograd.Function):
    """y = x . W + b   (compressor / decompressor / query_encoder / track_predictor)."""

    @staticmethod
    def forward(ctx, x, anchor, st, name, out_dtype):
        y = ops.gemm(x, st.c[name + ".Wt"], st.f32[name + ".b"], out_dtype=out_dtype)
        ctx.saved = x
        ctx.args = (st, name, x.requires_grad)
        return y

    @staticmethod
    def backward(ctx, dy):
        st, name, need_dx = ctx.args
        x = ctx.saved
        dy = dy.contiguous()
        dyc = _c(st, dy)
        st.accum_bias(name + ".b", dy)
        st.accum_dw(name + ".Wt", dyc, x)
        dx = ops.gemm(dyc, st.ct[name + ".Wt"], out_dtype=x.dtype) if need_dx else None
        ctx.saved = None
        st.finish([name + ".Wt", name + ".b"])
        return dx, None, None, None, None


class QuantFn(torch.autograd.Function):
    """clip + round + noise with the straight-through gradient (track_autoencoder_3d.py:251-260)."""

    @staticmethod
    def forward(ctx, z, anchor, st, noise, discretize):
        y, mask = ops.quantize_fwd(z, noise, discretize, save_mask=True)
        ctx.saved = mask
        ctx.cdt = st.cdt
        return y if st.cdt == torch.float32 else ops.convert(y, torch.empty_like(y, dtype=st.cdt))

    @staticmethod
    def backward(ctx, dy):
        d = dy.contiguous()
        if d.dtype != torch.float32:
            d = ops.convert(d, torch.empty(d.shape, device=d.device, dtype=torch.float32))
        return ops.quantize_bwd(d, ctx.saved), None, None, None, None


class TokensFn(torch.autograd.Function):
    """decoder token assembly (track_autoencoder_3d.py:276-284)."""

    @staticmethod
    def forward(ctx, lat, qe, anchor, st, qframe, B, Q, L, C):
        tokens = torch.empty(B * Q * (L + 1), C + 128, device=lat.device, dtype=torch.float32)
        ops.decoder_tokens_fwd(lat, qe, qframe, tokens, B, Q, L, C)
        ctx.args = (st, qframe, B, Q, L, C, lat.dtype)
        return tokens

    @staticmethod
    def backward(ctx, d):
        st, qframe, B, Q, L, C, lat_dtype = ctx.args
        d = d.contiguous()
        d_lat = torch.empty(B * L, C, device=d.device, dtype=torch.float32)
        d_qe = torch.empty(B * Q, C + 128, device=d.device, dtype=torch.float32)
        ops.decoder_tokens_bwd(d, qframe, d_lat, d_qe, B, Q, L, C)
        if lat_dtype != torch.float32:
            d_lat = ops.convert(d_lat, torch.empty_like(d_lat, dtype=lat_dtype))
        return d_lat, d_qe, None, None, None, None, None, None, None


# ---- training executor --------------------------------------------------------------------------------
class TrainEngine(Engine):
    """Forward with saved activations + backward + optimiser on a ParamStore."""

    def __init__(self, cfg, store: ParamStore):
        super().__init__(cfg, store)
        self.st = store
        self.prune_last = True   # token-0-only last layer of the read-out transformers (same dead-code elimination as inference)
        self.anchor = torch.zeros((), device=self.dev, requires_grad=True)

    def _transformer_t(self, short, x, batch, L, key_mask=None, kv=None, Lkv=0, first=False):
        m = self.w.meta[short]
        for i in range(m["layers"]):
            pre = f"{short}.{i}."
            if first and self.prune_last and i == m["layers"] - 1 and kv is None and L > 1:
                prev_b2 = f"{short}.{i - 1}.b2" if (_colsum_fusable(self.st, m["d"]) and i > 0) else None
                y0 = LastLayerFn.apply(x, self.anchor, self.st, pre, m, batch, L, key_mask, prev_b2)
                return FinalNormFn.apply(y0, self.anchor, self.st, f"{short}.norm_encoder", batch, 1, False)
            # bias gradients reduced inside the LayerNorm backward that produces the gradient they are column sums of:
            # this layer's attention block hands MLP_out's bias of the PREVIOUS layer to its norm_q backward, this layer's
            # MLP block hands the attention out-projection bias of the SAME layer to its norm_attn backward
            fuse = _colsum_fusable(self.st, m["d"])
            prev_b2 = f"{short}.{i - 1}.b2" if (fuse and i > 0) else None
            nxt_does_b2 = fuse and i + 1 < m["layers"]
            a = AttnBlockFn.apply(x, kv, self.anchor, self.st, pre, m, batch, L, Lkv, key_mask, prev_b2, fuse)
            x = MlpBlockFn.apply(a, self.anchor, self.st, pre, nxt_does_b2, fuse)
        return FinalNormFn.apply(x, self.anchor, self.st, f"{short}.norm_encoder", batch, L, first)

    def forward_train(self, inputs, noise, discretize=True):
        """Returns head_out [B*Q, 4T] (fp32) attached to the autograd graph."""
        cfg, meta, dev, st = self.cfg, self.w.meta, self.dev, self.st
        tracks = _as_dev(inputs["support_tracks"], torch.float32, dev)
        visible = _as_dev(inputs["support_tracks_visible"], torch.float32, dev)
        boundary = _as_dev(inputs["boundary_frame"], torch.int32, dev)
        dino = depth = None
        if meta["has_dino"] and cfg.use_dino and inputs.get("dino_features") is not None:
            dino = _as_dev(inputs["dino_features"], torch.float32, dev)
        if meta["has_depth"] and cfg.use_depth and inputs.get("depth_features") is not None:
            depth = _as_dev(inputs["depth_features"], torch.float32, dev)
        B, N, T, C3 = tracks.shape
        rows = B * N * T
        K = st.c["embed.Wt"].shape[1]
        a_cat = torch.empty(B * N * (T + 1), K, device=dev, dtype=self.cdt)
        a_cat.view(B * N, T + 1, K)[:, 0].zero_()
        bias_names = ["embed.b_track"] + (["embed.b_dino"] if dino is not None else []) + (["embed.b_depth"] if depth is not None else [])
        missing = (meta["has_dino"] and dino is None) or (meta["has_depth"] and depth is None)
        embed_bias = st.embed_bias
        if missing:
            # the tree has a projection the batch has no feature for: that Dense is skipped (:140,145) - zero feature columns
            # (so the concatenated GEMM adds nothing and its kernel gets a zero gradient) and no bias
            embed_bias = st.f32["embed.b_track"].clone()
            for n in bias_names[1:]:
                ops.axpy(embed_bias, st.f32[n])
            off = meta["fourier_in"]
            if meta["has_dino"]:
                if dino is None:
                    a_cat[:, off : off + meta["dino_dim"]].zero_()
                off += meta["dino_dim"]
            if meta["has_depth"] and depth is None:
                a_cat[:, off : off + meta["depth_dim"]].zero_()
        x0 = None
        W = st.c["embed.Wt"].shape[0]
        if (not missing and self.cdt == torch.bfloat16 and self.fused_embed and cfg.num_frequencies == 32 and C3 == 3
                and ops.embed_fused_applicable(W, K, dino.shape[-1] if dino is not None else 0, depth.shape[-1] if depth is not None else 0, C3)):
            # K1: one kernel produces the tokens AND the bf16 concatenated features the weight gradient needs
            x0 = torch.empty(B * N * (T + 1), W, device=dev, dtype=torch.float32)
            ops.embed_fused(tracks.view(rows, C3), dino.view(rows, -1) if dino is not None else None,
                            depth.view(rows, -1) if depth is not None else None, st.c["embed.Wt"], st.embed_bias, x0, T,
                            cfg.num_frequencies, cfg.track_scale_factor, a_cat=a_cat)
        else:
            ops.fourier_features(tracks.view(rows, C3), a_cat, cfg.num_frequencies, cfg.track_scale_factor, append_time=T,
                                 exact=self.exact, out_row_group=T)
            off = meta["fourier_in"]
            if dino is not None:
                ops.convert(dino.view(rows, -1), a_cat[:, off : off + meta["dino_dim"]], out_row_group=T)
            if meta["has_dino"]:
                off += meta["dino_dim"]
            if depth is not None:
                ops.convert(depth.view(rows, -1), a_cat[:, off : off + meta["depth_dim"]], out_row_group=T)
        x = EmbedFn.apply(self.anchor, st, a_cat, st.c["embed.Wt"], embed_bias, bias_names, B * N, T, True, x0)
        key_mask = ops.build_key_mask(visible, boundary, True)
        stok = self._transformer_t("itt", x, B * N, T + 1, key_mask, first=True)
        nl = meta["latent_tokens"]
        lat = LatentInitFn.apply(self.anchor, st, B)
        lat = self._transformer_t("t2l", lat, B, nl, kv=stok, Lkv=N)
        z = DenseFn.apply(lat, self.anchor, st, "compressor", torch.float32)
        noise_d = _as_dev(noise, torch.float32, dev).reshape(B * nl, -1) if discretize else None
        zq = QuantFn.apply(z, self.anchor, st, noise_d, discretize)
        xdec = DenseFn.apply(zq, self.anchor, st, "decompressor", torch.float32)
        latd = self._transformer_t("dec", xdec, B, nl)
        ctx = self.get_decoder_context(inputs)
        Q = ctx.query_frame.shape[1]
        qfeat = torch.empty(B * Q, meta["query_in"], device=dev, dtype=self.cdt)
        # tail_zero assumes query_frame // time_scale_factor == 0 (:268-269); Engine.decode raises in the same case
        lo_hi = torch.stack([ctx.query_frame.min(), ctx.query_frame.max()]).tolist()
        if lo_hi[0] < 0 or lo_hi[1] >= cfg.time_scale_factor:
            raise ValueError("query frames outside [0, time_scale_factor) are not supported")
        ops.fourier_features(ctx.decoder_query.reshape(B * Q, -1), qfeat, cfg.num_frequencies, cfg.track_scale_factor,
                             tail_zero=True, exact=self.exact)
        qe = DenseFn.apply(qfeat, self.anchor, st, "query_encoder", torch.float32)
        tokens = TokensFn.apply(latd, qe, self.anchor, st, ctx.query_frame, B, Q, nl, meta["D"] - 128)
        out = self._transformer_t("tra", tokens, B * Q, nl + 1, first=True)
        return DenseFn.apply(out, self.anchor, st, "track_predictor", torch.float32)

    def loss_and_backward(self, inputs, noise, denom, l1_weight=5000.0, bce_weight=1e-8, discretize=True, sums=None):
        """One micro-batch: forward, loss sums (position, bce, visible count), backward into st.grad.
        ``denom`` = max(global visible count, 1) (train.py:111-113 normalises over the whole batch)."""
        dev = self.dev
        T = self.cfg.num_output_frames
        self.st.lowp.clear()   # no bf16 gradient copy of an earlier backward may be matched by address
        with torch.enable_grad():
            head = self.forward_train(inputs, noise, discretize)
        tt = _as_dev(inputs["query_tracks"], torch.float32, dev).reshape(-1, T, 3)
        tv = _as_dev(inputs["query_tracks_visible"], torch.float32, dev).reshape(-1, T, 1)
        if sums is None:
            sums = torch.zeros(3, device=dev)
        ops.loss_fwd(head.detach(), tt, tv, sums, T)
        d_head = ops.loss_bwd(head.detach(), tt, tv, l1_weight, bce_weight, 1.0 / denom, T)
        head.backward(d_head)
        return sums


def learning_rate(step, base_lr=1e-4, warmup_steps=10000, total_steps=1000000):
    """create_learning_rate_schedule (train.py:41-57); ``step`` = number of updates already applied."""
    if step < warmup_steps:
        return base_lr * step / warmup_steps
    t = min(step - warmup_steps, total_steps - warmup_steps) / (total_steps - warmup_steps)
    return base_lr * 0.5 * (1.0 + math.cos(math.pi * t))


class Trainer:
    """Data-parallel training step over the clips of a global batch.

    Every rank holds the full parameters and optimiser state; rank r processes its slice of the
    global batch in micro-batches, gradients accumulate in ``store.grad`` and are summed across
    ranks by a bucketed NCCL all-reduce issued from a side stream, then the identical
    clip + AdamW update runs on every rank.
    """

    def __init__(self, model, tree, precision="bf16", device="cuda", base_lr=1e-4, warmup_steps=10000, total_steps=1000000,
                 weight_decay=0.01, clip_norm=1.0, micro_batch=2, bucket_mb=64, group=None, l1_weight=5000.0, bce_weight=1e-8):
        self.model = model
        self.store = ParamStore(tree, precision, device)
        self.engine = TrainEngine(model, self.store)
        self.hp = dict(base_lr=base_lr, warmup_steps=warmup_steps, total_steps=total_steps)
        self.wd, self.clip, self.micro = weight_decay, clip_norm, micro_batch
        self.l1_weight, self.bce_weight = l1_weight, bce_weight   # compute_loss_3d weights (train.py:96-129)
        self.unfinished = []
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.bucket_elems = bucket_mb * (1 << 20) // 4
        self.comm_stream = torch.cuda.Stream(device=device) if self.world > 1 and torch.device(device).type == "cuda" else None
        self.step_idx = 0
        self.overlap = True
        self._sent = 0
        self.debug_frontier = None   # set to [] to record (lo, hi, snapshot) of every region declared final (tests)

    # -- gradient all-reduce -----------------------------------------------------------------------------
    def _reduce_upto(self, frontier):
        """All-reduce (sum) the complete buckets below ``frontier`` on the side stream; they were
        finished by kernels already enqueued on the compute stream."""
        b = self.bucket_elems
        hi = self.store.total if frontier >= self.store.total else (frontier // b) * b
        if hi <= self._sent:
            return
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            dp.bucketed_allreduce(self.store.grad[self._sent : hi], self.bucket_elems, self.group)
        self._sent = hi

    def _progress(self, frontier):
        if self.debug_frontier is not None:
            # test hook: keep a copy of everything declared final; nothing may write there afterwards
            lo = self.debug_frontier[-1][1] if self.debug_frontier else 0
            if frontier > lo:
                self.debug_frontier.append((lo, frontier, self.store.grad[lo:frontier].clone()))
        if self.world > 1 and self.overlap:
            self._reduce_upto(frontier)

    def _allreduce_rest(self):
        if self.world == 1:
            return
        self._reduce_upto(self.store.total)
        torch.cuda.current_stream().wait_stream(self.comm_stream)

    def train_step(self, batch, noise):
        """batch: this rank's clips (dict of arrays with leading axis B_local, incl. query_tracks /
        query_tracks_visible targets); noise: [B_local,128,96] slice of the global noise tensor.
        Returns dict(total_loss, position_loss, visible_loss, learning_rate, grad_norm)."""
        st, dev = self.store, self.engine.dev
        ops.fill_zero(st.grad)
        Bl = batch["support_tracks"].shape[0]
        tv = _as_dev(batch["query_tracks_visible"], torch.float32, dev)
        cnt = torch.zeros(1, device=dev)
        ops.colsum(tv.reshape(-1, 1), cnt)  # local visible count; summed over ranks below
        denom = dp.global_denominator(cnt, self.group)
        sums = torch.zeros(3, device=dev)
        self._sent = 0
        for s in range(0, Bl, self.micro):
            mb = {k: (v[s : s + self.micro] if hasattr(v, "shape") and len(v.shape) > 0 and v.shape[0] == Bl else v) for k, v in batch.items()}
            # gradients are final only in the last micro-batch: overlap the all-reduce with ITS backward,
            # bucket by bucket, as the layers complete (the flat buffer is in backward-completion order)
            last = s + self.micro >= Bl
            track = last and ((self.world > 1 and self.overlap) or self.debug_frontier is not None)
            if track:
                st.progress_cb = self._progress
                st.begin_tracking()
            self.engine.loss_and_backward(mb, noise[s : s + self.micro], denom, l1_weight=self.l1_weight, bce_weight=self.bce_weight,
                                          sums=sums)
            if track:
                self.unfinished = sorted(st.end_tracking())   # parameters no backward block reported (reduced with the rest below)
        st.progress_cb = None
        self._allreduce_rest()
        if self.world > 1:
            dist.all_reduce(sums, group=self.group)
        lr = learning_rate(self.step_idx, **self.hp)
        ss = torch.zeros(1, device=dev)
        ops.sumsq(st.grad, ss)
        self.step_idx += 1
        ops.adamw_step(st.flat, st.grad, st.m, st.v, ss, self.clip, lr, 0.9, 0.999, 1e-8, self.wd, self.step_idx)
        st.refresh()
        s_host = sums.cpu()
        pos, bce = float(s_host[0]) / denom, float(s_host[1]) / denom
        return {"total_loss": self.l1_weight * pos + self.bce_weight * bce, "position_loss": pos, "visible_loss": bce, "learning_rate": lr,
                "grad_norm": float(ss.sqrt().item())}
