"""Forward pass of the 3D semantic point-track autoencoder on the B200 kernels.

Mirrors TrackAutoEncoder3D.encode / get_decoder_context / decode
(/root/reference/track_autoencoder_3d.py:123-307) and the transformer blocks of attention.py,
as a straight-line sequence of C-ABI kernel launches (ops.py).  Also runs TRAJAN 2D
(track_autoencoder.py:205-345) through the same kernels.

Precision modes
  "bf16"  activations/weights bf16 into tcgen05 GEMMs and mma attention, fp32 accumulate,
          fp32 residual stream and outputs            (north_star tolerance 2e-2)
  "fp32"  everything float32 on the SIMT kernels        (north_star tolerance 1e-4)

Data layout in HBM: every activation is a row-major [tokens, width] matrix; a "sequence" is a
run of consecutive rows (T+1 = 151 rows per support track, 129 per query, 128 latents per clip).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import ops, params as P, rng


@dataclass
class DeviceWeights:
    """Packed parameters resident on the GPU: fp32 masters + compute-dtype shadows."""

    meta: dict
    f32: Dict[str, torch.Tensor]   # packed layout, float32 (norm scales, biases, masters)
    c: Dict[str, torch.Tensor]     # matrices in the compute dtype ([out, in], K contiguous)
    cdt: torch.dtype
    embed_bias: torch.Tensor = None
    x3: bool = True                # fp32 mode: contractions on the tensor cores as "bf16 x 3" split products where the shape allows
    c3: Dict[str, torch.Tensor] = None   # fp32 mode: [out, 3*in] bf16 splits (hi | mid | lo) of the matrices ops.gemm_x3 can take

    def mats(self):
        """Weight lookup for the forward GEMMs: the compute-dtype matrix, or - fp32 mode - its three-term bf16 split."""
        if not self.c3:
            return self.c
        return _Prefer(self.c3, self.c)

    @staticmethod
    def from_tree(tree, precision="bf16", device="cuda"):
        meta = P.tree_meta(tree)
        packed = P.pack(tree)
        cdt = torch.bfloat16 if precision == "bf16" else torch.float32
        f32 = {k: torch.from_numpy(v).to(device) for k, v in packed.items()}
        w = DeviceWeights(meta, f32, {}, cdt)
        w.refresh()
        return w

    def refresh(self):
        """(Re)derive compute-dtype shadows from the fp32 masters (after an optimiser step)."""
        if self.c3 is None:
            self.c3 = {}
        for k, v in self.f32.items():
            if v.dim() == 2 and (k.endswith("_t") or k.endswith(".Wt")):
                if self.cdt == torch.float32:
                    self.c[k] = v
                    if self.x3 and ops.gemm_x3_applicable(1, v.shape[0], v.shape[1]):
                        self.c3[k] = ops.split3(v, self.c3.get(k))
                else:
                    dst = self.c.get(k)
                    if dst is None:
                        dst = torch.empty_like(v, dtype=self.cdt)
                        self.c[k] = dst
                    ops.convert(v, dst)
        self._refresh_embed_bias()

    def _refresh_embed_bias(self):
        """Sum of the three embedding biases (the concatenated-K GEMM adds them once)."""
        eb = self.f32["embed.b_track"].clone()
        for k in ("embed.b_dino", "embed.b_depth"):
            if k in self.f32:
                ops.axpy(eb, self.f32[k], 1.0)
        self.embed_bias = eb


class _Prefer:
    """Read-only mapping: ``first[k]`` when present, else ``second[k]``."""

    def __init__(self, first, second):
        self.first, self.second = first, second

    def __getitem__(self, k):
        hit = self.first.get(k)
        return hit if hit is not None else self.second[k]


def _feat_dev(x, device):
    """Per-track DINO / depth features on the device: float32 as in the reference's batch, or bfloat16 when the caller
    already holds them in bfloat16 (a torch tensor; halves the host->device bytes - the bf16 path rounds them to bf16 anyway)."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.bfloat16:
        return x.to(device=device, non_blocking=True).contiguous()
    return _as_dev(x, torch.float32, device)


def _as_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=device, dtype=dtype, non_blocking=True)


class Engine:
    """Inference-time executor (no autograd).  Training uses autograd_fns.TrainEngine."""

    def __init__(self, cfg, weights: DeviceWeights, exact_sin: Optional[bool] = None):
        self.cfg = cfg
        self.w = weights
        self.cdt = weights.cdt
        self.exact = (weights.cdt == torch.float32) if exact_sin is None else exact_sin
        self.fused_embed = True     # bf16 path: fused Fourier + feature conversion + first projection (K1)
        # bf16 path, d = 384: MLP_in -> GELU -> MLP_out as one kernel (spa3d_mlp_fused: the hidden activation stays on the SM).  Correct
        # (tests/test_gpu_kernels.py) but OFF by default: measured 0.805 ms against 0.783 ms for the two GEMMs on the per-track shape -
        # with the 96 KB token tile resident, shared memory leaves a 96 KB weight ring, about one L2 latency of look-ahead
        self.fused_mlp = False
        # dtype of the residual stream INSIDE the transformer stacks.  float32 (default) is the reference's; bfloat16 halves the bytes the
        # residual GEMMs and the LayerNorms move (they are HBM-bound at 8 bytes per output element) at the price of one more
        # bf16 rounding per residual add - an inference-only option (model.residual_dtype), measured in DESIGN 5, off by default
        self.sdt = torch.float32
        self.stream_chunk = 256     # support tracks per host->device pipeline stage (0 = one-shot upload)
        self._copy_stream = None
        self._noise_cache = None    # ((B, tokens, dim, layout), device tensor) of the default quantiser noise
        self._stage = {}            # (input key, slot) -> persistent device staging buffer
        self.dev = weights.f32["latents_init"].device

    # ---- transformer blocks (attention.py:11-185) ---------------------------------------------
    def transformer(self, short, x, batch, L, key_mask=None, kv=None, Lkv=0, out_rows="all"):
        """ImprovedTransformer (attention.py:22-53) on x [batch*L, d] (fp32 residual stream).

        kv: optional [batch*Lkv, d_kv] cross-attention inputs (compute dtype, un-normed).
        out_rows: "all" -> final LayerNorm of every token; "first" -> only token 0 of every sequence
        (the read-out token, track_autoencoder_3d.py:187,286).  With "first" nothing downstream
        reads tokens 1.. of the last layer, so that layer computes keys/values for every token but
        queries, output projection, MLP and final norm only for token 0 (dead-code elimination;
        bench.py reports executed FLOPs from the launches, not from the model formula).
        """
        m = self.w.meta[short]
        H, Dh = m["heads"], m["Dh"]
        A = H * Dh
        d = m["d"]
        w, f = self.w.mats(), self.w.f32
        sdt = self.sdt if self.cdt == torch.bfloat16 else torch.float32
        for i in range(m["layers"]):
            pre = f"{short}.{i}."
            prune = out_rows == "first" and i == m["layers"] - 1 and kv is None and L > 1
            xn = ops.layernorm_fwd(x, f[pre + "norm_q"], self.cdt)
            sq, sk = f[pre + "self.norm_query"], f[pre + "self.norm_key"]
            if prune:
                wqkv = w[pre + "self.Wqkv_t"]
                kvp = ops.gemm_rmsnorm(xn, wqkv[A:], Dh, 0, A, sq, sk)                      # keys, values: all tokens
                q0 = ops.gemm_rmsnorm(xn.view(batch, L * d)[:, :d], wqkv[:A], Dh, A, 0, sq, sk)  # queries: token 0
                o = torch.empty(batch, A, device=self.dev, dtype=self.cdt)
                ops.attention_fwd(q0, kvp[:, :A], kvp[:, A:], o, batch, H, 1, L, Dh, key_mask)
                x = ops.gemm(o, w[pre + "self.Wo_t"], f[pre + "self.bo"], residual=x.view(batch, L * d)[:, :d], out_dtype=sdt)
                del kvp, q0, o
                an = ops.layernorm_fwd(x, f[pre + "norm_attn"], self.cdt)
                h = ops.gemm(an, w[pre + "W1_t"], f[pre + "b1"], act=ops.ACT_GELU)
                x = ops.gemm(h, w[pre + "W2_t"], f[pre + "b2"], residual=x, out_dtype=sdt)
                return ops.layernorm_fwd(x, f[f"{short}.norm_encoder"], self.cdt)
            qkv = ops.gemm_rmsnorm(xn, w[pre + "self.Wqkv_t"], Dh, A, A, sq, sk)
            o = torch.empty(x.shape[0], A, device=self.dev, dtype=self.cdt)
            ops.attention_fwd(qkv[:, :A], qkv[:, A : 2 * A], qkv[:, 2 * A :], o, batch, H, L, L, Dh, key_mask)
            a = ops.gemm(o, w[pre + "self.Wo_t"], f[pre + "self.bo"], residual=x, out_dtype=sdt)
            del qkv, o
            if kv is not None:
                cq, ck = f[pre + "cross.norm_query"], f[pre + "cross.norm_key"]
                qc = ops.gemm_rmsnorm(xn, w[pre + "cross.Wq_t"], Dh, A, 0, cq, ck)
                kvp = ops.gemm_rmsnorm(kv, w[pre + "cross.Wkv_t"], Dh, 0, A, cq, ck)
                oc = torch.empty(x.shape[0], A, device=self.dev, dtype=self.cdt)
                ops.attention_fwd(qc, kvp[:, :A], kvp[:, A:], oc, batch, H, L, Lkv, Dh, None)
                a = ops.gemm(oc, w[pre + "cross.Wo_t"], f[pre + "cross.bo"], residual=a, out_dtype=sdt)
                del qc, kvp, oc
            an = ops.layernorm_fwd(a, f[pre + "norm_attn"], self.cdt)
            if (self.fused_mlp and self.cdt == torch.bfloat16 and a.shape[0] >= 4096
                    and ops.mlp_fused_applicable(d, w[pre + "W1_t"].shape[0])):
                # MLP_in -> GELU -> MLP_out in one kernel: the hidden activation never reaches HBM (per-track transformer, d = 384)
                x = ops.mlp_fused(an, w[pre + "W1_t"], f[pre + "b1"], w[pre + "W2_t"], f[pre + "b2"], a)
                del xn, an, a
                continue
            h = ops.gemm(an, w[pre + "W1_t"], f[pre + "b1"], act=ops.ACT_GELU)
            x = ops.gemm(h, w[pre + "W2_t"], f[pre + "b2"], residual=a, out_dtype=sdt)
            del xn, an, h, a
        if out_rows == "first":
            return ops.layernorm_fwd(x, f[f"{short}.norm_encoder"], self.cdt, rows=batch, ldx=L * d, d=d)
        return ops.layernorm_fwd(x, f[f"{short}.norm_encoder"], self.cdt)

    # ---- encoder ---------------------------------------------------------------------------------
    def embed_tracks(self, tracks, dino, depth, readout):
        """embed_track_pos_visible (+ read-out slot): returns the fp32 token matrix
        [B*N*(T+ro), W] (track_autoencoder_3d.py:123-165)."""
        cfg, meta = self.cfg, self.w.meta
        B, N, T, C = tracks.shape
        ro = 1 if readout else 0
        rows = B * N * T
        K = meta["fourier_in"] + (meta["dino_dim"] if dino is not None else 0) + (meta["depth_dim"] if depth is not None else 0)
        wt = self.w.c["embed.Wt"]
        bias = self.w.embed_bias
        if K != wt.shape[1]:
            # a feature the checkpoint has a projection for was not supplied (:140,145): use the
            # matching column block of the concatenated kernel and only the biases in play
            cols = list(range(meta["fourier_in"]))
            off = meta["fourier_in"]
            bias = self.w.f32["embed.b_track"].clone()
            if meta["has_dino"]:
                if dino is not None:
                    cols += list(range(off, off + meta["dino_dim"]))
                    ops.axpy(bias, self.w.f32["embed.b_dino"])
                off += meta["dino_dim"]
            if meta["has_depth"] and depth is not None:
                cols += list(range(off, off + meta["depth_dim"]))
                ops.axpy(bias, self.w.f32["embed.b_depth"])
            wt = wt[:, cols].contiguous()
        lowp_in = any(t is not None and t.dtype != torch.float32 for t in (dino, depth))   # bf16 features: converted / copied, then one GEMM
        if lowp_in and self.cdt != torch.bfloat16:
            raise ValueError("bfloat16 features need precision='bf16'")
        if (ro and not lowp_in and self.cdt == torch.bfloat16 and self.fused_embed and cfg.num_frequencies == 32 and C == 3
                and ops.embed_fused_applicable(wt.shape[0], K, dino.shape[-1] if dino is not None else 0,
                                               depth.shape[-1] if depth is not None else 0, C)):
            # K1: Fourier features + fp32->bf16 feature conversion inside the GEMM's operand producers
            x = torch.empty(B * N * (T + 1), wt.shape[0], device=self.dev, dtype=torch.float32)
            ops.embed_fused(tracks.view(rows, C), dino.view(rows, -1) if dino is not None else None,
                            depth.view(rows, -1) if depth is not None else None, wt, bias, x, T, cfg.num_frequencies,
                            cfg.track_scale_factor)
            ops.set_rows(x, T + 1, self.w.f32["readout_token"].view(-1), B * N)
            return x
        grp = T if ro else 0
        a = torch.empty(B * N * (T + ro), K, device=self.dev, dtype=self.cdt)
        if ro:
            a.view(B * N, T + 1, K)[:, 0].zero_()
        ops.fourier_features(tracks.view(rows, C), a, cfg.num_frequencies, cfg.track_scale_factor, append_time=T,
                             exact=self.exact, out_row_group=grp)
        off = meta["fourier_in"]
        if dino is not None:
            ops.convert(dino.view(rows, -1), a[:, off : off + meta["dino_dim"]], out_row_group=grp)
            off += meta["dino_dim"]
        if depth is not None:
            ops.convert(depth.view(rows, -1), a[:, off : off + meta["depth_dim"]], out_row_group=grp)
        if wt is self.w.c["embed.Wt"]:
            wt = self.w.mats()["embed.Wt"]       # fp32 mode: the three-term split when the shape allows
        x = ops.gemm(a, wt, bias, out_dtype=torch.float32)
        if ro:
            ops.set_rows(x, T + 1, self.w.f32["readout_token"].view(-1), B * N)
        return x

    def embed_tracks_from_maps(self, tracks_2d, depth_map, dino_map, video_hw, intrinsics=None):
        """The inference pipeline's lift -> sample -> embed (inference.py:543-557 + track_autoencoder_3d.py:123-165) without
        the per-track features: one clip, tracks_2d [N,T,2] px, depth_map [T,H,W,(1)], dino_map [T,Hp,Wp,D] (device f32).
        Returns (tokens fp32 [N*(T+1), W] with the read-out slot filled, xyz [N,T,3]).

        "Project, then sample": bilinear sampling and the DINO Dense commute, so the patch map is projected once
        (T*Hp*Wp rows instead of N*T) and each token adds a blend of four projected patch rows in the GEMM epilogue."""
        cfg, meta = self.cfg, self.w.meta
        if self.cdt != torch.bfloat16 or meta["coords"] != 3 or cfg.num_frequencies != 32:
            raise ValueError("embed_tracks_from_maps: bf16 3D path with 32 frequencies only")
        N, T = tracks_2d.shape[:2]
        W = meta["W"]
        wt = self.w.c["embed.Wt"]
        use_dino = dino_map is not None and meta["has_dino"] and cfg.use_dino
        use_depth = depth_map is not None and meta["has_depth"] and cfg.use_depth
        if depth_map is None or not use_dino:
            raise ValueError("embed_tracks_from_maps needs the depth map (lifting) and the DINO patch map")
        xyz, _, dfeat = ops.lift_sample(tracks_2d, depth=depth_map, intrinsics=intrinsics, depth_feature_dim=4, want_dino=False)
        off = meta["fourier_in"]
        Tm, Hp, Wp, D = dino_map.shape
        if Tm != T or D != meta["dino_dim"]:
            raise ValueError(f"dino_map {tuple(dino_map.shape)} does not match T={T}, dino_dim={meta['dino_dim']}")
        dm = dino_map.view(T * Hp * Wp, D)
        if dm.dtype != self.cdt:
            dm = ops.convert(dm, torch.empty(T * Hp * Wp, D, device=self.dev, dtype=self.cdt))
        proj = ops.gemm(dm, wt[:, off : off + D])           # [T*Hp*Wp, W] bf16, no bias (added once per token)
        del dm
        bias = self.w.f32["embed.b_track"].clone()
        ops.axpy(bias, self.w.f32["embed.b_dino"])
        wdep = None
        if use_depth:
            ops.axpy(bias, self.w.f32["embed.b_depth"])
            o2 = off + D
            wdep = wt[:, o2 : o2 + 3].float().t().contiguous()   # [3, W]: only channels 0..2 of the depth feature are non-zero
        x = torch.empty(N * (T + 1), W, device=self.dev, dtype=torch.float32)
        ops.embed_sampled(xyz.view(N * T, 3), tracks_2d.reshape(N * T, 2), dfeat.view(N * T, 4) if use_depth else None, proj, wt,
                          wdep, bias, x, T, Hp, Wp, video_hw, cfg.num_frequencies, cfg.track_scale_factor)
        ops.set_rows(x, T + 1, self.w.f32["readout_token"].view(-1), N)
        return x, xyz

    def encode_from_maps(self, inputs):
        """encode() for the inference pipeline's raw products (one clip): ``support_tracks_2d`` [N,T,2] px,
        ``support_tracks_visible`` [N,T,1], ``depth`` [T,H,W,1], ``dino_map`` [T,Hp,Wp,D], ``video_shape`` (T,H,W,3),
        optional ``intrinsics`` and ``boundary_frame``.  Returns (latents [1,128,latent] f32, support xyz [N,T,3])."""
        meta, dev = self.w.meta, self.dev
        tr2 = _as_dev(inputs["support_tracks_2d"], torch.float32, dev)
        N, T = tr2.shape[:2]
        vis = _as_dev(inputs["support_tracks_visible"], torch.float32, dev).reshape(1, N, T, 1)
        depth = _as_dev(inputs["depth"], torch.float32, dev)
        dino = _feat_dev(inputs["dino_map"], dev)      # float32 as the backbone delivers it, or bfloat16 when the caller holds it so
        _, H, Wv = inputs["video_shape"][:3]
        boundary = _as_dev(inputs.get("boundary_frame", np.array([T], np.int32)), torch.int32, dev)
        x, xyz = self.embed_tracks_from_maps(tr2, depth, dino, (H, Wv), inputs.get("intrinsics"))
        key_mask = ops.build_key_mask(vis, boundary, has_readout=True)
        st = self.transformer("itt", x, N, T + 1, key_mask, out_rows="first")
        nl, E = meta["latent_tokens"], meta["E"]
        lat = self.w.f32["latents_init"].unsqueeze(0).reshape(nl, E).contiguous()
        lat = self.transformer("t2l", lat, 1, nl, kv=st, Lkv=N)
        z = ops.gemm(lat, self.w.mats()["compressor.Wt"], self.w.f32["compressor.b"], out_dtype=torch.float32)
        return z.view(1, nl, meta["latent_dim"]), xyz

    def encode(self, inputs):
        """encode (track_autoencoder_3d.py:190-204 / track_autoencoder.py:234-246) -> [B,128,latent] f32."""
        meta = self.w.meta
        dev = self.dev
        three_d = meta["coords"] == 3
        boundary = _as_dev(inputs["boundary_frame"], torch.int32, dev)
        B, N, T, _ = inputs["support_tracks"].shape
        # host-resident inputs: overlap the host->device copies (1.26 GB of fp32 features per clip,
        # PCIe-bound) with the per-track work, chunk by chunk - tracks are independent until pooling
        stream_keys = ["support_tracks", "support_tracks_visible"] + [k for k in ("dino_features", "depth_features") if inputs.get(k) is not None]
        streamed = three_d and self.stream_chunk > 0 and N >= 2 * self.stream_chunk and all(
            isinstance(inputs.get(k), torch.Tensor) and not inputs[k].is_cuda for k in stream_keys)   # NumPy inputs: one-shot upload
        tracks = visible = None
        if not streamed:
            tracks = _as_dev(inputs["support_tracks"], torch.float32, dev)
            visible = _as_dev(inputs["support_tracks_visible"], torch.float32, dev)
        dino = depth = None
        if three_d and not streamed:
            if self.cfg.use_dino and meta["has_dino"] and inputs.get("dino_features") is not None:
                dino = _feat_dev(inputs["dino_features"], dev)
            if self.cfg.use_depth and meta["has_depth"] and inputs.get("depth_features") is not None:
                depth = _feat_dev(inputs["depth_features"], dev)
        if streamed:
            st = self._encode_tracks_streamed(inputs, boundary, B, N, T)
        elif three_d:
            x = self.embed_tracks(tracks, dino, depth, readout=True)
            key_mask = ops.build_key_mask(visible, boundary, has_readout=True)
            st = self.transformer("itt", x, B * N, T + 1, key_mask, out_rows="first")  # [B*N, W]
        else:
            x = self.embed_tracks(tracks, dino, depth, readout=False)
            key_mask = ops.build_key_mask(visible, boundary, has_readout=False)
            tok = self.transformer("itt", x, B * N, T, key_mask)  # [B*N*T, W]
            st = self._masked_mean(tok, visible, B * N, T)
            del tok
        nl, E = meta["latent_tokens"], meta["E"]
        lat = self.w.f32["latents_init"].unsqueeze(0).expand(B, nl, E).reshape(B * nl, E).contiguous()
        lat = self.transformer("t2l", lat, B, nl, kv=st, Lkv=N)
        z = ops.gemm(lat, self.w.mats()["compressor.Wt"], self.w.f32["compressor.b"], out_dtype=torch.float32)
        return z.view(B, nl, meta["latent_dim"])

    def _encode_tracks_streamed(self, inputs, boundary, B, N, T):
        """encode_tracks (track_autoencoder_3d.py:151-188) on host-resident inputs, in chunks of
        ``stream_chunk`` support tracks: the copy of chunk i+1 (side stream, pinned source) runs under
        the embedding + temporal transformer of chunk i.  Same arithmetic per track as the one-shot path."""
        meta, dev, cfg = self.w.meta, self.dev, self.cfg
        cur = torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        cs.wait_stream(cur)
        use_dino = cfg.use_dino and meta["has_dino"] and inputs.get("dino_features") is not None
        use_depth = cfg.use_depth and meta["has_depth"] and inputs.get("depth_features") is not None
        keys = ["support_tracks", "support_tracks_visible"] + (["dino_features"] if use_dino else []) + (["depth_features"] if use_depth else [])
        keep = lambda k, t: t.dtype == torch.float32 or (t.dtype == torch.bfloat16 and k in ("dino_features", "depth_features"))
        host = {k: (inputs[k] if keep(k, inputs[k]) else inputs[k].float()) for k in keys}
        pieces = [(b, n0, min(n0 + self.stream_chunk, N)) for b in range(B) for n0 in range(0, N, self.stream_chunk)]
        # taper the tail: whatever is computed after the LAST upload has landed is not hidden by any copy, so the final
        # chunk is split 1/2, 1/4, 1/4 (the last piece's transformer pass is then a quarter as long)
        b_, n0_, n1_ = pieces[-1]
        if n1_ - n0_ >= 128 and (n1_ - n0_) % 4 == 0:
            q = (n1_ - n0_) // 4
            pieces[-1:] = [(b_, n0_, n0_ + 2 * q), (b_, n0_ + 2 * q, n0_ + 3 * q), (b_, n0_ + 3 * q, n1_)]

        # two persistent staging slots per input (no allocator traffic across streams: a tensor
        # allocated on the copy stream and freed after use on the compute stream makes the caching
        # allocator wait on cross-stream events and fall back to cudaMalloc, which stalls everything)
        C = self.stream_chunk
        for slot in (0, 1):
            for k in keys:
                shp = (1, C) + tuple(host[k].shape[2:])
                cur_buf = self._stage.get((k, slot))
                if cur_buf is None or cur_buf.shape != shp or cur_buf.dtype != host[k].dtype:
                    self._stage[(k, slot)] = torch.empty(shp, device=dev, dtype=host[k].dtype)
        done = [None, None]   # compute-finished events per slot

        def upload(i):
            b, n0, n1 = pieces[i]
            slot = i & 1
            if done[slot] is not None:
                cs.wait_event(done[slot])          # the chunk that used this slot two steps ago has been consumed
            with torch.cuda.stream(cs):
                dv = {}
                for k in keys:
                    view = self._stage[(k, slot)][:, : n1 - n0]
                    view.copy_(host[k][b : b + 1, n0:n1], non_blocking=True)
                    dv[k] = view
                ev = torch.cuda.Event()
                ev.record(cs)
            return dv, ev

        out = torch.empty(B * N, meta["W"], device=dev, dtype=self.cdt)
        nxt = upload(0)
        for i, (b, n0, n1) in enumerate(pieces):
            dv, ev = nxt
            if i + 1 < len(pieces):
                nxt = upload(i + 1)
            cur.wait_event(ev)
            x = self.embed_tracks(dv["support_tracks"], dv.get("dino_features"), dv.get("depth_features"), readout=True)
            km = ops.build_key_mask(dv["support_tracks_visible"], boundary[b : b + 1], has_readout=True)
            st = self.transformer("itt", x, n1 - n0, T + 1, km, out_rows="first")
            out[b * N + n0 : b * N + n1].copy_(st)
            done[i & 1] = torch.cuda.Event()
            done[i & 1].record(cur)
            del x, st, dv
        return out

    def _masked_mean(self, tok, visible, seqs, T):
        """TRAJAN pooling (track_autoencoder.py:230-232): sum(tok*vis)/max(1,sum vis)."""
        return ops.masked_mean_fwd(tok, visible.reshape(seqs, T).contiguous(), seqs, T, self.cdt)

    # ---- decoder ---------------------------------------------------------------------------------
    def get_decoder_context(self, inputs):
        """get_decoder_context (track_autoencoder_3d.py:206-233): Fourier features of the query
        xyz (always the exact sine: they are re-embedded at 1290x gain) and int32 query frames."""
        cfg, dev = self.cfg, self.dev
        coords = self.w.meta["coords"]
        if "query_points" in inputs and inputs["query_points"] is not None:
            qp = _as_dev(inputs["query_points"], torch.float32, dev)
            B, Q, _ = qp.shape
            xyz = qp[..., 1:].reshape(B * Q, coords).contiguous()
            qframe = torch.round(qp[..., 0]).to(torch.int32).contiguous()  # half-to-even, like jnp.round
        else:
            B = inputs["support_tracks"].shape[0]
            gc = torch.arange(32, dtype=torch.float32) / 32.0 + 1.0 / 64.0
            qx, qy = torch.meshgrid(gc, gc, indexing="xy")
            comps = [qx, qy] + ([torch.zeros_like(qx)] if coords == 3 else [])
            grid = torch.stack(comps, dim=-1).reshape(-1, coords)
            Q = grid.shape[0]
            xyz = grid.unsqueeze(0).expand(B, Q, coords).reshape(B * Q, coords).contiguous().to(dev)
            qframe = torch.zeros(B, Q, dtype=torch.int32, device=dev)
        dq = torch.empty(B * Q, coords * 2 * cfg.num_frequencies, device=dev, dtype=torch.float32)
        ops.fourier_features(xyz, dq, cfg.num_frequencies, cfg.track_scale_factor, exact=True)
        return DecoderContext(dq.view(B, Q, -1), qframe, inputs.get("boundary_frame"))

    def decode(self, latents, ctx, noise=None, discretize=True):
        """decode (track_autoencoder_3d.py:248-307) -> head output [B*Q, 4T] f32."""
        cfg, meta, dev = self.cfg, self.w.meta, self.dev
        w, f = self.w.mats(), self.w.f32
        latents = _as_dev(latents, torch.float32, dev)
        B, nl, ld_ = latents.shape
        if discretize:
            if noise is None:
                # the reference's own draw: jax.random.uniform(PRNGKey(0), latents.shape), the same tensor on every call
                # (track_autoencoder_3d.py:254-257); Threefry restated in rng.py, counter layout chosen on the module
                key = (B, nl, ld_, getattr(cfg, "noise_layout", "original"))
                if self._noise_cache is None or self._noise_cache[0] != key:
                    self._noise_cache = (key, torch.from_numpy(rng.jax_uniform(0, (B, nl, ld_), key[3])).to(dev))
                noise = self._noise_cache[1]
            noise = _as_dev(noise, torch.float32, dev)
        zq = ops.quantize_fwd(latents.reshape(B * nl, ld_), noise.reshape(B * nl, ld_) if discretize else None, discretize)
        zc = zq if self.cdt == torch.float32 else ops.convert(zq, torch.empty_like(zq, dtype=self.cdt))
        x = ops.gemm(zc, w["decompressor.Wt"], f["decompressor.b"], out_dtype=torch.float32)
        C = meta["D"] - 128
        lat = self.transformer("dec", x, B, nl)  # [B*nl, C] cdt
        Q = ctx.query_frame.shape[1]
        dq = ctx.decoder_query.reshape(B * Q, -1)
        qfeat = torch.empty(B * Q, meta["query_in"], device=dev, dtype=self.cdt)
        # query_frame // time_scale_factor (:268-269) is 0 for every in-range frame: tail_zero
        # (validated on the host once; a stream that is being captured into a CUDA graph cannot synchronise,
        # and the token-assembly kernel is memory safe for any frame index)
        if not torch.cuda.is_current_stream_capturing():
            if int(ctx.query_frame.max().item()) >= cfg.time_scale_factor or int(ctx.query_frame.min().item()) < 0:
                raise ValueError("query frames outside [0, time_scale_factor) are not supported")
        ops.fourier_features(dq, qfeat, cfg.num_frequencies, cfg.track_scale_factor, tail_zero=True, exact=self.exact)
        qe = ops.gemm(qfeat, w["query_encoder.Wt"], f["query_encoder.b"], out_dtype=torch.float32)
        tokens = torch.empty(B * Q * (nl + 1), meta["D"], device=dev, dtype=torch.float32)
        ops.decoder_tokens_fwd(lat, qe, ctx.query_frame, tokens, B, Q, nl, C)
        out = self.transformer("tra", tokens, B * Q, nl + 1, out_rows="first")  # [B*Q, D]
        return ops.gemm(out, w["track_predictor.Wt"], f["track_predictor.b"], out_dtype=torch.float32)


@dataclass
class DecoderContext:
    """TrackAutoEncoderDecoderContext (track_autoencoder.py:108-114)."""

    decoder_query: torch.Tensor
    query_frame: torch.Tensor
    boundary_frame: object
