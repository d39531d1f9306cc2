"""Drop-in mirror of the reference module API for the 3DSPA hot path.

``TrackAutoEncoder3D`` keeps the constructor fields (track_autoencoder_3d.py:53-67), the batch-dict
keys (``TrackAutoEncoder3DInputs``, :23-40), the ``init`` / ``apply`` call shapes the reference's
scripts use (train.py:137-146,221-233; inference.py:594-623) and the result container
(``TrackAutoEncoderResults``, track_autoencoder.py:72-105), so a caller written against the Flax
module only swaps the import.  The body runs on the sm_100a kernels of lib3dspa_b200.so.

Differences that are deliberate and documented:
  * arrays are torch CUDA tensors (or NumPy on input), not jax Arrays;
  * ``apply(..., noise=...)``: the quantiser noise, drawn by the reference from
    ``jax.random.uniform(PRNGKey(0))`` (:254-258), is an explicit input because that bit stream
    cannot be generated without JAX;
  * ``precision`` selects "bf16" (tcgen05 throughput path) or "fp32" (accurate path);
  * the as-written defects R1 (mask shapes) and R2 (projection widths) are repaired.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np
import torch

from . import ops, params as P
from .engine import DecoderContext, DeviceWeights, Engine


@dataclass
class TrackAutoEncoderResults:
    """track_autoencoder.py:72-105."""

    tracks: torch.Tensor            # [B,Q,T,3] (2 for TRAJAN)
    visible_logits: torch.Tensor    # [B,Q,T,1]
    certain_logits: torch.Tensor    # [B,Q,T,1]

    @property
    def visible(self):
        return (self.visible_logits > 0).to(torch.float32)

    @property
    def certain(self):
        return (self.certain_logits > 0).to(torch.float32)

    @property
    def visible_and_certain(self):
        return ((torch.sigmoid(self.visible_logits) * torch.sigmoid(self.certain_logits)) > 0.5).to(torch.float32)


TrackAutoEncoderDecoderContext = DecoderContext


class _ModuleBase:
    _coords = 3
    _arch = None

    def __post_init__(self):
        self._bound = {}  # (id(tree), precision) -> (weakref-able holder, Engine)
        self.cuda_graph = False   # replay the whole forward as ONE CUDA graph when the same device buffers are passed again
        self._graphs = {}

    # -- Flax-style API ---------------------------------------------------------------------------
    def init(self, rng=0, batch=None, arch=None):
        """``model.init(rng, batch)`` -> {'params': tree}.  ``rng`` is an int seed (or anything with
        an int() / a JAX PRNGKey-like array whose last element is used)."""
        seed = int(np.asarray(rng).reshape(-1)[-1]) if not isinstance(rng, int) else rng
        has_dino = batch is not None and batch.get("dino_features") is not None
        has_depth = batch is not None and batch.get("depth_features") is not None
        tree = P.init_tree(self, seed, has_dino, has_depth, arch or self._arch, self._coords)
        return {"params": tree}

    def bind(self, params, precision="bf16", device="cuda") -> Engine:
        """Upload + pack a parameter tree once and return the executor bound to it."""
        tree = params["params"] if "params" in params and isinstance(params["params"], dict) else params
        key = (id(tree), precision, str(device))
        hit = self._bound.get(key)
        if hit is not None and hit[0] is tree:
            return hit[1]
        self._check_tree(tree)
        eng = Engine(self, DeviceWeights.from_tree(tree, precision, device))
        self._bound = {key: (tree, eng)}  # keep one binding alive (weights are 0.4-0.9 GB)
        return eng

    def _check_tree(self, tree):
        meta = P.tree_meta(tree)
        if meta["coords"] != self._coords:
            raise ValueError(f"parameter tree is for a {meta['coords']}D model")
        D = meta["D"]
        if np.shape(tree["decompressor"]["kernel"])[1] != D - 128:
            # append_time_feat asserts latents.shape[-1] == decoder_num_channels - 128 (:237)
            raise AssertionError("decompressor width must equal decoder_num_channels - 128")
        if meta["head_out"] != 4 * self.num_output_frames:
            raise ValueError("track_predictor width must be 4 * num_output_frames")

    def apply(self, variables, inputs, *, method=None, noise=None, discretize=True, precision="bf16", rngs=None,
              **method_kwargs):
        """``model.apply({'params': p}, batch)`` -> TrackAutoEncoderResults."""
        eng = self.bind(variables, precision)
        if method is not None:
            name = method if isinstance(method, str) else method.__name__
            if name == "encode":
                return eng.encode(inputs)
            if name == "get_decoder_context":
                return eng.get_decoder_context(inputs)
            if name == "decode":
                ctx = method_kwargs.get("decoder_context", inputs if isinstance(inputs, DecoderContext) else None)
                lat = method_kwargs.get("latents")
                return self._results(eng, eng.decode(lat, ctx, noise, method_kwargs.get("discretize", discretize)), ctx)
            raise ValueError(f"unknown method {name}")
        if self.cuda_graph and self._graphable(inputs, noise, discretize):
            return self._forward_graphed(eng, inputs, noise, discretize, precision)
        return self._forward(eng, inputs, noise, discretize)

    def __call__(self, variables, inputs, **kw):
        return self.apply(variables, inputs, **kw)

    # -- the reference's methods, bound to already-uploaded weights ----------------------------------
    def encode(self, variables, inputs, precision="bf16"):
        return self.bind(variables, precision).encode(inputs)

    def get_decoder_context(self, variables, inputs, precision="bf16"):
        return self.bind(variables, precision).get_decoder_context(inputs)

    def decode(self, variables, latents, decoder_context, discretize=True, noise=None, precision="bf16"):
        eng = self.bind(variables, precision)
        return self._results(eng, eng.decode(latents, decoder_context, noise, discretize), decoder_context)

    def apply_from_maps(self, variables, inputs, *, noise=None, discretize=True, precision="bf16"):
        """The reference's run_inference model call (inference.py:543-623) fed with the pipeline's RAW products instead of
        per-track features: ``support_tracks_2d`` [N,T,2] px, ``support_tracks_visible`` [N,T,1], ``depth`` [T,H,W,1],
        ``dino_map`` [T,Hp,Wp,768], ``video_shape``, ``query_points`` [1,Q,4] (t,x,y,z) (+ ``intrinsics``, ``boundary_frame``).
        Lifting, DINO / depth sampling and the embedding run fused: the [N,T,768] and [N,T,256] features are
        never written.  Same result as lifting.lift_and_sample + apply within the bf16 tolerance."""
        eng = self.bind(variables, precision)
        with torch.no_grad():
            latents, _ = eng.encode_from_maps(inputs)
            ctx = eng.get_decoder_context({"query_points": inputs["query_points"],
                                           "boundary_frame": inputs.get("boundary_frame", np.array([inputs["support_tracks_2d"].shape[1]], np.int32))})
            return self._results(eng, eng.decode(latents, ctx, noise, discretize), ctx)

    # -- CUDA-graph replay of the forward ---------------------------------------------------------------
    def _graphable(self, inputs, noise, discretize):
        if self.decoder_scan_chunk_size is not None or inputs.get("query_points") is None:
            return False
        ts = [v for v in inputs.values() if v is not None] + ([noise] if discretize else [])
        return all(isinstance(t, torch.Tensor) and t.is_cuda for t in ts)

    def _forward_graphed(self, eng, inputs, noise, discretize, precision):
        """The ~145 launches of a forward (90 of them under 20 us) captured once into a CUDA graph and
        replayed: the graph is keyed on the addresses and shapes of the device input buffers, so a
        pipeline that refills the same buffers pays one launch per clip.  The returned tensors are
        owned by the graph and are overwritten by the next replay with the same buffers."""
        ts = {k: v for k, v in inputs.items() if isinstance(v, torch.Tensor)}
        key = (id(eng), precision, bool(discretize), noise.data_ptr() if discretize else 0,
               tuple(sorted((k, t.data_ptr(), tuple(t.shape), str(t.dtype)) for k, t in ts.items())))
        hit = self._graphs.get(key)
        if hit is None:
            self._forward(eng, inputs, noise, discretize)   # eager once: lazy init, input validation
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._forward(eng, inputs, noise, discretize)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = self._forward(eng, inputs, noise, discretize)
            if len(self._graphs) >= 2:   # each graph pins its activations (a few GB): keep two
                self._graphs.pop(next(iter(self._graphs)))
            hit = (g, res, dict(inputs), noise)
            self._graphs[key] = hit
        hit[0].replay()
        return hit[1]

    # -- internals ------------------------------------------------------------------------------------
    def _results(self, eng, head_out, ctx):
        B, Q = ctx.query_frame.shape
        T = self.num_output_frames
        tracks, vis, cert = ops.split_outputs(head_out, T, self._coords)
        c = self._coords
        return TrackAutoEncoderResults(tracks.view(B, Q, T, c), vis.view(B, Q, T, 1), cert.view(B, Q, T, 1))

    def _forward(self, eng, inputs, noise, discretize):
        """__call__ (track_autoencoder_3d.py:309-357) incl. the chunked decode of :315-349."""
        with torch.no_grad():
            latents = eng.encode(inputs)
            h = self.decoder_scan_chunk_size
            if h is None or "query_points" not in inputs:
                ctx = eng.get_decoder_context(inputs)
                return self._results(eng, eng.decode(latents, ctx, noise, discretize), ctx)
            qp = inputs["query_points"]
            Q = qp.shape[-2]
            if Q % h:
                raise ValueError(f"decoder_scan_chunk_size={h} must divide the number of queries {Q}")
            outs = []
            for s in range(0, Q, h):
                sub = dict(inputs)
                sub["query_points"] = qp[..., s : s + h, :]
                ctx = eng.get_decoder_context(sub)
                outs.append(self._results(eng, eng.decode(latents, ctx, noise, discretize), ctx))
            return TrackAutoEncoderResults(
                torch.cat([o.tracks for o in outs], dim=1),
                torch.cat([o.visible_logits for o in outs], dim=1),
                torch.cat([o.certain_logits for o in outs], dim=1),
            )


@dataclass
class TrackAutoEncoder3D(_ModuleBase):
    """3DSPA (track_autoencoder_3d.py:43-357)."""

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

    _coords = 3
    _arch = None


@dataclass
class TrackAutoEncoder(_ModuleBase):
    """TRAJAN 2D (track_autoencoder.py:117-390) on the same kernels."""

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

    _coords = 2
    _arch = None
    use_dino = False
    use_depth = False
