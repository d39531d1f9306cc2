#!/usr/bin/env python
"""model.residual_dtype = "fp32" vs "bf16" on the full cfg2 clip: parity against the fp32 oracle (quantiser on, the metric and the
flip-aware comparison of tests/test_gpu_parity_full.py) and the device-resident forward time (CUDA-graph replay)."""
import importlib, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import model as om
from tests.helpers import make_inputs, quantiser_aware_reference, rel_err
spa = importlib.import_module("3dspa_code_b200")
dev = torch.device("cuda:0")
c = om.Config3D()
inp, noise = make_inputs(c, B=1, N=2048, Q=512, seed=21, vis_p=0.9)
out = {}
ref_cache = None
for rd in ("fp32", "bf16"):
    model = spa.TrackAutoEncoder3D()
    model.residual_dtype = rd
    variables = model.init(22, inp)
    om._randomize(variables["params"], np.random.RandomState(22))
    got = model.apply(variables, inp, noise=noise, discretize=True, precision="bf16")
    z_gpu = model.apply(variables, inp, method="encode", precision="bf16")
    ref_plain, ref_aware, z_ref, flipped = quantiser_aware_reference(variables["params"], c, inp, noise, z_gpu)
    out[rd] = {"latents": rel_err(z_gpu, z_ref), "flipped_pct": 100 * flipped,
               "tracks_same_rounding": rel_err(got.tracks, ref_aware.tracks), "logits_same_rounding": rel_err(got.visible_logits, ref_aware.visible_logits),
               "tracks_plain": rel_err(got.tracks, ref_plain.tracks), "logits_plain": rel_err(got.visible_logits, ref_plain.visible_logits)}
    # speed: bench's clip, graph replay
    m2 = spa.TrackAutoEncoder3D()
    m2.residual_dtype = rd
    m2.cuda_graph = True
    inputs, nz = bench.synth_clip(100, device=dev)
    v2 = m2.init(0, {"dino_features": 1, "depth_features": 1})
    for _ in range(5):
        m2.apply(v2, inputs, noise=nz, precision="bf16")
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m2.apply(v2, inputs, noise=nz, precision="bf16")
        e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1) / 20, 3))
    out[rd]["ms_per_clip"] = ts
    print(json.dumps({rd: out[rd]}), flush=True)
    del model, m2
    torch.cuda.empty_cache()
