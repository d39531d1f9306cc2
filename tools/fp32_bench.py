#!/usr/bin/env python
"""Development tool: time the accurate (precision="fp32") path on the cfg2 clip and report which contraction kernels it took."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
spa = importlib.import_module("3dspa_code_b200")
dev = torch.device("cuda")
model = spa.TrackAutoEncoder3D()
variables = model.init(0, {"dino_features": 1, "depth_features": 1})
inputs, noise = bench.synth_clip(100, device=dev)
spa.ops.stats(reset=True)
out = bench.run_fp32_leg(spa, model, variables, inputs, noise, dev)
out["dispatch"] = {k: v for k, v in spa.ops.stats().items() if v}
print(json.dumps(out))
