#!/usr/bin/env python
"""Development tool: the device-resident cfg2 forward (CUDA-graph replay, as bench.py's `value` leg) timed over N steps.
A/B of library builds: SPA3D_LIB_PATH=3dspa_code_b200/lib3dspa_b200_<tag>.so python tools/forward_time.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
spa = importlib.import_module("3dspa_code_b200")
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
model = spa.TrackAutoEncoder3D()
model.cuda_graph = True
inputs, noise = bench.synth_clip(100, device=dev)
variables = model.init(0, {"dino_features": 1, "depth_features": 1})
for _ in range(5):
    model.apply(variables, inputs, noise=noise, precision="bf16")
torch.cuda.synchronize()
ts = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.apply(variables, inputs, noise=noise, precision="bf16")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / steps)
print(json.dumps({"lib": os.path.basename(spa._lib.LIB_PATH), "ms_per_clip": [round(t, 3) for t in ts]}))
