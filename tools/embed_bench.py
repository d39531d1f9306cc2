#!/usr/bin/env python
"""Development tool: time the fused track-embedding kernel (K1) on one clip and report HBM GB/s."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spa = importlib.import_module("3dspa_code_b200")
ops = spa.ops
dev = torch.device("cuda")
N, T, Dd, Dz, W = 2048, 150, 768, 256, 384
R = N * T
tracks = torch.rand(R, 3, device=dev) * 2 - 1
dino = torch.randn(R, Dd, device=dev)
depth = torch.randn(R, Dz, device=dev)
wt = (torch.randn(W, 256 + Dd + Dz, device=dev) / 36).to(torch.bfloat16)
bias = torch.randn(W, device=dev)
out = torch.empty(N * (T + 1), W, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(8):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.embed_fused(tracks, dino, depth, wt, bias, out, T, 32, 1.0)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sorted(ts[2:])[len(ts[2:]) // 2]
nbytes = R * ((Dd + Dz) * 4 + 12) + R * W * 4   # fp32 features + xyz in, fp32 tokens out
print(json.dumps({"kernel": "embed_fused", "ms": ms, "bytes": nbytes, "GBps": nbytes / ms / 1e6, "frac_of_6544.7": nbytes / ms / 1e6 / 6544.7,
                  "tflops": 2.0 * R * W * (256 + Dd + Dz) / ms / 1e9}))
