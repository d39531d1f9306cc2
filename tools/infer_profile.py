#!/usr/bin/env python
"""Development tool: CUDA-event time per C-ABI entry point over one eager cfg2 inference forward."""
import collections
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from train_bench import synth_batch  # noqa: E402

spa = importlib.import_module("3dspa_code_b200")
ops = spa.ops
dev = torch.device("cuda")
model = spa.TrackAutoEncoder3D()
variables = model.init(0, {"dino_features": 1, "depth_features": 1})
batch, noise = synth_batch(1, 1, dev)
for _ in range(2):
    model.apply(variables, batch, noise=noise, precision="bf16")
prof = collections.defaultdict(list)
orig = ops._call


def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    orig(name, *a)
    e.record()
    prof[name].append((s, e))


ops._call = timed
model.apply(variables, batch, noise=noise, precision="bf16")
torch.cuda.synchronize()
tot = {k: sum(s.elapsed_time(e) for s, e in v) for k, v in prof.items()}
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:8.3f} ms  n={len(prof[k]):4d}  {k}")
print(f"{sum(tot.values()):8.3f} ms total")
big = sorted(((s.elapsed_time(e), k) for k, v in prof.items() for s, e in v if k in ("spa3d_gemm", "spa3d_fourier_features")), reverse=True)[:12]
print(big)
