#!/usr/bin/env python
"""Development tool: CUDA-event time of the phases of one cfg2 inference forward (eager launches)."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from train_bench import synth_batch  # noqa: E402


def main():
    spa = importlib.import_module("3dspa_code_b200")
    ops = spa.ops
    dev = torch.device("cuda")
    model = spa.TrackAutoEncoder3D()
    variables = model.init(0, {"dino_features": 1, "depth_features": 1})
    eng = model.bind(variables, "bf16", dev)
    batch, noise = synth_batch(1, 1, dev)
    marks = []
    orig_tr = eng.transformer

    def timed_tr(short, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = orig_tr(short, *a, **k)
        e.record()
        marks.append((short, s, e))
        return out

    eng.transformer = timed_tr
    for _ in range(3):
        marks.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.apply(variables, batch, noise=noise, precision="bf16")
        e1.record()
        torch.cuda.synchronize()
    out = {"total_ms": e0.elapsed_time(e1)}
    for short, s, e in marks:
        out[short] = out.get(short, 0.0) + s.elapsed_time(e)
    out["other"] = out["total_ms"] - sum(v for k, v in out.items() if k != "total_ms")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
