#!/usr/bin/env python
"""Development tool: time one full-size 3DSPA training step (fwd + bwd + AdamW) on one GPU.

  python tools/train_bench.py [--clips B] [--micro MB] [--steps K] [--profile]
--profile prints the CUDA-event time of every C-ABI entry point summed over the timed steps.
"""
import argparse
import collections
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

T, S, Q = 150, 2048, 512


def synth_batch(B, seed, dev, s=S, q=Q):
    g = torch.Generator(device=dev).manual_seed(seed)
    r = lambda *sh: torch.rand(*sh, generator=g, device=dev)
    batch = {
        "support_tracks": r(B, s, T, 3) * 2 - 1,
        "support_tracks_visible": (r(B, s, T, 1) < 0.9).float(),
        "query_points": torch.cat([torch.randint(0, T, (B, q, 1), generator=g, device=dev).float(), r(B, q, 3) * 2 - 1], -1),
        "boundary_frame": torch.full((B,), T, dtype=torch.int32, device=dev),
        "dino_features": torch.randn(B, s, T, 768, generator=g, device=dev),
        "depth_features": torch.randn(B, s, T, 256, generator=g, device=dev),
        "query_tracks": r(B, q, T, 3) * 2 - 1,
        "query_tracks_visible": (r(B, q, T, 1) < 0.9).float(),
    }
    return batch, r(B, 128, 96)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1)
    ap.add_argument("--micro", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--support", type=int, default=S)
    ap.add_argument("--query", type=int, default=Q)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--detail", default="", help="comma-separated entry points whose per-call times are printed (last step)")
    args = ap.parse_args()
    spa = importlib.import_module("3dspa_code_b200")
    te = importlib.import_module("3dspa_code_b200.train_engine")
    dev = torch.device("cuda")
    model = spa.TrackAutoEncoder3D()
    tree = model.init(0, {"dino_features": 1, "depth_features": 1})["params"]
    trainer = te.Trainer(model, tree, precision="bf16", micro_batch=args.micro)
    batch, noise = synth_batch(args.clips, 1, dev, args.support, args.query)
    trainer.train_step(batch, noise)  # warm-up
    torch.cuda.synchronize()
    prof = collections.defaultdict(list)
    if args.profile:
        ops = spa.ops
        orig = ops._call

        def timed_call(name, *a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *a)
            e.record()
            prof[name].append((s, e))

        ops._call = timed_call
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        log = trainer.train_step(batch, noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    out = {"clips": args.clips, "micro": args.micro, "ms_per_step": ms, "clips_per_s": args.clips / ms * 1e3,
           "model_tflops": 28.24 * args.clips * (args.support / S) / ms * 1e3 / 1e3 if args.query * 4 == args.support else None,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "loss": log["total_loss"], "grad_norm": log["grad_norm"]}
    print(json.dumps(out))
    if args.profile:
        tot = {k: sum(s.elapsed_time(e) for s, e in v) / args.steps for k, v in prof.items()}
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print(f"{v:10.3f} ms  n={len(prof[k]) // args.steps:5d}  {k}")
        print(f"{sum(tot.values()):10.3f} ms  total in kernels")
        for name in filter(None, args.detail.split(",")):
            v = prof["spa3d_" + name]
            per = len(v) // args.steps
            print(name, " ".join(f"{s.elapsed_time(e):.3f}" for s, e in v[-per:]))


if __name__ == "__main__":
    main()
