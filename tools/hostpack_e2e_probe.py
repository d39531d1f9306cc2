"""apply_stream(from_maps=True) with host-side bf16 staging: per-clip time and the duration of every staging call, by thread count."""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
spa = importlib.import_module("3dspa_code_b200")
import bench

dev = torch.device("cuda:0")
model = spa.TrackAutoEncoder3D()
model.cuda_graph = True
host, noise = bench.synth_maps_clip(200, spa, dev)
variables = model.init(0, {"dino_features": 1, "depth_features": 1})
ops = spa.ops
orig = ops.host_pack_bf16
log = []


def timed_pack(src, dst, threads):
    t = time.perf_counter()
    r = orig(src, dst, threads)
    log.append((time.perf_counter() - t) * 1e3)
    return r


ops.host_pack_bf16 = timed_pack
K = 20
for pack, th in ((None, None), ("bf16", 14), ("bf16", 12), ("bf16", 10), ("bf16", 8), (None, None), ("bf16", 12)):
    for n in (4, K):
        log.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        stamps = []
        for res in model.apply_stream(variables, (host for _ in range(n)), noises=(noise for _ in range(n)), from_maps=True, host_pack=pack, pack_threads=th):
            stamps.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
    if pack and th == 12 and os.environ.get("SPA3D_TRACE"):
        model.stream_trace = []
        list(model.apply_stream(variables, (host for _ in range(8)), noises=(noise for _ in range(8)), from_maps=True, host_pack=pack, pack_threads=th))
        torch.cuda.synchronize()
        tr = model.stream_trace
        del model.stream_trace
        t00 = tr[0][2]
        ev0 = next(e for _, _, _, e in tr if e is not None)
        for what, i, t, e in sorted(tr, key=lambda r: r[2]):
            print(f"  {what:14s} clip {i}  host {1e3 * (t - t00):8.2f} ms" + (f"   device {ev0.elapsed_time(e):8.2f} ms" if e is not None else ""))
    print(json.dumps({"host_pack": pack, "threads": th, "ms_per_clip": round(dt / K, 2), "pack_ms": [round(x, 1) for x in log],
                      "yield_ms": [round(x, 1) for x in stamps]}), flush=True)
