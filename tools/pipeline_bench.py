#!/usr/bin/env python
"""Development tool: the inference pipeline's lift -> sample -> embed stage (inference.py:543-557 + track_autoencoder_3d.py:123-149)
as two kernels with the per-track features in HBM (K0 + K1) against the fused "project, then sample" form.

  python tools/pipeline_bench.py [--tracks N] [--reps R]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    spa = importlib.import_module("3dspa_code_b200")
    ops = spa.ops
    dev = torch.device("cuda")
    N, T, H, W, Hp, Wp = args.tracks, 150, 518, 518, 37, 37
    g = torch.Generator(device=dev).manual_seed(4)
    model = spa.TrackAutoEncoder3D()
    variables = model.init(0, {"dino_features": 1, "depth_features": 1})
    eng = model.bind(variables, "bf16")
    tracks = (torch.rand(N, T, 2, generator=g, device=dev) * 530 - 6).contiguous()
    depth = torch.rand(T, H, W, 1, generator=g, device=dev) * 9.5 + 0.5
    dino = torch.randn(T, Hp, Wp, 768, generator=g, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def unfused():
        xyz, df, zf = ops.lift_sample(tracks, depth=depth, dino=dino, video_hw=(H, W))
        return eng.embed_tracks(xyz[None], df[None], zf[None], readout=True)

    def fused():
        return eng.embed_tracks_from_maps(tracks, depth, dino, (H, W))[0]

    out = {"tracks": N}
    for name, fn in (("unfused_ms", unfused), ("fused_ms", fused)):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
            del y
        out[name] = sorted(ts)[len(ts) // 2]
    # per entry point, one fused pass
    import collections
    prof = collections.OrderedDict()
    orig = ops._call

    def timed(name, *a_):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        orig(name, *a_)
        e_.record()
        prof.setdefault(name, []).append((s_, e_))

    ops._call = timed
    flush.zero_()
    fused()
    torch.cuda.synchronize()
    ops._call = orig
    out["fused_stages_ms"] = {k: round(sum(s_.elapsed_time(e_) for s_, e_ in v), 4) for k, v in prof.items()}
    a, b = unfused(), fused()
    out["rel_err"] = float((a - b).abs().max() / a.abs().max())
    out["feature_bytes_avoided"] = N * T * 1024 * 4 * 2
    print(json.dumps(out))


if __name__ == "__main__":
    main()
