#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM on the contraction shapes of the 3DSPA step (cfg2, one clip).

Development tool (not the bench): CUDA events around each launch, L2 flushed between launches,
cuBLAS (torch.matmul) on the same shape printed beside it as the practical ceiling.
  python tools/gemm_shapes.py [--clips B] [--reps R]
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
    ap.add_argument("--clips", type=int, default=1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    spa = importlib.import_module("3dspa_code_b200")
    ops = spa.ops
    dev = torch.device("cuda")
    Mi, Mr = 2048 * 151 * args.clips, 512 * 129 * args.clips
    # (name, M, K, N, kind)
    shapes = [
        ("itt.qkv", Mi, 384, 2304, "rms"),
        ("itt.out", Mi, 768, 384, "res"),
        ("itt.mlp1", Mi, 384, 1536, "gelu"),
        ("itt.mlp2", Mi, 1536, 384, "res"),
        ("tra.qkv", Mr, 1280, 2304, "rms"),
        ("tra.out", Mr, 768, 1280, "res"),
        ("tra.mlp1", Mr, 1280, 1536, "gelu"),
        ("tra.mlp2", Mr, 1536, 1280, "res"),
        ("itt.gbwd", Mi, 384, 1536, "gbwd"),
        ("itt.gfwd", Mi, 384, 1536, "gfwd"),
        ("embed", 2048 * 150 * args.clips, 1280, 384, "bias"),
        ("itt.out.resbf", Mi, 768, 384, "resbf"),
        ("itt.mlp2.resbf", Mi, 1536, 384, "resbf"),
        ("tra.mlp2.resbf", Mr, 1536, 1280, "resbf"),
        ("tra.out.f32", Mr, 768, 1280, "f32"),
        ("tra.out.resbf", Mr, 768, 1280, "resbf"),
        ("tra.out.bf16", Mr, 768, 1280, "bias"),
        ("plain.4096", 4096, 4096, 4096, "plain"),
    ]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, M, K, N, kind in shapes:
        if args.only and args.only not in name:
            continue
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        wt = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if kind in ("res", "resbf", "gbwd") else None
        if kind in ("resbf", "gbwd"):
            res = res.to(torch.bfloat16)
        sq = torch.ones(96, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if kind in ("res", "f32") else torch.bfloat16)

        def ours():
            if kind == "gbwd":   # dz = (dy . W2^T) * g'  (saved-derivative backward through MLP_out + GELU)
                return ops.gemm_gelu_bwd(a, wt, res, z_is_grad=True)
            if kind == "gfwd":   # training forward: h and gelu'(z)
                return ops.gemm_gelu(a, wt, bias, save_grad=True)
            if kind == "rms":
                return ops.gemm_rmsnorm(a, wt, 96, 768, 768, sq, sq)
            if kind in ("res", "resbf"):
                return ops.gemm(a, wt, bias, residual=res, out=out)
            if kind == "f32":
                return ops.gemm(a, wt, bias, out=out)
            if kind == "gelu":
                return ops.gemm(a, wt, bias, act=ops.ACT_GELU, out=out)
            if kind == "bias":
                return ops.gemm(a, wt, bias, out=out)
            return ops.gemm(a, wt, out=out)

        def cublas():
            return torch.matmul(a, wt.t())

        res_row = {"name": name, "M": M, "K": K, "N": N, "kind": kind}
        for label, fn in (("ours", ours), ("cublas", cublas)):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[len(ts) // 2]
            res_row[label + "_ms"] = round(ms, 4)
            res_row[label + "_tflops"] = round(2.0 * M * N * K / ms / 1e9, 1)
        rows.append(res_row)
        print(json.dumps(res_row), flush=True)
        del a, wt, res, out
    if not args.only or "fused" in args.only:
        # the fused MLP sub-block against MLP_in (GELU) + MLP_out (residual) as two kernels
        M, D, Hd = Mi, 384, 1536
        a = torch.randn(M, D, device=dev).to(torch.bfloat16)
        w1 = (torch.randn(Hd, D, device=dev) / D ** 0.5).to(torch.bfloat16)
        w2 = (torch.randn(D, Hd, device=dev) / Hd ** 0.5).to(torch.bfloat16)
        b1, b2 = torch.randn(Hd, device=dev), torch.randn(D, device=dev)
        res = torch.randn(M, D, device=dev)
        out = torch.empty(M, D, device=dev)

        def two():
            h = ops.gemm(a, w1, b1, act=ops.ACT_GELU)
            return ops.gemm(h, w2, b2, residual=res, out=out)

        def fused():
            return ops.mlp_fused(a, w1, b1, w2, b2, res)

        row = {"name": "itt.mlp.fused", "M": M, "D": D, "Hd": Hd}
        for label, fn in (("two_kernels", two), ("fused", fused)):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[len(ts) // 2]
            row[label + "_ms"] = round(ms, 4)
            row[label + "_tflops"] = round(4.0 * M * D * Hd / ms / 1e9, 1)
        print(json.dumps(row), flush=True)
    tot_o = sum(r["ours_ms"] for r in rows)
    tot_c = sum(r["cublas_ms"] for r in rows)
    print(json.dumps({"total_ours_ms": tot_o, "total_cublas_ms": tot_c}))


if __name__ == "__main__":
    main()
