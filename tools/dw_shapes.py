#!/usr/bin/env python
"""Development tool: time the weight-gradient GEMM (dW = dY^T X) on the shapes of one clip."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spa = importlib.import_module("3dspa_code_b200")
ops = spa.ops
dev = torch.device("cuda")
Mi, Mr = 2048 * 151, 512 * 129
shapes = [("itt.qkv", Mi, 2304, 384), ("itt.out", Mi, 384, 768), ("itt.mlp1", Mi, 1536, 384), ("itt.mlp2", Mi, 384, 1536),
          ("tra.qkv", Mr, 2304, 1280), ("tra.out", Mr, 1280, 768), ("tra.mlp1", Mr, 1536, 1280), ("tra.mlp2", Mr, 1280, 1536),
          ("embed", Mi, 384, 1280)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0
for name, M, N, K in shapes:
    dy = torch.randn(M, N, device=dev).to(torch.bfloat16)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    dw = torch.zeros(N, K, device=dev)
    ts = []
    for i in range(6):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm_dw(dy, x, dw, accumulate=True); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts[1:])[2]
    tot += ms
    print(json.dumps({"name": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1),
                      "hbm_floor_ms": round((M * (N + K) * 2) / 6544.7e6, 4)}))
print(json.dumps({"total_ms": tot}))
