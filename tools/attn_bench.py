#!/usr/bin/env python
"""Development tool: time the attention cores on the shapes of one clip (ITT L=151 x 2048 tracks, TRA L=129 x 512
queries, cross 128 x 2048) forward and backward, with the HBM floor of each call."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spa = importlib.import_module("3dspa_code_b200")
ops = spa.ops
dev = torch.device("cuda")
H, Dh = 8, 96
A = H * Dh
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]

for name, batch, Lq, Lk, masked in (("itt", 2048, 151, 151, True), ("tra", 512, 129, 129, False), ("t2l.cross", 1, 128, 2048, False)):
    qkv = torch.randn(batch * max(Lq, Lk), 3 * A, device=dev).to(torch.bfloat16)
    q = (qkv[: batch * Lq, :A] / Dh ** 0.5).contiguous() if Lq != Lk else qkv[:, :A]
    k, v = qkv[: batch * Lk, A : 2 * A], qkv[: batch * Lk, 2 * A :]
    mask = (torch.rand(batch, Lk, device=dev) < 0.9).to(torch.uint8) if masked else None
    if mask is not None:
        mask[:, 0] = 1
    o = torch.empty(batch * Lq, A, device=dev, dtype=torch.bfloat16)
    stats = ops.attention_fwd(q, k, v, o, batch, H, Lq, Lk, Dh, mask, save_stats=True)
    d_o = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    dq = torch.empty_like(q) if Lq != Lk else dqkv[:, :A]
    f_ms = timed(lambda: ops.attention_fwd(q, k, v, o, batch, H, Lq, Lk, Dh, mask, save_stats=True))
    b_ms = timed(lambda: ops.attention_bwd(q, k, v, o, d_o, dq, dqkv[: batch * Lk, A : 2 * A], dqkv[: batch * Lk, 2 * A :], stats, batch, H, Lq, Lk, Dh, mask))
    fbytes = (batch * Lq * A * 2 + 2 * batch * Lk * A) * 2
    bbytes = (2 * batch * Lq * A * 2 + 4 * batch * Lk * A) * 2
    flop = 4.0 * batch * H * Lq * Lk * Dh
    print(json.dumps({"shape": name, "fwd_ms": round(f_ms, 4), "fwd_hbm_floor_ms": round(fbytes / 6544.7e6, 4), "fwd_tflops": round(flop / f_ms / 1e9, 1),
                      "bwd_ms": round(b_ms, 4), "bwd_hbm_floor_ms": round(bbytes / 6544.7e6, 4), "bwd_tflops": round(2.5 * flop / b_ms / 1e9, 1)}))
