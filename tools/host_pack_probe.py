"""Host float32 -> bfloat16 staging throughput (spa3d_host_pack_bf16) at several thread counts, beside torch's own CPU conversion,
alone and while a host->device copy of the packed buffer is in flight (what model.apply_stream(host_pack="bf16") does)."""
import importlib
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
spa = importlib.import_module("3dspa_code_b200")
from importlib import import_module
L = import_module("3dspa_code_b200._lib").lib()

n = 150 * 37 * 37 * 768            # one clip's DINOv2 patch map
pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
src = pin(torch.randn(n))
src[:4] = torch.tensor([float('nan'), float('inf'), -0.0, 1e-40])
dst = pin(torch.empty(n, dtype=torch.bfloat16))
ref = src.to(torch.bfloat16)
L.spa3d_host_pack_bf16(src.data_ptr(), dst.data_ptr(), n, 4)
ok = (dst.view(torch.int16) == ref.view(torch.int16)) | (dst.isnan() & ref.isnan())
assert bool(ok.all())
cores = len(os.sched_getaffinity(0))
out = {"values": n, "cores": cores, "torch_threads": torch.get_num_threads()}


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


for th in (1, 2, 4, 6, 8, 12, 16, 24, 32):
    if th > 2 * cores:
        break
    out[f"pack_ms_t{th}"] = round(best(lambda: L.spa3d_host_pack_bf16(src.data_ptr(), dst.data_ptr(), n, th)), 3)
out["torch_copy_ms"] = round(best(lambda: dst.copy_(src)), 3)
if torch.cuda.is_available():
    dev = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    dev32 = torch.empty(n, dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()

    def h2d(buf, host):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s); buf.copy_(host, non_blocking=True); e1.record(s)
        return e0, e1
    e0, e1 = h2d(dev, dst); e1.synchronize(); out["h2d_bf16_alone_ms"] = round(e0.elapsed_time(e1), 3)
    e0, e1 = h2d(dev32, src); e1.synchronize(); out["h2d_f32_alone_ms"] = round(e0.elapsed_time(e1), 3)
    dst2 = torch.empty(n, dtype=torch.bfloat16).pin_memory()
    for th in (4, 8, 12, 16):
        if th > cores:
            break
        ts, cs = [], []
        for _ in range(4):
            e0, e1 = h2d(dev, dst2)
            t = time.perf_counter(); L.spa3d_host_pack_bf16(src.data_ptr(), dst.data_ptr(), n, th); ts.append(time.perf_counter() - t)
            e1.synchronize(); cs.append(e0.elapsed_time(e1))
        out[f"pack_ms_t{th}_under_h2d"] = round(min(ts) * 1e3, 3)
        out[f"h2d_bf16_ms_under_pack_t{th}"] = round(min(cs), 3)
print(json.dumps(out))
