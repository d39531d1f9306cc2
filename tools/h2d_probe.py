#!/usr/bin/env python
"""Development probe (launched with torchrun, one rank per GPU): host->device bandwidth per rank when all ranks upload at once,
for ordinary pinned memory and write-combined pinned memory, with and without binding the rank to its GPU's NUMA-local CPUs
(SPA3D_BIND=1).  Prints one JSON line per rank 0."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bound = bench.bind_to_gpu_numa_node(local) if os.environ.get("SPA3D_BIND") == "1" else None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 1 << 30
normal = torch.empty(N, dtype=torch.uint8).pin_memory()
normal.fill_(1)
cudart = ctypes.CDLL("libcudart.so.12")
ptr = ctypes.c_void_p()
rc = cudart.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(N), ctypes.c_uint(0x04))   # cudaHostAllocWriteCombined
wc = None
if rc == 0:
    wc = torch.frombuffer((ctypes.c_char * N).from_address(ptr.value), dtype=torch.uint8)
    wc.fill_(1)
dst = torch.empty(N, dtype=torch.uint8, device=dev)


def run(src, reps=8):
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([N * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(x.item()), 1) for x in out]
    return [round(float(t.item()), 1)]


res = {"world": world, "bound_cpus": bound, "cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
       "pinned_gbs_per_rank": run(normal), "wc_pinned": wc is not None and wc.is_pinned()}
if wc is not None:
    res["wc_gbs_per_rank"] = run(wc)
if rank == 0:
    res["aggregate_pinned"] = round(sum(res["pinned_gbs_per_rank"]), 1)
    if "wc_gbs_per_rank" in res:
        res["aggregate_wc"] = round(sum(res["wc_gbs_per_rank"]), 1)
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
