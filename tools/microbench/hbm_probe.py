"""HBM bandwidth by access mix on this GPU (torch kernels; sizes far above the 126 MB L2): pure write (fill), pure read (sum),
copy (1 read : 1 write) and a 1 : 3.3 read : write mix like the K0 lifting kernel's (0.74 GB read, 2.47 GB written)."""
import json

import torch

dev = torch.device("cuda")
n = 1 << 29   # 2 GiB of float32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
t = timed(lambda: a.fill_(1.0)); out["write_only_gbs"] = round(4 * n / t / 1e9, 1)
t = timed(lambda: a.sum()); out["read_only_gbs"] = round(4 * n / t / 1e9, 1)
t = timed(lambda: b.copy_(a)); out["copy_gbs"] = round(8 * n / t / 1e9, 1)
q = int(n * 0.3)
t = timed(lambda: (b.fill_(2.0), torch.add(a[:q], 1.0, out=b[:q]))); out["fill_plus_0.3_copy_gbs"] = round((4 * n + 8 * q) / t / 1e9, 1)
print(json.dumps(out))
