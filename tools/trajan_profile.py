#!/usr/bin/env python
"""Development tool: CUDA-event time per C-ABI entry point over one eager cfg1 (TRAJAN 2D) forward."""
import collections, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spa = importlib.import_module("3dspa_code_b200")
ops = spa.ops
dev = torch.device("cuda")
T, S, Q = 150, 2048, 512
rs = np.random.RandomState(11)
inp = {"support_tracks": rs.uniform(0, 1, (1, S, T, 2)).astype(np.float32),
       "support_tracks_visible": (rs.uniform(size=(1, S, T, 1)) < 0.9).astype(np.float32),
       "query_points": np.concatenate([rs.randint(0, T, (1, Q, 1)).astype(np.float32), rs.uniform(0, 1, (1, Q, 2)).astype(np.float32)], -1),
       "boundary_frame": np.array([T], np.int32)}
noise = torch.from_numpy(rs.uniform(size=(1, 128, 64)).astype(np.float32)).to(dev)
model = spa.TrackAutoEncoder()
variables = model.init(0, inp)
dev_inp = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
for _ in range(2):
    model.apply(variables, dev_inp, noise=noise, precision="bf16")
prof = collections.defaultdict(list)
orig = ops._call
def timed(name, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig(name, *a); e.record()
    prof[name].append((s, e, a))
ops._call = timed
ops.stats(reset=True)
model.apply(variables, dev_inp, noise=noise, precision="bf16")
torch.cuda.synchronize()
tot = {k: sum(s.elapsed_time(e) for s, e, _ in v) for k, v in prof.items()}
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:8.3f} ms  n={len(prof[k]):4d}  {k}")
print(f"{sum(tot.values()):8.3f} ms total", {k: v for k, v in ops.stats().items() if v})
big = sorted(((s.elapsed_time(e), k, [x for x in a if isinstance(x, int)][:12]) for k, v in prof.items() for s, e, a in v), key=lambda t: -t[0])[:10]
for b in big:
    print(b)
