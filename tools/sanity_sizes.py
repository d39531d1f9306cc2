#!/usr/bin/env python
"""Development check at the other BASELINE sizes: cfg5 clip shape (S=4096, Q=1024), batched inference equals
per-clip inference bit for bit, TRAJAN 2D at cfg1 size, training with micro-batches of 2."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
spa = importlib.import_module("3dspa_code_b200")
te = importlib.import_module("3dspa_code_b200.train_engine")
dev = torch.device("cuda")
model = spa.TrackAutoEncoder3D()
variables = model.init(0, {"dino_features": 1, "depth_features": 1})

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

# cfg5 clip shape
inp, noise = bench.synth_clip(7, device=dev, s=4096, q=1024)
res = model.apply(variables, inp, noise=noise, precision="bf16")
assert torch.isfinite(res.tracks).all() and res.tracks.shape == (1, 1024, 150, 3)
print("cfg5 clip (S=4096,Q=1024): %.2f ms" % timeit(lambda: model.apply(variables, inp, noise=noise, precision="bf16")))
# batch of 2 == two singles
a, na = bench.synth_clip(1, device=dev, s=512, q=128)
b, nb = bench.synth_clip(2, device=dev, s=512, q=128)
both = {k: torch.cat([a[k], b[k]], 0) for k in a}
r2 = model.apply(variables, both, noise=torch.cat([na, nb], 0), precision="bf16")
ra = model.apply(variables, a, noise=na, precision="bf16")
rb = model.apply(variables, b, noise=nb, precision="bf16")
assert torch.equal(r2.tracks[0], ra.tracks[0]) and torch.equal(r2.tracks[1], rb.tracks[0]), "batched != per clip"
print("batched inference equals per-clip inference bit for bit")
# TRAJAN cfg1 shape on the GPU
tj = spa.TrackAutoEncoder()
tv = tj.init(0, None)
rs = np.random.RandomState(0)
ti = {"support_tracks": rs.uniform(0, 1, (1, 2048, 150, 2)).astype(np.float32),
      "support_tracks_visible": (rs.uniform(size=(1, 2048, 150, 1)) < 0.9).astype(np.float32),
      "query_points": np.concatenate([rs.randint(0, 150, (1, 512, 1)).astype(np.float32), rs.uniform(0, 1, (1, 512, 2)).astype(np.float32)], -1),
      "boundary_frame": np.array([150], np.int32)}
ti = {k: torch.as_tensor(v).to(dev) for k, v in ti.items()}
tn = torch.rand(1, 128, 64, device=dev)
rt = tj.apply(tv, ti, noise=tn, precision="bf16")
assert torch.isfinite(rt.tracks).all() and rt.tracks.shape == (1, 512, 150, 2)
print("TRAJAN 2D cfg1 shape on GPU: %.2f ms" % timeit(lambda: tj.apply(tv, ti, noise=tn, precision="bf16")))
# training, micro-batch 2
tb = importlib.import_module("tools.train_bench") if False else None
g = torch.Generator(device=dev).manual_seed(3)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from train_bench import synth_batch
batch, nz = synth_batch(2, 5, dev, 1024, 256)
tr = te.Trainer(model, variables["params"], precision="bf16", micro_batch=2)
l1 = tr.train_step(batch, nz)
tr1 = te.Trainer(model, variables["params"], precision="bf16", micro_batch=1)
l2 = tr1.train_step(batch, nz)
print("train micro=2 vs micro=1: loss", l1["total_loss"], l2["total_loss"], "grad norm", l1["grad_norm"], l2["grad_norm"])
assert abs(l1["total_loss"] - l2["total_loss"]) < 2e-3 * abs(l2["total_loss"]) and abs(l1["grad_norm"] - l2["grad_norm"]) < 2e-2 * l2["grad_norm"]
print("ok")
