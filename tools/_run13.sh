cd $GRAFT_REPO_ROOT
python - <<'PY' 2>&1 | tail -20
import importlib, torch, time, bench
spa = importlib.import_module("3dspa_code_b200")
dev = torch.device("cuda")
r = bench.run_trajan_leg(spa, dev, cpu=False); print("isolated", r["ms_per_clip"])
model = spa.TrackAutoEncoder3D()
variables = model.init(0, {"dino_features": 1, "depth_features": 1})
inputs, noise = bench.synth_clip(100, device=dev)
f = bench.run_fp32_leg(spa, spa.TrackAutoEncoder3D(), variables, inputs, noise, dev); print("fp32", f["ms_per_clip"])
r = bench.run_trajan_leg(spa, dev, cpu=False); print("after fp32 leg", r["ms_per_clip"])
l = bench.run_lifting_leg(spa, dev, cpu=True); print("lifting", l["ms"])
r = bench.run_trajan_leg(spa, dev, cpu=False); print("after lifting cpu leg", r["ms_per_clip"])
v = bench.cpu_train_rate(); print("cpu train", v[1])
r = bench.run_trajan_leg(spa, dev, cpu=False); print("after cpu_train_rate", r["ms_per_clip"])
PY
