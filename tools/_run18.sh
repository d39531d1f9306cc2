cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -3
SPA3D_GEMM_EWD=0 timeout 300 python tools/gemm_shapes.py --reps 7 > gpurun_out/r2_gs18_ewd0.log 2>&1
timeout 300 python tools/gemm_shapes.py --reps 7 > gpurun_out/r2_gs18_ewd1.log 2>&1
SPA3D_GEMM_EWD=0 timeout 300 python tools/gemm_shapes.py --reps 7 > gpurun_out/r2_gs18_ewd0b.log 2>&1
python - <<'PY'
import json
def load(f): return {json.loads(l)['name']: json.loads(l) for l in open(f) if l.startswith('{"name')}
a,b,c=load('gpurun_out/r2_gs18_ewd0.log'),load('gpurun_out/r2_gs18_ewd1.log'),load('gpurun_out/r2_gs18_ewd0b.log')
for k in a:
    print(f"{k:14s} {a[k]['kind']:6s} ewd0 {a[k]['ours_ms']:.4f}  ewd1 {b[k]['ours_ms']:.4f}  ewd0-again {c[k]['ours_ms']:.4f}")
PY
