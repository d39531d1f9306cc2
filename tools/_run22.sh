cd $GRAFT_REPO_ROOT
timeout 200 python tools/gemm_shapes.py --only fused --reps 7 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-train --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'eager', d['roofline']['ms_per_step_eager'], 'e2e', d['e2e']['ms_per_step'])"
