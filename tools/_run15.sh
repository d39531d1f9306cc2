cd $GRAFT_REPO_ROOT
SPA3D_BENCH_DEBUG=1 python bench.py --steps 3 --warmup 3 --train-batch 4 --train-steps 1 --no-cpu 2> gpurun_out/dbg.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('final trajan leg', d['trajan']['ms_per_clip'])"
grep debug gpurun_out/dbg.err
python bench.py --steps 3 --warmup 3 --train-batch 4 --train-steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('no-debug trajan leg', d['trajan']['ms_per_clip'])"
