cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 --train-batch 4 --train-steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('trajan', d['trajan']['ms_per_clip'], 'fp32', d['fp32_path']['ms_per_clip'], 'train', d['train']['ms_per_step'], 'sweep', d['sweep']['ms_per_clip'])"
