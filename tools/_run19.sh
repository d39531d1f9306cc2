cd $GRAFT_REPO_ROOT
SPA3D_BENCH_PROFILE=1 python bench.py --steps 3 --warmup 3 --train-batch 4 --train-steps 1 --no-cpu 2>&1 >/dev/null | grep -A40 "\[profile\]" | cut -c1-170
