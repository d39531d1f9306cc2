set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -s -x > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo "bench rc=$?" >> gpurun_out/r2_b2.err
tail -3 gpurun_out/r2_t2.log
