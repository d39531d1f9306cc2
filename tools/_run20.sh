set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t20.log
grep -E "FAILED|passed|failed" gpurun_out/r2_t20.log
python bench.py --steps 10 --warmup 3 --no-train --no-cpu > gpurun_out/r2_b20.json 2> gpurun_out/r2_b20.err; tail -2 gpurun_out/r2_b20.err
python tools/trajan_profile.py 2>&1 | tail -14 | head -3
