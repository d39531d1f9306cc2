set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t17.log
grep -E "FAILED|passed|failed" gpurun_out/r2_t17.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke17.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke17.log; tail -2 gpurun_out/r2_smoke17.log
python bench.py > gpurun_out/r2_b17.json 2> gpurun_out/r2_b17.err; echo "bench rc=$?" >> gpurun_out/r2_b17.err; tail -2 gpurun_out/r2_b17.err
