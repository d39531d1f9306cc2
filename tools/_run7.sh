set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -5 gpurun_out/r2_t7.log
timeout 300 python tools/fp32_bench.py > gpurun_out/r2_fp32.json 2> gpurun_out/r2_fp32.err; cat gpurun_out/r2_fp32.json; tail -3 gpurun_out/r2_fp32.err
