set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5.log
timeout 300 python tools/attn_bench.py > gpurun_out/r2_attn5.log 2>&1
tail -5 gpurun_out/r2_t5.log; cat gpurun_out/r2_attn5.log
