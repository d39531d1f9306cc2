set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t8.log
tail -5 gpurun_out/r2_t8.log
SPA3D_GEMM_EW16=0 timeout 300 python tools/gemm_shapes.py --only mlp1 --reps 7 > gpurun_out/r2_gs8_ew8.log 2>&1
timeout 300 python tools/gemm_shapes.py --only mlp1 --reps 7 > gpurun_out/r2_gs8_ew16.log 2>&1
cat gpurun_out/r2_gs8_ew8.log gpurun_out/r2_gs8_ew16.log
timeout 300 python tools/gemm_shapes.py --reps 5 > gpurun_out/r2_gs8_all.log 2>&1
