set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t9.log
grep -E "FAILED|passed|failed" gpurun_out/r2_t9.log
SPA3D_GEMM_RMS12=0 SPA3D_GEMM_EW16=0 timeout 300 python tools/gemm_shapes.py --reps 7 > gpurun_out/r2_gs9_old.log 2>&1
timeout 300 python tools/gemm_shapes.py --reps 7 > gpurun_out/r2_gs9_new.log 2>&1
python bench.py --steps 10 --warmup 3 --no-train --no-cpu > gpurun_out/r2_b9.json 2> gpurun_out/r2_b9.err
