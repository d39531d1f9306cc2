set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_model.py tests/test_torch_ops.py -m gpu -q -x -k "stream or torch_ops or dispatch or opcheck or composed" > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; echo "bench rc=$?" >> gpurun_out/r2_b4.err
tail -3 gpurun_out/r2_t4.log
