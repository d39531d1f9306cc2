set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2_topo2.txt 2>&1
python -m pytest tests/test_gpu_dp.py tests/test_gpu_parity_full.py::test_bf16_gradients_at_reference_widths -m gpu -q -s > gpurun_out/r2_t3_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3_dp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_b3_n2.json 2> gpurun_out/r2_b3_n2.err; echo "bench rc=$?" >> gpurun_out/r2_b3_n2.err
tail -3 gpurun_out/r2_t3_dp.log
