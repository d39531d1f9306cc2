set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
(numactl -H; lscpu | head -30; free -g) > gpurun_out/r2_host8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 tools/h2d_probe.py > gpurun_out/r2_h2d8_unbound.json 2> gpurun_out/r2_h2d8.err
SPA3D_BIND=1 $TR --master-port 29522 tools/h2d_probe.py > gpurun_out/r2_h2d8_bound.json 2>> gpurun_out/r2_h2d8.err
cat gpurun_out/r2_h2d8_unbound.json gpurun_out/r2_h2d8_bound.json
$TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_b6_n8.json 2> gpurun_out/r2_b6_n8.err; echo "bench rc=$?" >> gpurun_out/r2_b6_n8.err
tail -2 gpurun_out/r2_b6_n8.err
