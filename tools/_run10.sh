set -x
cd $GRAFT_REPO_ROOT
NCUL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCUL -c 4000 --log-file gpurun_out/r2_launches_infer.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-train > gpurun_out/r2_ncu_l1.log 2>&1
$NCUL -c 5000 --log-file gpurun_out/r2_launches_train.csv python tools/train_bench.py --clips 1 --steps 1 > gpurun_out/r2_ncu_l2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r2_gemm_mlp1_ew16 -f python tools/gemm_shapes.py --only itt.mlp1 --reps 1 > gpurun_out/r2_ncu_f.log 2>&1
$NCU -k regex:attn_fwd_tc -s 4 -c 2 -o gpurun_out/r2_attn_cross_fwd -f python tools/attn_bench.py > gpurun_out/r2_ncu_g.log 2>&1
ls -la gpurun_out | grep -E "r2_launches|ew16|cross"
