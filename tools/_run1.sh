set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -s > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
python tools/gemm_shapes.py --reps 5 > gpurun_out/r2_gs1.log 2>&1
python tools/attn_bench.py > gpurun_out/r2_attn1.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r2_gemm_mlp1 -f python tools/gemm_shapes.py --only itt.mlp1 --reps 1 > gpurun_out/r2_ncu_a.log 2>&1
$NCU -k regex:attn_bwd_tc -c 1 -o gpurun_out/r2_attn_bwd -f python tools/attn_bench.py > gpurun_out/r2_ncu_b.log 2>&1
$NCU -k regex:attn_fwd_tc -c 1 -o gpurun_out/r2_attn_fwd -f python tools/attn_bench.py > gpurun_out/r2_ncu_c.log 2>&1
$NCU -k regex:lift_sample -s 2 -c 1 -o gpurun_out/r2_lift -f python tools/lift_bench.py > gpurun_out/r2_ncu_d.log 2>&1
$NCU -k regex:embed_fused -s 2 -c 1 -o gpurun_out/r2_embed -f python tools/embed_bench.py > gpurun_out/r2_ncu_e.log 2>&1
ls -la gpurun_out/ | tail -20
