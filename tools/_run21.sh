cd $GRAFT_REPO_ROOT
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mlp_fused" 2>&1 | tail -15
