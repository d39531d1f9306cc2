cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4
SPA3D_ATTN_NP=1 timeout 200 python tools/attn_bench.py 2>&1 | tail -3
timeout 200 python tools/attn_bench.py 2>&1 | tail -3
