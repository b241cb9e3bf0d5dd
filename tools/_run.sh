cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3
echo "--- MC on"; timeout 120 python tools/mlp_timing.py
echo "--- MC off"; V2S_MLP_MC=0 timeout 120 python tools/mlp_timing.py
V2S_GEMM_DEBUG=1 timeout 120 python tools/mlp_timing.py 2>&1 | tail -12
timeout 300 python tools/mlp_stress.py 2>&1 | tail -4
