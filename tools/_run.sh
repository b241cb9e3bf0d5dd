cd $GRAFT_REPO_ROOT
V2S_GEMM_DEBUG=1 timeout 120 python tools/attn_timing.py 2>&1 | tail -25
