cd $GRAFT_REPO_ROOT
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "eval1024" -s 2>&1 | grep -E "finetune B=128|passed|failed|assert |Error"; done
