cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_infonce.py tests/test_abi.py -q -s 2>&1 | grep -E "infonce|passed|failed|Error|assert" | head -40
