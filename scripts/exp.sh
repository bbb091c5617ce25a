cd "$GRAFT_REPO_ROOT"
timeout 100 bash scripts/ab.sh lib/libcmpc_b200.so | tail -1
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
