cd "$GRAFT_REPO_ROOT"
bash scripts/ab.sh lib/variants/lib_base4.so lib/libcmpc_b200.so lib/variants/lib_base4.so lib/libcmpc_b200.so
timeout 600 python -m pytest tests -x -q -m gpu -k "golden or parity" 2>&1 | tail -3
