cd "$GRAFT_REPO_ROOT"
bash scripts/ab.sh lib/variants/lib_prev.so lib/libcmpc_b200.so lib/variants/lib_z64.so
