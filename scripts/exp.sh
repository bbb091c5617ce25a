cd "$GRAFT_REPO_ROOT"
echo "### A/B skip"; bash scripts/ab.sh lib/variants/lib_noskip.so lib/libcmpc_b200.so lib/variants/lib_noskip.so lib/libcmpc_b200.so
echo "### profile"; CMPC_LIB=$PWD/lib/variants/lib_prof.so python scripts/phase_profile.py 20 4096 | tail -1
echo "### tests"; timeout 1500 python -m pytest tests -x -q -m gpu -k "golden or parity" 2>&1 | tail -4
