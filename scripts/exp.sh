cd "$GRAFT_REPO_ROOT"
echo "### tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "### rolling"; timeout 300 python scripts/rolling_probe.py 4 | head -11
echo "### bench"; timeout 300 bash scripts/ab.sh lib/libcmpc_b200.so
