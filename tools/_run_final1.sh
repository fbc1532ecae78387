cd $GRAFT_REPO_ROOT
echo "=== all gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== fuzz"; timeout 1500 python tools/fuzz_gpu.py 12 31 2>&1 | tail -2
