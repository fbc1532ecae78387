cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_guard_bands.py -m gpu -x -q 2>&1 | tail -15
