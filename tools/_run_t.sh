cd $GRAFT_REPO_ROOT
echo "=== new tests"; timeout 1200 python -m pytest tests -m gpu -x -q -k "walk_emit or collate_epoch_equals" 2>&1 | tail -15
echo "=== all gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== split"; timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3
