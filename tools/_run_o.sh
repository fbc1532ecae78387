cd $GRAFT_REPO_ROOT
echo "=== SENT tests"; timeout 1200 python -m pytest tests -m gpu -x -q -k "sent or shim or agtt or pin or smoke" 2>&1 | tail -3
echo "=== probe_sent"; timeout 300 python tools/probe_sent.py 2>&1 | tail -3
echo "=== fused"; timeout 600 python tools/probe_sent_collate.py 2>&1 | tail -4
