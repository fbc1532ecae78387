cd $GRAFT_REPO_ROOT
echo "=== SENT tests"; timeout 1200 python -m pytest tests -m gpu -x -q -k "sent or shim or agtt or pin or smoke" 2>&1 | tail -3
for pf in 0 1; do echo "=== probe_sent PF=$pf"; GTOK_WALK_PF=$pf timeout 300 python tools/probe_sent.py 2>&1 | tail -3; done
echo "=== fused"; timeout 600 python tools/probe_sent_collate.py 2>&1 | tail -4
