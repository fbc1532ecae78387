cd $GRAFT_REPO_ROOT
echo "=== SENT tests (new walk kernel default)"; timeout 1200 python -m pytest tests -m gpu -x -q -k "sent or shim or agtt or pin or smoke" 2>&1 | tail -8
for v in 1 2; do echo "=== probe_sent SMALL=$v"; GTOK_SENT_SMALL=$v timeout 300 python tools/probe_sent.py 2>&1 | tail -3; done
echo "=== ZINC balanced"; GTOK_SENT_SMALL=2 timeout 300 python tools/probe_sent.py 1000000 2048 zincb 2>&1 | tail -2
