cd $GRAFT_REPO_ROOT
echo "=== tests"; timeout 2400 python -m pytest tests -m gpu -x -q -k "walk_emit or collate_epoch_equals or sweep or guard or mid" 2>&1 | tail -3
for n in 128 256; do timeout 300 python tools/probe_sent_split.py 32000000 sparse $n 2>&1 | tail -1; done
