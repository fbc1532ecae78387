cd $GRAFT_REPO_ROOT
echo "=== tests"; timeout 2400 python -m pytest tests -m gpu -x -q -k "walk_emit or collate_epoch_equals or sweep or guard or mid or append_eos" 2>&1 | tail -3
for n in 128 256; do timeout 300 python tools/probe_sent_split.py 32000000 sparse $n 2>&1 | tail -1; done
echo "--- lists"; for n in 128 256; do GTOK_SENT_LISTS=1 timeout 300 python tools/probe_sent_split.py 32000000 sparse $n 2>&1 | tail -1; done
timeout 600 python tools/fuzz_gpu.py 40 21 sent 2>&1 | tail -2
