cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "collate_epoch_equals or guard or loader or agtt or smoke" 2>&1 | tail -2
timeout 600 python tools/probe_sent_collate.py 2>&1 | tail -3
