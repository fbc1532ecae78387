cd $GRAFT_REPO_ROOT
echo "=== base"; GTOK_LIB_PATH=$PWD/build_ab/libgtok_base.so timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3
echo "=== staged32"; timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q -k "sent or guard" 2>&1 | tail -2
