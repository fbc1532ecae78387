cd $GRAFT_REPO_ROOT
for v in pf0 pf4 pf8; do echo "=== $v"; GTOK_LIB_PATH=$PWD/build_ab/libgtok_$v.so timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3; done
timeout 900 python -m pytest tests -m gpu -x -q -k "sent or guard" 2>&1 | tail -2
