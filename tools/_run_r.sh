cd $GRAFT_REPO_ROOT
echo "=== A (first two-kernel build)"; GTOK_LIB_PATH=$PWD/build_ab/libgtok_A.so timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3
echo "=== current"; timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3
echo "=== current again"; timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3
