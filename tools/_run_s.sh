cd $GRAFT_REPO_ROOT
for v in A B0 B1 B2 B3 B4 B5; do echo "=== $v"; GTOK_LIB_PATH=$PWD/build_ab/libgtok_$v.so timeout 300 python tools/probe_sent_split.py 2>&1 | tail -3; done
