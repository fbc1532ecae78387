cd $GRAFT_REPO_ROOT
export GTOK_ZINC_TMA=1
echo "=== no group bounds"; GTOK_GROUP_BOUNDS=0 timeout 300 python tools/probe_zinc.py 2>&1 | tail -2
for pf in 160 288 416; do echo "=== group bounds PF=$pf"; GTOK_ZINC_PREFETCH=$pf timeout 300 python tools/probe_zinc.py 2>&1 | tail -3; done
echo "=== sleep256"; GTOK_LIB_PATH=$PWD/glearning-benchmark_b200/gtok/libgtok_sleep256.so timeout 300 python tools/probe_zinc.py 2>&1 | tail -2
echo "=== TMA=2 group bounds"; GTOK_ZINC_TMA=2 timeout 300 python tools/probe_zinc.py 2>&1 | tail -2
echo "=== TMA=0 group bounds"; GTOK_ZINC_TMA=0 timeout 300 python tools/probe_zinc.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_pyg.py tests/test_dropin_gpu.py -m gpu -x -q -k "zinc or pyg or ingest or dropin or fused_entry" 2>&1 | tail -3
