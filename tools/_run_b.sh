cd $GRAFT_REPO_ROOT
./tools/micro/wbw.bin
for t in 1 2; do
  echo "=== TMA=$t tests"; GTOK_ZINC_TMA=$t timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_pyg.py tests/test_dropin_gpu.py -m gpu -x -q -k "zinc or pyg or ingest or dropin or fused_entry" 2>&1 | tail -3
done
