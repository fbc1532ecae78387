cd $GRAFT_REPO_ROOT
for cfg in "0 1" "1 0" "1 1" "2 1"; do set -- $cfg
  echo "=== TMA=$1 L2HINT=$2"; GTOK_ZINC_TMA=$1 GTOK_ZINC_L2HINT=$2 timeout 300 python tools/probe_zinc.py 2>&1 | tail -2
done
for pf in 96 160 224 384; do echo "=== TMA=1 hint=1 PF=$pf";  GTOK_ZINC_PREFETCH=$pf GTOK_ZINC_TMA=1 timeout 300 python tools/probe_zinc.py 2>&1 | tail -2; done
GTOK_ZINC_TMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_pyg.py -m gpu -x -q -k "zinc or pyg" 2>&1 | tail -3
