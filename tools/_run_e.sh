cd $GRAFT_REPO_ROOT
export GTOK_ZINC_TMA=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_pyg.py tests/test_dropin_gpu.py -m gpu -x -q -k "zinc or pyg or ingest or dropin or fused_entry" 2>&1 | tail -3
python tools/probe_zinc.py > gpurun_out/r2e_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:zinc_tpm_kernel -s 27 -c 1 -f -o gpurun_out/r2e_tma1 python tools/probe_zinc.py > gpurun_out/r2e_ncu1.log 2>&1
tail -3 gpurun_out/r2e_plain1.log; tail -3 gpurun_out/r2e_ncu1.log
