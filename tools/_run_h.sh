cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_util_kernels.py tests/test_dropin_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "util or compact or gather or balance or shim or sent or dropin" 2>&1 | tail -5
python tools/sent_only.py > gpurun_out/r2h_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sent_small_kernel -s 1 -c 1 -f -o gpurun_out/r2h_sent_er python tools/sent_only.py > gpurun_out/r2h_ncu.log 2>&1
tail -2 gpurun_out/r2h_plain.log gpurun_out/r2h_ncu.log
