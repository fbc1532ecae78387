cd $GRAFT_REPO_ROOT
timeout 1500 python tools/fuzz_gpu.py 60 11 sent > gpurun_out/r2_fuzz_sent.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_fuzz_sent.log; grep -c ok gpurun_out/r2_fuzz_sent.log; grep MISMATCH gpurun_out/r2_fuzz_sent.log | head
timeout 900 python tools/fuzz_gpu.py 8 12 > gpurun_out/r2_fuzz_all.log 2>&1; echo rc=$?; tail -2 gpurun_out/r2_fuzz_all.log
