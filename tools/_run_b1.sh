cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo rc=$?
tail -3 gpurun_out/r2_bench.err
python tools/show_bench.py gpurun_out/r2_bench.json 2>&1 | head -80
