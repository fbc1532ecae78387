cd $GRAFT_REPO_ROOT
SECONDS=0; python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo rc=$?
echo "bench wall ${SECONDS}s"; tail -3 gpurun_out/r2_bench_full.err
python tools/show_bench.py gpurun_out/r2_bench_full.json 2>&1 | tail -22
