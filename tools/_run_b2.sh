cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,pstate,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo rc=$?
python tools/show_bench.py gpurun_out/r2_bench_b.json 2>&1 | grep -A6 "extra ibtt_graph_token"
python tools/probe_ibtt_sent.py 2>&1 | tail -8
