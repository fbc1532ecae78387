cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_N2.json 2> gpurun_out/r2_bench_N2.err; echo rc=$?
tail -2 gpurun_out/r2_bench_N2.err
python tools/show_bench.py gpurun_out/r2_bench_N2.json 2>&1 | head -12
python tools/show_bench.py gpurun_out/r2_bench_N2.json 2>&1 | grep -i "sharded"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo rc=$?; cat gpurun_out/r2_bench_ref.json | cut -c1-600
