cd $GRAFT_REPO_ROOT
N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_N$N.json 2> gpurun_out/r2_bench_N$N.err; echo rc=$?
tail -2 gpurun_out/r2_bench_N$N.err
python tools/show_bench.py gpurun_out/r2_bench_N$N.json 2>&1 | grep -v "^      \(ibtt\|scan\|batch\|pad\|zinc\)" | head -30
