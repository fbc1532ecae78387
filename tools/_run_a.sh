set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -2; nproc; lscpu | grep -E "Model name|Socket|NUMA" 
for t in 1 2 0; do
  echo "=== TMA=$t tests"; GTOK_ZINC_TMA=$t timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_e2e_pyg.py tests/test_dropin_gpu.py -m gpu -x -q -k "zinc or pyg or ingest or dropin or fused_entry" 2>&1 | tail -5
done
for t in 0 1 2; do echo "=== probe TMA=$t"; GTOK_ZINC_TMA=$t timeout 300 python tools/probe_zinc.py 2>&1 | tail -4; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --cpu-graphs-per-core 5000 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err
python tools/show_bench.py gpurun_out/r2a_bench.json 2>&1 | head -40
