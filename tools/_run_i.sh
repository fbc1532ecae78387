cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -30
lscpu | grep -E "Socket|NUMA|Model name|^CPU\(s\)"
nproc
for n in 2 4 8; do
  echo "=== probe_h2d N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/probe_h2d.py 512 2>/dev/null | tail -1 > gpurun_out/r2i_h2d_$n.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2i_h2d_$n.json"))
for ph in ("default_affinity","bound_to_gpu_node"):
    p=d[ph]; print(ph, "agg", round(p["aggregate_concurrent_GBps"],1), "conc", [round(x,1) for x in p["concurrent_GBps_per_rank"]], "alone", [round(x,1) for x in p["alone_GBps_per_rank"]])
print([ (t["gpu"], t["gpu_numa_node"], t["node_cpulist"], len(t["affinity"] or [])) for t in d["default_affinity"]["topology"]])
PY
done
for n in 8 4; do
  echo "=== bench N=$n"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2i_bench_$n.json 2> gpurun_out/r2i_bench_$n.err; echo rc=$?; tail -2 gpurun_out/r2i_bench_$n.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2i_bench_$n.json") if l.startswith("{")][-1]
e=d["e2e"]; print("N", d["n_gpus"], "value", round(d["value"]/1e6,1), "e2e", round(e["value"]/1e6,1), "M/s", round(e["ms_per_step"],2), "ms", round(e["h2d_GBps_per_gpu"],1), "GB/s/gpu", d.get("host_placement"))
for k,v in e["prepacked"].items(): print("  ", k, round(v["value"]/1e6,1), "M/s", round(v["ms_per_step"],3), "ms")
print(d.get("extras",{}).get("sent_zinc_10M_sharded"))
PY
done
