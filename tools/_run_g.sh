cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo bench rc=$?; tail -3 gpurun_out/r2g_bench.err
python tools/show_bench.py gpurun_out/r2g_bench.json 2>&1 | head -60
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2g_bench.json") if l.startswith("{")][-1]
e=d["e2e"]; print("e2e", e["value"]/1e6, "M/s", e["ms_per_step"], "ms", e["h2d_GBps_per_gpu"], "GB/s")
for k,v in e["prepacked"].items(): print("  ", k, round(v["value"]/1e6,1), "M/s", round(v["ms_per_step"],3), "ms pack_ms", round(v["pack_ms_untimed"],1))
print(d["cpu_baseline"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-graphs-per-core 4000 > gpurun_out/r2g_ref.json 2>gpurun_out/r2g_ref.err; cat gpurun_out/r2g_ref.json | cut -c1-600
