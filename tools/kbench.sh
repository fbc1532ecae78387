#!/bin/bash
# quick GPU iteration helper: parity tests for one family + per-kernel timings from bench.py
python -m pytest tests -m gpu -q -x -k "${1:-zinc}" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras ${2:-} > /tmp/b.json 2>/tmp/b.err || tail -5 /tmp/b.err
python - <<PY
import json
d=json.load(open("/tmp/b.json"))
print("value",round(d["value"]/1e6,1),"Mgraphs/s  ms/step",round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"]/1e6,1))
for k,v in list(d["kernels"].items())+list(d.get("kernels_two_step",{}).items()): print(f"  {k:18s} {v['ms']:.4f} ms  {v['GBps']:7.0f} GB/s  frac {v['frac']:.3f}")
print("  step_roofline", round(d["step_roofline"]["frac"],3), d["clocks"])
PY
