#!/usr/bin/env python
"""Pretty-print a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_X.json"""
import json
import sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
print("value", round(d["value"] / 1e6, 1), "Mgraphs/s  ms/step", round(d["ms_per_step"], 4), " e2e", round(d["e2e"]["value"] / 1e6, 1),
      "Mgraphs/s  launches", d["gpu_launches"])
for k, v in list(d["kernels"].items()) + list(d.get("kernels_two_step", {}).items()):
    print(f"  {k:18s} {v['ms']:.4f} ms  {v['GBps']:7.0f} GB/s  frac {v['frac']:.3f}")
print("  roofline", d["roofline"])
print("  step_roofline", round(d["step_roofline"]["frac"], 3), d["clocks"])
print("  cpu_baseline", d.get("cpu_baseline"))
for k, v in (d.get("extras") or {}).items():
    if "points" in v:
        print(f"  extra {k}: {v['workload']}")
        for p in v["points"]:
            print(f"      N={p['nodes_per_graph']:5d} G={p['graphs_per_gpu']:7d}  ibtt_collate {p['ibtt_collate_ms']:.3f} ms ({p['ibtt_collate_frac']:.2f})  "
                  f"sent {p['sent_ms']:.3f} ms ({p['sent_frac']:.3f})  sent epoch {p['sent_epoch_ms']:.3f} ms ({p['sent_epoch_frac']:.3f}, fused={p['sent_epoch_fused']})")
        continue
    if "ms_per_step" not in v:
        print(f"  extra {k}: " + "  ".join(f"{a}={b:.4g}" if isinstance(b, float) else f"{a}={b}" for a, b in v.items() if a != "workload"))
        continue
    print(f"  extra {k}: {v.get('ms_per_step'):.3f} ms/step  {v.get('graphs_per_s', 0) / 1e6:.0f} Mgraphs/s  frac {v.get('frac')}" +
          (f"  | +collate {v['collate_ms']:.3f} ms (frac {v['collate_frac']:.2f}) -> epoch {v['epoch_graphs_per_s'] / 1e6:.0f} Mgraphs/s" if 'collate_ms' in v else ""))
    for kk, vv in (v.get("kernels") or {}).items():
        print(f"      {kk:18s} {vv['ms']:.4f} ms  {vv['GBps']:7.0f} GB/s  frac {vv['frac']:.3f}")
