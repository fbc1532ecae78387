#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant
kernels, each entry stamped with the sha256 of the kernel's source file at capture time - bench.py reports
`roofline.traffic` only while the source still hashes to that value (a changed kernel gets null + a note, never a stale figure).
   python tools/update_traffic.py"""
import csv
import hashlib
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "glearning-benchmark_b200", "csrc")
# bench key -> (report, kernel-name substrings to sum, source files)
ENTRIES = {
    "zinc_collate": ("gpurun_out/r2_zinc_collate.ncu-rep", ["zinc_tpm_kernel"], ["gtok_zinc_tpm.cu"], "profiles/r2_zinc_collate_full.txt"),
    "sent_graph_token": ("gpurun_out/r2_sent_er.ncu-rep", ["sent_walk_kernel", "sent_emit_kernel"], ["gtok_sent_walk.cu"],
                         "profiles/r2_sent_er_full.txt"),
    "sent_zinc": ("gpurun_out/r2_sent_zinc.ncu-rep", ["sent_walk_kernel", "sent_emit_kernel"], ["gtok_sent_walk.cu"],
                  "profiles/r2_sent_zinc_full.txt"),
}


def sha(files):
    h = hashlib.sha256()
    for f in files:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]


def dram_bytes(rep, names):
    out = subprocess.run(["ncu", "-i", os.path.join(ROOT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    tot, seen = 0.0, set()
    for r in rows[2:]:
        nm = next((n for n in names if n in r[ki]), None)
        if nm is None or nm in seen:
            continue
        seen.add(nm)
        for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(col)
            tot += float(r[i]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return int(tot)


path = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
new = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (walk + emission kernels summed for SENT) from one `ncu --set "
                "full` capture of the 1M-graph workload; source_sha16 = sha256 of the kernel source at capture time, bench.py "
                "drops the figure when the source changed; algorithmic bytes are in bench.py's roofline.algo_bytes_per_launch"}
for key, (rep, names, files, txt) in ENTRIES.items():
    if os.path.exists(os.path.join(ROOT, rep)):
        new[key] = {"dram_bytes_per_launch": dram_bytes(rep, names), "source": txt, "source_files": files, "source_sha16": sha(files)}
    elif key in old:
        new[key] = old[key]
for k, v in old.items():
    if k not in new:
        new[k] = v
json.dump(new, open(path, "w"), indent=1)
print(json.dumps(new, indent=1))
