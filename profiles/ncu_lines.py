#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on.
   python profiles/ncu_lines.py gpurun_out/prof_X.ncu-rep [top_n]
Columns: stall samples (all), warp instructions executed, avg active threads, dominant stall reasons."""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, recs = "", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("Function Name",) or not r[0].strip().isdigit():
            continue
        d = dict(zip(hdr, r))
        # 'Source' appears twice (cuda, sass); csv dict keeps the last -> take by position
        src = r[1]
        try:
            smp = int(d.get("# Samples", "0") or 0)
            ins = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        thr = d.get("Avg. Threads Executed", "")
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        top3 = " ".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        recs.append((smp, ins, thr, cur_file, int(r[0]), src.strip()[:90], top3))
    tot_s = sum(x[0] for x in recs) or 1
    tot_i = sum(x[1] for x in recs) or 1
    print(f"# {path}: {tot_s} samples, {tot_i} warp instructions")
    print(f"{'samp%':>6s} {'inst%':>6s} {'thr':>4s}  location / source / top stalls")
    for smp, ins, thr, f, ln, src, st in sorted(recs, key=lambda x: -x[0])[:top]:
        print(f"{100 * smp / tot_s:6.2f} {100 * ins / tot_i:6.2f} {thr:>4s}  {f}:{ln}  {src}   [{st}]")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
