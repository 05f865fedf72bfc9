#!/usr/bin/env python
"""Per-source-line summary of an ncu report (CPU only): warp instructions executed per warp, share of the stall samples and
the top stall reasons, from `ncu -i REP --page source --csv --print-source cuda,sass` (needs -lineinfo at compile time).

    python scripts/ncu_lines.py gpurun_out/r2_step_v6.ncu-rep [n_warps] [kernel-name regex] > profiles/..._source_stalls.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_warps = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
kfilter = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kfilter, capture_output=True, text=True).stdout
# the output holds one table per (file, function); take the first function's table (launch 0)
lines = txt.splitlines()
start = next(i for i, ln in enumerate(lines) if ln.startswith('"Line No"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"File Path"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
i_line, i_src = 0, 1
i_inst, i_samp = col["Instructions Executed"], col["# Samples"]
stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
cur_line, cur_src = None, ""
for r in rows[1:]:
    if r[i_line] != "":
        cur_line, cur_src = int(r[i_line]), r[i_src]
    if cur_line is None:
        continue
    a = agg.setdefault(cur_line, dict(src=cur_src, inst=0.0, samp=0.0, stalls={}))
    try:
        a["inst"] += float(r[i_inst] or 0)
        a["samp"] += float(r[i_samp] or 0)
    except ValueError:
        continue
    for h, i in stall_cols:
        try:
            v = float(r[i] or 0)
        except ValueError:
            v = 0
        if v:
            a["stalls"][h[6:]] = a["stalls"].get(h[6:], 0) + v
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samp"] for a in agg.values())
print(f"# {rep}: total inst/warp {tot_i / n_warps:.1f}  samples {tot_s:.0f}")
for ln in sorted(agg):
    a = agg[ln]
    if a["inst"] == 0 and a["samp"] == 0:
        continue
    top = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{ln:5d} {a['inst'] / n_warps:8.1f} {100 * a['samp'] / max(tot_s, 1):5.1f}%  [{', '.join(f'{k}:{int(v)}' for k, v in top)}]  {a['src'].strip()[:90]}")
