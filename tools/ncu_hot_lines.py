#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples for one kernel of an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_hot_lines.py prof.ncu-rep kernel_regex [N]"""
import collections
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{rx}"], capture_output=True, text=True).stdout
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        stall_cols = [(h, i) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and cur and len(r) > hdr["Thread Instructions Executed"] and r[0].isdigit() and r[hdr["# Samples"]].isdigit():
        key = (cur, int(r[0]), r[1].strip()[:100])
        a = agg[key]
        a[0] += int(r[hdr["# Samples"]])
        a[1] += int(r[hdr["Instructions Executed"]])
        a[2] += int(r[hdr["Thread Instructions Executed"]])
        for h, i in stall_cols:
            if r[i].isdigit():
                a[3][h] += int(r[i])
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
allst = collections.Counter()
for v in agg.values():
    allst.update(v[3])
print("stall reasons:", ", ".join(f"{k[6:]} {100 * c / sum(allst.values()):.0f}%" for k, c in allst.most_common(7)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{h[6:]}:{c}" for h, c in v[3].most_common(2))
    print(f"{100 * v[0] / tot:5.1f}% smp {100 * v[1] / toti:5.1f}% inst lanes {v[2] / max(v[1], 1):4.1f} {k[0]}:{k[1]} [{st}] {k[2]}")
