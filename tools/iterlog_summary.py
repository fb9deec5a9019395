#!/usr/bin/env python
"""Summarise a JT_ITER_LOG file: time share and work share of a 512-spp step by extend-queue fill.
usage: python tools/iterlog_summary.py log.txt [slots_per_pipeline]"""
import sys

import numpy as np

runs, cur = [], []
for line in open(sys.argv[1]):
    if line.startswith("end"):
        runs.append(cur)
        cur = []
    else:
        cur.append([float(x) for x in line.split()])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 460800
r = np.array(max(runs, key=len))
print(f"{sys.argv[1]}: {len(runs)} ranges logged, longest has {len(r)} polls")
for k in sorted(set(r[:, 1].astype(int))):
    a = r[r[:, 1] == k]
    t, q = a[:, 4], a[:, 2]
    dt = np.diff(np.concatenate([[0], t]))
    fill = np.concatenate([[1.0], q[:-1] / n])
    print(f" pipeline {k}: {len(a)} polls of 4 iterations, {t[-1]:.1f} ms")
    for lo, hi in ((0.9, 1.01), (0.75, 0.9), (0.5, 0.75), (0.25, 0.5), (0.1, 0.25), (0.02, 0.1), (0, 0.02)):
        m = (fill >= lo) & (fill < hi)
        if m.any():
            print(f"   queue fill {lo:.2f}-{hi:.2f}: {m.sum():4d} polls  time {100 * dt[m].sum() / t[-1]:5.1f} %  "
                  f"work {100 * fill[m].sum() / fill.sum():5.1f} %  {dt[m].mean():.3f} ms per poll")
