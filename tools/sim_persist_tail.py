#!/usr/bin/env python
"""Model of the persistent extend kernel's launch tail (CPU only; TEST/ANALYSIS TOOL, uses the host emulation of the
device code to get per-ray step counts, then replays the warp scheduling of jt_dev_persist.cuh).
Question answered: how many warp-iterations of a launch run with few live lanes because the ray queue is empty, and
what would suspending / re-queueing the stragglers save?   usage: python tools/sim_persist_tail.py [scene] [nrays]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import emu  # noqa: E402
import orc  # noqa: E402
import raygen  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "classroom"
nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
sc = orc.jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{name}.jtscene"))
import importlib  # noqa: E402
bvhm = importlib.import_module("julia-raytracer_b200.bvh")
lm = importlib.import_module("julia-raytracer_b200.lights")
bvh, lights = bvhm.make_scene_bvh(sc), lm.make_trace_lights(sc)
o = orc.Oracle(sc, bvh, lights)
e = emu.Emu(sc, bvh, lights)
p = orc.make_params(camera=orc.jt.find_camera(sc, ""), resolution=1280)
w, h = o.make_state(p)
prim = raygen.camera_rays(o, p, w, h, nrays, seed=3)
hits = e.intersect(prim, 0)
sec = raygen.secondary_rays(prim, hits, seed=4)
mix = np.concatenate([prim[: nrays // 3], sec])  # ~1/3 camera rays, 2/3 bounce rays like the in-render mix
rng = np.random.default_rng(0)
mix = mix[rng.permutation(len(mix))]
steps = np.zeros(len(mix), np.int64)
emu.wide_counts(True)
for i in range(len(mix)):
    e.intersect(mix[i:i + 1], 3)
    c = emu.wide_counts(True)
    steps[i] = c["nodes"] + (c["prims"] + 2) // 3 + 1
print(f"{name}: {len(mix)} rays, steps per ray mean {steps.mean():.1f} median {np.median(steps):.0f} p90 {np.percentile(steps, 90):.0f} "
      f"p99 {np.percentile(steps, 99):.0f} max {steps.max()}")


def simulate(queue_len, resident_warps, threshold=20, abandon_below=0, seed=1):
    """All warps advance in lockstep, one loop iteration per tick. Returns warp-iterations by phase."""
    rng = np.random.default_rng(seed)
    q = rng.choice(steps, queue_len)
    qpos = 0
    rem = np.zeros((resident_warps, 32), np.int64)   # remaining iterations per lane (0 = idle)
    done_at = np.zeros((resident_warps, 32), np.int64)
    total_len = np.zeros((resident_warps, 32), np.int64)
    wi_steady = wi_tail = 0
    lanes_steady = lanes_tail = 0
    abandoned = redo = 0
    hist = np.zeros(33, np.int64)
    active = np.ones(resident_warps, bool)
    ticks = 0
    while active.any():
        live = (rem > 0).sum(axis=1)
        if qpos < queue_len:
            for wdx in np.nonzero(active & (live < threshold))[0]:
                idle = np.nonzero(rem[wdx] == 0)[0]
                k = min(len(idle), queue_len - qpos)
                rem[wdx, idle[:k]] = q[qpos:qpos + k]
                total_len[wdx, idle[:k]] = q[qpos:qpos + k]
                qpos += k
                if qpos >= queue_len:
                    break
            live = (rem > 0).sum(axis=1)
        more = qpos < queue_len
        if not more and abandon_below:
            ab = active & (live > 0) & (live < abandon_below)
            abandoned += int(live[ab].sum())
            redo += int(total_len[ab][rem[ab] > 0].sum())
            rem[ab] = 0
            live = (rem > 0).sum(axis=1)
        active = live > 0
        n_act = int(active.sum())
        if n_act == 0:
            break
        if more:
            wi_steady += n_act
            lanes_steady += int(live[active].sum())
        else:
            wi_tail += n_act
            lanes_tail += int(live[active].sum())
        np.add.at(hist, live[active], 1)
        rem[rem > 0] -= 1
        ticks += 1
    return dict(wi_steady=wi_steady, wi_tail=wi_tail, lanes_steady=lanes_steady / max(wi_steady, 1),
                lanes_tail=lanes_tail / max(wi_tail, 1), ticks=ticks, abandoned=abandoned, redo=redo, hist=hist)


for qlen, label in ((460800, "one pipeline of the 1280x720 image, full queue"), (115200, "quarter-full queue (late iterations)")):
    warps = 148 * 24
    base = simulate(qlen, warps)
    tot = base["wi_steady"] + base["wi_tail"]
    ideal = steps.mean() * qlen / 32
    print(f"[{label}] warp-iterations {tot} (ideal at 32 lanes {ideal:.0f}, x{tot / ideal:.2f}); steady {base['wi_steady']} at "
          f"{base['lanes_steady']:.1f} lanes, tail {base['wi_tail']} ({100 * base['wi_tail'] / tot:.0f} %) at {base['lanes_tail']:.1f} lanes; "
          f"ticks {base['ticks']}")
    for ab in (4, 8, 12, 16):
        r = simulate(qlen, warps, abandon_below=ab)
        t2 = r["wi_steady"] + r["wi_tail"]
        # a re-queued ray is traversed again from the root next iteration, at the steady-state lane count
        redo_wi = r["redo"] / max(base["lanes_steady"], 1)
        print(f"    re-queue warps below {ab:2d} lanes once the queue is empty: warp-iterations {t2} + redo {redo_wi:.0f} = "
              f"{(t2 + redo_wi) / tot:.3f} of baseline; {r['abandoned']} rays ({100 * r['abandoned'] / qlen:.2f} %) re-queued; ticks {r['ticks']}")
