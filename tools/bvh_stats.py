#!/usr/bin/env python
"""Wide-BVH quality on the CPU (host-stepped device code, tests/emu): nodes / triangle records / instance entries
per ray for camera rays and three generations of diffuse bounce rays. Used to judge builder changes before
spending GPU time.   usage: python tools/bvh_stats.py [scene ...]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import emu  # noqa: E402
import raygen  # noqa: E402

bvhm = importlib.import_module("julia-raytracer_b200.bvh")
lm = importlib.import_module("julia-raytracer_b200.lights")
for scene in sys.argv[1:] or ["classroom", "features1", "ecosys"]:
    sc = orc.jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{scene}.jtscene"))
    b = bvhm.make_scene_bvh(sc)
    lt = lm.make_trace_lights(sc)
    o = orc.Oracle(sc, b, lt)
    t0 = time.time()
    e = emu.Emu(sc, b, lt)
    build_s = time.time() - t0
    p = orc.make_params(resolution=320)
    w, h = o.make_state(p)
    cur = raygen.camera_rays(o, p, w, h, 40000, seed=9)
    gens = [cur]
    for g in range(3):
        cur = raygen.secondary_rays(cur, o.intersect(cur), seed=10 + g)
        if len(cur) == 0:
            break
        gens.append(cur)
    out = {"scene": scene, "stage_s": round(build_s, 2), **e.stats()}
    for label, rays in (("primary", gens[0]), ("secondary", np.concatenate(gens[1:]) if len(gens) > 1 else gens[0])):
        emu.wide_counts()
        hits = e.intersect(rays, 0)
        wc = emu.wide_counts()
        raygen.check_wide_vs_reference(hits, o.intersect(rays), max_rate=2e-4)
        out[label] = {k: round(v / len(rays), 2) for k, v in wc.items()}
    print(json.dumps(out), flush=True)
