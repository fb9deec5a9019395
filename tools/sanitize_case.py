#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both integrators, both traversal
modes, both samplers on the synthetic all-features scene + cornellbox at tiny resolution, plus the
identical-rays hooks. usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402

jt = importlib.import_module("julia-raytracer_b200")
bvh = importlib.import_module("julia-raytracer_b200.bvh")
lights = importlib.import_module("julia-raytracer_b200.lights")
trace = importlib.import_module("julia-raytracer_b200.trace")
A = importlib.import_module("julia-raytracer_b200._abi")

for name in ("synthetic_all", "cornellbox", "features1"):
    sc = synth.make_scene(name) if name.startswith("synthetic") else jt.load_scene(
        os.path.join(ROOT, "assets", "scenes", f"{name}.jtscene"))
    b = bvh.make_scene_bvh(sc)
    d = trace.DeviceScene(sc, b, lights.make_trace_lights(sc), 0)
    for sampler in (1, 2):
        for trav in ("wide", "reference"):
            for integ in ("wavefront", "megakernel"):
                p = jt.Params(scene=name, resolution=40, samples=2, batch=2, sampler=sampler, camera=1,
                              gpu_traversal=trav, gpu_integrator=integ)
                st = trace.make_trace_state(d, p)
                trace.trace_samples(st, d, None, None, p)
                st.sync()
                assert np.isfinite(st.image).all()
                st.close()
    rng = np.random.default_rng(0)
    rays = np.zeros(4096, A.RAY_DTYPE)
    rays["o"] = rng.normal(size=(4096, 3)) * 2 + [0, 1, 3]
    dd = rng.normal(size=(4096, 3))
    rays["d"] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    rays["tmin"], rays["tmax"] = 1e-4, np.inf
    for mode in (0, 1, 2):
        d.intersect(rays, mode)
    d.intersect_instance(rays, rng.integers(1, len(sc.instances) + 1, 4096), 0)
    d.close()
    print("ok", name, flush=True)
