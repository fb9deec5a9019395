#!/usr/bin/env python
"""Closest-hit microbenchmark (GPU): Mrays/s of jt_intersect_device for primary and diffuse-bounce rays
of a scene, per traversal mode (0 = persistent wide, 2 = plain wide, 1 = reference order).
usage: python tools/bench_traverse.py [scene ...]"""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
jt = importlib.import_module("julia-raytracer_b200")
bvh = importlib.import_module("julia-raytracer_b200.bvh")
lights = importlib.import_module("julia-raytracer_b200.lights")
trace = importlib.import_module("julia-raytracer_b200.trace")
A = importlib.import_module("julia-raytracer_b200._abi")
libmod = importlib.import_module("julia-raytracer_b200._lib")

scenes = sys.argv[1:] or ["classroom", "features1", "ecosys"]
for name in scenes:
    sc = jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{name}.jtscene"))
    b = bvh.make_scene_bvh(sc)
    d = trace.DeviceScene(sc, b, lights.make_trace_lights(sc), 0)
    p = jt.Params(scene=name, resolution=1280, camera=jt.find_camera(sc, ""))
    jp = trace.to_jt_params(p)
    w, h = jt.image_size(sc.cameras[jp.camera - 1], 1280)
    rng = np.random.default_rng(0)
    ii, jj = np.meshgrid(np.arange(w), np.arange(h))
    ij = np.stack([ii.ravel(), jj.ravel()], 1).astype(np.int32)
    ij = np.tile(ij, (2, 1))
    r = rng.random((len(ij), 4)).astype(np.float32)
    prim = d.sample_camera(jp, w, h, ij, r)
    hits = d.intersect(prim, 0)
    m = hits["hit"] != 0
    o = prim["o"][m] + prim["d"][m] * hits["distance"][m][:, None]
    dd = rng.normal(size=(int(m.sum()), 3)).astype(np.float32)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    sec = np.zeros(len(o), A.RAY_DTYPE)
    sec["o"], sec["d"], sec["tmin"], sec["tmax"] = o, dd, 1e-4, np.inf
    sec = sec[rng.permutation(len(sec))]  # incoherent order, like compacted path queues
    out = {}
    for label, rays in (("primary", prim), ("secondary", sec)):
        dr = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
        dh = torch.empty(len(rays) * 32, dtype=torch.uint8, device="cuda")
        for mode, mname in ((0, "persist"), (2, "plain"), (1, "reference")):
            for it in range(3):
                torch.cuda.synchronize()
                d.elapsed_ms()
                libmod.check(d.L.jt_intersect_device(d.h, dr.data_ptr(), len(rays), mode, dh.data_ptr()))
                d.synchronize()
                ms = d.elapsed_ms()
            out[f"{label}_{mname}_mrays"] = round(len(rays) / ms / 1e3, 1)
    print(json.dumps({"scene": name, "rays": {"primary": len(prim), "secondary": len(sec)}, **out}), flush=True)
    d.close()
