#!/usr/bin/env python
"""Freeze the algorithmic bytes per ray / per sample of SURVEY.md §8(d) from ORACLE counters:
    B_ray = 32 (ray in) + 24 (hit out) + 32*N_nodes + 48*N_inst + 36*N_tri + 48*N_quad
counted under the reference's traversal order on the reference's binary BVH over a full render at
reduced resolution/spp (per-sample figures do not depend on either). Writes
profiles/algorithmic_bytes.json, which bench.py multiplies by the samples one launch processes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402

out = {}
for scene, sampler, res, spp in [("cornellbox", "naive", 360, 8), ("cornellbox", "path", 360, 8),
                                 ("features1", "path", 640, 4), ("materials1", "path", 640, 4),
                                 ("classroom", "path", 640, 4), ("ecosys", "path", 480, 2)]:
    sc = orc.jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{scene}.jtscene"))
    o = orc.Oracle(sc)
    p = orc.make_params(resolution=res, samples=spp, batch=spp, sampler=1 if sampler == "path" else 2)
    w, h = o.make_state(p)
    o.trace_samples(p)
    c = o.counters()
    rays = c["scene_rays"] + c["light_rays"]
    b = orc.algorithmic_bytes(c)
    out[f"{scene}_{sampler}"] = {
        "bytes_per_sample": b / c["camera_paths"] + 16.0, "bytes_per_ray": b / rays,
        "scene_bytes_per_scene_ray": orc.algorithmic_bytes_scene(c) / c["scene_rays"],
        "probe_bytes_per_light_ray": (orc.algorithmic_bytes_probe(c) / c["light_rays"]) if c["light_rays"] else 0.0,
        "rays_per_sample": rays / c["camera_paths"], "scene_rays_per_sample": c["scene_rays"] / c["camera_paths"],
        "light_rays_per_sample": c["light_rays"] / c["camera_paths"],
        "nodes_per_ray": (c["tlas_nodes"] + c["blas_nodes"]) / rays, "instance_visits_per_ray": c["instance_visits"] / rays,
        "tri_tests_per_ray": c["tri_tests"] / rays, "quad_tests_per_ray": c["quad_tests"] / rays,
        "measured_on": f"{w}x{h} x {spp} spp, oracle, reference traversal order", "source": "profiles/algorithmic_bytes.json"}
    print(scene, sampler, json.dumps(out[f"{scene}_{sampler}"]), flush=True)
# bytes the wide-BVH kernel itself walks per scene ray (80 B nodes, 48 B triangle records, 112 B instance
# records, 56 B ray in / hit out), counted by the host-stepped device code on camera rays + 3 diffuse bounces
import importlib  # noqa: E402

import numpy as np  # noqa: E402

import emu  # noqa: E402
import raygen  # noqa: E402

bvhm = importlib.import_module("julia-raytracer_b200.bvh")
lm = importlib.import_module("julia-raytracer_b200.lights")
for key in list(out):
    scene = key.rsplit("_", 1)[0]
    sc = orc.jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{scene}.jtscene"))
    b = bvhm.make_scene_bvh(sc)
    lt = lm.make_trace_lights(sc)
    o, e = orc.Oracle(sc, b, lt), emu.Emu(sc, b, lt)
    p = orc.make_params(resolution=320)
    w, h = o.make_state(p)
    cur = raygen.camera_rays(o, p, w, h, 60000, seed=9)
    rays = [cur]
    for g in range(3):
        cur = raygen.secondary_rays(cur, o.intersect(cur), seed=10 + g)
        if len(cur) == 0:
            break
        rays.append(cur)
    rays = np.concatenate(rays)
    emu.wide_counts()
    e.intersect(rays, 0)
    wc = emu.wide_counts()
    out[key]["wide_bytes_per_scene_ray"] = 56 + (80 * wc["nodes"] + 48 * wc["prims"] + 112 * wc["instances"] + 48 * wc["xforms"]) / len(rays)
    out[key]["wide_nodes_per_scene_ray"] = wc["nodes"] / len(rays)
    out[key]["wide_prims_per_scene_ray"] = wc["prims"] / len(rays)
    print(key, "as implemented:", round(out[key]["wide_bytes_per_scene_ray"]), "B per scene ray", flush=True)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json"), "w"), indent=1)
