#!/usr/bin/env python
"""(experiment) How much does the ORDER of a mixed ray queue matter to the persistent extend kernel? classroom, 35 % camera rays +
65 % diffuse bounce rays (the in-render mix): fully shuffled (what atomics-built queues look like) vs camera rays grouped
(random order) vs camera rays grouped in pixel order."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
jt = importlib.import_module("julia-raytracer_b200")
bvh = importlib.import_module("julia-raytracer_b200.bvh")
lights = importlib.import_module("julia-raytracer_b200.lights")
trace = importlib.import_module("julia-raytracer_b200.trace")
A = importlib.import_module("julia-raytracer_b200._abi")
libmod = importlib.import_module("julia-raytracer_b200._lib")
name = sys.argv[1] if len(sys.argv) > 1 else "classroom"
sc = jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{name}.jtscene"))
b = bvh.make_scene_bvh(sc)
d = trace.DeviceScene(sc, b, lights.make_trace_lights(sc), 0)
p = jt.Params(scene=name, resolution=1280, camera=jt.find_camera(sc, ""))
jp = trace.to_jt_params(p)
w, h = jt.image_size(sc.cameras[jp.camera - 1], 1280)
rng = np.random.default_rng(0)
ii, jj = np.meshgrid(np.arange(w), np.arange(h))
ij = np.stack([ii.ravel(), jj.ravel()], 1).astype(np.int32)
r = rng.random((len(ij), 4)).astype(np.float32)
prim = d.sample_camera(jp, w, h, ij, r)
cur, secs, pix, curpix = prim, [], [], np.arange(len(prim))
for g in range(3):
    hits = d.intersect(cur, 0)
    m = hits["hit"] != 0
    curpix = curpix[m]
    pix.append(curpix)
    o = cur["o"][m] + cur["d"][m] * hits["distance"][m][:, None]
    dd = rng.normal(size=(int(m.sum()), 3)).astype(np.float32)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    s = np.zeros(len(o), A.RAY_DTYPE)
    s["o"], s["d"], s["tmin"], s["tmax"] = o, dd, 1e-4, np.inf
    secs.append(s)
    cur = s
sec = np.concatenate(secs)
secpix = np.concatenate(pix)
perm = rng.permutation(len(sec))
sec, secpix = sec[perm], secpix[perm]
n = len(prim)
nsec = min(len(sec), int(n * 0.65 / 0.35))
sec, secpix = sec[:nsec], secpix[:nsec]
sel = np.sort(rng.choice(n, n, replace=False))  # all pixels, pixel order
cases = {
    "shuffled": np.concatenate([prim, sec])[rng.permutation(n + nsec)],
    "camera_grouped_random_order": np.concatenate([prim[rng.permutation(n)], sec]),
    "camera_grouped_pixel_order": np.concatenate([prim[sel], sec]),
    # every ray (camera and bounce) ordered by the pixel it belongs to: what a queue compacted in slot order looks like
    "all_in_pixel_order": np.concatenate([prim, sec])[np.argsort(np.concatenate([np.arange(n), secpix]), kind="stable")],
    "camera_pixel_order_then_bounce_pixel_order": np.concatenate([prim, sec[np.argsort(secpix, kind="stable")]]),
    # bounce rays bucketed by direction octant (what an 8-way keyed append in shade / probe would give)
    "camera_pixel_order_then_bounce_by_octant": np.concatenate([prim, sec[np.argsort(
        (sec["d"][:, 0] < 0) * 1 + (sec["d"][:, 1] < 0) * 2 + (sec["d"][:, 2] < 0) * 4, kind="stable")]]),
    "camera_pixel_order_then_bounce_by_octant_and_pixel": np.concatenate([prim, sec[np.lexsort(
        (secpix, (sec["d"][:, 0] < 0) * 1 + (sec["d"][:, 1] < 0) * 2 + (sec["d"][:, 2] < 0) * 4))]]),
}
out = {"scene": name, "camera": n, "bounce": nsec}
for label, rays in cases.items():
    dr = torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(-1)).cuda()
    dh = torch.empty(len(rays) * 32, dtype=torch.uint8, device="cuda")
    for it in range(3):
        torch.cuda.synchronize(); d.elapsed_ms()
        libmod.check(d.L.jt_intersect_device(d.h, dr.data_ptr(), len(rays), 0, dh.data_ptr()))
        d.synchronize(); ms = d.elapsed_ms()
    out[label] = round(len(rays) / ms / 1e3, 1)
print(json.dumps(out))
