"""Pins the CPU oracle (test infrastructure) -- no GPU needed.

The reference ships no tests and cannot be executed here (no julia), so the oracle is pinned by
 (a) the reference's own shipped renders (tests/golden/ref_*.png, made by tools/make_golden.py): mean
     radiance and RMSE of an oracle render against them,
 (b) analytic known-answer tests of the geometric and BSDF primitives,
 (c) structural facts of the reference's scene files (tests/golden/scene_facts.json)."""
import json
import os

import numpy as np
import pytest
from PIL import Image

import orc

GOLDEN = os.path.join(orc.ROOT, "tests", "golden")


def _srgb8(img):
    return orc.jt.sceneio.image_to_srgb8(img)[..., :3].astype(np.float32)


# (scene, sampler, spp, max |mean difference| in 8-bit units, max RMSE in 8-bit units)
# RMSE bounds are loose: the goldens have unknown spp and an unseeded RNG; the mean is the sharp pin.
GOLDEN_CASES = [
    ("cornellbox", "path", 96, 2.0, 10.0),
    ("materials1", "path", 48, 2.5, 12.0),
    ("materials1", "naive", 128, 3.0, 13.0),
    ("features1", "path", 48, 2.5, 8.0),
    ("features1", "naive", 128, 3.0, 12.0),
    ("classroom", "path", 32, 4.0, 18.0),  # texture3.png absent from the checkout -> white
]


@pytest.mark.parametrize("scene,sampler,spp,mean_tol,rmse_tol", GOLDEN_CASES)
def test_oracle_matches_reference_render(scenes, scene, sampler, spp, mean_tol, rmse_tol):
    sc, bvh, lights = scenes(scene)
    o = orc.Oracle(sc)  # own BVH + own lights: the full restatement
    p = orc.make_params(resolution=160, samples=spp, batch=spp, sampler=1 if sampler == "path" else 2)
    o.make_state(p)
    o.trace_samples(p)
    # Compare after a 4x4 box filter in LINEAR space: the goldens are converged (unknown, large spp) while
    # this render is not, and encoding noisy pixels to clipped 8-bit sRGB biases their mean downwards.
    lin = o.get_state()["image"][..., :3]
    ref8 = np.asarray(Image.open(os.path.join(GOLDEN, f"ref_{scene}_{sampler}.png")).convert("RGB"), np.float32) / 255.0
    assert ref8.shape == lin.shape
    ref_lin = np.where(ref8 <= 0.04045, ref8 / 12.92, ((ref8 + 0.055) / 1.055) ** 2.4)

    def box4(a):
        h, w = a.shape[0] // 4 * 4, a.shape[1] // 4 * 4
        return a[:h, :w].reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))

    mine = _srgb8(np.concatenate([box4(lin), np.ones(box4(lin).shape[:2] + (1,), np.float32)], axis=2))
    ref = _srgb8(np.concatenate([box4(ref_lin), np.ones(box4(lin).shape[:2] + (1,), np.float32)], axis=2))
    assert abs(mine.mean() - ref.mean()) < mean_tol
    assert np.sqrt(((mine - ref) ** 2).mean()) < rmse_tol


def test_scene_facts(scenes):
    facts = json.load(open(os.path.join(GOLDEN, "scene_facts.json")))
    for name, f in facts.items():
        sc, bvh, lights = scenes(name)
        assert len(sc.instances) == f["instances"]
        assert sum(len(s.triangles) for s in sc.shapes) == f["triangles"]
        assert sum(len(s.quads) for s in sc.shapes) == f["quads"]
        assert [[t.width, t.height] for t in sc.textures] == f["texture_sizes"]


def _ray(o, d, tmin=1e-4, tmax=np.inf):
    r = np.zeros(1, orc.A.RAY_DTYPE)
    r["o"], r["d"], r["tmin"], r["tmax"] = o, d, tmin, tmax
    return r


def test_single_triangle_known_answer(scenes):
    sc, bvh, lights = scenes("synthetic_one")
    o = orc.Oracle(sc)
    # triangle (-1,-1,0) (1,-1,0) (0,1,0); ray from z=4 straight down -z through (0.25, -0.5, 0)
    h = o.intersect(_ray((0.25, -0.5, 4), (0, 0, -1)))[0]
    assert h["hit"] == 1 and h["instance"] == 1 and h["element"] == 1
    assert h["distance"] == pytest.approx(4.0, rel=1e-6)
    # barycentrics: p = p1 + u (p2-p1) + v (p3-p1) -> v = 0.25, u = (0.25 + 1 - 0.25)/2 = 0.5
    assert h["uv"][0] == pytest.approx(0.5, abs=1e-6) and h["uv"][1] == pytest.approx(0.25, abs=1e-6)
    # miss: SceneIntersection() = (-1, -1, (0,0), 0, false), src/shape.jl:68
    m = o.intersect(_ray((5, 5, 4), (0, 0, -1)))[0]
    assert (m["hit"], m["instance"], m["element"], m["distance"]) == (0, -1, -1, 0.0)
    # Q2: t == tmax is accepted, t < tmin rejected
    assert o.intersect(_ray((0.25, -0.5, 4), (0, 0, -1), tmax=4.0))[0]["hit"] == 1
    assert o.intersect(_ray((0.25, -0.5, 4), (0, 0, -1), tmax=3.999))[0]["hit"] == 0
    assert o.intersect(_ray((0.25, -0.5, 4), (0, 0, -1), tmin=4.001))[0]["hit"] == 0


def test_cornellbox_wall_hits(scenes):
    sc, bvh, lights = scenes("cornellbox")
    o = orc.Oracle(sc)
    # camera at (0,1,3.9) looking down -z: the back wall is the plane z = -1 -> distance 4.9
    h = o.intersect(_ray((0.3, 1.2, 3.9), (0, 0, -1)))[0]
    assert h["hit"] == 1 and h["distance"] == pytest.approx(4.9, rel=1e-5)
    # floor y = 0 from (0.5, 1, 0.9) straight down
    h = o.intersect(_ray((0.9, 1.0, 0.9), (0, -1, 0)))[0]
    assert h["hit"] == 1 and h["distance"] == pytest.approx(1.0, rel=1e-5)


def _bsdf_furnace(mat, n_dirs=200000, seed=0):
    """E[eval/pdf] over sample_bsdfcos must equal the directional albedo <= 1 (energy conservation),
    and pdf must integrate to ~1 over the sphere (checked by uniform sphere sampling)."""
    import ctypes as C
    L = orc.lib()
    rng = np.random.default_rng(seed)
    n = np.asarray([0, 0, 1], np.float32)
    o = np.asarray([0.3, 0.2, 0.9327379], np.float32)
    o /= np.linalg.norm(o)
    m = np.asarray(mat, np.float32)
    out = np.zeros(3, np.float32)
    tot = np.zeros(3)
    cnt = 0
    for _ in range(4000):
        r = rng.random(3).astype(np.float32)
        L.orc_bsdf_sample(m.ctypes.data, n.ctypes.data, o.ctypes.data, C.c_float(r[0]), C.c_float(r[1]), C.c_float(r[2]), 0, out.ctypes.data)
        i = out.copy()
        cnt += 1
        if not i.any():
            continue
        f = np.zeros(3, np.float32)
        p = np.zeros(3, np.float32)
        L.orc_bsdf_eval(m.ctypes.data, n.ctypes.data, o.ctypes.data, i.ctypes.data, 0, f.ctypes.data)
        L.orc_bsdf_eval(m.ctypes.data, n.ctypes.data, o.ctypes.data, i.ctypes.data, 1, p.ctypes.data)
        if p[0] > 0:
            tot += f / p[0]
    albedo = tot / cnt
    # pdf normalisation by uniform sphere sampling
    d = rng.normal(size=(20000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    s = 0.0
    for k in range(len(d)):
        p = np.zeros(3, np.float32)
        v = np.ascontiguousarray(d[k])
        L.orc_bsdf_eval(m.ctypes.data, n.ctypes.data, o.ctypes.data, v.ctypes.data, 1, p.ctypes.data)
        s += p[0]
    return albedo, s / len(d) * 4 * np.pi


@pytest.mark.parametrize("mat,name", [
    ((0, 0.8, 0.7, 0.6, 0.0009, 1.5), "matte"),
    ((1, 0.8, 0.7, 0.6, 0.25, 1.5), "glossy"),
    ((2, 0.9, 0.8, 0.7, 0.25, 1.5), "reflective"),
    ((3, 0.9, 0.9, 0.9, 0.25, 1.5), "transparent"),
    ((4, 0.9, 0.9, 0.9, 0.25, 1.5), "refractive"),
])
def test_bsdf_energy_and_pdf(mat, name):
    albedo, pdf_integral = _bsdf_furnace(mat)
    assert np.all(albedo <= 1.05), (name, albedo)
    assert np.all(albedo > 0.2), (name, albedo)
    if name == "matte":
        assert albedo == pytest.approx(np.asarray(mat[1:4]), rel=2e-2)
    # sampled lobes that can fail (return 0) integrate to <= 1; never above
    # (the rough-refraction Jacobian of src/shading.jl:529-533 is not normalised: a reference quirk we keep)
    assert 0.3 < pdf_integral < 1.1, (name, pdf_integral)


def test_rng_stream():
    v = np.array([orc.lib().orc_rng_float(7, 123, 5, k) for k in range(20000)], np.float32)
    assert v.min() >= 0.0 and v.max() < 1.0
    assert np.all(v * 2 ** 24 == np.floor(v * 2 ** 24))  # multiples of 2^-24 like Julia's rand(Float32)
    assert abs(v.mean() - 0.5) < 0.01 and abs(v.var() - 1 / 12) < 0.005
    w = np.array([orc.lib().orc_rng_float(7, 124, 5, k) for k in range(2000)], np.float32)
    assert abs(np.corrcoef(v[:2000], w)[0, 1]) < 0.08


def test_fmath_accuracy():
    """jt_fmath.h (shared numerical contract) against float64 libm, in ulps of the float result."""
    rng = np.random.default_rng(1)

    def ulps(got, want64):
        want = want64.astype(np.float32)
        ulp = np.spacing(np.maximum(np.abs(want), np.float32(1e-30)))
        return np.max(np.abs(got.astype(np.float64) - want64) / ulp)

    x = rng.uniform(-12, 12, 200000).astype(np.float32)
    assert ulps(orc.fmath(0, x), np.sin(x.astype(np.float64))) < 2.5
    assert ulps(orc.fmath(1, x), np.cos(x.astype(np.float64))) < 2.5
    x = rng.uniform(0, 50, 200000).astype(np.float32)
    assert ulps(orc.fmath(2, x), np.arctan(x.astype(np.float64))) < 3.5
    x = rng.uniform(-1, 1, 200000).astype(np.float32)
    assert ulps(orc.fmath(4, x), np.arccos(x.astype(np.float64))) < 2.5
    x = rng.uniform(-60, 5, 200000).astype(np.float32)
    assert ulps(orc.fmath(5, x), np.exp(x.astype(np.float64))) < 2.0
    x = rng.uniform(1e-6, 2, 200000).astype(np.float32)
    assert ulps(orc.fmath(6, x), np.log(x.astype(np.float64))) < 2.0
    y, xx = rng.uniform(-1, 1, 200000).astype(np.float32), rng.uniform(-1, 1, 200000).astype(np.float32)
    assert ulps(orc.fmath(3, y, xx), np.arctan2(y.astype(np.float64), xx.astype(np.float64))) < 4.5


def test_srgb_lut_matches_per_texel_formula():
    flatten = orc.flatten
    lut = flatten.srgb_to_rgb_lut()
    c = (np.arange(256, dtype=np.float32) / np.float32(255)).astype(np.float32)
    assert np.array_equal(lut, orc.fmath(7, c))


def test_counters_and_algorithmic_bytes(scenes):
    sc, bvh, lights = scenes("cornellbox")
    o = orc.Oracle(sc)
    p = orc.make_params(resolution=32, samples=2, batch=2, sampler=1)
    o.make_state(p)
    o.trace_samples(p)
    c = o.counters()
    assert c["camera_paths"] == 32 * 32 * 2 and c["scene_rays"] >= c["camera_paths"]
    assert orc.algorithmic_bytes(c) > 56 * (c["scene_rays"] + c["light_rays"])
