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


# ---- the HDR loader rule (Q9, third-party behaviour of the reference's image loader) -------------------------------
# The reference's environment maps reach the renderer as clamp01(sRGB_encode(rgbe)) (sceneio.hdr_as_reference_loader).
# Evidence 1 (direct): the sky of images/ecosys_path.png is seen by camera rays. Evidence 2 (indirect, two more
# scenes): features1 and materials1 have a ground plane under every camera ray -- NO pixel of their shipped renders
# shows the sky directly (asserted below) -- but the sky is their dominant light, so the rule decides their mean
# radiance: SURVEY 8c's rule (clamp01 of the LINEAR values) is rendered next to it and must miss the golden's mean
# by several times more.
ECOSYS_TREE_RECTS = [  # (x0, y0, x1, y1) as fractions of the reference's 1280x640 render: the three trees whose shape
    (215 / 1280, 210 / 640, 370 / 1280, 325 / 640),  # files (shape002/003.ply, 8 instances) are absent from the checkout,
    (485 / 1280, 90 / 640, 740 / 1280, 315 / 640),   # plus the reflection of the middle tree in the water
    (905 / 1280, 75 / 640, 1.0, 340 / 640),
    (520 / 1280, 480 / 640, 680 / 1280, 620 / 640),
]


def ecosys_masks(h, w):
    """(comparable, sky): pixels outside the missing trees, and the part of them above the horizon."""
    keep = np.ones((h, w), bool)
    for x0, y0, x1, y1 in ECOSYS_TREE_RECTS:
        keep[int(np.floor(y0 * h)):int(np.ceil(y1 * h)), int(np.floor(x0 * w)):int(np.ceil(x1 * w))] = False
    sky = keep.copy()
    sky[int(0.42 * h):] = False
    return keep, sky


def test_hdr_rule_directly_visible_sky_of_ecosys(scenes):
    sc, bvh, lights = scenes("ecosys")
    o = orc.Oracle(sc, bvh, lights)
    p = orc.make_params(resolution=320, samples=8, batch=8, sampler=1)
    o.make_state(p)
    o.trace_samples(p)
    ref = np.asarray(Image.open(os.path.join(GOLDEN, "ref_ecosys_path.png")).convert("RGB"), np.float32)
    enc = orc.jt.sceneio.image_to_srgb8(o.get_state()["image"])[..., :3].astype(np.uint8)
    mine = np.asarray(Image.fromarray(enc, "RGB").resize((ref.shape[1], ref.shape[0]), Image.BOX), np.float32)
    keep, sky = ecosys_masks(*ref.shape[:2])
    d = mine - ref
    sky_mean, sky_rmse = abs(mine[sky].mean() - ref[sky].mean()), float(np.sqrt((d[sky] ** 2).mean()))
    print(f"\n[ecosys sky, {sky.mean():.0%} of the image] mean diff {sky_mean:.2f}, RMSE {sky_rmse:.2f} (8-bit units)")
    assert sky_mean < 0.6 and sky_rmse < 2.0  # measured 0.13 / 0.95
    assert abs(mine[keep].mean() - ref[keep].mean()) < 6.0  # plants at 8 spp: only the mean is meaningful


@pytest.mark.parametrize("scene", ["materials1", "features1"])
def test_hdr_rule_discriminates_on_sky_lit_scenes(scenes, scene):
    import copy
    sc, _, _ = scenes(scene)
    ref8 = np.asarray(Image.open(os.path.join(GOLDEN, f"ref_{scene}_path.png")).convert("RGB"), np.float32)
    # the alternative: clamp01 of the linear radiance. The packed texture holds q = sRGB_encode(clamp01(c)), so
    # clamp01(c) = sRGB_decode(q) up to the 16-bit quantum.
    alt = copy.deepcopy(sc)
    n_hdr = 0
    for t in alt.textures:
        if t.pixelsf is not None:
            q = t.pixelsf[:, :3].astype(np.float64)
            t.pixelsf[:, :3] = np.where(q <= 0.04045, q / 12.92, ((q + 0.055) / 1.055) ** 2.4).astype(np.float32)
            n_hdr += 1
    assert n_hdr == 1
    means = {}
    for tag, s in (("rule", sc), ("linear-clamp", alt)):
        o = orc.Oracle(s)
        p = orc.make_params(resolution=160, samples=32, batch=32, sampler=1)
        o.make_state(p)
        o.trace_samples(p)
        st = o.get_state()
        if tag == "rule":  # no camera ray sees the sky: an environment miss at bounce 0 writes albedo (1, 1, 1)
            assert (st["albedo"].min(axis=-1) < 1.0).all() and (st["hits"] == 32).all()
        lin = st["image"][..., :3]
        h, w = lin.shape[0] // 4 * 4, lin.shape[1] // 4 * 4
        box = lin[:h, :w].reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))
        means[tag] = _srgb8(np.concatenate([box, np.ones(box.shape[:2] + (1,), np.float32)], axis=2)).mean()
    r8 = ref8 / 255.0
    rl = np.where(r8 <= 0.04045, r8 / 12.92, ((r8 + 0.055) / 1.055) ** 2.4)
    h, w = rl.shape[0] // 4 * 4, rl.shape[1] // 4 * 4
    rbox = rl[:h, :w].reshape(h // 4, 4, w // 4, 4, 3).mean(axis=(1, 3))
    gold = _srgb8(np.concatenate([rbox, np.ones(rbox.shape[:2] + (1,), np.float32)], axis=2)).mean()
    e_rule, e_alt = abs(means["rule"] - gold), abs(means["linear-clamp"] - gold)
    print(f"\n[{scene}] golden mean {gold:.1f}; loader rule {means['rule']:.1f} (off by {e_rule:.1f}); "
          f"linear clamp {means['linear-clamp']:.1f} (off by {e_alt:.1f})")
    assert e_rule < 2.5 and e_alt > 3 * max(e_rule, 1.0)
