"""-m gpu: the parity tests proper. Everything goes through the C ABI of libjtrace_b200.so (via the
ctypes host mirror) on a real B200 and is compared with the CPU oracle on the same seeded inputs."""
import ctypes as C
import importlib

import numpy as np
import pytest

import orc
import raygen

pytestmark = pytest.mark.gpu

jt = orc.jt
trace = importlib.import_module("julia-raytracer_b200.trace")
libmod = importlib.import_module("julia-raytracer_b200._lib")

SCENES = ["cornellbox", "materials1", "features1", "classroom", "ecosys", "synthetic_all", "synthetic_closed",
          "synthetic_one"]


@pytest.fixture(scope="module")
def pair(scenes):
    cache = {}

    def get(name):
        if name not in cache:
            sc, bvh, lights = scenes(name)
            cache[name] = (orc.Oracle(sc, bvh, lights), trace.DeviceScene(sc, bvh, lights, 0))
        return cache[name]

    yield get
    for _, d in cache.values():
        d.close()


# Wide-BVH mode vs the oracle on a fixed sample set: the only admissible source of a differing pixel is the residual
# class of DESIGN.md section 1 (the reference's non-conservative slab test culls a node that holds the true closest hit;
# the wide BVH's conservative boxes keep it). Measured on B200 (profiles/r02/pytest_gpu_*.log): the fraction of pixels
# differing by > 1e-4 is exactly 0 on every case, including C1-C5 at the BASELINE sizes (0 of 1 843 200 pixels on ecosys); the bound leaves room
# for other sample sets, NOT the 2e-3 of round 1.
WIDE_PIXEL_FRACTION_BOUND = 1e-5


def wide_mode_check(tag, o, d, op, img, ref_img, spp, sample_begin=0):
    """Print the measured fraction of pixels differing by more than 1e-4 (north-star gate 2), bound it, and EXPLAIN
    every differing pixel: replay the oracle's rays of that pixel's samples through both GPU traversals; the first
    closest-hit query on which the wide traversal departs from the reference must return a strictly closer hit."""
    err = np.abs(img - ref_img).max(axis=-1)
    bad = err > 1e-4
    frac = float(bad.mean())
    print(f"\n[{tag}] wide traversal vs oracle: {int(bad.sum())} of {bad.size} pixels differ by > 1e-4 "
          f"(fraction {frac:.2e}, max |diff| {float(err.max()):.3g})")
    assert frac <= WIDE_PIXEL_FRACTION_BOUND, frac
    h, w = bad.shape
    for j, i in np.argwhere(bad)[:64]:
        explained = False
        for s_ in range(sample_begin, sample_begin + spp):
            rays, inst = o.trace_pixel_rays(op, int(i), int(j), s_)
            scene_q = inst < 0
            got = np.zeros(len(rays), orc.A.HIT_DTYPE)
            want = np.zeros(len(rays), orc.A.HIT_DTYPE)
            if scene_q.any():
                got[scene_q] = d.intersect(rays[scene_q], 0)
                want[scene_q] = o.intersect(rays[scene_q])
            if (~scene_q).any():
                got[~scene_q] = d.intersect_instance(rays[~scene_q], inst[~scene_q], 0)
                want[~scene_q] = o.intersect_instance(rays[~scene_q], inst[~scene_q])
            differs = (got["instance"] != want["instance"]) | (got["element"] != want["element"]) | (got["hit"] != want["hit"])
            if differs.any():
                k = int(np.argmax(differs))  # the path is only comparable up to its first divergent query
                closer = got["hit"][k] == 1 and (want["hit"][k] == 0 or got["distance"][k] < want["distance"][k])
                assert closer, (tag, int(i), int(j), s_, k, got[k], want[k])
                explained = True
        assert explained, f"{tag}: pixel ({i}, {j}) differs but none of its rays is in the residual class"
    return frac


def _params(**kw):
    d = dict(scene="x", resolution=64, samples=3, batch=3, sampler=1, camera=1)
    d.update(kw)
    return jt.Params(**d)


@pytest.mark.parametrize("name", SCENES)
def test_identical_rays_ids_t_uv(pair, name):
    """North-star gate 1: hit ids bit-exact, t/uv <= 1e-5 relative (bit-exact expected)."""
    o, d = pair(name)
    op = orc.make_params(resolution=256)
    w, h = o.make_state(op)
    n = 200000 if name != "ecosys" else 60000
    rays = raygen.camera_rays(o, op, w, h, n, seed=21)
    allr = np.concatenate([rays, raygen.secondary_rays(rays, o.intersect(rays), seed=22)])
    ref = o.intersect(allr)
    got = d.intersect(allr, 1)
    r = raygen.compare_hits(got, ref)
    assert r["id_mismatch"] == 0 and r["t_mismatch"] == 0 and r["uv_mismatch"] == 0, ("reference-order mode", r)
    r = raygen.check_wide_vs_reference(d.intersect(allr, 0), ref)
    print(f"\n[{name}] wide-mode residual mismatches: {r['id_mismatch']} of {r['n']} rays")


@pytest.mark.parametrize("name", ["cornellbox", "features1", "synthetic_all"])
def test_instance_probes(pair, scenes, name):
    o, d = pair(name)
    sc, _, _ = scenes(name)
    op = orc.make_params(resolution=128)
    w, h = o.make_state(op)
    rays = raygen.camera_rays(o, op, w, h, 50000, seed=23)
    inst = np.random.default_rng(24).integers(1, len(sc.instances) + 1, len(rays))
    ref = o.intersect_instance(rays, inst)
    r = raygen.compare_hits(d.intersect_instance(rays, inst, 1), ref)
    assert r["id_mismatch"] == 0 and r["t_mismatch"] == 0 and r["uv_mismatch"] == 0, r
    raygen.check_wide_vs_reference(d.intersect_instance(rays, inst, 0), ref)


def test_sample_camera_bit_exact(pair):
    o, d = pair("features1")
    op = orc.make_params(resolution=320)
    w, h = o.make_state(op)
    rng = np.random.default_rng(25)
    ij = np.stack([rng.integers(0, w, 4096), rng.integers(0, h, 4096)], 1).astype(np.int32)
    r = (rng.integers(0, 1 << 24, (4096, 4)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    for tent in (0, 1):
        op.tentfilter = tent
        jp = trace.to_jt_params(_params(tentfilter=bool(tent)))
        assert d.sample_camera(jp, w, h, ij, r).tobytes() == o.sample_camera(op, w, h, ij, r).tobytes()


CASES = [("cornellbox", 2, {}), ("cornellbox", 1, {}), ("features1", 1, {}), ("features1", 2, {}),
         ("materials1", 1, {}), ("classroom", 1, {}), ("ecosys", 1, {}), ("synthetic_all", 1, {}),
         ("synthetic_all", 2, {}), ("synthetic_all", 1, dict(nocaustics=True, tentfilter=True, envhidden=True)),
         ("synthetic_closed", 1, {}), ("synthetic_closed", 2, dict(envhidden=True)), ("synthetic_one", 1, {}),
         ("synthetic_all", 1, dict(bounces=2, clamp=1))]


@pytest.mark.parametrize("name,sampler,extra", CASES)
def test_fixed_sample_set_images(pair, name, sampler, extra):
    """North-star gate 2: with the shared counter-based RNG the image matches the oracle within 1e-4
    per channel (reference-order mode: bit-exact; wide mode: report the fraction of pixels above it)."""
    o, d = pair(name)
    res = 96 if name != "ecosys" else 64
    okw = {k: int(v) for k, v in extra.items()}
    op = orc.make_params(resolution=res, samples=3, batch=3, sampler=sampler, seed=31, **okw)
    w, h = o.make_state(op)
    o.trace_samples(op)
    ref = o.get_state()
    oc = o.counters(reset=True)
    for traversal, integrator in (("reference", "wavefront"), ("wide", "wavefront"), ("reference", "megakernel"),
                                  ("wide", "megakernel")):
        p = _params(resolution=res, sampler=sampler, gpu_seed=31, gpu_traversal=traversal, gpu_integrator=integrator,
                    **extra)
        st = trace.make_trace_state(d, p)
        assert (st.width, st.height) == (w, h)
        d.counters(reset=True)
        trace.trace_samples(st, d, None, None, p)
        st.sync()
        c = d.counters()
        img = st.image.reshape(h, w, 4)
        err = np.abs(img - ref["image"]).max(axis=-1)
        if traversal == "reference":
            assert np.array_equal(img, ref["image"]), (integrator, float(err.max()))
            assert np.array_equal(st.albedo.reshape(h, w, 3), ref["albedo"])
            assert np.array_equal(st.normal.reshape(h, w, 3), ref["normal"])
            assert np.array_equal(st.hits.reshape(h, w), ref["hits"])
            assert (c["camera_paths"], c["scene_rays"], c["light_rays"]) == (oc["camera_paths"], oc["scene_rays"], oc["light_rays"])
        elif integrator == "wavefront":
            wide_mode_check(f"{name} sampler {sampler} {res}px", o, d, op, img, ref["image"], 3)
        else:
            assert (err > 1e-4).mean() <= WIDE_PIXEL_FRACTION_BOUND, (err > 1e-4).mean()
        st.close()


def test_batching_and_ranges_are_equivalent(pair):
    """trace_samples called samples/batch times == one call; sample ranges compose (sharding unit)."""
    o, d = pair("cornellbox")
    p1 = _params(samples=4, batch=4, sampler=2)
    a = trace.make_trace_state(d, p1)
    trace.trace_samples(a, d, None, None, p1)
    a.sync()
    pb = _params(samples=4, batch=1, sampler=2)
    b = trace.make_trace_state(d, pb)
    for _ in range(4):
        trace.trace_samples(b, d, None, None, pb)
    assert b.samples == 4
    trace.trace_samples(b, d, None, None, pb)  # state.samples >= params.samples: no-op (src/trace.jl:225)
    assert b.samples == 4
    b.sync()
    assert np.array_equal(a.image, b.image) and np.array_equal(a.hits, b.hits)
    c = trace.make_trace_state(d, p1)
    trace.trace_sample_range(c, d, p1, 0, 1)
    trace.trace_sample_range(c, d, p1, 1, 4)
    c.sync()
    assert np.array_equal(a.image, c.image)


def test_sum_mode_and_logical_shards(pair):
    """per-GPU sums over disjoint global sample indices, added, divided by N, equal the
    single-GPU sums bit-for-bit when added in the same order, and the running mean to ~1e-6."""
    o, d = pair("features1")
    p = _params(samples=4, batch=4, sampler=1, resolution=80)
    mean = trace.make_trace_state(d, p, accumulate=0)
    trace.trace_samples(mean, d, None, None, p)
    mean.sync()
    whole = trace.make_trace_state(d, p, accumulate=1)
    trace.trace_sample_range(whole, d, p, 0, 4)
    whole.sync()
    assert np.allclose(whole.image, mean.image, rtol=1e-5, atol=1e-6)
    shards = []
    for g in range(2):  # two logical GPUs on one device: samples {0,1} and {2,3}
        s = trace.make_trace_state(d, p, accumulate=1)
        trace.trace_sample_range(s, d, p, 2 * g, 2 * g + 2)
        s.set_samples(1)  # download raw sums (scale 1/1)
        s.sync()
        shards.append(s.image.copy())
    merged = (shards[0] + shards[1]) * np.float32(0.25)
    assert np.allclose(merged, whole.image, rtol=1e-6, atol=1e-7)
    assert np.array_equal(whole.hits, mean.hits)


def test_srgb8_output_matches_host_save_path(pair):
    """N3: GPU rgb_to_srgb + clamp + 8-bit quantise vs the host formula (+-1 LSB: third-party rounding)."""
    o, d = pair("features1")
    p = _params(samples=4, batch=4, sampler=1, resolution=96)
    st = trace.make_trace_state(d, p)
    trace.trace_samples(st, d, None, None, p)
    st.sync()
    gpu = st.srgb8().astype(np.int32)
    host = jt.sceneio.image_to_srgb8(st.image.reshape(st.height, st.width, 4)).astype(np.int32)
    assert gpu.shape == host.shape
    assert np.abs(gpu - host).max() <= 1 and (gpu != host).mean() < 0.01
    st.close()


def test_error_paths(pair, scenes):
    o, d = pair("cornellbox")
    L = d.L
    jp = trace.to_jt_params(_params())
    jp.camera = 7
    h = C.c_void_p()
    assert L.jt_state_create(d.h, C.byref(jp), C.byref(h)) == -1 and b"camera" in L.jt_last_error()
    jp = trace.to_jt_params(_params())
    jp.sampler = 3
    assert L.jt_state_create(d.h, C.byref(jp), C.byref(h)) == -1
    assert L.jt_scene_create(None, 0, C.byref(h)) == -1
    sc, bvh, lights = scenes("cornellbox")
    with pytest.raises(libmod.JtError):
        trace.DeviceScene(sc, bvh, lights, 99)  # device out of range
    # gltfpbr throws in the reference -> JT_ERR_UNSUPPORTED at the boundary
    import copy
    bad = copy.deepcopy(sc)
    bad.materials["type"][0] = 7
    with pytest.raises(libmod.JtError) as e:
        trace.DeviceScene(bad, bvh, lights, 0)
    assert e.value.code == -4
    # empty ray batch is fine
    assert len(d.intersect(np.zeros(0, orc.A.RAY_DTYPE))) == 0


def test_main_drop_in(tmp_path):
    """Jtrace.main with the reference's own command line renders on the GPU and writes the PNG."""
    import os
    j = importlib.import_module("julia-raytracer_b200.jtrace")
    out = str(tmp_path / "out.png")
    scene = os.path.join(orc.ROOT, "assets", "scenes", "cornellbox.jtscene")
    r = j.main(f"--scene {scene} --sampler naive --resolution 72 --samples 4 --batch 2 --output {out}")
    assert os.path.exists(out) and r["image"].shape == (72, 72, 4) and r["counters"]["camera_paths"] == 72 * 72 * 4


# (scene, sampler, spp, max |mean difference|, max RMSE) in 8-bit sRGB units after the same box reduction to 160 px the goldens got.
# The goldens are the reference's OWN shipped renders (1280 px, unknown spp, unseeded RNG, full asset set): classroom
# misses texture3.png here (white instead), hence its looser bound.
# Measured on a B200 (profiles/r01/pytest_gpu_*.log): mean diff 0.09 / 0.00 / 0.12 / 0.00 / 0.30 / 0.57, RMSE 0.84 / 0.60 /
# 1.26 / 0.83 / 1.77 / 2.49 -- the bounds are about twice that.
CONVERGED_CASES = [("cornellbox", "path", 1024, 0.5, 2.0), ("materials1", "path", 512, 0.5, 1.5),
                   ("materials1", "naive", 1024, 0.8, 3.0), ("features1", "path", 512, 0.5, 2.0),
                   ("features1", "naive", 1024, 1.0, 4.0), ("classroom", "path", 1024, 1.5, 5.0)]


@pytest.mark.parametrize("name,sampler,spp,mean_tol,rmse_tol", CONVERGED_CASES)
def test_converged_render_matches_the_reference_image(pair, name, sampler, spp, mean_tol, rmse_tol):
    """North-star gate 3 on the GPU itself: the full-resolution (1280 px) render, converged, against the reference's
    shipped image of the same scene and sampler."""
    import os
    from PIL import Image
    _, d = pair(name)
    ref8 = np.asarray(Image.open(os.path.join(orc.ROOT, "tests", "golden", f"ref_{name}_{sampler}.png")).convert("RGB"), np.float32)
    p = _params(scene=name, resolution=1280, samples=spp, batch=spp, sampler=1 if sampler == "path" else 2)
    st = trace.make_trace_state(d, p)
    trace.trace_samples(st, d, None, None, p)
    st.sync()
    lin = st.image.reshape(st.height, st.width, 4)[..., :3]
    # same reduction as tools/make_golden.py applied to the reference's PNG: PIL BOX resize of the 8-bit sRGB image
    enc = jt.sceneio.image_to_srgb8(np.concatenate([lin, np.ones(lin.shape[:2] + (1,), np.float32)], axis=2))[..., :3]
    mine = np.asarray(Image.fromarray(enc.astype(np.uint8), "RGB").resize((ref8.shape[1], ref8.shape[0]), Image.BOX), np.float32)
    dm, rmse = abs(mine.mean() - ref8.mean()), float(np.sqrt(((mine - ref8) ** 2).mean()))
    print(f"\n[{name} {sampler} {spp} spp] mean diff {dm:.2f}, RMSE {rmse:.2f} (8-bit units)")
    st.close() if hasattr(st, "close") else None
    assert dm < mean_tol and rmse < rmse_tol, (dm, rmse)


def test_baseline_config_c1_full_size_is_bit_exact(pair):
    """BASELINE config C1 at its full size -- cornellbox, naive sampler, 720 px, 16 spp (8.3 M camera paths): the GPU's
    reference-order mode equals the oracle bit for bit, the fast mode to <= 2e-3 of the pixels (residual class)."""
    o, d = pair("cornellbox")
    op = orc.make_params(resolution=720, samples=16, batch=16, sampler=2, seed=0)
    w, h = o.make_state(op)
    assert (w, h) == (720, 720)
    o.trace_samples(op)
    ref = o.get_state()
    for traversal in ("reference", "wide"):
        p = _params(scene="cornellbox", resolution=720, samples=16, batch=1, sampler=2, gpu_traversal=traversal)
        st = trace.make_trace_state(d, p)
        for _ in range(16):  # the reference's own call pattern: samples / batch calls (src/jtrace.jl:83-94)
            trace.trace_samples(st, d, None, None, p)
        st.sync()
        img = st.image.reshape(h, w, 4)
        if traversal == "reference":
            assert np.array_equal(img, ref["image"]) and np.array_equal(st.hits.reshape(h, w), ref["hits"])
            assert np.array_equal(st.albedo.reshape(h, w, 3), ref["albedo"])
            assert np.array_equal(st.normal.reshape(h, w, 3), ref["normal"])
        else:
            wide_mode_check("C1 cornellbox naive 720px 16spp", o, d, op, img, ref["image"], 16)
        st.close()


# BASELINE configs C2-C5 at their full image sizes; spp reduced (the properties below do not depend on it)
FULL_SIZE = [("features1", 1280, (1280, 533)), ("materials1", 1280, (1280, 533)), ("classroom", 1280, (1280, 720)),
             ("ecosys", 1920, (1920, 960))]


@pytest.mark.parametrize("name,res,size", FULL_SIZE)
def test_baseline_configs_full_size_properties(pair, name, res, size):
    """Size-independent properties at the BASELINE image sizes (path sampler):
    * the fast traversal (wide / flattened BVH, wavefront) differs from the reference-order traversal on at most
      2e-3 of the pixels (residual class) and both count the same camera paths;
    * determinism: the same sample range rendered twice is bit-identical;
    * sharding: two disjoint sample ranges rendered as separate sum-mode states add up to the one-state render;
    * alpha / hits bookkeeping: hits <= spp and alpha == hits / spp in running-mean mode."""
    _, d = pair(name)
    spp = 4
    p = _params(scene=name, resolution=res, samples=spp, batch=spp, sampler=1)
    imgs = {}
    for traversal in ("wide", "reference"):
        q = _params(scene=name, resolution=res, samples=spp, batch=spp, sampler=1, gpu_traversal=traversal)
        st = trace.make_trace_state(d, q)
        assert (st.width, st.height) == size
        d.counters(reset=True)
        trace.trace_samples(st, d, None, None, q)
        st.sync()
        assert d.counters()["camera_paths"] == size[0] * size[1] * spp
        imgs[traversal] = (st.image.copy(), st.hits.copy())
        st.close()
    img, hits = imgs["wide"]
    assert np.isfinite(img).all()
    differing = np.abs(img - imgs["reference"][0]).max(axis=-1) > 1e-4
    print(f"\n[{name} {size[0]}x{size[1]} {spp}spp] GPU wide vs GPU reference-order: fraction of differing pixels {differing.mean():.2e}")
    assert differing.mean() <= WIDE_PIXEL_FRACTION_BOUND, differing.mean()
    assert hits.max() <= spp and np.allclose(img[:, 3], hits / np.float32(spp), atol=1e-6)
    again = trace.make_trace_state(d, p)
    trace.trace_sample_range(again, d, p, 0, 1)  # different chunking, same samples
    trace.trace_sample_range(again, d, p, 1, spp)
    again.sync()
    assert np.array_equal(again.image, img) and np.array_equal(again.hits, hits)
    again.close()
    whole = trace.make_trace_state(d, p, accumulate=1)
    trace.trace_sample_range(whole, d, p, 0, spp)
    whole.set_samples(1)
    whole.sync()
    parts = []
    for g in range(2):
        s = trace.make_trace_state(d, p, accumulate=1)
        trace.trace_sample_range(s, d, p, g * spp // 2, (g + 1) * spp // 2)
        s.set_samples(1)
        s.sync()
        parts.append(s.image.copy())
        s.close()
    assert np.allclose(parts[0] + parts[1], whole.image, rtol=1e-6, atol=1e-6)
    whole.close()


@pytest.mark.parametrize("name,res,size", FULL_SIZE)
def test_baseline_configs_full_size_against_the_oracle(pair, name, res, size):
    """BASELINE configs C2-C5 at their full image sizes, path sampler, 2 spp, ORACLE vs GPU (not GPU vs GPU):
    * reference-order traversal + wavefront integrator: image, albedo, normal, hits and the ray counters bit-exact;
    * the BENCHMARKED mode (wide traversal + wavefront): measured fraction of differing pixels printed and bounded,
      every differing pixel traced back to a residual-class ray."""
    o, d = pair(name)
    spp = 2
    op = orc.make_params(resolution=res, samples=spp, batch=spp, sampler=1, seed=0)
    w, h = o.make_state(op)
    assert (w, h) == size
    o.counters(reset=True)
    o.trace_samples(op)
    ref = o.get_state()
    oc = o.counters(reset=True)
    for traversal in ("reference", "wide"):
        p = _params(scene=name, resolution=res, samples=spp, batch=spp, sampler=1, gpu_traversal=traversal)
        st = trace.make_trace_state(d, p)
        d.counters(reset=True)
        trace.trace_samples(st, d, None, None, p)
        st.sync()
        c = d.counters()
        img = st.image.reshape(h, w, 4)
        if traversal == "reference":
            assert np.array_equal(img, ref["image"]), float(np.abs(img - ref["image"]).max())
            assert np.array_equal(st.albedo.reshape(h, w, 3), ref["albedo"])
            assert np.array_equal(st.normal.reshape(h, w, 3), ref["normal"])
            assert np.array_equal(st.hits.reshape(h, w), ref["hits"])
            assert (c["camera_paths"], c["scene_rays"], c["light_rays"]) == (oc["camera_paths"], oc["scene_rays"], oc["light_rays"])
        else:
            assert c["camera_paths"] == oc["camera_paths"]
            wide_mode_check(f"{name} {w}x{h} path {spp}spp (BASELINE size)", o, d, op, img, ref["image"], spp)
        st.close()


def test_ecosys_converged_render_masked(pair):
    """North-star gate 3 for C5: the reference's shipped ecosys render (1280x640) against the GPU render at the same
    size. The checkout lacks shape002/003.ply (the three trees, 8 instances), so the trees and the middle tree's
    reflection are masked out (test_oracle.ECOSYS_TREE_RECTS); the sky (directly visible environment: also the pin of
    the HDR loader rule) and the plant-covered ground are gated separately."""
    import os
    from PIL import Image
    from test_oracle import ecosys_masks
    _, d = pair("ecosys")
    ref8 = np.asarray(Image.open(os.path.join(orc.ROOT, "tests", "golden", "ref_ecosys_path.png")).convert("RGB"), np.float32)
    spp = 256
    p = _params(scene="ecosys", resolution=1280, samples=spp, batch=spp, sampler=1)
    st = trace.make_trace_state(d, p)
    assert (st.width, st.height) == (1280, 640)
    trace.trace_samples(st, d, None, None, p)
    st.sync()
    lin = st.image.reshape(st.height, st.width, 4)[..., :3]
    enc = jt.sceneio.image_to_srgb8(np.concatenate([lin, np.ones(lin.shape[:2] + (1,), np.float32)], axis=2))[..., :3]
    mine = np.asarray(Image.fromarray(enc.astype(np.uint8), "RGB").resize((ref8.shape[1], ref8.shape[0]), Image.BOX), np.float32)
    keep, sky = ecosys_masks(*ref8.shape[:2])
    ground = keep & ~sky
    dd = mine - ref8
    out = {}
    for tag, m in (("sky", sky), ("ground", ground)):
        out[tag] = (abs(mine[m].mean() - ref8[m].mean()), float(np.sqrt((dd[m] ** 2).mean())))
        print(f"\n[ecosys path {spp} spp, {tag}: {m.mean():.0%} of the image] mean diff {out[tag][0]:.2f}, RMSE {out[tag][1]:.2f} (8-bit units)")
    st.close()
    assert out["sky"][0] < 0.6 and out["sky"][1] < 1.5
    assert out["ground"][0] < 4.0 and out["ground"][1] < 12.0  # plant placement is identical, leaf-level detail is not converged in either


# ---- in-library multi-GPU (jt_group) on ONE device: logical shards {0, 0} ---------------------------------------------
def test_group_logical_shards_match_single_device(scenes):
    """jt_group with devices = [0, 0, 0]: three members on one GPU render disjoint sample sub-ranges into sum buffers and
    the fused reduce + finalize kernel merges them. The union is the single-device sample set, so image / albedo /
    normal agree to float-addition order and hits exactly; the 8-bit path agrees to +-1 LSB; counters add up."""
    sc, bvh, lights = scenes("features1")
    g = trace.DeviceGroup(sc, bvh, lights, [0, 0, 0])
    single = trace.DeviceScene(sc, bvh, lights, 0)
    try:
        gs = g.stats()
        assert gs["members"] == 3 and gs["distinct_devices"] == 1 and gs["peer_members"] == 0 and gs["staged_members"] == 0
        p = _params(scene="features1", resolution=200, samples=7, batch=1, sampler=1)
        a = trace.make_trace_state(g, p)
        assert isinstance(a, trace.GroupState)
        for _ in range(7):  # the reference's call pattern: samples / batch calls on ONE thread
            trace.trace_samples(a, g, None, None, p)
        assert a.samples == 7
        trace.trace_samples(a, g, None, None, p)  # no-op once samples == params.samples
        assert a.samples == 7
        b = trace.make_trace_state(single, p, accumulate=1)
        trace.trace_sample_range(b, single, p, 0, 7)
        b.sync()
        assert (a.width, a.height) == (b.width, b.height)
        assert np.allclose(a.image, b.image, rtol=2e-6, atol=1e-7)
        assert np.allclose(a.albedo, b.albedo, rtol=2e-6, atol=1e-7) and np.allclose(a.normal, b.normal, rtol=2e-6, atol=1e-6)
        assert np.array_equal(a.hits, b.hits)
        c = g.counters()
        assert c["camera_paths"] == a.width * a.height * 7 == single.counters()["camera_paths"]
        assert np.abs(a.srgb8().astype(np.int32) - b.srgb8().astype(np.int32)).max() <= 1
        # member scenes are reachable for the parity hooks
        m1 = g.member(1)
        rays = raygen.camera_rays(orc.Oracle(sc, bvh, lights), orc.make_params(resolution=64), 64, 27, 2000, seed=5)
        assert m1.intersect(rays, 0).tobytes() == single.intersect(rays, 0).tobytes()
        # ranges, reset, errors
        a.reset()
        assert a.samples == 0
        trace.trace_sample_range(a, g, p, 0, 3)
        trace.trace_sample_range(a, g, p, 3, 7)
        a.sync()
        assert np.allclose(a.image, b.image, rtol=2e-6, atol=1e-7) and np.array_equal(a.hits, b.hits)
        bad = trace.to_jt_params(p, 1)
        bad.bounces = 300  # rejected by every member -> surfaces from the next blocking call
        assert g.L.jt_group_trace_sample_range(g.h, a.h, C.byref(bad), 7, 8) == 0
        assert g.L.jt_group_synchronize(g.h) == -1 and b"member" in g.L.jt_last_error()
    finally:
        g.close()
        single.close()
    with pytest.raises(libmod.JtError):
        trace.DeviceGroup(sc, bvh, lights, [0, 99])


def test_group_main_drop_in(tmp_path):
    """Jtrace.main --gpu-devices 0,0: the reference's command line, sharded inside the library."""
    import os
    j = importlib.import_module("julia-raytracer_b200.jtrace")
    out = str(tmp_path / "out.png")
    scene = os.path.join(orc.ROOT, "assets", "scenes", "cornellbox.jtscene")
    r = j.main(f"--scene {scene} --sampler path --resolution 96 --samples 6 --batch 2 --output {out} --gpu-devices 0,0")
    assert os.path.exists(out) and r["image"].shape == (96, 96, 4) and r["counters"]["camera_paths"] == 96 * 96 * 6
    one = j.main(f"--scene {scene} --sampler path --resolution 96 --samples 6 --batch 2 --output {out}")
    assert np.allclose(r["image"], one["image"], rtol=1e-5, atol=1e-6)


def test_state_outlives_scene_and_bounce_limits(scenes):
    """a state whose scene was destroyed is an orphan (entry points fail cleanly, destroy is safe); the
    wavefront integrator rejects bounce counts its packed control word cannot hold."""
    sc, bvh, lights = scenes("cornellbox")
    d = trace.DeviceScene(sc, bvh, lights, 0)
    L = d.L
    jp = trace.to_jt_params(_params())
    h = C.c_void_p()
    assert L.jt_state_create(d.h, C.byref(jp), C.byref(h)) == 0
    jp.bounces = 255
    h2 = C.c_void_p()
    assert L.jt_state_create(d.h, C.byref(jp), C.byref(h2)) == -1 and b"bounces" in L.jt_last_error()
    jp.bounces = -1
    assert L.jt_state_create(d.h, C.byref(jp), C.byref(h2)) == -1
    jp.integrator = 1  # the megakernel takes any value (the reference loop simply runs 0 times)
    assert L.jt_state_create(d.h, C.byref(jp), C.byref(h2)) == 0
    L.jt_state_destroy(h2)
    L.jt_scene_destroy(d.h)
    d.h = None
    img = np.zeros(64 * 64 * 4, np.float32)
    assert L.jt_state_download(h, img.ctypes.data, None, None, None) == -1 and b"destroyed" in L.jt_last_error()
    assert L.jt_state_reset(h) == -1
    L.jt_state_destroy(h)  # no use-after-free


# ---- "any scenes/*": all 19 scenes the reference ships, on the GPU ---------------------------------------------------
OTHER_SCENES = ["bathroom1", "bathroom2", "coffee", "features2", "kitchen", "livingroom1", "livingroom2", "livingroom3",
                "materials2", "materials4", "shapes1", "shapes2", "staircase1", "staircase2"]


@pytest.mark.parametrize("name", OTHER_SCENES)
def test_every_other_shipped_scene_on_the_gpu(scenes, name):
    """The 14 scenes beyond the BASELINE five (packed with the missing-asset rule, assets/scenes/README.md). They reach
    branches the BASELINE scenes do not: `transparent` lobes (src/shading.jl:323-446), stochastic opacity incl. 0.0
    (src/trace.jl:356-364), a 1 024-face area light and 13 lights (src/trace.jl:1019-1044). Per scene:
    identical rays (ids / t / uv bit-exact in reference order, residual class only in wide mode) and a fixed sample set
    with both samplers (reference order: image, albedo, normal, hits, counters bit-exact; wide + wavefront: explained)."""
    sc, bvh, lights = scenes(name)
    o = orc.Oracle(sc, bvh, lights)
    d = trace.DeviceScene(sc, bvh, lights, 0)
    cam = jt.find_camera(sc, "")
    try:
        op = orc.make_params(camera=cam, resolution=192)
        w, h = o.make_state(op)
        rays = raygen.camera_rays(o, op, w, h, 60000, seed=41)
        allr = np.concatenate([rays, raygen.secondary_rays(rays, o.intersect(rays), seed=42)])
        ref = o.intersect(allr)
        r = raygen.compare_hits(d.intersect(allr, 1), ref)
        assert r["id_mismatch"] == 0 and r["t_mismatch"] == 0 and r["uv_mismatch"] == 0, ("reference-order mode", r)
        r = raygen.check_wide_vs_reference(d.intersect(allr, 0), ref)
        print(f"\n[{name}] wide-mode residual mismatches: {r['id_mismatch']} of {r['n']} rays")
        for sampler in (1, 2):
            op = orc.make_params(camera=cam, resolution=96, samples=3, batch=3, sampler=sampler, seed=43)
            w, h = o.make_state(op)
            o.counters(reset=True)
            o.trace_samples(op)
            want = o.get_state()
            oc = o.counters(reset=True)
            for traversal in ("reference", "wide"):
                p = _params(scene=name, camera=cam, resolution=96, samples=3, batch=3, sampler=sampler, gpu_seed=43,
                            gpu_traversal=traversal)
                st = trace.make_trace_state(d, p)
                d.counters(reset=True)
                trace.trace_samples(st, d, None, None, p)
                st.sync()
                c = d.counters()
                img = st.image.reshape(h, w, 4)
                if traversal == "reference":
                    assert np.array_equal(img, want["image"]), (name, sampler, float(np.abs(img - want["image"]).max()))
                    assert np.array_equal(st.albedo.reshape(h, w, 3), want["albedo"])
                    assert np.array_equal(st.normal.reshape(h, w, 3), want["normal"])
                    assert np.array_equal(st.hits.reshape(h, w), want["hits"])
                    assert (c["camera_paths"], c["scene_rays"], c["light_rays"]) == (oc["camera_paths"], oc["scene_rays"], oc["light_rays"])
                else:
                    wide_mode_check(f"{name} sampler {sampler} 96px", o, d, op, img, want["image"], 3)
                st.close()
    finally:
        d.close()


@pytest.mark.parametrize("res", [33, 50, 301])
def test_odd_image_sizes_are_bit_exact(pair, res):
    """Slot counts that are not multiples of 16 / 32 / the pipeline split (the flag compaction's scalar tail, padded shade
    queues, the 2-pipeline split at 301 x 301 = 90 601 pixels): reference-order mode still equals the oracle bit for bit."""
    o, d = pair("cornellbox")
    for sampler in (1, 2):
        op = orc.make_params(resolution=res, samples=3, batch=3, sampler=sampler, seed=9)
        w, h = o.make_state(op)
        o.trace_samples(op)
        ref = o.get_state()
        for traversal in ("reference", "wide"):
            p = _params(scene="cornellbox", resolution=res, samples=3, batch=3, sampler=sampler, gpu_seed=9,
                        gpu_traversal=traversal)
            st = trace.make_trace_state(d, p)
            trace.trace_samples(st, d, None, None, p)
            st.sync()
            img = st.image.reshape(h, w, 4)
            if traversal == "reference":
                assert np.array_equal(img, ref["image"]) and np.array_equal(st.hits.reshape(h, w), ref["hits"])
            else:
                wide_mode_check(f"cornellbox {res}px sampler {sampler}", o, d, op, img, ref["image"], 3)
            st.close()


@pytest.mark.gpu
def test_scheduling_counters_and_their_invariance(pair):
    """Work stealing and launch-tail suspension are live on the GPU (the counters say so) and change no bit: a render
    whose ranges are cut differently -- so that other slots trace other samples and other rays get parked -- ends in the
    same accumulators, and both equal the reference-order traversal (bit-exact vs the oracle elsewhere in this file)."""
    o, d = pair("features1")
    p = _params(resolution=640, samples=24, batch=24)
    d.counters(reset=True)
    a = trace.make_trace_state(d, p)
    trace.trace_sample_range(a, d, p, 0, 24)
    a.sync()
    c = d.counters(reset=True)
    assert c["camera_paths"] == a.width * a.height * 24
    assert c["stolen_samples"] > 0.05 * c["camera_paths"], c  # the sky pixels finish early and help the others
    assert c["resumed_rays"] > 0, c
    b = trace.make_trace_state(d, p)
    for lo, hi in ((0, 5), (5, 6), (6, 19), (19, 24)):
        trace.trace_sample_range(b, d, p, lo, hi)
        d.synchronize()
    b.sync()
    for k in ("image", "albedo", "normal", "hits"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    pr = _params(resolution=640, samples=24, batch=24, gpu_traversal="reference")
    r = trace.make_trace_state(d, pr)
    trace.trace_sample_range(r, d, pr, 0, 24)
    r.sync()
    differing = np.abs(a.image - r.image).max(axis=-1) > 1e-4
    assert differing.mean() <= WIDE_PIXEL_FRACTION_BOUND, differing.mean()
    assert np.array_equal(a.hits, r.hits) or differing.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornellbox", "classroom", "features1", "materials1", "ecosys", "synthetic_all", "bathroom1"])
def test_device_light_setup_is_bit_identical(scenes, name):
    """N4: make_trace_lights on the GPU (jt_lights_create: parallel element weights, sequential Float32 prefix sums replayed
    by a warp) returns the host builder's arrays bit for bit: same lights, same order, same CDFs (src/trace.jl:117-187)."""
    lm = importlib.import_module("julia-raytracer_b200.lights")
    sc, _, host = scenes(name)
    dev = lm.make_trace_lights_device(sc, 0)
    assert [(l.instance, l.environment) for l in dev] == [(l.instance, l.environment) for l in host]
    assert len(dev) > 0
    for a, b in zip(dev, host):
        assert a.elements_cdf.dtype == np.float32 and a.elements_cdf.shape == b.elements_cdf.shape
        assert a.elements_cdf.tobytes() == b.elements_cdf.tobytes(), (name, a.instance, a.environment)


@pytest.mark.gpu
def test_env_importance_flag_is_unbiased_and_uses_the_image(scenes):
    """Quirk Q8 behind a flag (JT_LIGHTS_ENV_LUMINANCE): the environment CDF follows max(R, G, B) of the texels instead
    of max(R, G, B, A = 1). Another sample set, the same expectation: the converged means agree."""
    lm = importlib.import_module("julia-raytracer_b200.lights")
    sc, bvh, host = scenes("features1")
    flagged = lm.make_trace_lights_device(sc, 0, env_luminance=True)
    env = [i for i, l in enumerate(host) if l.environment != -1][0]
    ref_w = np.diff(host[env].elements_cdf.astype(np.float64))
    new_w = np.diff(flagged[env].elements_cdf.astype(np.float64))
    assert (new_w >= -1e-6).all() and not np.allclose(ref_w / ref_w.sum(), new_w / new_w.sum(), atol=1e-9)
    p = _params(resolution=160, samples=256, batch=256)
    images = []
    for lights in (host, flagged):
        d = trace.DeviceScene(sc, bvh, lights, 0)
        try:
            st = trace.make_trace_state(d, p)
            trace.trace_samples(st, d, None, None, p)
            images.append(st.image[:, :3].astype(np.float64).copy())
        finally:
            d.close()
    a, b = images
    assert abs(a.mean() - b.mean()) <= 0.01 * a.mean(), (a.mean(), b.mean())
