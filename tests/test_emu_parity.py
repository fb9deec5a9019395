"""No-GPU proof of the CUDA sources' logic: the device headers (traversal, shading, integrators,
wide-BVH builder) compiled for the host (tests/emu) must agree bit-for-bit with the independent CPU
oracle on identical rays and identical counter-based sample sets. The same comparisons run through
the real kernels and the C ABI in test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

import emu
import orc
import raygen

SCENES = ["cornellbox", "materials1", "features1", "classroom", "synthetic_all", "synthetic_closed", "synthetic_one"]


@pytest.fixture(scope="module")
def pair(scenes):
    cache = {}

    def get(name):
        if name not in cache:
            sc, bvh, lights = scenes(name)
            cache[name] = (orc.Oracle(sc, bvh, lights), emu.Emu(sc, bvh, lights))
        return cache[name]

    return get


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("traversal", [1, 0, 3])  # 3 = the persistent-warp state machine, single-lane
def test_identical_rays(pair, name, traversal):
    o, e = pair(name)
    p = orc.make_params(resolution=128)
    w, h = o.make_state(p)
    rays = raygen.camera_rays(o, p, w, h, 20000, seed=3)
    allr = np.concatenate([rays, raygen.secondary_rays(rays, o.intersect(rays), seed=4)])
    got, ref = e.intersect(allr, traversal), o.intersect(allr)
    if traversal == 1:  # reference-order mode: bit-exact, no exceptions
        r = raygen.compare_hits(got, ref)
        assert r["id_mismatch"] == 0 and r["t_mismatch"] == 0 and r["uv_mismatch"] == 0, r
    else:
        raygen.check_wide_vs_reference(got, ref)


@pytest.mark.parametrize("name", ["cornellbox", "features1", "synthetic_all"])
def test_instance_probes(pair, scenes, name):
    o, e = pair(name)
    sc, _, lights = scenes(name)
    p = orc.make_params(resolution=96)
    w, h = o.make_state(p)
    rays = raygen.camera_rays(o, p, w, h, 5000, seed=5)
    rng = np.random.default_rng(6)
    inst = rng.integers(1, len(sc.instances) + 1, len(rays))
    ref = o.intersect_instance(rays, inst)
    r = raygen.compare_hits(e.intersect_instance(rays, inst, 1), ref)
    assert r["id_mismatch"] == 0 and r["t_mismatch"] == 0 and r["uv_mismatch"] == 0, r
    raygen.check_wide_vs_reference(e.intersect_instance(rays, inst, 0), ref)


CASES = [("cornellbox", 2, {}), ("cornellbox", 1, {}), ("features1", 1, {}), ("features1", 2, {}),
         ("materials1", 1, {}), ("classroom", 1, {}), ("synthetic_all", 1, {}), ("synthetic_all", 2, {}),
         ("synthetic_all", 1, dict(nocaustics=1, tentfilter=1, envhidden=1)), ("synthetic_closed", 1, {}),
         ("synthetic_closed", 2, dict(envhidden=1)), ("synthetic_one", 1, {}),
         ("synthetic_all", 1, dict(accumulate=1)), ("synthetic_all", 1, dict(bounces=2, clamp=1))]


@pytest.mark.parametrize("name,sampler,extra", CASES)
@pytest.mark.parametrize("traversal", [1, 0])
@pytest.mark.parametrize("wavefront", [False, True])
def test_fixed_sample_set_images(pair, name, sampler, extra, traversal, wavefront):
    o, e = pair(name)
    p = orc.make_params(resolution=64, samples=3, batch=3, sampler=sampler, traversal=traversal, seed=11, **extra)
    w, h = o.make_state(p)
    o.trace_samples(p)
    ref = o.get_state()
    got = e.trace(p, w, h, 0, 3, wavefront=wavefront)
    c = o.counters(reset=True)
    img = got["image"] * np.float32(1.0 / 3.0) if extra.get("accumulate") else got["image"]
    if traversal == 1:
        assert np.array_equal(img, ref["image"])
        assert np.array_equal(got["hits"], ref["hits"])
        assert (got["scene_rays"], got["light_rays"]) == (c["scene_rays"], c["light_rays"])
    else:  # wide mode: identical except where a ray falls in the residual class (rate <= 1e-4 per ray)
        differing = np.abs(img - ref["image"]).max(axis=-1) > 1e-4
        assert differing.mean() <= 2e-3, differing.mean()


@pytest.mark.parametrize("name,sampler,extra", CASES)
def test_wide_wavefront_equals_wide_megakernel(pair, name, sampler, extra):
    """Same traversal on both sides, so there is no residual-class slack: the wavefront reorganisation (queues,
    MIS weights finished inside the shade kernel when no light walk is needed, CDF guide tables) must reproduce the
    per-thread integrator bit for bit, ray counters included."""
    o, e = pair(name)
    p = orc.make_params(resolution=64, samples=3, batch=3, sampler=sampler, traversal=0, seed=23, **extra)
    w, h = o.make_state(p)
    a = e.trace(p, w, h, 0, 3, wavefront=False)
    b = e.trace(p, w, h, 0, 3, wavefront=True)
    for k in ("image", "albedo", "normal", "hits"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["scene_rays"], a["light_rays"]) == (b["scene_rays"], b["light_rays"])


def test_cdf_guide_table_returns_the_reference_index():
    """sample_discrete_light (guide-table bracket + the reference bisection) against the plain bisection of
    src/sampling.jl:33-56 on adversarial CDFs: plateaus, duplicates, huge and tiny totals, r at the clamp edges."""
    import ctypes as C
    L = emu.lib()
    L.emu_sample_discrete.restype = C.c_int
    L.emu_sample_discrete.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(3)
    cases = []
    for n in (64, 65, 1000, 131072):
        w = rng.random(n).astype(np.float32)
        cases.append(w)                                                    # generic
        cases.append(np.where(rng.random(n) < 0.7, 0, w).astype(np.float32))  # long plateaus
        cases.append((w * np.float32(1e-9)).astype(np.float32))             # total below the 1e-5 clamp offset
        cases.append((w * np.float32(1e12)).astype(np.float32))
        cases.append(np.sin(np.pi * (np.arange(n) // 8 + 0.5) / (n // 8 + 1)).astype(np.float32))  # sky-like rows
    used = 0
    for w in cases:
        cdf = np.cumsum(w, dtype=np.float32)  # sequential Float32 sums, like make_trace_lights
        r = np.concatenate([rng.random(20000).astype(np.float32), np.float32([0, 1e-8, 0.5, 0.99999994]),
                            (cdf[rng.integers(0, len(cdf), 2000)] / cdf[-1]).astype(np.float32)])
        plain = np.zeros(len(r), np.int32)
        guided = np.zeros(len(r), np.int32)
        k = L.emu_sample_discrete(cdf.ctypes.data, len(cdf), r.ctypes.data, len(r), plain.ctypes.data, guided.ctypes.data)
        used += k > 0
        assert np.array_equal(plain, guided)
    assert used >= len(cases) // 2  # the table was really in play


@pytest.mark.parametrize("name", ["features1", "classroom", "synthetic_all"])
@pytest.mark.parametrize("braid", [1, 16])
def test_braided_and_flattened_instances_return_the_same_hits(scenes, name, braid, monkeypatch):
    """Instances with a non-identity frame can be opened into the top-level tree (JT_BRAID_MAX = 1: every triangle
    becomes a flattened record tested in instance space; 16: entry records at BLAS sub-trees of <= 16 triangles).
    The topology only decides which boxes are visited: hits must equal the two-level walk's bit for bit and obey the
    wide-vs-reference contract."""
    sc, b, lights = scenes(name)
    o = orc.Oracle(sc, b, lights)
    monkeypatch.setenv("JT_BRAID_MAX", "0")
    plain = emu.Emu(sc, b, lights)
    monkeypatch.setenv("JT_BRAID_MAX", str(braid))
    monkeypatch.setenv("JT_BRAID_MIN_INSTANCES", "1")
    opened = emu.Emu(sc, b, lights)
    assert plain.stats()["flattened"] == 0 and opened.stats()["flattened"] > 0
    p = orc.make_params(resolution=128)
    w, h = o.make_state(p)
    cur = raygen.camera_rays(o, p, w, h, 20000, seed=31)
    rays = [cur]
    for g in range(2):
        cur = raygen.secondary_rays(cur, o.intersect(cur), seed=32 + g)
        rays.append(cur)
    rays = np.concatenate(rays)
    ref = o.intersect(rays)
    for mode in (0, 3):  # plain wide walk, persistent-warp state machine
        a, c = plain.intersect(rays, mode), opened.intersect(rays, mode)
        assert np.array_equal(a.view(np.uint8), c.view(np.uint8))
        raygen.check_wide_vs_reference(c, ref)
    pp = orc.make_params(resolution=48, samples=2, batch=2, sampler=1, traversal=0, seed=5)
    w, h = o.make_state(pp)
    x, y = plain.trace(pp, w, h, 0, 2, wavefront=True), opened.trace(pp, w, h, 0, 2, wavefront=True)
    assert np.array_equal(x["image"], y["image"]) and np.array_equal(x["hits"], y["hits"])


def test_parallel_wide_bvh_build_is_deterministic(scenes, monkeypatch):
    """The task-parallel SAH build + parallel collapse (used for big scenes) must give the same wide BVH as the
    single-threaded build: same node / record counts, same per-ray work, bit-identical hits. Forced onto a small scene
    by lowering the record threshold; flattening is switched on so the braid / flatten record generation runs too."""
    sc, b, lights = scenes("features1")
    o = orc.Oracle(sc, b, lights)
    monkeypatch.setenv("JT_BRAID_MIN_INSTANCES", "1")
    monkeypatch.setenv("JT_BUILD_THREADS", "1")
    monkeypatch.setenv("JT_BUILD_PARALLEL_MIN", "1000000000")
    serial = emu.Emu(sc, b, lights)
    monkeypatch.setenv("JT_BUILD_THREADS", "6")
    monkeypatch.setenv("JT_BUILD_PARALLEL_MIN", "1000")
    parallel = emu.Emu(sc, b, lights)
    assert serial.stats() == parallel.stats()
    p = orc.make_params(resolution=128)
    w, h = o.make_state(p)
    rays = raygen.camera_rays(o, p, w, h, 20000, seed=41)
    rays = np.concatenate([rays, raygen.secondary_rays(rays, o.intersect(rays), seed=42)])
    emu.wide_counts()
    a = serial.intersect(rays, 0)
    ca = emu.wide_counts()
    c = parallel.intersect(rays, 0)
    cc = emu.wide_counts()
    assert np.array_equal(a.view(np.uint8), c.view(np.uint8)) and ca == cc


@pytest.mark.parametrize("name,sampler", [("cornellbox", 1), ("features1", 1), ("classroom", 1), ("synthetic_all", 1),
                                          ("synthetic_all", 2), ("materials1", 2)])
@pytest.mark.parametrize("every", [1, 3, 7])
def test_suspended_rays_resume_to_the_same_image(pair, name, sampler, every):
    """The extend kernel parks the stragglers of a launch tail and the next launch resumes them (jt_dev_persist.cuh).
    Forced here for EVERY ray, every `every` traversal iterations: accumulators and ray counts must not change."""
    o, e = pair(name)
    p = orc.make_params(resolution=48, samples=2, batch=2, sampler=sampler, traversal=0, seed=23)
    w, h = o.make_state(p)
    plain = e.trace(p, w, h, 0, 2, wavefront=True)
    emu.set_suspend_every(every)
    emu.resumed_rays()
    try:
        parked = e.trace(p, w, h, 0, 2, wavefront=True)
        resumed = emu.resumed_rays()
    finally:
        emu.set_suspend_every(0)
    # small trees finish in fewer than `every` iterations; with every = 1 all but the one-step rays are parked
    if every == 1:
        assert resumed > plain["scene_rays"] // 2, "the park / resume path did not run"
    elif name in ("classroom", "features1"):
        assert resumed > 0
    for k in ("image", "albedo", "normal", "hits"):
        assert np.array_equal(parked[k], plain[k]), k
    assert parked["scene_rays"] == plain["scene_rays"] and parked["light_rays"] == plain["light_rays"]


@pytest.mark.parametrize("name,sampler,traversal", [("features1", 1, 0), ("features1", 1, 1), ("materials1", 2, 0),
                                                    ("synthetic_all", 1, 0), ("classroom", 1, 0)])
def test_work_stealing_keeps_the_sample_order(pair, name, sampler, traversal):
    """Slots whose pixel has no sample left trace samples of other pixels (wf_regen_slot); the per-pixel commit counter
    keeps the accumulation in sample order, so a longer range with many stolen samples still equals the oracle."""
    o, e = pair(name)
    p = orc.make_params(resolution=40, samples=12, batch=12, sampler=sampler, traversal=traversal, seed=5)
    w, h = o.make_state(p)
    o.trace_samples(p)
    ref = o.get_state()
    emu.stolen_samples()
    got = e.trace(p, w, h, 0, 12, wavefront=True)
    stolen = emu.stolen_samples()
    assert stolen > 0.02 * w * h * 12, stolen
    if traversal == 1:
        for k in ("image", "albedo", "normal", "hits"):
            assert np.array_equal(got[k], ref[k]), k
    else:
        assert (np.abs(got["image"] - ref["image"]).max(axis=-1) > 1e-4).mean() <= 2e-3
    # split ranges (every range restarts the claim / commit counters) give the same accumulators
    a = e.trace(p, w, h, 0, 12, wavefront=True)
    assert np.array_equal(a["image"], got["image"])

