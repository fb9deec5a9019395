"""Host-side logic (no GPU): CLI, scene model, the library's exports, and the product's C++ BVH builder
(`jt_make_bvh`) + Python light builder against the oracle's independent restatements."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

import orc

jt = orc.jt
bvhm = importlib.import_module("julia-raytracer_b200.bvh")
libmod = importlib.import_module("julia-raytracer_b200._lib")
cli = importlib.import_module("julia-raytracer_b200.cli")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(orc.ROOT, "include", "jtrace_b200.h")).read()
    import re
    declared = set(re.findall(r"JT_API\s+[\w\s\*]+?\b(jt_\w+)\s*\(", hdr))
    assert declared == set(libmod.EXPORTS), declared ^ set(libmod.EXPORTS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", libmod.LIB_PATH]).decode()
    for sym in declared:
        assert f" T {sym}" in out, sym
    L = libmod.lib()
    assert b"sm_100a" in L.jt_version()


def test_no_cpu_fallback_without_device(scenes):
    L = libmod.lib()
    if L.jt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    sc, bvh, lights = scenes("cornellbox")
    trace = importlib.import_module("julia-raytracer_b200.trace")
    with pytest.raises(libmod.JtError) as e:
        trace.DeviceScene(sc, bvh, lights, 0)
    assert e.value.code == -3 and "no CPU fallback" in e.value.message


def test_product_does_not_reference_oracle():
    pkg = os.path.join(orc.ROOT, "julia-raytracer_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")) or f == "Makefile":
                text = open(os.path.join(root, f), errors="replace").read()
                assert "liboracle" not in text and "orc_" not in text and "/oracle/" not in text, (root, f)


def test_cli_defaults_and_aliases():
    p = cli.parse_cli_args("--scene scenes/cornellbox/cornellbox.json")
    assert (p.resolution, p.samples, p.bounces, p.sampler, p.clamp, p.batch, p.bvhstacksize) == (1280, 512, 8, 1, 10, 1, 128)
    assert p.output == "tests/test_scene.png" and p.camera == ""
    p = cli.parse_cli_args("--scene x.json --sampler naive --noparallel true --resolution 720 --samples 16")
    assert p.sampler == 2 and p.noparallel is True and p.resolution == 720
    assert cli.parse_cli_args("--scene x.json --shader naive").sampler == 2  # north-star spelling
    assert cli.parse_cli_args("--scene x.json --sampler bogus").sampler == 1  # src/cli.jl:111-116
    with pytest.raises(ValueError):
        cli.parse_cli_args("--scene x.json --clamp 2.5")  # Params.clamp::Int
    with pytest.raises(SystemExit):
        cli.parse_cli_args("--samples 4")  # --scene is required


def test_image_size_rounding():
    S = jt.scene
    cam = S.CameraData(frame=S.IDENTITY_FRAME, aspect=np.float32(2.4000000953674316))
    assert S.image_size(cam, 1280) == (1280, 533)
    cam.aspect = np.float32(1.7777777910232544)
    assert S.image_size(cam, 1280) == (1280, 720)
    cam.aspect = np.float32(0.5)
    assert S.image_size(cam, 101) == (50, 101)  # round half to even: 50.5 -> 50


def test_find_camera():
    S = jt.scene
    mk = lambda n: S.CameraData(frame=S.IDENTITY_FRAME, name=n)
    sc = S.SceneData([mk("a"), mk("camera"), mk("default")], None, None, [], [], None)
    assert S.find_camera(sc, "") == 3 and S.find_camera(sc, "a") == 1 and S.find_camera(sc, "zzz") == 3
    sc = S.SceneData([mk("x"), mk("y")], None, None, [], [], None)
    assert S.find_camera(sc, "") == 1


@pytest.mark.parametrize("name", ["cornellbox", "materials1", "features1", "classroom", "synthetic_all"])
@pytest.mark.parametrize("hq", [False, True])
def test_host_bvh_matches_oracle_builder(scenes, name, hq):
    if hq and name in ("features1", "classroom"):
        pytest.skip("SAH build of the large scenes is slow on CPU; covered by the smaller ones")
    sc, _, _ = scenes(name)
    b = bvhm.make_scene_bvh(sc, high_quality=hq)
    o = orc.Oracle(sc, high_quality=hq)
    n, p = o.get_bvh(0)
    assert n.tobytes() == b.bvh.nodes.tobytes() and np.array_equal(p, b.bvh.primitives)
    for k in range(len(sc.shapes)):
        n, p = o.get_bvh(k + 1)
        assert n.tobytes() == b.shapes[k].nodes.tobytes(), (name, k)
        assert np.array_equal(p, b.shapes[k].primitives)


def test_make_bvh_edge_cases():
    t = bvhm.make_bvh(np.zeros((0, 6), np.float32))
    assert len(t.nodes) == 1 and t.nodes[0]["num"] == 0 and not t.nodes[0]["internal"]
    one = np.asarray([[0, 0, 0, 1, 1, 1]], np.float32)
    t = bvhm.make_bvh(one)
    assert len(t.nodes) == 1 and t.nodes[0]["num"] == 1 and t.primitives.tolist() == [1]
    # identical centroids -> median split on axis 1 (src/bvh.jl:198-200)
    same = np.tile(one, (9, 1))
    t = bvhm.make_bvh(same)
    assert t.nodes[0]["internal"] and t.nodes[0]["axis"] == 1
    assert sorted(t.primitives.tolist()) == list(range(1, 10))
    leaves = [n for n in t.nodes if not n["internal"]]
    assert sum(int(n["num"]) for n in leaves) == 9 and max(int(n["num"]) for n in leaves) <= 4


@pytest.mark.parametrize("name", ["cornellbox", "materials1", "classroom", "synthetic_all"])
def test_lights_match_oracle(scenes, name):
    sc, bvh, lights = scenes(name)
    o = orc.Oracle(sc)
    ol = o.get_lights()
    assert [(l.instance, l.environment) for l in lights] == [(a, b) for a, b, _ in ol]
    for l, (_, _, cdf) in zip(lights, ol):
        assert len(cdf) == len(l.elements_cdf)
        if l.instance != -1:
            assert np.array_equal(cdf, l.elements_cdf)  # IEEE basic ops only: bit-exact
        else:
            assert np.allclose(cdf, l.elements_cdf, rtol=2e-6)  # sin() differs by libm


def test_packed_scene_roundtrip(tmp_path, scenes):
    sc, _, _ = scenes("synthetic_all")
    f = str(tmp_path / "s.jtscene")
    jt.save_packed(sc, f)
    back = jt.load_packed(f)
    assert back.instances.tobytes() == sc.instances.tobytes() and back.materials.tobytes() == sc.materials.tobytes()
    for a, b in zip(sc.shapes, back.shapes):
        assert np.array_equal(a.positions, b.positions) and np.array_equal(a.triangles, b.triangles)
        assert np.array_equal(a.quads, b.quads) and np.array_equal(a.colors, b.colors)
    for a, b in zip(sc.textures, back.textures):
        assert (a.width, a.height, a.linear) == (b.width, b.height, b.linear)


def test_ply_loader_quads_and_fans(tmp_path):
    # triangle + quad + pentagon in one shape -> stored as quads (src/shape.jl:302-369)
    import struct
    verts = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (2, 0, 0), (2, 1, 0), (3, 0.5, 0)]
    faces = [[0, 1, 2], [0, 1, 2, 3], [1, 4, 6, 5, 2]]
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
           "property float z\nproperty float u\nproperty float v\nelement face %d\n"
           "property list uchar int vertex_indices\nend_header\n" % (len(verts), len(faces))).encode()
    body = b"".join(struct.pack("<5f", *v, 0.25, 0.75) for v in verts)
    body += b"".join(struct.pack("<B%di" % len(f), len(f), *f) for f in faces)
    p = tmp_path / "m.ply"
    p.write_bytes(hdr + body)
    sh = jt.sceneio.load_shape(str(p))
    assert len(sh.triangles) == 0
    assert sh.quads.tolist() == [[1, 2, 3, 3], [1, 2, 3, 4], [2, 5, 7, 7], [2, 7, 6, 6], [2, 6, 3, 3]]
    assert np.allclose(sh.texcoords[:, 1], 0.25)  # v flipped: 1 - 0.75


def test_format_seconds():
    j = importlib.import_module("julia-raytracer_b200.jtrace")
    assert j.format_seconds(1.5) == "01.500" and j.format_seconds(75.25) == "01:15.250"
    assert j.format_seconds(3700.0) == "01:01:40.000"


def _header_prototypes():
    """name -> number of parameters, parsed from include/jtrace_b200.h."""
    import re
    hdr = open(os.path.join(orc.ROOT, "include", "jtrace_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"JT_API\s+[\w\s\*]+?\b(jt_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def _split_top_level(text):
    """Split on commas that are not nested inside (), [] or {}."""
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _julia_ccalls(src):
    """[(symbol, n_argtypes, n_args)] for every ccall((:sym, LIB), ret, (argtypes...), args...) in a Julia source."""
    import re
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\)", src):
        i, depth = m.start() + len("ccall"), 0
        j = i
        while True:  # matching parenthesis of the ccall
            if src[j] == "(":
                depth += 1
            elif src[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        parts = _split_top_level(src[i + 1:j])
        assert len(parts) >= 3 and parts[2].startswith("("), (m.group(1), parts)
        inner = parts[2][1:-1].strip()
        types = [t for t in _split_top_level(inner) if t]
        out.append((m.group(1), len(types), len(parts) - 3))
    return out


def test_julia_glue_matches_the_header():
    """The Julia binding cannot run here (no julia in the image): check it at the text level instead. Every ccall
    names a symbol the header declares, passes as many argument types as the C prototype has parameters, and as many
    values as types; the reference-facing surface INTEGRATION.md documents exists with those signatures."""
    import re
    protos = _header_prototypes()
    src = open(os.path.join(orc.ROOT, "julia-raytracer_b200", "julia", "JtraceB200.jl")).read()
    calls = _julia_ccalls(src)
    assert len(calls) >= 12
    for sym, ntypes, nargs in calls:
        assert sym in protos, f"ccall of undeclared symbol {sym}"
        assert ntypes == protos[sym], f"{sym}: {ntypes} argument types, C prototype has {protos[sym]}"
        assert nargs == ntypes, f"{sym}: {nargs} values for {ntypes} types"
    used = {c[0] for c in calls}
    assert {"jt_scene_create", "jt_state_create", "jt_trace_samples", "jt_state_download", "jt_group_create",
            "jt_group_state_create", "jt_group_trace_samples", "jt_group_state_download", "jt_last_error"} <= used
    # JtParams mirrors jt_params field for field (names and order)
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(orc.ROOT, "include", "jtrace_b200.h")).read(), flags=re.S)
    body = re.search(r"typedef struct \{([^}]*)\} jt_params;", hdr, flags=re.S).group(1)
    c_fields = [f for f in re.findall(r"\b(?:int32_t|uint64_t)\s+(\w+)", body)]
    jl = re.search(r"struct JtParams\n(.*?)\nend", src, flags=re.S).group(1)
    jl_fields = re.findall(r"(\w+)::(?:Int32|UInt64)", jl)
    assert jl_fields[:len(c_fields) - 1] == c_fields[:-1] and len(jl_fields) == len(c_fields) - 1 + 6  # _reserved[6]
    # the surface INTEGRATION.md's diff relies on
    assert "import ..Trace: trace_samples" in src
    assert re.search(r"function trace_samples\(state::TraceState, g::GpuScene, bvh, lights, params::Params", src)
    assert re.search(r"function gpu_scene\(scene::SceneData, bvh::SceneBvh, lights::TraceLights, state::TraceState, params::Params", src)
    integ = open(os.path.join(orc.ROOT, "INTEGRATION.md")).read()
    assert "gpu_scene(scene, bvh, lights, state, params" in integ and "gscene," in integ


def test_header_structs_match_ctypes_mirror():
    """sizeof of every POD struct in the header, compiled by gcc, equals the ctypes mirror's."""
    import tempfile
    A = orc.A
    names = ["jt_frame", "jt_bvh_node", "jt_bvh_desc", "jt_instance", "jt_material", "jt_environment", "jt_camera",
             "jt_texture_desc", "jt_shape_desc", "jt_light_desc", "jt_scene_desc", "jt_params", "jt_ray", "jt_hit",
             "jt_counters", "jt_scene_stats", "jt_group_stats"]
    prog = '#include "jtrace_b200.h"\n#include <stdio.h>\nint main(void){' + "".join(
        f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(orc.ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    for n, s in zip(names, sizes):
        assert C.sizeof(getattr(A, n)) == s, (n, s, C.sizeof(getattr(A, n)))


def test_wide_bvh_cache_round_trip(tmp_path, scenes, monkeypatch):
    """N1: jt_stage_scene stores the finished wide BVH under a hash of the builder's inputs and finds it again; a damaged
    or foreign file is ignored. Exercised through the host emulation (same jt_stage.cpp as the library)."""
    import emu
    import raygen
    import orc
    monkeypatch.setenv("JT_BVH_CACHE_DIR", str(tmp_path))
    sc, bvh, lights = scenes("features1")
    a = emu.Emu(sc, bvh, lights)
    assert a.stats()["from_cache"] == 0
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].startswith("jtwide_") and files[0].endswith(".bin")
    b = emu.Emu(sc, bvh, lights)
    assert b.stats()["from_cache"] == 1 and b.stats() == {**a.stats(), "from_cache": 1}
    o = orc.Oracle(sc, bvh, lights)
    p = orc.make_params(resolution=96)
    w, h = o.make_state(p)
    rays = raygen.camera_rays(o, p, w, h, 4000, seed=9)
    rays = np.concatenate([rays, raygen.secondary_rays(rays, o.intersect(rays), seed=10)])
    ha, hb = a.intersect(rays, 0), b.intersect(rays, 0)
    assert ha.tobytes() == hb.tobytes()
    # another scene gets another file; a truncated file is rebuilt, not trusted
    sc2, bvh2, lights2 = scenes("cornellbox")
    emu.Emu(sc2, bvh2, lights2)
    assert len(os.listdir(tmp_path)) == 2
    path = os.path.join(tmp_path, files[0])
    data = open(path, "rb").read()
    open(path, "wb").write(data[: len(data) // 2])
    c = emu.Emu(sc, bvh, lights)
    assert c.stats()["from_cache"] == 0 and os.path.getsize(path) == len(data)
    monkeypatch.delenv("JT_BVH_CACHE_DIR")
    assert emu.Emu(sc, bvh, lights).stats()["from_cache"] == 0


def test_collapse_rule_changes_the_tree_not_the_hits(scenes, monkeypatch):
    """The wide BVH's topology is free (tie ranks come from the reference tree): the dynamic-programming collapse and the
    greedy one give different trees and the same hits, bit for bit."""
    import emu
    import raygen
    sc, bvh, lights = scenes("features1")
    o = orc.Oracle(sc, bvh, lights)
    p = orc.make_params(resolution=96)
    w, h = o.make_state(p)
    rays = raygen.camera_rays(o, p, w, h, 6000, seed=21)
    rays = np.concatenate([rays, raygen.secondary_rays(rays, o.intersect(rays), seed=22)])
    dp = emu.Emu(sc, bvh, lights)
    monkeypatch.setenv("JT_COLLAPSE", "greedy")
    greedy = emu.Emu(sc, bvh, lights)
    assert dp.stats()["wide_nodes"] < greedy.stats()["wide_nodes"]
    assert dp.intersect(rays, 0).tobytes() == greedy.intersect(rays, 0).tobytes()
    raygen.check_wide_vs_reference(dp.intersect(rays, 0), o.intersect(rays))
