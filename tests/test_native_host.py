"""N2 (+ the host halves of N1 / N4) behind the C ABI: jt_host_scene_load / _build / _desc must hand the library the
same bytes as the Python host mirror (sceneio.py = src/sceneio.jl:25-93 + src/shape.jl:78-124, :302-446 +
src/scene.jl:164-189; bvh.py = src/bvh.jl:66-136; lights.py = src/trace.jl:117-187).

* a synthetic scene written on the fly (ascii + binary PLY, quads / n-gons, RGB / RGBA / palette / grey PNGs, an RLE and
  a flat Radiance file, a missing shape and a missing texture) runs everywhere;
* every scene the reference ships (19) runs where /root/reference exists (this container, not the GPU box).
No GPU: nothing here creates a device scene."""
import ctypes as C
import glob
import importlib
import json
import os
import struct
import zlib

import numpy as np
import pytest

A = importlib.import_module("julia-raytracer_b200._abi")
_lib = importlib.import_module("julia-raytracer_b200._lib")
sio = importlib.import_module("julia-raytracer_b200.sceneio")
scn = importlib.import_module("julia-raytracer_b200.scene")
bvhm = importlib.import_module("julia-raytracer_b200.bvh")
lm = importlib.import_module("julia-raytracer_b200.lights")
fl = importlib.import_module("julia-raytracer_b200.flatten")

REF_SCENES = "/root/reference/scenes"
ref_names = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(REF_SCENES, "*", "*.json")))


def _arr(ptr, count, dtype):
    """Copy `count` records of `dtype` from a C pointer (None -> empty)."""
    dtype = np.dtype(dtype)
    if not ptr or count == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (count * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def _bvh(d):
    return {"nodes": _arr(d.nodes, d.num_nodes, scn.BVHNODE_DTYPE), "prims": _arr(d.primitives, d.num_primitives, "<i8")}


CAMERA_DTYPE = np.dtype([("frame", "<f4", (12,)), ("orthographic", "<i4"), ("lens", "<f4"), ("film", "<f4"),
                         ("aspect", "<f4"), ("focus", "<f4"), ("aperture", "<f4")])


def desc_to_dict(d: A.jt_scene_desc) -> dict:
    """Every byte a jt_scene_desc points at, as numpy arrays."""
    out = {
        "cameras": _arr(d.cameras, d.num_cameras, CAMERA_DTYPE),
        "instances": _arr(d.instances, d.num_instances, np.dtype((np.void, 64))),
        "environments": _arr(d.environments, d.num_environments, scn.ENVIRONMENT_DTYPE),
        "materials": _arr(d.materials, d.num_materials, np.dtype((np.void, 104))),
        "lut": _arr(d.srgb_to_rgb_lut, 256, "<f4"),
        "bvh": _bvh(d.bvh),
        "shapes": [], "textures": [], "lights": [],
    }
    shapes = C.cast(d.shapes, C.POINTER(A.jt_shape_desc))
    for i in range(d.num_shapes):
        s = shapes[i]
        out["shapes"].append({
            "positions": _arr(s.positions, s.num_positions * 3, "<f4"), "normals": _arr(s.normals, s.num_normals * 3, "<f4"),
            "texcoords": _arr(s.texcoords, s.num_texcoords * 2, "<f4"), "colors": _arr(s.colors, s.num_colors * 4, "<f4"),
            "triangles": _arr(s.triangles, s.num_triangles * 3, "<i8"), "quads": _arr(s.quads, s.num_quads * 4, "<i8"),
            "bvh": _bvh(s.bvh)})
    texs = C.cast(d.textures, C.POINTER(A.jt_texture_desc))
    for i in range(d.num_textures):
        t = texs[i]
        n = t.width * t.height * 4
        out["textures"].append({"size": (t.width, t.height, t.linear), "f": _arr(t.pixelsf, n if t.pixelsf else 0, "<f4"),
                                "b": _arr(t.pixelsb, n if t.pixelsb else 0, "u1")})
    lts = C.cast(d.lights, C.POINTER(A.jt_light_desc))
    for i in range(d.num_lights):
        out["lights"].append({"ids": (lts[i].instance, lts[i].environment),
                              "cdf": _arr(lts[i].elements_cdf, lts[i].num_elements, "<f4")})
    return out


def assert_same(a, b, path="desc"):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            assert_same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, list):
        assert len(a) == len(b), f"{path}: {len(a)} vs {len(b)} entries"
        for i, (x, y) in enumerate(zip(a, b)):
            assert_same(x, y, f"{path}[{i}]")
    elif isinstance(a, np.ndarray):
        assert a.shape == b.shape and a.dtype == b.dtype, f"{path}: {a.shape} {a.dtype} vs {b.shape} {b.dtype}"
        if a.dtype.names:  # records: field by field (alignment padding carries no data)
            for name in a.dtype.names:
                assert np.ascontiguousarray(a[name]).tobytes() == np.ascontiguousarray(b[name]).tobytes(), f"{path}.{name}: bytes differ"
        else:
            assert a.tobytes() == b.tobytes(), f"{path}: bytes differ"
    else:
        assert a == b, f"{path}: {a} vs {b}"


class NativeScene:
    def __init__(self, path):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        _lib.check(self.L.jt_host_scene_load(path.encode(), C.byref(self.h)))

    def build(self, hq=False):
        _lib.check(self.L.jt_host_scene_build(self.h, int(hq)))

    def desc(self):
        p = C.POINTER(A.jt_scene_desc)()
        _lib.check(self.L.jt_host_scene_desc(self.h, C.byref(p)))
        return p.contents

    def notes(self):
        return [self.L.jt_host_scene_note(self.h, i).decode() for i in range(self.L.jt_host_scene_num_notes(self.h))]

    def find_camera(self, name):
        c = C.c_int32()
        _lib.check(self.L.jt_host_scene_find_camera(self.h, name.encode(), C.byref(c)))
        return c.value

    def close(self):
        if self.h:
            self.L.jt_host_scene_destroy(self.h)
            self.h = C.c_void_p()


def compare_scene(path, hq=False):
    py = sio.load_scene(path)
    nat = NativeScene(path)
    try:
        flat0 = fl.FlatScene(py, None, None)  # keep the owner alive while its desc is read
        assert_same(desc_to_dict(nat.desc()), desc_to_dict(flat0.desc), "loaded")
        assert nat.notes() == py.notes
        for name in ("", "default", "nonexistent"):
            assert nat.find_camera(name) == scn.find_camera(py, name)
        nat.build(hq)
        flat = fl.FlatScene(py, bvhm.make_scene_bvh(py, hq), lm.make_trace_lights(py))
        assert_same(desc_to_dict(nat.desc()), desc_to_dict(flat.desc), "built")
    finally:
        nat.close()
    return py


# ---------------------------------------------------------------------------------------------------------------------
# a synthetic scene directory that exercises every decoder branch
# ---------------------------------------------------------------------------------------------------------------------
def _png(path, w, h, ctype, rows, plte=None, trns=None, filters=None):
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    rows = np.asarray(rows, np.uint8).reshape(h, w * ch)
    raw = bytearray()
    prev = np.zeros(w * ch, np.int32)
    for y in range(h):
        cur = rows[y].astype(np.int32)
        f = (filters[y % len(filters)] if filters else 0)
        left = np.concatenate([np.zeros(ch, np.int32), cur[:-ch]])
        upleft = np.concatenate([np.zeros(ch, np.int32), prev[:-ch]])
        if f == 0:
            enc = cur
        elif f == 1:
            enc = cur - left
        elif f == 2:
            enc = cur - prev
        elif f == 3:
            enc = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            enc = cur - pred
        raw.append(f)
        raw += (enc & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)

    z = zlib.compress(bytes(raw))
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
    if plte is not None:
        data += chunk(b"PLTE", bytes(plte))
    if trns is not None:
        data += chunk(b"tRNS", bytes(trns))
    half = len(z) // 2
    data += chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:]) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


def _hdr(path, w, h, rgbe, rle):
    rgbe = np.asarray(rgbe, np.uint8).reshape(h, w, 4)
    body = bytearray()
    for y in range(h):
        if rle:
            body += bytes([2, 2, w >> 8, w & 255])
            for c in range(4):
                row = rgbe[y, :, c]
                x = 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and row[x + run] == row[x]:
                        run += 1
                    if run >= 3:
                        body += bytes([128 + run, int(row[x])])
                        x += run
                    else:
                        n = min(w - x, 5)
                        body += bytes([n]) + row[x:x + n].tobytes()
                        x += n
        else:
            body += rgbe[y].tobytes()
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode() + bytes(body))


def _ply_ascii(path, verts, faces, props):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by tests\n")
        f.write(f"element vertex {len(verts)}\n")
        for p in props:
            f.write(f"property float {p}\n")
        f.write(f"element face {len(faces)}\nproperty list uchar int vertex_indices\nend_header\n")
        for v in verts:
            f.write(" ".join(repr(float(x)) for x in v) + "\n")
        for fc in faces:
            f.write(f"{len(fc)} " + " ".join(str(i) for i in fc) + "\n")


def _ply_binary(path, verts, faces, props, index_type=("uint", "<I"), extra_uchar=None):
    with open(path, "wb") as f:
        hdr = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {len(verts)}\n"
        for p in props:
            hdr += f"property float {p}\n"
        if extra_uchar:
            hdr += f"property uchar {extra_uchar}\n"
        hdr += f"element face {len(faces)}\nproperty list uchar {index_type[0]} vertex_indices\nend_header\n"
        f.write(hdr.encode())
        for k, v in enumerate(verts):
            f.write(np.asarray(v, "<f4").tobytes())
            if extra_uchar:
                f.write(bytes([k & 255]))
        for fc in faces:
            f.write(bytes([len(fc)]) + np.asarray(fc, index_type[1]).tobytes())


@pytest.fixture(scope="module")
def synth_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("native_scene")
    os.makedirs(d / "shapes")
    os.makedirs(d / "textures")
    rng = np.random.default_rng(11)
    # a grid of quads with normals + uv (quad shape), a triangle soup with s,t + colours, a mixed tri/quad/pentagon mesh
    g = 5
    verts = [(x / g, 0.0, y / g, 0, 1, 0, x / g, y / g) for y in range(g + 1) for x in range(g + 1)]
    quads = [(y * (g + 1) + x, y * (g + 1) + x + 1, (y + 1) * (g + 1) + x + 1, (y + 1) * (g + 1) + x)
             for y in range(g) for x in range(g)]
    _ply_binary(d / "shapes/floor.ply", verts, quads, ["x", "y", "z", "nx", "ny", "nz", "u", "v"], extra_uchar="flag")
    tv = [tuple(rng.random(3) * 2 - 1) + tuple(rng.random(2)) + tuple(rng.random(4)) for _ in range(60)]
    tris = [tuple(rng.choice(60, 3, replace=False)) for _ in range(40)]
    _ply_ascii(d / "shapes/soup.ply", tv, tris, ["s", "t", "x", "y", "z", "red", "green", "blue", "alpha"][0:2] +
               ["x", "y", "z", "red", "green", "blue", "alpha"])
    # reorder so that 's' is the first property (get_tex_coords looks at the first vertex property only)
    tv2 = [(v[3], v[4], v[0], v[1], v[2]) + v[5:] for v in tv]
    _ply_ascii(d / "shapes/soup.ply", tv2, tris, ["s", "t", "x", "y", "z", "red", "green", "blue", "alpha"])
    mv = [tuple(rng.random(3)) for _ in range(12)]
    mixed = [(0, 1, 2), (2, 3, 4, 5), (5, 6, 7, 8, 9), (9, 10, 11), (1, 4, 7, 10, 2, 5)]
    _ply_binary(d / "shapes/mixed.ply", mv, mixed, ["x", "y", "z"], index_type=("int", "<i"))
    lv = [(0, 2, 0), (1, 2, 0), (1, 2, 1), (0, 2, 1), (0.5, 2.5, 0.5)]
    _ply_ascii(d / "shapes/light.ply", lv, [(0, 1, 2), (0, 2, 3), (0, 1, 4), (1, 2, 4, 3, 0)], ["x", "y", "z"])
    # textures: every colour type and every scanline filter
    _png(d / "textures/rgb.png", 7, 5, 2, rng.integers(0, 256, 7 * 5 * 3), filters=[0, 1, 2, 3, 4])
    _png(d / "textures/rgba.png", 4, 6, 6, rng.integers(0, 256, 4 * 6 * 4), filters=[4, 3, 1])
    _png(d / "textures/grey.png", 9, 3, 0, rng.integers(0, 256, 9 * 3), filters=[2, 4])
    _png(d / "textures/greya.png", 3, 3, 4, rng.integers(0, 256, 3 * 3 * 2), filters=[1])
    _png(d / "textures/pal.png", 6, 4, 3, rng.integers(0, 5, 6 * 4), plte=rng.integers(0, 256, 15), filters=[0, 3])
    _png(d / "textures/palt.png", 6, 4, 3, rng.integers(0, 5, 6 * 4), plte=rng.integers(0, 256, 15),
         trns=[0, 128, 255], filters=[4])
    e = rng.integers(120, 136, (6, 16, 1))
    m = rng.integers(0, 256, (6, 16, 3))
    m[:, 4:12, :] = m[:, 4:5, :]  # runs, so the RLE encoder emits both packet kinds
    rgbe = np.concatenate([m, np.broadcast_to(e[:, :1, :], (6, 16, 1))], axis=2)
    rgbe[0, 0] = 0  # a zero-exponent texel
    _hdr(d / "textures/sky_rle.hdr", 16, 6, rgbe, rle=True)
    _hdr(d / "textures/sky_flat.hdr", 5, 4, rng.integers(100, 140, 5 * 4 * 4), rle=False)
    tex = ["rgb.png", "rgba.png", "grey.png", "greya.png", "pal.png", "palt.png", "sky_rle.hdr", "sky_flat.hdr", "absent.png"]
    scene = {
        "asset": {"generator": "tests"},
        "cameras": [{"name": "other", "aspect": 0.75, "lens": 0.035}, {"name": "default", "frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 1, 4],
                                                                       "aperture": 0.01, "focus": 3.5, "orthographic": False}],
        "textures": [{"uri": "textures/" + t} for t in tex],
        "materials": [
            {"name": "floor", "type": "matte", "color": [0.7, 0.6, 0.5], "color_tex": 0, "normal_tex": 1},
            {"type": "glossy", "color": [0.2, 0.8, 0.3], "roughness": 0.15, "roughness_tex": 2},
            {"type": "refractive", "color": [1, 1, 1], "ior": 1.33, "scattering": [0.1, 0.2, 0.3], "trdepth": 0.5,
             "scanisotropy": 0.2, "scattering_tex": 4},
            {"type": "matte", "emission": [10, 9, 8], "emission_tex": 5, "opacity": 0.5},
            {"type": "transparent", "color": [0.9, 0.9, 1.0], "metallic": 0.25},
            {"type": "never-heard-of"},
        ],
        "shapes": [{"uri": "shapes/floor.ply"}, {"uri": "shapes/soup.ply"}, {"uri": "shapes/mixed.ply"},
                   {"uri": "shapes/light.ply"}, {"uri": "shapes/absent.ply"}],
        "instances": [
            {"shape": 0, "material": 0},
            {"shape": 1, "material": 1, "frame": [0.5, 0, 0, 0, 0.5, 0.1, 0, -0.1, 0.5, 0.2, 0.6, 0.3]},
            {"shape": 4, "material": 2},
            {"shape": 2, "material": 2, "frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, -1, 0, 0]},
            {"shape": 3, "material": 3},
            {"shape": 1, "material": 4, "frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, 1.5, 0.5, 0]},
            {"shape": 0, "material": 3, "frame": [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 3, 0]},
        ],
        "environments": [{"emission": [0.5, 0.5, 0.5], "emission_tex": 6}, {"emission": [0, 0, 0]},
                         {"emission": [1, 1, 1], "emission_tex": 7, "frame": [0, 0, 1, 0, 1, 0, -1, 0, 0, 0, 0, 0]}],
    }
    with open(d / "synth.json", "w") as f:
        json.dump(scene, f, indent=1)
    return str(d / "synth.json")


def test_symbols_exported():
    L = _lib.lib()
    for s in ("jt_host_scene_load", "jt_host_scene_build", "jt_host_scene_desc", "jt_host_scene_find_camera",
              "jt_host_scene_num_notes", "jt_host_scene_note", "jt_host_scene_destroy"):
        assert hasattr(L, s)


@pytest.mark.parametrize("hq", [False, True])
def test_synthetic_scene_is_byte_identical(synth_dir, hq):
    py = compare_scene(synth_dir, hq)
    assert len(py.notes) == 3 and len(py.instances) == 6  # one texture + one shape missing, one instance dropped
    assert len(py.shapes[0].quads) == 25 and len(py.shapes[2].quads) > 0 and len(py.shapes[1].triangles) == 40
    assert len(py.shapes[3].triangles) == 6  # the pentagon became a fan: no 4-index face -> triangles


def test_errors_are_reported_not_thrown(tmp_path):
    L = _lib.lib()
    h = C.c_void_p()
    assert L.jt_host_scene_load(str(tmp_path / "nope.json").encode(), C.byref(h)) != 0 and not h
    assert b"cannot open" in L.jt_last_error()
    bad = tmp_path / "bad.json"
    bad.write_text('{"cameras": [ {"lens": } ]}')
    assert L.jt_host_scene_load(str(bad).encode(), C.byref(h)) != 0 and not h
    look = tmp_path / "look.json"
    look.write_text('{"cameras": [{"lookat": [0,0,1,0,0,0,0,1,0]}]}')
    assert L.jt_host_scene_load(str(look).encode(), C.byref(h)) == -4  # JT_ERR_UNSUPPORTED, like the Python mirror
    nonply = tmp_path / "s.json"
    os.makedirs(tmp_path / "shapes")
    (tmp_path / "shapes/x.ply").write_bytes(b"not a ply")
    nonply.write_text('{"shapes": [{"uri": "shapes/x.ply"}]}')
    assert L.jt_host_scene_load(str(nonply).encode(), C.byref(h)) != 0 and not h
    assert L.jt_host_scene_build(None, 0) != 0
    L.jt_host_scene_destroy(None)


@pytest.mark.skipif(not ref_names, reason="reference scenes not available here")
@pytest.mark.parametrize("name", ref_names)
def test_reference_scene_is_byte_identical(name):
    compare_scene(os.path.join(REF_SCENES, name, f"{name}.json"))


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", [1, 2])
def test_native_host_scene_renders_the_same_image(synth_dir, sampler):
    """The library's own loader feeds the GPU the same scene: identical accumulators, both through DeviceScene and
    through `main(--gpu-native-host true)`."""
    tr = importlib.import_module("julia-raytracer_b200.trace")
    cli = importlib.import_module("julia-raytracer_b200.cli")
    py = sio.load_scene(synth_dir)
    p = cli.Params(scene=synth_dir, resolution=96, samples=4, batch=4, sampler=sampler, gpu_seed=3)
    p.camera = scn.find_camera(py, "")
    a = tr.DeviceScene(py, bvhm.make_scene_bvh(py), lm.make_trace_lights(py))
    nat = tr.NativeHostScene(synth_dir)
    assert nat.find_camera("") == p.camera
    b = tr.DeviceScene(nat)
    try:
        sa, sb = tr.make_trace_state(a, p), tr.make_trace_state(b, p)
        tr.trace_samples(sa, a, params=p)
        tr.trace_samples(sb, b, params=p)
        for k in ("image", "albedo", "normal", "hits"):
            assert np.array_equal(getattr(sa, k), getattr(sb, k)), k
        assert sa.image[:, :3].max() > 0
        ca, cb = a.counters(), b.counters()
        for k in ("camera_paths", "scene_rays", "light_rays"):  # launch counts depend on when the host polls
            assert ca[k] == cb[k], k
    finally:
        a.close()
        b.close()
        nat.close()


@pytest.mark.gpu
def test_main_with_native_host(synth_dir, tmp_path):
    jm = importlib.import_module("julia-raytracer_b200.jtrace")
    args = f"--scene {synth_dir} --resolution 64 --samples 2 --batch 2 --gpu-seed 1 "
    r1 = jm.main(args + f"--output {tmp_path / 'a.png'}")
    r2 = jm.main(args + f"--gpu-native-host true --output {tmp_path / 'b.png'}")
    assert np.array_equal(r1["image"], r2["image"])
    assert (tmp_path / "a.png").read_bytes() == (tmp_path / "b.png").read_bytes()


def _build_c_host(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "jtrace_c")
    libdir = os.path.join(root, "julia-raytracer_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "jtrace_c.c"), "-o", exe, "-L" + libdir, "-ljtrace_b200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_c_host_example_builds_and_fails_loudly_without_a_gpu(synth_dir, tmp_path):
    """examples/jtrace_c.c: a host in plain C over include/jtrace_b200.h alone. Without a CUDA device the library refuses
    (JT_ERR_NO_DEVICE): there is no CPU fallback to fall into."""
    import subprocess
    _lib.lib()
    exe = _build_c_host(tmp_path)
    if _lib.lib().jt_device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe, synth_dir, str(tmp_path / "o.ppm"), "32", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("device_lights", [0, 1])
def test_c_host_example_renders_the_same_image(synth_dir, tmp_path, device_lights):
    """The C host and the Python host are two bindings of one ABI: same scene file, same parameters, same bytes."""
    import subprocess
    jm = importlib.import_module("julia-raytracer_b200.jtrace")
    exe = _build_c_host(tmp_path)
    out = tmp_path / "c.ppm"
    r = subprocess.run([exe, synth_dir, str(out), "96", "3", "1", str(device_lights)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    head, size, maxv, body = out.read_bytes().split(b"\n", 3)
    w, h = (int(x) for x in size.split())
    assert head == b"P6" and maxv == b"255" and len(body) == 3 * w * h
    res = jm.main(f"--scene {synth_dir} --resolution 96 --samples 3 --batch 1 --output {tmp_path / 'py.png'}")
    srgb = res["state"].srgb8()
    assert srgb.shape == (h, w, 4)
    assert np.array_equal(np.frombuffer(body, np.uint8).reshape(h, w, 3), srgb[..., :3])
