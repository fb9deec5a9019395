"""Scene I/O on the host: the reference's `load_scene` / `load_shape` / `load_texture` /
`save_image` (src/sceneio.jl:25-123, src/shape.jl:78-446, src/scene.jl:164-189), plus a packed
single-file container (`.jtscene`, an uncompressed-or-deflated npz) so that scenes travel to
machines where the reference checkout does not exist.

This is host code (out of the GPU hot path); it produces exactly the arrays
the Julia host would own, in the Julia memory layouts of `scene.py`.

Missing-asset rule: a texture file that is absent becomes a 1x1 opaque white
RGBA8 texture; instances whose shape file is absent are dropped before BVH/light building.
Every substitution is recorded in `SceneData.notes`.
"""
from __future__ import annotations

import io
import json
import os
from typing import List

import numpy as np

from .scene import (CameraData, ENVIRONMENT_DTYPE, INSTANCE_DTYPE, MATERIAL_DTYPE,
                    MATERIAL_TYPES, MATTE, SceneData, ShapeData, TextureData, frame_from_json,
                    invalid_id)

_f32 = np.float32


# ----------------------------------------------------------------------------------------------
# PLY (binary little endian / ascii), src/shape.jl:78-124
# ----------------------------------------------------------------------------------------------
_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "<i2", "int16": "<i2",
    "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4", "uint": "<u4",
    "uint32": "<u4", "float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8",
}


def _parse_ply(path: str):
    """Returns {element: {property: ndarray | (flat_values, start_offsets)}}."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header")
    header = data[:end].decode("ascii", "replace").split("\n")
    body = end + len(b"end_header")
    # header terminator may be \n or \r\n
    if data[body:body + 2] == b"\r\n":
        body += 2
    else:
        body += 1
    fmt = None
    elements = []  # (name, count, [(kind, name, types...)])
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append(("list", tok[4], _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]]))
            else:
                elements[-1][2].append(("scalar", tok[2], _PLY_TYPES[tok[1]]))
    out = {}
    if fmt == "binary_little_endian":
        pos = body
        for name, count, props in elements:
            el = {}
            if all(p[0] == "scalar" for p in props):
                dt = np.dtype([(p[1], p[2]) for p in props])
                arr = np.frombuffer(data, dtype=dt, count=count, offset=pos)
                pos += dt.itemsize * count
                for p in props:
                    el[p[1]] = arr[p[1]]
            elif len(props) == 1 and props[0][0] == "list":
                _, pname, ctype, itype = props[0]
                csz, isz = np.dtype(ctype).itemsize, np.dtype(itype).itemsize
                # fast path: constant list length
                first = int(np.frombuffer(data, dtype=ctype, count=1, offset=pos)[0]) if count else 0
                rec = csz + first * isz
                ok = False
                if count and pos + rec * count <= len(data):
                    raw = np.frombuffer(data, dtype=np.uint8, count=rec * count, offset=pos)
                    raw = raw.reshape(count, rec)
                    counts = raw[:, :csz].copy().view(ctype).reshape(count)
                    if np.all(counts == first):
                        vals = raw[:, csz:].copy().view(itype).reshape(count * first)
                        starts = np.arange(count + 1, dtype=np.int64) * first
                        el[pname] = (vals.astype(np.int64), starts)
                        pos += rec * count
                        ok = True
                if not ok:
                    vals: List[int] = []
                    starts = [0]
                    for _ in range(count):
                        n = int(np.frombuffer(data, dtype=ctype, count=1, offset=pos)[0])
                        pos += csz
                        v = np.frombuffer(data, dtype=itype, count=n, offset=pos)
                        pos += n * isz
                        vals.extend(int(x) for x in v)
                        starts.append(len(vals))
                    el[pname] = (np.asarray(vals, dtype=np.int64), np.asarray(starts, dtype=np.int64))
            else:
                raise ValueError(f"{path}: mixed scalar/list element '{name}' not supported")
            out[name] = el
    elif fmt == "ascii":
        toks = data[body:].split()
        k = 0
        for name, count, props in elements:
            el = {p[1]: [] for p in props if p[0] == "scalar"}
            lists = {p[1]: ([], [0]) for p in props if p[0] == "list"}
            for _ in range(count):
                for p in props:
                    if p[0] == "scalar":
                        el[p[1]].append(float(toks[k])); k += 1
                    else:
                        n = int(toks[k]); k += 1
                        lists[p[1]][0].extend(int(x) for x in toks[k:k + n]); k += n
                        lists[p[1]][1].append(len(lists[p[1]][0]))
            for p in props:
                if p[0] == "scalar":
                    el[p[1]] = np.asarray(el[p[1]], dtype=p[2])
                else:
                    el[p[1]] = (np.asarray(lists[p[1]][0], dtype=np.int64),
                                np.asarray(lists[p[1]][1], dtype=np.int64))
            out[name] = el
    else:
        raise ValueError(f"{path}: unsupported PLY format {fmt}")
    return out, [(n, [p[1] for p in props]) for n, _, props in elements]


def _faces_to_elements(vals: np.ndarray, starts: np.ndarray):
    """`get_faces` (src/shape.jl:430-446): if any face has exactly 4 indices the whole shape is
    stored as quads (3-gon -> (a,b,c,c), n-gon -> fan of degenerate quads, `:323-369`), else as
    fan-triangulated triangles (`:371-405`). 0-based in, 0-based out."""
    sizes = np.diff(starts)
    has_quads = bool(np.any(sizes == 4))  # src/shape.jl:302-321
    width = 4 if has_quads else 3
    if len(sizes) and np.all(sizes == sizes[0]) and sizes[0] in (3, 4):
        n = int(sizes[0])
        v = vals.reshape(-1, n)
        if n == width:
            return has_quads, v.copy()
        # all triangles inside a quad shape cannot happen (has_quads would be False)
    rows = []
    for i in range(len(sizes)):
        s, n = int(starts[i]), int(sizes[i])
        d = vals[s:s + n]
        if n < 3:
            row = [-1] * width
            for k in range(n):
                row[k] = int(d[k])
            rows.append(row)
        elif n == 3:
            rows.append([d[0], d[1], d[2], d[2]] if has_quads else [d[0], d[1], d[2]])
        elif n == 4 and has_quads:
            rows.append([d[0], d[1], d[2], d[3]])
        else:
            for item in range(2, n):
                if has_quads:
                    rows.append([d[0], d[item - 1], d[item], d[item]])
                else:
                    rows.append([d[0], d[item - 1], d[item]])
    return has_quads, np.asarray(rows, dtype=np.int64).reshape(-1, width)


def load_shape(path: str) -> ShapeData:
    """src/shape.jl:78-124."""
    ply, order = _parse_ply(path)
    shape = ShapeData.empty()
    vert = ply.get("vertex", {})
    vnames = dict(order).get("vertex", [])

    def stack(names):
        if all(n in vert and not isinstance(vert[n], tuple) for n in names):
            return np.stack([np.asarray(vert[n], dtype=np.float32) for n in names], axis=1)
        return None

    p = stack(["x", "y", "z"])
    if p is not None:
        shape.positions = np.ascontiguousarray(p)
    n = stack(["nx", "ny", "nz"])
    if n is not None:
        shape.normals = np.ascontiguousarray(n)
    # get_tex_coords, src/shape.jl:265-278: only the FIRST vertex property decides s,t vs u,v
    if vnames:
        uvn = ["s", "t"] if vnames[0] == "s" else ["u", "v"]
        t = stack(uvn)
        if t is not None:
            t = t.copy()
            t[:, 1] = _f32(1) - t[:, 1]  # flip, src/shape.jl:233-235
            shape.texcoords = np.ascontiguousarray(t)
    c = stack(["red", "green", "blue", "alpha"])  # src/shape.jl:280-299 (rgb-only path is broken)
    if c is not None and "alpha" in vnames:
        shape.colors = np.ascontiguousarray(c)
    face = ply.get("face", {})
    if "vertex_indices" in face and isinstance(face["vertex_indices"], tuple):
        vals, starts = face["vertex_indices"]
        is_quads, elems = _faces_to_elements(vals, starts)
        elems = elems + 1  # src/shape.jl:101-105
        if is_quads:
            shape.quads = np.ascontiguousarray(elems)
        else:
            shape.triangles = np.ascontiguousarray(elems)
    for bad in ("line", "point"):
        if bad in ply and len(next(iter(ply[bad].values()), ())):
            raise NotImplementedError(
                f"{path}: '{bad}' elements crash the reference (SURVEY.md §2.3); not supported")
    return shape


# ----------------------------------------------------------------------------------------------
# textures, src/scene.jl:164-189
# ----------------------------------------------------------------------------------------------
def _decode_hdr(path: str) -> np.ndarray:
    """Radiance RGBE -> (H, W, 3) float32, value = mantissa * 2^(e-136) (exact)."""
    import cv2  # OpenCV's RGBE reader returns mantissa * 2^(e-136), BGR order
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise IOError(f"cannot decode {path}")
    return np.ascontiguousarray(img[:, :, ::-1].astype(np.float32))


def hdr_as_reference_loader(rgb_linear: np.ndarray) -> np.ndarray:
    """Linear RGBE radiance -> what the reference's HDR loader yields (see load_texture)."""
    c = np.maximum(np.asarray(rgb_linear, np.float64), 0.0)
    enc = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1.0 / 2.4) - 0.055)
    q = np.rint(np.clip(enc, 0.0, 1.0) * 65535.0) / 65535.0  # ImageMagick Q16 quantum
    return q.astype(np.float32)


def load_texture(path: str) -> TextureData:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".hdr":
        rgb = _decode_hdr(path)
        h, w = rgb.shape[:2]
        # Q9: the reference loads HDR through FileIO -> ImageMagick, whose values the
        # authors call "wrong" (src/scene.jl:167). Third-party behaviour, pinned empirically against
        # the reference's shipped images/ecosys_path.png: directly visible sky pixels match
        # clamp01(sRGB-ENCODED(rgbe)) with RMSE 0.0018 (linear-clamp hypothesis: 0.28) -- i.e. the
        # loader hands the renderer display-encoded, [0,1]-clamped, 16-bit quantum values which the
        # renderer then treats as linear radiance. Reproduced here; alpha = 1 (src/math.jl:28).
        rgb = hdr_as_reference_loader(rgb)
        px = np.concatenate([rgb.reshape(-1, 3), np.ones((h * w, 1), np.float32)], axis=1)
        return TextureData(w, h, True, np.ascontiguousarray(px, dtype=np.float32), None)
    if ext == ".png":
        from PIL import Image
        im = Image.open(path)
        if im.mode in ("RGBA", "LA", "PA") or "transparency" in im.info:
            px = np.asarray(im.convert("RGBA"), dtype=np.uint8)
        else:
            # Vec4b(::RGB) stores alpha = 1 (not 255), src/math.jl:39-44
            rgb = np.asarray(im.convert("RGB"), dtype=np.uint8)
            px = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 1, np.uint8)], axis=2)
        h, w = px.shape[:2]
        return TextureData(w, h, False, None, np.ascontiguousarray(px.reshape(-1, 4)))
    raise ValueError(f"unknown texture format: {ext}")


def white_texture() -> TextureData:
    return TextureData(1, 1, False, None, np.full((1, 4), 255, np.uint8))


# ----------------------------------------------------------------------------------------------
# JSON scene, src/sceneio.jl:25-93 and the constructors in src/scene.jl:58-263
# ----------------------------------------------------------------------------------------------
def _check_no_lookat(obj, what):
    if "lookat" in obj:
        raise NotImplementedError(f"{what} 'lookat' is not used by any shipped scene; unsupported")


def _camera(j) -> CameraData:
    _check_no_lookat(j, "camera")
    return CameraData(
        frame=frame_from_json(j.get("frame")),
        orthographic=bool(j.get("orthographic", False)),
        lens=_f32(j.get("lens", 0.050)), film=_f32(j.get("film", 0.036)),
        aspect=_f32(j.get("aspect", 1.5)), focus=_f32(j.get("focus", 10000)),
        aperture=_f32(j.get("aperture", 0)), name=str(j.get("name", "")))


def _material(j, out):
    out["type"] = MATERIAL_TYPES.get(j.get("type", "matte"), MATTE)
    out["emission"] = np.asarray(j.get("emission", [0, 0, 0]), np.float64).astype(np.float32)
    out["color"] = np.asarray(j.get("color", [0, 0, 0]), np.float64).astype(np.float32)
    out["roughness"] = _f32(j.get("roughness", 0))
    out["metallic"] = _f32(j.get("metallic", 0))
    out["ior"] = _f32(j.get("ior", 1.5))
    out["scattering"] = np.asarray(j.get("scattering", [0, 0, 0]), np.float64).astype(np.float32)
    out["scanisotropy"] = _f32(j.get("scanisotropy", 0))
    out["trdepth"] = _f32(j.get("trdepth", 0.01))
    out["opacity"] = _f32(j.get("opacity", 1))
    for k in ("emission_tex", "color_tex", "roughness_tex", "scattering_tex", "normal_tex"):
        out[k] = int(j.get(k, invalid_id - 1)) + 1  # 0-based -> 1-based, missing -> -1


def load_scene(filename: str, no_parallel: bool = False, verbose: bool = False) -> SceneData:
    """`load_scene`, src/sceneio.jl:25-93. Accepts a scene JSON or a packed `.jtscene`."""
    if filename.endswith(".jtscene"):
        return load_packed(filename)
    d = os.path.dirname(filename)
    with open(filename, "r") as f:
        js = json.load(f)
    notes: List[str] = []
    say = print if verbose else (lambda *a, **k: None)
    say("    loading cameras...")
    cameras = [_camera(c) for c in js.get("cameras", [])]
    say("    loading textures...")
    textures = []
    for t in js.get("textures", []):
        p = os.path.join(d, t["uri"])
        if os.path.exists(p):
            textures.append(load_texture(p))
        else:
            notes.append(f"missing texture {t['uri']} -> 1x1 opaque white")
            textures.append(white_texture())
    say("    loading materials...")
    jm = js.get("materials", [])
    materials = np.zeros(len(jm), MATERIAL_DTYPE)
    for i, m in enumerate(jm):
        _material(m, materials[i])
    say("    loading shapes...")
    shapes = []
    missing_shapes = set()
    for i, s in enumerate(js.get("shapes", [])):
        p = os.path.join(d, s["uri"])
        if os.path.exists(p):
            shapes.append(load_shape(p))
        else:
            notes.append(f"missing shape {s['uri']} -> its instances are dropped")
            missing_shapes.add(i + 1)
            shapes.append(ShapeData.empty())
    say("    loading instances...")
    ji = js.get("instances", [])
    inst = np.zeros(len(ji), INSTANCE_DTYPE)
    keep = np.ones(len(ji), bool)
    for i, x in enumerate(ji):
        _check_no_lookat(x, "instance")
        inst[i]["frame"] = frame_from_json(x.get("frame"))
        inst[i]["shape"] = int(x.get("shape", invalid_id - 1)) + 1
        inst[i]["material"] = int(x.get("material", invalid_id - 1)) + 1
        if int(inst[i]["shape"]) in missing_shapes:
            keep[i] = False
    if not keep.all():
        notes.append(f"dropped {int((~keep).sum())} of {len(ji)} instances (absent shape files)")
        inst = np.ascontiguousarray(inst[keep])
    say("    loading environments...")
    je = js.get("environments", [])
    envs = np.zeros(len(je), ENVIRONMENT_DTYPE)
    for i, e in enumerate(je):
        _check_no_lookat(e, "environment")
        envs[i]["frame"] = frame_from_json(e.get("frame"))
        envs[i]["emission"] = np.asarray(e.get("emission", [0, 0, 0]), np.float64).astype(np.float32)
        envs[i]["emission_tex"] = int(e.get("emission_tex", invalid_id - 1)) + 1
    return SceneData(cameras, inst, envs, shapes, textures, materials, notes)


# ----------------------------------------------------------------------------------------------
# packed container
# ----------------------------------------------------------------------------------------------
def save_packed(scene: SceneData, filename: str, compress: bool = True) -> None:
    arrs = {}
    cams = []
    for c in scene.cameras:
        cams.append(dict(frame=[float(x) for x in c.frame], orthographic=bool(c.orthographic),
                         lens=float(c.lens), film=float(c.film), aspect=float(c.aspect),
                         focus=float(c.focus), aperture=float(c.aperture), name=c.name))
    meta = dict(version=1, cameras=cams, notes=scene.notes, num_shapes=len(scene.shapes),
                textures=[dict(width=t.width, height=t.height, linear=bool(t.linear),
                               kind="f" if t.pixelsf is not None else "b") for t in scene.textures])
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    arrs["instances"] = scene.instances.view(np.uint8)
    arrs["materials"] = scene.materials.view(np.uint8)
    arrs["environments"] = scene.environments.view(np.uint8)
    for i, t in enumerate(scene.textures):
        arrs[f"tex{i}"] = t.pixelsf if t.pixelsf is not None else t.pixelsb
    for i, s in enumerate(scene.shapes):
        for k in ("positions", "normals", "texcoords", "colors"):
            a = getattr(s, k)
            if len(a):
                arrs[f"s{i}_{k}"] = a
        for k in ("triangles", "quads"):
            a = getattr(s, k)
            if len(a):
                arrs[f"s{i}_{k}"] = a.astype(np.int32)  # widened to Int64 again on load
    (np.savez_compressed if compress else np.savez)(filename + ".tmp.npz", **arrs)
    os.replace(filename + ".tmp.npz", filename)


def load_packed(filename: str) -> SceneData:
    with open(filename, "rb") as f:
        z = np.load(io.BytesIO(f.read()))
    meta = json.loads(bytes(z["meta"]).decode())
    cams = [CameraData(frame=np.asarray(c["frame"], np.float32), orthographic=c["orthographic"],
                       lens=_f32(c["lens"]), film=_f32(c["film"]), aspect=_f32(c["aspect"]),
                       focus=_f32(c["focus"]), aperture=_f32(c["aperture"]), name=c["name"])
            for c in meta["cameras"]]
    textures = []
    for i, t in enumerate(meta["textures"]):
        a = np.ascontiguousarray(z[f"tex{i}"])
        textures.append(TextureData(t["width"], t["height"], t["linear"],
                                    a if t["kind"] == "f" else None,
                                    a if t["kind"] == "b" else None))
    shapes = []
    for i in range(meta["num_shapes"]):
        s = ShapeData.empty()
        for k in ("positions", "normals", "texcoords", "colors"):
            if f"s{i}_{k}" in z:
                setattr(s, k, np.ascontiguousarray(z[f"s{i}_{k}"], dtype=np.float32))
        for k in ("triangles", "quads"):
            if f"s{i}_{k}" in z:
                setattr(s, k, np.ascontiguousarray(z[f"s{i}_{k}"].astype(np.int64)))
        shapes.append(s)
    inst = np.ascontiguousarray(z["instances"]).view(INSTANCE_DTYPE).copy()
    mats = np.ascontiguousarray(z["materials"]).view(MATERIAL_DTYPE).copy()
    envs = np.ascontiguousarray(z["environments"]).view(ENVIRONMENT_DTYPE).copy()
    return SceneData(cams, inst, envs, shapes, textures, mats, list(meta.get("notes", [])))


# ----------------------------------------------------------------------------------------------
# save_image, src/sceneio.jl:97-123 + src/color.jl:25-29
# ----------------------------------------------------------------------------------------------
def rgb_to_srgb(c: np.ndarray) -> np.ndarray:
    c = np.asarray(c, np.float32)
    with np.errstate(invalid="ignore"):
        hi = np.float32(1.055) * np.power(c.astype(np.float64), float(np.float32(1) / np.float32(2.4))
                                           ).astype(np.float32) - np.float32(0.055)
    return np.where(c <= np.float32(0.0031308), np.float32(12.92) * c, hi).astype(np.float32)


def image_to_srgb8(image_rgba: np.ndarray) -> np.ndarray:
    """Linear RGBA float (H,W,4) -> 8-bit sRGB RGBA as the reference's PNG writer produces
    (`clamp01nan` then N0f8 rounding; +-1 LSB is third-party behaviour)."""
    img = np.asarray(image_rgba, np.float32).copy()
    img[..., :3] = rgb_to_srgb(img[..., :3])
    img = np.nan_to_num(img, nan=0.0, posinf=1.0, neginf=0.0)
    img = np.clip(img, 0.0, 1.0)
    return np.rint(img * 255.0).astype(np.uint8)


def save_image(filename: str, image_rgba: np.ndarray) -> None:
    ext = os.path.splitext(filename)[1].lower()
    if ext != ".png":
        raise ValueError(f"{ext} is not supported")
    from PIL import Image
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    Image.fromarray(image_to_srgb8(image_rgba), "RGBA").save(filename)


def save_srgb8(filename: str, rgba8: np.ndarray) -> None:
    """Write an already display-encoded (H, W, 4) uint8 image (what jt_state_download_srgb8 returns)."""
    ext = os.path.splitext(filename)[1].lower()
    if ext != ".png":
        raise ValueError(f"{ext} is not supported")
    from PIL import Image
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    Image.fromarray(np.ascontiguousarray(rgba8, np.uint8), "RGBA").save(filename)
