"""Procedural scenes for the parity tests (TEST INFRASTRUCTURE): they exercise every branch of the
hot path that the five BASELINE scenes do not (transparent rough/delta, volumetric passthrough,
rough refractive, opacity < 1, vertex colours, degenerate quads, roughness / emission /
scattering textures, scaled + rotated instances of shared shapes, several lights)."""
from __future__ import annotations

import importlib

import numpy as np

jt = importlib.import_module("julia-raytracer_b200")
S = importlib.import_module("julia-raytracer_b200.scene")

f32 = np.float32


def _frame(scale=1.0, rot_y=0.0, rot_x=0.0, o=(0, 0, 0), shear=0.0):
    cy, sy, cx, sx = np.cos(rot_y), np.sin(rot_y), np.cos(rot_x), np.sin(rot_x)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    m = ry @ rx @ np.diag(np.broadcast_to(np.asarray(scale, float), (3,)))
    m[0, 1] += shear
    cols = [m[:, 0], m[:, 1], m[:, 2], np.asarray(o, float)]
    return np.concatenate(cols).astype(np.float32)


def _sphere(nu=16, nv=8, quads=False, normals=True, uvs=True, colors=False):
    pos, nrm, uv = [], [], []
    for j in range(nv + 1):
        th = np.pi * j / nv
        for i in range(nu + 1):
            ph = 2 * np.pi * i / nu
            p = [np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)]
            pos.append(p)
            nrm.append(p)
            uv.append([i / nu, j / nv])
    tris, qs = [], []
    for j in range(nv):
        for i in range(nu):
            a = j * (nu + 1) + i
            b, c, d = a + 1, a + nu + 2, a + nu + 1
            if quads:
                if j == 0:
                    qs.append([a, c, d, d])  # degenerate quad (p3 == p4 by index)
                else:
                    qs.append([a, b, c, d])
            else:
                if j != 0:
                    tris.append([a, b, c])
                if j != nv - 1:
                    tris.append([a, c, d])
    sh = S.ShapeData.empty()
    sh.positions = np.asarray(pos, np.float32)
    if normals:
        sh.normals = np.asarray(nrm, np.float32)
    if uvs:
        sh.texcoords = np.asarray(uv, np.float32)
    if colors:
        rng = np.random.default_rng(5)
        sh.colors = (0.5 + 0.5 * rng.random((len(pos), 4))).astype(np.float32)
        sh.colors[:, 3] = 1.0
    if quads:
        sh.quads = np.asarray(qs, np.int64) + 1
    else:
        sh.triangles = np.asarray(tris, np.int64) + 1
    return sh


def _quad(size=1.0, uvscale=1.0, as_tris=False, colors=False):
    sh = S.ShapeData.empty()
    sh.positions = np.asarray([[-size, 0, -size], [size, 0, -size], [size, 0, size], [-size, 0, size]], np.float32)
    sh.normals = np.asarray([[0, 1, 0]] * 4, np.float32)
    sh.texcoords = np.asarray([[0, 0], [uvscale, 0], [uvscale, uvscale], [0, uvscale]], np.float32)
    if colors:
        sh.colors = np.asarray([[1, 0.8, 0.8, 1], [0.8, 1, 0.8, 1], [0.8, 0.8, 1, 1], [1, 1, 1, 1]], np.float32)
    if as_tris:
        sh.triangles = np.asarray([[1, 2, 3], [1, 3, 4]], np.int64)
    else:
        sh.quads = np.asarray([[1, 2, 3, 4]], np.int64)
    return sh


def _tex_rgba8(w, h, seed, kind="noise"):
    rng = np.random.default_rng(seed)
    if kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        c = (((xx // 4) + (yy // 4)) % 2).astype(np.uint8)
        px = np.stack([60 + 180 * c, 200 - 120 * c, 90 + 100 * c, np.full_like(c, 255)], axis=-1)
    elif kind == "normal":
        n = rng.normal(size=(h, w, 3)) * 0.25 + np.array([0, 0, 1.0])
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        px = np.concatenate([(n * 0.5 + 0.5) * 255, np.full((h, w, 1), 255.0)], axis=-1)
    else:
        px = rng.integers(30, 256, (h, w, 4))
        px[..., 3] = 255
    return S.TextureData(w, h, False, None, np.ascontiguousarray(px.reshape(-1, 4).astype(np.uint8)))


def _tex_env(w, h, seed):
    rng = np.random.default_rng(seed)
    yy = np.linspace(0, 1, h)[:, None]
    base = 0.2 + 0.8 * (1 - yy) * np.ones((h, w))
    px = np.stack([base * 0.9, base, np.minimum(1.0, base * 1.1), np.ones_like(base)], axis=-1)
    px[..., :3] *= 0.6 + 0.4 * rng.random((h, w, 1))
    px = np.clip(px, 0, 1)
    return S.TextureData(w, h, True, np.ascontiguousarray(px.reshape(-1, 4).astype(np.float32)), None)


def _materials(specs):
    m = np.zeros(len(specs), S.MATERIAL_DTYPE)
    for i, sp in enumerate(specs):
        m[i]["type"] = sp.get("type", S.MATTE)
        m[i]["emission"] = sp.get("emission", (0, 0, 0))
        m[i]["color"] = sp.get("color", (0.8, 0.8, 0.8))
        m[i]["roughness"] = sp.get("roughness", 0)
        m[i]["metallic"] = 0
        m[i]["ior"] = sp.get("ior", 1.5)
        m[i]["scattering"] = sp.get("scattering", (0, 0, 0))
        m[i]["scanisotropy"] = sp.get("scanisotropy", 0)
        m[i]["trdepth"] = sp.get("trdepth", 0.01)
        m[i]["opacity"] = sp.get("opacity", 1)
        for k in ("emission_tex", "color_tex", "roughness_tex", "scattering_tex", "normal_tex"):
            m[i][k] = sp.get(k, -1)
    return m


def make_scene(name: str) -> "S.SceneData":
    """synthetic_all: every material type / texture role / primitive kind; with environment.
    synthetic_closed: area lights only, no environment (envhidden / alpha paths).
    synthetic_one: a single triangle + one light triangle (smallest possible scene)."""
    cam = S.CameraData(frame=_frame(o=(0.0, 1.2, 5.0), rot_x=-0.15), aspect=f32(1.6), lens=f32(0.05),
                       film=f32(0.036), focus=f32(5.0), aperture=f32(0.0), name="default")
    if name == "synthetic_one":
        tri = S.ShapeData.empty()
        tri.positions = np.asarray([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32)
        tri.triangles = np.asarray([[1, 2, 3]], np.int64)
        light = S.ShapeData.empty()
        light.positions = np.asarray([[-1, 3, 2], [1, 3, 2], [0, 3, 0]], np.float32)
        light.triangles = np.asarray([[1, 3, 2]], np.int64)
        mats = _materials([dict(color=(0.7, 0.6, 0.5)), dict(emission=(10, 10, 10), color=(0, 0, 0))])
        inst = np.zeros(2, S.INSTANCE_DTYPE)
        inst[0]["frame"], inst[0]["shape"], inst[0]["material"] = S.IDENTITY_FRAME, 1, 1
        inst[1]["frame"], inst[1]["shape"], inst[1]["material"] = S.IDENTITY_FRAME, 2, 2
        cam.frame = _frame(o=(0, 0, 4))
        return S.SceneData([cam], inst, np.zeros(0, S.ENVIRONMENT_DTYPE), [tri, light], [], mats)

    textures = [_tex_env(64, 32, 1), _tex_rgba8(32, 32, 2, "checker"), _tex_rgba8(16, 16, 3, "normal"),
                _tex_rgba8(8, 8, 4, "noise")]
    T_ENV, T_CHECK, T_NORMAL, T_NOISE = 1, 2, 3, 4
    shapes = [_quad(6.0, 4.0, colors=True),            # 1 floor (quad, vertex colours, texcoords)
              _sphere(16, 8),                           # 2 tri sphere with normals + uvs
              _sphere(12, 6, quads=True),               # 3 quad sphere incl. degenerate quads
              _sphere(10, 5, normals=False, uvs=False), # 4 faceted sphere, no attributes
              _quad(0.7, as_tris=True),                 # 5 light (2 triangles)
              _quad(0.5),                               # 6 light (1 quad)
              _sphere(8, 4, colors=True)]               # 7 sphere with vertex colours
    specs = [
        dict(color=(0.8, 0.8, 0.8), color_tex=T_CHECK),                                   # 1 floor matte textured
        dict(type=S.GLOSSY, color=(0.7, 0.3, 0.3), roughness=0.3, normal_tex=T_NORMAL),   # 2 glossy normal-mapped
        dict(type=S.REFLECTIVE, color=(0.9, 0.8, 0.5), roughness=0.0),                    # 3 mirror
        dict(type=S.REFLECTIVE, color=(0.6, 0.7, 0.9), roughness=0.25, roughness_tex=T_NOISE),  # 4 rough metal
        dict(type=S.TRANSPARENT, color=(0.9, 0.9, 1.0), roughness=0.0),                   # 5 thin glass
        dict(type=S.TRANSPARENT, color=(0.9, 1.0, 0.9), roughness=0.2),                   # 6 rough thin glass
        dict(type=S.REFRACTIVE, color=(0.95, 0.7, 0.7), roughness=0.0, scattering=(0.3, 0.3, 0.3), trdepth=0.5),  # 7
        dict(type=S.REFRACTIVE, color=(0.8, 0.9, 0.95), roughness=0.15, trdepth=1.0),     # 8 rough glass
        dict(type=S.VOLUMETRIC, color=(0.6, 0.6, 0.9), scattering=(0.8, 0.8, 0.8), scanisotropy=0.3, trdepth=0.7),  # 9
        dict(color=(0.3, 0.8, 0.4), opacity=0.5),                                         # 10 half-transparent matte
        dict(emission=(12, 11, 9), color=(0, 0, 0)),                                      # 11 light
        dict(emission=(3, 5, 9), color=(0, 0, 0), emission_tex=T_CHECK),                  # 12 textured light
        dict(type=S.SUBSURFACE, color=(0.8, 0.6, 0.5), roughness=0.3, scattering=(0.5, 0.3, 0.2), trdepth=0.3),  # 13
        dict(type=S.GLOSSY, color=(0.5, 0.5, 0.9), roughness=0.0),                        # 14 glossy r=0 -> min_roughness
        dict(color=(0.9, 0.9, 0.9), scattering_tex=T_NOISE),                              # 15 matte with vertex colours
    ]
    mats = _materials(specs)
    placements = [  # (shape, material, frame)
        (1, 1, _frame(o=(0, 0, 0))),
        (2, 2, _frame(0.6, 0.3, 0.0, (-2.4, 0.6, 0.0))),
        (2, 3, _frame(0.5, 0.0, 0.0, (-1.2, 0.5, 0.6))),
        (3, 4, _frame((0.5, 0.7, 0.5), 0.7, 0.2, (0.0, 0.7, 0.0))),
        (2, 5, _frame(0.45, 0.0, 0.0, (1.2, 0.45, 0.8))),
        (3, 6, _frame(0.45, 1.1, 0.0, (2.3, 0.45, 0.2))),
        (2, 7, _frame(0.5, 0.0, 0.4, (-1.8, 0.5, 1.8))),
        (4, 8, _frame(0.45, 0.0, 0.0, (-0.5, 0.45, 1.9))),
        (2, 9, _frame(0.5, 0.2, 0.0, (0.8, 0.5, 2.0))),
        (4, 10, _frame(0.4, 0.0, 0.0, (2.0, 0.4, 1.7), shear=0.2)),
        (5, 11, _frame(1.0, 0.0, np.pi, (0.0, 3.2, 0.5))),
        (6, 12, _frame(1.0, 0.0, np.pi * 0.9, (-2.5, 2.5, 1.0))),
        (2, 13, _frame(0.35, 0.0, 0.0, (0.2, 0.35, 3.0))),
        (4, 14, _frame(0.3, 0.0, 0.0, (-1.0, 0.3, 3.0))),
        (7, 15, S.IDENTITY_FRAME.copy() + np.array([0] * 9 + [1.4, 1.0, 3.0], np.float32)),
        (7, 15, S.IDENTITY_FRAME.copy()),  # exact identity frame: inlined by the wide BVH builder
    ]
    inst = np.zeros(len(placements), S.INSTANCE_DTYPE)
    for i, (s, m, fr) in enumerate(placements):
        inst[i]["frame"], inst[i]["shape"], inst[i]["material"] = fr, s, m
    envs = np.zeros(1, S.ENVIRONMENT_DTYPE)
    envs[0]["frame"] = _frame(rot_y=0.8)
    envs[0]["emission"] = (0.7, 0.7, 0.8)
    envs[0]["emission_tex"] = T_ENV
    if name == "synthetic_closed":
        envs = np.zeros(0, S.ENVIRONMENT_DTYPE)
    elif name != "synthetic_all":
        raise KeyError(name)
    return S.SceneData([cam], inst, envs, shapes, textures, mats)
