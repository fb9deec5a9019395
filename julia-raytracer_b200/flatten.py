"""The "new flattening pass" of the north star: packs the host-side SceneData + SceneBvh +
TraceLights into the flat `jt_scene_desc` of include/jtrace_b200.h (pointers + counts over
arrays that already have the Julia isbits layouts, so nothing is converted or copied here).

The Julia twin of this file is julia/JtraceB200.jl (`flatten_scene`)."""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional

import numpy as np

from . import _abi as A
from .scene import SceneData


@dataclasses.dataclass
class BvhTree:  # src/bvh.jl:46-49
    nodes: np.ndarray       # BVHNODE_DTYPE
    primitives: np.ndarray  # int64, 1-based


@dataclasses.dataclass
class SceneBvh:  # src/bvh.jl:57-62
    bvh: BvhTree
    shapes: List[BvhTree]


@dataclasses.dataclass
class TraceLight:  # src/trace.jl:102-105
    instance: int
    environment: int
    elements_cdf: np.ndarray  # float32 inclusive prefix sums


def _ptr(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _bvh_desc(t: Optional[BvhTree]) -> A.jt_bvh_desc:
    d = A.jt_bvh_desc()
    if t is not None:
        d.nodes = _ptr(t.nodes)
        d.num_nodes = len(t.nodes)
        d.primitives = _ptr(t.primitives)
        d.num_primitives = len(t.primitives)
    return d


def srgb_to_rgb_lut() -> np.ndarray:
    """srgb_to_rgb(b / 255f0) for b = 0..255 (src/color.jl:12-23); Julia evaluates
    `x^2.4f0` as Float32(exp2(log2(Float64(x)) * Float64(2.4f0)))."""
    c = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    lo = (c / np.float32(12.92)).astype(np.float32)
    b = ((c + np.float32(0.055)) / np.float32(1.055)).astype(np.float32)
    with np.errstate(divide="ignore"):
        hi = np.exp2(np.log2(b.astype(np.float64)) * float(np.float32(2.4))).astype(np.float32)
    return np.where(c <= np.float32(0.04045), lo, hi).astype(np.float32)


class FlatScene:
    """Owns a jt_scene_desc and keeps every array it points into alive."""

    def __init__(self, scene: SceneData, bvh: Optional[SceneBvh], lights: Optional[List[TraceLight]]):
        self._keep = [scene, bvh, lights]
        d = A.jt_scene_desc()

        cams = (A.jt_camera * max(1, len(scene.cameras)))()
        for i, c in enumerate(scene.cameras):
            fr = np.ascontiguousarray(c.frame, np.float32)
            C.memmove(C.byref(cams[i].frame), fr.ctypes.data, 48)
            cams[i].orthographic = int(bool(c.orthographic))
            cams[i].lens, cams[i].film, cams[i].aspect = float(c.lens), float(c.film), float(c.aspect)
            cams[i].focus, cams[i].aperture = float(c.focus), float(c.aperture)
        self._keep.append(cams)
        d.num_cameras, d.cameras = len(scene.cameras), C.cast(cams, C.c_void_p)

        d.num_instances, d.instances = len(scene.instances), _ptr(scene.instances)
        d.num_environments, d.environments = len(scene.environments), _ptr(scene.environments)
        d.num_materials, d.materials = len(scene.materials), _ptr(scene.materials)

        texs = (A.jt_texture_desc * max(1, len(scene.textures)))()
        for i, t in enumerate(scene.textures):
            texs[i].width, texs[i].height, texs[i].linear = t.width, t.height, int(bool(t.linear))
            texs[i].pixelsf = _ptr(t.pixelsf)
            texs[i].pixelsb = _ptr(t.pixelsb)
        self._keep.append(texs)
        d.num_textures, d.textures = len(scene.textures), C.cast(texs, C.c_void_p)

        shp = (A.jt_shape_desc * max(1, len(scene.shapes)))()
        for i, s in enumerate(scene.shapes):
            shp[i].positions, shp[i].num_positions = _ptr(s.positions), len(s.positions)
            shp[i].normals, shp[i].num_normals = _ptr(s.normals), len(s.normals)
            shp[i].texcoords, shp[i].num_texcoords = _ptr(s.texcoords), len(s.texcoords)
            shp[i].colors, shp[i].num_colors = _ptr(s.colors), len(s.colors)
            shp[i].triangles, shp[i].num_triangles = _ptr(s.triangles), len(s.triangles)
            shp[i].quads, shp[i].num_quads = _ptr(s.quads), len(s.quads)
            shp[i].bvh = _bvh_desc(bvh.shapes[i] if bvh is not None else None)
        self._keep.append(shp)
        d.num_shapes, d.shapes = len(scene.shapes), C.cast(shp, C.c_void_p)

        nl = len(lights) if lights is not None else 0
        lts = (A.jt_light_desc * max(1, nl))()
        for i in range(nl):
            lts[i].instance, lts[i].environment = lights[i].instance, lights[i].environment
            lts[i].elements_cdf = _ptr(lights[i].elements_cdf)
            lts[i].num_elements = len(lights[i].elements_cdf)
        self._keep.append(lts)
        d.num_lights, d.lights = nl, C.cast(lts, C.c_void_p)

        d.bvh = _bvh_desc(bvh.bvh if bvh is not None else None)
        self._lut = srgb_to_rgb_lut()
        d.srgb_to_rgb_lut = self._lut.ctypes.data
        self.desc = d

    def byref(self):
        return C.byref(self.desc)
