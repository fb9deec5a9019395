"""`make_scene_bvh` for non-Julia hosts (src/bvh.jl:66-136): primitive boxes are computed here
with numpy (Float32 min/max are exact), the tree itself by `jt_make_bvh`, the library's C++
restatement of `make_bvh` / `split_middle` / `split_sah` / `partition` (src/bvh.jl:138-304)."""
from __future__ import annotations

import numpy as np

from . import _lib
from .flatten import BvhTree, SceneBvh
from .scene import BVHNODE_DTYPE, SceneData, ShapeData


def make_bvh(bboxes: np.ndarray, high_quality: bool = False) -> BvhTree:
    """bboxes: (n, 6) float32 rows {min.xyz, max.xyz}."""
    L = _lib.lib()
    bboxes = np.ascontiguousarray(bboxes, np.float32).reshape(-1, 6)
    n = len(bboxes)
    nodes = np.zeros(2 * n + 1, BVHNODE_DTYPE)
    prims = np.zeros(max(n, 1), np.int64)
    count = _lib.C.c_int64(0)
    _lib.check(L.jt_make_bvh(bboxes.ctypes.data if n else None, n, int(high_quality), nodes.ctypes.data,
                             _lib.C.byref(count), prims.ctypes.data))
    return BvhTree(np.ascontiguousarray(nodes[:count.value]), np.ascontiguousarray(prims[:n]))


def shape_bboxes(shape: ShapeData) -> np.ndarray:
    """triangle_bounds / quad_bounds (src/geometry.jl:64-68); triangles take precedence."""
    if len(shape.triangles):
        idx = shape.triangles - 1
    elif len(shape.quads):
        idx = shape.quads - 1
    else:
        return np.zeros((0, 6), np.float32)
    p = shape.positions[idx]  # (n, k, 3)
    return np.concatenate([p.min(axis=1), p.max(axis=1)], axis=1).astype(np.float32)


def transform_bbox(frame: np.ndarray, lo: np.ndarray, hi: np.ndarray):
    """src/geometry.jl:70-86: min/max over the 8 transformed corners, Float32 op order
    ((x*p1 + y*p2) + z*p3) + o."""
    f = np.asarray(frame, np.float32)
    x, y, z, o = f[0:3], f[3:6], f[6:9], f[9:12]
    out_lo = np.full(3, np.inf, np.float32)
    out_hi = np.full(3, -np.inf, np.float32)
    for cx in (lo[0], hi[0]):
        for cy in (lo[1], hi[1]):
            for cz in (lo[2], hi[2]):
                p = ((x * np.float32(cx) + y * np.float32(cy)) + z * np.float32(cz)) + o
                out_lo = np.minimum(out_lo, p)
                out_hi = np.maximum(out_hi, p)
    return out_lo, out_hi


def make_scene_bvh(scene: SceneData, high_quality: bool = False, no_parallel: bool = False) -> SceneBvh:
    shapes = [make_bvh(shape_bboxes(s), high_quality) for s in scene.shapes]
    boxes = np.zeros((len(scene.instances), 6), np.float32)
    for i, inst in enumerate(scene.instances):
        sb = shapes[int(inst["shape"]) - 1]
        root = sb.nodes[0]
        if shape_is_empty(scene.shapes[int(inst["shape"]) - 1]):
            raise ValueError(f"instance {i + 1} references an element-less shape: its Inf box hangs the "
                             f"reference's partition (SURVEY.md App. D); drop it first")
        lo, hi = transform_bbox(inst["frame"], root["bbox_min"], root["bbox_max"])
        boxes[i, :3], boxes[i, 3:] = lo, hi
    return SceneBvh(make_bvh(boxes, high_quality), shapes)


def shape_is_empty(shape: ShapeData) -> bool:
    return len(shape.triangles) == 0 and len(shape.quads) == 0
