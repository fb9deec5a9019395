"""`make_trace_lights` (src/trace.jl:117-187) for non-Julia hosts: area CDFs per emissive
instance (shape-local areas, sequential Float32 prefix sums) and the sin(theta)*max(texel) CDF per
textured environment (Q8). On the Julia host this stays Julia code; the resulting
`elements_cdf` vectors are an INPUT of the C ABI either way."""
from __future__ import annotations

from typing import List

import numpy as np

from .flatten import TraceLight
from .scene import SceneData, invalid_id

_f32 = np.float32


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(np.float32)


def _tri_area(p0, p1, p2):
    c = _cross((p1 - p0).astype(np.float32), (p2 - p0).astype(np.float32))
    d = ((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]).astype(np.float32) + c[:, 2] * c[:, 2]).astype(np.float32)
    return (np.sqrt(d).astype(np.float32) / _f32(2)).astype(np.float32)


def _seq_cumsum(a: np.ndarray) -> np.ndarray:
    # np.cumsum accumulates sequentially in the array dtype: cdf[i] = cdf[i-1] + a[i] in Float32
    return np.cumsum(np.asarray(a, np.float32), dtype=np.float32)


def _sin_f32(x: np.ndarray) -> np.ndarray:
    return np.sin(x.astype(np.float64)).astype(np.float32)


def make_trace_lights(scene: SceneData, params=None) -> List[TraceLight]:
    lights: List[TraceLight] = []
    for handle, inst in enumerate(scene.instances):
        mat = scene.materials[int(inst["material"]) - 1]
        if np.all(mat["emission"] == 0):
            continue
        shape = scene.shapes[int(inst["shape"]) - 1]
        if len(shape.triangles) == 0 and len(shape.quads) == 0:
            continue
        P = shape.positions
        cdf = None
        if len(shape.triangles):
            t = shape.triangles - 1
            cdf = _seq_cumsum(_tri_area(P[t[:, 0]], P[t[:, 1]], P[t[:, 2]]))
        if len(shape.quads):  # a second `if`, like the reference: quads overwrite
            q = shape.quads - 1
            area = (_tri_area(P[q[:, 0]], P[q[:, 1]], P[q[:, 3]]) +
                    _tri_area(P[q[:, 2]], P[q[:, 3]], P[q[:, 1]])).astype(np.float32)
            cdf = _seq_cumsum(area)
        lights.append(TraceLight(handle + 1, invalid_id, np.ascontiguousarray(cdf)))
    for handle, env in enumerate(scene.environments):
        if np.all(env["emission"] == 0):
            continue
        cdf = np.zeros(0, np.float32)
        tex_id = int(env["emission_tex"])
        if tex_id != invalid_id:
            tex = scene.textures[tex_id - 1]
            if tex.pixelsf is not None:
                texels = tex.pixelsf
            else:
                texels = (tex.pixelsb.astype(np.float32) / _f32(255)).astype(np.float32)
            j = (np.arange(tex.width * tex.height) // tex.width).astype(np.float32)
            th = ((j + _f32(0.5)) * _f32(np.pi)).astype(np.float32) / _f32(tex.height)
            value = texels.max(axis=1).astype(np.float32)  # maximum over RGBA (Q8)
            cdf = _seq_cumsum((value * _sin_f32(th.astype(np.float32))).astype(np.float32))
        lights.append(TraceLight(invalid_id, handle + 1, np.ascontiguousarray(cdf)))
    return lights


JT_LIGHTS_ENV_LUMINANCE = 1


def make_trace_lights_device(scene: SceneData, device: int = 0, env_luminance: bool = False) -> List[TraceLight]:
    """`make_trace_lights` on the GPU (jt_lights_create): element weights in parallel, the CDF as the
    reference's sequential Float32 prefix sum, so the arrays equal make_trace_lights' bit for bit. `env_luminance`
    weights environment texels by max(R, G, B) instead of max(R, G, B, A) (quirk Q8; not the reference's sample set)."""
    import ctypes as C

    from . import _abi as A
    from . import _lib
    from .flatten import FlatScene

    L = _lib.lib()
    flat = FlatScene(scene, None, None)
    h = C.c_void_p()
    _lib.check(L.jt_lights_create(flat.byref(), device, JT_LIGHTS_ENV_LUMINANCE if env_luminance else 0, C.byref(h)))
    try:
        descs, n = C.c_void_p(), C.c_int64()
        _lib.check(L.jt_lights_desc(h, C.byref(descs), C.byref(n)))
        arr = C.cast(descs, C.POINTER(A.jt_light_desc))
        out = []
        for i in range(n.value):
            d = arr[i]
            cdf = np.zeros(d.num_elements, np.float32)
            if d.num_elements:
                C.memmove(cdf.ctypes.data, d.elements_cdf, 4 * d.num_elements)
            out.append(TraceLight(int(d.instance), int(d.environment), cdf))
        return out
    finally:
        L.jt_lights_destroy(h)
