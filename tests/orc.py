"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE. Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
jt = importlib.import_module("julia-raytracer_b200")
A = importlib.import_module("julia-raytracer_b200._abi")
flatten = importlib.import_module("julia-raytracer_b200.flatten")

_LIB = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_bvh_num_nodes.restype = C.c_int64
        L.orc_bvh_num_nodes.argtypes = [C.c_void_p, C.c_int64]
        L.orc_bvh_num_primitives.restype = C.c_int64
        L.orc_bvh_num_primitives.argtypes = [C.c_void_p, C.c_int64]
        L.orc_bvh_get.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_num_lights.restype = C.c_int64
        L.orc_num_lights.argtypes = [C.c_void_p]
        L.orc_light_info.restype = C.c_int64
        L.orc_light_info.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_light_cdf.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_intersect_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_sample_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_make_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_trace_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int]
        L.orc_trace_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_trace_pixel.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_trace_pixel_rays.restype = C.c_int64
        L.orc_trace_pixel_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_int64]
        L.orc_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_bsdf_eval.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
        L.orc_bsdf_sample.argtypes = [C.c_void_p] * 3 + [C.c_float] * 3 + [C.c_int, C.c_void_p]
        L.orc_rng_float.restype = C.c_float
        L.orc_rng_float.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_fmath_array.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _LIB = L
    return _LIB


COUNTER_NAMES = ["scene_rays", "light_rays", "camera_paths", "tlas_nodes", "blas_nodes",
                 "instance_visits", "tri_tests", "quad_tests", "probe_blas_nodes", "probe_tri_tests",
                 "probe_quad_tests"]


def make_params(**kw) -> "A.jt_params":
    p = A.jt_params()
    p.camera, p.resolution, p.samples, p.bounces, p.sampler, p.clamp = 1, 1280, 512, 8, 1, 10
    p.batch, p.bvhstacksize = 1, 128
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Oracle:
    """The CPU restatement bound to one scene."""

    def __init__(self, scene, bvh=None, lights=None, high_quality=False):
        self.flat = flatten.FlatScene(scene, bvh, lights)
        self.L = lib()
        self.h = self.L.orc_create(self.flat.byref(), int(bvh is not None), int(lights is not None),
                                   int(high_quality))
        self.scene = scene
        self.width = self.height = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    # -- host-step restatements -----------------------------------------------------------------
    def get_bvh(self, shape=0):
        n = self.L.orc_bvh_num_nodes(self.h, shape)
        m = self.L.orc_bvh_num_primitives(self.h, shape)
        nodes = np.zeros(n, jt.scene.BVHNODE_DTYPE)
        prims = np.zeros(m, np.int64)
        self.L.orc_bvh_get(self.h, shape, nodes.ctypes.data, prims.ctypes.data)
        return nodes, prims

    def get_lights(self):
        out = []
        for i in range(self.L.orc_num_lights(self.h)):
            a, b = C.c_int64(), C.c_int64()
            n = self.L.orc_light_info(self.h, i, C.byref(a), C.byref(b))
            cdf = np.zeros(n, np.float32)
            self.L.orc_light_cdf(self.h, i, cdf.ctypes.data)
            out.append((a.value, b.value, cdf))
        return out

    # -- identical rays -------------------------------------------------------------------------
    def intersect(self, rays: np.ndarray, threads=0, counters=False):
        rays = np.ascontiguousarray(rays, dtype=A.RAY_DTYPE)
        hits = np.zeros(len(rays), A.HIT_DTYPE)
        cnt = np.zeros(11, np.uint64)
        self.L.orc_intersect(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, cnt.ctypes.data, threads)
        return (hits, dict(zip(COUNTER_NAMES, (int(x) for x in cnt)))) if counters else hits

    def intersect_instance(self, rays: np.ndarray, instances: np.ndarray):
        rays = np.ascontiguousarray(rays, dtype=A.RAY_DTYPE)
        instances = np.ascontiguousarray(instances, dtype=np.int64)
        hits = np.zeros(len(rays), A.HIT_DTYPE)
        self.L.orc_intersect_instance(self.h, rays.ctypes.data, instances.ctypes.data, len(rays), hits.ctypes.data)
        return hits

    def sample_camera(self, params, width, height, ij, puv_luv):
        ij = np.ascontiguousarray(ij, np.int32)
        r = np.ascontiguousarray(puv_luv, np.float32)
        rays = np.zeros(len(ij), A.RAY_DTYPE)
        self.L.orc_sample_camera(self.h, params.camera, params.tentfilter, width, height, ij.ctypes.data,
                                 r.ctypes.data, len(ij), rays.ctypes.data)
        return rays

    # -- render loop ----------------------------------------------------------------------------
    def make_state(self, params):
        w, h = C.c_int32(), C.c_int32()
        self.L.orc_make_state(self.h, C.byref(params), C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value
        return self.width, self.height

    def trace_range(self, params, begin, end, threads=0):
        self.L.orc_trace_range(self.h, C.byref(params), begin, end, threads)

    def trace_samples(self, params, threads=0):
        self.L.orc_trace_samples(self.h, C.byref(params), threads)

    def trace_pixel(self, params, i, j, sample):
        out = np.zeros(5, np.float32)
        self.L.orc_trace_pixel(self.h, C.byref(params), i, j, sample, out.ctypes.data)
        return out

    def trace_pixel_rays(self, params, i, j, sample, max_rays=4096):
        """(rays, instances): every closest-hit query of one (pixel, sample) path, in call order; instances[k] is the
        probed instance (1-based, intersect_instance_bvh) or -1 for a scene query."""
        rays = np.zeros(max_rays, A.RAY_DTYPE)
        inst = np.zeros(max_rays, np.int64)
        n = self.L.orc_trace_pixel_rays(self.h, C.byref(params), i, j, sample, rays.ctypes.data, inst.ctypes.data,
                                        max_rays)
        n = min(int(n), max_rays)
        return rays[:n], inst[:n]

    def get_state(self):
        n = self.width * self.height
        image = np.zeros((self.height, self.width, 4), np.float32)
        albedo = np.zeros((self.height, self.width, 3), np.float32)
        normal = np.zeros((self.height, self.width, 3), np.float32)
        hits = np.zeros((self.height, self.width), np.int64)
        s = C.c_int32()
        self.L.orc_get_state(self.h, image.ctypes.data, albedo.ctypes.data, normal.ctypes.data,
                             hits.ctypes.data, C.byref(s))
        return dict(image=image, albedo=albedo, normal=normal, hits=hits, samples=s.value)

    def counters(self, reset=False):
        cnt = np.zeros(11, np.uint64)
        self.L.orc_get_counters(self.h, cnt.ctypes.data, int(reset))
        return dict(zip(COUNTER_NAMES, (int(x) for x in cnt)))


def algorithmic_bytes(c: dict) -> int:
    """32 (ray) + 24 (hit) per query + 32/node + 48/instance visit + 36/tri + 48/quad."""
    rays = c["scene_rays"] + c["light_rays"]
    return (56 * rays + 32 * (c["tlas_nodes"] + c["blas_nodes"]) + 48 * c["instance_visits"]
            + 36 * c["tri_tests"] + 48 * c["quad_tests"])


def algorithmic_bytes_probe(c: dict) -> int:
    """The part of algorithmic_bytes() spent in intersect_instance_bvh (light-pdf probes)."""
    return (56 * c["light_rays"] + 32 * c["probe_blas_nodes"] + 48 * c["light_rays"]
            + 36 * c["probe_tri_tests"] + 48 * c["probe_quad_tests"])


def algorithmic_bytes_scene(c: dict) -> int:
    """The part spent in intersect_scene_bvh (what the extend kernel replaces)."""
    return algorithmic_bytes(c) - algorithmic_bytes_probe(c)


def fmath(fn: int, x, y=None):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(x)
    yy = np.ascontiguousarray(y, np.float32) if y is not None else None
    lib().orc_fmath_array(fn, x.ctypes.data, yy.ctypes.data if yy is not None else None, x.size, out.ctypes.data)
    return out
