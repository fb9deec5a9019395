"""ctypes wrapper of tests/emu/libjt_emu.so -- TEST INFRASTRUCTURE: the device headers compiled
for the host so the no-GPU tier can single-step the CUDA code paths against the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import orc

A = orc.A
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        d = os.path.join(orc.ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", d])
        L = C.CDLL(os.path.join(d, "libjt_emu.so"))
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_last_error.restype = C.c_char_p
        L.emu_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.emu_intersect_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.emu_trace_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.emu_trace_wavefront.restype = C.c_int
        L.emu_trace_wavefront.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.emu_wide_counts.argtypes = [C.c_void_p, C.c_int]
        L.emu_set_suspend_every.argtypes = [C.c_int]
        L.emu_resumed_rays.argtypes = [C.c_int]
        L.emu_resumed_rays.restype = C.c_ulonglong
        L.emu_stolen_samples.argtypes = [C.c_int]
        L.emu_stolen_samples.restype = C.c_ulonglong
        _LIB = L
    return _LIB


class Emu:
    def __init__(self, scene, bvh, lights):
        self.flat = orc.flatten.FlatScene(scene, bvh, lights)
        self.L = lib()
        self.h = self.L.emu_create(self.flat.byref())
        if not self.h:
            raise RuntimeError(self.L.emu_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_destroy(self.h)
            self.h = None

    def stats(self):
        out = np.zeros(8, np.int64)
        self.L.emu_stats(self.h, out.ctypes.data)
        return dict(zip(["wide_nodes", "wide_tris", "inlined", "instanced", "depth_top", "depth_blas", "flattened", "from_cache"], map(int, out)))

    def intersect(self, rays, traversal=0):
        rays = np.ascontiguousarray(rays, dtype=A.RAY_DTYPE)
        hits = np.zeros(len(rays), A.HIT_DTYPE)
        self.L.emu_intersect(self.h, rays.ctypes.data, len(rays), traversal, hits.ctypes.data)
        return hits

    def intersect_instance(self, rays, instances, traversal=0):
        rays = np.ascontiguousarray(rays, dtype=A.RAY_DTYPE)
        instances = np.ascontiguousarray(instances, np.int64)
        hits = np.zeros(len(rays), A.HIT_DTYPE)
        self.L.emu_intersect_instance(self.h, rays.ctypes.data, instances.ctypes.data, len(rays), traversal, hits.ctypes.data)
        return hits

    def trace(self, params, width, height, begin, end, wavefront=False):
        n = width * height
        image = np.zeros((n, 4), np.float32)
        albedo = np.zeros((n, 4), np.float32)
        normal = np.zeros((n, 4), np.float32)
        hits = np.zeros(n, np.int32)
        cnt = np.zeros(2, np.uint64)
        cnt = np.zeros(3, np.uint64)
        fn = self.L.emu_trace_wavefront if wavefront else self.L.emu_trace_range
        fn(self.h, C.byref(params), width, height, begin, end, image.ctypes.data,
           albedo.ctypes.data, normal.ctypes.data, hits.ctypes.data, cnt.ctypes.data)
        return dict(image=image.reshape(height, width, 4), albedo=albedo.reshape(height, width, 4)[..., :3],
                    normal=normal.reshape(height, width, 4)[..., :3], hits=hits.reshape(height, width).astype(np.int64),
                    scene_rays=int(cnt[0]), light_rays=int(cnt[1]))


def wide_counts(reset=True):
    out = np.zeros(4, np.uint64)
    lib().emu_wide_counts(out.ctypes.data, int(reset))
    return dict(nodes=int(out[0]), prims=int(out[1]), instances=int(out[2]), xforms=int(out[3]))


def set_suspend_every(n: int):
    """> 0: wide-mode wavefront traces use the persistent extend kernel and park / resume every ray after n iterations."""
    lib().emu_set_suspend_every(int(n))


def resumed_rays(reset=True) -> int:
    return int(lib().emu_resumed_rays(int(reset)))


def stolen_samples(reset=True) -> int:
    """Samples traced by a slot that started the range on another pixel (sample-level work stealing)."""
    return int(lib().emu_stolen_samples(int(reset)))
