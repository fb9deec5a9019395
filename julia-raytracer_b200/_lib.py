"""Loader of libjtrace_b200.so (the C ABI of include/jtrace_b200.h) through ctypes.

There is NO CPU fallback: if the shared library is missing this raises, and every compute entry
point returns JT_ERR_NO_DEVICE when no CUDA device is visible."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JTRACE_B200_LIB") or os.path.join(_HERE, "libjtrace_b200.so")  # override: tuning variants

# every symbol include/jtrace_b200.h declares
EXPORTS = [
    "jt_last_error", "jt_version", "jt_device_count", "jt_scene_create", "jt_scene_destroy",
    "jt_scene_counters", "jt_scene_get_stats", "jt_state_create", "jt_state_destroy",
    "jt_state_size", "jt_state_reset", "jt_state_download", "jt_state_device_buffers",
    "jt_state_set_samples", "jt_trace_samples", "jt_trace_sample_range", "jt_synchronize",
    "jt_elapsed_ms", "jt_intersect", "jt_intersect_instance", "jt_sample_camera",
    "jt_intersect_device", "jt_make_bvh", "jt_state_download_srgb8",
    "jt_group_create", "jt_group_destroy", "jt_group_get_stats", "jt_group_scene", "jt_group_counters",
    "jt_group_state_create", "jt_group_state_destroy", "jt_group_state_size", "jt_group_state_reset",
    "jt_group_trace_samples", "jt_group_trace_sample_range", "jt_group_synchronize",
    "jt_group_state_download", "jt_group_state_download_srgb8", "jt_probe_read_bandwidth",
    "jt_host_scene_load", "jt_host_scene_build", "jt_host_scene_desc", "jt_host_scene_find_camera",
    "jt_host_scene_num_notes", "jt_host_scene_note", "jt_host_scene_destroy", "jt_set_bvh_cache_dir",
    "jt_lights_create", "jt_lights_desc", "jt_lights_destroy",
]


class JtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libjtrace_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JtError(-100, f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                            f"or `make -C julia-raytracer_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float)
    L.jt_last_error.restype = C.c_char_p
    L.jt_version.restype = C.c_char_p
    L.jt_device_count.restype = C.c_int
    L.jt_scene_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.jt_scene_destroy.argtypes = [vp]
    L.jt_scene_destroy.restype = None
    L.jt_scene_counters.argtypes = [vp, C.POINTER(A.jt_counters), C.c_int]
    L.jt_scene_get_stats.argtypes = [vp, C.POINTER(A.jt_scene_stats)]
    L.jt_state_create.argtypes = [vp, C.POINTER(A.jt_params), C.POINTER(vp)]
    L.jt_state_destroy.argtypes = [vp]
    L.jt_state_destroy.restype = None
    L.jt_state_size.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.jt_state_reset.argtypes = [vp]
    L.jt_state_download.argtypes = [vp, vp, vp, vp, vp]
    L.jt_state_download_srgb8.argtypes = [vp, vp]
    L.jt_state_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                          C.POINTER(i64)]
    L.jt_state_set_samples.argtypes = [vp, i32]
    L.jt_trace_samples.argtypes = [vp, vp, C.POINTER(A.jt_params)]
    L.jt_trace_sample_range.argtypes = [vp, vp, C.POINTER(A.jt_params), i32, i32]
    L.jt_synchronize.argtypes = [vp]
    L.jt_elapsed_ms.argtypes = [vp, f32p]
    L.jt_intersect.argtypes = [vp, vp, i64, C.c_int, vp]
    L.jt_intersect_instance.argtypes = [vp, vp, vp, i64, C.c_int, vp]
    L.jt_sample_camera.argtypes = [vp, C.POINTER(A.jt_params), i32, i32, vp, vp, i64, vp]
    L.jt_intersect_device.argtypes = [vp, vp, i64, C.c_int, vp]
    L.jt_make_bvh.argtypes = [vp, i64, C.c_int, vp, C.POINTER(i64), vp]
    L.jt_set_bvh_cache_dir.argtypes = [C.c_char_p]
    L.jt_lights_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.jt_lights_desc.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.jt_lights_destroy.argtypes = [vp]
    L.jt_lights_destroy.restype = None
    L.jt_host_scene_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.jt_host_scene_build.argtypes = [vp, C.c_int]
    L.jt_host_scene_desc.argtypes = [vp, C.POINTER(C.POINTER(A.jt_scene_desc))]
    L.jt_host_scene_find_camera.argtypes = [vp, C.c_char_p, C.POINTER(i32)]
    L.jt_host_scene_num_notes.argtypes = [vp]
    L.jt_host_scene_note.argtypes = [vp, C.c_int]
    L.jt_host_scene_note.restype = C.c_char_p
    L.jt_host_scene_destroy.argtypes = [vp]
    L.jt_host_scene_destroy.restype = None
    L.jt_probe_read_bandwidth.argtypes = [C.c_int, i64, C.c_int, f32p]
    L.jt_group_create.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.jt_group_destroy.argtypes = [vp]
    L.jt_group_destroy.restype = None
    L.jt_group_get_stats.argtypes = [vp, C.POINTER(A.jt_group_stats)]
    L.jt_group_scene.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.jt_group_counters.argtypes = [vp, C.POINTER(A.jt_counters), C.c_int]
    L.jt_group_state_create.argtypes = [vp, C.POINTER(A.jt_params), C.POINTER(vp)]
    L.jt_group_state_destroy.argtypes = [vp]
    L.jt_group_state_destroy.restype = None
    L.jt_group_state_size.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.jt_group_state_reset.argtypes = [vp]
    L.jt_group_trace_samples.argtypes = [vp, vp, C.POINTER(A.jt_params)]
    L.jt_group_trace_sample_range.argtypes = [vp, vp, C.POINTER(A.jt_params), i32, i32]
    L.jt_group_synchronize.argtypes = [vp]
    L.jt_group_state_download.argtypes = [vp, vp, vp, vp, vp]
    L.jt_group_state_download_srgb8.argtypes = [vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("jt_device_count",):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise JtError(rc, lib().jt_last_error().decode("utf-8", "replace"))
