"""ctypes mirror of include/jtrace_b200.h (struct layouts only; no library is loaded here)."""
from __future__ import annotations

import ctypes as C

c_f3 = C.c_float * 3


class jt_frame(C.Structure):
    _fields_ = [("x", c_f3), ("y", c_f3), ("z", c_f3), ("o", c_f3)]


class jt_bvh_node(C.Structure):
    _fields_ = [("bbox_min", c_f3), ("bbox_max", c_f3), ("start", C.c_int64), ("num", C.c_int16),
                ("axis", C.c_int8), ("internal", C.c_uint8), ("_pad", C.c_uint8 * 4)]


class jt_bvh_desc(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("num_nodes", C.c_int64), ("primitives", C.c_void_p),
                ("num_primitives", C.c_int64)]


class jt_instance(C.Structure):
    _fields_ = [("frame", jt_frame), ("shape", C.c_int64), ("material", C.c_int64)]


class jt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("emission", c_f3), ("color", c_f3), ("roughness", C.c_float),
                ("metallic", C.c_float), ("ior", C.c_float), ("scattering", c_f3),
                ("scanisotropy", C.c_float), ("trdepth", C.c_float), ("opacity", C.c_float),
                ("emission_tex", C.c_int64), ("color_tex", C.c_int64), ("roughness_tex", C.c_int64),
                ("scattering_tex", C.c_int64), ("normal_tex", C.c_int64)]


class jt_environment(C.Structure):
    _fields_ = [("frame", jt_frame), ("emission", c_f3), ("_pad", C.c_int32),
                ("emission_tex", C.c_int64)]


class jt_camera(C.Structure):
    _fields_ = [("frame", jt_frame), ("orthographic", C.c_int32), ("lens", C.c_float),
                ("film", C.c_float), ("aspect", C.c_float), ("focus", C.c_float),
                ("aperture", C.c_float)]


class jt_texture_desc(C.Structure):
    _fields_ = [("width", C.c_int64), ("height", C.c_int64), ("linear", C.c_int32),
                ("_pad", C.c_int32), ("pixelsf", C.c_void_p), ("pixelsb", C.c_void_p)]


class jt_shape_desc(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("num_positions", C.c_int64),
                ("normals", C.c_void_p), ("num_normals", C.c_int64),
                ("texcoords", C.c_void_p), ("num_texcoords", C.c_int64),
                ("colors", C.c_void_p), ("num_colors", C.c_int64),
                ("triangles", C.c_void_p), ("num_triangles", C.c_int64),
                ("quads", C.c_void_p), ("num_quads", C.c_int64),
                ("bvh", jt_bvh_desc)]


class jt_light_desc(C.Structure):
    _fields_ = [("instance", C.c_int64), ("environment", C.c_int64), ("elements_cdf", C.c_void_p),
                ("num_elements", C.c_int64)]


class jt_scene_desc(C.Structure):
    _fields_ = [("num_cameras", C.c_int64), ("cameras", C.c_void_p),
                ("num_instances", C.c_int64), ("instances", C.c_void_p),
                ("num_environments", C.c_int64), ("environments", C.c_void_p),
                ("num_shapes", C.c_int64), ("shapes", C.c_void_p),
                ("num_textures", C.c_int64), ("textures", C.c_void_p),
                ("num_materials", C.c_int64), ("materials", C.c_void_p),
                ("num_lights", C.c_int64), ("lights", C.c_void_p),
                ("bvh", jt_bvh_desc),
                ("srgb_to_rgb_lut", C.c_void_p)]


class jt_params(C.Structure):
    _fields_ = [("camera", C.c_int32), ("resolution", C.c_int32), ("samples", C.c_int32),
                ("bounces", C.c_int32), ("sampler", C.c_int32), ("clamp", C.c_int32),
                ("nocaustics", C.c_int32), ("envhidden", C.c_int32), ("tentfilter", C.c_int32),
                ("batch", C.c_int32), ("bvhstacksize", C.c_int32), ("traversal", C.c_int32),
                ("seed", C.c_uint64), ("accumulate", C.c_int32), ("integrator", C.c_int32),
                ("_reserved", C.c_int32 * 6)]


class jt_ray(C.Structure):
    _fields_ = [("o", c_f3), ("d", c_f3), ("tmin", C.c_float), ("tmax", C.c_float)]


class jt_hit(C.Structure):
    _fields_ = [("instance", C.c_int64), ("element", C.c_int64), ("uv", C.c_float * 2),
                ("distance", C.c_float), ("hit", C.c_uint8), ("_pad", C.c_uint8 * 3)]


class jt_counters(C.Structure):
    _fields_ = [("camera_paths", C.c_uint64), ("scene_rays", C.c_uint64), ("light_rays", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("extend_kernel_us", C.c_uint64),
                ("extend_launches", C.c_uint64), ("stolen_samples", C.c_uint64), ("resumed_rays", C.c_uint64)]


class jt_scene_stats(C.Structure):
    _fields_ = [("wide_nodes", C.c_int64), ("wide_node_bytes", C.c_int64),
                ("prim_records", C.c_int64), ("prim_record_bytes", C.c_int64),
                ("inlined_instances", C.c_int64), ("instanced_instances", C.c_int64),
                ("texture_bytes", C.c_int64), ("total_device_bytes", C.c_int64),
                ("wide_depth_top", C.c_int64), ("wide_depth_blas", C.c_int64), ("opened_instances", C.c_int64),
                ("wide_bvh_from_cache", C.c_int64), ("_reserved", C.c_int64 * 4)]


class jt_group_stats(C.Structure):
    _fields_ = [("members", C.c_int32), ("distinct_devices", C.c_int32), ("peer_members", C.c_int32),
                ("staged_members", C.c_int32), ("reduce_bytes_remote", C.c_int64), ("downloads", C.c_int64),
                ("stage_seconds", C.c_double), ("upload_seconds", C.c_double), ("_reserved", C.c_int64 * 4)]


assert C.sizeof(jt_frame) == 48 and C.sizeof(jt_bvh_node) == 40 and C.sizeof(jt_instance) == 64
assert C.sizeof(jt_material) == 104 and C.sizeof(jt_environment) == 72
assert C.sizeof(jt_ray) == 32 and C.sizeof(jt_hit) == 32 and C.sizeof(jt_params) == 88
assert C.sizeof(jt_counters) == 64 and C.sizeof(jt_scene_stats) == 128 and C.sizeof(jt_group_stats) == 80

import numpy as np  # noqa: E402

RAY_DTYPE = np.dtype([("o", "<f4", (3,)), ("d", "<f4", (3,)), ("tmin", "<f4"), ("tmax", "<f4")])
HIT_DTYPE = np.dtype({"names": ["instance", "element", "uv", "distance", "hit"],
                      "formats": ["<i8", "<i8", ("<f4", (2,)), "<f4", "u1"],
                      "offsets": [0, 8, 16, 24, 28], "itemsize": 32})
