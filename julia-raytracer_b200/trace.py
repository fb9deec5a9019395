"""Host mirror of the reference's `Trace` module surface for the hot path (src/trace.jl):
`make_trace_lights` (:117), `make_trace_state` (:189), `trace_samples` (:215), `get_image` (:676).

`trace_samples` keeps the reference's signature (the three scratch-stack arguments are accepted
and ignored) but the loop body runs in libjtrace_b200.so: accumulators stay on the device between
calls and host arrays are refreshed on the final call (or on `state.sync()`)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _abi as A
from . import _lib
from .cli import Params
from .flatten import FlatScene, SceneBvh, TraceLight
from .lights import make_trace_lights  # noqa: F401  (re-exported like Trace.make_trace_lights)
from .scene import SceneData


def to_jt_params(params: Params, accumulate: int = 0) -> A.jt_params:
    p = A.jt_params()
    p.camera = int(params.camera) if not isinstance(params.camera, str) else 1
    p.resolution, p.samples, p.bounces = params.resolution, params.samples, params.bounces
    p.sampler, p.clamp = params.sampler, int(params.clamp)
    p.nocaustics, p.envhidden, p.tentfilter = int(params.nocaustics), int(params.envhidden), int(params.tentfilter)
    p.batch, p.bvhstacksize = params.batch, params.bvhstacksize
    p.traversal = 1 if getattr(params, "gpu_traversal", "wide") == "reference" else 0
    p.integrator = 1 if getattr(params, "gpu_integrator", "wavefront") == "megakernel" else 0
    p.seed = int(getattr(params, "gpu_seed", 0))
    p.accumulate = accumulate
    return p


def _counters_dict(c: A.jt_counters) -> dict:
    return dict(camera_paths=int(c.camera_paths), scene_rays=int(c.scene_rays),
                light_rays=int(c.light_rays), kernel_launches=int(c.kernel_launches),
                extend_us=int(c.extend_kernel_us), extend_launches=int(c.extend_launches),
                stolen_samples=int(c.stolen_samples), resumed_rays=int(c.resumed_rays))


def set_bvh_cache_dir(path: Optional[str]) -> None:
    """Directory in which jt_scene_create keeps finished wide BVHs (None / "" = off). SURVEY.md 8f N1."""
    _lib.check(_lib.lib().jt_set_bvh_cache_dir(str(path).encode() if path else None))


class NativeHostScene:
    """load_scene + make_scene_bvh + make_trace_lights done inside the library (jt_host_scene_*, SURVEY.md 8f N2 and
    the host halves of N1 / N4): what a C or C++ host uses instead of sceneio.py / bvh.py / lights.py. The arrays are
    byte-identical to the Python mirror's (tests/test_native_host.py). Pass it to DeviceScene / DeviceGroup in place
    of `scene` (with bvh = lights = None)."""

    def __init__(self, filename: str, high_quality_bvh: bool = False):
        self.L = _lib.lib()
        self.h = C.c_void_p()
        _lib.check(self.L.jt_host_scene_load(str(filename).encode(), C.byref(self.h)))
        _lib.check(self.L.jt_host_scene_build(self.h, int(bool(high_quality_bvh))))
        d = C.POINTER(A.jt_scene_desc)()
        _lib.check(self.L.jt_host_scene_desc(self.h, C.byref(d)))
        self.desc = d.contents
        self.notes = [self.L.jt_host_scene_note(self.h, i).decode()
                      for i in range(self.L.jt_host_scene_num_notes(self.h))]

    def byref(self):
        return C.byref(self.desc)

    def find_camera(self, name: str = "") -> int:
        c = C.c_int32()
        _lib.check(self.L.jt_host_scene_find_camera(self.h, str(name).encode(), C.byref(c)))
        return c.value

    def close(self):
        if getattr(self, "h", None):
            self.L.jt_host_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _flat(scene, bvh, lights):
    return scene if isinstance(scene, NativeHostScene) else FlatScene(scene, bvh, lights)


class DeviceScene:
    """(scene, bvh, lights) resident on one GPU: what `trace_samples` consumes."""

    def __init__(self, scene: SceneData, bvh: SceneBvh = None, lights: List[TraceLight] = None, device: int = 0):
        self.L = _lib.lib()
        self.flat = _flat(scene, bvh, lights)
        h = C.c_void_p()
        _lib.check(self.L.jt_scene_create(self.flat.byref(), device, C.byref(h)))
        self.h = h
        self.device = device
        self.scene = scene
        self._states = []
        self._owned = True

    @classmethod
    def borrowed(cls, handle, device: int, scene: SceneData) -> "DeviceScene":
        """A member scene of a DeviceGroup (owned by the group): parity hooks and per-device counters."""
        self = cls.__new__(cls)
        self.L = _lib.lib()
        self.flat = None
        self.h, self.device, self.scene = handle, device, scene
        self._states, self._owned = [], False
        return self

    def close(self):
        # states first (the library also tolerates the other order: they become orphans)
        for st in list(getattr(self, "_states", [])):
            st.close()
        if getattr(self, "h", None) and getattr(self, "_owned", True):
            self.L.jt_scene_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        s = A.jt_scene_stats()
        _lib.check(self.L.jt_scene_get_stats(self.h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in A.jt_scene_stats._fields_ if k != "_reserved"}

    def counters(self, reset: bool = False) -> dict:
        c = A.jt_counters()
        _lib.check(self.L.jt_scene_counters(self.h, C.byref(c), int(reset)))
        return _counters_dict(c)

    def synchronize(self):
        _lib.check(self.L.jt_synchronize(self.h))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        _lib.check(self.L.jt_elapsed_ms(self.h, C.byref(ms)))
        return float(ms.value)

    # -- parity hooks -------------------------------------------------------------------------
    def intersect(self, rays: np.ndarray, traversal: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=A.RAY_DTYPE)
        hits = np.zeros(len(rays), A.HIT_DTYPE)
        _lib.check(self.L.jt_intersect(self.h, rays.ctypes.data, len(rays), traversal, hits.ctypes.data))
        return hits

    def intersect_instance(self, rays: np.ndarray, instances: np.ndarray, traversal: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=A.RAY_DTYPE)
        instances = np.ascontiguousarray(instances, dtype=np.int64)
        hits = np.zeros(len(rays), A.HIT_DTYPE)
        _lib.check(self.L.jt_intersect_instance(self.h, rays.ctypes.data, instances.ctypes.data, len(rays),
                                                traversal, hits.ctypes.data))
        return hits

    def sample_camera(self, jp: A.jt_params, width: int, height: int, ij, puv_luv) -> np.ndarray:
        ij = np.ascontiguousarray(ij, np.int32)
        r = np.ascontiguousarray(puv_luv, np.float32)
        rays = np.zeros(len(ij), A.RAY_DTYPE)
        _lib.check(self.L.jt_sample_camera(self.h, C.byref(jp), width, height, ij.ctypes.data, r.ctypes.data,
                                           len(ij), rays.ctypes.data))
        return rays


class TraceState:
    """`TraceState` (src/trace.jl:87-96): host arrays in the reference's layouts + the device twin."""

    def __init__(self, dscene: DeviceScene, params: Params, accumulate: int = 0):
        self.dscene = dscene
        self.L = dscene.L
        self.jp = to_jt_params(params, accumulate)
        h = C.c_void_p()
        _lib.check(self.L.jt_state_create(dscene.h, C.byref(self.jp), C.byref(h)))
        self.h = h
        dscene._states.append(self)
        w, hh, s = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self.L.jt_state_size(self.h, C.byref(w), C.byref(hh), C.byref(s)))
        self.width, self.height = w.value, hh.value
        n = self.width * self.height
        self.image = np.zeros((n, 4), np.float32)   # Vector{Vec4f}
        self.albedo = np.zeros((n, 3), np.float32)  # Vector{Vec3f}
        self.normal = np.zeros((n, 3), np.float32)
        self.hits = np.zeros(n, np.int64)           # Vector{Int}
        self.denoised = np.zeros((0, 4), np.float32)

    @property
    def samples(self) -> int:
        s = C.c_int32()
        _lib.check(self.L.jt_state_size(self.h, None, None, C.byref(s)))
        return s.value

    def sync(self) -> "TraceState":
        """Refresh the host arrays from the device accumulators."""
        _lib.check(self.L.jt_state_download(self.h, self.image.ctypes.data, self.albedo.ctypes.data,
                                            self.normal.ctypes.data, self.hits.ctypes.data))
        return self

    def srgb8(self) -> np.ndarray:
        """(H, W, 4) uint8 sRGB image converted on the GPU (the reference's save path)."""
        out = np.zeros((self.height, self.width, 4), np.uint8)
        _lib.check(self.L.jt_state_download_srgb8(self.h, out.ctypes.data))
        return out

    def reset(self):
        _lib.check(self.L.jt_state_reset(self.h))

    def device_buffers(self):
        a, b, c, d, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self.L.jt_state_device_buffers(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(n)))
        return dict(image=a.value, albedo=b.value, normal=c.value, hits=d.value, count=n.value)

    def set_samples(self, samples: int):
        _lib.check(self.L.jt_state_set_samples(self.h, samples))

    def close(self):
        if getattr(self, "h", None):
            self.L.jt_state_destroy(self.h)
            self.h = None
            try:
                self.dscene._states.remove(self)
            except ValueError:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceGroup:
    """(scene, bvh, lights) replicated on several GPUs behind ONE host thread (jt_group): the scene is
    staged once, every device gets a worker thread, sample ranges are split over the members and the download merges
    the members' sum buffers with one fused peer-to-peer reduce + finalize kernel on the first device."""

    def __init__(self, scene: SceneData, bvh: SceneBvh, lights: List[TraceLight], devices):
        self.L = _lib.lib()
        self.flat = _flat(scene, bvh, lights)
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _lib.check(self.L.jt_group_create(self.flat.byref(), arr, len(self.devices), C.byref(h)))
        self.h = h
        self.scene = scene
        self._states = []

    def member(self, k: int) -> DeviceScene:
        h = C.c_void_p()
        _lib.check(self.L.jt_group_scene(self.h, k, C.byref(h)))
        return DeviceScene.borrowed(h, self.devices[k], self.scene)

    def stats(self) -> dict:
        s = A.jt_group_stats()
        _lib.check(self.L.jt_group_get_stats(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in A.jt_group_stats._fields_ if k != "_reserved"}
        d["scene"] = self.member(0).stats()
        return d

    def counters(self, reset: bool = False) -> dict:
        c = A.jt_counters()
        _lib.check(self.L.jt_group_counters(self.h, C.byref(c), int(reset)))
        return _counters_dict(c)

    def synchronize(self):
        _lib.check(self.L.jt_group_synchronize(self.h))

    def close(self):
        for st in list(getattr(self, "_states", [])):
            st.close()
        if getattr(self, "h", None):
            self.L.jt_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupState:
    """TraceState of a DeviceGroup: same host arrays, merged on download."""

    def __init__(self, group: DeviceGroup, params: Params):
        self.dscene = group
        self.L = group.L
        self.jp = to_jt_params(params, 1)
        h = C.c_void_p()
        _lib.check(self.L.jt_group_state_create(group.h, C.byref(self.jp), C.byref(h)))
        self.h = h
        group._states.append(self)
        w, hh = C.c_int32(), C.c_int32()
        _lib.check(self.L.jt_group_state_size(self.h, C.byref(w), C.byref(hh), None))
        self.width, self.height = w.value, hh.value
        n = self.width * self.height
        self.image = np.zeros((n, 4), np.float32)
        self.albedo = np.zeros((n, 3), np.float32)
        self.normal = np.zeros((n, 3), np.float32)
        self.hits = np.zeros(n, np.int64)
        self.denoised = np.zeros((0, 4), np.float32)

    @property
    def samples(self) -> int:
        s = C.c_int32()
        _lib.check(self.L.jt_group_state_size(self.h, None, None, C.byref(s)))
        return s.value

    def sync(self) -> "GroupState":
        _lib.check(self.L.jt_group_state_download(self.h, self.image.ctypes.data, self.albedo.ctypes.data,
                                                  self.normal.ctypes.data, self.hits.ctypes.data))
        return self

    def srgb8(self) -> np.ndarray:
        out = np.zeros((self.height, self.width, 4), np.uint8)
        _lib.check(self.L.jt_group_state_download_srgb8(self.h, out.ctypes.data))
        return out

    def reset(self):
        _lib.check(self.L.jt_group_state_reset(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.jt_group_state_destroy(self.h)
            self.h = None
            try:
                self.dscene._states.remove(self)
            except ValueError:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_trace_state(dscene: DeviceScene, params: Params, accumulate: int = 0) -> TraceState:
    """src/trace.jl:189-213 (sizes from resolution and the camera aspect; zeroed buffers)."""
    if isinstance(dscene, DeviceGroup):
        return GroupState(dscene, params)
    return TraceState(dscene, params, accumulate)


def trace_samples(state: TraceState, scene: DeviceScene, bvh=None, lights=None, params: Optional[Params] = None,
                  bvh_stacks=None, bvh_sub_stacks=None, volume_stacks=None) -> None:
    """src/trace.jl:215-274. `bvh` and `lights` already live inside `scene` (a DeviceScene); the
    scratch stacks are ignored. Enqueues `params.batch` more samples per pixel and returns."""
    jp = to_jt_params(params, state.jp.accumulate) if params is not None else state.jp
    L = state.L
    if isinstance(state, GroupState):
        _lib.check(L.jt_group_trace_samples(scene.h, state.h, C.byref(jp)))
    else:
        _lib.check(L.jt_trace_samples(scene.h, state.h, C.byref(jp)))
    if state.samples >= jp.samples:
        state.sync()


def trace_sample_range(state: TraceState, scene: DeviceScene, params: Params, begin: int, end: int) -> None:
    """The sharding unit: samples [begin, end) of every pixel."""
    jp = to_jt_params(params, state.jp.accumulate)
    if isinstance(state, GroupState):
        _lib.check(state.L.jt_group_trace_sample_range(scene.h, state.h, C.byref(jp), begin, end))
    else:
        _lib.check(state.L.jt_trace_sample_range(scene.h, state.h, C.byref(jp), begin, end))


def get_image(state: TraceState) -> np.ndarray:
    """src/trace.jl:676-690: (H, W, 4) linear RGBA view of state.image."""
    return state.image.reshape(state.height, state.width, 4)
