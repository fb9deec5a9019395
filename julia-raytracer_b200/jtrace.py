"""`Jtrace.main` (src/jtrace.jl:31-116) with the render loop on the GPU.

Same phases, same prints: load scene -> find camera -> build bvh -> make lights -> make state ->
`samples / batch` calls of `trace_samples` -> get_image -> save_image."""
from __future__ import annotations

import math
import time
from typing import Optional, Union

from .bvh import make_scene_bvh
from .cli import Params, parse_cli_args
from .lights import make_trace_lights, make_trace_lights_device
from .scene import find_camera
from .sceneio import load_scene, save_image, save_srgb8  # noqa: F401
from . import _lib
from .trace import DeviceGroup, DeviceScene, NativeHostScene, set_bvh_cache_dir, get_image, make_trace_state, trace_samples


def format_seconds(seconds: float) -> str:
    """src/utils.jl:10-32."""
    hours = math.floor(seconds / 3600)
    minutes = math.floor((seconds - hours * 3600) / 60)
    seconds = seconds - hours * 3600 - minutes * 60
    i_seconds = math.floor(seconds)
    ms = int(round((seconds - i_seconds) * 1000))
    if hours == 0:
        if minutes == 0:
            return f"{i_seconds:02d}.{ms:03d}"
        return f"{minutes:02d}:{i_seconds:02d}.{ms:03d}"
    return f"{hours:02d}:{minutes:02d}:{i_seconds:02d}.{ms:03d}"


def parse_devices(spec) -> list:
    """'' -> [], 'all' -> every visible device, '0,2,3' -> [0, 2, 3]; lists pass through."""
    if isinstance(spec, (list, tuple)):
        return [int(d) for d in spec]
    spec = str(spec).strip()
    if spec == "":
        return []
    if spec == "all":
        return list(range(_lib.lib().jt_device_count()))
    return [int(x) for x in spec.split(",") if x.strip() != ""]


def main(params: Union[str, Params, None], device: int = 0, devices=None) -> Optional[dict]:
    if isinstance(params, str):
        params = parse_cli_args(params)
    if params is None:
        return None
    if params.addsky:
        print("addsky is not yet supported")
        params.addsky = False
    if params.envname != "":
        print("envname is not yet supported")
        params.envname = ""
    if params.denoise:
        print("denoise is not yet supported")
        params.denoise = False
    render_start = time.time()
    print(f"loading scene {params.scene}...")
    t0 = time.time()
    native = bool(getattr(params, "gpu_native_host", False)) and not str(params.scene).endswith(".jtscene")
    if native:  # load_scene + make_scene_bvh + make_trace_lights inside the library (jt_host_scene_*)
        scene = NativeHostScene(params.scene, params.highqualitybvh)
        bvh = lights = None
        print(f"loaded scene, built bvh, made lights (native host) in {format_seconds(time.time() - t0)}")
        for note in scene.notes:
            print(f"    note: {note}")
        print("finding camera...")
        params.camera = scene.find_camera(params.camera if isinstance(params.camera, str) else "")
    else:
        scene = load_scene(params.scene, params.noparallel, verbose=True)
        print(f"loaded scene in {format_seconds(time.time() - t0)}")
        for note in scene.notes:
            print(f"    note: {note}")
        print("finding camera...")
        params.camera = find_camera(scene, params.camera if isinstance(params.camera, str) else "")
        print("building bvh...")
        t0 = time.time()
        bvh = make_scene_bvh(scene, params.highqualitybvh, params.noparallel)
        print(f"built bvh in {format_seconds(time.time() - t0)}")
        print("making lights...")
        if getattr(params, "gpu_device_lights", False):  # N4: element weights + sequential CDFs on the GPU
            lights = make_trace_lights_device(scene, device, env_luminance=bool(getattr(params, "gpu_env_importance", False)))
        else:
            lights = make_trace_lights(scene, params)
    print("uploading scene to the GPU...")
    t0 = time.time()
    if getattr(params, "gpu_bvh_cache", ""):
        set_bvh_cache_dir(params.gpu_bvh_cache)
    devs = parse_devices(devices if devices is not None else getattr(params, "gpu_devices", ""))
    if len(devs) > 1:  # one host thread, N devices: the sample axis is sharded inside the library (jt_group)
        dscene = DeviceGroup(scene, bvh, lights, devs)
        print(f"    sharding samples over devices {devs}")
    else:
        dscene = DeviceScene(scene, bvh, lights, devs[0] if devs else device)
    print(f"uploaded in {format_seconds(time.time() - t0)}")
    print("making state...")
    state = make_trace_state(dscene, params)
    print("tracing samples...")
    sampling_start = time.time()
    for _sample in range(1, params.samples + 1, max(params.batch, 1)):
        batch_start = time.time()
        trace_samples(state, dscene, bvh, lights, params, None, None, None)
        now = time.time()
        done = state.samples
        print("sample %3d/%3d in %s ETC: %s" % (
            done, params.samples, format_seconds(now - batch_start),
            format_seconds((now - sampling_start) / max(done, 1) * (params.samples - done))))
    dscene.synchronize()
    render_s = time.time() - sampling_start
    print("rendered in %s (%.3fs)" % (format_seconds(render_s), render_s))
    state.sync()
    c = dscene.counters()
    rays = c["scene_rays"] + c["light_rays"]
    print("    %.1f Msamples/s, %.1f Mrays/s (%d scene + %d light-probe rays)" % (
        c["camera_paths"] / render_s / 1e6, rays / render_s / 1e6, c["scene_rays"], c["light_rays"]))
    print("saving image...")
    image = get_image(state)
    save_srgb8(params.output, state.srgb8())  # rgb_to_srgb + 8-bit quantisation done on the GPU (N3)
    print("saved image to", params.output)
    print(f"total time: {format_seconds(time.time() - render_start)}")
    return dict(image=image, state=state, counters=c, render_seconds=render_s)
