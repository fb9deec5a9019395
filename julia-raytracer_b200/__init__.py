"""jtrace-b200: B200-native drop-in for julia-raytracer's per-pixel render loop.

Host-side mirror (Python) of the reference's `Jtrace` module surface for the hot path:
`load_scene`, `find_camera`, `make_scene_bvh`, `make_trace_lights`, `make_trace_state`,
`trace_samples`, `get_image`, `save_image`, `main` (src/jtrace.jl:23-30). The render loop
itself runs in `libjtrace_b200.so` (hand-written CUDA for sm_100a) behind the C ABI declared
in `include/jtrace_b200.h`; there is no CPU fallback.
"""
from .scene import (SceneData, ShapeData, TextureData, CameraData, find_camera, image_size,
                    invalid_id)
from .sceneio import load_scene, save_image, load_packed, save_packed
from .cli import Params, parse_cli_args

__all__ = [
    "SceneData", "ShapeData", "TextureData", "CameraData", "find_camera", "image_size",
    "invalid_id", "load_scene", "save_image", "load_packed", "save_packed", "Params",
    "parse_cli_args",
]
