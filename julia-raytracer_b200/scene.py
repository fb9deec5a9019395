"""Host-side scene data model.

Mirrors the reference's Julia structs (`src/scene.jl:45-370`, `src/shape.jl:13-23`) as numpy
arrays whose memory layouts are the ones Julia uses for the corresponding isbits structs
(SURVEY.md Appendix B), so the same flat buffers can cross the C ABI from either host.

All ids stored here are 1-based with ``invalid_id = -1`` exactly like the reference
(`src/scene.jl:45`, `:95-96`); conversion to 0-based happens inside the library.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

invalid_id = -1  # src/scene.jl:45
min_roughness = np.float32(0.03) * np.float32(0.03)  # src/scene.jl:46

# MaterialType enum, src/scene.jl:191-200
MATTE, GLOSSY, REFLECTIVE, TRANSPARENT, REFRACTIVE, SUBSURFACE, VOLUMETRIC, GLTFPBR = range(8)
MATERIAL_TYPES = {  # src/scene.jl:201-211
    "matte": MATTE,
    "glossy": GLOSSY,
    "reflective": REFLECTIVE,
    "transparent": TRANSPARENT,
    "refractive": REFRACTIVE,
    "subsurface": SUBSURFACE,
    "volume": VOLUMETRIC,
    "volumetric": VOLUMETRIC,
    "gltfpbr": GLTFPBR,
}

# InstanceData, src/scene.jl:88-91 (64 B)
INSTANCE_DTYPE = np.dtype(
    {"names": ["frame", "shape", "material"],
     "formats": [("<f4", (12,)), "<i8", "<i8"],
     "offsets": [0, 48, 56], "itemsize": 64})

# MaterialData, src/scene.jl:213-229 (104 B)
MATERIAL_DTYPE = np.dtype(
    {"names": ["type", "emission", "color", "roughness", "metallic", "ior", "scattering",
               "scanisotropy", "trdepth", "opacity", "emission_tex", "color_tex",
               "roughness_tex", "scattering_tex", "normal_tex"],
     "formats": ["<i4", ("<f4", (3,)), ("<f4", (3,)), "<f4", "<f4", "<f4", ("<f4", (3,)),
                 "<f4", "<f4", "<f4", "<i8", "<i8", "<i8", "<i8", "<i8"],
     "offsets": [0, 4, 16, 28, 32, 36, 40, 52, 56, 60, 64, 72, 80, 88, 96],
     "itemsize": 104})

# EnvironmentData, src/scene.jl:117-120 (72 B)
ENVIRONMENT_DTYPE = np.dtype(
    {"names": ["frame", "emission", "emission_tex"],
     "formats": [("<f4", (12,)), ("<f4", (3,)), "<i8"],
     "offsets": [0, 48, 64], "itemsize": 72})

# BvhNode, src/bvh.jl:34-39 (40 B)
BVHNODE_DTYPE = np.dtype(
    {"names": ["bbox_min", "bbox_max", "start", "num", "axis", "internal"],
     "formats": [("<f4", (3,)), ("<f4", (3,)), "<i8", "<i2", "<i1", "u1"],
     "offsets": [0, 12, 24, 32, 34, 35], "itemsize": 40})

IDENTITY_FRAME = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float32)


def frame_from_json(values) -> np.ndarray:
    """`Frame3f(array)` (src/math.jl:47-60): 12 floats -> x,y,z,o columns; any other length
    gives identity axes and a zero origin."""
    if values is None or len(values) != 12:
        return IDENTITY_FRAME.copy()
    return np.asarray(values, dtype=np.float64).astype(np.float32)


@dataclasses.dataclass
class CameraData:  # src/scene.jl:48-56
    frame: np.ndarray
    orthographic: bool = False
    lens: np.float32 = np.float32(0.050)
    film: np.float32 = np.float32(0.036)
    aspect: np.float32 = np.float32(1.5)
    focus: np.float32 = np.float32(10000)
    aperture: np.float32 = np.float32(0)
    name: str = ""


@dataclasses.dataclass
class TextureData:  # src/scene.jl:146-151
    width: int
    height: int
    linear: bool
    pixelsf: Optional[np.ndarray]  # (W*H, 4) float32, row-major, top row first
    pixelsb: Optional[np.ndarray]  # (W*H, 4) uint8


@dataclasses.dataclass
class ShapeData:  # src/shape.jl:13-23 (points/lines/radius/tangents are never populated)
    positions: np.ndarray  # (n,3) f32
    normals: np.ndarray    # (n,3) f32 or (0,3)
    texcoords: np.ndarray  # (n,2) f32 or (0,2)
    colors: np.ndarray     # (n,4) f32 or (0,4)
    triangles: np.ndarray  # (m,3) int64, 1-based
    quads: np.ndarray      # (m,4) int64, 1-based

    @staticmethod
    def empty() -> "ShapeData":
        z = np.zeros
        return ShapeData(z((0, 3), np.float32), z((0, 3), np.float32), z((0, 2), np.float32),
                         z((0, 4), np.float32), z((0, 3), np.int64), z((0, 4), np.int64))

    @property
    def num_elements(self) -> int:
        return len(self.triangles) if len(self.triangles) else len(self.quads)


@dataclasses.dataclass
class SceneData:  # src/scene.jl:337-356
    cameras: List[CameraData]
    instances: np.ndarray       # INSTANCE_DTYPE
    environments: np.ndarray    # ENVIRONMENT_DTYPE
    shapes: List[ShapeData]
    textures: List[TextureData]
    materials: np.ndarray       # MATERIAL_DTYPE
    notes: List[str] = dataclasses.field(default_factory=list)  # missing-asset substitutions


def find_camera(scene: SceneData, name: str) -> int:
    """src/scene.jl:358-370. Returns a 1-based index (or invalid_id)."""
    if len(scene.cameras) == 0:
        return invalid_id
    for n in [name, "default", "camera", "camera0", "camera1"]:
        for i, cam in enumerate(scene.cameras):
            if cam.name == n:
                return i + 1
    return 1


def image_size(camera: CameraData, resolution: int):
    """`make_trace_state` sizing, src/trace.jl:189-197: Float32 division, round half-to-even."""
    res = np.float32(resolution)
    if camera.aspect >= 1:
        width = int(resolution)
        height = int(np.rint(res / np.float32(camera.aspect)))
    else:
        height = int(resolution)
        width = int(np.rint(res * np.float32(camera.aspect)))
    return width, height
