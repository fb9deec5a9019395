"""Command line of the reference, kept intact (src/cli.jl:12-138).

Same 18 options, same defaults, Bool options take a value (`--noparallel true`) like the
reference's `arg_type = Bool`. `--shader` is accepted as an alias of `--sampler` because
BASELINE.json's north_star spells it that way. Extra, GPU-only options are
prefixed `--gpu-` so the reference's own command lines parse unchanged.
"""
from __future__ import annotations

import argparse
import dataclasses
import shlex
from typing import Any, Optional, Sequence, Union

SAMPLER_TYPES = ["path", "naive"]  # src/cli.jl:88


def _bool(v: str) -> bool:
    s = str(v).strip().lower()
    if s in ("true", "1", "yes"):
        return True
    if s in ("false", "0", "no"):
        return False
    raise argparse.ArgumentTypeError(f"invalid Bool value: {v!r}")


def make_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="Jtrace.main", allow_abbrev=False)
    p.add_argument("--scene", type=str, required=True, help="scene filename")
    p.add_argument("--output", type=str, default="tests/test_scene.png", help="output filename")
    p.add_argument("--camera", type=str, default="", help="camera name")
    p.add_argument("--addsky", type=_bool, default=False, help="add sky")
    p.add_argument("--envname", type=str, default="", help="add environment")
    p.add_argument("--resolution", type=int, default=1280, help="image resolution")
    p.add_argument("--samples", type=int, default=512, help="number of samples")
    p.add_argument("--bounces", type=int, default=8, help="number of bounces")
    p.add_argument("--denoise", type=_bool, default=False, help="enable denoiser")
    p.add_argument("--noparallel", type=_bool, default=False, help="disable threading")
    p.add_argument("--highqualitybvh", type=_bool, default=False, help="enable high quality bvh")
    p.add_argument("--envhidden", type=_bool, default=False, help="hide environment")
    p.add_argument("--tentfilter", type=_bool, default=False, help="filter image")
    p.add_argument("--sampler", "--shader", dest="sampler", type=str, default="path",
                   help="sampler type")
    p.add_argument("--clamp", type=float, default=10.0, help="clamp image")
    p.add_argument("--nocaustics", type=_bool, default=False, help="disable caustics")
    p.add_argument("--batch", type=int, default=1, help="run samples in batches")
    p.add_argument("--bvhstacksize", type=int, default=128, help="max depth of bvh exploration")
    # GPU-side extras (not in the reference)
    p.add_argument("--gpu-seed", dest="gpu_seed", type=int, default=0,
                   help="seed of the counter-based RNG")
    p.add_argument("--gpu-traversal", dest="gpu_traversal", type=str, default="wide",
                   choices=["wide", "reference"],
                   help="wide = quantised 8-wide BVH (fast); reference = the host-built binary "
                        "BVH walked in the reference's own order (parity mode)")
    p.add_argument("--gpu-integrator", dest="gpu_integrator", type=str, default="wavefront",
                   choices=["wavefront", "megakernel"],
                   help="wavefront = staged kernels over compacted, material-sorted queues (default); "
                        "megakernel = one thread per pixel running the whole trace_sample")
    p.add_argument("--gpu-devices", dest="gpu_devices", type=str, default="",
                   help="comma-separated CUDA device ids (or 'all') to shard the sample axis over, inside one "
                        "process (jt_group: scene replicated, fused peer-to-peer merge); empty = one device")
    p.add_argument("--gpu-bvh-cache", dest="gpu_bvh_cache", type=str, default="",
                   help="directory for finished wide BVHs (keyed by a hash of the scene geometry); empty = no cache")
    p.add_argument("--gpu-device-lights", dest="gpu_device_lights", type=_bool, default=False,
                   help="build the light CDFs on the GPU (jt_lights_create: bit-identical to make_trace_lights)")
    p.add_argument("--gpu-env-importance", dest="gpu_env_importance", type=_bool, default=False,
                   help="with --gpu-device-lights: weight environment texels by max(R, G, B) instead of the reference's "
                        "max(R, G, B, A = 1) (quirk Q8); unbiased, but no longer the reference's sample set")
    p.add_argument("--gpu-native-host", dest="gpu_native_host", type=_bool, default=False,
                   help="load the scene, build the BVH and the light CDFs inside the library (jt_host_scene_*) "
                        "instead of with the Python mirror; same bytes either way")
    return p


@dataclasses.dataclass
class Params:  # src/cli.jl:90-108
    scene: str = ""
    output: str = "tests/test_scene.png"
    camera: Any = ""
    addsky: bool = False
    envname: str = ""
    resolution: int = 1280
    samples: int = 512
    bounces: int = 8
    denoise: bool = False
    noparallel: bool = False
    highqualitybvh: bool = False
    envhidden: bool = False
    tentfilter: bool = False
    sampler: int = 1  # 1 = path, 2 = naive (index into SAMPLER_TYPES, 1-based)
    clamp: int = 10
    nocaustics: bool = False
    batch: int = 1
    bvhstacksize: int = 128
    gpu_seed: int = 0
    gpu_traversal: str = "wide"
    gpu_integrator: str = "wavefront"
    gpu_devices: str = ""
    gpu_native_host: bool = False
    gpu_device_lights: bool = False
    gpu_env_importance: bool = False
    gpu_bvh_cache: str = ""

    @staticmethod
    def from_args(ns: argparse.Namespace) -> "Params":
        sampler = SAMPLER_TYPES.index(ns.sampler) + 1 if ns.sampler in SAMPLER_TYPES else 1
        clamp = ns.clamp
        if float(clamp) != int(clamp):
            # Params.clamp::Int receives a Float32 (src/cli.jl:70-73,105): InexactError
            raise ValueError(f"InexactError: Int64({clamp}) -- the reference's Params.clamp is an Int")
        d = {f.name: getattr(ns, f.name) for f in dataclasses.fields(Params)
             if f.name not in ("sampler", "clamp")}
        return Params(sampler=sampler, clamp=int(clamp), **d)


def parse_cli_args(args: Union[str, Sequence[str]]) -> Optional[Params]:
    """src/cli.jl:140-147. Accepts the reference's single-string form too (`main(::String)`)."""
    if isinstance(args, str):
        args = shlex.split(args)
    ns = make_parser().parse_args(list(args))
    return Params.from_args(ns)
