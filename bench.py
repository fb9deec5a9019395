#!/usr/bin/env python
"""bench.py -- throughput of the per-pixel render loop on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (`trace_samples`) over one batch of `--spp-per-step` samples
for every pixel of the workload image. The workload is BASELINE.json's headline configuration,
classroom / path sampler / 1280 px (C4), from the packed reference scene (assets/scenes). One process
per GPU; the scene is replicated, global sample indices are sharded across ranks (no data-path
collective), and the per-rank sum buffers are merged with ONE NCCL reduce at the end of the job.

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field."""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "samples/s per image (camera paths traced per second), classroom path-traced 1280px"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="classroom")
    ap.add_argument("--sampler", default="path", choices=["path", "naive"])
    ap.add_argument("--resolution", type=int, default=1280)
    ap.add_argument("--spp-per-step", type=int, default=512)
    ap.add_argument("--traversal", default="wide", choices=["wide", "reference"])
    ap.add_argument("--integrator", default="wavefront", choices=["wavefront", "megakernel"])
    ap.add_argument("--cpu-spp", type=int, default=2, help="samples per pixel of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_workload(args):
    jt = importlib.import_module("julia-raytracer_b200")
    bvh = importlib.import_module("julia-raytracer_b200.bvh")
    lights = importlib.import_module("julia-raytracer_b200.lights")
    scene = jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{args.scene}.jtscene"))
    sbvh = bvh.make_scene_bvh(scene)
    lts = lights.make_trace_lights(scene)
    return jt, scene, sbvh, lts


def workload_name(args, w, h):
    return (f"{args.scene}, {args.sampler} sampler, {w}x{h}, bounces 8, clamp 10, "
            f"{args.spp_per_step} spp per step (BASELINE config C4 image; spp-independent metric)")


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle port (test infrastructure) timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(args, scene, sbvh, lts, spp, steps=1, warmup=0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    # all host threads: torchrun exports OMP_NUM_THREADS=1 to every rank, which would cripple the CPU arm (libgomp
    # reads the variable when the oracle library is loaded, i.e. at the import below)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import orc  # the ONLY place bench.py touches oracle/: cpu_baseline and --impl reference
    o = orc.Oracle(scene, sbvh, lts)
    p = orc.make_params(resolution=args.resolution, samples=1 << 30, batch=spp,
                        sampler=1 if args.sampler == "path" else 2)
    w, h = o.make_state(p)
    cores = os.cpu_count() or 1  # passed explicitly: independent of OMP_NUM_THREADS and of who loaded libgomp first
    for _ in range(warmup):
        o.trace_samples(p, threads=cores)
    o.counters(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.trace_samples(p, threads=cores)
    dt = time.perf_counter() - t0
    c = o.counters()
    return dict(seconds=dt, paths=c["camera_paths"], counters=c, cores=cores, width=w, height=h,
                alg_bytes=orc.algorithmic_bytes(c))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    jt, scene, sbvh, lts = load_workload(args)
    spp = 1
    r = cpu_run(args, scene, sbvh, lts, spp, steps=args.steps, warmup=min(args.warmup, 1))
    value = r["paths"] / r["seconds"] / 1e6
    rays = r["counters"]["scene_rays"] + r["counters"]["light_rays"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": r["seconds"] / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference scene (packed asset), counter-based RNG",
        "config": {"workload": workload_name(args, r["width"], r["height"]),
                   "note": "julia is not installable here: the reference arm is the C++ oracle port of the "
                           "same algorithm on the same binary BVH in the reference's traversal order, OpenMP over "
                           "rows, all host threads; each step = 1 spp over the full image"},
        "mrays_per_s": rays / r["seconds"] / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{args.steps} x 1 spp over the full {r['width']}x{r['height']} image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class _DevBuf:
    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def frozen_algorithmic_bytes(scene, sampler):
    p = os.path.join(ROOT, "profiles", "algorithmic_bytes.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get(f"{scene}_{sampler}")
    return None


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libjtrace_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at first use; stdout must carry exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    jt, scene, sbvh, lts = load_workload(args)
    trace = importlib.import_module("julia-raytracer_b200.trace")
    t0 = time.perf_counter()
    dscene = trace.DeviceScene(scene, sbvh, lts, local_rank)
    upload_s = time.perf_counter() - t0
    stats = dscene.stats()
    spp = args.spp_per_step
    total_steps = args.warmup + args.steps
    params = jt.Params(scene=args.scene, resolution=args.resolution, samples=1 << 30, batch=spp,
                       sampler=1 if args.sampler == "path" else 2, camera=jt.find_camera(scene, ""),
                       gpu_traversal=args.traversal, gpu_integrator=args.integrator)
    # sum mode: what the cross-GPU reduce adds up; rank r takes global sample indices
    # [ (step*world + r)*spp, +spp ): disjoint counter-RNG streams by construction
    state = trace.make_trace_state(dscene, params, accumulate=1)
    w, h = state.width, state.height
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    shard = importlib.import_module("julia-raytracer_b200.shard")

    def step(k):
        begin, end = shard.step_range(k, world, rank, spp)
        trace.trace_sample_range(state, dscene, params, begin, end)

    def barrier():
        dscene.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for k in range(args.warmup):
        step(k)
        dscene.synchronize()
        flush.zero_()
    barrier()
    dscene.counters(reset=True)
    dscene.elapsed_ms()
    clocks = ClockSampler(local_rank)
    clocks.start()
    kernel_ms = 0.0
    t0 = time.perf_counter()
    for k in range(args.warmup, total_steps):
        step(k)
        dscene.synchronize()
        kernel_ms += dscene.elapsed_ms()  # CUDA events on the library's launch stream
        flush.zero_()                     # L2 flush between timed iterations
    bufs = state.device_buffers()
    if world > 1:  # the single end-of-job merge of the accumulation buffers (NCCL reduce over NVLink)
        img = torch.as_tensor(_DevBuf(bufs["image"], bufs["count"] * 4), device="cuda")
        shard.reduce_sums(img, dst=0)
    barrier()
    wall_s = time.perf_counter() - t0
    clk = clocks.stop()
    c = dscene.counters()

    # max over ranks of the timed region
    t = torch.tensor([wall_s, kernel_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([c["camera_paths"], c["scene_rays"], c["light_rays"], c["kernel_launches"]],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    wall_s, kernel_ms = float(t[0]), float(t[1])
    paths, scene_rays, light_rays, launches = (float(x) for x in cnt)

    # ---- e2e: the public API with host buffers -- trace_samples + sync() (D2H of the whole TraceState)
    e2e_state = trace.make_trace_state(dscene, params, accumulate=0)
    e2e_params = jt.Params(**{**params.__dict__})
    for _ in range(1):
        trace.trace_sample_range(e2e_state, dscene, e2e_params, 0, spp)
        e2e_state.sync()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 4))
    for k in range(e2e_steps):
        trace.trace_sample_range(e2e_state, dscene, e2e_params, (k + 1) * spp, (k + 2) * spp)
        e2e_state.sync()  # device -> host: image 16 B + albedo 12 B + normal 12 B + hits 8 B per pixel
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * w * h * spp / float(te[0]) / 1e6

    if rank == 0:
        value = paths / wall_s / 1e6
        rays = scene_rays + light_rays
        peak, peak_src = peak_hbm()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "reference scene (packed asset assets/scenes, missing-asset rule of SURVEY §8d), counter-based RNG",
            "config": {"workload": workload_name(args, w, h), "traversal": args.traversal, "integrator": args.integrator,
                       "sharding": f"global sample indices strided over {world} rank(s); scene replicated; "
                                   f"one NCCL reduce of the RGBA sum buffer at job end" if world > 1 else "single GPU",
                       "l2": "256 MB buffer written between timed iterations (L2 flush); the scene itself "
                             f"({stats['total_device_bytes'] / 1e6:.0f} MB) is L2-resident by design"},
            "mrays_per_s": rays / wall_s / 1e6,
            "rays_per_sample": rays / max(paths, 1),
            "scene_rays": scene_rays, "light_probe_rays": light_rays,
            "kernel_ms_per_step": kernel_ms / args.steps,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(trace.A.jt_params),
                    "d2h_bytes_per_step": w * h * 48,
                    "note": "trace_sample_range + TraceState.sync() per step; the scene is uploaded once per "
                            "render (see scene_upload), like the reference loads it once"},
            "scene_upload": {"seconds": upload_s, "device_bytes": stats["total_device_bytes"],
                             "includes": "host-side wide-BVH build + cudaMemcpy of every array"},
            "gpu_launches": int(launches),
            "scene_stats": stats,
        }
        # roofline of the dominant kernel (k_wf_extend_persist = intersect_scene_bvh): algorithmic bytes per launch
        # (frozen oracle figure per scene ray x rays per launch) / CUDA-event duration of that kernel's launches
        per_sample = frozen_algorithmic_bytes(args.scene, args.sampler)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_run(args, scene, sbvh, lts, args.cpu_spp)
            line["cpu_baseline"] = {
                "value": r["paths"] / r["seconds"] / 1e6, "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": f"{args.cpu_spp} spp over the full {r['width']}x{r['height']} image "
                          f"({r['paths']} camera paths, {r['seconds']:.1f} s); C++ oracle port, OpenMP",
                "mrays_per_s": (r["counters"]["scene_rays"] + r["counters"]["light_rays"]) / r["seconds"] / 1e6}
        if per_sample is not None:
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(f"{args.scene}_{args.sampler}_{args.traversal}")
            step_bytes = per_sample["bytes_per_sample"] * (paths / world) / args.steps
            step_gbs = step_bytes / ((kernel_ms / args.steps) / 1e3) / 1e9
            if args.integrator == "wavefront" and c.get("extend_launches", 0) > 0:
                ext_s = c["extend_us"] / 1e6
                alg = per_sample["scene_bytes_per_scene_ray"] * c["scene_rays"]  # rank 0's launches
                achieved = alg / ext_s / 1e9
                kernel = "k_wf_extend_persist" if args.traversal == "wide" else "k_wf_extend<reference>"
                # extend launches of the two image-half pipelines run on separate streams and overlap other kernels:
                # the sum of their per-launch durations is compared with the step time, it is not a wall-clock share
                extra = {"launches_timed": c["extend_launches"], "kernel_share_of_step": ext_s * 1e3 / kernel_ms,
                         "kernel_share_note": "sum of per-launch CUDA-event durations over ALL pipelines / step time; "
                                              "pipelines overlap, so shares of different kernels can add up to > 1",
                         "algorithmic_bytes_per_launch": alg / c["extend_launches"],
                         "avg_launch_ms": ext_s * 1e3 / c["extend_launches"],
                         "algorithmic_bytes_per_scene_ray": per_sample["scene_bytes_per_scene_ray"]}
                if "wide_bytes_per_scene_ray" in per_sample and args.traversal == "wide":
                    wb = per_sample["wide_bytes_per_scene_ray"] * c["scene_rays"] / ext_s / 1e9
                    extra["as_implemented"] = {
                        "bytes_per_scene_ray": per_sample["wide_bytes_per_scene_ray"], "achieved": wb, "unit": "GB/s",
                        "frac_of_hbm_peak": wb / peak,
                        "note": "bytes the wide-BVH kernel itself requests (80 B nodes, 48 B triangle records); served "
                                "by L2 -- the HBM peak is only a yardstick here"}
            else:
                achieved, kernel, extra = step_gbs, "k_trace_mega", {}
            line["roofline"] = {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": kernel, **extra,
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak,
                               "algorithmic_bytes_per_sample": per_sample["bytes_per_sample"]},
                "algorithmic_bytes_source": "profiles/algorithmic_bytes.json (oracle counters, SURVEY 8d formula)",
                "note": "work-normalised to the REFERENCE algorithm's node/primitive visits; every array the kernel "
                        "walks is L2-resident and the wide BVH visits ~7x fewer nodes, so DRAM traffic is a small "
                        "fraction of the algorithmic bytes by design"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
