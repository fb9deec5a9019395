#!/usr/bin/env python
"""bench.py -- throughput of the per-pixel render loop on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (`trace_samples`) over one batch of `--spp-per-step` samples
for every pixel of the workload image. The workload is BASELINE.json's headline configuration,
classroom / path sampler / 1280 px (C4), from the packed reference scene (assets/scenes). One process
per GPU; the scene is replicated, global sample indices are sharded across ranks (no data-path
collective while rendering), and the per-rank SUM buffers (RGBA, albedo, normal, hits) are merged on
rank 0 with NCCL reduces over NVLink, divided by the sample count and downloaded.

Legs (all in one JSON line, rank 0):
  value   device-resident throughput: K sharded steps + the end-of-job merge and download, max over ranks
  e2e     every step: reset, render this rank's share of the step's global samples, reduce all four buffers to
          rank 0, rank 0 divides and downloads the merged TraceState to host arrays (pinned staging)
  multi_gpu_check   rank 0 re-renders the global samples of the last e2e step on ONE GPU and compares
  strong  fixed total samples per step (--strong-spp, default 512) split over the ranks, merged and downloaded
  in_library_group  rank 0 ALONE drives all N GPUs through jt_group (one host thread, fused P2P merge kernel);
          the other ranks wait on a CPU (gloo) barrier -- the path a Julia host would use

See DESIGN.md "Measurement" for every field."""
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
import tempfile
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
    ap.add_argument("--cpu-spp", type=int, default=16, help="samples per pixel of the bounded CPU-baseline sample (16 spp of the full image = ~12 s on 16 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-spp", type=int, default=512, help="total samples per step of the strong-scaling leg (N > 1)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the strong-scaling and in-library-group legs")
    return ap.parse_args()


def load_workload(args, product=True):
    """product=True: the host steps of the product (BVH via libjtrace_b200's jt_make_bvh, Python light builder).
    product=False (reference arm / cpu_baseline): only the packed-scene loader (pure Python); the oracle builds its
    OWN BVH and lights (oracle/orc_scene.h), so that arm stands on oracle/ alone and never maps libjtrace_b200.so."""
    jt = importlib.import_module("julia-raytracer_b200")
    scene = jt.load_scene(os.path.join(ROOT, "assets", "scenes", f"{args.scene}.jtscene"))
    if not product:
        return jt, scene, None, None
    bvh = importlib.import_module("julia-raytracer_b200.bvh")
    lights = importlib.import_module("julia-raytracer_b200.lights")
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
    o = orc.Oracle(scene)  # the oracle's own make_scene_bvh + make_trace_lights restatements (no product code)
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
    jt, scene, sbvh, lts = load_workload(args, product=False)
    spp = 1
    r = cpu_run(args, scene, sbvh, lts, spp, steps=args.steps, warmup=args.warmup)
    value = r["paths"] / r["seconds"] / 1e6
    rays = r["counters"]["scene_rays"] + r["counters"]["light_rays"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference scene (packed asset), counter-based RNG",
        "config": {"workload": workload_name(args, r["width"], r["height"])},
        "reference_arm_note": "julia is not installable here: the reference arm is the C++ oracle port of the same "
                              "algorithm (its own make_scene_bvh / make_trace_lights restatements, the reference's "
                              "traversal order), OpenMP over rows, all host threads; each step = 1 spp over the full "
                              "image (samples/s is spp-independent)",
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
    """SM clock + throttle reasons DURING the timed region (the profiling recipe's clocks line), every 200 ms.
    In-process NVML (pynvml) when available: a handful of cheap driver queries per sample. The nvidia-smi fallback is
    started BEFORE the warm-up steps (its start-up enumerates every GPU of the box and was seen to stall kernel
    submission of the first timed step by ~0.2 s); only samples taken between mark_begin() and stop() are used."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []   # (time, sm_mhz, max_mhz, set of reasons)
        self.proc = None
        self.nvml = None
        self.t_begin = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the CUDA device's UUID when torch knows it
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nvml = (pynvml, h)
            threading.Thread(target=self._poll_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv, h = self.nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.perf_counter(), sm, mx, {k for k, b in bits.items() if mask & b}))
            except Exception:
                pass
            self._stop.wait(0.2)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) > 7 and r[1].replace(".", "").isdigit():
                mx = float(r[2]) if r[2].replace(".", "").isdigit() else None
                self.rows.append((time.perf_counter(), float(r[1]), mx,
                                  {nm for k, nm in enumerate(names) if r[4 + k].lower().startswith("active")}))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None and self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        time.sleep(0.25)
        t_end = time.perf_counter()
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for r in self.rows if self.t_begin is None or self.t_begin <= r[0] <= t_end]
        sm = [r[1] for r in rows]
        mx = [r[2] for r in rows if r[2] is not None]
        reasons = set()
        for r in rows:
            reasons |= r[3]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "via": "nvml" if self.nvml else "nvidia-smi"}


class _DevBuf:
    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


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


def static_issue_stats(scene, sampler):
    """Issue-slot utilisation of the dominant kernel from the committed ncu capture of this workload
    (profiles/r02/issue.json, made by tools/ncu_summary.py from a --set full capture): not measurable live."""
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", rnd, "issue.json")
        if os.path.exists(p):
            d = json.load(open(p)).get(f"{scene}_{sampler}")
            if d:
                return {**d, "source": f"profiles/{rnd}/issue.json"}
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
    cpu_group = None
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
            cpu_group = dist.new_group(backend="gloo")  # CPU-side barrier for the single-process group leg
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    jt, scene, sbvh, lts = load_workload(args)
    trace = importlib.import_module("julia-raytracer_b200.trace")
    libmod = importlib.import_module("julia-raytracer_b200._lib")
    # N1: the first upload builds the wide BVH and stores it in the cache directory; for scenes whose build takes long
    # (ecosys: 16.8 M flattened records) the upload is repeated once to time the cached path as well
    cache_dir = os.environ.get("JT_BVH_CACHE_DIR") or os.path.join(tempfile.gettempdir(), "jtrace_b200_bvh_cache")
    try:
        os.makedirs(cache_dir, exist_ok=True)
        trace.set_bvh_cache_dir(cache_dir)
    except OSError:
        trace.set_bvh_cache_dir(None)  # nowhere to cache: every upload builds
    t0 = time.perf_counter()
    dscene = trace.DeviceScene(scene, sbvh, lts, local_rank)
    upload_s = time.perf_counter() - t0
    stats = dscene.stats()
    upload = {"seconds": upload_s, "device_bytes": stats["total_device_bytes"],
              "wide_bvh_from_cache": bool(stats["wide_bvh_from_cache"]),
              "includes": "host-side staging (wide-BVH build, or its load from the cache directory) + cudaMemcpy of every array"}
    if not stats["wide_bvh_from_cache"] and upload_s > 1.0:
        dscene.close()
        t0 = time.perf_counter()
        dscene = trace.DeviceScene(scene, sbvh, lts, local_rank)
        upload["seconds_cold"] = upload_s
        upload["seconds"] = time.perf_counter() - t0
        stats = dscene.stats()
        upload["wide_bvh_from_cache"] = bool(stats["wide_bvh_from_cache"])
    spp = args.spp_per_step
    total_steps = args.warmup + args.steps
    params = jt.Params(scene=args.scene, resolution=args.resolution, samples=1 << 30, batch=spp,
                       sampler=1 if args.sampler == "path" else 2, camera=jt.find_camera(scene, ""),
                       gpu_traversal=args.traversal, gpu_integrator=args.integrator)
    # sum mode: what the cross-GPU reduce adds up; rank r takes global sample indices
    # [ (step*world + r)*spp, +spp ): disjoint counter-RNG streams by construction
    state = trace.make_trace_state(dscene, params, accumulate=1)
    w, h = state.width, state.height
    npix = w * h
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    shard = importlib.import_module("julia-raytracer_b200.shard")

    def barrier():
        dscene.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def merge_on_rank0(st, n_samples):
        """The data-path collective: sum-reduce ALL FOUR accumulation buffers (RGBA, albedo, normal: float4 per pixel;
        hits: int32) onto rank 0 over NCCL/NVLink, then tell rank 0's state how many samples the sums hold.
        jt_state_device_buffers has flushed and synchronised the library's streams; the reduces run on torch's current
        stream and are synchronised before the library touches the buffers again (include/jtrace_b200.h)."""
        if world > 1:
            bufs = st.device_buffers()
            for key, dt, per in (("image", "<f4", 4), ("albedo", "<f4", 4), ("normal", "<f4", 4), ("hits", "<i4", 1)):
                t = torch.as_tensor(_DevBuf(bufs[key], bufs["count"] * per, dt), device="cuda")
                shard.reduce_sums(t, dst=0)
            torch.cuda.synchronize()
        st.set_samples(n_samples)

    # ================================ leg 1: value (inputs resident, one merge + download at job end) ===============
    clocks = ClockSampler(local_rank)
    clocks.start()  # before the warm-up: its start-up cost stays out of the timed region
    for k in range(args.warmup):
        b, e = shard.step_range(k, world, rank, spp)
        trace.trace_sample_range(state, dscene, params, b, e)
        dscene.synchronize()
        flush.zero_()
    barrier()
    dscene.counters(reset=True)
    dscene.elapsed_ms()
    clocks.mark_begin()
    kernel_ms = 0.0
    t0 = time.perf_counter()
    for k in range(args.warmup, total_steps):
        b, e = shard.step_range(k, world, rank, spp)
        trace.trace_sample_range(state, dscene, params, b, e)
        dscene.synchronize()
        kernel_ms += dscene.elapsed_ms()  # CUDA events on the library's launch stream
        flush.zero_()                     # L2 flush between timed iterations
    merge_on_rank0(state, total_steps * world * spp)  # the single end-of-job merge: 52 B per pixel per rank
    if rank == 0:
        state.sync()                      # divide by N + device -> host of the merged TraceState (48 B per pixel)
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

    # ================================ leg 2: e2e (host buffers; every step ends with a merged image on the host) =====
    # Through the public host API: trace_sample_range + (N > 1: reduce of all four buffers) + TraceState.sync().
    # Every step renders world * spp NEW global samples (rank r its step_range share) into freshly reset sum buffers,
    # so what reaches rank 0's host arrays each step is ONE image of world * spp samples.
    e2e_state = trace.make_trace_state(dscene, params, accumulate=1)
    e2e_steps = max(2, min(args.steps, 4))

    def e2e_step(k, st, per_rank_spp):
        st.reset()
        b, e = shard.step_range(k, world, rank, per_rank_spp)
        trace.trace_sample_range(st, dscene, params, b, e)
        merge_on_rank0(st, world * per_rank_spp)
        if rank == 0:
            st.sync()  # device -> host: image 16 B + albedo 12 B + normal 12 B + hits 8 B per pixel

    e2e_step(0, e2e_state, spp)  # warm-up (allocates the staging buffers)
    barrier()
    t0 = time.perf_counter()
    for k in range(1, e2e_steps + 1):
        e2e_step(k, e2e_state, spp)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * npix * spp / float(te[0]) / 1e6

    # ---- multi_gpu_check: the merged image of the LAST e2e step vs one GPU rendering the same global samples -------
    multi_gpu_check = None
    if world > 1 and rank == 0:
        merged = {k: getattr(e2e_state, k).copy() for k in ("image", "albedo", "normal", "hits")}
        chk = trace.make_trace_state(dscene, params, accumulate=1)
        trace.trace_sample_range(chk, dscene, params, e2e_steps * world * spp, (e2e_steps + 1) * world * spp)
        chk.sync()
        scale = np.maximum(np.abs(chk.image), 1e-3)
        rel = float((np.abs(merged["image"] - chk.image) / scale).max())
        multi_gpu_check = {
            "max_rel_err": rel, "max_abs_err": float(np.abs(merged["image"] - chk.image).max()),
            "albedo_max_abs_err": float(np.abs(merged["albedo"] - chk.albedo).max()),
            "normal_max_abs_err": float(np.abs(merged["normal"] - chk.normal).max()),
            "hits_equal": bool(np.array_equal(merged["hits"], chk.hits)),
            "samples": world * spp, "tolerance": 1e-4,
            "what": f"merged {world}-rank image of global samples [{e2e_steps * world * spp}, {(e2e_steps + 1) * world * spp}) "
                    "on rank 0's host vs the same samples rendered by one GPU (differences = float addition order)"}
        multi_gpu_check["ok"] = bool(rel <= 1e-4 and multi_gpu_check["hits_equal"]
                                     and multi_gpu_check["albedo_max_abs_err"] <= 1e-4
                                     and multi_gpu_check["normal_max_abs_err"] <= 1e-4)
        chk.close()
    if world > 1:
        dist.barrier()

    # ================================ leg 3: strong scaling (fixed total samples per step) ============================
    strong = None
    if world > 1 and not args.no_extra_legs and args.strong_spp % world == 0:
        per = args.strong_spp // world
        e2e_step(0, e2e_state, per)
        barrier()
        t0 = time.perf_counter()
        s_steps = max(2, min(args.steps, 8))
        for k in range(1, s_steps + 1):
            e2e_step(k, e2e_state, per)
        barrier()
        ts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"total_spp_per_step": args.strong_spp, "spp_per_rank": per, "steps": s_steps,
                  "ms_per_step": float(ts[0]) / s_steps * 1e3,
                  "value": s_steps * npix * args.strong_spp / float(ts[0]) / 1e6, "unit": UNIT,
                  "includes": "reset + render + NCCL reduce of 4 buffers + download on rank 0, every step (e2e-style)",
                  "note": "the reference's default --samples 512 split over the ranks: each rank's chunk ends in its own "
                          "drain phase, so this regime is drain-heavier than the weak-scaling legs"}
    e2e_state.close()

    # ================================ leg 4: in-library group (one process, N devices) ==============================
    group_leg = None
    if world > 1 and not args.no_extra_legs:
        if rank == 0:
            try:
                g0 = time.perf_counter()
                group = trace.DeviceGroup(scene, sbvh, lts, list(range(world)))
                g_upload = time.perf_counter() - g0
                gstate = trace.make_trace_state(group, params)
                gspp = world * spp

                def gstep(k, download):
                    if download:
                        gstate.reset()
                    trace.trace_sample_range(gstate, group, params, k * gspp, (k + 1) * gspp)
                    if download:
                        gstate.sync()
                    else:
                        group.synchronize()

                gstep(0, True)
                group.counters(reset=True)
                g_steps = max(2, min(args.steps, 4))
                t0 = time.perf_counter()
                for k in range(1, g_steps + 1):
                    gstep(k, False)
                gstate.sync()
                g_dev = time.perf_counter() - t0
                gc_ = group.counters()
                t0 = time.perf_counter()
                for k in range(g_steps + 1, 2 * g_steps + 1):
                    gstep(k, True)
                g_e2e = time.perf_counter() - t0
                # check the last merged image against one GPU
                chk = trace.make_trace_state(dscene, params, accumulate=1)
                trace.trace_sample_range(chk, dscene, params, 2 * g_steps * gspp, (2 * g_steps + 1) * gspp)
                chk.sync()
                rel = float((np.abs(gstate.image - chk.image) / np.maximum(np.abs(chk.image), 1e-3)).max())
                gst = group.stats()
                group_leg = {
                    "value": gc_["camera_paths"] / g_dev / 1e6, "unit": UNIT, "steps": g_steps,
                    "e2e": {"value": g_steps * npix * gspp / g_e2e / 1e6, "unit": UNIT,
                            "d2h_bytes_per_step": npix * 48, "nvlink_bytes_per_step": npix * 52 * gst["peer_members"]},
                    "multi_gpu_check": {"max_rel_err": rel, "hits_equal": bool(np.array_equal(gstate.hits, chk.hits)),
                                        "ok": bool(rel <= 1e-4 and np.array_equal(gstate.hits, chk.hits))},
                    "members": gst["members"], "peer_members": gst["peer_members"], "staged_members": gst["staged_members"],
                    "stage_seconds": gst["stage_seconds"], "upload_seconds": gst["upload_seconds"],
                    "create_seconds": g_upload, "reduce_bytes_remote": gst["reduce_bytes_remote"],
                    "what": "ONE process / ONE host thread driving all GPUs through jt_group_* (worker thread per "
                            "device inside the library, fused peer-to-peer reduce + finalize kernel on device 0, no "
                            "NCCL); the other torchrun ranks idle on a CPU barrier meanwhile"}
                chk.close()
                gstate.close()
                group.close()
            except Exception as ex:  # reported, never fatal for the contract line
                group_leg = {"error": repr(ex)}
        dist.barrier(group=cpu_group)  # CPU wait: the idle ranks launch nothing while rank 0 drives their GPUs

    if rank == 0:
        value = paths / wall_s / 1e6
        rays = scene_rays + light_rays
        peak, peak_src = peak_hbm()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "reference scene (packed asset assets/scenes, missing-asset rule of SURVEY 8d), counter-based RNG",
            "config": {"workload": workload_name(args, w, h)},
            "config_detail": {
                "traversal": args.traversal, "integrator": args.integrator,
                "sharding": (f"global sample indices strided over {world} rank(s); scene replicated; NCCL reduce of the "
                             f"four sum buffers (RGBA, albedo, normal, hits) onto rank 0 + download, inside the timed region")
                if world > 1 else "single GPU",
                "l2": "256 MB buffer written between timed iterations (L2 flush); the scene itself "
                      f"({stats['total_device_bytes'] / 1e6:.0f} MB) is L2-resident by design"},
            "mrays_per_s": rays / wall_s / 1e6,
            "rays_per_sample": rays / max(paths, 1),
            "scene_rays": scene_rays, "light_probe_rays": light_rays,
            "kernel_ms_per_step": kernel_ms / args.steps,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(trace.A.jt_params),
                    "d2h_bytes_per_step": npix * 48, "steps": e2e_steps,
                    "reduce_bytes_per_step": npix * 52 * (world - 1),
                    "note": "per step: reset + trace_sample_range (this rank's share of world*spp new global samples)"
                            + (" + NCCL reduce of RGBA/albedo/normal/hits onto rank 0" if world > 1 else "")
                            + " + TraceState.sync() (merged image, albedo, normal, hits to host arrays); the scene is "
                              "uploaded once per render (see scene_upload), like the reference loads it once"},
            "scene_upload": upload,
            "gpu_launches": int(launches),
            "scheduling": {"stolen_samples": int(c.get("stolen_samples", 0)), "resumed_rays": int(c.get("resumed_rays", 0)),
                           "camera_paths": int(c["camera_paths"]), "scene_rays": int(c["scene_rays"]),
                           "what": "rank 0, timed steps: samples traced by a path slot that started on another pixel "
                                   "(work stealing) and closest-hit queries parked in a launch tail and resumed by the "
                                   "next launch; neither changes a result bit (DESIGN.md 1.5)"},
            "scene_stats": stats,
        }
        if multi_gpu_check is not None:
            line["multi_gpu_check"] = multi_gpu_check
        if strong is not None:
            line["strong_scaling"] = strong
        if group_leg is not None:
            line["in_library_group"] = group_leg
        per_sample = frozen_algorithmic_bytes(args.scene, args.sampler)
        if world == 1 and not args.no_cpu_baseline:
            _, cscene, _, _ = load_workload(args, product=False)
            r = cpu_run(args, cscene, None, None, args.cpu_spp)
            line["cpu_baseline"] = {
                "value": r["paths"] / r["seconds"] / 1e6, "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": f"{args.cpu_spp} spp over the full {r['width']}x{r['height']} image "
                          f"({r['paths']} camera paths, {r['seconds']:.1f} s); C++ oracle port, OpenMP",
                "mrays_per_s": (r["counters"]["scene_rays"] + r["counters"]["light_rays"]) / r["seconds"] / 1e6}
        # ---- roofline of the dominant kernel (k_wf_extend_persist = intersect_scene_bvh) -----------------------------
        # `achieved` / `frac` are PHYSICAL: the bytes this kernel itself requests (80 B wide nodes + 48 B triangle
        # records per scene ray, counted on the CPU emulation of the same traversal and frozen in
        # profiles/algorithmic_bytes.json) x the rays its launches retired / the CUDA-event duration of those launches,
        # against the measured HBM peak -- never above 1 by construction. The figure SURVEY 8d prescribes (bytes of the
        # REFERENCE algorithm for the same rays) is kept under `work_normalised`; it can exceed 1 because the wide BVH
        # does not move those bytes. What really bounds the kernel is instruction issue: `issue`.
        if per_sample is not None:
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(f"{args.scene}_{args.sampler}_{args.traversal}")
            step_bytes = per_sample["bytes_per_sample"] * (paths / world) / args.steps
            step_gbs = step_bytes / ((kernel_ms / args.steps) / 1e3) / 1e9
            roof = None
            if args.integrator == "wavefront" and c.get("extend_launches", 0) > 0:
                ext_s = c["extend_us"] / 1e6
                alg = per_sample["scene_bytes_per_scene_ray"] * c["scene_rays"]  # rank 0's launches
                wn = {"achieved": alg / ext_s / 1e9, "frac": alg / ext_s / 1e9 / peak, "unit": "GB/s",
                      "algorithmic_bytes_per_scene_ray": per_sample["scene_bytes_per_scene_ray"],
                      "algorithmic_bytes_per_launch": alg / c["extend_launches"],
                      "note": "SURVEY 8d normalisation: bytes of the REFERENCE algorithm (binary BVH, reference order) for "
                              "the rays retired; > 1 is possible and means the wide BVH avoided those bytes"}
                kernel = "k_wf_extend_persist" if args.traversal == "wide" else "k_wf_extend<reference>"
                if "wide_bytes_per_scene_ray" in per_sample and args.traversal == "wide":
                    own = per_sample["wide_bytes_per_scene_ray"] * c["scene_rays"]
                else:
                    own = alg
                achieved = own / ext_s / 1e9
                roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": peak_src, "kernel": kernel,
                        "bytes_per_scene_ray": own / max(c["scene_rays"], 1),
                        "bytes_per_launch": own / c["extend_launches"],
                        "launches_timed": c["extend_launches"], "avg_launch_ms": ext_s * 1e3 / c["extend_launches"],
                        "kernel_share_of_step": ext_s * 1e3 / kernel_ms,
                        "kernel_share_note": "sum of per-launch CUDA-event durations over ALL pipelines / step time; "
                                             "pipelines overlap, so shares of different kernels can add up to > 1",
                        "work_normalised": wn}
                # L2 roofline: everything this kernel walks is L2-resident -> measure the L2 read peak on this box
                try:
                    gbs = C.c_float()
                    l2 = {}
                    for mb in (24, 48, 96):
                        libmod.check(libmod.lib().jt_probe_read_bandwidth(local_rank, mb << 20, 40, C.byref(gbs)))
                        l2[f"{mb}MB"] = float(gbs.value)
                    libmod.check(libmod.lib().jt_probe_read_bandwidth(local_rank, 2048 << 20, 2, C.byref(gbs)))
                    l2_peak = max(l2.values())
                    roof["l2"] = {"read_peak_gbs": l2_peak, "by_working_set": l2, "hbm_read_gbs_2GB": float(gbs.value),
                                  "frac_of_l2_peak": achieved / l2_peak,
                                  "how": "jt_probe_read_bandwidth: one resident wave of 128-bit ld.global.cg over a buffer "
                                         "that fits L2, 40 passes, best of 5, CUDA events"}
                except Exception as ex:
                    roof["l2"] = {"error": repr(ex)}
                issue = static_issue_stats(args.scene, args.sampler)
                if issue:
                    roof["issue"] = issue
            else:
                roof = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": min(step_gbs / peak, 1.0),
                        "traffic": traffic, "peak_source": peak_src, "kernel": "k_trace_mega"}
            roof["whole_step_work_normalised"] = {"achieved": step_gbs, "frac": step_gbs / peak,
                                                  "algorithmic_bytes_per_sample": per_sample["bytes_per_sample"]}
            roof["algorithmic_bytes_source"] = "profiles/algorithmic_bytes.json (oracle + emulation counters, SURVEY 8d formula)"
            line["roofline"] = roof
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
