"""The N > 1 path on CPU: world_size-2 gloo processes shard the global sample indices, render their share
into SUM buffers (with the oracle standing in for the device here -- allowed in tests/), reduce once, and
must reproduce the single-rank sample set. The same sharding code drives bench.py --gpus N."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc

shard = importlib.import_module("julia-raytracer_b200.shard")


def test_step_ranges_partition_the_sample_axis():
    for world in (1, 2, 4, 8):
        seen = []
        for step in range(5):
            for rank in range(world):
                b, e = shard.step_range(step, world, rank, 3)
                seen.extend(range(b, e))
        assert sorted(seen) == list(range(5 * world * 3))
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            for b, e in shard.split_samples(37, world, rank, chunk=4):
                seen.extend(range(b, e))
        assert sorted(seen) == list(range(37))


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(orc.ROOT, "tests"))
    sc = orc.jt.load_scene(os.path.join(orc.ROOT, "assets", "scenes", "cornellbox.jtscene"))
    o = orc.Oracle(sc)
    p = orc.make_params(resolution=48, samples=1 << 20, batch=1, sampler=1, accumulate=1)
    o.make_state(p)
    steps, spp = 2, 2
    for k in range(steps):
        b, e = shard.step_range(k, world, rank, spp)
        o.trace_range(p, b, e, threads=2)
    st = o.get_state()                      # sum mode: get_state divides by this rank's sample count
    sums = torch.from_numpy(st["image"] * np.float32(steps * spp))
    shard.reduce_sums(sums, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), sums.numpy() / np.float32(steps * spp * world))
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_matches_single_rank(tmp_path):
    world, port = 2, 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    merged = np.load(tmp_path / "merged.npy")
    sc = orc.jt.load_scene(os.path.join(orc.ROOT, "assets", "scenes", "cornellbox.jtscene"))
    o = orc.Oracle(sc)
    p = orc.make_params(resolution=48, samples=1 << 20, batch=1, sampler=1, accumulate=1)
    o.make_state(p)
    o.trace_range(p, 0, 8)                  # the same 8 global sample indices on one rank
    single = o.get_state()["image"]
    assert merged.shape == single.shape
    assert np.allclose(merged, single, rtol=1e-5, atol=1e-6)
