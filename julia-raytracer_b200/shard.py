"""Sharding of the render loop across GPUs: the scene is replicated, every rank
renders ALL pixels for a disjoint set of GLOBAL sample indices into sum buffers, and one reduce merges
them. RNG streams are keyed by the global sample index, so the union over ranks is exactly the sample
set a single GPU would have used (results agree up to the order of float additions)."""
from __future__ import annotations

from typing import List, Tuple


def step_range(step: int, world: int, rank: int, spp: int) -> Tuple[int, int]:
    """Global sample indices [begin, end) rank `rank` renders in step `step` (weak scaling: every rank
    gets `spp` samples per step; ranks interleave so any prefix of steps is a contiguous sample set)."""
    begin = (step * world + rank) * spp
    return begin, begin + spp


def split_samples(total: int, world: int, rank: int, chunk: int = 8) -> List[Tuple[int, int]]:
    """Strong-scaling split of samples [0, total): chunks of `chunk` samples dealt round-robin."""
    out = []
    k = 0
    begin = 0
    while begin < total:
        end = min(begin + chunk, total)
        if k % world == rank:
            out.append((begin, end))
        begin = end
        k += 1
    return out


def reduce_sums(tensor, dst: int = 0):
    """The single data-path collective: sum-reduce an accumulation buffer onto rank `dst` (NCCL on GPUs,
    gloo in the CPU tests). No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
    return tensor
