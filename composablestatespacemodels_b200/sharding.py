"""Host-side plumbing of a particle cloud sharded over several GPUs (one process per GPU).

There is no reference analogue (a reference filter is one JVM thread); the
device side is described in include/cssm.h (cssm_filter_create_sharded).  The only thing the host
has to do is to carry each rank's connection blob to every other rank once, at creation; after
that the ranks talk to each other from inside the kernels, over NVLink.  `torch.distributed` is
used for that one all-gather (any transport would do: the blobs are plain bytes).
"""
import numpy as np

from . import _abi


def local_count(n_global, world):
    """Particles per rank; the cloud must split evenly (rank r owns slots [r*n, (r+1)*n))."""
    n_global, world = int(n_global), int(world)
    if world < 1 or world > _abi.MAX_RANKS:
        raise ValueError(f"world must be in [1, {_abi.MAX_RANKS}]")
    if n_global <= 0 or n_global % world != 0:
        raise ValueError("the global particle count must be a positive multiple of the number of ranks")
    if n_global > 2 ** 31 - 1:
        raise ValueError("at most 2^31-1 particles (ancestor indices are Int, as in the reference)")
    return n_global // world


def slot_range(rank, world, n_global):
    n = local_count(n_global, world)
    return rank * n, (rank + 1) * n


def all_gather_blobs(blob, group=None):
    """Every rank's blob, in rank order, on every rank (works with the gloo and nccl backends)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


def create_sharded(mod, resample_kind, n_global, dtype=_abi.F32, device=0, seed=0, stream_id=0, group=None):
    """This process's shard of ONE filter of `n_global` particles, connected to its peers.
    Every rank must call this (and every later method of the handle) in the same order."""
    import torch.distributed as dist
    from .filter import GpuFilterHandle
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    h = GpuFilterHandle(mod, resample_kind, local_count(n_global, world), dtype, device, seed, stream_id, rank=rank, world=world)
    if world > 1:
        h.shard_connect(all_gather_blobs(h.shard_export(), group))
        dist.barrier(group)
    return h
