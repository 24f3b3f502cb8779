"""Multi-GPU plumbing for the trace path (no reference counterpart -- RadeonRays is single GPU).

Rays are independent units, so a batch shards into contiguous slices (keeps primary-ray coherence), one per rank;
every rank needs the whole BVH.  A BLAS holds node INDICES, not pointers, so the bytes built on one rank are valid
on every other: broadcast them once (NCCL over NVLink on the GPU box, gloo in the CPU tests) instead of rebuilding.
Hits are gathered back with one all_gather.  There is no collective inside the traversal itself.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(count, rank, world, granule=32):
    """Contiguous [begin, end) slice of `count` rays for `rank`; slice boundaries are multiples of `granule`
    (a warp's worth of rays) so that no warp straddles two ranks' data."""
    per = -(-count // world)
    per = -(-per // granule) * granule
    begin = min(count, rank * per)
    return begin, min(count, begin + per)


def broadcast_bytes(tensor, src=0):
    """Broadcast a built BLAS (uint8 tensor, device or host) from `src` to every rank, in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src=src)
    return tensor


def gather_hits(local_hits, count, item_bytes=16):
    """all_gather per-rank hit slices (uint8 tensors of shard_range sizes) into one [count*item_bytes] tensor."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_hits
    world = dist.get_world_size()
    b0, e0 = shard_range(count, 0, world)
    per = (e0 - b0) * item_bytes
    padded = torch.zeros(per, dtype=torch.uint8, device=local_hits.device)
    padded[: local_hits.numel()] = local_hits
    out = torch.empty(per * world, dtype=torch.uint8, device=local_hits.device)
    dist.all_gather_into_tensor(out, padded)
    return out[: count * item_bytes]


def trace_sharded(trace_fn, rays, item_bytes=16):
    """Strong-scaling trace of one batch: rank r traces rays[shard_range(r)] with `trace_fn(ray_slice) -> uint8 tensor`
    and every rank receives all hits.  `rays` is a numpy structured array (RAY_DTYPE) present on every rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    b, e = shard_range(rays.shape[0], rank, world)
    local = trace_fn(rays[b:e])
    return gather_hits(local, rays.shape[0], item_bytes)
