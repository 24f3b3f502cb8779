"""Multi-GPU plumbing for the trace path (no reference counterpart -- RadeonRays is single GPU).

Rays are independent units, so a batch shards into contiguous slices (keeps primary-ray coherence), one per rank; every rank
needs the whole BVH.  A BLAS holds node INDICES, not pointers, and a scene buffer describes itself, so the bytes built on one rank
are valid on every other: broadcast them once (NCCL over NVLink on the GPU box, gloo in the CPU tests) instead of rebuilding; a
scene additionally gets its instance records re-pointed at the local copy of the geometry (rrCudaCmdRebindSceneGeometry).

Hits come back in one of two ways:
  * PeerHitBuffer (the product path on a GPU box): the destination buffer lives on the root rank and is mapped into every other
    rank's address space (rrCudaExportDeviceMemory / rrCudaImportDeviceMemory = CUDA IPC + NVLink peer access).  Rank r passes
    `buffer.slice_ptr(r)` as the `hits` argument of rrCmdIntersect and the traversal kernels store each hit straight into the
    root's memory while they trace: compute and gather are one kernel, there is no staging buffer and no separate collective.
  * gather_hits: trace into a local buffer, then one NCCL / gloo collective -- the comparison arm of bench.py and the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(count, rank, world, granule=64):
    """Contiguous [begin, end) slice of `count` rays for `rank`; slice boundaries are multiples of `granule` (a packet's worth of
    rays, rr_trace.cu k_trace_packet) so that no warp straddles two ranks' data."""
    per = -(-count // world)
    per = -(-per // granule) * granule
    begin = min(count, rank * per)
    return begin, min(count, begin + per)


def broadcast_bytes(tensor, src=0):
    """Broadcast a built BLAS / scene (uint8 tensor, device or host) from `src` to every rank, in place."""
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


class PeerHitBuffer:
    """`total_bytes` of device memory on rank `root`, mapped into every rank.  ptr(byte_offset) is an RRDevicePtr usable as the
    `hits` argument of rrCmdIntersect on the calling rank (local memory on the root, NVLink peer memory elsewhere)."""

    def __init__(self, ctx, total_bytes, root=0):
        self.ctx, self.root, self.total = ctx, root, int(total_bytes)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self._ptrs = []
        if self.rank == root:
            self.base = ctx.allocate(self.total)            # cudaMalloc inside the library: a whole allocation, exportable
            blob = [ctx.export_memory(self.base)]
        else:
            self.base, blob = None, [None]
        if self.world > 1:
            dist.broadcast_object_list(blob, src=root)
        if self.rank != root:
            handle, offset = blob[0]
            self.base = ctx.import_memory(handle, offset)
        self.address = ctx.raw_address(self.base)

    def ptr(self, byte_offset=0):
        p = self.ctx.device_ptr(self.address, int(byte_offset))
        self._ptrs.append(p)
        return p

    def read(self, dtype=np.uint8):
        """Root only: the buffer's current contents as a numpy array (rrMapDevicePtr / rrUnmapDevicePtr)."""
        assert self.rank == self.root
        arr, m = self.ctx.map(self.base, self.total, np.uint8)
        out = np.array(arr, copy=True).view(dtype)
        self.ctx.unmap(self.base, m)
        return out

    def close(self):
        for p in self._ptrs:
            self.ctx.release_ptr(p)
        self._ptrs = []
        if self.base is not None:
            self.ctx.release_ptr(self.base)
            self.base = None


def build_scene_sharded(engine, meshes, instance_mesh, transforms, build_flags=0):
    """Multi-mesh scene on N ranks ("BLAS i on GPU i mod G, then all-gather of BLAS bytes and TLAS build
    everywhere"): rank r builds the geometries i with i % world == r through rrCmdBuildGeometry, every rank then receives the bytes
    of every geometry from its builder (a BLAS holds indices, not pointers, so the bytes are valid anywhere), and builds the
    scene over its local copies with rrCmdBuildScene (a TLAS build is a few microseconds; a scene built elsewhere would instead be
    broadcast and re-pointed with rrCudaCmdRebindSceneGeometry).  `meshes`: list of (positions, indices) present on every rank.
    Returns (geometries, scene) as radeonrays_sdk_b200.host objects; with one rank it is a plain local build."""
    from .host import Geometry, _dev_bytes
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    ctx = engine.ctx
    geoms = []
    for i, (pos, idx) in enumerate(meshes):
        if i % world == rank:
            g = engine.build_geometry(pos, idx, build_flags=build_flags)
        else:   # a receive-only geometry: the result buffer of the right size, no build
            g = Geometry()
            g.engine, g.triangle_count = engine, int(idx.shape[0])
            probe = ctx.geometry_input(ctx.device_ptr(256), int(pos.shape[0]), 12, ctx.device_ptr(256), int(idx.shape[0]))
            from . import api
            req = ctx.geometry_requirements(probe, api.RRBuildOptions(build_flags, None))
            g.d_nodes = _dev_bytes(req.result_buffer_size, engine.device)
            g.p_nodes = ctx.tensor_ptr(g.d_nodes)
        geoms.append(g)
    torch.cuda.synchronize(engine.device)
    if world > 1:
        on_gpu = dist.get_backend() == "nccl"
        for i, g in enumerate(geoms):
            if on_gpu:
                dist.broadcast(g.d_nodes, src=i % world)
            else:   # gloo (tests): stage through the host
                host = g.d_nodes.cpu()
                dist.broadcast(host, src=i % world)
                g.d_nodes.copy_(host)
    scene = engine.build_scene(geoms, list(instance_mesh), transforms)
    return geoms, scene
