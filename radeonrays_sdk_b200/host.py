"""Host-side helpers that drive the rr* C ABI the way the reference's tests do (test/test_vk/basic_test.h):
query memory requirements -> allocate client buffers -> wrap them in RRDevicePtr -> record -> submit -> wait.

torch is used ONLY as the client-side device allocator / copy engine (the role Vulkan buffers play in the
reference tests); all BVH and ray work happens inside libradeonrays_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import api
from .workloads import HIT_DTYPE, NODE_DTYPE, RAY_DTYPE


def _dev_bytes(nbytes, device):
    # 256-byte aligned by the caching allocator; never zero-sized
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _upload(arr, device):
    a = np.ascontiguousarray(arr)
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(device)


class Engine:
    """One RRContext bound to torch's current CUDA stream on `device`."""

    def __init__(self, device=0):
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        # One stream for both the client-side torch copies and the library's kernels: make a real (non-default)
        # stream current on this device and hand its handle to rrCreateContextCuda, so that CUDA events recorded
        # through torch bracket the library's work (torch.cuda.Event only sees torch's current stream).
        self.torch_stream = torch.cuda.Stream(self.device)
        torch.cuda.set_stream(self.torch_stream)
        self.ctx = api.Context(device=device, cuda_stream=self.torch_stream.cuda_stream)
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count

    def close(self):
        torch.cuda.synchronize(self.device)
        self.ctx.destroy()

    # ---- geometry ----------------------------------------------------------------------------------------
    def build_geometry(self, positions, indices, build_flags=api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, vertex_stride=None,
                       vertex_byte_offset=0):
        """build_flags=None passes build_options == NULL (reference: no restructure, vlk/intersector.cpp:170).
        vertex_byte_offset: the vertices start that many bytes into their device buffer (interop pointer + offset,
        radeonrays_vlk.h:62-70), e.g. 4 for a buffer that is only 4-byte aligned."""
        g = Geometry()
        g.engine = self
        positions = np.ascontiguousarray(positions, np.float32)
        index16 = np.asarray(indices).dtype == np.uint16                 # RR_INDEX_TYPE_UINT16 (beyond the reference)
        indices = np.ascontiguousarray(indices, np.uint16 if index16 else np.uint32)
        g.triangle_count = int(indices.shape[0])
        g.vertex_count = int(positions.shape[0])
        g.vertex_stride = int(vertex_stride or positions.shape[1] * 4)
        g.vertex_byte_offset = int(vertex_byte_offset)
        g.d_vertices = _upload(np.concatenate([np.zeros(g.vertex_byte_offset, np.uint8), positions.view(np.uint8).reshape(-1)]), self.device)
        g.d_indices = _upload(indices, self.device)
        g.options = api.RRBuildOptions(build_flags, None) if build_flags is not None else None
        ctx = self.ctx
        g.p_vertices = ctx.tensor_ptr(g.d_vertices, g.vertex_byte_offset)
        g.p_indices = ctx.tensor_ptr(g.d_indices)
        g.input = ctx.geometry_input(g.p_vertices, g.vertex_count, g.vertex_stride, g.p_indices, g.triangle_count,
                                     api.RR_INDEX_TYPE_UINT16 if index16 else api.RR_INDEX_TYPE_UINT32)
        g.req = ctx.geometry_requirements(g.input, g.options)
        g.d_temp = _dev_bytes(g.req.temporary_build_buffer_size, self.device)
        g.d_nodes = _dev_bytes(g.req.result_buffer_size, self.device)
        g.p_temp = ctx.tensor_ptr(g.d_temp)
        g.p_nodes = ctx.tensor_ptr(g.d_nodes)
        self.rebuild(g)
        return g

    def rebuild(self, g):
        self.ctx.run(lambda s: self.ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, g.input, g.options, g.p_temp, g.p_nodes, s))

    def update_geometry(self, g, positions):
        g.d_vertices[getattr(g, "vertex_byte_offset", 0):].copy_(_upload(np.ascontiguousarray(positions, np.float32), self.device))
        self.ctx.run(lambda s: self.ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, g.input, g.options, g.p_temp, g.p_nodes, s))

    # ---- scene ---------------------------------------------------------------------------------------------
    def build_scene(self, geometries, instance_geometry, transforms):
        sc = Scene()
        sc.engine = self
        sc.geometries = geometries
        sc.instance_count = len(instance_geometry)
        ptrs = [geometries[i].p_nodes for i in instance_geometry]
        sc.input = self.ctx.scene_input(ptrs, transforms)
        sc.req = self.ctx.scene_requirements(sc.input)
        sc.d_temp = _dev_bytes(sc.req.temporary_build_buffer_size, self.device)
        sc.d_scene = _dev_bytes(sc.req.result_buffer_size, self.device)
        sc.p_temp = self.ctx.tensor_ptr(sc.d_temp)
        sc.p_nodes = self.ctx.tensor_ptr(sc.d_scene)
        self.ctx.run(lambda s: self.ctx.cmd_build_scene(sc.input, None, sc.p_temp, sc.p_nodes, s))
        return sc

    # ---- trace ---------------------------------------------------------------------------------------------
    def make_ray_buffers(self, ray_count, output=api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT):
        b = RayBuffers()
        b.ray_count = int(ray_count)
        b.output = output
        b.d_rays = _dev_bytes(32 * ray_count, self.device)
        b.hit_bytes = (16 if output == api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT else 4) * ray_count
        b.d_hits = _dev_bytes(b.hit_bytes, self.device)
        b.scratch_size = self.ctx.trace_requirements(ray_count)
        b.d_scratch = _dev_bytes(b.scratch_size, self.device)
        b.p_rays = self.ctx.tensor_ptr(b.d_rays)
        b.p_hits = self.ctx.tensor_ptr(b.d_hits)
        b.p_scratch = self.ctx.tensor_ptr(b.d_scratch)
        return b

    def intersect(self, target, rays, query=api.RR_INTERSECT_QUERY_CLOSEST, output=api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT,
                  init_hits=None, indirect_count=None):
        """Upload rays, trace, read hits back as a numpy structured array (HIT_DTYPE) or uint32 ids."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = rays.shape[0]
        b = self.make_ray_buffers(n, output)
        self.last_ray_buffers = b   # (tests look at the scratch header afterwards)
        b.d_rays[: 32 * n].copy_(_upload(rays, self.device))
        if init_hits is not None:
            b.d_hits[: b.hit_bytes].copy_(_upload(init_hits, self.device))
        else:
            b.d_hits.zero_()
        p_ind = None
        if indirect_count is not None:
            d_ind = torch.tensor([indirect_count], dtype=torch.int32, device=self.device)
            p_ind = self.ctx.tensor_ptr(d_ind)
        self.ctx.run(lambda s: self.ctx.cmd_intersect(target.p_nodes, query, b.p_rays, n, p_ind, output, b.p_hits, b.p_scratch, s))
        raw = b.d_hits[: b.hit_bytes].cpu().numpy()
        return raw.view(HIT_DTYPE) if output == api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT else raw.view(np.uint32)


class Geometry:
    def nodes(self):
        """Read the BLAS back: a VkBvhNode[2N-1] array (bvh_analyzer/transform.h:31-41)."""
        n = 2 * self.triangle_count - 1
        return self.d_nodes[: 64 * n].cpu().numpy().view(NODE_DTYPE)

    def scratch_u32(self, offset, count):
        return self.d_temp[offset: offset + 4 * count].cpu().numpy().view(np.uint32)


class Scene:
    def layout(self):
        return self.engine.ctx.scene_layout(self.instance_count)

    def nodes(self):
        L = self.layout()
        n = 2 * self.instance_count - 1
        return self.d_scene[L.nodes_offset: L.nodes_offset + 64 * n].cpu().numpy().view(NODE_DTYPE)

    def inverse_transforms(self):
        L = self.layout()
        rec = self.d_scene[L.records_offset: L.records_offset + 64 * self.instance_count].cpu().numpy().view(np.float32)
        return rec.reshape(self.instance_count, 16)[:, :12].copy()

    def forward_transforms(self):
        L = self.layout()
        f = self.d_scene[L.forward_transforms_offset: L.forward_transforms_offset + 48 * self.instance_count].cpu().numpy()
        return f.view(np.float32).reshape(self.instance_count, 12).copy()


class RayBuffers:
    pass


class HostTracePipeline:
    """rrCmdIntersect over a batch whose rays and hits live in HOST (pinned) memory.

    The batch is cut into `chunks` contiguous slices (multiples of 32 rays, sharding.shard_range); slice c is copied
    host->device on a copy stream, traced by its own pre-recorded command stream on the context's stream, and its hits
    are copied device->host on a third stream, so the PCIe transfers in both directions overlap each other and the
    kernel.  run() returns with the context stream waiting on the last device->host copy: an event recorded on it
    afterwards (or rrSumbitCommandStream's event of a later submit) covers the whole batch.
    """

    def __init__(self, engine, target, ray_count, query=api.RR_INTERSECT_QUERY_CLOSEST, output=api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT,
                 chunks=8):
        from .sharding import shard_range
        self.engine, self.ctx, self.n = engine, engine.ctx, int(ray_count)
        self.hit_size = 16 if output == api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT else 4
        self.buffers = engine.make_ray_buffers(ray_count, output)
        self.copy_in, self.copy_out = torch.cuda.Stream(engine.device), torch.cuda.Stream(engine.device)
        self.spans, self.streams, self.ev_in, self.ev_done, self.ev_out = [], [], [], [], []
        b = self.buffers
        for c in range(chunks):
            lo, hi = shard_range(self.n, c, chunks)
            if hi <= lo:
                continue
            cs = self.ctx.allocate_command_stream()
            p_rays = self.ctx.tensor_ptr(b.d_rays, 32 * lo)
            p_hits = self.ctx.tensor_ptr(b.d_hits, self.hit_size * lo)
            self.ctx.cmd_intersect(target.p_nodes, query, p_rays, hi - lo, None, output, p_hits, b.p_scratch, cs)
            self.spans.append((lo, hi))
            self.streams.append(cs)
            self.ev_in.append(torch.cuda.Event())
            self.ev_done.append(torch.cuda.Event())
            self.ev_out.append(torch.cuda.Event())
        self.first = True

    def run(self, h_rays, h_hits):
        """h_rays: pinned uint8 tensor [32*n]; h_hits: pinned uint8 tensor [hit_size*n]."""
        b, compute = self.buffers, self.engine.torch_stream
        for c, (lo, hi) in enumerate(self.spans):
            if not self.first:
                self.copy_in.wait_event(self.ev_done[c])      # the previous batch's trace has consumed this ray slice
                compute.wait_event(self.ev_out[c])            # ... and its hits have left the device
            with torch.cuda.stream(self.copy_in):
                b.d_rays[32 * lo: 32 * hi].copy_(h_rays[32 * lo: 32 * hi], non_blocking=True)
                self.ev_in[c].record(self.copy_in)
            compute.wait_event(self.ev_in[c])
            self.ctx.release_event(self.ctx.submit(self.streams[c]))
            self.ev_done[c].record(compute)
            self.copy_out.wait_event(self.ev_done[c])
            with torch.cuda.stream(self.copy_out):
                h_hits[self.hit_size * lo: self.hit_size * hi].copy_(b.d_hits[self.hit_size * lo: self.hit_size * hi], non_blocking=True)
                self.ev_out[c].record(self.copy_out)
        compute.wait_stream(self.copy_out)
        self.first = False

    def close(self):
        torch.cuda.synchronize(self.engine.device)
        for cs in self.streams:
            self.ctx.release_command_stream(cs)
        self.streams = []
