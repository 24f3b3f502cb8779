"""-m gpu: BASELINE.json's full-size shapes, checked through size-independent properties (the oracle would take minutes
at these sizes): sortedness / permutation / root box of a 50 M-triangle build (C5), refit == rebuild boxes, 16 Mi-ray
batches (C3) agreeing with the oracle on a random sample and with themselves across query modes, and the 4K primary batch
(C2) being identical between a fresh build, a rebuild into the same buffers and a BLAS copied to another buffer."""
import numpy as np
import pytest
import torch

from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from radeonrays_sdk_b200.host import Geometry, _dev_bytes
from helpers import assert_hits_equal

pytestmark = pytest.mark.gpu
CLOSEST, ANY = api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_ANY
FULL, IDS = api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID


def _device_heightfield(nx, nz, t, dev):
    xs = torch.arange(nx + 1, dtype=torch.float32, device=dev)
    zs = torch.arange(nz + 1, dtype=torch.float32, device=dev)
    Z, X = torch.meshgrid(zs, xs, indexing="ij")
    Y = 2.0 * torch.sin(0.05 * X + 1.0 * t) * torch.cos(0.07 * Z)
    pos = torch.stack([X, Y, Z], -1).reshape(-1, 3).contiguous()
    i = (torch.arange(nz, dtype=torch.int64, device=dev)[:, None] * (nx + 1) + torch.arange(nx, dtype=torch.int64, device=dev)[None, :]).reshape(-1)
    a, b, c, d = i, i + 1, i + nx + 1, i + nx + 2
    idx = torch.stack([torch.stack([a, c, b], -1), torch.stack([b, c, d], -1)], 1).reshape(-1, 3).to(torch.int32).contiguous()
    return pos, idx


def _device_geometry(engine, pos, idx):
    ctx, dev, n = engine.ctx, engine.device, idx.shape[0]
    g = Geometry()
    g.engine, g.triangle_count, g.vertex_count, g.vertex_stride = engine, n, pos.shape[0], 12
    g.d_vertices, g.d_indices = pos.view(torch.uint8).reshape(-1), idx.view(torch.uint8).reshape(-1)
    g.options = api.RRBuildOptions(api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, None)
    g.p_vertices, g.p_indices = ctx.tensor_ptr(g.d_vertices), ctx.tensor_ptr(g.d_indices)
    g.input = ctx.geometry_input(g.p_vertices, g.vertex_count, 12, g.p_indices, n)
    g.req = ctx.geometry_requirements(g.input, g.options)
    g.d_temp, g.d_nodes = _dev_bytes(g.req.temporary_build_buffer_size, dev), _dev_bytes(g.req.result_buffer_size, dev)
    g.p_temp, g.p_nodes = ctx.tensor_ptr(g.d_temp), ctx.tensor_ptr(g.d_nodes)
    return g


def _node_view(g):
    n = 2 * g.triangle_count - 1
    return g.d_nodes[: 64 * n].view(torch.int32).reshape(n, 16)


def _check_tree_on_device(g, pos):
    """Structure of a (2N-1)-node BLAS without leaving the GPU: every non-root node is the child of its parent, the
    children partition the node set, parents contain their children's boxes, the root box is the mesh box."""
    n = g.triangle_count
    nodes = _node_view(g)
    f = nodes.view(torch.float32)
    c0, c1, parent = nodes[: n - 1, 3].long(), nodes[: n - 1, 7].long(), nodes[:, 11].long()
    assert int(parent[0].item()) == -1 and bool((nodes[n - 1:, 3] == -1).all())        # root / leaf markers
    kids = torch.cat([c0, c1])
    assert bool((torch.sort(kids).values == torch.arange(1, 2 * n - 1, device=kids.device)).all()), "children must be a partition of the non-root nodes"
    idx = torch.arange(n - 1, device=kids.device)
    assert bool((parent[c0] == idx).all()) and bool((parent[c1] == idx).all()), "parent words"
    prims = nodes[n - 1:, 7].long()
    assert bool((torch.sort(prims).values == torch.arange(n, device=prims.device)).all()), "leaf prim ids must be a permutation"

    def box_of(node_ids):
        leaf = node_ids >= n - 1
        q = f[node_ids]
        tri_lo = torch.minimum(torch.minimum(q[:, 0:3], q[:, 4:7]), q[:, 8:11])
        tri_hi = torch.maximum(torch.maximum(q[:, 0:3], q[:, 4:7]), q[:, 8:11])
        int_lo, int_hi = torch.minimum(q[:, 0:3], q[:, 8:11]), torch.maximum(q[:, 4:7], q[:, 12:15])
        return torch.where(leaf[:, None], tri_lo, int_lo), torch.where(leaf[:, None], tri_hi, int_hi)

    for child, lo_cols, hi_cols in ((c0, slice(0, 3), slice(4, 7)), (c1, slice(8, 11), slice(12, 15))):
        lo, hi = box_of(child)
        assert bool((f[: n - 1, lo_cols] == lo).all()) and bool((f[: n - 1, hi_cols] == hi).all()), "stored child box == child's own box, bit for bit"
    rlo, rhi = box_of(torch.zeros(1, dtype=torch.long, device=kids.device))
    assert bool((rlo[0] == pos.min(0).values).all()) and bool((rhi[0] == pos.max(0).values).all())


@pytest.mark.parametrize("nx,nz", [(4000, 1000), (5000, 5000)])
def test_c5_large_build_and_refit(engine, nx, nz):
    """8 M and 50 M triangles (BASELINE config C5): build, refit to the next frame, rebuild of that frame."""
    free, _ = torch.cuda.mem_get_info()
    n = 2 * nx * nz
    if free < 260 * n + (2 << 30):
        pytest.skip("not enough free device memory for this size")
    dev = engine.device
    pos, idx = _device_heightfield(nx, nz, 0.0, dev)
    g = _device_geometry(engine, pos, idx)
    engine.rebuild(g)
    L = engine.ctx.build_scratch_layout(n)
    codes = g.d_temp[L.sorted_codes_offset: L.sorted_codes_offset + 4 * n].view(torch.int32)
    refs = g.d_nodes[L.sorted_refs_offset: L.sorted_refs_offset + 4 * n].view(torch.int32)   # kept in the geometry buffer's tail
    assert bool((codes[1:] >= codes[:-1]).all()), "sorted Morton codes"
    ties = codes[1:] == codes[:-1]
    assert bool((refs[1:][ties] > refs[:-1][ties]).all()), "equal codes keep ascending primitive order (stable sort)"
    del codes, refs, ties
    _check_tree_on_device(g, pos)
    topo = _node_view(g)[:, [3, 7, 11]].clone()
    # frame t = 1: refit in place, then compare with a fresh build of the same frame wherever the topology agrees
    pos1, _ = _device_heightfield(nx, nz, 1.0, dev)
    g.d_vertices.copy_(pos1.view(torch.uint8).reshape(-1))
    engine.ctx.run(lambda s: engine.ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, g.input, g.options, g.p_temp, g.p_nodes, s))
    assert bool((_node_view(g)[:, [3, 7, 11]] == topo).all()), "refit must not touch the topology"
    _check_tree_on_device(g, pos1)
    # refit twice == refit once (the parity rendezvous needs no reset)
    once = _node_view(g)[:, :15].clone()
    engine.ctx.run(lambda s: engine.ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, g.input, g.options, g.p_temp, g.p_nodes, s))
    assert bool((_node_view(g)[:, :15] == once).all())


def test_c3_16m_ray_batches(engine, sponza):
    """16 Mi shadow (ANY, ids) and diffuse-bounce (CLOSEST, full hit) rays on Sponza: a random 100 000-ray sample equals the
    oracle bit for bit; any-hit and closest-hit agree on WHICH rays hit; tracing the batch twice is idempotent."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    nodes = g.nodes()
    prim = W.sponza_primary_rays(1024, 1024)
    hits = engine.intersect(g, prim)
    count = 1 << 24
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(count, 100_000, replace=False))
    for name, rays in (("shadow", W.shadow_rays(pos, idx, prim, hits, count=count)), ("diffuse", W.diffuse_rays(pos, idx, prim, hits, count=count))):
        closest = engine.intersect(g, rays, CLOSEST, FULL)
        anyids = engine.intersect(g, rays, ANY, IDS)
        assert np.array_equal(closest["inst_id"] != O.INVALID, anyids != O.INVALID), f"{name}: any-hit and closest-hit must agree on hit / miss"
        assert_hits_equal(closest[sel], O.trace(nodes, rays[sel], init=np.zeros(sel.size, W.HIT_DTYPE)), what=f"{name} sample", mesh=(pos, idx), rays=rays[sel])
        assert np.array_equal(anyids[sel], O.trace(nodes, rays[sel], O.QUERY_ANY, O.OUTPUT_INSTANCE_ID))
        again = engine.intersect(g, rays, CLOSEST, FULL)
        assert np.array_equal(again.view(np.uint8), closest.view(np.uint8)), f"{name}: not deterministic"
        ids = engine.intersect(g, rays, CLOSEST, IDS)
        ok = closest["inst_id"] != O.INVALID
        assert np.array_equal(ids[ok], closest["prim_id"][ok]) and np.all(ids[~ok] == O.INVALID)


def test_c2_4k_primary_is_reproducible_and_position_independent(engine, sponza):
    """3840x2160 primary rays (C2): identical hits from a fresh build, a rebuild into the same buffers, and a byte copy
    of the BLAS at another address (what the multi-GPU broadcast relies on); a 1/64 sample equals the oracle."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    rays = W.sponza_primary_rays(3840, 2160)
    first = engine.intersect(g, rays)
    sel = np.arange(0, rays.shape[0], 64)
    assert_hits_equal(first[sel], O.trace(g.nodes(), rays[sel], init=np.zeros(sel.size, W.HIT_DTYPE)), what="4K sample", mesh=(pos, idx), rays=rays[sel])
    engine.rebuild(g)
    assert np.array_equal(engine.intersect(g, rays).view(np.uint8), first.view(np.uint8))
    clone = Geometry()
    clone.d_nodes = g.d_nodes.clone()
    clone.p_nodes = engine.ctx.tensor_ptr(clone.d_nodes)
    assert np.array_equal(engine.intersect(clone, rays).view(np.uint8), first.view(np.uint8))
    assert (first["inst_id"] != O.INVALID).all()          # the camera is inside the atrium: every ray hits


def test_c4_thousand_instances_full_size(engine, sponza):
    """BASELINE config C4 at full size (test/test_vk/basic_test.h:752-1069 is the reference's two-level scenario): 1 000 rotated
    instances of the Sponza BLAS under one TLAS, 3840 x 2160 primary rays from outside the grid.  TLAS nodes and transforms equal
    the oracle's bit for bit, a 200 000-ray sample equals the oracle's two-level walk (closest FULL_HIT and ANY ids), the device
    ray count (isect_2l.comp:146-151) limits a two-level trace, and the batch is reproducible."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    blas = [g.nodes()]
    xf = W.grid_instances(10, 250.0, 7.0)
    inst = [0] * xf.shape[0]
    sc = engine.build_scene([g], inst, xf)
    tlas, oxf = O.build_tlas(blas, inst, xf)
    from helpers import assert_nodes_equal
    assert_nodes_equal(sc.nodes(), tlas, what="C4 tlas")
    assert np.array_equal(sc.inverse_transforms().view(np.uint32), oxf[0::2].view(np.uint32))
    rays = W.grid_camera_rays(3840, 2160)
    n = rays.shape[0]
    hits = engine.intersect(sc, rays)
    assert np.unique(hits["inst_id"][hits["inst_id"] != O.INVALID]).size > 300
    sel = np.sort(np.random.default_rng(0).choice(n, 200_000, replace=False))
    want = O.trace_2l(tlas, oxf, blas, inst, rays[sel], init=np.zeros(sel.size, W.HIT_DTYPE))
    assert_hits_equal(hits[sel], want, what="C4 sample")
    anyids = engine.intersect(sc, rays, ANY, IDS)
    assert np.array_equal(anyids[sel], O.trace_2l(tlas, oxf, blas, inst, rays[sel], O.QUERY_ANY, O.OUTPUT_INSTANCE_ID))
    assert np.array_equal(anyids != O.INVALID, hits["inst_id"] != O.INVALID)
    assert np.array_equal(engine.intersect(sc, rays).view(np.uint8), hits.view(np.uint8)), "C4 trace is not deterministic"
    # two-level trace limited by a device-side ray count
    init = np.zeros(n, W.HIT_DTYPE)
    init["inst_id"] = 0xABCD
    part = engine.intersect(sc, rays, init_hits=init, indirect_count=100_001)
    assert np.array_equal(part[:100_001].view(np.uint8), hits[:100_001].view(np.uint8))   # (a miss writes inst_id only; uv / prim were 0 in both runs)
    assert np.all(part[100_001:]["inst_id"] == 0xABCD), "rays past the device count must not be traced"
