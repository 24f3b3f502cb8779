"""-m gpu: top-level (instance) BVH build and two-level traversal parity through rrCmdBuildScene / rrCmdIntersect."""
import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from helpers import assert_hits_equal, assert_nodes_equal

pytestmark = pytest.mark.gpu
CLOSEST, ANY = api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_ANY
FULL, IDS = api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID


def _identity(n):
    m = np.zeros((n, 3, 4), np.float32)
    m[:, 0, 0] = m[:, 1, 1] = m[:, 2, 2] = 1
    return m


def _check_scene(engine, geoms, blas_nodes, inst_geom, xf, rays, quirk=False):
    engine.ctx.set_option(api.RR_CUDA_OPTION_REFERENCE_TRANSFORM_AABB_QUIRK, int(quirk))
    try:
        sc = engine.build_scene(geoms, inst_geom, xf)
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_REFERENCE_TRANSFORM_AABB_QUIRK, 0)
    tlas, out_xf = O.build_tlas(blas_nodes, inst_geom, xf, reference_corner_quirk=quirk)
    assert_nodes_equal(sc.nodes(), tlas, what="tlas")
    n = len(inst_geom)
    assert np.array_equal(sc.inverse_transforms().view(np.uint32), out_xf[0::2].view(np.uint32)), "inverse transforms"
    assert np.array_equal(sc.forward_transforms().view(np.uint32), out_xf[1::2].view(np.uint32)), "forward transforms"
    for query in (CLOSEST, ANY):
        got = engine.intersect(sc, rays, query, FULL)
        want = O.trace_2l(tlas, out_xf, blas_nodes, inst_geom, rays, query, O.OUTPUT_FULL_HIT)
        assert_hits_equal(got, want, what=f"2-level q={query}")
        got = engine.intersect(sc, rays, query, IDS)
        want = O.trace_2l(tlas, out_xf, blas_nodes, inst_geom, rays, query, O.OUTPUT_INSTANCE_ID)
        assert np.array_equal(got, want)
    return sc


def test_sponza_per_shape_instances(engine, sponza):
    """basic_test.h:752-1069 BuildObj2Level: one BLAS per OBJ shape (390), identity transforms."""
    pos, idx, first = sponza
    geoms, blas = [], []
    for s in range(len(first) - 1):
        sub = idx[first[s]:first[s + 1]]
        g = engine.build_geometry(pos, sub)
        geoms.append(g)
        blas.append(g.nodes())
    inst = list(range(len(geoms)))
    rays = W.sponza_primary_rays(320, 320)
    sc = _check_scene(engine, geoms, blas, inst, _identity(len(inst)), rays)
    # two-level over per-shape BLASes sees the same surfaces as the one-level BVH over the whole mesh
    whole = engine.build_geometry(pos, idx)
    one = engine.intersect(whole, rays)
    two = engine.intersect(sc, rays)
    assert np.array_equal(one["inst_id"] == O.INVALID, two["inst_id"] == O.INVALID)


def test_instanced_grid_rotated(engine, cornell):
    """BASELINE config C4 shape at test size: one BLAS instanced on a grid with per-instance Y rotations."""
    pos, idx, _ = cornell
    g = engine.build_geometry(pos, idx)
    blas = [g.nodes()]
    xf = W.grid_instances(n_side=4, spacing=3.0, degrees_per_instance=7.0)
    inst = [0] * xf.shape[0]
    rng = np.random.default_rng(11)
    rays = W.random_rays(100_000, (-2, -2, -2), (12, 12, 12), seed=7)
    _check_scene(engine, [g], blas, inst, xf, rays)
    _check_scene(engine, [g], blas, inst, xf, rays[:20000], quirk=True)


def test_single_instance_and_single_triangle_blas(engine):
    pos, idx = W.single_triangle()
    g = engine.build_geometry(pos, idx)
    rays = np.zeros(64, W.RAY_DTYPE)
    rng = np.random.default_rng(2)
    rays["origin"] = np.c_[rng.random(64) * 2 - 1, rng.random(64) * 2 - 1, np.zeros(64)]
    rays["direction"] = (0, 0, 1)
    rays["min_t"], rays["max_t"] = 0.001, 1000.0
    xf = _identity(1)
    xf[0, 0, 3] = 0.25
    _check_scene(engine, [g], [g.nodes()], [0], xf, rays)
    xf3 = _identity(3)
    xf3[1, 0, 3], xf3[2, 1, 3] = 0.5, -0.5
    _check_scene(engine, [g], [g.nodes()], [0, 0, 0], xf3, rays)


def test_scene_buffer_is_self_describing(engine, cornell):
    """SURVEY.md section 8b "Ownership": the reference keeps "this buffer is a scene" in host state keyed by the buffer
    (vlk/intersector.cpp:86,263,289-290); here the buffer carries a header and the kernels tell scene from geometry on the device.
    So: a scene built through context A is traced through context B; a byte copy of the scene at another address traces
    identically; an intersect recorded BEFORE the scene build in the same stream still runs two-level; and a scene whose geometry
    was copied elsewhere is re-pointed with rrCudaCmdRebindSceneGeometry."""
    import torch
    from radeonrays_sdk_b200.host import Scene, Geometry
    pos, idx, _ = cornell
    g = engine.build_geometry(pos, idx)
    xf = W.grid_instances(n_side=3, spacing=3.0, degrees_per_instance=11.0)
    inst = [0] * xf.shape[0]
    sc = engine.build_scene([g], inst, xf)
    rays = W.random_rays(50_000, (-2, -2, -2), (9, 9, 9), seed=21)
    want = engine.intersect(sc, rays)
    assert (want["inst_id"] != O.INVALID).sum() > 5000
    # (1) another context on the same device
    other = api.Context(device=engine.device.index, cuda_stream=engine.torch_stream.cuda_stream)
    try:
        rb = engine.make_ray_buffers(rays.shape[0])
        rb.d_rays[: 32 * rays.shape[0]].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(engine.device))
        rb.d_hits.zero_()
        p = [other.tensor_ptr(t) for t in (sc.d_scene, rb.d_rays, rb.d_hits, rb.d_scratch)]
        other.run(lambda s: other.cmd_intersect(p[0], CLOSEST, p[1], rays.shape[0], None, FULL, p[2], p[3], s))
        got = rb.d_hits[: 16 * rays.shape[0]].cpu().numpy().view(W.HIT_DTYPE)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), "scene traced through another context"
    finally:
        other.destroy()
    # (2) a byte copy of the scene buffer
    clone = Scene()
    clone.d_scene = sc.d_scene.clone()
    clone.p_nodes = engine.ctx.tensor_ptr(clone.d_scene)
    assert np.array_equal(engine.intersect(clone, rays).view(np.uint8), want.view(np.uint8)), "cloned scene"
    # (3) the geometry moves too: copy the BLAS, re-point the clone's instances, scribble over the old BLAS
    g2 = Geometry()
    g2.d_nodes = g.d_nodes.clone()
    g2.p_nodes = engine.ctx.tensor_ptr(g2.d_nodes)
    engine.ctx.run(lambda s: engine.ctx.cmd_rebind_scene_geometry(clone.p_nodes, g.d_nodes.data_ptr(), g2.p_nodes, s))
    saved = g.d_nodes.clone()
    g.d_nodes.fill_(0xFF)
    try:
        assert np.array_equal(engine.intersect(clone, rays).view(np.uint8), want.view(np.uint8)), "rebound scene"
    finally:
        g.d_nodes.copy_(saved)
    # (4) record order: intersect recorded before the build that makes the buffer a scene, same command stream
    ctx = engine.ctx
    fresh = torch.zeros_like(sc.d_scene)
    p_fresh = ctx.tensor_ptr(fresh)
    rb = engine.make_ray_buffers(rays.shape[0])
    rb.d_rays[: 32 * rays.shape[0]].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(engine.device))
    rb.d_hits.zero_()
    cs = ctx.allocate_command_stream()
    ctx.cmd_build_scene(sc.input, None, sc.p_temp, p_fresh, cs)
    ctx.cmd_intersect(p_fresh, CLOSEST, rb.p_rays, rays.shape[0], None, FULL, rb.p_hits, rb.p_scratch, cs)
    ev = ctx.submit(cs)
    ctx.wait(ev)
    ctx.release_event(ev)
    ctx.release_command_stream(cs)
    assert np.array_equal(rb.d_hits[: 16 * rays.shape[0]].cpu().numpy().view(W.HIT_DTYPE).view(np.uint8), want.view(np.uint8))


def test_external_command_stream_interop(engine, cornell):
    """rrGetCommandStreamFromCudaStream / rrReleaseExternalCommandStream (reference: rrGetCommandStreamFromVkCommandBuffer,
    src/core/src/radeonrays.cpp:686-735): build + intersect recorded into a command stream that wraps the CLIENT's own
    cudaStream_t run on that stream, in order with the client's own work on it."""
    import torch
    pos, idx, _ = cornell
    ctx = engine.ctx
    g = engine.build_geometry(pos, idx)
    want_nodes = g.nodes()
    rays = W.cornell_primary_rays(96)
    n = rays.shape[0]
    mine = torch.cuda.Stream(engine.device)
    rb = engine.make_ray_buffers(n)
    torch.cuda.synchronize()
    ext = ctx.command_stream_from_cuda_stream(mine.cuda_stream)
    g.d_nodes.zero_()
    torch.cuda.synchronize()
    with torch.cuda.stream(mine):
        # client work on its own stream before the library's: upload the rays there
        rb.d_rays[: 32 * n].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory(), non_blocking=True)
        rb.d_hits.zero_()
    ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, g.input, g.options, g.p_temp, g.p_nodes, ext)
    ctx.cmd_intersect(g.p_nodes, CLOSEST, rb.p_rays, n, None, FULL, rb.p_hits, rb.p_scratch, ext)
    ev = ctx.submit(ext)
    with torch.cuda.stream(mine):
        hits_dev = rb.d_hits[: 16 * n].clone()          # client work after the library's, same stream: ordered behind it
    ctx.wait(ev)
    ctx.release_event(ev)
    mine.synchronize()
    assert ctx.lib.rrReleaseExternalCommandStream(ctx.handle, ext) == api.RR_SUCCESS
    assert_nodes_equal(g.nodes(), want_nodes, what="build on an external stream")
    got = hits_dev.cpu().numpy().view(W.HIT_DTYPE)
    assert_hits_equal(got, O.trace(want_nodes, rays, init=np.zeros(n, W.HIT_DTYPE)), what="external stream", mesh=(pos, idx), rays=rays)
