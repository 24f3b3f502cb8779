"""-m gpu: closest / any-hit traversal parity (ids bit-exact, uv 1e-6 relative) through rrCmdIntersect."""
import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from helpers import assert_hits_equal, assert_nodes_equal

pytestmark = pytest.mark.gpu

CLOSEST, ANY = api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_ANY
FULL, IDS = api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID


def _all_modes(engine, g, nodes, rays, what, mesh):
    rng = np.random.default_rng(1)
    init = np.zeros(rays.shape[0], W.HIT_DTYPE)
    init["uv"] = rng.random((rays.shape[0], 2), dtype=np.float32)
    init["prim_id"] = 12345
    init["inst_id"] = 777
    for query in (CLOSEST, ANY):
        got = engine.intersect(g, rays, query, FULL, init_hits=init)
        want = O.trace(nodes, rays, query, O.OUTPUT_FULL_HIT, init=init)
        closest = mesh if query == CLOSEST else None   # ANY keeps the reference's visit order: bit-exact, no tie allowance
        assert_hits_equal(got, want, what=f"{what} q={query} full", mesh=closest, rays=rays)
        got = engine.intersect(g, rays, query, IDS)
        want = O.trace(nodes, rays, query, O.OUTPUT_INSTANCE_ID)
        if closest:
            assert_hits_equal(got, want, what=f"{what} q={query} ids", mesh=closest, rays=rays)
        else:
            assert np.array_equal(got, want), f"{what} q={query} ids"


def test_single_triangle_hit_and_miss(engine):
    pos, idx = W.single_triangle()
    g = engine.build_geometry(pos, idx)
    rays = np.zeros(4, W.RAY_DTYPE)
    rays["origin"] = [(0, 0, 0), (0, 0, 0), (5, 5, 0), (0, 0, 2)]
    rays["direction"] = [(0, 0, 1), (0, 0, -1), (0, 0, 1), (0, 0, -1)]
    rays["min_t"], rays["max_t"] = 0.001, 100000.0
    _all_modes(engine, g, g.nodes(), rays, "single triangle", (pos, idx))
    hits = engine.intersect(g, rays)
    assert list(hits["inst_id"]) == [0, O.INVALID, O.INVALID, 0]


def test_cornell_1024(engine, cornell):
    """BASELINE config C1: Cornell box, 1024x1024 primary closest-hit rays."""
    pos, idx, _ = cornell
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    rays = W.cornell_primary_rays(1024)
    got = engine.intersect(g, rays)
    want = O.trace(nodes, rays)
    ties = assert_hits_equal(got, want, what="cornell 1024^2", mesh=(pos, idx), rays=rays)
    assert ties <= 16
    bf, _ = O.brute_force(pos, idx, rays[::97])
    ok = bf["inst_id"] != O.INVALID
    # exhaustive search agrees except possibly on exact-t ties across a culled subtree (see test_oracle_cpu.py)
    assert (got["prim_id"][::97][ok] != bf["prim_id"][ok]).sum() <= 2
    _all_modes(engine, g, nodes, W.cornell_primary_rays(128), "cornell 128^2", (pos, idx))


@pytest.mark.parametrize("flags", [api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, 0])
def test_sponza_primary_640(engine, sponza, flags):
    """internal_resources_test.h: 640x640 canonical primary rays; fast and quality builds."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=flags)
    nodes = g.nodes()
    rays = W.sponza_primary_rays(640, 640)
    _all_modes(engine, g, nodes, rays, f"sponza flags={flags}", (pos, idx))


def test_sponza_tie_rules(engine, sponza):
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    rays = W.sponza_primary_rays(512, 512)
    engine.ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 1)
    try:
        got = engine.intersect(g, rays)
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 0)
    assert_hits_equal(got, O.trace(nodes, rays, tie=O.TIE_FIRST_FOUND), what="first-found rule")


def test_sponza_incoherent_and_secondary(engine, sponza):
    """BASELINE config C3 shapes at test size: shadow (ANY, ids) and diffuse bounce (CLOSEST, full hit)."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    prim = W.sponza_primary_rays(256, 256)
    hits = engine.intersect(g, prim)
    shadow = W.shadow_rays(pos, idx, prim, hits, count=100_000)
    diffuse = W.diffuse_rays(pos, idx, prim, hits, count=100_000)
    got = engine.intersect(g, shadow, ANY, IDS)
    assert np.array_equal(got, O.trace(nodes, shadow, O.QUERY_ANY, O.OUTPUT_INSTANCE_ID))
    assert_hits_equal(engine.intersect(g, diffuse), O.trace(nodes, diffuse), what="diffuse", mesh=(pos, idx), rays=diffuse)
    lo, hi = pos.min(0), pos.max(0)
    rnd = W.random_rays(200_000, lo, hi)
    _all_modes(engine, g, nodes, rnd, "random rays", (pos, idx))


def test_indirect_ray_count_and_ragged_sizes(engine, cornell):
    pos, idx, _ = cornell
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    for n in (1, 31, 33, 127, 129, 1000):
        rays = W.cornell_primary_rays(32)[:n]
        assert_hits_equal(engine.intersect(g, rays), O.trace(nodes, rays), what=f"n={n}", mesh=(pos, idx), rays=rays)
    rays = W.cornell_primary_rays(32)
    init = np.zeros(rays.shape[0], W.HIT_DTYPE)
    init["inst_id"] = 0xABCD
    got = engine.intersect(g, rays, init_hits=init, indirect_count=100)   # isect.comp:98-103
    want = O.trace(nodes, rays[:100], init=init[:100])
    assert_hits_equal(got[:100], want, mesh=(pos, idx), rays=rays[:100])
    assert np.all(got["inst_id"][100:] == 0xABCD)


def test_deep_stack_spill(engine):
    """A degenerate mesh (all Morton codes equal) whose rays defer more nodes than the shared-memory stack holds."""
    n = 4096
    rng = np.random.default_rng(3)
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    pos = np.concatenate([base + np.array([0, 0, 1e-4 * i], np.float32) for i in range(n)]).astype(np.float32)
    # squash the centroids so that codes collide but every triangle still spans the same footprint
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    rays = np.zeros(256, W.RAY_DTYPE)
    rays["origin"] = np.c_[rng.random(256) * 0.4 + 0.05, rng.random(256) * 0.4 + 0.05, -np.ones(256)]
    rays["direction"] = (0, 0, 1)
    rays["min_t"], rays["max_t"] = 0.0, 1000.0
    got = engine.intersect(g, rays)
    want, st = O.trace(nodes, rays, want_stats=True)
    assert_hits_equal(got, want, what="deep stack", mesh=(pos, idx), rays=rays)


def test_misaligned_buffers_are_rejected(engine, cornell):
    """Interop pointer + offset (radeonrays_vlk.h:62-70 style): node arrays must be 64-byte aligned, rays / full hits /
    scratch 16-byte aligned; anything else is RR_ERROR_INVALID_PARAMETER at record time, never a misaligned-address fault."""
    import ctypes as C
    import torch
    pos, idx, _ = cornell
    g = engine.build_geometry(pos, idx)
    ctx = engine.ctx
    rays = W.cornell_primary_rays(8)
    rb = engine.make_ray_buffers(rays.shape[0])
    cs = ctx.allocate_command_stream()

    def intersect(scene, rays_p, hits_p, scratch_p):
        return ctx.lib.rrCmdIntersect(ctx.handle, scene, api.RR_INTERSECT_QUERY_CLOSEST, rays_p, rays.shape[0], None,
                                      api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, hits_p, scratch_p, cs)

    assert intersect(g.p_nodes, rb.p_rays, rb.p_hits, rb.p_scratch) == api.RR_SUCCESS
    assert intersect(ctx.tensor_ptr(g.d_nodes, 32), rb.p_rays, rb.p_hits, rb.p_scratch) == api.RR_ERROR_INVALID_PARAMETER
    assert intersect(g.p_nodes, ctx.tensor_ptr(rb.d_rays, 8), rb.p_hits, rb.p_scratch) == api.RR_ERROR_INVALID_PARAMETER
    assert intersect(g.p_nodes, rb.p_rays, ctx.tensor_ptr(rb.d_hits, 4), rb.p_scratch) == api.RR_ERROR_INVALID_PARAMETER
    assert intersect(g.p_nodes, rb.p_rays, rb.p_hits, ctx.tensor_ptr(rb.d_scratch, 4)) == api.RR_ERROR_INVALID_PARAMETER
    # a geometry buffer that is not 64-byte aligned is rejected by the build as well
    spare = torch.empty(g.req.result_buffer_size + 64, dtype=torch.uint8, device=engine.device)
    rc = ctx.lib.rrCmdBuildGeometry(ctx.handle, api.RR_BUILD_OPERATION_BUILD, C.byref(g.input), C.byref(g.options), g.p_temp,
                                    ctx.tensor_ptr(spare, 16), cs)
    assert rc == api.RR_ERROR_INVALID_PARAMETER
    ctx.release_command_stream(cs)


def test_resubmitted_command_stream_is_replayed_as_a_graph(engine, sponza):
    """A renderer's per-frame list -- update the BLAS, trace -- recorded once and submitted every frame: from the second
    submit on the library replays it as a CUDA graph (rr_api.cpp rrSumbitCommandStream); every frame must see that frame's
    vertices, and appending a command afterwards must re-capture."""
    import torch
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx)
    ctx = engine.ctx
    rays = W.sponza_primary_rays(160, 90)
    rb = engine.make_ray_buffers(rays.shape[0])
    rb.d_rays[: 32 * rays.shape[0]].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)))
    before = g.nodes()
    cs = ctx.allocate_command_stream()
    ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, g.input, g.options, g.p_temp, g.p_nodes, cs)
    ctx.cmd_intersect(g.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST, rb.p_rays, rays.shape[0], None,
                      api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, rb.p_hits, rb.p_scratch, cs)
    l0 = ctx.launch_count()
    per_submit = None
    for frame in range(4):
        moved = (pos + np.float32(0.5 * frame)).astype(np.float32)
        g.d_vertices.copy_(torch.from_numpy(moved.view(np.uint8).reshape(-1)))
        rb.d_hits.zero_()
        torch.cuda.synchronize()
        e = ctx.submit(cs)
        ctx.wait(e)
        ctx.release_event(e)
        n_launch = ctx.launch_count() - l0
        l0 = ctx.launch_count()
        per_submit = per_submit or n_launch
        assert n_launch == per_submit > 0                      # the replay accounts for the same kernels
        want_nodes = O.refit(before, moved, idx)
        assert_nodes_equal(g.nodes(), want_nodes, what=f"frame {frame}")
        got = rb.d_hits[: 16 * rays.shape[0]].cpu().numpy().view(W.HIT_DTYPE)
        assert_hits_equal(got, O.trace(want_nodes, rays, init=np.zeros(rays.shape[0], W.HIT_DTYPE)), what=f"frame {frame}", mesh=(moved, idx), rays=rays)
    # one more command: the stream is captured again, both traces run
    ids = torch.zeros(rays.shape[0], dtype=torch.int32, device=engine.device)
    ctx.cmd_intersect(g.p_nodes, api.RR_INTERSECT_QUERY_ANY, rb.p_rays, rays.shape[0], None,
                      api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID, ctx.tensor_ptr(ids), rb.p_scratch, cs)
    for _ in range(2):
        ids.fill_(-2)
        e = ctx.submit(cs)
        ctx.wait(e)
        ctx.release_event(e)
        assert int((ids == -2).sum().item()) == 0
    ctx.release_command_stream(cs)


def test_ray_binning_option_is_bit_identical(engine, sponza, cornell):
    """RR_CUDA_OPTION_SORT_RAYS: rays binned on the device (octant | origin cell | coarse direction, 3-pass radix sort) and traced in
    that order; every mode must return exactly what the unsorted trace returns, in the client's order -- one level, two level,
    ragged counts, a device-side ray count, misses left untouched."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    prim = W.sponza_primary_rays(256, 256)
    hits = engine.intersect(g, prim)
    diffuse = W.diffuse_rays(pos, idx, prim, hits, count=150_001)
    rng = np.random.default_rng(9)
    init = np.zeros(diffuse.shape[0], W.HIT_DTYPE)
    init["uv"] = rng.random((diffuse.shape[0], 2), dtype=np.float32)
    init["prim_id"], init["inst_id"] = 4242, 77
    cpos, cidx, _ = cornell
    cg = engine.build_geometry(cpos, cidx)
    xf = W.grid_instances(n_side=3, spacing=3.0, degrees_per_instance=9.0)
    sc = engine.build_scene([cg], [0] * xf.shape[0], xf)
    rnd = W.random_rays(70_003, (-2, -2, -2), (9, 9, 9), seed=5)
    cases = [(g, diffuse, CLOSEST, FULL, init, None), (g, diffuse, ANY, IDS, None, None), (g, diffuse, CLOSEST, IDS, None, 100_000),
             (sc, rnd, CLOSEST, FULL, None, None), (sc, rnd, ANY, FULL, None, 12_345), (g, diffuse[:31], CLOSEST, FULL, None, None)]
    engine.ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 1)     # per-ray kernel in both runs: bit for bit
    try:
        plain = [engine.intersect(t, r, q, o, init_hits=(i[: r.shape[0]] if i is not None else None), indirect_count=c) for t, r, q, o, i, c in cases]
        engine.ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 1)
        binned = [engine.intersect(t, r, q, o, init_hits=(i[: r.shape[0]] if i is not None else None), indirect_count=c) for t, r, q, o, i, c in cases]
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 0)
        engine.ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 0)
    for k, (a, b) in enumerate(zip(plain, binned)):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"case {k}: binned trace differs"
    # default tie rule + binning (per-ray kernel) against the oracle
    engine.ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 1)
    try:
        got = engine.intersect(g, diffuse)
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 0)
    assert_hits_equal(got, O.trace(g.nodes(), diffuse), what="binned diffuse", mesh=(pos, idx), rays=diffuse)


def _grid_words(engine, with_mixed=False):
    """Scratch header words 5 / 6 (/ 7) after an rrCmdIntersect: the row length k_detect_grid settled on (0: none), the phase of
    row 0 (and whether it found the batch incoherent, so that the packet kernel stood aside)."""
    w = engine.last_ray_buffers.d_scratch[20:32].cpu().numpy().view(np.int32)
    return (int(w[0]), int(w[1]), int(w[2])) if with_mixed else (int(w[0]), int(w[1]))


def _normalised(rays):
    out = rays.copy()
    d = out["direction"].astype(np.float64)
    out["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return out


def _orthographic(w, h):
    """Parallel rays through Sponza's nave from a w x h grid of origins: all directions equal, the ORIGINS step along a row."""
    rays = np.zeros(w * h, W.RAY_DTYPE)
    x = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0)
    y = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1)
    rays["origin"] = np.stack([np.full((h, w), 12.0, np.float32), 1.0 + 12.0 * y / h, -4.0 + 8.0 * x / w], -1).reshape(-1, 3)
    rays["direction"] = (-1.0, 0.02, 0.01)
    rays["min_t"], rays["max_t"] = 0.001, 1e5
    return rays


def test_ray_grid_tiles_are_bit_identical(engine, sponza):
    """Closest-hit packets are 8 x 8 tiles of the ray grid when the batch is an image in row order (k_detect_grid finds the row
    length on the device; RR_CUDA_OPTION_RAY_GRID_WIDTH).  The grouping must never change a hit: tiles (detected, and told), and
    64-ray strips return the same bytes -- for widths that are no multiple of 8, a ragged last row, a batch that starts in mid-row
    (a shard of a frame), a device-side ray count, rows too short or too few to tile, and rays with no grid at all."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    ctx = engine.ctx
    rng = np.random.default_rng(3)

    def run(rays, mode, **kw):
        ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, mode)
        try:
            return engine.intersect(g, rays, **kw), _grid_words(engine)
        finally:
            ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 0)

    img = W.sponza_primary_rays(331, 203)                      # 331 = 41 * 8 + 3 columns, 203 = 25 * 8 + 3 rows
    cases = [("full frame", img, 331, 0, {}),
             ("ragged last row", img[: 331 * 150 + 77], 331, 0, {}),
             ("starts in mid-row", img[1000:], 331, -(1000 % 331), {}),
             ("device-side count", img, 331, 0, {"indirect_count": 331 * 90 + 5}),
             ("wide and short", W.sponza_primary_rays(1024, 9), 1024, 0, {}),
             ("seven rows: strips", W.sponza_primary_rays(1024, 7), 0, 0, {}),
             ("63 columns: strips", W.sponza_primary_rays(63, 200), 0, 0, {}),
             ("shuffled rows: strips", img[rng.permutation(img.shape[0])], 0, 0, {}),
             ("normalised directions", _normalised(img), 331, 0, {}),
             ("orthographic camera", _orthographic(200, 120), 200, 0, {}),
             ("8K rows: second search round", W.sponza_primary_rays(7680, 9), 7680, 0, {})]
    for what, rays, want_w, want_base, kw in cases:
        strips, (w1, _) = run(rays, 1, **kw)
        tiles, (w0, b0) = run(rays, 0, **kw)
        assert w1 == 0 and (w0, b0) == (want_w, want_base), f"{what}: k_detect_grid said W = {w0}, row 0 starts at ray {b0}"
        assert np.array_equal(strips.view(np.uint8), tiles.view(np.uint8)), f"{what}: tiles and strips differ"
        # the per-ray kernel takes its 32-ray chunks as 8 x 4 half tiles of the same grid: any-hit ids, bit for bit
        a, _ = run(rays, 1, query=ANY, output=IDS, **kw)
        b, (w0, _) = run(rays, 0, query=ANY, output=IDS, **kw)
        assert w0 == want_w and np.array_equal(a, b), f"{what}: any-hit ids differ between 8 x 4 tiles and 32-ray chunks"
    # the client may state the row length (then row 0 starts at ray 0); a wrong one only costs coherence
    strips, _ = run(img, 1)
    for w in (331, 300, 4096):
        told, (w0, b0) = run(img, w)
        assert (w0, b0) == (w if (img.shape[0] + w - 1) // w >= 8 else 0, 0)
        assert np.array_equal(strips.view(np.uint8), told.view(np.uint8)), f"told W = {w}"
    assert ctx.lib.rrCudaSetOption(ctx.handle, api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 17) == api.RR_ERROR_INVALID_PARAMETER
    # tiles against the oracle (the parity tests above run with detection on as well: it is the default)
    assert_hits_equal(run(img, 0)[0], O.trace(g.nodes(), img), what="tiles", mesh=(pos, idx), rays=img)
    # incoherent batches are recognised beforehand (the packet kernel stands aside), camera rays are not mistaken for one
    assert _grid_words(engine, True)[2] == 0
    prim = W.sponza_primary_rays(256, 256)
    diffuse = W.diffuse_rays(pos, idx, prim, engine.intersect(g, prim), count=100_003)
    a, _ = run(diffuse, 1)
    b, _ = run(diffuse, 0)
    assert _grid_words(engine, True) == (0, 0, 1), "the diffuse batch should be found incoherent"
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert_hits_equal(b, O.trace(g.nodes(), diffuse), what="incoherent shortcut", mesh=(pos, idx), rays=diffuse)
    # jittered camera rays: consecutive steps may turn back anywhere, nothing to tile, same hits
    jit = img.copy()
    jit["direction"] += rng.normal(0, 2e-3, jit["direction"].shape).astype(np.float32)
    a, (w1, _) = run(jit, 1)
    b, (w0, _) = run(jit, 0)
    assert w0 == 0 and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_node_array_without_builder_tags(engine, sponza):
    """A VkBvhNode array that this builder did not write (a dump of the reference's, uploaded by the client) has small refit
    counters in its `update` words, not the builder's tag / leaf flags / child order (rr_internal.h node_update_word): the packet
    kernel must stand aside and the per-ray kernel must return the oracle's walk bit for bit, in tiles and in strips."""
    import torch
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    n_nodes = 2 * idx.shape[0] - 1
    words = g.d_nodes[: 64 * n_nodes].view(torch.int32).view(n_nodes, 16)
    saved = words[:, 15].clone()
    rays = W.sponza_primary_rays(320, 200)
    want = O.trace(g.nodes(), rays)
    try:
        for filler in (0, 2):                                  # what the reference leaves there: 0 after a build, 2 after an update
            words[:, 15] = filler
            for mode in (0, 1):
                engine.ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, mode)
                assert_hits_equal(engine.intersect(g, rays), want, what=f"untagged nodes, update = {filler}, grid mode {mode}")
            assert np.array_equal(engine.intersect(g, rays, ANY, IDS), O.trace(g.nodes(), rays, O.QUERY_ANY, O.OUTPUT_INSTANCE_ID))
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 0)
        words[:, 15] = saved
    # ... and a refit of such an array puts the builder's words back (the generic refit rewrites every parent it fits)
    words[: idx.shape[0] - 1, 15] = 0
    engine.update_geometry(g, pos)
    u = g.nodes()["update"][: idx.shape[0] - 1]
    assert np.all(u >> 16 == 0x52A5) and np.all(u & 1 == 0)
    assert_hits_equal(engine.intersect(g, rays), want, what="after refit", mesh=(pos, idx), rays=rays)


def test_degenerate_rays_and_triangles(engine, cornell):
    """Edge cases the arithmetic must survive exactly like the oracle: rays with zero / denormal / huge direction components
    (safe_invdir, common.h:166-183), axis-parallel rays, origins on a vertex / edge / face, max_t below min_t, infinite max_t,
    negative min_t, NaN and infinite components (those packets fail the coherence test and take the per-ray kernel), mixed with
    ordinary rays in the same 64-ray packets; zero-area and duplicated triangles in the mesh."""
    pos, idx, _ = cornell
    pos = np.concatenate([pos, np.array([[0.3, 0.5, 0.1], [0.3, 0.5, 0.1], [0.3, 0.5, 0.1], [0.1, 0.2, 0.0], [0.4, 0.8, 0.0]], np.float32)])
    nv = pos.shape[0]
    extra = np.array([[nv - 5, nv - 4, nv - 3],           # a point (zero area)
                      [nv - 2, nv - 1, nv - 1],           # a segment
                      idx[3], idx[3], idx[7]], np.uint32)  # duplicates of existing triangles: exact-t ties
    idx = np.concatenate([idx, extra]).astype(np.uint32)
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    want_nodes, _, _ = O.build_blas(pos, idx)
    assert_nodes_equal(nodes, want_nodes, what="mesh with degenerate triangles")
    base = W.cornell_primary_rays(64)                      # 4096 ordinary rays
    rays = base.copy()
    rng = np.random.default_rng(12)
    special = []
    def ray(o, d, tmin=0.001, tmax=1e5):
        r = np.zeros(1, W.RAY_DTYPE); r["origin"] = o; r["direction"] = d; r["min_t"] = tmin; r["max_t"] = tmax
        special.append(r)
    for d in ((0, 0, -1), (0, 0, 1), (1, 0, 0), (0, -1, 0), (1e-7, 1e-30, -1), (-0.0, 0.0, -1), (1e-5, -1e-5, -1), (3e38, 1, -1),
              (1e-40, 0, -1), (0, 0, 0), (0, 0, -1e-20)):
        ray((0.1, 1.0, 3.0), d)
    ray(pos[0], (0.3, 0.2, -1)); ray(0.5 * (pos[idx[0, 0]] + pos[idx[0, 1]]), (0, 0, -1)); ray(pos[idx[2]].mean(0), (0, 1, 0))
    ray((0, 1, 3.5), (0, 0, -1), tmin=5.0, tmax=1.0)            # empty interval
    ray((0, 1, 3.5), (0, 0, -1), tmin=0.0, tmax=np.inf)
    ray((0, 1, 0.0), (0, 0, -1), tmin=-10.0, tmax=10.0)         # negative min_t: hits behind the origin count
    ray((0, 1, 3.5), (0, 0, -1), tmin=0.0, tmax=0.0)
    ray((np.nan, 1, 3.5), (0, 0, -1)); ray((0, 1, 3.5), (np.nan, 0, -1)); ray((0, 1, 3.5), (0, np.inf, -1)); ray((np.inf, 1, 3.5), (0, 0, -1))
    ray((0, 1, 3.5), (0, 0, -1), tmin=np.nan); ray((0, 1, 3.5), (0, 0, -1), tmax=np.nan)
    special = np.concatenate(special)
    at = rng.choice(rays.shape[0], special.shape[0], replace=False)      # scattered among ordinary rays, inside their packets
    rays[at] = special
    for query in (CLOSEST, ANY):
        got = engine.intersect(g, rays, query, FULL)
        want = O.trace(nodes, rays, query, O.OUTPUT_FULL_HIT)
        # NaN uv of a hit compares equal bit for bit; ids must agree everywhere
        assert np.array_equal(got["inst_id"], want["inst_id"]), f"q={query}: hit / miss differs at {np.nonzero(got['inst_id'] != want['inst_id'])[0][:8]}"
        ok = want["inst_id"] != O.INVALID
        differ = ok & (got["prim_id"] != want["prim_id"])
        if query == CLOSEST and differ.any():              # exact-t ties between the duplicated triangles: brute force decides
            bf, _ = O.brute_force(pos, idx, rays[differ])
            assert np.array_equal(got["prim_id"][differ], bf["prim_id"])
        else:
            assert not differ.any()
        same = ok & ~differ
        assert np.array_equal(got["uv"][same].view(np.uint32), want["uv"][same].view(np.uint32))
    engine.ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 1)
    try:
        got = engine.intersect(g, rays)
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 0)
    want = O.trace(nodes, rays, tie=O.TIE_FIRST_FOUND)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), "first-found rule: bit for bit, degenerate rays included"
