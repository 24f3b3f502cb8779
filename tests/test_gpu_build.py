"""-m gpu: HLBVH build / refit / treelet parity of the CUDA path (through the rr* C ABI) against the oracle."""
import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from helpers import assert_hits_equal, assert_nodes_equal

pytestmark = pytest.mark.gpu


def _check_build(engine, pos, idx, flags=api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD):
    g = engine.build_geometry(pos, idx, build_flags=flags)
    n = idx.shape[0]
    restructure = flags is not None and not (flags & api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD)
    want, sc, sr = O.build_blas(pos, idx, restructure=restructure)
    L = engine.ctx.build_scratch_layout(n)
    if not restructure:  # treelet scratch aliases the build scratch
        lo, hi = O.scene_aabb(pos, idx)
        aabb = g.scratch_u32(L.scene_aabb_offset, 8)
        dec = lambda v: (v ^ (((v >> 31) - 1) | 0x80000000)).astype(np.uint32).view(np.float32)
        assert np.array_equal(dec(aabb[0:3]).view(np.uint32), lo.view(np.uint32))
        assert np.array_equal(dec(aabb[4:7]).view(np.uint32), hi.view(np.uint32))
        codes = g.scratch_u32(L.morton_codes_offset, n)
        assert np.array_equal(codes, O.morton_codes(pos, idx, lo, hi)), "Morton codes must be bit-exact"
        assert np.array_equal(g.scratch_u32(L.sorted_codes_offset, n), sc), "sorted codes"
        assert np.array_equal(g.d_nodes[L.sorted_refs_offset: L.sorted_refs_offset + 4 * n].cpu().numpy().view(np.uint32), sr), "sorted primitive order (stable)"
    got = g.nodes()
    assert_nodes_equal(got, want)
    assert O.check_consistency(got)
    return g, got


def test_single_triangle(engine):
    pos, idx = W.single_triangle()  # basic_test.h:283-292
    g, nodes = _check_build(engine, pos, idx)
    assert nodes.shape[0] == 1 and nodes["child0"][0] == O.INVALID and nodes["parent"][0] == O.INVALID


@pytest.mark.parametrize("n", [2, 3, 5, 31, 32, 33, 255, 1000])
def test_small_random_meshes(engine, n):
    rng = np.random.default_rng(n)
    pos = rng.random((3 * n, 3), dtype=np.float32)
    idx = rng.permutation(3 * n).astype(np.uint32).reshape(n, 3)
    _check_build(engine, pos, idx)


@pytest.mark.parametrize("n", [511, 512, 513, 544, 1023, 1025, 4097, 20000])
def test_sizes_around_group_and_window_borders(engine, n):
    """k_emit_leaves works on 32-leaf groups, k_emit_window on 512-leaf windows: sizes on either side of both."""
    rng = np.random.default_rng(n)
    pos = rng.random((n + 2, 3), dtype=np.float32)
    idx = np.stack([np.arange(n), np.arange(n) + 1, np.arange(n) + 2], 1).astype(np.uint32)
    _check_build(engine, pos, idx)


def _clustered_mesh(sizes, seed):
    """Tiny triangles in clusters that share a Morton cell: runs of equal codes of the given lengths (index tie-break
    subtrees, some longer than a 512-leaf window) next to isolated triangles -- deep, unbalanced hierarchies."""
    rng = np.random.default_rng(seed)
    centres = rng.random((len(sizes), 3)).astype(np.float32)
    pos, idx = [], []
    for c, k in zip(centres, sizes):
        for _ in range(k):
            b = len(pos)
            d = (rng.random((3, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(1e-5)
            pos.extend((c + d).astype(np.float32))
            idx.append((b, b + 1, b + 2))
    pos.append(np.zeros(3, np.float32)); pos.append(np.ones(3, np.float32))     # pin the scene box
    idx.append((len(pos) - 2, len(pos) - 1, len(pos) - 2))
    return np.asarray(pos, np.float32), np.asarray(idx, np.uint32)


@pytest.mark.parametrize("seed", [1, 2])
def test_clustered_codes_deep_hierarchy(engine, seed):
    rng = np.random.default_rng(100 + seed)
    sizes = [1, 1, 2, 3, 5, 8, 40, 100, 700, 1500, 33, 31, 32, 1, 513] + list(rng.integers(1, 60, 80))
    pos, idx = _clustered_mesh(sizes, seed)
    _check_build(engine, pos, idx)
    _check_build(engine, pos, idx, flags=0)


def test_octree_corner_chain(engine):
    """One triangle per octree level along the diagonal: the radix tree is a chain (depth ~ code bits), the case in which
    a group or window merges one node per pass."""
    pos, idx = [], []
    for lvl in range(1, 11):
        for rep in range(3 * lvl):                      # a few per level, spread so codes differ in the low bits
            c = np.float32(2.0 ** -lvl) * (np.float32(1.0) + np.float32(0.3) * np.float32(rep) / np.float32(3 * lvl))
            b = len(pos)
            pos.extend([np.array([c, c, c], np.float32), np.array([c, c, c], np.float32) * np.float32(1.0001), np.array([c, c * np.float32(1.0002), c], np.float32)])
            idx.append((b, b + 1, b + 2))
    pos.append(np.zeros(3, np.float32)); pos.append(np.ones(3, np.float32))
    idx.append((len(pos) - 2, len(pos) - 1, len(pos) - 2))
    pos, idx = np.asarray(pos, np.float32), np.asarray(idx, np.uint32)
    _check_build(engine, pos, idx)
    reps = np.tile(idx, (40, 1))                        # the same chain with every code repeated 40 times (6 640 triangles)
    _check_build(engine, pos, reps)


def test_cornell_box(engine, cornell):
    pos, idx, _ = cornell
    _check_build(engine, pos, idx)
    _check_build(engine, pos, idx, flags=None)           # build_options == NULL
    _check_build(engine, pos, idx, flags=0)               # quality build (no-op below 64 triangles)


def test_sponza_fast_build(engine, sponza):
    pos, idx, _ = sponza
    _check_build(engine, pos, idx)


def test_update_words_after_builds(engine, sponza, cornell):
    """Leaf flags / child order / tag in the update word after a fast build, a quality build and a Karras-tree refit."""
    for (pos, idx, _), flags in ((sponza, api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD), (sponza, 0), (cornell, 0)):
        g = engine.build_geometry(pos, idx, build_flags=flags)
        vote = flags != 0 or idx.shape[0] < 64          # (the treelet pass leaves trees below its 64-primitive floor alone)
        _check_update_words(g.nodes(), idx.shape[0], vote=vote)
        engine.update_geometry(g, (pos * np.float32(2.0)).astype(np.float32))
        _check_update_words(g.nodes(), idx.shape[0], vote=vote)


def test_sponza_quality_build_matches_oracle_treelets(engine, sponza):
    pos, idx, _ = sponza
    g, nodes = _check_build(engine, pos, idx, flags=0)
    fast, _, _ = O.build_blas(pos, idx)
    assert O.sah(nodes) <= O.sah(fast)                   # hlbvh_test.h:417


def test_duplicate_and_degenerate_inputs(engine):
    # all centroids identical (every Morton code equal -> index tie-break everywhere), flat extent (0/0 -> NaN -> 0)
    n = 777
    pos = np.zeros((3, 3), np.float32)
    pos[1] = (1, 0, 0)
    pos[2] = (0, 1, 0)                                     # z extent is zero
    idx = np.tile(np.array([[0, 1, 2]], np.uint32), (n, 1))
    _check_build(engine, pos, idx)
    # vertex stride larger than 12 bytes
    rng = np.random.default_rng(5)
    pos4 = rng.random((300, 4), dtype=np.float32)
    idx = rng.integers(0, 300, (500, 3)).astype(np.uint32)
    g = engine.build_geometry(pos4, idx, vertex_stride=16)
    want, _, _ = O.build_blas(pos4, idx)
    assert_nodes_equal(g.nodes(), want)


def test_uint16_indices(engine, cornell):
    """RR_INDEX_TYPE_UINT16 (beyond the reference, which reads 32-bit indices whatever index_type says):
    same BVH, refit and hits as the 32-bit mesh."""
    rng = np.random.default_rng(16)
    n_v, n = 60000, 70000                                  # > 65535 triangles, all vertex ids fit 16 bits; refit takes the staged path
    pos = rng.random((n_v, 3), dtype=np.float32)
    idx = rng.integers(0, n_v, (n, 3)).astype(np.uint32)
    g16 = engine.build_geometry(pos, idx.astype(np.uint16))
    want, _, _ = O.build_blas(pos, idx)
    assert_nodes_equal(g16.nodes(), want, what="uint16 build")
    moved = (pos * np.float32(0.5)).astype(np.float32)
    engine.update_geometry(g16, moved)
    assert_nodes_equal(g16.nodes(), O.refit(want, moved, idx), what="uint16 refit")
    cpos, cidx, _ = cornell
    g = engine.build_geometry(cpos, cidx.astype(np.uint16), build_flags=0)
    rays = W.cornell_primary_rays(64)
    ref = O.trace(O.build_blas(cpos, cidx, restructure=True)[0], rays)
    got = engine.intersect(g, rays)
    assert_hits_equal(got, ref, what="uint16 cornell", mesh=(cpos, cidx), rays=rays)   # (t, prim) minimum: tie-aware comparison


def test_vertex_buffer_at_4_byte_offset(engine):
    """A vertex pointer that is only 4-byte aligned (interop pointer + offset): the 8+4-byte vertex loads of k_emit_leaves /
    k_refit_leaves must fall back to scalar loads."""
    rng = np.random.default_rng(44)
    n_v, n = 30000, 600000                                 # large enough for the warp-cooperative refit path
    pos = rng.random((n_v, 3), dtype=np.float32)
    idx = rng.integers(0, n_v, (n, 3)).astype(np.uint32)
    g = engine.build_geometry(pos, idx, vertex_byte_offset=4)
    want, _, _ = O.build_blas(pos, idx)
    assert_nodes_equal(g.nodes(), want, what="offset build")
    moved = (pos + np.float32(0.25)).astype(np.float32)
    engine.update_geometry(g, moved)
    assert_nodes_equal(g.nodes(), O.refit(want, moved, idx), what="offset refit")


def test_update_refit(engine, sponza):
    """hlbvh_test.h:445-493 UpdateTest: move every vertex by +10 in y, UPDATE, topology untouched."""
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx)
    before = g.nodes()
    moved = pos.copy()
    moved[:, 1] += np.float32(10.0)
    for rep in range(2):                                   # twice: the parity rendezvous needs no reset pass
        engine.update_geometry(g, moved)
        got = g.nodes()
        want = O.refit(before, moved, idx)
        assert_nodes_equal(got, want, what=f"refit #{rep}")
        assert O.check_consistency(got)
        for f in ("child0", "child1", "parent"):
            assert np.array_equal(got[f], before[f])


def test_update_after_quality_build(engine, sponza):
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=0)
    before = g.nodes()
    moved = (pos * np.float32(1.5)).astype(np.float32)
    engine.update_geometry(g, moved)
    assert_nodes_equal(g.nodes(), O.refit(before, moved, idx))
    _check_update_words(g.nodes(), idx.shape[0])
    # ... and twice more (the rendezvous parity toggles; the builder's flags in the same word must survive), then traced by the
    # packet kernel, which relies on those flags
    for scale in (0.5, 1.25):
        moved = (pos * np.float32(scale)).astype(np.float32)
        engine.update_geometry(g, moved)
    nodes = g.nodes()
    assert_nodes_equal(nodes, O.refit(before, moved, idx))
    _check_update_words(nodes, idx.shape[0])
    rays = W.sponza_primary_rays(320, 200)
    rays["origin"] *= np.float32(1.25)
    assert_hits_equal(engine.intersect(g, rays), O.trace(nodes, rays), what="after refit", mesh=(moved, idx), rays=rays)


def _check_update_words(nodes, n, vote=False):
    """The builder's side of VkBvhNode::update (rr_internal.h node_update_word): tag, leaf flags that match the children, an even
    rendezvous parity after a complete refit, and a child order that is a function of one axis (a table of the form M or ~M)."""
    u = nodes["update"][: n - 1]
    assert np.all(u >> 16 == 0x52A5), "tag"
    assert np.all(u & 1 == 0), "rendezvous parity must be even between refits"
    assert np.all((u >> 3) & 1 == int(vote)), "vote-order flag: set by fast builds (plain LBVH), clear on treelet-optimised trees"
    assert np.array_equal((u >> 1) & 1, (nodes["child0"][: n - 1] >= n - 1).astype(np.uint32)), "child0-is-leaf flag"
    assert np.array_equal((u >> 2) & 1, (nodes["child1"][: n - 1] >= n - 1).astype(np.uint32)), "child1-is-leaf flag"
    order = (u >> 8) & 0xFF
    if vote:
        assert np.all(order == 0), "no order table where the traversal votes"
        return
    assert np.all(np.isin(order, [0xAA, 0x55, 0xCC, 0x33, 0xF0, 0x0F])), "child order table"
    # the order follows the boxes: child1 first for the octants that look from its side
    c = (nodes["aabb1_min_or_v2"][: n - 1] + nodes["aabb1_max_or_v3"][: n - 1]) - (nodes["aabb0_min_or_v0"][: n - 1] + nodes["aabb0_max_or_v1"][: n - 1])
    axis = np.where((np.abs(c[:, 0]) >= np.abs(c[:, 1])) & (np.abs(c[:, 0]) >= np.abs(c[:, 2])), 0, np.where(np.abs(c[:, 1]) >= np.abs(c[:, 2]), 1, 2))
    neg = np.array([0xAA, 0xCC, 0xF0], np.uint32)[axis]
    want = np.where(c[np.arange(n - 1), axis] < 0, ~neg & 0xFF, neg)
    assert np.array_equal(order, want), "child order does not follow the boxes"


def test_update_after_debug_restructure(engine, sponza):
    """hlbvh_test.h:350-420 drives RestructureHlBvh on an already built BVH; an update afterwards must take the generic refit
    (the tree is no longer the Karras tree a re-emission would produce)."""
    import torch
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx)                       # fast build: Karras-numbered, tail header = 1
    ctx = engine.ctx
    scratch = torch.empty(g.req.temporary_build_buffer_size + 64 * idx.shape[0], dtype=torch.uint8, device=engine.device)
    api.check(ctx.lib.rrCudaDebugRestructure(ctx.handle, g.p_nodes, idx.shape[0], ctx.tensor_ptr(scratch)))
    want, _, _ = O.build_blas(pos, idx, restructure=True)
    before = g.nodes()
    assert_nodes_equal(before, want, what="restructured")
    moved = (pos * np.float32(0.75)).astype(np.float32)
    engine.update_geometry(g, moved)
    assert_nodes_equal(g.nodes(), O.refit(before, moved, idx), what="refit after restructure")


def test_heightfield_1m(engine):
    """A 1 M-triangle slice of config C5's animated height field: rebuild and refit agree with the oracle."""
    pos, idx = W.heightfield_mesh(1000, 500, t=0.0)
    g, nodes = _check_build(engine, pos, idx)
    pos2, _ = W.heightfield_mesh(1000, 500, t=1.0)
    engine.update_geometry(g, pos2)
    assert_nodes_equal(g.nodes(), O.refit(nodes, pos2, idx))


def test_sort_pairs_stable(engine):
    """algos_test.h:322-390 SortTest, plus the value/stability check the reference omits."""
    import torch
    rng = np.random.default_rng(0)
    for n, hi in ((1, 10), (1000, 16), (8192, 1000), (8193, 1 << 30), (1 << 20, 1000), (3_000_001, 0xFFFFFFFF)):
        keys = rng.integers(0, hi, n, dtype=np.uint64).astype(np.uint32)
        vals = rng.integers(0, 0xFFFFFFFF, n, dtype=np.uint64).astype(np.uint32)
        dk, dv = torch.from_numpy(keys.view(np.int32)).cuda(), torch.from_numpy(vals.view(np.int32)).cuda()
        ok, ov = torch.empty_like(dk), torch.empty_like(dv)
        api.check(engine.ctx.lib.rrCudaDebugSortPairs(engine.ctx.handle, dk.data_ptr(), dv.data_ptr(), ok.data_ptr(), ov.data_ptr(), n))
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ok.cpu().numpy().view(np.uint32), keys[order])
        assert np.array_equal(ov.cpu().numpy().view(np.uint32), vals[order])
        api.check(engine.ctx.lib.rrCudaDebugSortPairs(engine.ctx.handle, dk.data_ptr(), None, ok.data_ptr(), ov.data_ptr(), n))
        assert np.array_equal(ov.cpu().numpy().view(np.uint32), order.astype(np.uint32))


@pytest.mark.parametrize("shape", ["sponza", "heightfield_540k"])
def test_refit_hand_over_lists_never_drop_a_subtree(engine, sponza, shape):
    """the staged refit hands finished subtrees from stage to stage through lists sized for the common
    case; an entry past the capacity used to be discarded, leaving its ancestors stale.  Every hand-over now falls back to
    finishing the climb in place.  RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY shrinks the lists to 64 entries so that almost every
    hand-over overflows, on a treelet-restructured tree, for both the small-mesh stages (Sponza) and the warp-cooperative stages
    (>= 500 000 triangles): the refitted tree must still equal the oracle's bit for bit."""
    if shape == "sponza":
        pos, idx, _ = sponza
        moved = (pos * np.float32(1.25)).astype(np.float32)
    else:
        pos, idx = W.heightfield_mesh(600, 450, t=0.0)
        moved, _ = W.heightfield_mesh(600, 450, t=1.0)
    g = engine.build_geometry(pos, idx, build_flags=0)          # quality build: restructured, generic refit path
    before = g.nodes()
    engine.ctx.set_option(api.RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY, 64)
    try:
        engine.update_geometry(g, moved)
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY, 0)
    assert_nodes_equal(g.nodes(), O.refit(before, moved, idx), what="refit with overflowing hand-over lists")
    engine.update_geometry(g, pos)                              # and back, with the default capacities
    assert_nodes_equal(g.nodes(), O.refit(before, pos, idx), what="refit back")


@pytest.mark.parametrize("mesh", ["cornell", "sponza", "duplicates", "heightfield_300k"])
def test_morton63_build(engine, sponza, cornell, mesh):
    """RR_CUDA_OPTION_MORTON_BITS = 63 (BASELINE north_star "30/63-bit Morton codes"; an extension, the reference ships 30-bit codes
    only): sorted 63-bit codes, sorted primitive order and every node bit-equal to the oracle's rro_build_blas63; a refit afterwards
    re-emits from the same deltas; the quality build restructures it like any other tree; traces agree with the oracle."""
    import torch
    if mesh == "cornell":
        pos, idx, _ = cornell
    elif mesh == "sponza":
        pos, idx, _ = sponza
    elif mesh == "duplicates":       # every triangle twice + a cluster of identical centroids: equal 63-bit codes, index tie-break
        rng = np.random.default_rng(4)
        base = rng.random((3000, 3), dtype=np.float32)
        tri = rng.integers(0, 3000, (5000, 3)).astype(np.uint32)
        pos, idx = base, np.concatenate([tri, tri, np.repeat(tri[:1], 700, 0)]).astype(np.uint32)
    else:
        pos, idx = W.heightfield_mesh(500, 300, t=0.0)
    n = idx.shape[0]
    ctx = engine.ctx
    ctx.set_option(api.RR_CUDA_OPTION_MORTON_BITS, 63)
    try:
        g = engine.build_geometry(pos, idx)                       # fast build
        L = ctx.build_scratch_layout(n)
        codes = g.d_temp[L.sorted_codes_offset: L.sorted_codes_offset + 8 * n].cpu().numpy().view(np.uint64)
        refs = g.d_nodes[L.sorted_refs_offset: L.sorted_refs_offset + 4 * n].cpu().numpy().view(np.uint32)
        want, wc, wr = O.build_blas63(pos, idx)
        assert np.array_equal(codes, wc), "sorted 63-bit Morton codes"
        assert np.array_equal(refs, wr), "sorted primitive order"
        assert_nodes_equal(g.nodes(), want, what="63-bit build")
        assert O.check_consistency(g.nodes())
        moved = (pos * np.float32(1.1) + np.float32(0.3)).astype(np.float32)
        engine.update_geometry(g, moved)
        assert_nodes_equal(g.nodes(), O.refit(want, moved, idx), what="refit of a 63-bit build")
        if mesh in ("sponza", "duplicates"):
            gq = engine.build_geometry(pos, idx, build_flags=0)   # quality build on top of the 63-bit tree
            wq, _, _ = O.build_blas63(pos, idx, restructure=True)
            assert_nodes_equal(gq.nodes(), wq, what="63-bit quality build")
    finally:
        ctx.set_option(api.RR_CUDA_OPTION_MORTON_BITS, 30)
    if mesh == "sponza":
        rays = W.sponza_primary_rays(200, 120)
        engine.update_geometry(g, pos)
        assert_hits_equal(engine.intersect(g, rays), O.trace(want, rays), what="trace of a 63-bit build", mesh=(pos, idx), rays=rays)
        # and the option is back at 30: a fresh build equals the reference-order tree again
        g30 = engine.build_geometry(pos, idx)
        assert_nodes_equal(g30.nodes(), O.build_blas(pos, idx)[0], what="30-bit build after the option was reset")


def test_multi_mesh_geometry(engine, sponza, cornell):
    """RRGeometryBuildInput::primitive_count > 1 (the reference asserts it out, vlk/intersector.cpp:110): three
    meshes with their own vertex buffers, strides and index types in ONE geometry.  The BLAS must equal the oracle's build of the
    concatenated mesh, prim_id is the running triangle index, and UPDATE refits from the (moved) separate buffers."""
    import torch
    from radeonrays_sdk_b200.host import Geometry, _dev_bytes, _upload
    pos, idx, first = sponza
    cpos, cidx, _ = cornell
    a_idx = idx[first[73]:first[74]]                                             # 13 988 triangles, u32 indices, stride 12
    b_pos = (cpos * np.float32(20.0) + np.float32([0, 0, 30])).astype(np.float32)  # Cornell box, u16 indices, stride 16
    b_pad = np.zeros((b_pos.shape[0], 4), np.float32); b_pad[:, :3] = b_pos; b_pad[:, 3] = 7.0
    c_pos, c_idx = W.single_triangle()
    c_pos = (c_pos * np.float32(50.0)).astype(np.float32)
    ctx, dev = engine.ctx, engine.device

    def make(posa, posb_pad, posc):
        bufs = [_upload(posa, dev), _upload(a_idx, dev), _upload(posb_pad, dev), _upload(cidx.astype(np.uint16), dev), _upload(posc, dev), _upload(c_idx, dev)]
        p = [ctx.tensor_ptr(t) for t in bufs]
        gi = ctx.geometry_input_multi([(p[0], posa.shape[0], 12, p[1], a_idx.shape[0], api.RR_INDEX_TYPE_UINT32),
                                       (p[2], posb_pad.shape[0], 16, p[3], cidx.shape[0], api.RR_INDEX_TYPE_UINT16),
                                       (p[4], posc.shape[0], 12, p[5], c_idx.shape[0], api.RR_INDEX_TYPE_UINT32)])
        return gi, bufs

    def concat(posa, posb, posc):
        allpos = np.concatenate([posa, posb, posc]).astype(np.float32)
        allidx = np.concatenate([a_idx, cidx + np.uint32(posa.shape[0]), c_idx + np.uint32(posa.shape[0] + posb.shape[0])]).astype(np.uint32)
        return allpos, allidx

    for flags in (api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, 0):
        gi, bufs = make(pos, b_pad, c_pos)
        opts = api.RRBuildOptions(flags, None)
        req = ctx.geometry_requirements(gi, opts)
        g = Geometry()
        g.engine = engine
        g.d_temp, g.d_nodes = _dev_bytes(max(req.temporary_build_buffer_size, req.temporary_update_buffer_size), dev), _dev_bytes(req.result_buffer_size, dev)
        g.p_temp, g.p_nodes = ctx.tensor_ptr(g.d_temp), ctx.tensor_ptr(g.d_nodes)
        ctx.run(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, gi, opts, g.p_temp, g.p_nodes, s))
        allpos, allidx = concat(pos, b_pos, c_pos)
        g.triangle_count = allidx.shape[0]
        want, _, _ = O.build_blas(allpos, allidx, restructure=(flags == 0))
        assert_nodes_equal(g.nodes(), want, what=f"multi-mesh build flags={flags}")
        rays = W.random_rays(50_000, allpos.min(0) - 5, allpos.max(0) + 5, seed=23)
        got = engine.intersect(g, rays)
        assert_hits_equal(got, O.trace(want, rays), what="multi-mesh trace", mesh=(allpos, allidx), rays=rays)
        hit = got["inst_id"] != O.INVALID
        assert (got["prim_id"][hit] >= a_idx.shape[0]).sum() > 50, "hits on the second / third mesh carry running triangle indices"
        # UPDATE from moved separate buffers
        moved_a = (pos + np.float32(0.5)).astype(np.float32)
        moved_b = b_pad.copy(); moved_b[:, :3] *= np.float32(1.1)
        moved_c = (c_pos * np.float32(0.9)).astype(np.float32)
        bufs[0].copy_(_upload(moved_a, dev)); bufs[2].copy_(_upload(moved_b, dev)); bufs[4].copy_(_upload(moved_c, dev))
        ctx.run(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, gi, opts, g.p_temp, g.p_nodes, s))
        mpos, _ = concat(moved_a, moved_b[:, :3], moved_c)
        assert_nodes_equal(g.nodes(), O.refit(want, mpos, allidx), what=f"multi-mesh refit flags={flags}")
    # more meshes than the backend merges, and a non-triangle primitive type: NOT_IMPLEMENTED like the reference
    many = ctx.geometry_input_multi([(ctx.tensor_ptr(bufs[4]), 3, 12, ctx.tensor_ptr(bufs[5]), 1, api.RR_INDEX_TYPE_UINT32)] * 17)
    req = api.RRMemoryRequirements()
    assert ctx.lib.rrGetGeometryBuildMemoryRequirements(ctx.handle, many, None, req) == api.RR_ERROR_NOT_IMPLEMENTED
