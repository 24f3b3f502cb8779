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
