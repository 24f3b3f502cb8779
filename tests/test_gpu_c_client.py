"""-m gpu: the reference's API-level test scenarios (test/test_vk/basic_test.h, internal_resources_test.h) driven from a
C++ client (tests/c/test_rr_c_api.cpp) that links libradeonrays_b200.so like any RadeonRays user would; this wrapper
feeds it the Sponza fixture and checks every hit buffer it dumps against the CPU oracle -- the value-level assertions
the reference's own tests do not make."""
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import workloads as W
from helpers import assert_hits_equal, assert_nodes_equal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "c", "test_rr_c_api")
RES = 320


def _rays(res, dy=0.0):
    """tests/c/test_rr_c_api.cpp sponza_rays(): the C++ evaluation `-1.f + (2.f / res) * y` in binary32."""
    r = W.sponza_primary_rays(res, res)
    r["origin"][:, 1] += np.float32(dy)
    return r


@pytest.fixture(scope="module")
def dumps(tmp_path_factory, sponza):
    if not os.path.exists(EXE):
        pytest.fail(f"{EXE} is missing: run __graft_entry__.build()")
    d = tmp_path_factory.mktemp("c_client")
    pos, idx, first = sponza
    pos.tofile(d / "positions.bin"); idx.tofile(d / "indices.bin"); first.tofile(d / "shapes.bin")
    out = subprocess.run([EXE, str(d), str(RES)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    for name in ("CreateContext", "BuildSingleTriangle", "BuildObj", "UpdateObj", "BuildObj2Level", "InternalResources"):
        assert f"[ OK ] {name}" in out.stdout
    return d


def _load(d, name, dtype):
    return np.fromfile(d / name, dtype=dtype)


def test_build_obj_matches_oracle(dumps, sponza):
    pos, idx, _ = sponza
    rays = _rays(RES)
    for tag, restructure in (("obj_fast", False), ("obj_quality", True)):
        want_nodes, _, _ = O.build_blas(pos, idx, restructure=restructure)
        assert_nodes_equal(_load(dumps, tag + ".nodes", W.NODE_DTYPE), want_nodes, what=tag)
        assert_hits_equal(_load(dumps, tag + ".hits", W.HIT_DTYPE), O.trace(want_nodes, rays, init=np.zeros(rays.shape[0], W.HIT_DTYPE)), what=tag,
                          mesh=(pos, idx), rays=rays)
        assert_hits_equal(_load(dumps, tag + ".ids", np.uint32), O.trace(want_nodes, rays, O.QUERY_CLOSEST, O.OUTPUT_INSTANCE_ID), what=tag + " ids",
                          mesh=(pos, idx), rays=rays)
        assert_hits_equal(_load(dumps, tag + "_any.hits", W.HIT_DTYPE),
                          O.trace(want_nodes, rays, O.QUERY_ANY, O.OUTPUT_FULL_HIT, init=np.zeros(rays.shape[0], W.HIT_DTYPE)), what=tag + " any")


def test_update_obj_matches_oracle(dumps, sponza):
    pos, idx, _ = sponza
    before, _, _ = O.build_blas(pos, idx)
    moved = pos.copy()
    moved[:, 1] -= np.float32(40.0)
    want = O.refit(before, moved, idx)
    assert_nodes_equal(_load(dumps, "update.nodes", W.NODE_DTYPE), want, what="update")
    rays = _rays(RES, dy=-40.0)
    assert_hits_equal(_load(dumps, "update.hits", W.HIT_DTYPE), O.trace(want, rays, init=np.zeros(rays.shape[0], W.HIT_DTYPE)), what="update",
                      mesh=(moved, idx), rays=rays)


def test_two_level_matches_oracle(dumps, sponza):
    pos, idx, first = sponza
    blas = [O.build_blas(pos, idx[first[s]:first[s + 1]])[0] for s in range(len(first) - 1)]
    n = len(blas)
    xf = np.zeros((n, 3, 4), np.float32)
    xf[:, 0, 0] = xf[:, 1, 1] = xf[:, 2, 2] = 1
    tlas, out_xf = O.build_tlas(blas, list(range(n)), xf)
    rays = _rays(RES)
    want = O.trace_2l(tlas, out_xf, blas, list(range(n)), rays, init=np.zeros(rays.shape[0], W.HIT_DTYPE))
    assert_hits_equal(_load(dumps, "two_level.hits", W.HIT_DTYPE), want, what="two level")
    assert np.array_equal(_load(dumps, "two_level.ids", np.uint32),
                          O.trace_2l(tlas, out_xf, blas, list(range(n)), rays, O.QUERY_CLOSEST, O.OUTPUT_INSTANCE_ID))


def test_internal_resources_matches_oracle(dumps, sponza):
    pos, idx, _ = sponza
    nodes, _, _ = O.build_blas(pos, idx, restructure=True)       # build_flags = 0
    rays = _rays(RES)
    # the library-owned hit buffer is not cleared by the client: compare hits only (a miss leaves uv / prim_id untouched)
    got = _load(dumps, "internal.hits", W.HIT_DTYPE)
    want = O.trace(nodes, rays)
    ok = (want["inst_id"] != O.INVALID) & (got["inst_id"] != O.INVALID)
    got, want = got.copy(), want.copy()
    for f in ("uv", "prim_id"):
        got[f][~ok] = 0
        want[f][~ok] = 0
    assert_hits_equal(got, want, what="internal resources", mesh=(pos, idx), rays=rays)
