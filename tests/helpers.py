"""Shared assertions for the parity tests (CUDA path vs the CPU oracle)."""
import numpy as np

from oracle import binding as O

FIELDS_EXACT = ("child0", "child1", "parent", "aabb0_min_or_v0", "aabb0_max_or_v1", "aabb1_min_or_v2", "aabb1_max_or_v3")


def assert_nodes_equal(got, want, what="nodes"):
    """Bit-exact comparison of every node field except `update` (rendezvous scratch state)."""
    assert got.shape == want.shape, (got.shape, want.shape)
    for f in FIELDS_EXACT:
        a = np.ascontiguousarray(got[f]).view(np.uint32)
        b = np.ascontiguousarray(want[f]).view(np.uint32)
        bad = np.nonzero((a != b).reshape(a.shape[0], -1).any(axis=1))[0]
        assert bad.size == 0, f"{what}: field {f} differs at {bad.size} nodes, first {bad[:5]}: {got[f][bad[:3]]} vs {want[f][bad[:3]]}"


def assert_hits_equal(got, want, rel_uv=1e-6, what="hits"):
    """ids bit-exact; uv within 1e-6 relative (BASELINE.json north_star) -- in practice they are bit-exact too."""
    assert got.shape == want.shape
    assert np.array_equal(got["inst_id"], want["inst_id"]), f"{what}: inst_id mismatch at {np.nonzero(got['inst_id'] != want['inst_id'])[0][:8]}"
    ok = want["inst_id"] != O.INVALID
    assert np.array_equal(got["prim_id"][ok], want["prim_id"][ok]), \
        f"{what}: prim_id mismatch at {np.nonzero(ok & (got['prim_id'] != want['prim_id']))[0][:8]}"
    a, b = got["uv"][ok].astype(np.float64), want["uv"][ok].astype(np.float64)
    assert np.all(np.abs(a - b) <= rel_uv * np.maximum(np.abs(b), 1e-30) + 1e-12), f"{what}: uv beyond 1e-6 relative"
    # untouched-on-miss semantics: uv / prim_id of missed rays keep whatever was there before (isect.comp:238-245)
    assert np.array_equal(got["prim_id"][~ok], want["prim_id"][~ok]), f"{what}: miss must not touch prim_id"
    assert np.array_equal(got["uv"][~ok].view(np.uint32), want["uv"][~ok].view(np.uint32)), f"{what}: miss must not touch uv"
