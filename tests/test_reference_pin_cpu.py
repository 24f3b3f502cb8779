"""The oracle pinned to values PRODUCED BY REFERENCE CODE RUN HERE.

The only part of RadeonRays that compiles in this image is its CPU tool bvh_analyzer (oracle/Makefile builds it from the
sources under /root/reference into oracle/_ref/).  It has no builder, but it holds the reference's triangle test
(bvh_analyzer/triangle.h:34-70 -- the same Moeller-Trumbore expression as vlk/kernels/common.h:103-137), its BVH2 validator
and SAH (bvh.h:130-223) and a BVH2 tracer (bvh.h:226-319).  These tests feed it the oracle's VkBvhNode dump and compare values:
hit/miss masks, primitive ids, closest t (bit for bit), uv, SAH.  tests/test_gpu_reference_pin.py repeats them on the GPU's dump
and the GPU's hits.  The prebuilt binaries travel to the GPU box; where they are absent the tests skip.
"""
import os

import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import workloads as W
from helpers import assert_matches_reference_tracer, pow2_scaled_rays

REF = os.path.join(os.path.dirname(O.__file__), "_ref")
needs_ref = pytest.mark.skipif(not (os.path.exists(os.path.join(REF, "bvh_analyzer_trace")) and os.path.exists(os.path.join(REF, "bvh_analyzer"))),
                               reason="oracle/_ref not built (needs /root/reference at build time)")


@needs_ref
def test_c1_cornell_1024x1024_matches_reference_values(cornell):
    """BASELINE config C1 at full size: 32 triangles, 1024 x 1024 primary rays."""
    pos, idx, _ = cornell
    nodes, _, _ = O.build_blas(pos, idx)
    rays, kept = pow2_scaled_rays(W.cornell_primary_rays(1024))
    assert kept.size > 1_040_000
    ref = O.ref_bvh_analyzer_trace(nodes, rays, want_brute=True)
    assert ref["is_valid"]
    hits = O.trace(nodes, rays)
    ties, single, same = assert_matches_reference_tracer(hits, ref, "C1 oracle")
    assert ties <= 64 and single > 700_000 and same > 1_000_000
    # closest t of the oracle's brute force == the reference triangle test's minimum, bit for bit
    _, t = O.brute_force(pos, idx, rays)
    ok = ref["brute"]["count"] > 0
    assert np.array_equal(t[ok].view(np.uint32), ref["brute"]["t"][ok].view(np.uint32))
    # power-of-two direction scaling does not change a single bit of the result (what makes the scaled set a fair stand-in)
    plain = O.trace(nodes, W.cornell_primary_rays(1024)[kept])
    assert np.array_equal(plain.view(np.uint8), hits.view(np.uint8))


@needs_ref
@pytest.mark.parametrize("restructure", [False, True])
def test_c2_sponza_sample_matches_reference_values(sponza, restructure):
    """BASELINE config C2 (Sponza, fast and treelet-restructured BVH): a 96 x 54 sample of the primary rays against the reference
    triangle test over all 262 267 triangles, and a 480 x 270 sample against the reference's BVH tracer."""
    pos, idx, _ = sponza
    nodes, _, _ = O.build_blas(pos, idx, restructure=restructure)
    rays, _ = pow2_scaled_rays(W.sponza_primary_rays(96, 54))
    ref = O.ref_bvh_analyzer_trace(nodes, rays, want_brute=True)
    assert ref["is_valid"]
    assert abs(ref["sah"] - O.sah(nodes)) / O.sah(nodes) < 0.02
    hits = O.trace(nodes, rays)
    ties, single, same = assert_matches_reference_tracer(hits, ref, "C2 oracle")
    assert ties <= 8   # (no Sponza primary ray crosses exactly one triangle: the atrium is closed and max_t is 1e5)
    _, t = O.brute_force(pos, idx, rays)
    ok = ref["brute"]["count"] > 0
    assert np.array_equal(t[ok].view(np.uint32), ref["brute"]["t"][ok].view(np.uint32))
    # larger sample, BVH tracer only (the brute-force pass is O(rays x triangles))
    rays, _ = pow2_scaled_rays(W.sponza_primary_rays(480, 270))
    ref = O.ref_bvh_analyzer_trace(nodes, rays, want_hits=True)
    hits = O.trace(nodes, rays)
    ours, theirs = hits["inst_id"] != O.INVALID, ref["hits"]["inst_id"] != O.INVALID
    assert np.count_nonzero(ours != theirs) <= 4
    both = ours & theirs
    assert (hits["prim_id"][both] == ref["hits"]["prim_id"][both]).mean() > 0.97   # the rest: several triangles accepted, ref keeps the last
    # the same rays clipped just behind their closest hit: (almost) every ray now crosses one triangle only, so the reference
    # tracer's "last accepted" IS the closest hit and the ids must agree ray by ray
    _, stats = O.trace(nodes, rays, want_stats=True)
    clipped = rays.copy()
    clipped["max_t"] = np.where(ours, stats["t"] * np.float32(1.0005), rays["max_t"])
    ref = O.ref_bvh_analyzer_trace(nodes, clipped, want_hits=True)
    hits = O.trace(nodes, clipped)
    ours, theirs = hits["inst_id"] != O.INVALID, ref["hits"]["inst_id"] != O.INVALID
    assert np.count_nonzero(ours != theirs) <= 4
    both = ours & theirs
    assert (hits["prim_id"][both] == ref["hits"]["prim_id"][both]).mean() > 0.995   # the rest: coplanar duplicates inside the 0.05 % clip margin


@needs_ref
def test_stock_bvh_analyzer_end_to_end(cornell, sponza):
    """The UNMODIFIED reference binary on the 7-line config (bvh_analyzer/config.h:46-63, main.cpp:56-88): accepts the dump, reports
    the SAH the oracle computes, and its traversal statistics are what the trace-loop driver reports for the same input."""
    for (pos, idx, _), rays, w, h in ((cornell, W.cornell_primary_rays(256), 256, 256), (sponza, W.sponza_primary_rays(160, 90), 160, 90)):
        nodes, _, _ = O.build_blas(pos, idx)
        rays, _ = pow2_scaled_rays(rays)
        rays = rays[: (rays.shape[0] // w) * w]
        res = O.ref_bvh_analyzer_stock(nodes, rays, w, rays.shape[0] // w)
        assert res["returncode"] == 0 and res["is_valid"] is True and res["wrote_jpegs"]
        assert abs(res["sah"] - O.sah(nodes)) / O.sah(nodes) < 0.02
        drv = O.ref_bvh_analyzer_trace(nodes, rays)
        assert abs(res["avg_primary_node_tests"] - drv["avg_node_tests"]) <= 1e-3 * drv["avg_node_tests"]
        assert abs(res["avg_primary_triangle_tests"] - drv["avg_tri_tests"]) <= 1e-3 * drv["avg_tri_tests"]
