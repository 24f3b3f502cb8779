"""-m gpu: the CUDA path against values produced by REFERENCE code (the prebuilt oracle/_ref binaries, compiled from
/root/reference/bvh_analyzer by oracle/Makefile; they travel to the GPU box, /root/reference itself does not).

The GPU's geometry buffer is dumped as VkBvhNode[] (bvh_analyzer/transform.h:31-41), validated and traced by the reference tool,
and the hits rrCmdIntersect returned through the C ABI are compared with the reference's: see helpers.assert_matches_reference_tracer."""
import os

import numpy as np
import pytest

from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from helpers import assert_matches_reference_tracer, pow2_scaled_rays

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(O.__file__), "_ref")
needs_ref = pytest.mark.skipif(not (os.path.exists(os.path.join(REF, "bvh_analyzer_trace")) and os.path.exists(os.path.join(REF, "bvh_analyzer"))),
                               reason="oracle/_ref not built")


@needs_ref
def test_c1_cornell_1024x1024_gpu_vs_reference(engine, cornell):
    pos, idx, _ = cornell
    g = engine.build_geometry(pos, idx)
    nodes = g.nodes()
    plain = W.cornell_primary_rays(1024)
    rays, kept = pow2_scaled_rays(plain)
    ref = O.ref_bvh_analyzer_trace(nodes, rays, want_brute=True)
    assert ref["is_valid"]
    hits = engine.intersect(g, rays)
    ties, single, same = assert_matches_reference_tracer(hits, ref, "C1 gpu")
    assert ties <= 64 and single > 700_000 and same > 1_000_000
    # the unscaled C1 batch gives the same bits (power-of-two scaling is exact) -- with the same grouping of rays into packets:
    # the scaled directions hide the image's rows from k_detect_grid, the plain ones do not, and on Cornell's flat boxes the rays
    # that tie at the ulp level resolve by packet membership (helpers.assert_closest_hits_equal); so strips for both
    engine.ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 1)
    try:
        a, b = engine.intersect(g, plain[kept]), engine.intersect(g, rays)
    finally:
        engine.ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 0)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    ties2, _, _ = assert_matches_reference_tracer(engine.intersect(g, plain[kept]), ref, "C1 gpu, tiles")
    assert ties2 <= 64
    # stock binary, 7-line config, on the GPU's dump
    w = 1024
    rows = rays[: (rays.shape[0] // w) * w]
    res = O.ref_bvh_analyzer_stock(nodes, rows, w, rows.shape[0] // w)
    assert res["returncode"] == 0 and res["is_valid"] is True
    assert abs(res["sah"] - O.sah(nodes)) / O.sah(nodes) < 0.02


@needs_ref
@pytest.mark.parametrize("flags", [api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, 0])
def test_c2_sponza_gpu_vs_reference(engine, sponza, flags):
    pos, idx, _ = sponza
    g = engine.build_geometry(pos, idx, build_flags=flags)
    nodes = g.nodes()
    rays, _ = pow2_scaled_rays(W.sponza_primary_rays(96, 54))
    ref = O.ref_bvh_analyzer_trace(nodes, rays, want_brute=True)
    assert ref["is_valid"]
    hits = engine.intersect(g, rays)
    ties, _, _ = assert_matches_reference_tracer(hits, ref, "C2 gpu")
    assert ties <= 8
    # 480 x 270 sample clipped just behind the closest hit: the reference tracer's last accepted triangle is then the closest one
    rays, _ = pow2_scaled_rays(W.sponza_primary_rays(480, 270))
    _, stats = O.trace(nodes, rays, want_stats=True)
    clipped = rays.copy()
    clipped["max_t"] = np.where(stats["t"] < rays["max_t"], stats["t"] * np.float32(1.0005), rays["max_t"])
    ref = O.ref_bvh_analyzer_trace(nodes, clipped, want_hits=True)
    hits = engine.intersect(g, clipped)
    ours, theirs = hits["inst_id"] != O.INVALID, ref["hits"]["inst_id"] != O.INVALID
    assert np.count_nonzero(ours != theirs) <= 4
    both = ours & theirs
    assert (hits["prim_id"][both] == ref["hits"]["prim_id"][both]).mean() > 0.995
