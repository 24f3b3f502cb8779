"""Shared assertions for the parity tests (CUDA path vs the CPU oracle)."""
import numpy as np

from oracle import binding as O

HIT_DTYPE_ = O.HIT_DTYPE

FIELDS_EXACT = ("child0", "child1", "parent", "aabb0_min_or_v0", "aabb0_max_or_v1", "aabb1_min_or_v2", "aabb1_max_or_v3")


def assert_nodes_equal(got, want, what="nodes"):
    """Bit-exact comparison of every node field except `update` (rendezvous scratch state)."""
    assert got.shape == want.shape, (got.shape, want.shape)
    for f in FIELDS_EXACT:
        a = np.ascontiguousarray(got[f]).view(np.uint32)
        b = np.ascontiguousarray(want[f]).view(np.uint32)
        bad = np.nonzero((a != b).reshape(a.shape[0], -1).any(axis=1))[0]
        assert bad.size == 0, f"{what}: field {f} differs at {bad.size} nodes, first {bad[:5]}: {got[f][bad[:3]]} vs {want[f][bad[:3]]}"


def assert_hits_equal(got, want, rel_uv=1e-6, what="hits", mesh=None, rays=None):
    """ids bit-exact; uv within 1e-6 relative (BASELINE.json north_star) -- in practice they are bit-exact too.
    mesh=(positions, indices) and rays: the query was CLOSEST under the default (t, prim) rule, which the packet traversal
    answers -- rays that differ from the BVH2 walk must then carry the brute-force minimum (assert_closest_hits_equal)."""
    if mesh is not None:
        return assert_closest_hits_equal(got, want, mesh, rays, what=what)
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


def pow2_scaled_rays(rays, max_shift=12):
    """See radeonrays_sdk_b200.workloads.pow2_scaled_rays (every |direction component| >= 1, results unchanged bit for bit)."""
    from radeonrays_sdk_b200.workloads import pow2_scaled_rays as f
    r, kept = f(rays, max_shift)
    assert (np.abs(r["direction"]).min(1) >= 1).all()
    return r, kept


def assert_matches_reference_tracer(hits, ref, what, max_mask_mismatch=4):
    """`hits` (ours: oracle or GPU, closest-hit FULL_HIT) against what the REFERENCE's bvh_analyzer produced on the same node
    dump and rays (oracle.binding.ref_bvh_analyzer_trace(want_hits=True, want_brute=True)).

    * ref["brute"]: the reference's Triangle::Intersect (triangle.h:34-70) over every triangle -- the set of accepted
      triangles is traversal independent, so: hit/miss mask equal (up to `max_mask_mismatch` grazing rays lost by the slab
      test); prim id equal on every ray we hit that has exactly one accepted triangle; prim id equal to the reference's (t, prim) minimum on all rays but the exact-t ties the BVH walk
      resolves by visit order (counted, returned).
    * ref["hits"]: BvhIntersect<2> (bvh.h:226-319) keeps the LAST accepted triangle in its own visit order and tests its boxes
      with an unfused multiply-add, so: mask equal up to `max_mask_mismatch` grazing rays, prim equal wherever exactly one
      triangle is accepted, and its Moeller-Trumbore (u, v) close to our recomputed barycentrics when the prim agrees."""
    brute, rh = ref["brute"], ref["hits"]
    ours_hit = hits["inst_id"] != O.INVALID
    # a hit of ours is always a triangle the reference test accepts; the converse may fail on a few grazing rays whose triangle
    # lies in a face of a flat box the slab test (common.h:150-164) rejects by one ulp -- the reference's own BVH walk loses them too
    assert not np.any(ours_hit & (brute["count"] == 0)), f"{what}: hit on a ray the reference triangle test rejects everywhere"
    lost = np.count_nonzero(~ours_hit & (brute["count"] > 0))
    assert lost <= max_mask_mismatch, f"{what}: {lost} rays miss although the reference triangle test accepts a triangle"
    one = ours_hit & (brute["count"] == 1)
    assert np.array_equal(hits["prim_id"][one], brute["prim_id"][one]), f"{what}: prim id differs on single-intersection rays"
    differs = ours_hit & (hits["prim_id"] != brute["prim_id"])
    ref_hit = rh["inst_id"] != O.INVALID
    assert np.count_nonzero(ref_hit != ours_hit) <= max_mask_mismatch, f"{what}: mask vs BvhIntersect differs on {np.count_nonzero(ref_hit != ours_hit)} rays"
    both = one & ref_hit
    assert np.array_equal(hits["prim_id"][both], rh["prim_id"][both]), f"{what}: prim id differs from BvhIntersect on single-intersection rays"
    same = ours_hit & ref_hit & (hits["prim_id"] == rh["prim_id"])
    duv = np.abs(hits["uv"][same].astype(np.float64) - rh["uv"][same].astype(np.float64))
    assert duv.size and np.median(duv) < 1e-5 and duv.max() < 5e-2, f"{what}: uv vs reference (u, v): median {np.median(duv)}, max {duv.max()}"
    return int(np.count_nonzero(differs)), int(np.count_nonzero(one)), int(np.count_nonzero(same))


def assert_closest_hits_equal(got, want, mesh, rays, what="closest hits", max_fraction=2e-4):
    """Closest-hit parity under the default (t, prim) tie rule, for the packet traversal (rr_trace.cu k_trace_packet).

    `want` is the oracle's BVH2 walk (isect.comp's visit order).  The result of a closest-hit query is the (t, prim) minimum over
    the triangles the ray accepts, which does not depend on the visit order -- except where the slab test, by an ulp, hides a
    triangle from one walk and not from the other (exact-t ties across a culled subtree, rays grazing a flat box).  Those rays
    must then carry the order-INDEPENDENT answer, the brute-force (t, prim) minimum over all triangles (same triangle arithmetic),
    and there may only be a handful of them.  Everything else is compared bit for bit as before.  Returns the number of such rays."""
    pos, idx = mesh
    full = got.dtype == HIT_DTYPE_
    gi = got["inst_id"] if full else got
    wi = want["inst_id"] if full else want
    differ = gi != wi
    if full:
        differ |= (wi != O.INVALID) & (got["prim_id"] != want["prim_id"])
    n = int(np.count_nonzero(differ))
    if n:
        assert n <= max(2, int(max_fraction * got.shape[0])), f"{what}: {n} of {got.shape[0]} rays differ from the BVH2 walk"
        bf, _ = O.brute_force(pos, idx, rays[differ])
        if full:
            assert np.array_equal(got["inst_id"][differ], bf["inst_id"]) and np.array_equal(got["prim_id"][differ], bf["prim_id"]), \
                f"{what}: a ray that differs from the BVH2 walk does not carry the brute-force (t, prim) minimum"
            a, b = got["uv"][differ].astype(np.float64), bf["uv"].astype(np.float64)
            assert np.all(np.abs(a - b) <= 1e-6 * np.maximum(np.abs(b), 1e-30) + 1e-12), f"{what}: uv of a tie ray"
        else:
            assert np.array_equal(got[differ], np.where(bf["inst_id"] != O.INVALID, bf["prim_id"], O.INVALID)), \
                f"{what}: a ray that differs from the BVH2 walk does not carry the brute-force (t, prim) minimum"
    same = ~differ
    if full:
        assert_hits_equal(got[same], want[same], what=what)
    else:
        assert np.array_equal(got[same], want[same]), what
    return n
