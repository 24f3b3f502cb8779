"""Small driver for ncu: build Sponza, then launch the requested trace flavour a few times.
  python tools/profile_trace.py [--bvh quality|fast] [--query closest|any] [--rays primary|diffuse|shadow] [--reps 4]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from radeonrays_sdk_b200 import api, workloads as W
from radeonrays_sdk_b200.host import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--bvh", default="quality")
ap.add_argument("--query", default="closest")
ap.add_argument("--rays", default="primary")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--presort", default="", help="experiment: reorder the rays on the host first: octant | octant+cell")
a = ap.parse_args()
eng = Engine(0)
ctx = eng.ctx
pos, idx, _ = W.load_mesh("sponza")
g = eng.build_geometry(pos, idx, build_flags=0 if a.bvh == "quality" else api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD)
rays = W.sponza_primary_rays(a.width, a.height)
if a.rays != "primary":
    prim = W.sponza_primary_rays(1024, 1024)
    hits = eng.intersect(g, prim)
    rays = (W.diffuse_rays if a.rays == "diffuse" else W.shadow_rays)(pos, idx, prim, hits, count=1 << 24)
if a.presort:
    f = rays.view(np.float32).reshape(-1, 8)
    o, d = f[:, 0:3], f[:, 4:7]
    octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2))
    key = octant << 40
    if "cell" in a.presort:
        lo, hi = pos.min(0), pos.max(0)
        q = np.clip(((o - lo) / (hi - lo) * 64).astype(np.int64), 0, 63).astype(np.uint64)
        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v
        key |= (spread(q[:, 0]) << 2 | spread(q[:, 1]) << 1 | spread(q[:, 2])) << 8
    if "dir" in a.presort:
        dn = d / np.linalg.norm(d, axis=1, keepdims=True)
        dq = np.clip((np.abs(dn) * 4).astype(np.int64), 0, 3).astype(np.uint64)
        key |= dq[:, 0] << 4 | dq[:, 1] << 2 | dq[:, 2]
    order = np.argsort(key, kind="stable")
    rays = np.ascontiguousarray(rays[order])
    print("presorted by", a.presort)
n = rays.shape[0]
full = a.query == "closest"
rb = eng.make_ray_buffers(n, api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT if full else api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID)
rb.d_rays[: 32 * n].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)))
cs = ctx.allocate_command_stream()
ctx.cmd_intersect(g.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST if full else api.RR_INTERSECT_QUERY_ANY, rb.p_rays, n, None, rb.output, rb.p_hits, rb.p_scratch, cs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(a.reps):
    e0.record()
    ctx.release_event(ctx.submit(cs))
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {r}: {e0.elapsed_time(e1):.3f} ms  {n / e0.elapsed_time(e1) / 1e3:.1f} Mrays/s")
ctx.release_command_stream(cs)
eng.close()
