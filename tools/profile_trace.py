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
