"""Timing of RR_CUDA_OPTION_SORT_RAYS on the C3 batches (16 Mi rays): python tools/binning_check.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from radeonrays_sdk_b200 import api, workloads as W
from radeonrays_sdk_b200.host import Engine
eng = Engine(0); ctx = eng.ctx
pos, idx, _ = W.load_mesh("sponza")
g = eng.build_geometry(pos, idx, build_flags=0)
prim = W.sponza_primary_rays(1024, 1024)
hits = eng.intersect(g, prim)
for name, rays, q, o in (("diffuse closest", W.diffuse_rays(pos, idx, prim, hits, count=1 << 24), api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT),
                         ("shadow any", W.shadow_rays(pos, idx, prim, hits, count=1 << 24), api.RR_INTERSECT_QUERY_ANY, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID)):
    n = rays.shape[0]
    res = {}
    for sort in (0, 1):
        ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, sort)
        rb = eng.make_ray_buffers(n, o)
        rb.d_rays[: 32 * n].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)))
        cs = ctx.allocate_command_stream()
        ctx.cmd_intersect(g.p_nodes, q, rb.p_rays, n, None, o, rb.p_hits, rb.p_scratch, cs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for r in range(5):
            e0.record(); ctx.release_event(ctx.submit(cs)); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[sort] = (best, rb.d_hits[: rb.hit_bytes].cpu().numpy().copy())
        ctx.release_command_stream(cs)
        print(f"{name}: sort={sort} best {best:.3f} ms  {n / best / 1e3:.1f} Mrays/s", flush=True)
    print(name, "identical:", bool(np.array_equal(res[0][1], res[1][1])), flush=True)
ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 0)
eng.close()
