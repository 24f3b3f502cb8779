"""Dev check of the packet kernel: differences against the oracle walk / brute force, and timing (python tools/packet_check.py)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import binding as O
from radeonrays_sdk_b200 import api, workloads as W
from radeonrays_sdk_b200.host import Engine
from helpers import assert_closest_hits_equal

eng = Engine(0)
ctx = eng.ctx
for name in ("cornell_box", "sponza"):
    pos, idx, _ = W.load_mesh(name)
    for flags in (api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, 0):
        g = eng.build_geometry(pos, idx, build_flags=flags)
        nodes = g.nodes()
        rays = W.cornell_primary_rays(512) if name == "cornell_box" else W.sponza_primary_rays(640, 360)
        got = eng.intersect(g, rays)
        want = O.trace(nodes, rays)
        try:
            n = assert_closest_hits_equal(got, want, (pos, idx), rays, what=name)
            print(name, flags, "ok, rays differing from the walk:", n, "of", rays.shape[0], flush=True)
        except AssertionError as e:
            print(name, flags, "FAIL", e, flush=True)
pos, idx, _ = W.load_mesh("sponza")
for flags in (0, api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD):
    g = eng.build_geometry(pos, idx, build_flags=flags)
    rays = W.sponza_primary_rays(3840, 2160)
    n = rays.shape[0]
    rb = eng.make_ray_buffers(n)
    rb.d_rays[: 32 * n].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)))
    cs = ctx.allocate_command_stream()
    ctx.cmd_intersect(g.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST, rb.p_rays, n, None, rb.output, rb.p_hits, rb.p_scratch, cs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for r in range(6):
        e0.record(); ctx.release_event(ctx.submit(cs)); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"flags {flags}: 4K primary closest best {best:.3f} ms  {n / best / 1e3:.1f} Mrays/s", flush=True)
    ctx.release_command_stream(cs)
eng.close()
