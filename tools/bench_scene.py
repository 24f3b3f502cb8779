"""BASELINE config C4: 1 000 instances of the Sponza BLAS on a 10x10x10 grid under one TLAS (rrCmdBuildScene) + 4K primary
rays from a camera outside the grid (two-level rrCmdIntersect).  Prints one JSON line.
  python tools/bench_scene.py [--side 10] [--reps 5] [--check N]   (--check: compare N rays with the CPU oracle)"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radeonrays_sdk_b200 import api, workloads as W
from radeonrays_sdk_b200.host import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--check", type=int, default=0)
a = ap.parse_args()
eng = Engine(0)
ctx = eng.ctx
pos, idx, _ = W.load_mesh("sponza")
g = eng.build_geometry(pos, idx, build_flags=0)
xf = W.grid_instances(a.side, 250.0, 7.0)
n_inst = xf.shape[0]
sc = eng.build_scene([g], [0] * n_inst, xf)


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cs = ctx.allocate_command_stream()
ctx.cmd_build_scene(sc.input, None, sc.p_temp, sc.p_nodes, cs)
ms_scene = timed(lambda: ctx.release_event(ctx.submit(cs)), a.reps)
ctx.release_command_stream(cs)
# camera outside the grid looking at its centre
extent = 250.0 * (a.side - 1)
centre = np.array([extent / 2, extent / 2 + 30.0, extent / 2], np.float32)
eye = centre + np.array([-1.2 * extent - 400.0, 0.35 * extent, -0.9 * extent - 300.0], np.float32)
fwd = (centre - eye) / np.linalg.norm(centre - eye)
right = np.cross(fwd, [0, 1, 0]); right /= np.linalg.norm(right)
up = np.cross(right, fwd)
x = (np.arange(a.width, dtype=np.float32) + 0.5) / a.width * 2 - 1
y = (np.arange(a.height, dtype=np.float32) + 0.5) / a.height * 2 - 1
d = fwd[None, None, :] + 0.6 * x[None, :, None] * right[None, None, :] + 0.6 * (a.height / a.width) * y[:, None, None] * up[None, None, :]
rays = W._pack_rays(eye, d.reshape(-1, 3).astype(np.float32), min_t=0.001, max_t=1e6)
n = rays.shape[0]
rb = eng.make_ray_buffers(n)
rb.d_rays[: 32 * n].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)))
cs = ctx.allocate_command_stream()
ctx.cmd_intersect(sc.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST, rb.p_rays, n, None, api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, rb.p_hits, rb.p_scratch, cs)
ms_trace = timed(lambda: ctx.release_event(ctx.submit(cs)), a.reps)
ctx.release_command_stream(cs)
hits = rb.d_hits[: 16 * n].cpu().numpy().view(W.HIT_DTYPE)
out = {"workload": f"C4: {n_inst} instances of Sponza ({idx.shape[0]} tris each), rotated grid", "scene_build_ms": round(ms_scene, 4),
       "instances_per_s_M": round(n_inst / ms_scene / 1e3, 2), "rays": n, "trace_ms": round(ms_trace, 3),
       "trace_mrays_per_s": round(n / ms_trace / 1e3, 1), "hit_fraction": round(float((hits["inst_id"] != W.INVALID).mean()), 4),
       "distinct_instances_hit": int(np.unique(hits["inst_id"][hits["inst_id"] != W.INVALID]).size)}
if a.check:
    from oracle import binding as O
    blas = [g.nodes()]
    tlas, oxf = O.build_tlas(blas, [0] * n_inst, xf)
    sel = np.sort(np.random.default_rng(0).choice(n, a.check, replace=False))
    want = O.trace_2l(tlas, oxf, blas, [0] * n_inst, rays[sel], init=np.zeros(a.check, W.HIT_DTYPE))
    ok = want["inst_id"] != O.INVALID
    out["oracle_rays_checked"] = int(a.check)
    out["ids_bit_exact"] = bool(np.array_equal(hits["inst_id"][sel], want["inst_id"]) and np.array_equal(hits["prim_id"][sel][ok], want["prim_id"][ok]))
    out["uv_bit_exact"] = bool(np.array_equal(hits["uv"][sel][ok].view(np.uint32), want["uv"][ok].view(np.uint32)))
print(json.dumps(out), flush=True)
eng.close()
