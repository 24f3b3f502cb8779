"""HLBVH build / refit throughput on large synthetic meshes (BASELINE config C5: animated height field).

  python tools/bench_build.py [--sizes 1000x500,4000x1000,5000x5000] [--reps 5] [--check]

Mesh vertices/indices are generated on the device with torch (client-side data, like a renderer would own);
the build and refit run through rrCmdBuildGeometry.  Prints one JSON line per size.
"""
import argparse, json, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radeonrays_sdk_b200 import api
from radeonrays_sdk_b200.host import Engine, Geometry, _dev_bytes

BUILD_BYTES_PER_TRI = 348
REFIT_BYTES_PER_TRI = 176


def heightfield_device(nx, nz, t, dev):
    xs = torch.arange(nx + 1, dtype=torch.float32, device=dev)
    zs = torch.arange(nz + 1, dtype=torch.float32, device=dev)
    Z, X = torch.meshgrid(zs, xs, indexing="ij")
    Y = 2.0 * torch.sin(0.05 * X + 1.0 * t) * torch.cos(0.07 * Z)
    pos = torch.stack([X, Y, Z], -1).reshape(-1, 3).contiguous()
    i = (torch.arange(nz, dtype=torch.int64, device=dev)[:, None] * (nx + 1) + torch.arange(nx, dtype=torch.int64, device=dev)[None, :]).reshape(-1)
    a, b, c, d = i, i + 1, i + nx + 1, i + nx + 2
    idx = torch.stack([torch.stack([a, c, b], -1), torch.stack([b, c, d], -1)], 1).reshape(-1, 3).to(torch.int32).contiguous()
    return pos, idx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000x500,4000x1000,5000x5000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="download the BVH and run the oracle's structural check")
    ap.add_argument("--no-refit", action="store_true", help="time the build only")
    ap.add_argument("--quality", action="store_true", help="build_flags = 0: treelet-restructured tree, generic staged refit")
    args = ap.parse_args()
    eng = Engine(0)
    ctx, dev = eng.ctx, eng.device
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for size in args.sizes.split(","):
        nx, nz = (int(v) for v in size.split("x"))
        pos, idx = heightfield_device(nx, nz, 0.0, dev)
        n = idx.shape[0]
        g = Geometry()
        g.engine, g.triangle_count, g.vertex_count, g.vertex_stride = eng, n, pos.shape[0], 12
        g.d_vertices, g.d_indices = pos.view(torch.uint8).reshape(-1), idx.view(torch.uint8).reshape(-1)
        g.options = api.RRBuildOptions(0 if args.quality else api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, None)
        g.p_vertices, g.p_indices = ctx.tensor_ptr(g.d_vertices), ctx.tensor_ptr(g.d_indices)
        g.input = ctx.geometry_input(g.p_vertices, g.vertex_count, 12, g.p_indices, n)
        g.req = ctx.geometry_requirements(g.input, g.options)
        g.d_temp, g.d_nodes = _dev_bytes(max(g.req.temporary_build_buffer_size, g.req.temporary_update_buffer_size), dev), _dev_bytes(g.req.result_buffer_size, dev)
        g.p_temp, g.p_nodes = ctx.tensor_ptr(g.d_temp), ctx.tensor_ptr(g.d_nodes)

        def timed(op):
            cs = ctx.allocate_command_stream()
            ctx.cmd_build_geometry(op, g.input, g.options, g.p_temp, g.p_nodes, cs)
            ctx.release_event(ctx.submit(cs))            # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                ctx.release_event(ctx.submit(cs))
            e1.record()
            torch.cuda.synchronize()
            ctx.release_command_stream(cs)
            return e0.elapsed_time(e1) / args.reps

        ms_build = timed(api.RR_BUILD_OPERATION_BUILD)
        # device-side sanity: sorted codes ascending, refs a permutation, root box == mesh bounds
        L = ctx.build_scratch_layout(n)
        sc = g.d_temp[L.sorted_codes_offset: L.sorted_codes_offset + 4 * n].view(torch.int32)
        sr = g.d_nodes[L.sorted_refs_offset: L.sorted_refs_offset + 4 * n].view(torch.int32)   # geometry buffer tail
        sorted_ok = bool((sc[1:] >= sc[:-1]).all().item())
        perm_ok = bool((torch.sort(sr.to(torch.int64)).values == torch.arange(n, device=dev)).all().item())
        root = g.d_nodes[:64].view(torch.float32).cpu().numpy()
        lo = np.minimum(root[0:3], root[8:11]); hi = np.maximum(root[4:7], root[12:15])
        box_ok = bool(np.array_equal(lo, pos.min(0).values.cpu().numpy()) and np.array_equal(hi, pos.max(0).values.cpu().numpy()))
        pos2, _ = heightfield_device(nx, nz, 1.0, dev)
        g.d_vertices.copy_(pos2.view(torch.uint8).reshape(-1))
        ms_refit = timed(api.RR_BUILD_OPERATION_UPDATE) if not args.no_refit else float('nan')
        out = {"workload": f"height field {nx}x{nz}" + (" (quality build)" if args.quality else ""), "triangles": n, "build_ms": round(ms_build, 4), "build_mtris_per_s": round(n / ms_build / 1e3, 1),
               "build_gbs": round(BUILD_BYTES_PER_TRI * n / ms_build / 1e6, 1), "build_roofline_frac": round(BUILD_BYTES_PER_TRI * n / ms_build / 1e6 / peak, 4),
               "refit_ms": round(ms_refit, 4), "refit_mtris_per_s": round(n / ms_refit / 1e3, 1),
               "refit_gbs": round(REFIT_BYTES_PER_TRI * n / ms_refit / 1e6, 1), "refit_roofline_frac": round(REFIT_BYTES_PER_TRI * n / ms_refit / 1e6 / peak, 4),
               "sorted_ok": sorted_ok, "perm_ok": perm_ok, "root_box_ok": box_ok,
               "scratch_bytes": g.req.temporary_build_buffer_size, "result_bytes": g.req.result_buffer_size}
        if args.check:
            from oracle import binding as O
            t0 = time.time()
            nodes = g.nodes()
            out["consistent"] = O.check_consistency(nodes)
            out["check_s"] = round(time.time() - t0, 1)
            del nodes
        print(json.dumps(out), flush=True)
        del g, pos, idx, pos2
        torch.cuda.empty_cache()
    eng.close()


if __name__ == "__main__":
    main()
