#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: Mrays/s closest-hit on Sponza (+ HLBVH build Mtris/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N=1 workload = BASELINE config C2: Sponza (262 267 triangles) HLBVH build + 3840x2160 coherent primary
closest-hit rays, FULL_HIT output.  A "step" is one rrCmdIntersect pass over the whole
8 294 400-ray batch.  `value` is timed with CUDA events around K steps with rays/BVH resident in HBM;
`e2e` is the same step through the rr* C ABI with HOST (pinned) ray and hit buffers, copies inside the
timed region.  N>1: one process per GPU (torchrun), BLAS built on rank 0 and broadcast over NCCL (it holds
indices, not pointers), every rank traces its own full batch (weak scaling, no data-path collective).

--impl reference times the reference's own CPU tracer (bvh_analyzer, compiled from /root/reference into
oracle/_ref by oracle/Makefile; falls back to the oracle port if that binary is absent) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 3840, 2160
BYTES_PER_RAY = 48           # 32 B ray read + 16 B RRHit write: the compulsory HBM traffic
TRACE_DRAM_BYTES_PER_LAUNCH = 357_213_952   # dram__bytes_read+write of one C2 launch, ncu --set full (profiles/round1_summary.md)
LANE_VISITS_PER_RAY = 66.6   # warp-iterations x 32 / rays of the C2 batch on the quality BVH (ncu instruction counts, same file)
BUILD_BYTES_PER_TRI = 348    # 48 aabb + 48 morton + 4 code + 60 sort + 12 emit reads + 48 gather + 128 nodes


BOUND_CPUS = 0


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process
    every ~2 ms (the timed region of a 20-step run is ~30 ms, shorter than one `nvidia-smi -lms` period); nvidia-smi is
    the fallback when the NVML binding is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self.stop_flag, self.thread, self.proc, self.rows, self.how = threading.Event(), None, None, [], None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.how = "nvml"

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                            else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in self.REASONS:
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.how = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.how == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml, 2 ms period, during the timed trace steps"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def bind_to_gpu_cpus(local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, before any pinned host buffer is allocated, so the
    H2D / D2H copies of the e2e leg do not cross sockets.  Returns how many CPUs the rank may use (0: left unbound)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1 and c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    global BOUND_CPUS
    BOUND_CPUS = bind_to_gpu_cpus(local) if world > 1 else 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def run_ours(args):
    import torch
    import torch.distributed as dist
    from radeonrays_sdk_b200 import api, workloads as W
    from radeonrays_sdk_b200.host import Engine

    rank, world, local = dist_setup(args.gpus)
    eng = Engine(local)
    ctx, dev = eng.ctx, eng.device
    pos, idx, _ = W.load_mesh("sponza")
    n_tris = idx.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- build (timed separately: HLBVH build Mtris/s, fast and quality) --------------------------------
    g_fast = eng.build_geometry(pos, idx, build_flags=api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD)
    g_qual = eng.build_geometry(pos, idx, build_flags=0)

    def resubmitter(record):
        cs = ctx.allocate_command_stream()
        record(cs)

        def go():
            ctx.release_event(ctx.submit(cs))
        return go, cs

    build_fast, cs1 = resubmitter(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, g_fast.input, g_fast.options, g_fast.p_temp, g_fast.p_nodes, s))
    build_qual, cs2 = resubmitter(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, g_qual.input, g_qual.options, g_qual.p_temp, g_qual.p_nodes, s))
    refit, cs3 = resubmitter(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, g_fast.input, g_fast.options, g_fast.p_temp, g_fast.p_nodes, s))
    l0 = ctx.launch_count()
    build_fast()
    launches_per_fast_build = ctx.launch_count() - l0
    bsteps = max(args.steps, 10)
    ms_fast = timed(build_fast, bsteps, args.warmup) / bsteps
    ms_qual = timed(build_qual, bsteps, args.warmup) / bsteps
    ms_refit = timed(refit, bsteps, args.warmup) / bsteps

    # N>1: rank 0's BLAS is the one everybody traces (broadcast over NCCL/NVLink; position independent)
    geom = g_qual if args.bvh == "quality" else g_fast
    if world > 1:
        dist.broadcast(geom.d_nodes, src=0)

    # ---- trace: device-resident ---------------------------------------------------------------------------
    rays = W.sponza_primary_rays(WIDTH, HEIGHT)
    n_rays = rays.shape[0]
    rb = eng.make_ray_buffers(n_rays)
    h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()
    h_hits = torch.empty(16 * n_rays, dtype=torch.uint8).pin_memory()
    rb.d_rays[: 32 * n_rays].copy_(h_rays)
    trace, cs4 = resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST, rb.p_rays, n_rays, None,
                                                         api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, rb.p_hits, rb.p_scratch, s))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms_total = timed(trace, args.steps, args.warmup)
    launches = (ctx.launch_count() - l0) * args.steps // (args.steps + args.warmup)   # k_trace + k_trace_deep per step
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    mrays = world * n_rays / (ms_step * 1e-3) / 1e6

    # the other trace flavours of the metric (any-hit; fast-build BVH), device resident, same batch
    other = g_fast if geom is g_qual else g_qual
    trace_other, cs5 = resubmitter(lambda s: ctx.cmd_intersect(other.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST, rb.p_rays, n_rays, None,
                                                               api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, rb.p_hits, rb.p_scratch, s))
    ms_other = timed(trace_other, args.steps, 1) / args.steps
    trace_any, cs6 = resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, api.RR_INTERSECT_QUERY_ANY, rb.p_rays, n_rays, None,
                                                             api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID, rb.p_hits, rb.p_scratch, s))
    ms_any = timed(trace_any, args.steps, 1) / args.steps

    # ---- trace: end to end through the C ABI with host buffers ------------------------------------------------
    from radeonrays_sdk_b200.host import HostTracePipeline
    pipe = HostTracePipeline(eng, geom, n_rays, chunks=args.e2e_chunks)

    def e2e_step():
        pipe.run(h_rays, h_hits)

    h_hits.zero_()
    ms_e2e = timed(e2e_step, args.steps, 2) / args.steps
    e2e_mrays = world * n_rays / (ms_e2e * 1e-3) / 1e6
    hits = h_hits.numpy().view(W.HIT_DTYPE)
    hit_fraction = float((hits["inst_id"] != W.INVALID).mean())

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(geom.nodes(), rays)

    # ---- config C5 on one GPU: HLBVH rebuild vs refit of a 50 M-triangle height field, as a fraction of the HBM roofline -------
    c5 = None
    if rank == 0 and world == 1 and not args.no_c5:
        c5 = build_c5(eng)

    peak, peak_src = measured_peak_gbs()
    l1_peak_visits = 2.0 * eng.sm_count * (clocks["sm_max_mhz"] if clocks and clocks.get("sm_max_mhz") else 1965.0) * 1e6
    achieved = BYTES_PER_RAY * n_rays / (ms_step * 1e-3) / 1e9
    build_gbs = BUILD_BYTES_PER_TRI * n_tris / (ms_fast * 1e-3) / 1e9
    out = {
        "metric": "Mrays/s closest-hit (Sponza)", "value": round(mrays, 2), "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic rays on the Sponza fixture (tests/golden/sponza.npz)",
        "config": {"workload": f"C2: Sponza {n_tris} tris, HLBVH build + {WIDTH}x{HEIGHT} coherent primary closest-hit rays, FULL_HIT",
                   "rays_per_step_per_gpu": n_rays, "bvh": f"{args.bvh} build", "parallelism": f"ray shards x{world}, BLAS broadcast",
                   "l2": "ray+hit buffers per step (398 MB) exceed the 126 MB L2; the 33.6 MB BVH is meant to stay L2 resident"},
        "e2e": {"value": round(e2e_mrays, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n_rays * world, "d2h_bytes_per_step": 16 * n_rays * world,
                "ms_per_step": round(ms_e2e, 4), "pipeline": f"{args.e2e_chunks} slices, H2D / rrCmdIntersect / D2H on three streams",
                "cpus_bound_to_gpu": BOUND_CPUS},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": TRACE_DRAM_BYTES_PER_LAUNCH if (n_rays, args.bvh) == (WIDTH * HEIGHT, "quality") else None,
                     "peak_source": peak_src, "kernel": "k_trace<closest,full_hit,one_level>",
                     "algorithmic_bytes_per_launch": BYTES_PER_RAY * n_rays,
                     "note": "48 B/ray compulsory (32 B ray + 16 B hit), BVH L2-resident: HBM is not the bound",
                     # the binding unit: every lane receives 64 B per visited node through the SM's
                     # 128 B/clk L1 data pipe => 2 node visits/clk/SM
                     "l1_data_pipe": {"lane_visits_per_ray": LANE_VISITS_PER_RAY, "peak_visits_per_s": round(l1_peak_visits, 1),
                                      "achieved_visits_per_s": round(LANE_VISITS_PER_RAY * n_rays / (ms_step * 1e-3), 1),
                                      "frac": round(LANE_VISITS_PER_RAY * n_rays / (ms_step * 1e-3) / l1_peak_visits, 4),
                                      "ncu_l1tex_data_pipe_pct": 91.2, "source": "profiles/round1_summary.md section 2"}},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "build": {"fast_ms": round(ms_fast, 4), "fast_mtris_per_s": round(n_tris / ms_fast / 1e3, 1), "quality_ms": round(ms_qual, 4),
                  "quality_mtris_per_s": round(n_tris / ms_qual / 1e3, 1), "refit_ms": round(ms_refit, 4),
                  "refit_mtris_per_s": round(n_tris / ms_refit / 1e3, 1), "launches_per_fast_build": int(launches_per_fast_build),
                  "fast_build_hbm_gbs": round(build_gbs, 1), "fast_build_roofline_frac": round(build_gbs / peak, 4)},
        "build_c5": c5,
        "trace_variants": {"closest_full_hit_other_bvh_mrays": round(n_rays / ms_other / 1e3, 1),
                           "any_hit_ids_mrays": round(n_rays / ms_any / 1e3, 1), "hit_fraction": round(hit_fraction, 4)},
    }
    if rank == 0:
        print(json.dumps(out), flush=True)
    pipe.close()
    for cs in (cs1, cs2, cs3, cs4, cs5, cs6):
        ctx.release_command_stream(cs)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def build_c5(eng, nx=5000, nz=5000, reps=5):
    """BASELINE config C5: 50 M-triangle animated height field, full rebuild vs RR_BUILD_OPERATION_UPDATE per frame (mesh made
    on the device, builds through rrCmdBuildGeometry, CUDA events); algorithmic bytes per triangle as in DESIGN.md section 4."""
    import torch
    from radeonrays_sdk_b200 import api
    from radeonrays_sdk_b200.host import Geometry, _dev_bytes
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_build import heightfield_device, REFIT_BYTES_PER_TRI
    ctx, dev = eng.ctx, eng.device
    pos, idx = heightfield_device(nx, nz, 0.0, dev)
    n = idx.shape[0]
    g = Geometry()
    g.engine, g.triangle_count, g.vertex_count, g.vertex_stride = eng, n, pos.shape[0], 12
    g.d_vertices, g.d_indices = pos.view(torch.uint8).reshape(-1), idx.view(torch.uint8).reshape(-1)
    g.options = api.RRBuildOptions(api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, None)
    g.p_vertices, g.p_indices = ctx.tensor_ptr(g.d_vertices), ctx.tensor_ptr(g.d_indices)
    g.input = ctx.geometry_input(g.p_vertices, g.vertex_count, 12, g.p_indices, n)
    g.req = ctx.geometry_requirements(g.input, g.options)
    g.d_temp, g.d_nodes = _dev_bytes(g.req.temporary_build_buffer_size, dev), _dev_bytes(g.req.result_buffer_size, dev)
    g.p_temp, g.p_nodes = ctx.tensor_ptr(g.d_temp), ctx.tensor_ptr(g.d_nodes)

    def timed(op):
        cs = ctx.allocate_command_stream()
        ctx.cmd_build_geometry(op, g.input, g.options, g.p_temp, g.p_nodes, cs)
        ctx.release_event(ctx.submit(cs))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.release_event(ctx.submit(cs))
        e1.record()
        torch.cuda.synchronize()
        ctx.release_command_stream(cs)
        return e0.elapsed_time(e1) / reps

    ms_build = timed(api.RR_BUILD_OPERATION_BUILD)
    pos2, _ = heightfield_device(nx, nz, 1.0, dev)
    g.d_vertices.copy_(pos2.view(torch.uint8).reshape(-1))
    ms_refit = timed(api.RR_BUILD_OPERATION_UPDATE)
    peak, _ = measured_peak_gbs()
    out = {"workload": f"C5: height field {nx}x{nz}, {n} triangles", "build_ms": round(ms_build, 4),
           "build_mtris_per_s": round(n / ms_build / 1e3, 1), "build_gbs": round(BUILD_BYTES_PER_TRI * n / ms_build / 1e6, 1),
           "build_roofline_frac": round(BUILD_BYTES_PER_TRI * n / ms_build / 1e6 / peak, 4), "refit_ms": round(ms_refit, 4),
           "refit_mtris_per_s": round(n / ms_refit / 1e3, 1), "refit_gbs": round(REFIT_BYTES_PER_TRI * n / ms_refit / 1e6, 1),
           "refit_roofline_frac": round(REFIT_BYTES_PER_TRI * n / ms_refit / 1e6 / peak, 4),
           "algorithmic_bytes_per_triangle": {"build": BUILD_BYTES_PER_TRI, "refit": REFIT_BYTES_PER_TRI}}
    del g, pos, idx, pos2
    torch.cuda.empty_cache()
    return out


def host_threads():
    """All the host threads this process may use (torchrun pins OMP_NUM_THREADS=1 for its ranks; the CPU arm undoes that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(nodes, rays, budget_rays=None):
    """Reference CPU tracer (oracle/_ref/bvh_analyzer_trace) or the oracle port on a bounded sample of the batch."""
    from oracle import binding as O
    sample = rays.reshape(HEIGHT, WIDTH)[::4, ::2].reshape(-1)   # 1 036 800 rays, every 4th row / 2nd column
    desc = f"{sample.shape[0]} rays = rows[::4], cols[::2] of the {WIDTH}x{HEIGHT} batch"
    res = O.ref_bvh_analyzer_trace(nodes, sample, repeats=3, threads=host_threads())
    if res is not None and res.get("is_valid"):
        return {"value": round(res["mrays_per_s"], 3), "unit": "Mrays/s", "cores": res["threads"], "kind": "reference",
                "sample": desc + "; bvh_analyzer BvhIntersect<2> loop only (bvh.h:87-93), 3 repeats, mean",
                "bvh_is_valid": True, "reference_sah": res["sah"]}
    t0 = time.time()
    O.trace(nodes, sample)
    dt = time.time() - t0
    return {"value": round(sample.shape[0] / dt / 1e6, 3), "unit": "Mrays/s", "cores": O.num_threads(), "kind": "port",
            "sample": desc + "; oracle/rr_oracle.c rro_trace (OpenMP)"}


def run_reference(args):
    """The reference arm: RadeonRays' own CPU path (bvh_analyzer) on this box's host cores, same metric/config.
    The BVH it traces is built by the CPU oracle (the reference has no CPU builder)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())   # before the OpenMP oracle library is loaded
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    pos, idx, _ = W.load_mesh("sponza")
    nodes, _, _ = O.build_blas(pos, idx, restructure=(args.bvh == "quality"))
    rays = W.sponza_primary_rays(WIDTH, HEIGHT)
    sample = rays.reshape(HEIGHT, WIDTH)[::4, ::2].reshape(-1)
    exe = os.path.join(ROOT, "oracle", "_ref", "bvh_analyzer_trace")
    kind = "reference" if os.path.exists(exe) else "port"
    times = []
    cores = O.num_threads()
    for step in range(args.warmup + args.steps):
        if kind == "reference":
            res = O.ref_bvh_analyzer_trace(nodes, sample, repeats=1, threads=host_threads())
            dt, cores = res["mean_s"], res["threads"]
        else:
            t0 = time.time()
            O.trace(nodes, sample)
            dt = time.time() - t0
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = sample.shape[0] / (ms * 1e-3) / 1e6
    desc = f"{sample.shape[0]} rays per step = rows[::4], cols[::2] of the {WIDTH}x{HEIGHT} batch"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s closest-hit (Sponza)", "value": round(v, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic rays on the Sponza fixture",
        "config": {"workload": f"C2: Sponza {idx.shape[0]} tris, {WIDTH}x{HEIGHT} coherent primary closest-hit rays (bounded sample per step)",
                   "bvh": f"{args.bvh} build (CPU oracle)"},
        "cpu_baseline": {"value": round(v, 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": round(v, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bvh", default="quality", choices=["quality", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the 50 M-triangle build / refit leg (N=1 only; ~2 s, 9 GB)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="slices of the host batch pipelined through H2D / trace / D2H")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
