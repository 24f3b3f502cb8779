#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: Mrays/s closest-hit on Sponza (+ HLBVH build Mtris/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N=1 workload = BASELINE config C2: Sponza (262 267 triangles) HLBVH build + 3840x2160 coherent primary
closest-hit rays, FULL_HIT output.  A "step" is one rrCmdIntersect pass over the whole
8 294 400-ray batch.  `value` is timed with CUDA events around K steps with rays/BVH resident in HBM;
`e2e` is the same step through the rr* C ABI with HOST (pinned) ray and hit buffers, copies inside the
timed region.  N>1: one process per GPU (torchrun), BLAS built on rank 0 and broadcast over NCCL (it holds
indices, not pointers); the global batch is N such frames sharded contiguously, rank r traces frame r and its
traversal kernels store the hits straight into rank 0's buffer over NVLink (CUDA IPC peer mapping), inside the timed
region (weak scaling in the batch size, with the gather of every hit to one rank included).  `c3_strong` in the same
line is BASELINE config C3: ONE 16 Mi-ray batch (diffuse CLOSEST, shadow ANY) sharded over the ranks, strong scaling.

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
PCIE_GEN5_X16_GBS = 63.0     # per direction
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


def workload_config(n_tris, n_rays, bvh, world):
    """`config` of the JSON line: the same dict in both arms (the reference arm always reports the N=1 form)."""
    return {"workload": f"C2: Sponza {n_tris} tris, HLBVH build + {WIDTH}x{HEIGHT} coherent primary closest-hit rays, FULL_HIT",
            "rays_per_step_per_gpu": n_rays, "bvh": f"{bvh} build",
            "parallelism": "1 GPU" if world == 1 else f"global batch of {world} frames sharded over {world} ranks (one frame each), BLAS broadcast over NCCL, "
                           "hits of all ranks gathered into rank 0's buffer by peer stores from the traversal kernels, inside the timed region",
            "l2": "ray+hit buffers per step (398 MB) exceed the 126 MB L2; the 33.6 MB BVH is meant to stay L2 resident"}


def load_ncu_summary():
    """profiles/round2_ncu_trace.json, written by tools/ncu_extract.py from the committed ncu capture of the dominant kernel."""
    p = os.path.join(ROOT, "profiles", "round2_ncu_trace.json")
    return json.load(open(p)) if os.path.exists(p) else None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from radeonrays_sdk_b200 import api, sharding, workloads as W
    from radeonrays_sdk_b200.host import Engine

    rank, world, local = dist_setup(args.gpus)
    eng = Engine(local)
    ctx, dev = eng.ctx, eng.device
    pos, idx, _ = W.load_mesh("sponza")
    n_tris = idx.shape[0]
    CLOSEST, ANY = api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_ANY
    FULL, IDS = api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID

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

    streams = []

    def resubmitter(record):
        cs = ctx.allocate_command_stream()
        record(cs)
        streams.append(cs)

        def go():
            ctx.release_event(ctx.submit(cs))
        return go

    # ---- build (timed separately: HLBVH build Mtris/s, fast and quality) --------------------------------
    g_fast = eng.build_geometry(pos, idx, build_flags=api.RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD)
    g_qual = eng.build_geometry(pos, idx, build_flags=0)
    build_fast = resubmitter(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, g_fast.input, g_fast.options, g_fast.p_temp, g_fast.p_nodes, s))
    build_qual = resubmitter(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_BUILD, g_qual.input, g_qual.options, g_qual.p_temp, g_qual.p_nodes, s))
    refit = resubmitter(lambda s: ctx.cmd_build_geometry(api.RR_BUILD_OPERATION_UPDATE, g_fast.input, g_fast.options, g_fast.p_temp, g_fast.p_nodes, s))
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

    # ---- trace: rays resident in HBM ------------------------------------------------------------------------
    # Global batch = `world` camera frames of 3840x2160 primary rays (frame r: the canonical camera raised by r/2 units), sharded
    # contiguously: rank r traces frame r.  N=1: hits go to a local buffer.  N>1: the hit buffer of the WHOLE batch lives on rank 0
    # and is mapped into every rank (sharding.PeerHitBuffer); each rank's traversal kernels store their hits straight into it
    # over NVLink while they trace -- the gather is inside the timed region, fused with the compute.
    def frame_rays(r):
        rays = W.sponza_primary_rays(WIDTH, HEIGHT)
        rays["origin"][:, 1] += np.float32(0.5 * r)
        return rays

    rays = frame_rays(rank)
    n_rays = rays.shape[0]
    rb = eng.make_ray_buffers(n_rays)
    h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()
    h_hits = torch.empty(16 * n_rays, dtype=torch.uint8).pin_memory()
    rb.d_rays[: 32 * n_rays].copy_(h_rays)
    peer, ring = None, []
    p_hits = rb.p_hits
    if world > 1:
        peer = sharding.PeerHitBuffer(ctx, world * 16 * n_rays, root=0)
        p_hits = peer.ptr(rank * 16 * n_rays)
    trace = resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, CLOSEST, rb.p_rays, n_rays, None, FULL, p_hits, rb.p_scratch, s))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms_total = timed(trace, args.steps, args.warmup)
    launches = (ctx.launch_count() - l0) // (args.steps + args.warmup)   # kernels per step (packet + per-ray + deep, one- and two-level)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    mrays = world * n_rays / (ms_step * 1e-3) / 1e6

    multi = None
    if world > 1:
        # (a) what rank 0 now holds must be what it gets by tracing every frame itself
        ok = True
        if rank == 0:
            got = peer.read(W.HIT_DTYPE).reshape(world, n_rays)
            for r in range(world):
                want = eng.intersect(geom, frame_rays(r))
                ok = ok and bool(np.array_equal(got[r].view(np.uint8), want.view(np.uint8)))
        # (b) comparison arm: trace into a local buffer, then one NCCL gather to rank 0
        local_trace = resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, CLOSEST, rb.p_rays, n_rays, None, FULL, rb.p_hits, rb.p_scratch, s))
        dst = [torch.empty(16 * n_rays, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        src = rb.d_hits[: 16 * n_rays]

        def nccl_step():
            local_trace()
            torch.cuda.current_stream(dev).wait_stream(eng.torch_stream)
            dist.gather(src, dst, dst=0)

        def nccl_only():
            dist.gather(src, dst, dst=0)
        # (c) the same fused stores with a DISTRIBUTED destination: the hits of frame r are delivered to rank (r + 1) mod N, so every
        # hit still crosses NVLink inside the timed region but every GPU takes in one frame instead of rank 0 taking in N - 1
        ring = [sharding.PeerHitBuffer(ctx, 16 * n_rays, root=r) for r in range(world)]
        p_next = ring[(rank + 1) % world].ptr(0)
        ring_trace = resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, CLOSEST, rb.p_rays, n_rays, None, FULL, p_next, rb.p_scratch, s))
        ms_ring = timed(ring_trace, max(5, args.steps // 2), 3) / max(5, args.steps // 2)
        mine = ring[rank].read(W.HIT_DTYPE)                                 # what the previous rank delivered: its frame
        ring_ok = torch.tensor([int(np.array_equal(mine.view(np.uint8), eng.intersect(geom, frame_rays((rank - 1) % world)).view(np.uint8)))], device=dev)
        dist.all_reduce(ring_ok, op=dist.ReduceOp.MIN)
        ms_nccl = timed(nccl_step, max(3, args.steps // 4), 2) / max(3, args.steps // 4)
        ms_gather = timed(nccl_only, max(3, args.steps // 4), 2) / max(3, args.steps // 4)
        ms_local = timed(local_trace, max(3, args.steps // 4), 2) / max(3, args.steps // 4)
        multi = {"hits_gathered_on_rank0_inside_timed_region": True, "gather": "traversal kernels store hits into rank 0's buffer through CUDA IPC / NVLink peer mappings",
                 "rank0_buffer_equals_local_trace_of_every_frame": ok, "nvlink_bytes_per_step": (world - 1) * 16 * n_rays,
                 "nvlink_gbs_into_rank0": round((world - 1) * 16 * n_rays / (ms_step * 1e-3) / 1e9, 1),
                 "local_trace_only_ms": round(ms_local, 4), "nccl_gather_only_ms": round(ms_gather, 4),
                 "trace_then_nccl_gather_ms": round(ms_nccl, 4), "fused_ms": round(ms_step, 4),
                 # one GPU can take in (world - 1) frames of 16-byte hits no faster than its NVLink ports deliver them
                 "mrays_per_s_if_hits_stayed_local": round(n_rays * world / ms_local / 1e3, 1),
                 "delivered_to_next_rank": {"what": "hits of frame r stored into rank (r + 1) mod N's buffer by the traversal kernels: every hit crosses NVLink, every GPU takes in one frame",
                                            "ms": round(ms_ring, 4), "mrays_per_s": round(n_rays * world / ms_ring / 1e3, 1),
                                            "nvlink_gbs_into_each_rank": round(16 * n_rays / (ms_ring * 1e-3) / 1e9, 1),
                                            "every_rank_holds_the_previous_ranks_frame": bool(ring_ok.item())},
                 "limiter": ("NVLink into rank 0: %d MB of hits per step, %.0f GB/s" % ((world - 1) * 16 * n_rays // 1000000, (world - 1) * 16 * n_rays / (ms_step * 1e-3) / 1e9)
                             if ms_step > 1.05 * ms_local else "the trace itself (the gather hides behind it)")}

    # the other trace flavours of the metric (any-hit; the other BVH; the per-ray kernel), device resident, same batch (N=1 only)
    variants = None
    if world == 1:
        other = g_fast if geom is g_qual else g_qual
        ms_other = timed(resubmitter(lambda s: ctx.cmd_intersect(other.p_nodes, CLOSEST, rb.p_rays, n_rays, None, FULL, rb.p_hits, rb.p_scratch, s)), args.steps, 1) / args.steps
        ms_any = timed(resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, ANY, rb.p_rays, n_rays, None, IDS, rb.p_hits, rb.p_scratch, s)), args.steps, 1) / args.steps
        ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 1)      # the reference's tie rule: per-ray kernel (k_trace), no packets
        ms_ff = timed(resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, CLOSEST, rb.p_rays, n_rays, None, FULL, rb.p_hits, rb.p_scratch, s)), args.steps, 1) / args.steps
        ctx.set_option(api.RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, 0)
        ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 1)                    # packets of 64 consecutive rays instead of 8 x 8 tiles of the image
        ms_strips = timed(resubmitter(lambda s: ctx.cmd_intersect(geom.p_nodes, CLOSEST, rb.p_rays, n_rays, None, FULL, rb.p_hits, rb.p_scratch, s)), args.steps, 1) / args.steps
        ctx.set_option(api.RR_CUDA_OPTION_RAY_GRID_WIDTH, 0)
        variants = {"closest_full_hit_other_bvh_mrays": round(n_rays / ms_other / 1e3, 1), "any_hit_ids_mrays": round(n_rays / ms_any / 1e3, 1),
                    "closest_first_found_rule_per_ray_kernel_mrays": round(n_rays / ms_ff / 1e3, 1),
                    "closest_full_hit_64x1_strip_packets_mrays": round(n_rays / ms_strips / 1e3, 1)}

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
    pcie_gbs = 48 * n_rays / (ms_e2e * 1e-3) / 1e9

    # ---- config C3: one 16 Mi-ray batch (diffuse CLOSEST / FULL_HIT and shadow ANY / ids) sharded over the ranks, strong scaling ----
    c3 = None
    if not args.no_c3:
        c3 = c3_strong(eng, geom, pos, idx, rank, world, timed, resubmitter, args_blocks=args.c3_blocks)

    # ---- config C4: 1 000 instanced Sponza BLASes under a TLAS (N=1) ------------------------------------------------
    c4 = None
    if world == 1 and not args.no_c4:
        c4 = scene_c4(eng, geom, timed, resubmitter)

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(geom.nodes(), rays)

    # ---- config C5 on one GPU: HLBVH rebuild vs refit of a 50 M-triangle height field, as a fraction of the HBM roofline -------
    c5 = None
    if rank == 0 and world == 1 and not args.no_c5:
        c5 = build_c5(eng)

    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_RAY * n_rays / (ms_step * 1e-3) / 1e9
    build_gbs = BUILD_BYTES_PER_TRI * n_tris / (ms_fast * 1e-3) / 1e9
    ncu = load_ncu_summary()
    same_kernel_config = ncu is not None and (n_rays, args.bvh) == (WIDTH * HEIGHT, "quality")
    out = {
        "metric": "Mrays/s closest-hit (Sponza)", "value": round(mrays, 2), "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic rays on the Sponza fixture (tests/golden/sponza.npz)",
        "config": workload_config(n_tris, n_rays, args.bvh, world),
        "e2e": {"value": round(e2e_mrays, 2), "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n_rays * world, "d2h_bytes_per_step": 16 * n_rays * world,
                "ms_per_step": round(ms_e2e, 4), "pipeline": f"{args.e2e_chunks} slices, H2D / rrCmdIntersect / D2H on three streams",
                "cpus_bound_to_gpu": BOUND_CPUS, "bound": "pcie", "pcie_gbs_per_gpu_both_directions": round(pcie_gbs, 1),
                "pcie_wire_fraction": round(pcie_gbs / 2 / PCIE_GEN5_X16_GBS, 3),
                "note": "48 B per ray cross PCIe (32 in, 16 out); wire fraction = per-direction average against Gen5 x16 (63 GB/s)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": ncu["dram_bytes"] if same_kernel_config else None,
                     "peak_source": peak_src, "kernel": "k_trace_packet<full_hit> (8 x 8 ray tiles found by k_detect_grid; + k_trace for declined chunks)",
                     "algorithmic_bytes_per_launch": BYTES_PER_RAY * n_rays,
                     "note": "48 B/ray compulsory (32 B ray + 16 B hit), BVH L2-resident: HBM is not the bound; the kernel is bound by the SM's "
                             "ALU pipe / issue slots (profiles/round2_summary.md)",
                     "ncu": ncu if same_kernel_config else None},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "build": {"fast_ms": round(ms_fast, 4), "fast_mtris_per_s": round(n_tris / ms_fast / 1e3, 1), "quality_ms": round(ms_qual, 4),
                  "quality_mtris_per_s": round(n_tris / ms_qual / 1e3, 1), "refit_ms": round(ms_refit, 4),
                  "refit_mtris_per_s": round(n_tris / ms_refit / 1e3, 1), "launches_per_fast_build": int(launches_per_fast_build),
                  "fast_build_hbm_gbs": round(build_gbs, 1), "fast_build_roofline_frac": round(build_gbs / peak, 4)},
        "build_c5": c5,
        "c3_strong": c3,
        "c4_two_level": c4,
        "multi_gpu": multi,
        "trace_variants": dict(variants or {}, hit_fraction=round(hit_fraction, 4)),
    }
    if rank == 0:
        print(json.dumps(out), flush=True)
    pipe.close()
    if peer is not None:
        barrier()
        peer.close()
        for b in ring:
            b.close()
    for cs in streams:
        ctx.release_command_stream(cs)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def c3_strong(eng, geom, pos, idx, rank, world, timed, resubmitter, count=1 << 24, steps=5, args_blocks=1):
    """BASELINE config C3: ONE batch of 16 Mi rays (shadow ANY / ids, diffuse-bounce CLOSEST / FULL_HIT) sharded contiguously over
    the ranks (sharding.shard_range), every rank's hits landing in rank 0's buffer by peer stores from the traversal kernels inside
    the timed region.  Strong scaling: total work is fixed, the time at N ranks is what a client on rank 0 waits for all hits."""
    import torch
    import torch.distributed as dist
    from radeonrays_sdk_b200 import api, sharding, workloads as W
    ctx, dev = eng.ctx, eng.device
    out = {"rays_per_batch": count, "ranks": world}
    d_rays = {}
    for name in ("diffuse", "shadow"):
        t = torch.empty(32 * count, dtype=torch.uint8, device=dev)
        if rank == 0:
            prim = W.sponza_primary_rays(1024, 1024)
            hits = eng.intersect(geom, prim)
            r = (W.diffuse_rays if name == "diffuse" else W.shadow_rays)(pos, idx, prim, hits, count=count)
            t.copy_(torch.from_numpy(r.view(np.uint8).reshape(-1)))
        if world > 1:
            dist.broadcast(t, src=0)
        d_rays[name] = t
    # One contiguous slice per rank by default.  --c3-blocks k cuts the batch into world x k blocks dealt round-robin (rank r traces
    # blocks r, r + world, ...; one rrCmdIntersect each in one command stream) to even out the cost of image regions -- measured at
    # N = 8: 1.56 ms (k = 1), 2.49 ms (k = 4), 3.76 ms (k = 8) for the diffuse batch: every extra call pays the tail of a persistent
    # grid that 512 Ki incoherent rays under-fill (3 chunks per warp), far more than the imbalance it removes.
    blocks_per_rank = 1 if world == 1 else args_blocks
    nblocks = world * blocks_per_rank
    spans = [sharding.shard_range(count, rank + world * k, nblocks) for k in range(blocks_per_rank)]
    biggest = max(e - b for b, e in spans)
    scratch = torch.empty(max(ctx.trace_requirements(biggest), 256), dtype=torch.uint8, device=dev)
    p_scratch = ctx.tensor_ptr(scratch)
    out["blocks_per_rank"] = blocks_per_rank
    for name, query, output, item in (("diffuse", api.RR_INTERSECT_QUERY_CLOSEST, api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, 16),
                                      ("shadow", api.RR_INTERSECT_QUERY_ANY, api.RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID, 4)):
        peer, local = None, None
        if world > 1:
            peer = sharding.PeerHitBuffer(ctx, item * count, root=0)
        else:
            local = torch.empty(item * count, dtype=torch.uint8, device=dev)

        def record(s):
            for b, e in spans:
                if e > b:
                    p_rays = ctx.tensor_ptr(d_rays[name], 32 * b)
                    p_hits = peer.ptr(item * b) if peer is not None else ctx.tensor_ptr(local, item * b)
                    ctx.cmd_intersect(geom.p_nodes, query, p_rays, e - b, None, output, p_hits, p_scratch, s)
        go = resubmitter(record)
        ms = timed(go, steps, 2) / steps
        out[name + "_ms"] = round(ms, 4)
        out[name + "_mrays_per_s"] = round(count / ms / 1e3, 1)
        if name == "diffuse":
            # the same batch / shards with on-device ray binning (RR_CUDA_OPTION_SORT_RAYS: key pass + radix sort inside the timed
            # call, bit-identical hits in the client's order; tests/test_gpu_trace.py::test_ray_binning_option_is_bit_identical)
            ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 1)
            big = torch.empty(ctx.trace_requirements(biggest), dtype=torch.uint8, device=dev)
            p_big = ctx.tensor_ptr(big)

            def record2(s):
                for b, e in spans:
                    if e > b:
                        p_rays = ctx.tensor_ptr(d_rays[name], 32 * b)
                        p_hits = peer.ptr(item * b) if peer is not None else ctx.tensor_ptr(local, item * b)
                        ctx.cmd_intersect(geom.p_nodes, query, p_rays, e - b, None, output, p_hits, p_big, s)
            go2 = resubmitter(record2)
            ctx.set_option(api.RR_CUDA_OPTION_SORT_RAYS, 0)
            ms2 = timed(go2, steps, 2) / steps
            out["diffuse_binned_ms"] = round(ms2, 4)
            out["diffuse_binned_mrays_per_s"] = round(count / ms2 / 1e3, 1)
            del big
        if world > 1:
            out[name + "_nvlink_bytes_into_rank0"] = item * (count - sum(e - b for b, e in [sharding.shard_range(count, world * k, nblocks) for k in range(blocks_per_rank)]))
            if dist.get_rank() == 0:
                got = peer.read(np.uint8)
                sel = np.arange(0, count, 4099)
                full = got.view(W.HIT_DTYPE)["inst_id"] if item == 16 else got.view(np.uint32)
                out[name + "_hit_fraction"] = round(float((full[sel] != W.INVALID).mean()), 4)
            dist.barrier()
            torch.cuda.synchronize(dev)
            peer.close()
    return out


def scene_c4(eng, geom, timed, resubmitter, steps=5):
    """BASELINE config C4: 1 000 rotated instances of the Sponza BLAS under one TLAS, rrCmdBuildScene + 3840x2160 primary rays
    from outside the grid, two-level rrCmdIntersect (closest, FULL_HIT)."""
    import torch
    from radeonrays_sdk_b200 import api, workloads as W
    ctx, dev = eng.ctx, eng.device
    xf = W.grid_instances(10, 250.0, 7.0)
    sc = eng.build_scene([geom], [0] * xf.shape[0], xf)
    # scene builds stage the host instance array: submitted directly, not replayed as a graph
    cs = ctx.allocate_command_stream()
    ctx.cmd_build_scene(sc.input, None, sc.p_temp, sc.p_nodes, cs)
    ms_scene = timed(lambda: ctx.release_event(ctx.submit(cs)), steps, 1) / steps
    ctx.release_command_stream(cs)
    rays = W.grid_camera_rays(WIDTH, HEIGHT)
    n = rays.shape[0]
    rb = eng.make_ray_buffers(n)
    rb.d_rays[: 32 * n].copy_(torch.from_numpy(rays.view(np.uint8).reshape(-1)))
    go = resubmitter(lambda s: ctx.cmd_intersect(sc.p_nodes, api.RR_INTERSECT_QUERY_CLOSEST, rb.p_rays, n, None, api.RR_INTERSECT_QUERY_OUTPUT_FULL_HIT,
                                                 rb.p_hits, rb.p_scratch, s))
    ms = timed(go, steps, 1) / steps
    hits = rb.d_hits[: 16 * n].cpu().numpy().view(W.HIT_DTYPE)
    return {"workload": f"C4: {xf.shape[0]} instances of the Sponza BLAS, rotated 10x10x10 grid, {WIDTH}x{HEIGHT} primary rays", "scene_build_ms": round(ms_scene, 4),
            "trace_ms": round(ms, 4), "trace_mrays_per_s": round(n / ms / 1e3, 1), "hit_fraction": round(float((hits["inst_id"] != W.INVALID).mean()), 4)}


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
    # the same mesh with 63-bit Morton codes (RR_CUDA_OPTION_MORTON_BITS = 63: two 4-pass sorts instead of one)
    ctx.set_option(api.RR_CUDA_OPTION_MORTON_BITS, 63)
    req63 = ctx.geometry_requirements(g.input, g.options)
    g.d_temp = None
    torch.cuda.empty_cache()
    g.d_temp = _dev_bytes(req63.temporary_build_buffer_size, dev)
    g.p_temp = ctx.tensor_ptr(g.d_temp)
    ms_build63 = timed(api.RR_BUILD_OPERATION_BUILD)
    ctx.set_option(api.RR_CUDA_OPTION_MORTON_BITS, 30)
    peak, _ = measured_peak_gbs()
    out = {"workload": f"C5: height field {nx}x{nz}, {n} triangles", "build_ms": round(ms_build, 4), "build_morton63_ms": round(ms_build63, 4),
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


def reference_rays():
    """The C2 batch as the reference's CPU tracer must be given it: every direction multiplied by a power of two so that all
    |components| >= 1 (workloads.pow2_scaled_rays: bit-identical hits; bvh_analyzer's `abs` is the integer overload under g++ and
    mis-traverses any ray with a |component| < 1, which would make it look several times faster than it is)."""
    from radeonrays_sdk_b200 import workloads as W
    rays, _ = W.pow2_scaled_rays(W.sponza_primary_rays(WIDTH, HEIGHT), keep_all=True)
    return rays


def cpu_baseline(nodes, rays):
    """Reference CPU tracer (oracle/_ref/bvh_analyzer_trace: the `#pragma omp parallel for` loop over BvhIntersect<2> of
    bvh_analyzer/bvh.h:87-93 with its stock scheduling) on the FULL C2 batch, or the oracle port if that binary is absent; plus one
    run of the unmodified stock binary on the 7-line config (a 960x540 sub-sample: it serialises every ray on an omp critical)."""
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    full = reference_rays()
    desc = f"all {full.shape[0]} rays of the {WIDTH}x{HEIGHT} batch (directions scaled by powers of two, same hits)"
    res = O.ref_bvh_analyzer_trace(nodes, full, repeats=2, threads=host_threads(), stock_schedule=True)
    if res is not None and res.get("is_valid"):
        out = {"value": round(res["mrays_per_s"], 3), "unit": "Mrays/s", "cores": res["threads"], "kind": "reference",
               "sample": desc + "; bvh_analyzer BvhIntersect<2> loop only (bvh.h:87-93), stock `omp parallel for`, 2 repeats, mean",
               "bvh_is_valid": True, "reference_sah": res["sah"], "hit_fraction": round(res["hit_count"] / full.shape[0], 4),
               "avg_node_tests": res["avg_node_tests"], "avg_triangle_tests": res["avg_tri_tests"]}
        sub, _ = W.pow2_scaled_rays(W.sponza_primary_rays(960, 540), keep_all=True)
        stock = O.ref_bvh_analyzer_stock(nodes, sub, 960, 540, threads=host_threads())
        if stock is not None:
            out["stock_binary"] = {"config": "7-line config (bvh_analyzer/config.h:46-63), 960x540 rays", "returncode": stock["returncode"],
                                   "is_valid": stock.get("is_valid"), "sah": stock.get("sah"), "avg_primary_node_tests": stock.get("avg_primary_node_tests"),
                                   "avg_primary_triangle_tests": stock.get("avg_primary_triangle_tests"), "wall_s": round(stock["wall_s"], 2),
                                   "mrays_per_s_end_to_end": round(sub.shape[0] / stock["wall_s"] / 1e6, 3)}
        return out
    sample = full.reshape(HEIGHT, WIDTH)[::4, ::2].reshape(-1)
    t0 = time.time()
    O.trace(nodes, sample)
    dt = time.time() - t0
    return {"value": round(sample.shape[0] / dt / 1e6, 3), "unit": "Mrays/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{sample.shape[0]} rays = rows[::4], cols[::2] of the batch; oracle/rr_oracle.c rro_trace (OpenMP)"}


def run_reference(args):
    """The reference arm: RadeonRays' own CPU path (bvh_analyzer) on this box's host cores, same metric / config: every step
    traces the FULL 3840x2160 batch with the stock `#pragma omp parallel for` (same_config).  The BVH it traces is built by the
    CPU oracle (the reference has no CPU builder)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_threads())   # before the OpenMP oracle library is loaded
    from oracle import binding as O
    from radeonrays_sdk_b200 import workloads as W
    pos, idx, _ = W.load_mesh("sponza")
    nodes, _, _ = O.build_blas(pos, idx, restructure=(args.bvh == "quality"))
    rays = reference_rays()
    exe = os.path.join(ROOT, "oracle", "_ref", "bvh_analyzer_trace")
    kind = "reference" if os.path.exists(exe) else "port"
    cores, unscaled = O.num_threads(), None
    if kind == "reference":
        if args.warmup:
            O.ref_bvh_analyzer_trace(nodes, rays, repeats=args.warmup, threads=host_threads(), stock_schedule=True)
        res = O.ref_bvh_analyzer_trace(nodes, rays, repeats=args.steps, threads=host_threads(), stock_schedule=True)
        ms, cores = 1e3 * res["mean_s"], res["threads"]
        # for the record: the same binary on the batch as generated (|direction components| < 1): most rays are mis-traversed and miss
        raw = O.ref_bvh_analyzer_trace(nodes, W.sponza_primary_rays(WIDTH, HEIGHT), repeats=1, threads=host_threads(), stock_schedule=True)
        unscaled = {"mrays_per_s": round(raw["mrays_per_s"], 3), "hit_fraction": round(raw["hit_count"] / rays.shape[0], 4),
                    "hit_fraction_scaled": round(res["hit_count"] / rays.shape[0], 4)}
        desc = f"all {rays.shape[0]} rays per step, bvh_analyzer BvhIntersect<2> loop (bvh.h:87-93), stock omp scheduling; directions scaled by powers of two"
    else:
        rays = rays.reshape(HEIGHT, WIDTH)[::4, ::2].reshape(-1)
        times = []
        for step in range(args.warmup + args.steps):
            t0 = time.time()
            O.trace(nodes, rays)
            if step >= args.warmup:
                times.append(time.time() - t0)
        ms = 1e3 * float(np.mean(times))
        desc = f"{rays.shape[0]} rays per step = rows[::4], cols[::2] of the batch; oracle port (oracle/_ref absent)"
    v = rays.shape[0] / (ms * 1e-3) / 1e6
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s closest-hit (Sponza)", "value": round(v, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic rays on the Sponza fixture",
        "config": workload_config(idx.shape[0], rays.shape[0], args.bvh, 1),
        "cpu_baseline": {"value": round(v, 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": desc,
                         "bvh_built_by": "CPU oracle (the reference has no CPU builder)", "unscaled_rays": unscaled},
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
    ap.add_argument("--no-c3", action="store_true", help="skip the 16 Mi-ray shadow / diffuse strong-scaling leg")
    ap.add_argument("--c3-blocks", type=int, default=1, help="blocks per rank of the block-cyclic C3 shards (N>1)")
    ap.add_argument("--no-c4", action="store_true", help="skip the 1 000-instance two-level leg (N=1 only)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="slices of the host batch pipelined through H2D / trace / D2H")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
