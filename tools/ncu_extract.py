"""Summarise one kernel launch of an .ncu-rep into a small JSON that bench.py reads (no literals in bench.py):
  python tools/ncu_extract.py gpurun_out/prof.ncu-rep k_trace_packet profiles/round2_ncu_trace.json "workload description"
Needs the `ncu` CLI only (no GPU)."""
import csv, io, json, subprocess, sys

rep, rx, out, desc = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: (float(vals[hdr.index(k)].replace(",", "")), units[hdr.index(k)]) if k in hdr else (None, None)


def to_bytes(v, u):
    return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


keys = {"gpu__time_duration.sum": "duration", "smsp__inst_executed.sum": "warp_instructions",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "l1_writeback_pct",
        "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
        "launch__grid_size": "grid", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction"}
res = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else rx, "workload": desc, "report": rep}
for k, name in keys.items():
    v, u = get(k)
    res[name] = v
    if name == "duration":
        res["duration_unit"] = u
r, ru = get("dram__bytes_read.sum")
w, wu = get("dram__bytes_write.sum")
res["dram_bytes_read"], res["dram_bytes_write"] = to_bytes(r, ru), to_bytes(w, wu)
res["dram_bytes"] = None if r is None or w is None else res["dram_bytes_read"] + res["dram_bytes_write"]
try:
    res["git_sha"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
except OSError:
    res["git_sha"] = None
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
