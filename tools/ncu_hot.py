"""Top stalled SASS instructions of one kernel from an .ncu-rep (needs the `ncu` CLI; no GPU).
  python tools/ncu_hot.py report.ncu-rep kernel_regex [launch_index] [top_n]"""
import csv, io, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:200])
end = next((i for i in range(1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:end]))))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"] or 0) for r in rows)
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print("total samples", tot, "instructions", len(rows), "warp-instr executed", sum(int(r["Instructions Executed"] or 0) for r in rows))
print("stall totals:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot})
order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(order):
    r = rows[i]
    s = int(r["# Samples"] or 0)
    why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100*s/max(tot,1):5.1f}% exec={r['Instructions Executed']:>9s} thr={r['Avg. Threads Executed']:>5s} {r['Source'].strip()[:80]:80s} {[(w, n) for n, w in why if n]}")
