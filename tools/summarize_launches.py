"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).
  python tools/summarize_launches.py gpurun_out/launches.csv"""
import collections, csv, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if "Metric Value" not in row or row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
    a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0, []])
    a[0] += 1; a[1] += v; a[2].append(v)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':90s} {'n':>5s} {'total us':>12s} {'avg us':>10s} {'share':>7s} {'min':>9s} {'max':>9s}")
for k, a in agg.items():
    print(f"{k:90s} {a[0]:5d} {a[1]:12.1f} {a[1]/a[0]:10.1f} {a[1]/tot*100:6.1f}% {min(a[2]):9.1f} {max(a[2]):9.1f}")
