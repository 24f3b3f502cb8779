"""Per-opcode summary of an `ncu --page source --csv --print-source sass` dump: instructions, stall samples, stall reasons.
  ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_source_ops.py src.csv <units>"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for b in blocks:
    hdr, data = b["hdr"], b["data"]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ci["# Samples"]]) for r in data)
    inst = sum(int(r[ci["Instructions Executed"]]) for r in data)
    print("==", b["name"][:90]); print("samples", tot, "warp-instructions", inst, "per unit", round(inst / units, 1))
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print({h[6:]: sum(int(r[ci[h]]) for r in data) for h in reasons})
    byop, sm = collections.Counter(), collections.Counter()
    for r in data:
        toks = [o for o in r[ci["Source"]].strip().split() if not o.startswith("@")]
        op = ".".join(toks[0].split(".")[:3]) if toks[0].startswith(("LDG", "STG", "LDS", "STS", "SHFL")) else toks[0].split(".")[0]
        byop[op] += int(r[ci["Instructions Executed"]]); sm[op] += int(r[ci["# Samples"]])
    for op, c in byop.most_common(22):
        print(f"  {op:22s} {c / units:8.1f} /unit   samples {sm[op]}")
