"""Aggregate the per-instruction stall samples of an `ncu --page source --csv` dump (gz) per kernel."""
import csv, gzip, io, sys
from collections import defaultdict
txt = gzip.open(sys.argv[1], "rt").read()
kern = None
hdr = None
agg = {}
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern = row[1][:70]
        hdr = None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if hdr is None or kern is None:
        continue
    d = agg.setdefault(kern, defaultdict(float))
    for h, v in zip(hdr, row):
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                d[h] += float(v)
            except ValueError:
                pass
    try:
        d["_inst"] += float(row[hdr.index("Instructions Executed")])
    except ValueError:
        pass
seen = set()
for k, d in agg.items():
    if k in seen:
        continue
    seen.add(k)
    tot = sum(v for h, v in d.items() if h.startswith("stall_"))
    top = sorted(((v, h) for h, v in d.items() if h.startswith("stall_")), reverse=True)[:6]
    print(k)
    print("   warp-instr %.3g  " % d["_inst"] + "  ".join(f"{h[6:]} {100*v/tot:.0f}%" for v, h in top))
