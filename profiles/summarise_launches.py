"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.
usage: python profiles/summarise_launches.py gpurun_out/launches.csv "note" > profiles/<name>.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) <= iV:
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(r[iU], 1.0)
    k = r[iK].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("#", sys.argv[2] if len(sys.argv) > 2 else "")
print("# cold-cache, serialised launches under ncu: compare SHARES, not absolutes")
print("kernel,launches,total_ms,share")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k},{n},{ms:.4f},{ms / tot:.4f}")
