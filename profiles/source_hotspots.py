"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line.
usage: python profiles/source_hotspots.py export.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
fname, h = None, None
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h = r
        iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
        continue
    if h is None or len(r) <= iI or r[0] == "":
        continue
    try:
        ins, sm, ti = int(r[iI] or 0), int(r[iS] or 0), int(r[iT] or 0)
    except ValueError:
        continue
    a = agg.setdefault((fname, int(r[0]), r[1].strip()), [0, 0, 0])
    a[0] += ins; a[1] += sm; a[2] += ti
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"warp instructions {tot}, samples {ts}")
print("-- by instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / tot * 100:5.1f}%i {a[1] / ts * 100:5.1f}%s lanes {a[2] / max(a[0], 1):4.1f} {k[0][:24]:24s} L{k[1]:>4} {k[2][:90]}")
print("-- by stall samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
    print(f"{a[0] / tot * 100:5.1f}%i {a[1] / ts * 100:5.1f}%s lanes {a[2] / max(a[0], 1):4.1f} {k[0][:24]:24s} L{k[1]:>4} {k[2][:90]}")
