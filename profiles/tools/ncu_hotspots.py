"""Per-source-line summary of an `ncu --page source --csv --print-source sass,cuda` dump:
instructions executed, stall samples and their dominant reasons, average active lanes.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python profiles/tools/ncu_hotspots.py src.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
rows = list(csv.reader(open(path)))
files = {}
cur_file = None
hdr = None
agg = defaultdict(lambda: defaultdict(float))
sass_rows = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No" or (r and r[0] == "Address"):
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    d = dict(zip(hdr, r))
    if "Line No" in d and d["Line No"].isdigit() and cur_file:
        key = (cur_file, int(d["Line No"]), d.get("Source", "")[:90].strip())
        a = agg[key]
        for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                  "stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_mio", "stall_branch_resolving",
                  "stall_not_selected", "stall_lg", "stall_dispatch", "stall_no_inst"):
            try:
                a[k] += float(d.get(k, 0) or 0)
            except ValueError:
                pass
tot_i = sum(a["Instructions Executed"] for a in agg.values())
tot_s = sum(a["# Samples"] for a in agg.values())
print(f"warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
def show(items):
    for (f, ln, src), a in items:
        lanes = a["Thread Instructions Executed"] / a["Instructions Executed"] if a["Instructions Executed"] else 0
        reasons = sorted(((a[k], k[6:]) for k in a if k.startswith("stall_")), reverse=True)[:3]
        rs = " ".join(f"{n}:{100 * v / max(a['# Samples'], 1):.0f}%" for v, n in reasons if v > 0)
        print(f"{100 * a['Instructions Executed'] / tot_i:5.1f}%i {100 * a['# Samples'] / tot_s:5.1f}%s lanes {lanes:4.1f} smem {a['L1 Wavefronts Shared']:.2e}/{a['L1 Wavefronts Shared Ideal']:.2e} {f[13:]:14s} L{ln:4d} {src[:70]:70s} | {rs}")
print("-- by stall samples")
show(sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top])
print("-- by instructions")
show(sorted(agg.items(), key=lambda kv: -kv[1]["Instructions Executed"])[:top])
