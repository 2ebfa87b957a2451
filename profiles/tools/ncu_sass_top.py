"""Top SASS instructions of an `ncu --page source --csv --print-source sass,cuda` dump by stall samples,
plus every synchronisation instruction.  usage: python profiles/tools/ncu_sass_top.py src.csv [top]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr, line, out = None, None, {}
def I(x):
    try:
        return int(x)
    except ValueError:
        return 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0].isdigit():
        line = int(r[0])
        continue
    try:
        ad = int(r[2], 16) & 0xfffff
    except ValueError:
        continue
    if ad not in out:
        st = {hdr[i][6:]: I(r[i]) for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not" not in hdr[i]}
        out[ad] = (ad, r[3], I(r[4]), I(r[7]), line, sorted(st.items(), key=lambda kv: -kv[1])[:2])
out = sorted(out.values())
tot, toti = sum(o[2] for o in out), sum(o[3] for o in out)
print(f"samples {tot}, warp instructions {toti}")
sync = [o for o in out if "PHASECHK" in o[1] or "NANOSLEEP" in o[1] or "BAR." in o[1]]
print("-- synchronisation")
for o in sync:
    if o[3] > 1000:
        print(f"{o[0]:05x} {o[1][:56]:56s} samples {o[2]:6d} ({100 * o[2] / tot:4.1f}%) inst {o[3]:10d} L{o[4]}")
print("-- top by samples")
for o in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"{o[0]:05x} {o[1][:56]:56s} samples {o[2]:6d} ({100 * o[2] / tot:4.1f}%) inst {o[3]:10d} L{o[4]} {o[5]}")
