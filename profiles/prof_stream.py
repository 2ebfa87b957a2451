"""One short run of the polyhedral stream kernel for ncu (options from the environment).
usage (GPU box): [SAB_OPTS="key=value,..."] python profiles/prof_stream.py [frames] [passes]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from site_analysis_b200 import AssignmentEngine, synthetic as syn

F = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, F, device="cuda")
e = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, **syn.polyhedral_engine_kwargs(g, d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy()))
for kv in filter(None, os.environ.get("SAB_OPTS", "").split(",")):
    k, v = kv.split("=")
    e.set_option(int(k), int(v))
for i in range(P):
    r = e.assign_device(d_frac, d_lat)
    e._pending = []
    print(f"pass {i}: assign {e.last_info['assign_ns'] * 1e-6:.3f} ms, deferred {e.last_info['deferred']}, fallback frames {e.last_info['full27']}", flush=True)
print("checksum", int(r.to(dtype=__import__('torch').int64).sum().item()))
