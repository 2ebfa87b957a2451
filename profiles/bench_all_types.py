"""Secondary measurement: all four site types on the synthetic argyrodite 2x2x2 geometry
(S = 1056 sites, A = 192 ions), device-resident frames, CUDA-event kernel time from the library,
next to the single-thread C oracle port on a sample of the same frames, with a parity check.
usage (GPU box): python profiles/bench_all_types.py [frames]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import torch
import oracle
from site_analysis_b200 import AssignmentEngine, synthetic as syn

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
NPAR = 300
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, F, device="cuda")
frame0, lat0 = d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy()
S = len(g.site_centres)
slots = [list(map(int, t)) for t in g.tetrahedra]
h_frac, h_lat = d_frac[:NPAR].cpu().numpy(), d_lat[:NPAR].cpu().numpy()
cases = {
    "polyhedral": (dict(**syn.polyhedral_engine_kwargs(g, frame0, lat0)),
                   lambda: __import__("bench").oracle_for(g, frame0, lat0)()),
    "voronoi": (dict(centres=g.site_centres), lambda: oracle.SiteOracle("voronoi", g.mobile_idx, centres=g.site_centres)),
    "dynamic_voronoi": (dict(slots=slots, reference_centres=list(g.site_centres)),
                        lambda: oracle.SiteOracle("dynamic_voronoi", g.mobile_idx, ref_idx=np.array(slots, dtype=np.int32),
                                                  n_ref=np.full(S, 4, np.int32), has_ref_center=np.ones(S, np.uint8),
                                                  ref_center=g.site_centres)),
    "spherical": (dict(centres=g.site_centres, rcut=1.5),
                  lambda: oracle.SiteOracle("spherical", g.mobile_idx, centres=g.site_centres, rcut=1.5)),
}
out = {}
for kind, (kw, make_oracle) in cases.items():
    eng = AssignmentEngine(kind, g.n_atoms, g.mobile_idx, **kw)
    got = eng.assign_device(d_frac[:NPAR].contiguous(), d_lat[:NPAR].contiguous()).cpu().numpy()
    o = make_oracle()
    t0 = time.perf_counter(); want = o.run(h_frac, h_lat, positions=True); t_cpu = time.perf_counter() - t0
    eng = AssignmentEngine(kind, g.n_atoms, g.mobile_idx, **kw)
    ts, tot = [], []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.assign_device(d_frac, d_lat); eng._pending = []
        torch.cuda.synchronize(); tot.append(time.perf_counter() - t0)
        ts.append((eng.last_info["assign_ns"] + eng.last_info["wrap_ns"]) * 1e-9)
    out[kind] = {"frames": F, "mismatches_vs_oracle": int((got != want).sum()), "parity_frames": NPAR,
                 "gpu_kernel_ion_frames_per_s": F * g.n_mobile / min(ts[1:]),
                 "gpu_call_ion_frames_per_s": F * g.n_mobile / min(tot[1:]),
                 "ambiguous_atom_frames": eng.last_info["ambiguous"],
                 "cpu_port_1thread_ion_frames_per_s": NPAR * g.n_mobile / t_cpu}
    print(kind, json.dumps(out[kind]), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "all_types.json"), "w"), indent=1)
