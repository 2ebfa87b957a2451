"""Where the history fold's time goes (bench.py: value_with_history): device fold kernels vs host bookkeeping.
usage (GPU box): python profiles/diag_fold.py [frames]"""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
wl = bench.Workload("polyhedral")
d_frac, d_lat = wl.trajectory(F, 0, "cuda")
eng = wl.engine(d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy(), 0)
for i in range(3):
    eng.assign_device(d_frac, d_lat); eng.sync_history()
torch.cuda.synchronize()
ts = []
for i in range(5):
    eng.assign_device(d_frac, d_lat)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.sync_history()
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("sync_history wall ms:", [f"{t:.3f}" for t in ts])
eng.assign_device(d_frac, d_lat); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); eng.sync_history(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue())
from site_analysis_b200.folds import fold_rows
r = eng.assign_device(d_frac, d_lat); eng._pending = []
torch.cuda.synchronize()
for kw in (dict(want_pairs=True, want_recent=True, want_runs=False), dict(want_pairs=True, want_recent=False, want_runs=False),
           dict(want_pairs=False, want_recent=False, want_runs=False), dict(want_pairs=True, want_recent=True, want_runs=True)):
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fold_rows(r, eng.S, prev=eng.prev, **kw)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(kw, f"wall {dt:.3f} ms", out["info"], "runs", None if out["runs"] is None else len(out["runs"][0]), "pairs", None if out["pairs"] is None else len(out["pairs"][0]))
