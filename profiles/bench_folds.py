"""Timing of the post-processing fold (sab_fold_result, csrc/k_fold.cuh) on the bench workload's own result array.
usage (GPU box): python profiles/bench_folds.py [frames] > gpurun_out/folds.json

Reports, for the polyhedral argyrodite 2x2x2 result of `frames` frames (default 125000) and for the same array
tiled to 1,000,000 frames (BASELINE config 5 after the all-gather): runs / pairs found, device time of the whole
fold (CUDA events inside the library), algorithmic bytes (4 B per ion-frame read) per second against the measured HBM
peak, and the host alternatives it replaces (numpy fold_history of the engine; the per-site list walk of the
reference is O(sites x frames) and is only timed on a 2000-frame sample)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import torch

import site_analysis_b200 as sab
from site_analysis_b200 import AssignmentEngine, synthetic as syn
from site_analysis_b200.engine import fold_history

F = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, F, device="cuda")
kw = syn.polyhedral_engine_kwargs(g, d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy())
e = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, **kw)
rows = e.assign_device(d_frac, d_lat)
S, A = e.S, e.A
peak = 6543.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
out = {}
for label, d_rows in (("bench_shard", rows), ("config5_1M_frames", rows.repeat((1000000 + F - 1) // F, 1)[:1000000].contiguous())):
    n = d_rows.shape[0]
    ts, res = [], None
    for i in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sab.fold_rows(d_rows, S)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    kernel_ms = res["info"]["kernel_ns"] * 1e-6
    rec = dict(frames=n, atoms=A, sites=S, runs=int(len(res["runs"][0])), pairs=int(len(res["pairs"][0])),
               fold_device_ms=kernel_ms, fold_call_ms_incl_alloc_and_d2h=min(ts[1:]), launches=res["info"]["launches"],
               roofline=dict(bound="hbm", bytes=4 * n * A, achieved=4 * n * A / (kernel_ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                             frac=4 * n * A / (kernel_ms * 1e-3) / 1e9 / peak))
    h = d_rows.cpu().numpy()
    t0 = time.perf_counter()
    recent = np.full((A, 2), -1, dtype=np.int32); prev = np.full(A, -2, dtype=np.int32)
    tr = [dict() for _ in range(S)]
    fold_history(recent, prev, tr, h, True)
    rec["host_numpy_transition_fold_ms"] = (time.perf_counter() - t0) * 1e3
    got = {}
    for a, b, c in zip(*[x.tolist() for x in res["pairs"][:3]]):
        got.setdefault(a, {})[b] = c
    rec["pairs_equal_host_fold"] = got == {s: d for s, d in enumerate(tr) if d} and \
        all(list(got[s]) == list(tr[s]) for s in got)
    if label == "bench_shard":
        import folds_oracle as fo
        m = min(n, 2000)
        t0 = time.perf_counter()
        want = fo.runs_from_rows(h[:m])
        rec["python_run_extraction_ms_per_2000_frames"] = (time.perf_counter() - t0) * 1e3
        small = sab.fold_rows(h[:m], S)
        rec["runs_equal_oracle_on_sample"] = all(np.array_equal(x, y) for x, y in zip(small["runs"], want))
    out[label] = rec
    print(label, json.dumps(rec), file=sys.stderr, flush=True)
print(json.dumps(out))
