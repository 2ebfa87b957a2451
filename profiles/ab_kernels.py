"""A/B timing of the polyhedral kernel variants on identical device-resident data.
usage (GPU box): python profiles/ab_kernels.py [frames]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from site_analysis_b200 import AssignmentEngine, _native as nat, synthetic as syn

F = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, F, device="cuda")
kw = syn.polyhedral_engine_kwargs(g, d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy())
ref = None
VARIANTS = [("stream gen 4 (round 1)", {nat.OPT_STREAM_GEN: 4}), ("stream gen 5 (default)", {}),
            ("gen 5, 3 producer warps (288-thread build)", {nat.OPT_STREAM_PRODUCERS: 3}),
            ("gen 5, 4 raw, 3 rows", {nat.OPT_STREAM5_RAW: 4, nat.OPT_STREAM5_ROWS: 3}),
            ("gen 5, 4 raw, 4 rows, records via L1", {nat.OPT_STREAM5_RAW: 4, nat.OPT_STREAM5_ROWS: 4, nat.OPT_STREAM_REC_SMEM: 0}),
            ("gen 5, 3 raw, 5 rows", {nat.OPT_STREAM5_RAW: 3, nat.OPT_STREAM5_ROWS: 5}),
            ("gen 5, 3 raw, 3 rows", {nat.OPT_STREAM5_RAW: 3, nat.OPT_STREAM5_ROWS: 3}),
            ("gen 5, records via L1", {nat.OPT_STREAM_REC_SMEM: 0}),
            ("gen 5, 1 producer warp", {nat.OPT_STREAM_PRODUCERS: 1}),
            ("gen 5, 2 CTAs/SM, 4 raw, 6 rows", {nat.OPT_STREAM5_BLOCKS: 2}),
            ("gen 5, chunk 128", {nat.OPT_SEQ_CHUNK: 128}),
            ("tile kernel + wrap scan", {nat.OPT_POLY_KERNEL: 5}), ("stream kernel gen 4 again", {nat.OPT_STREAM_GEN: 4}),
            ("gen 5, fp64 only", {nat.OPT_FP32_FILTER: 0}),
            ("seq1 + wrap scan", {nat.OPT_POLY_KERNEL: 4})]
if len(sys.argv) > 2:
    VARIANTS = VARIANTS[:int(sys.argv[2])]
for label, opts in VARIANTS:
    e = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, **kw)
    for k, v in opts.items():
        e.set_option(k, v)
    ts = []
    for i in range(5):
        r = e.assign_device(d_frac, d_lat)
        e._pending = []
        ts.append((e.last_info["assign_ns"] + e.last_info["wrap_ns"]) * 1e-6)
    r = r.cpu().numpy()
    if ref is None:
        ref = r
    print(f"{label:32s} wrap scan + assign ms: min {min(ts[1:]):.3f} median {np.median(ts[1:]):.3f}   identical={np.array_equal(r, ref)}  fallback frames={e.last_info['full27']} deferred={e.last_info.get('deferred')} scan_redo={e.last_info.get('scan_redo')}", flush=True)
