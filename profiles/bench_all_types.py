"""Secondary measurement: every site type, on the BASELINE.json configurations that name them.

  polyhedral / voronoi / dynamic_voronoi / spherical   synthetic argyrodite 2x2x2 (S = 1056, A = 192, N = 416)
  config3_voronoi_4x4x4                                 BASELINE config 3: N = 3328, A = 1536, S = 8448, 100k frames
  config4_garnet_dynvor                                 BASELINE config 4: LLZO 2x2x2, N = 1536, A = 448, S = 576, 200k frames

Device-resident frames, CUDA-event kernel time from the library, whole-call wall time, a parity check of the
first frames against the C oracle, and the HBM roofline of the assignment pass: algorithmic bytes per frame
24 (A + M) + 72 + 4 A (M = referenced framework atoms) over the kernel time, against the
measured HBM peak in MEASURED_PEAKS.json.

usage (GPU box): python profiles/bench_all_types.py [frames-for-the-2x2x2-cases] [case-name ...]"""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import oracle
from site_analysis_b200 import AssignmentEngine, synthetic as syn

F2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ONLY = sys.argv[2:]
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6543.1


def padded(slots):
    r = np.zeros((len(slots), max(map(len, slots))), dtype=np.int32)
    for i, s in enumerate(slots):
        r[i, :len(s)] = s
    return r, np.array([len(s) for s in slots], dtype=np.int32)


def cases():
    g = syn.argyrodite_geometry(2)
    S = len(g.site_centres)
    slots = [list(map(int, t)) for t in g.tetrahedra]

    def poly_kw(frame0, lat0):
        return syn.polyhedral_engine_kwargs(g, frame0, lat0)

    def poly_oracle(frame0, lat0):
        return __import__("bench").oracle_for(g, frame0, lat0)()
    yield ("polyhedral", g, F2, "polyhedral", poly_kw, poly_oracle, 300, 192, 20260105)
    yield ("voronoi", g, F2, "voronoi", lambda *_: dict(centres=g.site_centres),
           lambda *_: oracle.SiteOracle("voronoi", g.mobile_idx, centres=g.site_centres), 300, 0, 20260105)
    yield ("dynamic_voronoi", g, F2, "dynamic_voronoi", lambda *_: dict(slots=slots, reference_centres=list(g.site_centres)),
           lambda *_: oracle.SiteOracle("dynamic_voronoi", g.mobile_idx, ref_idx=np.array(slots, dtype=np.int32),
                                        n_ref=np.full(S, 4, np.int32), has_ref_center=np.ones(S, np.uint8),
                                        ref_center=g.site_centres), 300, 192, 20260105)
    yield ("spherical", g, F2, "spherical", lambda *_: dict(centres=g.site_centres, rcut=1.5),
           lambda *_: oracle.SiteOracle("spherical", g.mobile_idx, centres=g.site_centres, rcut=1.5), 300, 0, 20260105)
    if not ONLY or "config3_voronoi_4x4x4" in ONLY:
        g3 = syn.argyrodite_geometry(4)
        yield ("config3_voronoi_4x4x4", g3, 100000, "voronoi", lambda *_: dict(centres=g3.site_centres),
               lambda *_: oracle.SiteOracle("voronoi", g3.mobile_idx, centres=g3.site_centres), 6, 0, 20260103)
    if not ONLY or "config4_garnet_dynvor" in ONLY:
        g4 = syn.garnet_geometry(2)
        ref, n_ref = padded(g4.reference_slots)
        S4 = len(g4.site_centres)
        m4 = len({a for s in g4.reference_slots for a in s})
        yield ("config4_garnet_dynvor", g4, 200000, "dynamic_voronoi",
               lambda *_: dict(slots=g4.reference_slots, reference_centres=list(g4.site_centres)),
               lambda *_: oracle.SiteOracle("dynamic_voronoi", g4.mobile_idx, ref_idx=ref, n_ref=n_ref,
                                            has_ref_center=np.ones(S4, np.uint8), ref_center=g4.site_centres), 100, m4, 20260104)


out = {}
for name, g, F, kind, make_kw, make_oracle, npar, n_fw, seed in cases():
    if ONLY and name not in ONLY:
        continue
    d_frac, d_lat = syn.synthetic_trajectory(g, F, device="cuda", seed=seed)
    frame0, lat0 = d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy()
    kw = make_kw(frame0, lat0)
    h_frac, h_lat = d_frac[:npar].cpu().numpy(), d_lat[:npar].cpu().numpy()
    eng = AssignmentEngine(kind, g.n_atoms, g.mobile_idx, **kw)
    got = eng.assign_device(d_frac[:npar].contiguous(), d_lat[:npar].contiguous()).cpu().numpy()
    o = make_oracle(frame0, lat0)
    t0 = time.perf_counter(); want = o.run(h_frac, h_lat, positions=True); t_cpu = time.perf_counter() - t0
    eng = AssignmentEngine(kind, g.n_atoms, g.mobile_idx, **kw)
    ts, tot = [], []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.assign_device(d_frac, d_lat); eng._pending = []
        torch.cuda.synchronize(); tot.append(time.perf_counter() - t0)
        ts.append((eng.last_info["assign_ns"] + eng.last_info["wrap_ns"]) * 1e-9)
    A = g.n_mobile
    bytes_per_frame = 24 * (A + n_fw) + 72 + 4 * A
    t_k = min(ts[1:])
    out[name] = {"kind": kind, "frames": F, "atoms_per_frame": g.n_atoms, "mobile_atoms": A, "sites": len(g.site_centres),
                 "mismatches_vs_oracle": int((got != want).sum()), "parity_frames": npar,
                 "gpu_kernel_ms": t_k * 1e3,
                 "gpu_kernel_ion_frames_per_s": F * A / t_k,
                 "gpu_call_ion_frames_per_s": F * A / min(tot[1:]),
                 "roofline": {"bound": "hbm", "bytes_per_frame": bytes_per_frame, "achieved": F * bytes_per_frame / t_k / 1e9,
                              "peak": PEAK, "unit": "GB/s", "frac": F * bytes_per_frame / t_k / 1e9 / PEAK},
                 "ambiguous_atom_frames": eng.last_info["ambiguous"], "exhaustive_fallback": eng.last_info["full27"],
                 "cell_lists": eng._cells,
                 "cpu_port_1thread_ion_frames_per_s": npar * A / t_cpu}
    print(name, json.dumps(out[name]), flush=True)
    del d_frac, d_lat, eng
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "all_types.json"), "w"), indent=1)
