"""A/B timing of differently COMPILED builds of the backend (SAB200_LIB) on the bench workload's polyhedral shard.
usage (GPU box): python profiles/ab_libs.py lib1.so lib2.so:14=0,11=4 ...   (each build runs in its own process; results must be identical)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch, zlib
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    F = int(os.environ.get("AB_FRAMES", "125000"))
    g = syn.argyrodite_geometry(2)
    d_frac, d_lat = syn.argyrodite_trajectory(g, F, device="cuda")
    e = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, **syn.polyhedral_engine_kwargs(g, d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy()))
    for kv in filter(None, os.environ.get("SAB_OPTS", "").split(",")):
        k, v = kv.split("="); e.set_option(int(k), int(v))
    ts = []
    for i in range(8):
        r = e.assign_device(d_frac, d_lat); e._pending = []
        ts.append(e.last_info["assign_ns"] * 1e-6)
    crc = zlib.crc32(r.cpu().numpy().tobytes())
    print(f"min {min(ts[2:]):.4f} ms  median {float(np.median(ts[2:])):.4f} ms  deferred {e.last_info.get('deferred')} flagged {e.last_info['full27']} crc {crc:08x}", flush=True)
    sys.exit(0)
for spec in sys.argv[1:]:                       # path[:key=value,key=value]  (engine options, include/sab200.h SAB_OPT_*)
    lib, _, opts = spec.partition(":")
    env = dict(os.environ, SAB200_LIB=os.path.abspath(lib), SAB_OPTS=opts)
    out = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib) + (':' + opts if opts else ''):40s} {out.stdout.strip() or out.stderr.strip()[-400:]}", flush=True)
