"""Time the UNMODIFIED reference (Python + numba) and the C oracle port on the same synthetic
argyrodite frames, single core, in the build container.  TEST INFRASTRUCTURE.  Writes
oracle/REFERENCE_SPEED.json, which DESIGN.md and bench.py quote to relate the port-based CPU
baseline to the real reference."""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent))
os.environ.setdefault("NUMBA_NUM_THREADS", "1")
import oracle  # noqa: E402
import refcases  # noqa: E402


def main(n_frames=1500):
    refcases.import_reference()
    from site_analysis.atom import Atom
    from site_analysis.polyhedral_site import PolyhedralSite
    from site_analysis.site import Site
    from site_analysis.trajectory import Trajectory
    from site_analysis_b200 import synthetic as syn
    geom = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(geom, n_frames, device="cpu")
    frac, lat = frac.numpy(), lat.numpy()
    structures = refcases.structures_from_arrays(geom.species, lat, frac)
    Site.reset_index()
    sites = [PolyhedralSite(vertex_indices=[int(v) for v in t], label=l, reference_center=c)
             for t, l, c in zip(geom.tetrahedra, geom.site_labels, geom.site_centres)]
    for s in sites:
        s.assign_vertex_coords(frac[0], lat[0])
    atoms = [Atom(index=int(i)) for i in geom.mobile_idx]
    traj = Trajectory(sites=sites, atoms=atoms)
    traj.append_timestep(structures[0], t=0)                    # warm-up: JIT + caches (benchmark_all_site_types.py:95)
    t0 = time.perf_counter()
    for s in structures[1:]:
        traj.analyse_structure(s)
    t_ref = time.perf_counter() - t0
    ref_rate = (n_frames - 1) * len(atoms) / t_ref

    case = refcases.site_definitions(Trajectory(sites=[PolyhedralSite(vertex_indices=[int(v) for v in t], reference_center=c)
                                                       for t, c in zip(geom.tetrahedra, geom.site_centres)], atoms=atoms))
    o = oracle.SiteOracle("polyhedral", geom.mobile_idx, vert_idx=case["vert_idx"], n_vert=case["n_vert"],
                          has_ref_center=case["has_ref_center"], ref_center=case["ref_center"])
    o.run(frac[:1], lat[:1])
    t0 = time.perf_counter()
    o.run_timed(frac[1:], lat[1:])
    t_port = time.perf_counter() - t0
    port_rate = (n_frames - 1) * len(atoms) / t_port
    out = {"workload": "synthetic argyrodite 2x2x2 polyhedral, %d frames" % n_frames, "cores": 1,
           "reference_python_numba_ion_frames_per_s": ref_rate, "c_oracle_port_ion_frames_per_s": port_rate,
           "port_over_reference": port_rate / ref_rate, "host": "build container (8 vCPU Xeon VM)",
           "numba_threads": 1}
    (HERE / "REFERENCE_SPEED.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1500)
