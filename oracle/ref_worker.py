"""Worker process of bench.py's reference arm: runs the UNMODIFIED reference (``oracle/_ref/site_analysis``, a verbatim copy
of /root/reference made by oracle/build_ref.py) on one contiguous slice of frames, the way the reference's own benchmark does
(tests/benchmark_all_site_types.py:92-103: one warm-up frame for the numba JIT and the caches, then
``trajectory.analyse_structure(structure)`` per frame), and prints one JSON line {"frames", "atoms", "seconds"}.

MEASUREMENT INFRASTRUCTURE ONLY (see oracle/build_ref.py).  The parent passes the frames as .npy files and sets
NUMBA_NUM_THREADS=1, so P workers use P cores (CPU baseline (ii)).

usage: python oracle/ref_worker.py <site kind> <tables.npz> <frac.npy> <lattice.npy> <first frame> <frames>
"""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def main():
    kind, tables, frac_path, lat_path, first, count = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/sa_numba_cache")
    for p in (str(HERE / "refharness"), str(HERE / "_ref")):
        sys.path.insert(0, p)
    import site_analysis  # noqa: F401  -- the reference
    from pymatgen.core import Lattice, Structure
    from site_analysis.atom import Atom
    from site_analysis.site import Site
    from site_analysis.trajectory import Trajectory
    t = np.load(tables, allow_pickle=False)
    frac = np.load(frac_path, mmap_mode="r")[first:first + count]
    lat = np.load(lat_path, mmap_mode="r")[first:first + count]
    frame0, lat0 = np.load(frac_path, mmap_mode="r")[0], np.load(lat_path, mmap_mode="r")[0]
    species = [str(s) for s in t["species"]]
    Site.reset_index()
    if kind == "polyhedral":
        from site_analysis.polyhedral_site import PolyhedralSite
        sites = [PolyhedralSite(vertex_indices=[int(v) for v in vs], reference_center=c) for vs, c in zip(t["slots"], t["centres"])]
        for s in sites:                       # the builder primes the PBC-shift caches from the build structure (site_factory.py:248-257)
            s.assign_vertex_coords(np.asarray(frame0), np.asarray(lat0))
    elif kind == "voronoi":
        from site_analysis.voronoi_site import VoronoiSite
        sites = [VoronoiSite(frac_coords=c) for c in t["centres"]]
    elif kind == "spherical":
        from site_analysis.spherical_site import SphericalSite
        sites = [SphericalSite(frac_coords=c, rcut=float(t["rcut"])) for c in t["centres"]]
    elif kind == "dynamic_voronoi":
        from site_analysis.dynamic_voronoi_site import DynamicVoronoiSite
        n_ref = t["n_ref"]
        sites = [DynamicVoronoiSite(reference_indices=[int(v) for v in vs[:n]], reference_center=c)
                 for vs, n, c in zip(t["slots"], n_ref, t["centres"])]
    else:
        raise SystemExit(f"unknown site kind {kind}")
    atoms = [Atom(index=int(i)) for i in t["mobile"]]
    traj = Trajectory(sites=sites, atoms=atoms)
    structures = [Structure(Lattice(np.asarray(lat[f])), species, np.asarray(frac[f])) for f in range(count)]
    traj.append_timestep(structures[0], t=0)  # warm-up: JIT + caches (benchmark_all_site_types.py:95)
    t0 = time.perf_counter()
    for s in structures[1:]:
        traj.analyse_structure(s)
    dt = time.perf_counter() - t0
    from site_analysis._compat import HAS_NUMBA
    print(json.dumps({"frames": count - 1, "atoms": len(atoms), "seconds": dt, "numba": bool(HAS_NUMBA)}))


if __name__ == "__main__":
    main()
