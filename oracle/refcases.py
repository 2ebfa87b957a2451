"""Helpers that run the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (writes the committed
fixtures under ``tests/golden/``) and ``oracle/validate_oracle.py`` (pins the C
oracle against the live reference).  Nothing here is imported by the product
package, ``bench.py`` or the ``-m gpu`` tests: ``/root/reference`` does not exist
on the GPU box.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE_ROOT = Path(os.environ.get("SA_REFERENCE_ROOT", "/root/reference"))


def import_reference():
    """Put the stubbed pymatgen/monty and the reference on sys.path; return the module."""
    if not REFERENCE_ROOT.exists():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/sa_numba_cache")
    for p in (str(REFERENCE_ROOT), str(HERE / "refharness")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import site_analysis  # noqa: F401  (the reference)
    return site_analysis


def structures_from_arrays(species, lattice, frac):
    from pymatgen.core import Lattice, Structure
    lattice = np.asarray(lattice, dtype=np.float64)
    out = []
    for f in range(frac.shape[0]):
        lat = lattice if lattice.ndim == 2 else lattice[f]
        out.append(Structure(Lattice(lat), species, frac[f]))
    return out


def argyrodite_reference_structure():
    """F-4 3m Li6PS5Cl reference with all five Li site types (reference
    tests/benchmark_all_site_types.py:29-46), 2x2x2 supercell."""
    from pymatgen.core import Lattice, Structure
    coords = np.array([
        [0.5, 0.5, 0.5], [0.9, 0.9, 0.6], [0.77, 0.585, 0.585], [0.25, 0.25, 0.25],
        [0.15, 0.15, 0.15], [0.0, 0.183, 0.183], [0.0, 0.0, 0.0], [0.75, 0.25, 0.25],
        [0.11824, 0.11824, 0.38176]])
    return Structure.from_spacegroup(
        sg="F-43m", lattice=Lattice.cubic(a=10.155),
        species=["P", "Li", "Mg", "Na", "Be", "K", "S", "S", "S"], coords=coords) * [2, 2, 2]


# ---------------------------------------------------------------------------
# extracting site definitions / state / results from reference objects
# ---------------------------------------------------------------------------

def _pad(rows, fill=-1, dtype=np.int32):
    width = max(len(r) for r in rows)
    out = np.full((len(rows), width), fill, dtype=dtype)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def site_definitions(trajectory) -> dict:
    """Snapshot of the static site tables + initial PBC-cache state of a reference Trajectory."""
    sites = trajectory.sites
    kind = type(sites[0]).__name__
    d: dict = {
        "kind": np.array(kind),
        "site_index": np.array([s.index for s in sites], dtype=np.int32),
        "labels": np.array([("" if s.label is None else str(s.label)) for s in sites]),
        "atom_idx": np.array([a.index for a in trajectory.atoms], dtype=np.int32),
    }
    if kind == "SphericalSite":
        d["centres"] = np.array([s.frac_coords for s in sites], dtype=np.float64)
        d["rcut"] = np.array([s.rcut for s in sites], dtype=np.float64)
    elif kind == "VoronoiSite":
        d["centres"] = np.array([s.frac_coords for s in sites], dtype=np.float64)
    elif kind == "PolyhedralSite":
        d["vert_idx"] = _pad([list(s.vertex_indices) for s in sites])
        d["n_vert"] = np.array([len(s.vertex_indices) for s in sites], dtype=np.int32)
        vmax = d["vert_idx"].shape[1]
        has_rc = np.array([s.reference_center is not None for s in sites])
        rc = np.full((len(sites), 3), np.nan)
        for i, s in enumerate(sites):
            if s.reference_center is not None:
                rc[i] = s.reference_center
        d["has_ref_center"] = has_rc
        d["ref_center"] = rc
        primed = np.zeros(len(sites), dtype=bool)
        shift0 = np.zeros((len(sites), vmax, 3), dtype=np.int64)
        raw0 = np.zeros((len(sites), vmax, 3), dtype=np.float64)
        for i, s in enumerate(sites):
            if s._pbc_image_shifts is not None and s._pbc_cached_raw_frac is not None:
                primed[i] = True
                shift0[i, :len(s.vertex_indices)] = s._pbc_image_shifts
                raw0[i, :len(s.vertex_indices)] = s._pbc_cached_raw_frac
        d["primed"] = primed
        d["shift0"] = shift0
        d["cached_raw0"] = raw0
    elif kind == "DynamicVoronoiSite":
        d["ref_idx"] = _pad([list(s.reference_indices) for s in sites])
        d["n_ref"] = np.array([len(s.reference_indices) for s in sites], dtype=np.int32)
        has_rc = np.array([s.reference_center is not None for s in sites])
        rc = np.full((len(sites), 3), np.nan)
        for i, s in enumerate(sites):
            if s.reference_center is not None:
                rc[i] = s.reference_center
        d["has_ref_center"] = has_rc
        d["ref_center"] = rc
    else:
        raise TypeError(kind)
    return d


def rank_table(trajectory):
    """The reference's own distance-ranked table (list positions), or None."""
    coll = trajectory.site_collection
    ranked = getattr(coll, "_distance_ranked_sites", None)
    if ranked is None:
        return None
    pos = {s.index: i for i, s in enumerate(trajectory.sites)}
    n = len(trajectory.sites)
    tab = np.zeros((n, max(n - 1, 0)), dtype=np.int32)
    for s in trajectory.sites:
        tab[pos[s.index]] = [pos[j] for j in ranked[s.index]]
    return tab


def results(trajectory) -> dict:
    """Per-atom site trajectories (site.index, -1 = None) and per-site transition counters."""
    at = np.array([[(-1 if v is None else v) for v in a.trajectory] for a in trajectory.atoms],
                  dtype=np.int32).T.copy()
    src, dst, cnt, order = [], [], [], []
    for s in trajectory.sites:
        for k, (dest, c) in enumerate(s.transitions.items()):
            src.append(s.index); dst.append(dest); cnt.append(c); order.append(k)
    out = {
        "atoms_traj": at,
        "tr_src": np.array(src, dtype=np.int32),
        "tr_dst": np.array(dst, dtype=np.int32),
        "tr_cnt": np.array(cnt, dtype=np.int64),
        "tr_order": np.array(order, dtype=np.int32),
        "recent": np.array([[(-1 if r is None else r) for r in a._recent_sites]
                            for a in trajectory.atoms], dtype=np.int32),
    }
    kind = type(trajectory.sites[0]).__name__
    if kind == "PolyhedralSite":
        nf = []
        simp = []
        for s in trajectory.sites:
            c = s._face_topology_cache
            if c is None:
                nf.append(-1); simp.append(np.zeros((0, 3), dtype=np.int32))
            else:
                nf.append(len(c.face_simplices)); simp.append(np.asarray(c.face_simplices, dtype=np.int32))
        fmax = max([len(x) for x in simp] + [1])
        arr = np.full((len(simp), fmax, 3), -1, dtype=np.int32)
        for i, x in enumerate(simp):
            arr[i, :len(x)] = x
        out["n_face"] = np.array(nf, dtype=np.int32)
        out["simplices"] = arr
    if kind == "DynamicVoronoiSite":
        out["final_centres"] = np.array([s.centre for s in trajectory.sites], dtype=np.float64)
    return out


def run_case(trajectory, species, lattice, frac, per_frame_hook=None) -> dict:
    """site defs (before) + run ``trajectory_from_structures`` + results (after)."""
    case = site_definitions(trajectory)
    rt = rank_table(trajectory)
    if rt is not None:
        case["rank_table"] = rt
    structures = structures_from_arrays(species, lattice, frac)
    if per_frame_hook is None:
        trajectory.trajectory_from_structures(structures)
    else:
        for t, s in enumerate(structures, 1):
            trajectory.append_timestep(s, t=t)
            per_frame_hook(t - 1, trajectory)
    case.update(results(trajectory))
    case["lattice"] = np.asarray(lattice, dtype=np.float64)
    case["species"] = np.array(species)
    return case
