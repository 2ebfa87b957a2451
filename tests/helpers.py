"""Build site_analysis_b200 objects from golden fixtures (shared by CPU and GPU tests)."""
from __future__ import annotations

import numpy as np


def sites_and_atoms_from_case(case, seed_topology=True):
    """``seed_topology``: give polyhedral sites the Qhull face triangles the reference ended up with
    (the fixture's ``simplices``), i.e. sites that already own a face-topology cache.  For
    non-simplicial polyhedra (cubes) Qhull's choice of quad diagonals depends on the geometry of
    the frame in which the reference happened to query the site first; seeding removes that
    history dependence from the comparison (guard G3)."""
    import site_analysis_b200 as sab
    kind = str(case["kind"])
    labels = [str(l) or None for l in case["labels"]]
    sites = []
    sab.Site.reset_index()
    S = len(case["site_index"])
    for s in range(S):
        if kind == "SphericalSite":
            site = sab.SphericalSite(case["centres"][s].copy(), float(case["rcut"][s]), label=labels[s])
        elif kind == "VoronoiSite":
            site = sab.VoronoiSite(case["centres"][s].copy(), label=labels[s])
        elif kind == "PolyhedralSite":
            nv = int(case["n_vert"][s])
            rc = case["ref_center"][s].copy() if case["has_ref_center"][s] else None
            site = sab.PolyhedralSite([int(v) for v in case["vert_idx"][s, :nv]], label=labels[s], reference_center=rc)
            if case["primed"][s]:
                site._pbc_image_shifts = case["shift0"][s, :nv].copy()
                site._pbc_cached_raw_frac = case["cached_raw0"][s, :nv].copy()
            if seed_topology and "n_face" in case and case["n_face"][s] > 0:
                site._face_topology_cache = sab.polyhedral_site.FaceTopology(case["simplices"][s, :case["n_face"][s]])
        else:
            nr = int(case["n_ref"][s])
            rc = case["ref_center"][s].copy() if case["has_ref_center"][s] else None
            site = sab.DynamicVoronoiSite([int(v) for v in case["ref_idx"][s, :nr]], label=labels[s], reference_center=rc)
        site.index = int(case["site_index"][s])
        sites.append(site)
    atoms = [sab.Atom(index=int(i)) for i in case["atom_idx"]]
    return sites, atoms


def trajectory_from_case(case, use_case_rank=True, seed_topology=True):
    import site_analysis_b200 as sab
    sites, atoms = sites_and_atoms_from_case(case, seed_topology)
    t = sab.Trajectory(sites=sites, atoms=atoms)
    if use_case_rank and "rank_table" in case:
        t.site_collection.rank_table = np.asarray(case["rank_table"], dtype=np.int32)
    return t


def transitions_of(trajectory) -> dict:
    return {s.index: dict(s.transitions) for s in trajectory.sites if s.transitions}


def has_jump(case) -> bool:
    return "variant" in case and "jump=True" in str(case["variant"])
