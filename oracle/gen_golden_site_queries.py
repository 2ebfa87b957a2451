"""Golden answers of the reference's single-site queries, from the UNMODIFIED reference:

* ``PolyhedralSite.contains_point`` (polyhedral_site.py:431-485) for tetrahedra, octahedra and cubes against points
  inside, outside, on faces / edges / vertices and 1e-4 .. 1e-16 off them, through the real call sequence
  ``notify_structure_changed`` -> ``contains_point`` (PBC-shift caches and face topology as the reference leaves them);
* ``PolyhedralSiteCollection.sites_contain_points`` (polyhedral_site_collection.py:113-141);
* ``DynamicVoronoiSite.calculate_centre`` (dynamic_voronoi_site.py:88-104), centres as float64 bit patterns.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/gen_golden_site_queries.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refcases  # noqa: E402

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def shapes(rng):
    tet = np.array([[0.4, 0.4, 0.4], [0.4, 0.6, 0.6], [0.6, 0.6, 0.4], [0.6, 0.4, 0.6]])          # tests/test_polyhedral_site.py:680
    octa = np.array([[0.6, 0.5, 0.5], [0.4, 0.5, 0.5], [0.5, 0.6, 0.5], [0.5, 0.4, 0.5], [0.5, 0.5, 0.6], [0.5, 0.5, 0.4]])
    cube = np.array([[x, y, z] for x in (0.35, 0.65) for y in (0.35, 0.65) for z in (0.35, 0.65)])
    out = []
    for base in (tet, octa, cube):
        for k in range(6):
            v = base + rng.normal(scale=0.01, size=base.shape)
            v = v + rng.uniform(-0.5, 0.5, size=3)              # anywhere in the cell, across its faces
            out.append(v)
    return out


def main():
    refcases.import_reference()
    from site_analysis.polyhedral_site import PolyhedralSite
    from site_analysis.polyhedral_site_collection import PolyhedralSiteCollection
    from site_analysis.dynamic_voronoi_site import DynamicVoronoiSite
    from site_analysis.site import Site
    rng = np.random.default_rng(20261020)
    lattice = np.array([[9.7, 0.3, -0.2], [0.1, 10.4, 0.5], [-0.3, 0.2, 11.1]])
    verts = shapes(rng)
    n_vert = np.array([len(v) for v in verts], dtype=np.int32)
    vmax = int(n_vert.max())
    # one "structure" holding every polyhedron's vertices (wrapped into [0, 1)), two frames (the second a small displacement)
    frames = []
    for step in range(2):
        allc = np.vstack([v + (rng.normal(scale=0.004, size=v.shape) if step else 0.0) for v in verts]) % 1.0
        frames.append(allc)
    offs = np.concatenate([[0], np.cumsum(n_vert)])
    Site.reset_index()
    sites = []
    for i, v in enumerate(verts):
        centre = v.mean(axis=0) % 1.0
        rc = centre if i % 3 != 2 else None                   # every third site: legacy (no reference centre)
        sites.append(PolyhedralSite(list(range(int(offs[i]), int(offs[i + 1]))), reference_center=rc))
    queries, answers, simp_out = [], [], []
    for step, allc in enumerate(frames):
        for i, site in enumerate(sites):
            site.notify_structure_changed(allc, lattice)
            v = allc[offs[i]:offs[i + 1]]
            rel = v - v[0]; rel -= np.round(rel)
            pts = []
            for q in range(24):
                w = rng.random(len(v))
                kind = q % 6
                if kind == 1: w[rng.integers(len(v))] = 0.0
                elif kind == 2: w = np.eye(len(v))[rng.integers(len(v))]
                elif kind == 3: w[rng.choice(len(v), len(v) - 2, replace=False)] = 0.0
                w = w / w.sum()
                p = v[0] + w @ rel
                if kind == 4: p = p + rng.normal(size=3) * 10.0 ** (-rng.integers(2, 17))
                if kind == 5: p = rng.random(3)
                pts.append(p % 1.0)
            for p in pts:
                queries.append((step, i, p))
                answers.append(bool(site.contains_point(p)))
            if step == 0:
                simp_out.append(np.asarray(site._face_topology_cache.face_simplices, dtype=np.int32))
    # sites_contain_points: all centres (True), then one displaced point (False)
    coll = PolyhedralSiteCollection(sites)
    centres = np.array([(frames[1][offs[i]:offs[i + 1]] - np.round(frames[1][offs[i]:offs[i + 1]] - frames[1][offs[i]])).mean(axis=0) % 1.0
                        for i in range(len(sites))])
    scp_true = bool(coll.sites_contain_points(centres, frames[1], lattice))
    moved = centres.copy(); moved[3] = (moved[3] + 0.31) % 1.0
    scp_false = bool(coll.sites_contain_points(moved, frames[1], lattice))
    fmax = max(len(s) for s in simp_out)
    simp = np.zeros((len(sites), fmax, 3), dtype=np.int32)
    for i, s in enumerate(simp_out):
        simp[i, :len(s)] = s
    # dynamic Voronoi centres
    dv_centres = []
    for step, allc in enumerate(frames):
        for i in range(len(sites)):
            rc = sites[i].reference_center
            dv = DynamicVoronoiSite(list(range(int(offs[i]), int(offs[i + 1]))), reference_center=rc)
            dv.calculate_centre(allc, lattice)
            dv_centres.append(dv._centre_coords.copy())
    np.savez_compressed(
        GOLDEN / "queries_sites.npz", lattice=lattice, frames=np.array(frames), n_vert=n_vert, offsets=offs.astype(np.int32),
        has_ref_center=np.array([s.reference_center is not None for s in sites]),
        ref_center=np.array([s.reference_center if s.reference_center is not None else np.full(3, np.nan) for s in sites]),
        query_step=np.array([q[0] for q in queries], dtype=np.int32), query_site=np.array([q[1] for q in queries], dtype=np.int32),
        query_point=np.array([q[2] for q in queries]), contains=np.array(answers), simplices=simp,
        n_face=np.array([len(s) for s in simp_out], dtype=np.int32), scp_centres=centres, scp_moved=moved,
        scp_true=np.array(scp_true), scp_false=np.array(scp_false),
        dv_centres_bits=np.array(dv_centres).view(np.int64))
    print(f"wrote queries_sites.npz: {len(queries)} point queries ({sum(answers)} inside), sites_contain_points {scp_true}/{scp_false}, "
          f"{len(dv_centres)} dynamic-Voronoi centres")


if __name__ == "__main__":
    main()
