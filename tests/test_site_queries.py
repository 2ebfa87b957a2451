"""Single-site queries on the GPU against answers of the unmodified reference (tests/golden/queries_sites.npz, made by
oracle/gen_golden_site_queries.py): PolyhedralSite.contains_point / contains_atom (reference polyhedral_site.py:431-520),
PolyhedralSiteCollection.sites_contain_points (polyhedral_site_collection.py:113-141), DynamicVoronoiSite.calculate_centre
(dynamic_voronoi_site.py:88-104); plus the known answers of reference tests/test_polyhedral_site.py:707-820."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"

_TETRAHEDRON = np.array([[0.4, 0.4, 0.4], [0.4, 0.6, 0.6], [0.6, 0.6, 0.4], [0.6, 0.4, 0.6]])
_OCTAHEDRON = np.array([[0.6, 0.5, 0.5], [0.4, 0.5, 0.5], [0.5, 0.6, 0.5], [0.5, 0.4, 0.5], [0.5, 0.5, 0.6], [0.5, 0.5, 0.4]])


def _sites(g):
    import site_analysis_b200 as sab
    sab.Site.reset_index()
    sites = []
    for i in range(len(g["n_vert"])):
        rc = g["ref_center"][i] if g["has_ref_center"][i] else None
        sites.append(sab.PolyhedralSite(list(range(int(g["offsets"][i]), int(g["offsets"][i + 1]))), reference_center=rc))
    return sites


@pytest.mark.gpu
def test_contains_point_matches_the_reference_on_every_golden_query():
    g = np.load(GOLDEN / "queries_sites.npz")
    sites = _sites(g)
    got = []
    k = 0
    for step in range(2):
        for i, site in enumerate(sites):
            site.notify_structure_changed(g["frames"][step], g["lattice"])
            while k < len(g["contains"]) and g["query_step"][k] == step and g["query_site"][k] == i:
                got.append(site.contains_point(g["query_point"][k]))
                k += 1
            if step == 0:          # the hull found at the first query is the reference's (same Qhull call on the same vertices)
                assert np.array_equal(site._face_topology_cache.face_simplices, g["simplices"][i, :g["n_face"][i]])
    assert k == len(g["contains"])
    assert np.array_equal(np.array(got), g["contains"])


@pytest.mark.gpu
def test_sites_contain_points_and_contains_atom():
    import site_analysis_b200 as sab
    g = np.load(GOLDEN / "queries_sites.npz")
    sites = _sites(g)
    coll = sab.site_collection.PolyhedralSiteCollection(sites)
    for step in range(2):          # the same call sequence as the generator: PBC caches advance through both frames
        for s in sites:
            s.notify_structure_changed(g["frames"][step], g["lattice"]); s._query_ready()
    assert coll.sites_contain_points(g["scp_centres"], g["frames"][1], g["lattice"]) is bool(g["scp_true"]) is True
    assert coll.sites_contain_points(g["scp_moved"], g["frames"][1], g["lattice"]) is bool(g["scp_false"]) is False
    with pytest.raises(ValueError):
        coll.sites_contain_points(g["scp_centres"][:-1], g["frames"][1], g["lattice"])
    atom = sab.Atom(index=0)
    atom._frac_coords = g["scp_centres"][0]
    assert sites[0].contains_atom(atom) is True
    with pytest.warns(DeprecationWarning):
        sites[0].contains_point(g["scp_centres"][0], algo="sn")


@pytest.mark.gpu
def test_known_answers_of_the_reference_tests():
    """reference tests/test_polyhedral_site.py:707-762 (inside / outside a tetrahedron and an octahedron, 4 and 8 faces,
    answers after a coordinate update) and :465-481 (no vertex coordinates -> RuntimeError)."""
    import site_analysis_b200 as sab
    for verts, n_faces in ((_TETRAHEDRON, 4), (_OCTAHEDRON, 8)):
        site = sab.PolyhedralSite(list(range(len(verts))))
        site.vertex_coords = verts.copy()
        assert site.contains_point(np.array([0.5, 0.5, 0.5])) is True
        assert site.contains_point(np.array([0.1, 0.1, 0.1])) is False
        assert site._face_topology_cache.face_simplices.shape == (n_faces, 3)
    site = sab.PolyhedralSite([0, 1, 2, 3])
    site.vertex_coords = _TETRAHEDRON.copy()
    site.contains_point(np.array([0.5, 0.5, 0.5]))
    site.vertex_coords = _TETRAHEDRON + np.array([0.1, 0.0, 0.0])           # topology kept, geometry updated
    assert site.contains_point(np.array([0.6, 0.5, 0.5])) is True
    assert site.contains_point(np.array([0.4, 0.5, 0.5])) is False
    assert site.contains_point(None, pbc_images=np.array([[0.1, 0.1, 0.1], [0.6, 0.5, 0.5]])) is True       # any image inside
    assert site.contains_point(None, pbc_images=np.array([[0.1, 0.1, 0.1], [0.8, 0.8, 0.9]])) is False
    with pytest.raises(RuntimeError):
        sab.PolyhedralSite([0, 1, 2, 3]).contains_point(np.array([0.5, 0.5, 0.5]))
    flat = sab.PolyhedralSite([0, 1, 2, 3])
    flat.vertex_coords = np.array([[0.1, 0.1, 0.1], [0.2, 0.1, 0.1], [0.3, 0.1, 0.1], [0.4, 0.1, 0.1]])
    with pytest.raises(RuntimeError, match="Degenerate vertex geometry"):
        flat.contains_point(np.array([0.2, 0.1, 0.1]))


@pytest.mark.gpu
def test_dynamic_voronoi_calculate_centre_bit_patterns():
    import site_analysis_b200 as sab
    g = np.load(GOLDEN / "queries_sites.npz")
    got = []
    for step in range(2):
        for i in range(len(g["n_vert"])):
            rc = g["ref_center"][i] if g["has_ref_center"][i] else None
            dv = sab.DynamicVoronoiSite(list(range(int(g["offsets"][i]), int(g["offsets"][i + 1]))), reference_center=rc)
            with pytest.raises(RuntimeError):
                dv.centre
            dv.calculate_centre(g["frames"][step], g["lattice"])
            got.append(dv.centre.copy())
    assert np.array_equal(np.array(got).view(np.int64), g["dv_centres_bits"])
