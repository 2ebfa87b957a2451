"""Known-answer tests of the reference's own test-suite, re-expressed against the public API of
site_analysis_b200 (each cites the reference test it mirrors, paths relative to /root/reference/tests).
GPU tests run the assignment on the device; the object-level ones are CPU-only."""
import numpy as np
import pytest

import site_analysis_b200 as sab
from site_analysis_b200 import SimpleStructure as Structure

CUBIC5 = np.eye(3) * 5.0
CUBIC10 = np.eye(3) * 10.0


# ------------------------------------------------------------------ CPU: object behaviour

def test_polyhedral_shift_cache_first_call_and_small_move():
    """test_polyhedral_site.py:465-490: first call computes shifts, a small vibration keeps them."""
    sab.Site.reset_index()
    site = sab.PolyhedralSite(vertex_indices=[0, 1, 2, 3])
    frac1 = np.array([[0.1] * 3, [0.2] * 3, [0.3] * 3, [0.4] * 3])
    assert site._pbc_image_shifts is None
    site.assign_vertex_coords(frac1, CUBIC10)
    assert site._pbc_image_shifts is not None and site.vertex_coords is not None
    shifts = site._pbc_image_shifts.copy()
    site.assign_vertex_coords(frac1 + 0.005, CUBIC10)
    assert np.array_equal(site._pbc_image_shifts, shifts)
    assert np.allclose(site.vertex_coords, frac1 + 0.005)


def test_polyhedral_shift_cache_wrap_adjusts_shift_by_one():
    """test_polyhedral_site.py:505-521: 0.99 -> 0.01 changes the cached shift by +1, no recomputation."""
    site = sab.PolyhedralSite(vertex_indices=[0, 1, 2, 3])
    frac1 = np.array([[0.99, 0.5, 0.5], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]])
    site.assign_vertex_coords(frac1, CUBIC10)
    original = site._pbc_image_shifts.copy()
    frac2 = frac1.copy(); frac2[0, 0] = 0.01
    site.assign_vertex_coords(frac2, CUBIC10)
    assert site._pbc_image_shifts[0, 0] == original[0, 0] + 1
    assert np.isclose(site.vertex_coords[0, 0], 1.01)


def test_polyhedral_shift_cache_large_jump_recomputes():
    """test_polyhedral_site.py:492-503: a 0.4 displacement invalidates the cache (full recomputation)."""
    site = sab.PolyhedralSite(vertex_indices=[0, 1, 2, 3])
    frac1 = np.array([[0.1] * 3, [0.2] * 3, [0.3] * 3, [0.4] * 3])
    site.assign_vertex_coords(frac1, CUBIC10)
    site.assign_vertex_coords(frac1 + 0.4, CUBIC10)
    assert np.array_equal(site._pbc_image_shifts, np.zeros((4, 3), dtype=np.int64))      # legacy rule on the new coords
    assert np.allclose(site._pbc_cached_raw_frac, frac1 + 0.4)


def test_polyhedral_site_constructor_errors():
    """test_polyhedral_site.py (constructor validation): empty list, non-integers."""
    with pytest.raises(ValueError):
        sab.PolyhedralSite(vertex_indices=[])
    with pytest.raises(TypeError):
        sab.PolyhedralSite(vertex_indices=[0, 1.5, 2, 3])
    assert sab.PolyhedralSite(vertex_indices=np.array([4, 5, 6, 7])).vertex_indices == [4, 5, 6, 7]


def test_atom_recent_site_history():
    """test_atom.py (update_recent_site): only changes shift the two-deep history; reset clears it."""
    a = sab.Atom(index=3)
    assert a._recent_sites == [None, None] and a.most_recent_site is None
    a.update_recent_site(5); a.update_recent_site(5); a.update_recent_site(7)
    assert a._recent_sites == [7, 5]
    a.assign_coords(np.array([[0, 0, 0], [0, 0, 0], [0, 0, 0], [1.25, -0.25, 2.0]]))
    assert np.allclose(a.frac_coords, [0.25, 0.75, 0.0])
    a.reset()
    assert a._recent_sites == [None, None] and a.trajectory == [] and a.in_site is None
    with pytest.raises(AttributeError):
        a.frac_coords


# ------------------------------------------------------------------ GPU: assignments

@pytest.mark.gpu
def test_spherical_site_with_unwrapped_coordinates():
    """test_coordinate_integration.py:20-41."""
    st = Structure(CUBIC5, ["Li"], [[1.1, 1.2, 1.3]])
    t = (sab.TrajectoryBuilder().with_structure(st).with_mobile_species("Li")
         .with_spherical_sites(centres=[[0.1, 0.2, 0.3]], radii=0.1).build())
    t.analyse_structure(st)
    assert t.atoms[0].in_site == t.sites[0].index and t.sites[0].contains_atoms == [0]


@pytest.mark.gpu
def test_voronoi_site_with_unwrapped_coordinates():
    """test_coordinate_integration.py:43-74."""
    st = Structure(CUBIC5, ["Li"], [[1.1, 1.2, 1.3]])
    atoms = sab.atoms_from_structure(st, "Li")
    sites = [sab.VoronoiSite(np.array(c)) for c in ([0.1, 0.2, 0.3], [0.5, 0.5, 0.5], [0.8, 0.8, 0.8])]
    coll = sab.VoronoiSiteCollection(sites)
    coll.analyse_structure(atoms, st)
    assert [len(s.contains_atoms) for s in sites] == [1, 0, 0] and atoms[0].index in sites[0].contains_atoms


@pytest.mark.gpu
def test_polyhedral_site_with_unwrapped_coordinates():
    """test_coordinate_integration.py:76-110: Li at the centroid of an O4 tetrahedron, unwrapped by +1."""
    st = Structure(CUBIC5, ["Li", "O", "O", "O", "O"],
                   [[1.1, 1.1, 1.25], [0.0, 0.0, 0.0], [0.2, 0.2, 0.2], [0.2, 0.0, 0.4], [0.0, 0.2, 0.4]])
    atoms = sab.atoms_from_structure(st, "Li")
    site = sab.PolyhedralSite(vertex_indices=[1, 2, 3, 4])
    coll = sab.PolyhedralSiteCollection([site])
    coll.analyse_structure(atoms, st)
    assert site.contains_atoms == [atoms[0].index]


@pytest.mark.gpu
def test_dynamic_voronoi_site_with_unwrapped_coordinates():
    """test_coordinate_integration.py:112-150."""
    st = Structure(CUBIC5, ["Li", "O", "O", "O", "O"],
                   [[1.1, 1.2, 1.3], [1.0, 1.2, 1.3], [1.2, 1.2, 1.3], [1.5, 1.5, 1.5], [1.7, 1.5, 1.5]])
    atoms = sab.atoms_from_structure(st, "Li")
    target, other = sab.DynamicVoronoiSite([1, 2]), sab.DynamicVoronoiSite([3, 4])
    coll = sab.DynamicVoronoiSiteCollection([target, other])
    coll.analyse_structure(atoms, st)
    assert len(target.contains_atoms) == 1 and len(other.contains_atoms) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("re, o, centre", [
    ([[0.3, 0.5, 0.5], [0.7, 0.5, 0.5]], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]),
    ([[0.0, 0.0, 0.0], [0.0, 0.5, 0.0]], [0.0, 0.25, 0.0], [0.0, 0.25, 0.0]),
])
def test_dynamic_voronoi_sites_simple(re, o, centre):
    """test_dynamic_voronoi_trajectory.py:11-75."""
    st = Structure(CUBIC10, ["Re", "Re", "O"], re + [o])
    atom = sab.Atom(index=2)
    site = sab.DynamicVoronoiSite(reference_indices=[0, 1])
    t = sab.Trajectory(sites=[site], atoms=[atom])
    t.analyse_structure(st)
    assert atom.in_site == site.index and atom.index in site.contains_atoms
    np.testing.assert_array_almost_equal(site.centre, np.array(centre))


@pytest.mark.gpu
def test_dynamic_voronoi_sites_multiple():
    """test_dynamic_voronoi_trajectory.py:77-131: V-shaped Re3, two O between the pairs."""
    st = Structure(CUBIC10, ["Re", "Re", "Re", "O", "O"],
                   [[0.3, 0.3, 0.5], [0.5, 0.5, 0.5], [0.7, 0.3, 0.5], [0.4, 0.4, 0.5], [0.6, 0.4, 0.5]])
    s1, s2 = sab.DynamicVoronoiSite([0, 1]), sab.DynamicVoronoiSite([1, 2])
    t = sab.Trajectory(sites=[s1, s2], atoms=[sab.Atom(index=3), sab.Atom(index=4)])
    t.analyse_structure(st)
    assert t.atom_by_index(3).in_site == s1.index and 3 in s1.contains_atoms
    assert t.atom_by_index(4).in_site == s2.index and 4 in s2.contains_atoms
    np.testing.assert_array_almost_equal(s1.centre, [0.4, 0.4, 0.5])
    np.testing.assert_array_almost_equal(s2.centre, [0.6, 0.4, 0.5])


@pytest.mark.gpu
def test_spherical_overlap_assignment_with_history():
    """test_spherical_site_collection.py:179-210: hand-set history, assignment from the atoms' own
    coordinates through assign_site_occupations, transition recorded on the old site."""
    sab.Site.reset_index()
    s1, s2, s3 = (sab.SphericalSite(np.array(c), rcut=1.5) for c in ([0, 0, 0], [0.1, 0, 0], [0.5, 0.5, 0.5]))
    coll = sab.SphericalSiteCollection([s1, s2, s3])
    a1, a2 = sab.Atom(index=0), sab.Atom(index=1)
    for a, x in ((a1, [0.05, 0.05, 0.05]), (a2, [0.15, 0.05, 0.05])):
        a.trajectory = [s1.index]
        a._recent_sites = [s1.index, None]
        a._frac_coords = np.array(x)
    coll.assign_site_occupations([a1, a2], CUBIC10)
    assert a1.in_site == s1.index and a2.in_site == s2.index
    assert s1.contains_atoms == [a1.index] and s2.contains_atoms == [a2.index]
    assert s1.transitions[s2.index] == 1
    assert a2._recent_sites == [s2.index, s1.index]


@pytest.mark.gpu
def test_append_timestep_and_from_structures_bookkeeping():
    """test_trajectory.py:159-234: t=None records no timestep, t=0 is recorded, from_structures APPENDS
    and labels 1..F; reset clears everything."""
    st = Structure(np.eye(3) * 8.0, ["O", "Li"], [[0, 0, 0], [0.26, 0.25, 0.25]])
    t = (sab.TrajectoryBuilder().with_structure(st).with_mobile_species("Li")
         .with_spherical_sites(centres=[[0.25, 0.25, 0.25], [0.75, 0.25, 0.25]], radii=1.5).build())
    t.append_timestep(st)
    assert t.timesteps == [] and t.atoms[0].trajectory == [0] and t.sites[0].trajectory == [[1]]
    t.append_timestep(st, t=0)
    assert t.timesteps == [0] and len(t) == 1
    t.trajectory_from_structures([st, st, st])
    assert t.timesteps == [0, 1, 2, 3] and t.atoms[0].trajectory == [0] * 5
    assert t.atoms_trajectory == [[0]] * 5 and t.sites_trajectory == [[[1], []]] * 5
    t.reset()
    assert t.timesteps == [] and t.atoms[0].trajectory == [] and t.sites[0].trajectory == [] and len(t) == 0
