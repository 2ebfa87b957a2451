"""Generate the committed golden fixtures under ``tests/golden/`` by running the
UNMODIFIED reference (imported from ``/root/reference`` through the pymatgen/monty
stand-ins in ``oracle/refharness``).

TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py fuzz       # only cases whose name contains "fuzz"

Each fixture is an ``.npz`` holding the inputs (fractional coordinates, lattice,
static site tables, initial PBC-cache state), and the reference's outputs
(per-atom site trajectories, per-site transition counters, recent-site history,
Qhull face topology as cached by the reference, dynamic-Voronoi centres).
XDATCAR-derived coordinates are stored as ``int32 = round(frac * 1e8)`` (the files
carry 8 decimals, so ``i / 1e8`` reproduces the parsed float64 exactly).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refcases  # noqa: E402

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"
REF = refcases.REFERENCE_ROOT


def save(name, case):
    GOLDEN.mkdir(parents=True, exist_ok=True)
    path = GOLDEN / f"{name}.npz"
    np.savez_compressed(path, **case)
    print(f"  wrote {path.name}: {path.stat().st_size / 1024:.0f} KiB, "
          f"traj {case['atoms_traj'].shape if 'atoms_traj' in case else '-'}")


def frames_fixture(name, xdatcar_path):
    from pymatgen.io.vasp import read_xdatcar_arrays
    species, lattice, frac = read_xdatcar_arrays(xdatcar_path)
    fi = np.round(frac * 1e8).astype(np.int32)
    assert np.array_equal(fi / 1e8, frac), "1e8 integer encoding is not exact"
    path = GOLDEN / f"{name}.npz"
    np.savez_compressed(path, frac_i32=fi, lattice=lattice, species=np.array(species))
    print(f"  wrote {path.name}: {path.stat().st_size / 1024:.0f} KiB, frames {frac.shape}")
    return species, lattice, frac


# ---------------------------------------------------------------------------
# cases from the reference's own documentation / tests
# ---------------------------------------------------------------------------

def case_readme():
    """README.md:33-41 quick start: expected [[0],[0],[None],[None],[1]]."""
    from site_analysis import TrajectoryBuilder
    species, lattice, frac = frames_fixture("frames_readme", REF / "example_data" / "XDATCAR")
    st = refcases.structures_from_arrays(species, lattice, frac)
    t = (TrajectoryBuilder().with_structure(st[0]).with_mobile_species("Li")
         .with_spherical_sites(centres=[[0.25, 0.25, 0.25], [0.75, 0.25, 0.25]], radii=1.5).build())
    case = refcases.run_case(t, species, lattice, frac)
    assert case["atoms_traj"].T.tolist() == [[0, 0, -1, -1, 1]]
    case["frames_file"] = np.array("frames_readme")
    save("readme_spherical", case)


def case_simple_cubic():
    """docs quickstart + tests/test_numba_validation.py:27-74 (64 cubic sites, 1 Li)."""
    from site_analysis import TrajectoryBuilder
    species, lattice, frac = frames_fixture(
        "frames_simple_cubic", REF / "docs" / "source" / "examples" / "simple_cubic_li.XDATCAR")
    st = refcases.structures_from_arrays(species, lattice, frac)
    centres = [list((np.array([i, j, k]) + 0.5) / 4) for i in range(4) for j in range(4) for k in range(4)]

    t = (TrajectoryBuilder().with_structure(st[0]).with_mobile_species("Li")
         .with_spherical_sites(centres=centres, radii=2.0).build())
    case = refcases.run_case(t, species, lattice, frac)
    assert case["atoms_traj"][:, 0].tolist() == [21] * 6 + [37] * 5
    case["frames_file"] = np.array("frames_simple_cubic")
    save("simple_cubic_spherical", case)

    t = (TrajectoryBuilder().with_structure(st[0]).with_mobile_species("Li")
         .with_voronoi_sites(centres=centres).build())
    case = refcases.run_case(t, species, lattice, frac)
    assert case["atoms_traj"][:, 0].tolist() == [21] * 6 + [37] * 5
    case["frames_file"] = np.array("frames_simple_cubic")
    save("simple_cubic_voronoi", case)

    reference_structure = st[0].copy()
    reference_structure.remove_species(["Li"])
    for c in centres:
        reference_structure.append(species="Li", coords=c)
    t = (TrajectoryBuilder().with_structure(st[0]).with_reference_structure(reference_structure)
         .with_mobile_species("Li")
         .with_polyhedral_sites(centre_species="Li", vertex_species="O", cutoff=4.0,
                                n_vertices=8, label="Cubic")
         .with_structure_alignment(align_species="O").build())
    case = refcases.run_case(t, species, lattice, frac)
    case["frames_file"] = np.array("frames_simple_cubic")
    print("   simple cubic polyhedral:", case["atoms_traj"][:, 0].tolist())
    save("simple_cubic_polyhedral", case)


def _argyrodite_builder(st0, reference_structure):
    from site_analysis import TrajectoryBuilder
    return (TrajectoryBuilder().with_structure(st0).with_reference_structure(reference_structure)
            .with_mobile_species("Li").with_structure_alignment(align_species="P")
            .with_site_mapping(mapping_species=["S", "Cl"]))


_LABELS = [("Li", "1"), ("Mg", "2"), ("Na", "3"), ("Be", "4"), ("K", "5")]


def _occupation_percent(case):
    labels = case["labels"]
    idx2label = {int(i): str(l) for i, l in zip(case["site_index"], labels)}
    flat = case["atoms_traj"].ravel()
    tot = len(flat)
    out = {}
    for lab in "12345":
        out[lab] = 100.0 * sum(1 for v in flat if v >= 0 and idx2label[int(v)] == lab) / tot
    return out


def case_argyrodite():
    """tutorials/argyrodite_site_analysis.ipynb + tests/benchmark_all_site_types.py:49-89."""
    ref = refcases.argyrodite_reference_structure()
    golden_pct = {  # notebook cells 9/11/13/14: types 5,4,3,2,1
        "0p": {"5": 80.20, "4": 0.02, "3": 0.00, "2": 19.78, "1": 0.01},
        "50p": {"5": 65.92, "4": 2.63, "3": 0.00, "2": 31.43, "1": 0.02},
        "100p": {"5": 53.39, "4": 7.33, "3": 0.00, "2": 39.28, "1": 0.00},
    }
    for tag in ("0p", "50p", "100p"):
        species, lattice, frac = frames_fixture(
            f"frames_argyrodite_{tag}", REF / "tutorials" / "data" / f"Li6PS5Cl_{tag}_XDATCAR.gz")
        st = refcases.structures_from_arrays(species, lattice, frac)

        b = _argyrodite_builder(st[0], ref)
        for sp, label in _LABELS:
            b.with_polyhedral_sites(centre_species=sp, vertex_species="S", cutoff=3.0,
                                    n_vertices=4, label=label)
        case = refcases.run_case(b.build(), species, lattice, frac)
        pct = _occupation_percent(case)
        print("   polyhedral occupation %:", {k: round(v, 2) for k, v in pct.items()})
        for k, v in golden_pct[tag].items():
            assert abs(pct[k] - v) < 0.006, (tag, k, pct[k], v)
        case["frames_file"] = np.array(f"frames_argyrodite_{tag}")
        case.pop("rank_table", None)   # recomputed locally; never consulted (count == 1 everywhere)
        save(f"argyrodite_{tag}_polyhedral", case)

        if tag == "100p":
            continue
        b = _argyrodite_builder(st[0], ref)
        for sp, label in _LABELS:
            b.with_dynamic_voronoi_sites(centre_species=sp, reference_species="S", cutoff=3.0,
                                         n_reference=4, label=label)
        centre_frames = [0, 1, 70, 139]
        centres = {}

        def hook(f, traj):
            if f in centre_frames:
                centres[f] = np.array([s.centre for s in traj.sites])
        case = refcases.run_case(b.build(), species, lattice, frac, per_frame_hook=hook)
        case["centre_frames"] = np.array(centre_frames)
        case["centres_at"] = np.array([centres[f] for f in centre_frames])
        case["frames_file"] = np.array(f"frames_argyrodite_{tag}")
        save(f"argyrodite_{tag}_dynvor", case)

        if tag != "0p":
            continue
        li = [i for i, s in enumerate(ref) if s.species_string == "Li"]
        b = _argyrodite_builder(st[0], ref)
        b.with_voronoi_sites(centres=[ref[i].frac_coords.tolist() for i in li])
        case = refcases.run_case(b.build(), species, lattice, frac)
        case["frames_file"] = np.array(f"frames_argyrodite_{tag}")
        save(f"argyrodite_{tag}_voronoi", case)

        cs = []
        for sp, _ in _LABELS:
            cs.extend(ref[i].frac_coords.tolist() for i, s in enumerate(ref) if s.species_string == sp)
        b = _argyrodite_builder(st[0], ref)
        b.with_spherical_sites(centres=cs, radii=1.5)
        case = refcases.run_case(b.build(), species, lattice, frac)
        case["frames_file"] = np.array(f"frames_argyrodite_{tag}")
        case["rank_table"] = case["rank_table"].astype(np.int16)
        save(f"argyrodite_{tag}_spherical", case)


# ---------------------------------------------------------------------------
# seeded fuzz cases: random triclinic cells, boundary coordinates, wraps, jumps
# ---------------------------------------------------------------------------

def _triclinic(rng, lo=6.0, hi=10.0, skew=25.0):
    from pymatgen.core import Lattice
    a, b, c = rng.uniform(lo, hi, 3)
    al, be, ga = 90.0 + rng.uniform(-skew, skew, 3)
    m = Lattice.from_parameters(a, b, c, al, be, ga).matrix
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))      # random rigid rotation: no zero entries
    return m @ q


def _walk(rng, start, n_frames, sigma, wrap):
    steps = rng.normal(scale=sigma, size=(n_frames,) + start.shape)
    steps[0] = 0.0
    pos = start[None] + np.cumsum(steps, axis=0)
    return pos % 1.0 if wrap else pos


def _edge_coordinates(rng, frac, atom_rows, n):
    """Overwrite a few coordinates with exact boundary values."""
    specials = [0.0, 1.0, -0.0, np.nextafter(1.0, 0.0), -1e-17, -1e-9, 0.5, 1.0 + 1e-12, -0.25, 2.75]
    F = frac.shape[0]
    for _ in range(n):
        f = rng.integers(0, F); a = rng.choice(atom_rows); d = rng.integers(0, 3)
        frac[f, a, d] = specials[rng.integers(0, len(specials))]


def _per_frame_lattice(rng, base, n_frames, eps=2e-3):
    return np.array([base * (1.0 + rng.uniform(-eps, eps)) + rng.normal(scale=1e-4, size=(3, 3))
                     for _ in range(n_frames)])


def fuzz_spherical_voronoi(seed, big=False):
    """``big``: 40 mobile atoms, 60 sites, 400 frames (round 2: the small cases have 6 atoms and 80 frames)."""
    from site_analysis import TrajectoryBuilder
    rng = np.random.default_rng(seed + (7000 if big else 0))
    n_fw, n_mob, F, S = (8, 40, 400, 60) if big else (5, 6, 80, 14)
    tag = f"big{seed}" if big else f"{seed}"
    species = ["O"] * n_fw + ["Li"] * n_mob
    base = _triclinic(rng)
    lattice = _per_frame_lattice(rng, base, F) if seed % 2 else base
    frac = np.empty((F, n_fw + n_mob, 3))
    frac[:, :n_fw] = _walk(rng, rng.uniform(0, 1, (n_fw, 3)), F, 0.01, True)
    frac[:, n_fw:n_fw + 3] = _walk(rng, rng.uniform(0, 1, (3, 3)), F, 0.06, True)
    frac[:, n_fw + 3:] = _walk(rng, rng.uniform(0, 1, (n_mob - 3, 3)), F, 0.06, False)
    _edge_coordinates(rng, frac, list(range(n_fw, n_fw + n_mob)), 25)
    if seed % 3 == 0:   # symmetric centres + atoms on exact midpoints: argmin / overlap ties
        g = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)]) / 2.0 + 0.25
        centres = np.vstack([g, rng.uniform(0, 1, (S - 8, 3))])
        for f in range(0, F, 7):
            frac[f, n_fw] = [0.5, 0.25, 0.25]
            frac[f, n_fw + 1] = [0.5, 0.5, 0.5]
            frac[f, n_fw + 2] = [0.0, 0.75, 0.25]
    else:
        centres = rng.uniform(0, 1, (S, 3))
    st = refcases.structures_from_arrays(species, lattice, frac)

    radii = list(rng.uniform(1.2, 2.4 if big else 3.2, S))
    t = (TrajectoryBuilder().with_structure(st[0]).with_mobile_species("Li")
         .with_spherical_sites(centres=centres.tolist(), radii=radii).build())
    case = refcases.run_case(t, species, lattice, frac)
    case["frac"] = frac
    save(f"fuzz_spherical_{tag}", case)

    t = (TrajectoryBuilder().with_structure(st[0]).with_mobile_species("Li")
         .with_voronoi_sites(centres=centres.tolist()).build())
    case = refcases.run_case(t, species, lattice, frac)
    case["frac"] = frac
    save(f"fuzz_voronoi_{tag}", case)


def _grid_framework(rng, n, jitter):
    g = np.array([[i, j, k] for i in range(n) for j in range(n) for k in range(n)], dtype=float) / n
    return g + rng.normal(scale=jitter, size=g.shape)


def _cube_sites(n):
    """Vertex index lists of the n^3 cubes of an n x n x n periodic grid + their centres."""
    def idx(i, j, k):
        return ((i % n) * n + (j % n)) * n + (k % n)
    verts, centres = [], []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                verts.append([idx(i + a, j + b, k + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)])
                centres.append((np.array([i, j, k]) + 0.5) / n)
    return verts, np.array(centres)


def _tet_sites(n):
    """Five tetrahedra per cube (a proper decomposition: 4 corner + 1 central)."""
    def idx(i, j, k):
        return ((i % n) * n + (j % n)) * n + (k % n)
    corner_sets = [
        [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)],
        [(1, 1, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1)],
        [(1, 0, 1), (1, 0, 0), (0, 0, 1), (1, 1, 1)],
        [(0, 1, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1)],
        [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)],
    ]
    verts, centres = [], []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                for cs in corner_sets:
                    verts.append([idx(i + a, j + b, k + c) for a, b, c in cs])
                    centres.append((np.array([i, j, k]) + np.mean(cs, axis=0)) / n)
    return verts, np.array(centres)


def fuzz_polyhedral(seed, big=False):
    """Hand-made PolyhedralSite objects (the reference's low-level API) on a jittered grid.

    Variants by seed: cubes / tetrahedra; with / without reference centres; primed from the
    first frame (as SiteFactory does, site_factory.py:248-257) or unprimed; a >0.3 jump of
    one framework atom in one case (forces the cache invalidation of polyhedral_site.py:398-416).
    """
    from site_analysis.atom import Atom
    from site_analysis.polyhedral_site import PolyhedralSite
    from site_analysis.site import Site
    from site_analysis.trajectory import Trajectory
    rng = np.random.default_rng(seed + (7000 if big else 0))
    n = 4 if big else 3                    # big: 64 cubes / 320 tetrahedra, 24 mobile atoms, 300 frames
    F = 300 if big else 70
    tets = bool(seed % 2)
    with_rc = (seed // 2) % 2 == 0
    primed = (seed // 4) % 2 == 0
    jump = seed % 5 == 4
    verts, centres = _tet_sites(n) if tets else _cube_sites(n)
    n_fw = n ** 3
    n_mob = 24 if big else 7
    species = ["O"] * n_fw + ["Li"] * n_mob
    base = _triclinic(rng, 8.0, 11.0, 12.0)
    lattice = _per_frame_lattice(rng, base, F) if seed % 3 == 0 else base
    fw0 = _grid_framework(rng, n, 0.012)
    fw = fw0[None] + rng.normal(scale=0.008, size=(F, n_fw, 3))     # vibrating framework
    fw = fw % 1.0                                                   # wraps at the cell faces
    if jump:
        fw[40:, 5] = (fw[40:, 5] + np.array([0.41, 0.0, 0.0])) % 1.0
    frac = np.empty((F, n_fw + n_mob, 3))
    frac[:, :n_fw] = fw
    frac[:, n_fw:n_fw + 4] = _walk(rng, rng.uniform(0, 1, (4, 3)), F, 0.05, True)
    frac[:, n_fw + 4:] = _walk(rng, rng.uniform(0, 1, (n_mob - 4, 3)), F, 0.05, False)
    _edge_coordinates(rng, frac, list(range(n_fw, n_fw + n_mob)), 20)
    # put a mobile atom exactly on a framework vertex / on the plane of a shared face
    frac[10, n_fw] = frac[10, 0]
    frac[11, n_fw] = (frac[11, 0] + frac[11, 1]) / 2
    st = refcases.structures_from_arrays(species, lattice, frac)

    Site.reset_index()
    sites = [PolyhedralSite(vertex_indices=v, label=("tet" if tets else "cube"),
                            reference_center=(c if with_rc else None))
             for v, c in zip(verts, centres)]
    if primed:
        for s in sites:
            s.assign_vertex_coords(st[0].frac_coords, st[0].lattice.matrix)
    atoms = [Atom(index=i) for i in range(n_fw, n_fw + n_mob)]
    t = Trajectory(sites=sites, atoms=atoms)
    case = refcases.run_case(t, species, lattice, frac)
    case["frac"] = frac
    case["variant"] = np.array(f"tets={tets} rc={with_rc} primed={primed} jump={jump}")
    n_none = int((case["atoms_traj"] < 0).sum())
    print(f"   {case['variant']}: None assignments {n_none}/{case['atoms_traj'].size}")
    save(f"fuzz_polyhedral_big{seed}" if big else f"fuzz_polyhedral_{seed}", case)


def fuzz_dynvor(seed, big=False):
    """DynamicVoronoiSite objects with two reference-count groups (R=4 tetrahedra and R=8 cubes)."""
    from site_analysis.atom import Atom
    from site_analysis.dynamic_voronoi_site import DynamicVoronoiSite
    from site_analysis.site import Site
    from site_analysis.trajectory import Trajectory
    rng = np.random.default_rng(1000 + seed + (7000 if big else 0))
    n, F = (4, 300) if big else (3, 70)
    with_rc = seed % 2 == 0
    jump = seed % 3 == 2
    cube_v, cube_c = _cube_sites(n)
    tet_v, tet_c = _tet_sites(n)
    pick_t = rng.choice(len(tet_v), 60 if big else 20, replace=False)
    pick_c = rng.choice(len(cube_v), 24 if big else 9, replace=False)
    half = len(pick_t) // 2
    defs = [(tet_v[i], tet_c[i]) for i in pick_t[:half]] + [(cube_v[i], cube_c[i]) for i in pick_c] + \
           [(tet_v[i], tet_c[i]) for i in pick_t[half:]]                   # interleaved groups
    n_fw, n_mob = n ** 3, (20 if big else 6)
    species = ["O"] * n_fw + ["Li"] * n_mob
    base = _triclinic(rng, 8.0, 11.0, 12.0)
    lattice = _per_frame_lattice(rng, base, F) if seed % 2 else base
    fw0 = _grid_framework(rng, n, 0.012)
    fw = (fw0[None] + rng.normal(scale=0.008, size=(F, n_fw, 3))) % 1.0
    if jump:
        fw[33:, 7] = (fw[33:, 7] + np.array([0.0, 0.38, 0.0])) % 1.0
    frac = np.empty((F, n_fw + n_mob, 3))
    frac[:, :n_fw] = fw
    frac[:, n_fw:] = _walk(rng, rng.uniform(0, 1, (n_mob, 3)), F, 0.05, seed % 2 == 0)
    _edge_coordinates(rng, frac, list(range(n_fw, n_fw + n_mob)), 15)
    Site.reset_index()
    sites = [DynamicVoronoiSite(reference_indices=v, label=f"R{len(v)}",
                                reference_center=(c if with_rc else None)) for v, c in defs]
    atoms = [Atom(index=i) for i in range(n_fw, n_fw + n_mob)]
    t = Trajectory(sites=sites, atoms=atoms)
    all_centres = []

    def hook(f, traj):
        all_centres.append(np.array([s.centre for s in traj.sites]))
    case = refcases.run_case(t, species, lattice, frac, per_frame_hook=hook)
    case["frac"] = frac
    case["centre_frames"] = np.arange(F)
    case["centres_at"] = np.array(all_centres)
    case["variant"] = np.array(f"rc={with_rc} jump={jump}")
    save(f"fuzz_dynvor_big{seed}" if big else f"fuzz_dynvor_{seed}", case)


CASES = {
    "readme": case_readme,
    "simple_cubic": case_simple_cubic,
    "argyrodite": case_argyrodite,
    **{f"fuzz_sphvor_{s}": (lambda s=s: fuzz_spherical_voronoi(s)) for s in range(4)},
    **{f"fuzz_polyhedral_{s}": (lambda s=s: fuzz_polyhedral(s)) for s in range(10)},
    **{f"fuzz_dynvor_{s}": (lambda s=s: fuzz_dynvor(s)) for s in range(4)},
    # round 2: larger systems ("fuzz fixtures are tiny"): 20-40 mobile atoms, 300-400 frames
    "fuzzbig_sphvor_1": lambda: fuzz_spherical_voronoi(1, big=True),
    "fuzzbig_sphvor_3": lambda: fuzz_spherical_voronoi(3, big=True),
    "fuzzbig_polyhedral_1": lambda: fuzz_polyhedral(1, big=True),     # tetrahedra, reference centres, primed
    "fuzzbig_polyhedral_4": lambda: fuzz_polyhedral(4, big=True),     # cubes, reference centres, unprimed, a > 0.3 jump
    "fuzzbig_polyhedral_7": lambda: fuzz_polyhedral(7, big=True),     # tetrahedra, no reference centres, unprimed
    "fuzzbig_dynvor_1": lambda: fuzz_dynvor(1, big=True),
    "fuzzbig_dynvor_2": lambda: fuzz_dynvor(2, big=True),             # a 0.38 jump: group re-initialisation
}


def main():
    refcases.import_reference()
    pats = sys.argv[1:]
    for name, fn in CASES.items():
        if pats and not any(p in name for p in pats):
            continue
        print(name)
        fn()


if __name__ == "__main__":
    main()
