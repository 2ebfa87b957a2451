"""Golden fixtures for the reference-structure workflow, produced by the UNMODIFIED reference
(``TrajectoryBuilder.with_polyhedral_sites`` / ``with_dynamic_voronoi_sites``, builders.py:385-636, on top of
``reference_workflow/``).  TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python oracle/gen_golden_workflow.py

tests/golden/workflow_*.npz hold the two input structures as arrays, the builder arguments and, per generated site, the
vertex / reference atom indices, the reference centre, the label and (polyhedra) the primed PBC image shifts, plus the
alignment translation vector and metrics of ``ReferenceBasedSites``.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refcases  # noqa: E402

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"
REF = refcases.REFERENCE_ROOT
LABELS = [("Li", "1"), ("Mg", "2"), ("Na", "3"), ("Be", "4"), ("K", "5")]


def arrays(structure):
    return (np.array([s.species_string for s in structure]), np.array(structure.lattice.matrix, dtype=np.float64),
            np.array(structure.frac_coords, dtype=np.float64))


def snapshot(sites):
    slots = [list(getattr(s, "vertex_indices", None) or s.reference_indices) for s in sites]
    width = max(len(v) for v in slots)
    idx = np.full((len(sites), width), -1, dtype=np.int32)
    for i, v in enumerate(slots):
        idx[i, :len(v)] = v
    out = dict(slot_idx=idx, n_slot=np.array([len(v) for v in slots], dtype=np.int32),
               labels=np.array(["" if s.label is None else str(s.label) for s in sites]),
               site_index=np.array([s.index for s in sites], dtype=np.int32),
               has_centre=np.array([s.reference_center is not None for s in sites]),
               centre=np.array([np.full(3, np.nan) if s.reference_center is None else s.reference_center for s in sites]))
    if hasattr(sites[0], "vertex_indices"):
        sh = np.zeros((len(sites), width, 3), dtype=np.int64)
        raw = np.zeros((len(sites), width, 3))
        vc = np.zeros((len(sites), width, 3))
        for i, s in enumerate(sites):
            n = len(s.vertex_indices)
            sh[i, :n] = s._pbc_image_shifts; raw[i, :n] = s._pbc_cached_raw_frac; vc[i, :n] = s.vertex_coords
        out.update(shift0=sh, cached_raw0=raw, vertex_coords=vc)
    return out


def save(name, ref, target, kind, calls, builder_kwargs, trajectory, rbs):
    case = dict(kind=np.array(kind))
    for tag, st in (("ref", ref), ("tgt", target)):
        sp, lat, fr = arrays(st)
        case[f"{tag}_species"], case[f"{tag}_lattice"], case[f"{tag}_frac"] = sp, lat, fr
    case["calls"] = np.array([repr(c) for c in calls])
    case["builder_kwargs"] = np.array(repr(builder_kwargs))
    case.update(snapshot(trajectory.sites))
    case["atom_idx"] = np.array([a.index for a in trajectory.atoms], dtype=np.int32)
    if rbs is not None and rbs.translation_vector is not None:
        case["translation"] = np.array(rbs.translation_vector, dtype=np.float64)
        case["metrics"] = np.array([rbs.alignment_metrics[k] for k in ("rmsd", "max_dist", "mean_dist")])
        case["aligned_frac"] = np.array(rbs.aligned_structure.frac_coords, dtype=np.float64)
    np.savez_compressed(GOLDEN / f"{name}.npz", **case)
    print(f"  wrote {name}.npz: {len(trajectory.sites)} sites")


def main():
    refcases.import_reference()
    from pymatgen.core import Lattice, Structure
    from pymatgen.io.vasp import read_xdatcar_arrays
    from site_analysis import TrajectoryBuilder
    from site_analysis.reference_workflow import ReferenceBasedSites

    # --- argyrodite (tutorials/argyrodite_site_analysis.ipynb; tests/benchmark_all_site_types.py:49-89)
    ref = refcases.argyrodite_reference_structure()
    species, lattice, frac = read_xdatcar_arrays(REF / "tutorials" / "data" / "Li6PS5Cl_0p_XDATCAR.gz")
    target = Structure(Lattice(lattice), species, frac[0])
    kw = dict(align_species="P", mapping_species=["S", "Cl"])
    for kind in ("polyhedral", "dynamic_voronoi"):
        b = (TrajectoryBuilder().with_structure(target).with_reference_structure(ref).with_mobile_species("Li")
             .with_structure_alignment(align_species="P").with_site_mapping(mapping_species=["S", "Cl"]))
        calls = []
        for sp, label in LABELS:
            if kind == "polyhedral":
                b.with_polyhedral_sites(centre_species=sp, vertex_species="S", cutoff=3.0, n_vertices=4, label=label)
            else:
                b.with_dynamic_voronoi_sites(centre_species=sp, reference_species="S", cutoff=3.0, n_reference=4, label=label)
            calls.append((sp, "S", 3.0, 4, label))
        rbs = ReferenceBasedSites(ref, target, align=True, align_species=["P"])
        save(f"workflow_argyrodite_{kind}", ref, target, kind, calls, kw, b.build(), rbs)

    # --- simple cubic Li in an O cage (docs quickstart / tests/test_numba_validation.py:27-74): 8-vertex polyhedra
    species, lattice, frac = read_xdatcar_arrays(REF / "docs" / "source" / "examples" / "simple_cubic_li.XDATCAR")
    target = Structure(Lattice(lattice), species, frac[0])
    ref_sc = Structure.from_spacegroup(sg="Pm-3m", lattice=Lattice.cubic(a=float(lattice[0, 0]) / 4), species=["Li", "O"],
                                       coords=[[0.5, 0.5, 0.5], [0.0, 0.0, 0.0]]) * [4, 4, 4]
    b = (TrajectoryBuilder().with_structure(target).with_reference_structure(ref_sc).with_mobile_species("Li")
         .with_polyhedral_sites(centre_species="Li", vertex_species="O", cutoff=4.0, n_vertices=8, label="Cubic")
         .with_structure_alignment(align_species="O"))
    rbs = ReferenceBasedSites(ref_sc, target, align=True, align_species=["O"])
    save("workflow_simple_cubic_polyhedral", ref_sc, target, "polyhedral", [("Li", "O", 4.0, 8, "Cubic")],
         dict(align_species="O", mapping_species=None), b.build(), rbs)

    # --- no alignment, spread-based PBC handling (use_reference_centers=False), rock-salt cell with jitter
    rng = np.random.default_rng(20260111)
    ref_rs = Structure.from_spacegroup(sg="Fm-3m", lattice=Lattice.cubic(5.6), species=["Na", "Cl"],
                                       coords=[[0, 0, 0], [0.5, 0.5, 0.5]]) * [2, 2, 2]
    tgt_frac = (np.array(ref_rs.frac_coords) + rng.normal(0, 0.01, size=(len(ref_rs), 3))) % 1.0
    target = Structure(ref_rs.lattice, [s.species_string for s in ref_rs], tgt_frac)
    b = (TrajectoryBuilder().with_structure(target).with_reference_structure(ref_rs).with_mobile_species("Na")
         .with_structure_alignment(align=False)
         .with_dynamic_voronoi_sites(centre_species="Na", reference_species="Cl", cutoff=3.0, n_reference=6, label="oct",
                                     use_reference_centers=False))
    save("workflow_rocksalt_dynvor_noalign", ref_rs, target, "dynamic_voronoi", [("Na", "Cl", 3.0, 6, "oct")],
         dict(align=False, use_reference_centers=False), b.build(), None)


if __name__ == "__main__":
    main()
