"""Golden fixtures for the post-processing folds, produced by the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/gen_golden_folds.py

A finished assignment history (rows[F][A] of site indices, -1 = None) is replayed through the reference's own
objects -- ``SiteCollection.update_occupation`` (site_collection.py:279-310), the appends of
``Trajectory.append_timestep`` (trajectory.py:357-360) -- and the reference's ``Site.residence_times`` /
``Site.average_occupation`` (site.py:225-325) and ``Site.transitions`` are recorded.  ``tests/golden/folds_*.npz``
hold rows, the atom / site indices and those outputs.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import folds_oracle  # noqa: E402
import refcases  # noqa: E402

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"
PARAMS = [(0, False), (0, True), (1, False), (2, True), (5, False), (40, True)]     # (filter_length, include_edge_runs)


def replay(rows, atom_index, site_index, prev=None):
    refcases.import_reference()
    from site_analysis.atom import Atom
    from site_analysis.spherical_site import SphericalSite
    from site_analysis.spherical_site_collection import SphericalSiteCollection
    from site_analysis.site import Site
    Site._newid = 0
    sites = []
    for s in site_index:
        site = SphericalSite(frac_coords=np.zeros(3), rcut=1.0)
        site.index = int(s)
        sites.append(site)
    coll = SphericalSiteCollection(sites)
    atoms = [Atom(index=int(a)) for a in atom_index]
    for a, atom in enumerate(atoms):
        atom._frac_coords = np.zeros(3)
        if prev is not None and prev[a] != -2:
            atom.trajectory.append(None if prev[a] < 0 else int(prev[a]))
    n0 = {a.index: len(a.trajectory) for a in atoms}
    by_index = {s.index: s for s in sites}
    for row in rows:
        coll.reset_site_occupations()
        for atom, v in zip(atoms, row):
            atom.in_site = None
            if v >= 0:
                coll.update_occupation(by_index[int(v)], atom)
        for atom in atoms:                                   # trajectory.py:357-360
            atom.trajectory.append(atom.in_site)
        for site in sites:
            site.trajectory.append(site.contains_atoms)
    assert all(len(a.trajectory) - n0[a.index] == len(rows) for a in atoms)
    return sites


def make(name, rows, atom_index, site_index, prev=None):
    sites = replay(rows, atom_index, site_index, prev)
    case = dict(rows=rows.astype(np.int32), atom_index=np.asarray(atom_index, dtype=np.int32),
                site_index=np.asarray(site_index, dtype=np.int32),
                prev=np.full(len(atom_index), -2, dtype=np.int32) if prev is None else np.asarray(prev, dtype=np.int32),
                params=np.array([[p[0], int(p[1])] for p in PARAMS], dtype=np.int32))
    occ = np.array([np.nan if s.average_occupation is None else s.average_occupation for s in sites])
    case["average_occupation"] = occ
    for k, (fl, edge) in enumerate(PARAMS):
        flat, off = [], [0]
        for s in sites:
            rt = s.residence_times(filter_length=fl, include_edge_runs=edge)
            flat.extend(rt)
            off.append(len(flat))
        case[f"rt{k}_flat"] = np.array(flat, dtype=np.int64)
        case[f"rt{k}_off"] = np.array(off, dtype=np.int64)
    tr = [(s.index, int(k), int(v)) for s in sites for k, v in s.transitions.items()]       # Counter insertion order
    case["transitions"] = np.array(tr, dtype=np.int64).reshape(-1, 3)
    # Site.trajectory as CSR over (site, frame): atoms in the reference's order
    flat, off = [], [0]
    for s in sites:
        for ts in s.trajectory:
            flat.extend(ts)
            off.append(len(flat))
    case["site_traj_flat"] = np.array(flat, dtype=np.int32)
    case["site_traj_off"] = np.array(off, dtype=np.int64)
    # cross-check the plain-Python restatement while the reference is at hand
    st = folds_oracle.site_trajectories(rows, atom_index, site_index)
    for s in sites:
        assert st[s.index] == s.trajectory
        for fl, edge in PARAMS:
            assert folds_oracle.residence_times(s.trajectory, fl, edge) == s.residence_times(fl, edge)
        assert folds_oracle.average_occupation(s.trajectory) == s.average_occupation
    want = folds_oracle.transitions(rows, None if prev is None else prev)
    got = {s.index: dict(s.transitions) for s in sites if s.transitions}
    assert got == want and all(list(got[k]) == list(want[k]) for k in got)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **case)
    print(f"  wrote {name}.npz rows {rows.shape} transitions {len(tr)}")


def main():
    rng = np.random.default_rng(20260107)
    # site indices deliberately not 0..S-1 and atoms not sorted by index
    for k, (F, A, S, stay, none) in enumerate([(60, 5, 4, 0.7, 0.1), (300, 12, 9, 0.9, 0.05), (500, 7, 3, 0.97, 0.01),
                                                (40, 3, 2, 0.5, 0.3)]):
        pos = folds_oracle.random_rows(rng, F, A, S, stay, none)
        site_index = np.arange(S) * 3 + 10
        atom_index = rng.permutation(np.arange(A) + 100)
        rows = np.where(pos >= 0, site_index[np.clip(pos, 0, None)], -1)
        prev = None
        if k % 2 == 1:
            p = rng.integers(-2, S, size=A)
            prev = np.where(p >= 0, site_index[np.clip(p, 0, None)], p)
        make(f"folds_random_{k}", rows, atom_index, site_index, prev)
    # docstring examples of site.py:268-290
    make("folds_docstring", np.array([[-1], [7], [-1], [7], [-1], [7], [7], [7]]), [1], [7])
    # several atoms in one site in one frame, an always-occupied and a never-occupied site
    rows = np.array([[0, 0, 2], [0, 0, 2], [1, 0, 2], [0, 1, 2], [0, 0, 2], [0, 0, 2]])
    make("folds_shared_site", rows, [4, 2, 9], [0, 1, 2, 3])


if __name__ == "__main__":
    main()
