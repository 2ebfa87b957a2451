"""Trajectory driver (mirror of reference ``site_analysis/trajectory.py``).

Same constructor checks, attributes and methods as the reference ``Trajectory``
(trajectory.py:54-432).  ``trajectory_from_structures`` stacks the structures into
``frac[F, N, 3]`` / ``lattice[F, 3, 3]`` arrays and runs the whole trajectory on the GPU in one
call (:meth:`trajectory_from_arrays`, the array-native entry SURVEY.md H7 asks for); the
per-frame methods ``analyse_structure`` / ``append_timestep`` go through the same kernels one
frame at a time.  Per-atom / per-site trajectories are kept as int32 arrays and turned into the
reference's nested lists lazily.
"""
from __future__ import annotations

import warnings
from collections import Counter
from typing import Sequence

import numpy as np

from .atom import Atom
from .folds import BlockFolds
from .site import Site
from .transition_table import TransitionTable
from .site_collection import POINTS_LIMIT  # noqa: E402
from .site_collection import (DynamicVoronoiSiteCollection, PolyhedralSiteCollection, SiteCollection,
                              SphericalSiteCollection, VoronoiSiteCollection)

# site class NAME -> collection: sites made by the reference package itself are accepted as well
_COLLECTIONS = {
    "PolyhedralSite": PolyhedralSiteCollection,
    "VoronoiSite": VoronoiSiteCollection,
    "SphericalSite": SphericalSiteCollection,
    "DynamicVoronoiSite": DynamicVoronoiSiteCollection,
}


class Trajectory:
    """Site analysis of a simulation trajectory (reference trajectory.py:54)."""

    def __init__(self, sites: Sequence[Site], atoms: list[Atom]) -> None:
        if not sites:
            raise ValueError("Cannot initialize Trajectory with empty sites list")
        if not atoms:
            raise ValueError("Cannot initialize Trajectory with empty atoms list")
        if len(set(type(s) for s in sites)) > 1:
            raise TypeError("A Trajectory cannot be initialised with mixed Site types")
        site_type = type(sites[0])
        try:
            collection_class = _COLLECTIONS[site_type.__name__]
        except KeyError:
            raise TypeError(f"Site type {site_type} not recognised for Trajectory initialisation")
        self.site_collection: SiteCollection = collection_class(sites)
        self.sites = sites
        self.atoms = atoms
        self.timesteps: list[int] = []
        self.atom_lookup = {a.index: i for i, a in enumerate(atoms)}
        self.site_lookup = {s.index: i for i, s in enumerate(sites)}

    def atom_by_index(self, i: int) -> Atom:
        return self.atoms[self.atom_lookup[i]]

    def site_by_index(self, i: int) -> Site:
        return self.sites[self.site_lookup[i]]

    def analyse_structure(self, structure) -> None:
        self.site_collection.analyse_structure(self.atoms, structure)

    def assign_site_occupations(self, structure) -> None:
        self.site_collection.assign_site_occupations(self.atoms, structure.lattice.matrix)

    def site_coordination_numbers(self) -> Counter:
        return Counter([s.coordination_number for s in self.sites])

    def site_labels(self) -> list[str | None]:
        return [s.label for s in self.sites]

    # -- transition tables (reference trajectory.py:168-321) ---------------------------------------
    def _transition_entries(self):
        """(source site, destination index, count) of every Counter entry; unknown destinations raise as the
        reference does (trajectory.py:179-186)."""
        known = {s.index for s in self.sites}
        for site in self.sites:
            for dest, count in site.transitions.items():
                if dest not in known:
                    raise ValueError(f"Site {site.index} has a transition to unknown site index {dest}.")
                yield site, dest, count

    def transition_counts_by_site(self, *, keys: Sequence[int] | None = None) -> TransitionTable:
        """Per-site transition counts (trajectory.py:188-219): rows = source site index, sorted unless ``keys``."""
        order = tuple(sorted(s.index for s in self.sites))
        where = {k: i for i, k in enumerate(order)}
        counts = np.zeros((len(order), len(order)), dtype=int)
        for site, dest, count in self._transition_entries():
            counts[where[site.index], where[dest]] = count
        table = TransitionTable(keys=order, matrix=counts)
        return table if keys is None else table.reorder(keys)

    def transition_counts_by_label(self, *, keys: Sequence[str] | None = None) -> TransitionTable:
        """Label-aggregated counts (trajectory.py:221-274); transitions touching unlabelled sites are dropped with a
        warning."""
        label_of = {s.index: s.label for s in self.sites if s.label is not None}
        order = tuple(sorted(set(label_of.values())))
        where = {k: i for i, k in enumerate(order)}
        counts = np.zeros((len(order), len(order)), dtype=int)
        dropped = 0
        for site, dest, count in self._transition_entries():
            a, b = label_of.get(site.index), label_of.get(dest)
            if a is None or b is None:
                dropped += count
            else:
                counts[where[a], where[b]] += count
        if dropped > 0:
            warnings.warn(f"{dropped} transition(s) involving unlabelled sites were "
                          f"excluded from the label-aggregated counts.", stacklevel=2)
        table = TransitionTable(keys=order, matrix=counts)
        return table if keys is None else table.reorder(keys)

    @staticmethod
    def _normalise_counts(counts: TransitionTable) -> TransitionTable:
        """Row-normalise (trajectory.py:168-176); rows without transitions stay zero."""
        c = counts.matrix.astype(float)
        total = c.sum(axis=1, keepdims=True)
        return TransitionTable(keys=counts.keys, matrix=np.divide(c, total, out=np.zeros_like(c), where=total > 0))

    def transition_probabilities_by_site(self, *, keys: Sequence[int] | None = None) -> TransitionTable:
        return self._normalise_counts(self.transition_counts_by_site(keys=keys))

    def transition_probabilities_by_label(self, *, keys: Sequence[str] | None = None) -> TransitionTable:
        return self._normalise_counts(self.transition_counts_by_label(keys=keys))

    @property
    def atom_sites(self) -> list[int | None]:
        return [atom.in_site for atom in self.atoms]

    @property
    def site_occupations(self) -> list[list[int]]:
        return [s.contains_atoms for s in self.sites]

    def append_timestep(self, structure, t: int | None = None) -> None:
        """trajectory.py:342-362."""
        self.analyse_structure(structure)
        for atom in self.atoms:
            atom.trajectory.append(atom.in_site)
        for site in self.sites:
            site.trajectory.append(site.contains_atoms)
        if t is not None:
            self.timesteps.append(t)

    def reset(self) -> None:
        """trajectory.py:364-373."""
        for atom in self.atoms:
            atom.reset()
        self.site_collection.reset()
        self.timesteps = []

    # -- whole trajectories ---------------------------------------------------------------------
    def trajectory_from_arrays(self, frac_coords: np.ndarray, lattice: np.ndarray,
                               timesteps: Sequence[int] | None = None) -> np.ndarray:
        """Analyse F frames given as arrays and append them to the trajectory.

        Args:
            frac_coords: (F, N, 3) fractional coordinates, ``structure.frac_coords`` per frame.
            lattice: (3, 3) or (F, 3, 3) lattice matrices (rows are lattice vectors).
            timesteps: labels appended to ``self.timesteps`` (default 1..F, as
                ``trajectory_from_structures`` does, trajectory.py:425).

        Returns:
            (F, A) int32 array of ``Site.index`` per frame and atom (-1 = no site).
        """
        frac_coords = np.asarray(frac_coords)
        rows = self.site_collection.analyse_trajectory(self.atoms, frac_coords, lattice)
        F = rows.shape[0]
        if F == 0:
            return rows
        atom_index = np.array([a.index for a in self.atoms])
        foreign = [o for o in list(self.atoms) + list(self.sites) if not hasattr(o, "_lazy_blocks")]
        if foreign:                                             # objects of the reference package itself: plain lists
            for a, atom in enumerate(self.atoms):
                atom.trajectory.extend(None if v < 0 else int(v) for v in rows[:, a].tolist())
            index_of = [a.index for a in self.atoms]
            for site in self.sites:
                hit = rows == site.index
                site.trajectory.extend([[index_of[a] for a in np.nonzero(r)[0]] for r in hit])
            self.timesteps.extend(range(1, F + 1) if timesteps is None else timesteps)
            return rows
        for a, atom in enumerate(self.atoms):
            atom._lazy_blocks.append(rows[:, a])
        # one shared, lazily GPU-folded form of the batch answers every site's trajectory / residence times
        folds = BlockFolds(rows, atom_index, self.site_collection._index_of_pos)
        keep_points = rows.size > POINTS_LIMIT          # short batches got their Site.points eagerly (site_collection._run)
        if keep_points:
            folds.frac = frac_coords
        for p, site in enumerate(self.sites):
            site._lazy_blocks.append((folds, p))
            if keep_points and hasattr(site, "_lazy_points"):
                site._lazy_points.append((folds, p))
        self.timesteps.extend(range(1, F + 1) if timesteps is None else timesteps)
        return rows

    def trajectory_from_structures(self, structures, progress: bool = False) -> None:
        """trajectory.py:413-432: appends one timestep per structure (labels 1..F)."""
        structures = list(structures)
        if not structures:
            return
        frac = np.stack([np.asarray(s.frac_coords, dtype=np.float64) for s in structures])
        lats = np.stack([np.asarray(s.lattice.matrix, dtype=np.float64) for s in structures])
        lattice = lats[0] if np.all(lats == lats[0]) else lats
        self.trajectory_from_arrays(frac, lattice)

    def trajectory_from_xdatcar(self, path, n_threads: int | None = None) -> np.ndarray:
        """Read an XDATCAR (``.gz`` accepted) with the native parser (:mod:`site_analysis_b200.io`) and append its frames;
        equivalent to ``trajectory_from_structures(Xdatcar(path).structures)`` without building ``Structure`` objects."""
        from .io import read_xdatcar
        species, lattice, frac = read_xdatcar(path, n_threads=n_threads, pinned=True)
        return self.trajectory_from_arrays(frac, lattice)

    @property
    def atoms_trajectory(self):
        return list(map(list, zip(*[atom.trajectory for atom in self.atoms])))

    @property
    def sites_trajectory(self):
        return list(map(list, zip(*[site.trajectory for site in self.sites])))

    @property
    def at(self):
        return self.atoms_trajectory

    @property
    def st(self):
        return self.sites_trajectory

    def atoms_trajectory_array(self) -> np.ndarray:
        """(F, A) int32 of ``Site.index`` (-1 = None) without building nested lists."""
        cols = []
        for atom in self.atoms:
            parts = [np.array([-1 if v is None else v for v in atom._trajectory], dtype=np.int32)] + \
                    [np.asarray(b, dtype=np.int32) for b in atom._lazy_blocks]
            cols.append(np.concatenate(parts))
        return np.stack(cols, axis=1)

    def __len__(self):
        return len(self.timesteps)

    def site_summaries(self, metrics: list[str] | None = None) -> list[dict]:
        return self.site_collection.summaries(metrics=metrics)

    def write_site_summaries(self, filename: str, metrics: list[str] | None = None) -> None:
        import json
        with open(filename, "w") as f:
            json.dump(self.site_summaries(metrics=metrics), f, indent=2)
