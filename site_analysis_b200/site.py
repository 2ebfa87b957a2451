"""Site base class (mirror of reference ``site_analysis/site.py``).

Attributes as in the reference (site.py:24-88): ``index`` (global counter), ``label``,
``contains_atoms``, ``trajectory``, ``points``, ``transitions``.  ``trajectory`` is
materialised lazily from the array-backed assignment results of GPU batches.
"""
from __future__ import annotations

from collections import Counter
from typing import Any

import numpy as np

from .folds import (BlockFolds, occupied_frames_from_runs, residence_times_from_runs,
                    site_runs_over_blocks)


class Site:
    _newid = 0

    def __init__(self, label: str | None = None) -> None:
        self.index = Site._newid
        Site._newid += 1
        self.label = label
        self.contains_atoms: list[int] = []
        self._trajectory: list[list[int]] = []
        self._lazy_blocks: list = []                 # pending (results (F, A), atom_index (A,)) pairs
        self._points: list[np.ndarray] = []
        self._lazy_points: list = []                 # pending (BlockFolds with .frac, site position) pairs of long batches
        self._transitions: Counter = Counter()
        self._lazy_transitions = None                # (collection, state token): the Counter lives in the GPU engine's store

    @property
    def transitions(self) -> Counter:
        """Counter {destination Site.index: count} in insertion order (reference site.py:71, site_collection.py:300-302).
        After a GPU batch it is built from the engine's folded store on first access."""
        if self._lazy_transitions is not None:
            collection, token = self._lazy_transitions
            self._lazy_transitions = None
            self._transitions = collection._site_transitions(self, token)
        return self._transitions

    @transitions.setter
    def transitions(self, value) -> None:
        self._lazy_transitions = None
        self._transitions = value

    @property
    def trajectory(self) -> list[list[int]]:
        """Atom indices occupying this site at each appended timestep (atom list order)."""
        if self._lazy_blocks:
            for block, second in self._lazy_blocks:
                if isinstance(block, BlockFolds):            # GPU-folded batch: only this site's own runs are touched
                    per_frame = [[] for _ in range(block.n_frames)]
                    col, start, length = block.site_runs(second)
                    for a, f0, n in zip(block.atom_index[col].tolist(), start.tolist(), length.tolist()):
                        for f in range(f0, f0 + n):          # runs come in atom LIST order, so every frame's list does too
                            per_frame[f].append(a)
                else:                                        # hand-made (rows, atom_index) block
                    frames, cols = np.nonzero(block == self.index)
                    per_frame = [[] for _ in range(block.shape[0])]
                    for f, a in zip(frames.tolist(), second[cols].tolist()):
                        per_frame[f].append(a)
                self._trajectory.extend(per_frame)
            self._lazy_blocks = []
        return self._trajectory

    @property
    def points(self) -> list[np.ndarray]:
        """Wrapped fractional coordinates of every atom-frame assigned to this site (reference site.py:70,
        site_collection.py:308: one ``atom.frac_coords`` per ``update_occupation``, frames in order, atoms in list order).
        Long batches (``Trajectory.trajectory_from_arrays`` above ``POINTS_LIMIT`` atom-frames) are kept as a reference to
        the caller's coordinate array plus the GPU-folded runs and expanded here, on first access."""
        if self._lazy_points:
            pending, self._lazy_points = self._lazy_points, []
            for block, pos in pending:
                col, start, length = block.site_runs(pos)
                if len(col) == 0:
                    continue
                frames = np.concatenate([np.arange(f0, f0 + n) for f0, n in zip(start.tolist(), length.tolist())])
                cols = np.repeat(col, length)
                order = np.lexsort((cols, frames))                   # frame by frame, atoms in list order
                xyz = block.frac[frames[order], block.atom_index[cols[order]]] % 1.0      # atom.py:104
                self._points.extend(xyz)
        return self._points

    @points.setter
    def points(self, value) -> None:
        self._lazy_points = []
        self._points = value

    def _folded_runs(self):
        """(atom index, start, end, n_frames) of this site's runs when its whole history sits in GPU-folded batches,
        else ``None`` (then the list form is authoritative)."""
        if self._trajectory or not self._lazy_blocks or not all(isinstance(b, BlockFolds) for b, _ in self._lazy_blocks):
            return None
        return site_runs_over_blocks([b for b, _ in self._lazy_blocks], [p for _, p in self._lazy_blocks])

    @trajectory.setter
    def trajectory(self, value) -> None:
        self._lazy_blocks = []
        self._trajectory = value

    def reset(self) -> None:
        """site.py:75-88 (``points`` is deliberately left alone, as in the reference)."""
        self.contains_atoms = []
        self._trajectory = []
        self._lazy_blocks = []
        self.transitions = Counter()

    def contains_point(self, x: np.ndarray) -> bool:
        raise NotImplementedError("contains_point should be implemented in the derived class")

    def contains_atom(self, atom) -> bool:
        return self.contains_point(atom.frac_coords)

    @property
    def centre(self) -> np.ndarray:
        raise NotImplementedError("centre should be implemented in the derived class")

    @property
    def coordination_number(self) -> int:
        raise NotImplementedError("coordination_number should be implemented in the derived class")

    @classmethod
    def reset_index(cls, newid: int = 0) -> None:
        Site._newid = newid

    def most_frequent_transitions(self) -> list[int]:
        """site.py:216-223: destinations by descending count (stable)."""
        t = self.transitions
        return sorted(t.keys(), key=t.get, reverse=True)

    @property
    def average_occupation(self) -> float | None:
        """site.py:225-239: fraction of appended timesteps with at least one atom in the site."""
        runs = self._folded_runs()
        if runs is not None:
            _, start, end, n_frames = runs
            return occupied_frames_from_runs(start, end) / n_frames if n_frames else None
        traj = self.trajectory
        if not traj:
            return None
        return sum(1 for ts in traj if ts) / len(traj)

    def residence_times(self, filter_length: int = 0, include_edge_runs: bool = False) -> tuple[int, ...]:
        """Per-atom run lengths of consecutive occupation (reference site.py:241-325; gap filter :327-352).

        Same arguments, result order (atoms by ascending index, runs in time order) and exceptions as the
        reference.  Histories appended by ``trajectory_from_arrays`` are answered from their GPU-folded run-length
        form; anything else from the list form."""
        if not isinstance(filter_length, int):
            raise TypeError(f"filter_length must be an integer, got {type(filter_length).__name__}")
        if filter_length < 0:
            raise ValueError(f"filter_length must be non-negative, got {filter_length}")
        runs = self._folded_runs()
        if runs is None:
            traj = self.trajectory
            if not traj:
                return ()
            n = len(traj)
            first: dict[int, list[list[int]]] = {}
            for f, ts in enumerate(traj):                    # list form -> runs [start, end) per atom
                for a in ts:
                    r = first.setdefault(a, [])
                    if r and r[-1][1] == f:
                        r[-1][1] = f + 1
                    elif not r or r[-1][1] < f:
                        r.append([f, f + 1])
            atoms = sorted(first)
            atom = np.array([a for a in atoms for _ in first[a]], dtype=np.int64)
            start = np.array([r[0] for a in atoms for r in first[a]], dtype=np.int64)
            end = np.array([r[1] for a in atoms for r in first[a]], dtype=np.int64)
            runs = (atom, start, end, n)
        atom, start, end, n_frames = runs
        return residence_times_from_runs(atom, start, end, n_frames, filter_length, include_edge_runs)

    def as_dict(self) -> dict:
        d = {"index": self.index, "contains_atoms": self.contains_atoms, "trajectory": self.trajectory,
             "points": self.points, "transitions": self.transitions}
        if self.label:
            d["label"] = self.label
        return d

    def summary(self, metrics: list[str] | None = None) -> dict:
        """site.py:353-419."""
        available = ["index", "label", "site_type", "average_occupation", "transitions"]
        defaults = metrics is None
        if metrics is None:
            metrics = ["index", "site_type", "average_occupation", "transitions"]
            if self.label is not None:
                metrics.insert(1, "label")
        invalid = set(metrics) - set(available)
        if invalid:
            raise ValueError(f"Invalid metric(s): {invalid}. Available metrics: {available}")
        out: dict[str, Any] = {}
        for m in metrics:
            if m == "site_type":
                v: Any = self.__class__.__name__
            elif m == "transitions":
                v = dict(self.transitions)
            else:
                v = getattr(self, m)
            if v is not None or not defaults:
                out[m] = v
        return out
