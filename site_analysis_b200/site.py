"""Site base class (mirror of reference ``site_analysis/site.py``).

Attributes as in the reference (site.py:24-88): ``index`` (global counter), ``label``,
``contains_atoms``, ``trajectory``, ``points``, ``transitions``.  ``trajectory`` is
materialised lazily from the array-backed assignment results of GPU batches.
"""
from __future__ import annotations

from collections import Counter
from typing import Any

import numpy as np


class Site:
    _newid = 0

    def __init__(self, label: str | None = None) -> None:
        self.index = Site._newid
        Site._newid += 1
        self.label = label
        self.contains_atoms: list[int] = []
        self._trajectory: list[list[int]] = []
        self._lazy_blocks: list = []                 # pending (results (F, A), atom_index (A,)) pairs
        self.points: list[np.ndarray] = []
        self.transitions: Counter = Counter()

    @property
    def trajectory(self) -> list[list[int]]:
        """Atom indices occupying this site at each appended timestep (atom list order)."""
        if self._lazy_blocks:
            for block, atom_index in self._lazy_blocks:
                frames, cols = np.nonzero(block == self.index)
                per_frame: list[list[int]] = [[] for _ in range(block.shape[0])]
                for f, a in zip(frames.tolist(), atom_index[cols].tolist()):
                    per_frame[f].append(a)
                self._trajectory.extend(per_frame)
            self._lazy_blocks = []
        return self._trajectory

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
        traj = self.trajectory
        if not traj:
            return None
        return sum(1 for ts in traj if ts) / len(traj)

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
