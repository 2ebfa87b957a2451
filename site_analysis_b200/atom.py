"""Mobile-ion record (mirror of reference ``site_analysis/atom.py``).

Same attributes and behaviour as the reference ``Atom`` (atom.py:25-192).  The only
difference is that ``trajectory`` is materialised lazily from the int32 result columns the
GPU path produces (no per-frame Python objects in the hot loop).
"""
from __future__ import annotations

import gzip
import json
from typing import Any

import numpy as np


class Atom:
    def __init__(self, index: int, species_string: str | None = None) -> None:
        self.index = index
        self.species_string = species_string
        self.in_site: int | None = None
        self._frac_coords: np.ndarray | None = None
        self._trajectory: list[int | None] = []
        self._lazy_blocks: list[np.ndarray] = []     # pending (F,) int32 columns, -1 = None
        self._recent_sites: list[int | None] = [None, None]

    @property
    def trajectory(self) -> list[int | None]:
        """Site index per appended timestep (``None`` = unassigned), as in the reference.

        Batches analysed on the GPU arrive as int32 columns and are turned into list entries
        only when somebody looks."""
        if self._lazy_blocks:
            for col in self._lazy_blocks:
                self._trajectory.extend(None if v < 0 else v for v in col.tolist())
            self._lazy_blocks = []
        return self._trajectory

    @trajectory.setter
    def trajectory(self, value) -> None:
        self._lazy_blocks = []
        self._trajectory = value

    def _last_trajectory_entry(self):
        """``atom.trajectory[-1]`` without materialising: -2 empty, -1 None, else site index."""
        if self._lazy_blocks:
            return int(self._lazy_blocks[-1][-1])
        if not self._trajectory:
            return -2
        v = self._trajectory[-1]
        return -1 if v is None else int(v)

    def __str__(self) -> str:
        return f"Atom: {self.index}"

    def __repr__(self) -> str:
        return (f"site_analysis.Atom(index={self.index}, in_site={self.in_site}, "
                f"frac_coords={self._frac_coords}, species_string={self.species_string})")

    def reset(self) -> None:
        """atom.py:81-93."""
        self.in_site = None
        self._frac_coords = None
        self._trajectory = []
        self._lazy_blocks = []
        self._recent_sites = [None, None]

    def assign_coords(self, frac_coords: np.ndarray) -> None:
        """atom.py:95-104: ``frac_coords[self.index] % 1.0`` (numpy remainder semantics)."""
        self._frac_coords = frac_coords[self.index] % 1.0

    @property
    def frac_coords(self) -> np.ndarray:
        if self._frac_coords is None:
            raise AttributeError("Coordinates not set for atom {}".format(self.index))
        return self._frac_coords

    @property
    def most_recent_site(self) -> int | None:
        return self._recent_sites[0]

    def update_recent_site(self, site_index: int) -> None:
        """atom.py:181-192."""
        if site_index != self._recent_sites[0]:
            self._recent_sites[1] = self._recent_sites[0]
            self._recent_sites[0] = site_index

    # -- serialisation (atom.py:118-166) ---------------------------------------------------
    def as_dict(self) -> dict[str, Any]:
        d: dict[str, Any] = {"index": int(self.index),
                             "in_site": None if self.in_site is None else int(self.in_site)}
        if self._frac_coords is not None:
            d["frac_coords"] = self._frac_coords.tolist()
        if self.species_string is not None:
            d["species_string"] = self.species_string
        return d

    @classmethod
    def from_dict(cls, d: dict) -> "Atom":
        atom = cls(index=d["index"], species_string=d.get("species_string"))
        atom.in_site = None if d["in_site"] is None else int(d["in_site"])
        if "frac_coords" in d:
            atom._frac_coords = np.array(d["frac_coords"])
        return atom

    def to(self, filename: str | None = None) -> str:
        s = json.dumps(self.as_dict())
        if filename:
            opener = gzip.open if str(filename).endswith(".gz") else open
            with opener(filename, "wb") as f:
                f.write(s.encode("utf-8"))
        return s

    @classmethod
    def from_str(cls, input_string: str) -> "Atom":
        return cls.from_dict(json.loads(input_string))

    @classmethod
    def from_file(cls, filename) -> "Atom":
        opener = gzip.open if str(filename).endswith(".gz") else open
        with opener(filename, "rt") as f:
            return cls.from_str(f.read())


def atoms_from_species_string(structure, species_string: str) -> list[Atom]:
    return [Atom(index=i) for i, s in enumerate(structure) if s.species_string == species_string]


def atoms_from_structure(structure, species_string) -> list[Atom]:
    if isinstance(species_string, str):
        species_string = [species_string]
    atoms = [Atom(index=i, species_string=s.species_string)
             for i, s in enumerate(structure) if s.species_string in species_string]
    frac = structure.frac_coords
    for a in atoms:
        a.assign_coords(frac)
    return atoms


def atoms_from_indices(indices: list[int]) -> list[Atom]:
    return [Atom(index=i) for i in indices]
