"""Minimal duck-typed structure for array data (what the hot path reads from a pymatgen
``Structure``: ``frac_coords``, ``lattice.matrix``, iteration over sites with ``species_string``).
pymatgen ``Structure`` objects work unchanged wherever a structure is accepted."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SimpleLattice:
    matrix: np.ndarray


class _SiteView:
    def __init__(self, species: str, frac: np.ndarray):
        self.species_string = species
        self.specie = species
        self.species = species
        self.frac_coords = frac


class SimpleStructure:
    def __init__(self, lattice, species, frac_coords):
        self.lattice = SimpleLattice(np.asarray(lattice, dtype=np.float64).reshape(3, 3))
        self._species = [str(s) for s in species]
        self.frac_coords = np.asarray(frac_coords, dtype=np.float64).reshape(-1, 3)

    def __len__(self):
        return len(self._species)

    def __bool__(self):
        return True

    def __iter__(self):
        for s, f in zip(self._species, self.frac_coords):
            yield _SiteView(s, f)

    def __getitem__(self, i):
        return _SiteView(self._species[i], self.frac_coords[i])
