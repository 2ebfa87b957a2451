"""Duck-typed stand-in for ``pymatgen.core`` (TEST INFRASTRUCTURE ONLY).

pymatgen is not installed in this image.  The reference package at
``/root/reference`` imports ``pymatgen.core.Structure`` for type hints and touches
only ``structure.frac_coords``, ``structure.lattice.matrix``, iteration over
sites (``.species_string`` / ``.specie`` / ``.species`` / ``.frac_coords``),
``len``, item get/set, ``copy``, ``append`` and ``remove_species``.  This module
provides exactly that surface so that the UNMODIFIED reference can be imported
by ``oracle/gen_golden.py`` and ``oracle/validate_oracle.py`` in the build
container.  Nothing in the product package imports it.
"""
from __future__ import annotations

import itertools

import numpy as np


class Lattice:
    def __init__(self, matrix):
        self._matrix = np.array(matrix, dtype=np.float64).reshape(3, 3)

    @property
    def matrix(self) -> np.ndarray:
        return self._matrix

    @classmethod
    def cubic(cls, a: float) -> "Lattice":
        return cls([[a, 0.0, 0.0], [0.0, a, 0.0], [0.0, 0.0, a]])

    @classmethod
    def from_parameters(cls, a, b, c, alpha, beta, gamma) -> "Lattice":
        al, be, ga = np.radians([alpha, beta, gamma])
        cos_al, cos_be, cos_ga = np.cos([al, be, ga])
        sin_al, sin_be = np.sin([al, be])
        val = (cos_al * cos_be - cos_ga) / (sin_al * sin_be)
        val = max(min(val, 1.0), -1.0)
        gamma_star = np.arccos(val)
        va = [a * sin_be, 0.0, a * cos_be]
        vb = [-b * sin_al * np.cos(gamma_star), b * sin_al * np.sin(gamma_star), b * cos_al]
        vc = [0.0, 0.0, float(c)]
        return cls([va, vb, vc])

    @property
    def abc(self):
        return tuple(np.linalg.norm(self._matrix, axis=1))

    @property
    def volume(self) -> float:
        return float(abs(np.linalg.det(self._matrix)))

    def get_cartesian_coords(self, frac):
        return np.asarray(frac) @ self._matrix

    def get_fractional_coords(self, cart):
        return np.asarray(cart) @ np.linalg.inv(self._matrix)

    def copy(self):
        return Lattice(self._matrix.copy())


class Species(str):
    """A species is just its string here."""

    @property
    def symbol(self):
        return str(self)


class Site:
    pass


class PeriodicSite(Site):
    def __init__(self, species, frac_coords, lattice, index=None):
        self.species_string = str(species)
        self.frac_coords = np.array(frac_coords, dtype=np.float64)
        self.lattice = lattice
        self.index = index

    @property
    def specie(self):
        return Species(self.species_string)

    @property
    def species(self):
        return Species(self.species_string)

    @property
    def coords(self):
        return self.frac_coords @ self.lattice.matrix


class Structure:
    def __init__(self, lattice, species, coords, coords_are_cartesian=False, **_):
        self.lattice = lattice if isinstance(lattice, Lattice) else Lattice(lattice)
        coords = np.array(coords, dtype=np.float64).reshape(-1, 3)
        if coords_are_cartesian:
            coords = coords @ np.linalg.inv(self.lattice.matrix)
        self._species = [str(s) for s in species]
        self._frac = coords

    # -- the surface the reference uses ---------------------------------
    @property
    def frac_coords(self) -> np.ndarray:
        return self._frac.copy()

    @property
    def cart_coords(self) -> np.ndarray:
        return self._frac @ self.lattice.matrix

    @property
    def species(self):
        return [Species(s) for s in self._species]

    def __len__(self):
        return len(self._species)

    def __bool__(self):
        return True

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __getitem__(self, i):
        return PeriodicSite(self._species[i], self._frac[i], self.lattice, index=i)

    def __setitem__(self, i, value):
        species, frac = value
        self._species[i] = str(species)
        self._frac[i] = np.asarray(frac, dtype=np.float64)

    def copy(self):
        return Structure(self.lattice.copy(), list(self._species), self._frac.copy())

    def append(self, species, coords, coords_are_cartesian=False, **_):
        c = np.asarray(coords, dtype=np.float64)
        if coords_are_cartesian:
            c = c @ np.linalg.inv(self.lattice.matrix)
        self._species.append(str(species))
        self._frac = np.vstack([self._frac, c[None, :]])

    def remove_species(self, species):
        keep = [i for i, s in enumerate(self._species) if s not in set(map(str, species))]
        self._species = [self._species[i] for i in keep]
        self._frac = self._frac[keep]

    def __mul__(self, scale):
        """Supercell: grouped by original site, translations innermost."""
        scale = np.array(scale, dtype=int).reshape(3)
        trans = list(itertools.product(range(scale[0]), range(scale[1]), range(scale[2])))
        new_species, new_frac = [], []
        for sp, f in zip(self._species, self._frac):
            for t in trans:
                new_species.append(sp)
                new_frac.append((f + np.array(t, dtype=float)) / scale)
        return Structure(Lattice(self.lattice.matrix * scale[:, None]), new_species, new_frac)

    # -- minimal space-group expansion -----------------------------------
    @classmethod
    def from_spacegroup(cls, sg, lattice, species, coords, tol=1e-5, **_):
        ops = _SPACEGROUP_OPS[sg.replace(" ", "")]()
        all_species, all_coords = [], []
        for sp, c in zip(species, np.array(coords, dtype=float).reshape(-1, 3)):
            orbit: list[np.ndarray] = []
            for rot, tr in ops:
                p = (rot @ c + tr) % 1.0
                p[np.abs(p - 1.0) < tol] = 0.0
                dup = False
                for q in orbit:
                    d = p - q
                    d -= np.round(d)
                    if np.all(np.abs(d) < tol):
                        dup = True
                        break
                if not dup:
                    orbit.append(p)
            all_species.extend([sp] * len(orbit))
            all_coords.extend(orbit)
        return cls(lattice, all_species, all_coords)


def _point_group_m3m():
    """All 48 signed permutation matrices (O_h)."""
    mats = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            m = np.zeros((3, 3))
            for r in range(3):
                m[r, perm[r]] = signs[r]
            mats.append(m)
    return mats


def _ops_f43m():
    """F-4 3m (216): signed permutations with det(perm)*prod(signs) pattern of T_d."""
    centring = [np.zeros(3), np.array([0, .5, .5]), np.array([.5, 0, .5]), np.array([.5, .5, 0])]
    ops = []
    for m in _point_group_m3m():
        n_neg = int((m.sum(axis=1) < 0).sum())
        if n_neg % 2 == 0:  # T_d: even number of sign changes for every permutation
            for t in centring:
                ops.append((m, t))
    return ops


def _ops_fm3m():
    centring = [np.zeros(3), np.array([0, .5, .5]), np.array([.5, 0, .5]), np.array([.5, .5, 0])]
    return [(m, t) for m in _point_group_m3m() for t in centring]


def _ops_pm3m():
    return [(m, np.zeros(3)) for m in _point_group_m3m()]


_SPACEGROUP_OPS = {"F-43m": _ops_f43m, "Fm-3m": _ops_fm3m, "Pm-3m": _ops_pm3m}
