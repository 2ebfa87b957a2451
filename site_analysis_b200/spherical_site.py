"""Spherical site (mirror of reference ``site_analysis/spherical_site.py``)."""
from __future__ import annotations

import numpy as np

from .site import Site


class SphericalSite(Site):
    """Sphere of radius ``rcut`` (Angstrom) centred at fractional ``frac_coords``."""

    def __init__(self, frac_coords: np.ndarray, rcut: float, label: str | None = None) -> None:
        super().__init__(label=label)
        self.frac_coords = frac_coords
        self.rcut = rcut

    def __repr__(self) -> str:
        return (f"site_analysis.SphericalSite(index={self.index}, label={self.label}, "
                f"frac_coords={self.frac_coords}, rcut={self.rcut}, contains_atoms={self.contains_atoms})")

    @property
    def centre(self) -> np.ndarray:
        return self.frac_coords

    def as_dict(self) -> dict:
        d = super().as_dict()
        d["frac_coords"] = self.frac_coords
        d["rcut"] = self.rcut
        return d

    @classmethod
    def from_dict(cls, d: dict) -> "SphericalSite":
        site = cls(frac_coords=d["frac_coords"], rcut=d["rcut"])
        site.label = d.get("label")
        return site

    def contains_point(self, x, *, lattice_matrix=None) -> bool:
        """``mic_distance(centre, x) <= rcut`` (reference spherical_site.py:126-146), the distance computed by the GPU
        operator :func:`site_analysis_b200.distances.mic_distance` (bit-identical to the reference's numba kernel).
        Whole trajectories go through ``SphericalSiteCollection`` instead -- this is the single-point query."""
        if lattice_matrix is None:
            raise ValueError("lattice_matrix is required for SphericalSite.contains_point()")
        from .distances import mic_distance
        return bool(mic_distance(np.asarray(self.frac_coords, dtype=np.float64), np.asarray(x, dtype=np.float64),
                                 lattice_matrix) <= self.rcut)

    def contains_atom(self, atom, *, lattice_matrix=None) -> bool:
        if lattice_matrix is None:
            raise ValueError("lattice_matrix is required for SphericalSite.contains_atom()")
        return self.contains_point(atom.frac_coords, lattice_matrix=lattice_matrix)
