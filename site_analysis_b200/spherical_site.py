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
        """Single-point queries are answered by the collection on the GPU
        (``SphericalSiteCollection.sites_containing``); the per-site Python test of the
        reference (spherical_site.py:126-146) is not part of the accelerated path."""
        if lattice_matrix is None:
            raise ValueError("lattice_matrix is required for SphericalSite.contains_point()")
        raise NotImplementedError("use SphericalSiteCollection.assign_site_occupations (GPU path)")

    def contains_atom(self, atom, *, lattice_matrix=None) -> bool:
        if lattice_matrix is None:
            raise ValueError("lattice_matrix is required for SphericalSite.contains_atom()")
        return self.contains_point(atom.frac_coords, lattice_matrix=lattice_matrix)
