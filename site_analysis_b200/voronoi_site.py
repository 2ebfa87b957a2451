"""Voronoi site (mirror of reference ``site_analysis/voronoi_site.py``)."""
from __future__ import annotations

import numpy as np

from .site import Site


class VoronoiSite(Site):
    """Voronoi cell around a fixed fractional centre; membership is decided by the collection."""

    def __init__(self, frac_coords: np.ndarray, label: str | None = None) -> None:
        super().__init__(label=label)
        self.frac_coords = frac_coords

    def __repr__(self) -> str:
        return (f"site_analysis.VoronoiSite(index={self.index}, label={self.label}, "
                f"frac_coords={self.frac_coords}, contains_atoms={self.contains_atoms})")

    def as_dict(self) -> dict:
        d = super().as_dict()
        d["frac_coords"] = self.frac_coords
        return d

    @classmethod
    def from_dict(cls, d):
        site = cls(frac_coords=d["frac_coords"])
        site.label = d.get("label")
        return site

    @property
    def centre(self) -> np.ndarray:
        return self.frac_coords

    def contains_point(self, x: np.ndarray) -> bool:
        raise NotImplementedError
