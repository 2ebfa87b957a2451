"""Dynamic Voronoi site (mirror of reference ``site_analysis/dynamic_voronoi_site.py``)."""
from __future__ import annotations

import numpy as np

from .site import Site


class DynamicVoronoiSite(Site):
    """Voronoi cell whose centre is the PBC-aware mean of ``reference_indices`` atoms."""

    def __init__(self, reference_indices: list[int], label: str | None = None,
                 reference_center: np.ndarray | None = None) -> None:
        super().__init__(label=label)
        self.reference_indices = reference_indices
        self._centre_coords: np.ndarray | None = None
        self.reference_center = reference_center

    def __repr__(self) -> str:
        return (f"site_analysis.DynamicVoronoiSite(index={self.index}, label={self.label}, "
                f"reference_indices={self.reference_indices}, contains_atoms={self.contains_atoms})")

    def reset(self) -> None:
        super().reset()
        self._centre_coords = None

    def calculate_centre(self, all_frac_coords: np.ndarray, lattice_matrix: np.ndarray) -> None:
        """Centre of this site from its reference atoms in one structure (dynamic_voronoi_site.py:88-104):
        ``mean(correct_pbc(ref_coords, reference_center, lattice)) % 1.0``.  The integer image shifts come from the same
        numpy calls as the reference's (``_tables.full_pbc_unwrap``); the mean is taken on the GPU by the engine's centre
        kernel (``sab_dynvor_centres``), the code path whole trajectories use."""
        import torch
        from .engine import AssignmentEngine
        frac = np.ascontiguousarray(all_frac_coords, dtype=np.float64)
        lat = np.ascontiguousarray(lattice_matrix, dtype=np.float64)
        eng = AssignmentEngine("dynamic_voronoi", len(frac), [int(self.reference_indices[0])],
                               slots=[list(self.reference_indices)], reference_centres=[self.reference_center])
        eng._init_pbc(frac, lat)
        d = torch.from_numpy(frac[None]).to(torch.device("cuda", eng.device))
        self._centre_coords = eng.dynvor_centres(d)[0, 0].cpu().numpy()

    @property
    def centre(self) -> np.ndarray:
        if self._centre_coords is None:
            raise RuntimeError("Centre coordinates for this DynamicVoronoiSite have not been calculated yet.")
        return self._centre_coords

    def contains_point(self, x: np.ndarray) -> bool:
        raise NotImplementedError

    def as_dict(self) -> dict:
        d = super().as_dict()
        d["reference_indices"] = self.reference_indices
        if self._centre_coords is not None:
            d["centre_coords"] = self._centre_coords.tolist()
        return d

    @classmethod
    def from_dict(cls, d: dict) -> "DynamicVoronoiSite":
        site = cls(reference_indices=d["reference_indices"])
        site.label = d.get("label")
        site.contains_atoms = d.get("contains_atoms", [])
        site.trajectory = d.get("trajectory", [])
        site.points = d.get("points", [])
        if "transitions" in d:
            site.transitions = d["transitions"]
        if "centre_coords" in d:
            site._centre_coords = np.array(d["centre_coords"])
        return site

    @property
    def coordination_number(self) -> int:
        return len(self.reference_indices)

    @property
    def cn(self) -> int:
        return self.coordination_number

    @classmethod
    def sites_from_reference_indices(cls, reference_indices_list, label=None):
        return [cls(reference_indices=ri, label=label) for ri in reference_indices_list]
