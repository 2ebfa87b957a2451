"""Polyhedral site (mirror of reference ``site_analysis/polyhedral_site.py``).

Holds the same definition and cache attributes as the reference class
(polyhedral_site.py:189-255) -- ``vertex_indices``, ``reference_center``, ``vertex_coords``,
``_pbc_image_shifts``, ``_pbc_cached_raw_frac``, ``_face_topology_cache`` -- so that sites
built by the reference's own ``TrajectoryBuilder`` can be handed over unchanged (duck typing).
Vertex gathering, face normals and containment run on the GPU for whole frames
(``csrc/k_polyhedral.cuh``); this class is data only.
"""
from __future__ import annotations

import numpy as np

from .site import Site


class FaceTopology:
    """Qhull face triangles of a polyhedron (reference ``FaceTopologyCache.face_simplices``)."""

    def __init__(self, face_simplices: np.ndarray) -> None:
        self.face_simplices = np.asarray(face_simplices, dtype=np.int32)


class PolyhedralSite(Site):
    def __init__(self, vertex_indices: list[int], label: str | None = None,
                 reference_center: np.ndarray | None = None):
        if isinstance(vertex_indices, np.ndarray):
            vertex_indices = vertex_indices.tolist()
        if not vertex_indices:
            raise ValueError("vertex_indices cannot be empty")
        if not all(isinstance(idx, int) for idx in vertex_indices):
            raise TypeError("All vertex indices must be integers")
        super().__init__(label=label)
        self.vertex_indices = vertex_indices
        self.vertex_coords: np.ndarray | None = None
        self._face_topology_cache: FaceTopology | None = None
        self._pbc_image_shifts: np.ndarray | None = None
        self._pbc_cached_raw_frac: np.ndarray | None = None
        self.reference_center = reference_center

    def __repr__(self) -> str:
        return (f"site_analysis.PolyhedralSite(index={self.index}, label={self.label}, "
                f"vertex_indices={self.vertex_indices}, contains_atoms={self.contains_atoms})")

    def reset(self) -> None:
        """polyhedral_site.py:264-279: PBC caches cleared, face topology kept."""
        super().reset()
        self.vertex_coords = None
        self._pbc_image_shifts = None
        self._pbc_cached_raw_frac = None

    @property
    def coordination_number(self) -> int:
        return len(self.vertex_indices)

    @property
    def cn(self) -> int:
        return self.coordination_number

    def assign_vertex_coords(self, all_frac_coords: np.ndarray, lattice_matrix: np.ndarray) -> None:
        """Prime the PBC-shift cache from one structure, as ``SiteFactory`` does
        (polyhedral_site.py:340-360, site_factory.py:248-257)."""
        from . import _tables as tab
        raw = np.asarray(all_frac_coords, dtype=np.float64)[self.vertex_indices]
        if self._pbc_image_shifts is not None and self._pbc_cached_raw_frac is not None:
            diff = raw - self._pbc_cached_raw_frac
            w = np.round(diff)
            if np.all(np.abs(diff - w) < 0.3):
                self._pbc_image_shifts = self._pbc_image_shifts - w.astype(np.int64)
                self._pbc_cached_raw_frac = raw.copy()
                self.vertex_coords = tab.unwrapped_vertices(raw, self._pbc_image_shifts)
                return
        shifts = tab.full_pbc_unwrap(raw, self.reference_center, lattice_matrix)
        self._pbc_image_shifts = shifts
        self._pbc_cached_raw_frac = raw.copy()
        self.vertex_coords = tab.unwrapped_vertices(raw, shifts, non_negative=self.reference_center is not None)

    def get_vertex_species(self, species: list[str]) -> list[str]:
        return [species[i] for i in self.vertex_indices]

    def contains_point(self, x, **kwargs) -> bool:
        raise NotImplementedError("use PolyhedralSiteCollection.assign_site_occupations (GPU path)")

    def as_dict(self) -> dict:
        d = super().as_dict()
        d["vertex_indices"] = self.vertex_indices
        d["vertex_coords"] = self.vertex_coords
        return d

    @classmethod
    def from_dict(cls, d):
        site = cls(vertex_indices=d["vertex_indices"])
        site.vertex_coords = d["vertex_coords"]
        site.contains_atoms = d["contains_atoms"]
        site.label = d.get("label")
        return site

    @property
    def centre(self) -> np.ndarray:
        if self.vertex_coords is None:
            raise RuntimeError("Vertex coordinates have not been assigned.")
        return np.array(np.mean(self.vertex_coords, axis=0))

    @classmethod
    def sites_from_vertex_indices(cls, vertex_indices, label=None):
        return [cls(vertex_indices=vi, label=label) for vi in vertex_indices]
