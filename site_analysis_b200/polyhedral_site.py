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
        self._pending_frac_coords: np.ndarray | None = None
        self._pending_lattice_matrix: np.ndarray | None = None
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

    def notify_structure_changed(self, all_frac_coords: np.ndarray, lattice_matrix: np.ndarray) -> None:
        """Mark the vertex coordinates stale; they are re-derived when the site is next queried
        (polyhedral_site.py:322-338)."""
        self._pending_frac_coords = all_frac_coords
        self._pending_lattice_matrix = lattice_matrix

    def _query_ready(self):
        """Vertex coordinates and face topology for a query, as ``contains_point`` prepares them
        (polyhedral_site.py:459-471): pending coordinates applied, topology hulled once and kept."""
        if self._pending_frac_coords is not None and self._pending_lattice_matrix is not None:
            frac, lat = self._pending_frac_coords, self._pending_lattice_matrix
            self._pending_frac_coords = self._pending_lattice_matrix = None
            self.assign_vertex_coords(frac, lat)
        if self.vertex_coords is None:
            raise RuntimeError(f"no vertex coordinates set for polyhedral_site {self.index}")
        if self._face_topology_cache is None:
            from . import _tables as tab
            try:
                self._face_topology_cache = FaceTopology(tab.hull_simplices(self.vertex_coords, self.index))
            except RuntimeError as exc:
                raise RuntimeError(f"Degenerate vertex geometry for polyhedral_site {self.index} "
                                   f"(vertices {self.vertex_indices})") from exc.__cause__
        return self.vertex_coords, self._face_topology_cache.face_simplices

    def contains_point(self, x: np.ndarray, *, algo: str | None = None, pbc_images: np.ndarray | None = None) -> bool:
        """Is the point (or any of ``pbc_images``) inside this polyhedron?  Reference polyhedral_site.py:431-485; the
        face-plane test runs on the GPU (``sab_polyhedra_contain``, float64, the reference's operation order).  Whole
        trajectories go through ``PolyhedralSiteCollection`` instead -- this is the single-point query."""
        if algo is not None:
            import warnings
            warnings.warn("The 'algo' parameter is deprecated and will be removed in a future version. The best available "
                          "containment algorithm is now selected automatically.", DeprecationWarning, stacklevel=2)
        from .distances import polyhedra_contain, x_pbc
        vc, simp = self._query_ready()
        images = np.asarray(pbc_images, dtype=np.float64).reshape(-1, 3) if pbc_images is not None else x_pbc(x)
        return bool(polyhedra_contain([vc], [simp], [images])[0])

    def contains_atom(self, atom, *, algo: str | None = None, pbc_images: np.ndarray | None = None) -> bool:
        """polyhedral_site.py:497-520."""
        if algo is not None:
            import warnings
            warnings.warn("The 'algo' parameter is deprecated and will be removed in a future version. The best available "
                          "containment algorithm is now selected automatically.", DeprecationWarning, stacklevel=2)
        return self.contains_point(atom.frac_coords, pbc_images=pbc_images)

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
