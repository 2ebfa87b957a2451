"""Fluent builder (mirror of the per-frame-relevant part of reference ``site_analysis/builders.py``).

Kept: ``with_structure``, ``with_reference_structure``, ``with_mobile_species``,
``with_spherical_sites``, ``with_voronoi_sites``, ``with_existing_sites``, ``with_existing_atoms``,
``build`` and the two centre-based factory functions, with the reference's validation and error
messages (builders.py:141-175, 312-383, 638-658, 696-800).

Out of scope here: the reference-structure workflow that *derives*
polyhedral / dynamic-Voronoi sites (alignment, coordination search, index mapping; 1.4 kLoC that
run once per analysis).  Sites of those types are passed in ready-made --
``with_polyhedral_sites_from_indices`` / ``with_dynamic_voronoi_sites_from_indices``, or
``with_existing_sites`` with sites produced by the reference's own builder.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from .atom import Atom, atoms_from_structure
from .dynamic_voronoi_site import DynamicVoronoiSite
from .polyhedral_site import PolyhedralSite
from .site import Site
from .spherical_site import SphericalSite
from .trajectory import Trajectory
from .voronoi_site import VoronoiSite

_WORKFLOW_MSG = ("deriving {0} sites from a reference structure is a one-off host step outside the "
                 "accelerated path; build the sites with the reference package and pass them via "
                 "with_existing_sites(), or use with_{1}_sites_from_indices()")


class TrajectoryBuilder:
    def __init__(self) -> None:
        self.reset()

    def reset(self) -> "TrajectoryBuilder":
        self._structure = None
        self._reference_structure = None
        self._mobile_species: str | list[str] | None = None
        self._atoms: list[Atom] | None = None
        self._site_generators: list[Callable] = []
        return self

    def with_structure(self, structure) -> "TrajectoryBuilder":
        self._structure = structure
        return self

    def with_reference_structure(self, reference_structure) -> "TrajectoryBuilder":
        self._reference_structure = reference_structure
        return self

    def with_mobile_species(self, species) -> "TrajectoryBuilder":
        self._mobile_species = species
        return self

    def with_spherical_sites(self, centres, radii, labels=None) -> "TrajectoryBuilder":
        if isinstance(radii, (float, int)):
            radii = [radii] * len(centres)
        if isinstance(labels, str):
            labels = [labels] * len(centres)
        if len(centres) != len(radii):
            raise ValueError("Number of centres must match number of radii")
        if labels is not None and len(centres) != len(labels):
            raise ValueError("Number of centres must match number of labels")

        def create() -> Sequence[SphericalSite]:
            return [SphericalSite(frac_coords=np.array(c), rcut=r,
                                  label=(labels[i] if labels and i < len(labels) else None))
                    for i, (c, r) in enumerate(zip(centres, radii))]
        self._site_generators.append(create)
        return self

    def with_voronoi_sites(self, centres, labels=None) -> "TrajectoryBuilder":
        def create() -> Sequence[VoronoiSite]:
            return [VoronoiSite(frac_coords=np.array(c), label=(labels[i] if labels and i < len(labels) else None))
                    for i, c in enumerate(centres)]
        self._site_generators.append(create)
        return self

    def with_polyhedral_sites_from_indices(self, vertex_indices, reference_centers=None, labels=None,
                                           prime: bool = True) -> "TrajectoryBuilder":
        """Polyhedral sites from explicit vertex atom indices.  ``prime`` assigns vertex coordinates
        from the builder's structure, as the reference's SiteFactory does (site_factory.py:110-111)."""
        if isinstance(labels, str):
            labels = [labels] * len(vertex_indices)

        def create() -> Sequence[PolyhedralSite]:
            sites = []
            for i, v in enumerate(vertex_indices):
                rc = None if reference_centers is None else np.asarray(reference_centers[i], dtype=np.float64)
                site = PolyhedralSite(vertex_indices=[int(x) for x in v], label=(labels[i] if labels else None),
                                      reference_center=rc)
                if prime:
                    site.assign_vertex_coords(np.asarray(self._structure.frac_coords),
                                              np.asarray(self._structure.lattice.matrix))
                sites.append(site)
            return sites
        self._site_generators.append(create)
        return self

    def with_dynamic_voronoi_sites_from_indices(self, reference_indices, reference_centers=None,
                                                labels=None) -> "TrajectoryBuilder":
        if isinstance(labels, str):
            labels = [labels] * len(reference_indices)

        def create() -> Sequence[DynamicVoronoiSite]:
            return [DynamicVoronoiSite(reference_indices=[int(x) for x in r], label=(labels[i] if labels else None),
                                       reference_center=(None if reference_centers is None
                                                         else np.asarray(reference_centers[i], dtype=np.float64)))
                    for i, r in enumerate(reference_indices)]
        self._site_generators.append(create)
        return self

    def with_polyhedral_sites(self, *args, **kwargs):
        raise NotImplementedError(_WORKFLOW_MSG.format("polyhedral", "polyhedral"))

    def with_dynamic_voronoi_sites(self, *args, **kwargs):
        raise NotImplementedError(_WORKFLOW_MSG.format("dynamic Voronoi", "dynamic_voronoi"))

    def with_existing_sites(self, sites: list) -> "TrajectoryBuilder":
        self._site_generators.append(lambda: sites)
        return self

    def with_existing_atoms(self, atoms: list) -> "TrajectoryBuilder":
        self._atoms = atoms
        return self

    @staticmethod
    def _validate_unique_sites(sites) -> None:
        if not sites:
            return
        attr = {"PolyhedralSite": "vertex_indices", "DynamicVoronoiSite": "reference_indices"}.get(type(sites[0]).__name__)
        if attr is None:
            return
        seen: dict[frozenset, int] = {}
        for site in sites:
            key = frozenset(getattr(site, attr))
            if key in seen:
                raise ValueError(f"Duplicate sites: site {seen[key]} and site {site.index} share the same {attr} "
                                 f"{sorted(key)}. This typically means the reference structure has multiple atoms "
                                 f"inside the same coordination environment.")
            seen[key] = site.index

    def build(self) -> Trajectory:
        if not self._structure:
            raise ValueError("Structure must be set")
        if not self._site_generators:
            raise ValueError("Site type must be defined using one of the with_*_sites methods")
        Site.reset_index()
        sites: list[Site] = []
        site_type = None
        for generator in self._site_generators:
            generated = generator()
            if generated:
                current = type(generated[0])
                if site_type is None:
                    site_type = current
                elif site_type != current:
                    raise TypeError(f"Cannot mix site types: {site_type.__name__} and {current.__name__}")
            sites.extend(generated)
        self._validate_unique_sites(sites)
        if not self._atoms:
            if not self._mobile_species:
                raise ValueError("Mobile species must be set")
            self._atoms = atoms_from_structure(self._structure, self._mobile_species)
            if not self._atoms:
                available = sorted(set(str(site.specie) for site in self._structure))
                raise ValueError(f"No atoms found matching mobile species '{self._mobile_species}'. "
                                 f"Available species in structure: {available}")
        trajectory = Trajectory(sites=sites, atoms=self._atoms)
        self.reset()
        return trajectory


def create_trajectory_with_spherical_sites(structure, mobile_species, centres, radii, labels=None) -> Trajectory:
    return (TrajectoryBuilder().with_structure(structure).with_mobile_species(mobile_species)
            .with_spherical_sites(centres, radii, labels).build())


def create_trajectory_with_voronoi_sites(structure, mobile_species, centres, labels=None) -> Trajectory:
    return (TrajectoryBuilder().with_structure(structure).with_mobile_species(mobile_species)
            .with_voronoi_sites(centres, labels).build())
