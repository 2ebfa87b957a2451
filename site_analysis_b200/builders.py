"""Fluent builder (mirror of reference ``site_analysis/builders.py``).

``with_structure``, ``with_reference_structure``, ``with_mobile_species``, ``with_structure_alignment``,
``with_site_mapping``, ``with_min_atom_distance``, ``with_spherical_sites``, ``with_voronoi_sites``,
``with_polyhedral_sites``, ``with_dynamic_voronoi_sites``, ``with_existing_sites``, ``with_existing_atoms``, ``build`` and
the four factory functions, with the reference's arguments, validation and error messages (builders.py:99-1047).
Polyhedral / dynamic-Voronoi sites are derived from the reference structure by
:mod:`site_analysis_b200.reference_workflow` (alignment, coordination search, index mapping -- every distance matrix
on the GPU).  ``with_polyhedral_sites_from_indices`` / ``with_dynamic_voronoi_sites_from_indices`` take ready-made index
lists, and ``with_existing_sites`` accepts sites produced by the reference's own builder.
"""
from __future__ import annotations

from typing import Any, Callable, Sequence

import numpy as np

from .atom import Atom, atoms_from_structure
from .dynamic_voronoi_site import DynamicVoronoiSite
from .polyhedral_site import PolyhedralSite
from .site import Site
from .spherical_site import SphericalSite
from .trajectory import Trajectory
from .voronoi_site import VoronoiSite


class TrajectoryBuilder:
    def __init__(self) -> None:
        self.reset()

    def reset(self) -> "TrajectoryBuilder":
        self._structure = None
        self._reference_structure = None
        self._mobile_species: str | list[str] | None = None
        self._atoms: list[Atom] | None = None
        # alignment / mapping / validation options (builders.py:123-137)
        self._align = True
        self._align_species: list[str] | None = None
        self._align_metric = "rmsd"
        self._align_algorithm = "Nelder-Mead"
        self._align_minimizer_options: dict[str, Any] | None = None
        self._align_tolerance = 1e-4
        self._mapping_species: list[str] | None = None
        self._min_atom_distance: float = 0.5
        self._site_generators: list[Callable] = []
        return self

    def with_structure_alignment(self, align: bool = True, align_species: str | list[str] | None = None,
                                 align_metric: str = "rmsd", align_algorithm: str = "Nelder-Mead",
                                 align_minimizer_options: dict[str, Any] | None = None,
                                 align_tolerance: float = 1e-4) -> "TrajectoryBuilder":
        """builders.py:178-248.  Alignment is ON by default for reference-derived sites."""
        self._align = align
        self._align_species = [align_species] if isinstance(align_species, str) else align_species
        self._align_metric = align_metric
        self._align_algorithm = align_algorithm
        self._align_minimizer_options = align_minimizer_options
        self._align_tolerance = align_tolerance
        return self

    def with_site_mapping(self, mapping_species: str | list[str] | None) -> "TrajectoryBuilder":
        """builders.py:250-275."""
        self._mapping_species = [mapping_species] if isinstance(mapping_species, str) else mapping_species
        return self

    def with_min_atom_distance(self, distance: float) -> "TrajectoryBuilder":
        """builders.py:277-310: smallest allowed same-species distance in the reference structure (0 disables)."""
        if distance < 0:
            raise ValueError(f"min_atom_distance must be non-negative, got {distance}")
        self._min_atom_distance = distance
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

    def _reference_based_sites(self, what: str):
        """The configured workflow object for this builder's (reference, target) pair (builders.py:460-483)."""
        from .reference_workflow import ReferenceBasedSites
        if not self._structure or not self._reference_structure:
            raise ValueError(f"Both structure and reference_structure must be set for {what} sites")
        align_species = self._align_species
        if self._align and align_species is None and self._mapping_species is not None:
            align_species = self._mapping_species
        rbs = ReferenceBasedSites(reference_structure=self._reference_structure, target_structure=self._structure,
                                  align=self._align, align_species=align_species, align_metric=self._align_metric,
                                  align_algorithm=self._align_algorithm,
                                  align_minimizer_options=self._align_minimizer_options,
                                  align_tolerance=self._align_tolerance)
        target_species = self._mapping_species if self._mapping_species is not None else self._align_species
        return rbs, target_species

    def with_polyhedral_sites(self, centre_species: str, vertex_species: str | list[str], cutoff: float, n_vertices: int,
                              label: str | None = None, use_reference_centers: bool = True) -> "TrajectoryBuilder":
        """builders.py:385-510: polyhedra around every ``centre_species`` atom of the reference that has exactly
        ``n_vertices`` ``vertex_species`` atoms within ``cutoff``, mapped onto the target structure at ``build()``."""
        def create() -> Sequence[PolyhedralSite]:
            rbs, target_species = self._reference_based_sites("polyhedral")
            sites = rbs.create_polyhedral_sites(center_species=centre_species, vertex_species=vertex_species, cutoff=cutoff,
                                                n_vertices=n_vertices, label=label, target_species=target_species,
                                                use_reference_centers=use_reference_centers)
            if not sites:
                raise ValueError(f"No polyhedral sites found for centre_species='{centre_species}', "
                                 f"vertex_species='{vertex_species}', cutoff={cutoff}, n_vertices={n_vertices}. "
                                 f"Try adjusting these parameters or verify that the specified species exist "
                                 f"in the structure.")
            return sites
        self._site_generators.append(create)
        return self

    def with_dynamic_voronoi_sites(self, centre_species: str, reference_species: str | list[str], cutoff: float,
                                   n_reference: int, label: str | None = None,
                                   use_reference_centers: bool = True) -> "TrajectoryBuilder":
        """builders.py:512-636."""
        def create() -> Sequence[DynamicVoronoiSite]:
            rbs, target_species = self._reference_based_sites("dynamic Voronoi")
            sites = rbs.create_dynamic_voronoi_sites(center_species=centre_species, reference_species=reference_species,
                                                     cutoff=cutoff, n_reference=n_reference, label=label,
                                                     target_species=target_species,
                                                     use_reference_centers=use_reference_centers)
            if not sites:
                raise ValueError(f"No dynamic Voronoi sites found for centre_species='{centre_species}', "
                                 f"reference_species='{reference_species}', cutoff={cutoff}, n_reference={n_reference}. "
                                 f"Try adjusting these parameters or verify that the specified species exist "
                                 f"in the structure.")
            return sites
        self._site_generators.append(create)
        return self

    def _validate_reference_atom_distances(self) -> None:
        """builders.py:660-694: no two same-species atoms of the reference closer than ``min_atom_distance``."""
        from .distances import all_mic_distances
        ref = self._reference_structure
        lattice = np.array(ref.lattice.matrix)
        species = [s.species_string for s in ref]
        frac = np.array(ref.frac_coords)
        for sp in set(species):
            idx = [i for i, s in enumerate(species) if s == sp]
            if len(idx) < 2:
                continue
            d = all_mic_distances(frac[idx], frac[idx], lattice)
            np.fill_diagonal(d, np.inf)
            nearest = float(np.min(d))
            if nearest < self._min_atom_distance:
                i, j = np.unravel_index(int(np.argmin(d)), d.shape)
                raise ValueError(f"Reference structure has {sp} atoms at indices {idx[i]} and {idx[j]} that are "
                                 f"only {nearest:.3f} apart (threshold: {self._min_atom_distance}). This typically means "
                                 f"the atom coordinate is on a general Wyckoff position instead of the correct special "
                                 f"position, which produces duplicate coordination environments. To disable this check, "
                                 f"call .with_min_atom_distance(0) on the builder.")

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
        if self._reference_structure is not None and self._min_atom_distance > 0:
            self._validate_reference_atom_distances()
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


def create_trajectory_with_polyhedral_sites(structure, reference_structure, mobile_species, centre_species, vertex_species,
                                            cutoff, n_vertices, label=None, align=True, align_species=None,
                                            align_metric="rmsd", align_algorithm="Nelder-Mead",
                                            align_minimizer_options=None, mapping_species=None, align_tolerance=1e-4,
                                            use_reference_centers=True) -> Trajectory:
    """builders.py:878-962."""
    return (TrajectoryBuilder().with_structure(structure).with_reference_structure(reference_structure)
            .with_mobile_species(mobile_species)
            .with_structure_alignment(align=align, align_species=align_species, align_metric=align_metric,
                                      align_algorithm=align_algorithm, align_minimizer_options=align_minimizer_options,
                                      align_tolerance=align_tolerance)
            .with_site_mapping(mapping_species)
            .with_polyhedral_sites(centre_species, vertex_species, cutoff, n_vertices, label, use_reference_centers)
            .build())


def create_trajectory_with_dynamic_voronoi_sites(structure, reference_structure, mobile_species, centre_species,
                                                 reference_species, cutoff, n_reference, label=None, align=True,
                                                 align_species=None, align_metric="rmsd", align_algorithm="Nelder-Mead",
                                                 align_minimizer_options=None, mapping_species=None, align_tolerance=1e-4,
                                                 use_reference_centers=True) -> Trajectory:
    """builders.py:964-1047."""
    return (TrajectoryBuilder().with_structure(structure).with_reference_structure(reference_structure)
            .with_mobile_species(mobile_species)
            .with_structure_alignment(align=align, align_species=align_species, align_metric=align_metric,
                                      align_algorithm=align_algorithm, align_minimizer_options=align_minimizer_options,
                                      align_tolerance=align_tolerance)
            .with_site_mapping(mapping_species)
            .with_dynamic_voronoi_sites(centre_species, reference_species, cutoff, n_reference, label,
                                        use_reference_centers)
            .build())
