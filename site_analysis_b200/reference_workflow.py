"""Deriving polyhedral / dynamic-Voronoi sites from a reference structure.

Same workflow, arguments, results and ``ValueError`` conditions as the reference package's
``site_analysis/reference_workflow`` (``ReferenceBasedSites`` reference_based_sites.py:36-415, ``StructureAligner``
structure_aligner.py:25-364, ``IndexMapper`` index_mapper.py:26-209, ``CoordinationEnvironmentFinder``
coord_finder.py:28-106, ``SiteFactory`` site_factory.py:28-257, and the helpers ``get_coordination_indices`` /
``calculate_species_distances`` of tools.py:32-111, 390-452), written around arrays:

1. *align*: find the translation of the reference that minimises the distances between same-species atoms
   (Nelder-Mead or differential evolution on the rmsd / largest distance);
2. *coordination search*: centre atoms of the reference with exactly ``n`` coordinating atoms inside a cut-off, nearest
   first;
3. *index mapping*: every coordinating reference atom -> the nearest atom (optionally of given species) of the target;
4. *site construction*: ``PolyhedralSite`` / ``DynamicVoronoiSite`` objects, polyhedra primed with the target structure.

Every distance in steps 1-3 is a minimum-image distance matrix computed ON THE GPU by the stand-alone operator
``site_analysis_b200.distances.all_mic_distances`` (``sab_mic_distances``), bit-identical to the reference's numba
kernel -- so the optimiser sees the same objective values and all integer decisions (who is inside the cut-off, which
atom is nearest) are the reference's.  A structure is anything with ``frac_coords``, ``lattice.matrix`` and iteration
over sites carrying ``species_string`` (pymatgen ``Structure`` or :class:`~site_analysis_b200.structure.SimpleStructure`).
"""
from __future__ import annotations

from typing import Any, Callable

import numpy as np

from . import distances
from .dynamic_voronoi_site import DynamicVoronoiSite
from .polyhedral_site import PolyhedralSite
from .structure import SimpleStructure


def _species_list(structure) -> list[str]:
    return [site.species_string for site in structure]


def indices_for_species(species: list[str], wanted: str) -> list[int]:
    return [i for i, s in enumerate(species) if s == wanted]


def get_coordination_indices(frac_coords, lattice_matrix, species, centre_species, coordination_species, cutoff,
                             n_coord) -> dict[int, list[int]]:
    """Centre atoms with exactly ``n_coord`` coordinating atoms within ``cutoff``, nearest first (tools.py:32-111)."""
    if len(species) != len(frac_coords):
        raise ValueError(f"species length ({len(species)}) does not match frac_coords rows ({len(frac_coords)})")
    wanted = {coordination_species} if isinstance(coordination_species, str) else set(coordination_species)
    centres = indices_for_species(species, centre_species)
    if not centres:
        raise ValueError(f"No atoms of species '{centre_species}' found in species list")
    shell = [i for i, s in enumerate(species) if s in wanted]
    if isinstance(n_coord, int):
        required = [n_coord] * len(centres)
    else:
        if len(n_coord) != len(centres):
            raise ValueError(f"Length of n_coord list ({len(n_coord)}) does not match "
                             f"number of {centre_species} atoms ({len(centres)})")
        required = list(n_coord)
    frac_coords = np.asarray(frac_coords, dtype=np.float64)
    if shell:
        d = distances.all_mic_distances(frac_coords[centres], frac_coords[shell], lattice_matrix)
    else:
        d = np.empty((len(centres), 0))
    shell_arr = np.array(shell, dtype=np.int64)
    found: dict[int, list[int]] = {}
    for row, (centre, need) in enumerate(zip(centres, required)):
        inside = np.flatnonzero((d[row] <= cutoff) & (shell_arr != centre)) if shell else np.empty(0, dtype=np.int64)
        if len(inside) == need:
            order = np.argsort(d[row][inside], kind="stable")        # ties keep atom order, as list.sort does
            found[centre] = [int(shell_arr[j]) for j in inside[order]]
    return found


def calculate_species_distances(frac_coords1, frac_coords2, lattice_matrix, species1, species2, species=None):
    """Per species, the distance from every atom of set 1 to the nearest same-species atom of set 2 (tools.py:390-452)."""
    if len(species1) != len(frac_coords1):
        raise ValueError(f"species1 length ({len(species1)}) does not match frac_coords1 rows ({len(frac_coords1)})")
    if len(species2) != len(frac_coords2):
        raise ValueError(f"species2 length ({len(species2)}) does not match frac_coords2 rows ({len(frac_coords2)})")
    if species is None:
        species = sorted(set(species1) & set(species2))
    per_species: dict[str, list[float]] = {}
    flat: list[float] = []
    for sp in species:
        a, b = indices_for_species(species1, sp), indices_for_species(species2, sp)
        if not a or not b:
            continue
        nearest = np.min(distances.all_mic_distances(frac_coords1[a], frac_coords2[b], lattice_matrix), axis=1).tolist()
        per_species[sp] = nearest
        flat.extend(nearest)
    return per_species, flat


class StructureAligner:
    """Translation that superimposes a reference on a target (structure_aligner.py:25-364)."""

    def align(self, reference, target, species: list[str] | None = None, metric: str = "rmsd", tolerance: float = 1e-4,
              algorithm: str = "Nelder-Mead", minimizer_options: dict[str, Any] | None = None):
        ref_frac = np.asarray(reference.frac_coords, dtype=np.float64)
        tgt_frac = np.asarray(target.frac_coords, dtype=np.float64)
        lattice = np.asarray(reference.lattice.matrix, dtype=np.float64)
        ref_species, tgt_species = _species_list(reference), _species_list(target)
        use = self._validate_structures(ref_species, tgt_species, species)

        def objective(t: np.ndarray) -> float:
            moved = (ref_frac + t % 1.0) % 1.0
            _, dist = calculate_species_distances(moved, tgt_frac, lattice, ref_species, tgt_species, species=use)
            if not dist:
                return float("inf")
            if metric == "rmsd":
                return float(np.sqrt(np.mean(np.array(dist) ** 2)))
            if metric == "max_dist":
                return float(np.max(dist))
            raise ValueError(f"Unknown metric: {metric}")

        runners = {"Nelder-Mead": self._run_nelder_mead, "differential_evolution": self._run_differential_evolution}
        if algorithm not in runners:
            raise ValueError(f"Unsupported algorithm: {algorithm}. Supported algorithms: {', '.join(runners)}")
        translation = runners[algorithm](objective, tolerance, minimizer_options)
        aligned_frac = (ref_frac + translation) % 1.0
        aligned = SimpleStructure(lattice, ref_species, aligned_frac)
        _, dist = calculate_species_distances(aligned_frac, tgt_frac, lattice, ref_species, tgt_species, species=use)
        inf = float("inf")
        metrics = {"rmsd": float(np.sqrt(np.mean(np.array(dist) ** 2))) if dist else inf,
                   "max_dist": float(np.max(dist)) if dist else inf,
                   "mean_dist": float(np.mean(dist)) if dist else inf}
        return aligned, translation, metrics

    @staticmethod
    def _validate_structures(ref_species, tgt_species, species=None) -> list[str]:
        def count(items):
            c: dict[str, int] = {}
            for s in items:
                c[s] = c.get(s, 0) + 1
            return c
        ref_n, tgt_n = count(ref_species), count(tgt_species)
        if species is None:
            if ref_n != tgt_n:
                raise ValueError(f"Structures have different compositions: {ref_n} vs {tgt_n}")
            use = sorted(ref_n)
        else:
            use = species
        for sp in use:
            if ref_n.get(sp, 0) == 0:
                raise ValueError(f"Species {sp} not found in reference structure")
            if tgt_n.get(sp, 0) == 0:
                raise ValueError(f"Species {sp} not found in target structure")
            if ref_n[sp] != tgt_n[sp]:
                raise ValueError(f"Different number of {sp} atoms: {ref_n[sp]} in reference vs {tgt_n[sp]} in target")
        return use

    @staticmethod
    def _run_nelder_mead(objective: Callable[[np.ndarray], float], tolerance: float, minimizer_options=None) -> np.ndarray:
        from scipy.optimize import minimize
        options: dict[str, Any] = {"xatol": tolerance, "fatol": tolerance}
        options.update(minimizer_options or {})
        result = minimize(objective, x0=np.array([0, 0, 0]), method="Nelder-Mead", options=options)
        if not result.success:
            raise ValueError(f"Optimization failed: {result.message}")
        return np.array(result.x) % 1.0

    @staticmethod
    def _run_differential_evolution(objective, tolerance: float, minimizer_options=None) -> np.ndarray:
        from scipy.optimize import differential_evolution
        options: dict[str, Any] = {"tol": tolerance, "popsize": 15, "maxiter": 1000, "strategy": "best1bin",
                                   "updating": "immediate", "workers": 1}
        bounds = [(0, 1), (0, 1), (0, 1)]
        if minimizer_options:
            options.update(minimizer_options)
            if "bounds" in minimizer_options:
                bounds = options.pop("bounds")
        result = differential_evolution(objective, bounds=bounds, **options)
        if not result.success:
            raise ValueError(f"Differential evolution optimization failed: {result.message}")
        return np.array(result.x) % 1.0


class IndexMapper:
    """Reference atom index -> nearest target atom index, one to one (index_mapper.py:26-209)."""

    def map_coordinating_atoms(self, ref_frac_coords, target_frac_coords, lattice_matrix, ref_coordinating,
                               target_species=None, species_filter=None) -> list[list[int]]:
        if target_species is not None and len(target_species) != len(target_frac_coords):
            raise ValueError(f"target_species length ({len(target_species)}) does not match "
                             f"target_frac_coords rows ({len(target_frac_coords)})")
        wanted = sorted({i for env in ref_coordinating for i in env})
        table = self._nearest_target_atoms(np.asarray(ref_frac_coords), np.asarray(target_frac_coords), lattice_matrix,
                                           wanted, target_species, species_filter)
        return [[int(table[i]) for i in env] for env in ref_coordinating]

    @staticmethod
    def _nearest_target_atoms(ref_frac, tgt_frac, lattice, ref_indices, target_species, species_filter) -> dict[int, int]:
        if not ref_indices:
            return {}
        if isinstance(species_filter, str):
            species_filter = [species_filter]
        if species_filter is not None:
            if target_species is None:
                raise ValueError("target_species must be provided when using species_filter")
            candidates = np.flatnonzero(np.array([s in species_filter for s in target_species]))
            if len(candidates) == 0:
                raise ValueError(f"No atoms of species {species_filter} found in target structure")
        else:
            candidates = np.arange(len(tgt_frac))
        d = distances.all_mic_distances(ref_frac[ref_indices], tgt_frac[candidates], lattice)
        chosen = candidates[np.argmin(d, axis=1)]
        if len(np.unique(chosen)) != len(chosen):
            seen, twice = set(), []
            for idx in chosen.tolist():
                if idx in seen:
                    twice.append(int(idx))
                seen.add(idx)
            raise ValueError(f"1:1 mapping violation: Multiple reference atoms map to the same target atom(s) at indices {twice}")
        return {ref: int(t) for ref, t in zip(ref_indices, chosen)}


class CoordinationEnvironmentFinder:
    """Coordination environments of one structure (coord_finder.py:28-106)."""

    def __init__(self, structure) -> None:
        self._frac_coords = np.asarray(structure.frac_coords, dtype=np.float64)
        self._lattice_matrix = np.asarray(structure.lattice.matrix, dtype=np.float64)
        self._species = _species_list(structure)

    def find_environments(self, center_species: str, coordination_species, n_coord: int, cutoff: float) -> dict[int, list[int]]:
        present = set(self._species)
        if center_species not in present:
            raise ValueError(f"Center species '{center_species}' not found in structure")
        if isinstance(coordination_species, str):
            coordination_species = [coordination_species]
        for sp in coordination_species:
            if sp not in present:
                raise ValueError(f"Coordinating species '{sp}' not found in structure")
        return get_coordination_indices(self._frac_coords, self._lattice_matrix, self._species, center_species,
                                        coordination_species, cutoff, n_coord)


class SiteFactory:
    """Site objects from lists of atom indices of the target structure (site_factory.py:28-257)."""

    def __init__(self, structure) -> None:
        self.structure = structure
        self._frac_coords = np.asarray(structure.frac_coords, dtype=np.float64)
        self._lattice_matrix = np.asarray(structure.lattice.matrix, dtype=np.float64)

    def _check(self, environments, reference_centers, label, labels) -> None:
        if not isinstance(environments, list):
            raise ValueError(f"Environments must be a list, got {type(environments)}")
        for i, env in enumerate(environments):
            if not isinstance(env, list):
                raise ValueError(f"Environment {i} must be a list, got {type(env)}")
            for j, idx in enumerate(env):
                if not isinstance(idx, int):
                    raise ValueError(f"Index {j} in environment {i} must be an integer, got {type(idx)}")
                if idx < 0 or idx >= len(self.structure):
                    raise ValueError(f"Index {idx} in environment {i} is out of range "
                                     f"(structure has {len(self.structure)} atoms)")
        if label is not None and labels is not None:
            raise ValueError("Cannot provide both 'label' and 'labels' arguments")
        if labels is not None and len(labels) != len(environments):
            raise ValueError(f"Number of labels ({len(labels)}) does not match number of environments ({len(environments)})")

    @staticmethod
    def _check_centres(environments, reference_centers) -> None:
        if reference_centers is not None and len(reference_centers) != len(environments):
            raise ValueError(f"Length of reference_centers ({len(reference_centers)}) must match "
                             f"length of environments ({len(environments)})")

    def create_polyhedral_sites(self, environments, reference_centers=None, label=None, labels=None) -> list[PolyhedralSite]:
        self._check(environments, reference_centers, label, labels)
        for i, env in enumerate(environments):
            if len(env) < 4:
                raise ValueError(f"Environment {i} has {len(env)} vertices, but PolyhedralSite "
                                 f"requires at least 4 vertices to form a polyhedron.")
        self._check_centres(environments, reference_centers)
        sites = []
        for i, env in enumerate(environments):
            site = PolyhedralSite(vertex_indices=env, label=label if label is not None else (labels[i] if labels is not None else None),
                                  reference_center=None if reference_centers is None else reference_centers[i])
            site.assign_vertex_coords(self._frac_coords, self._lattice_matrix)      # primes the PBC image shifts
            sites.append(site)
        return sites

    def create_dynamic_voronoi_sites(self, environments, reference_centers=None, label=None, labels=None) -> list[DynamicVoronoiSite]:
        self._check(environments, reference_centers, label, labels)
        self._check_centres(environments, reference_centers)
        return [DynamicVoronoiSite(reference_indices=env,
                                   label=label if label is not None else (labels[i] if labels is not None else None),
                                   reference_center=None if reference_centers is None else reference_centers[i])
                for i, env in enumerate(environments)]


class ReferenceBasedSites:
    """The whole workflow for one (reference, target) pair (reference_based_sites.py:36-415)."""

    def __init__(self, reference_structure, target_structure, align: bool = True, align_species: list[str] | None = None,
                 align_metric: str = "rmsd", align_algorithm: str = "Nelder-Mead",
                 align_minimizer_options: dict[str, Any] | None = None, align_tolerance: float = 1e-4) -> None:
        self.reference_structure = reference_structure
        self.target_structure = target_structure
        self._ref_frac_coords = np.asarray(reference_structure.frac_coords, dtype=np.float64)
        self._ref_lattice_matrix = np.asarray(reference_structure.lattice.matrix, dtype=np.float64)
        self._ref_species = _species_list(reference_structure)
        self._target_frac_coords = np.asarray(target_structure.frac_coords, dtype=np.float64)
        self._target_lattice_matrix = np.asarray(target_structure.lattice.matrix, dtype=np.float64)
        self._target_species = _species_list(target_structure)
        self.aligned_structure = None
        self.translation_vector: np.ndarray | None = None
        self.alignment_metrics: dict[str, float] | None = None
        self._aligned_frac_coords: np.ndarray | None = None
        if align:
            try:
                aligned, translation, metrics = StructureAligner().align(
                    reference_structure, target_structure, species=align_species, metric=align_metric,
                    tolerance=align_tolerance, algorithm=align_algorithm, minimizer_options=align_minimizer_options)
            except ValueError as e:
                raise ValueError(f"Structure alignment failed: {str(e)}") from e
            self.aligned_structure, self.translation_vector, self.alignment_metrics = aligned, translation, metrics
            self._aligned_frac_coords = np.asarray(aligned.frac_coords)
        self._coord_finder: CoordinationEnvironmentFinder | None = None
        self._index_mapper: IndexMapper | None = None
        self._site_factory: SiteFactory | None = None

    @property
    def _effective_ref_coords(self) -> np.ndarray:
        return self._ref_frac_coords if self._aligned_frac_coords is None else self._aligned_frac_coords

    def _environments(self, center_species, coordination_species, cutoff, n_coord) -> dict[int, list[int]]:
        try:
            if self._coord_finder is None:
                self._coord_finder = CoordinationEnvironmentFinder(self.reference_structure)
            found = self._coord_finder.find_environments(center_species=center_species,
                                                         coordination_species=coordination_species, n_coord=n_coord, cutoff=cutoff)
        except ValueError as e:
            raise ValueError(f"Failed to find coordination environments for {center_species} centers "
                             f"and {coordination_species} coordinating atoms: {str(e)}") from e
        for centre, env in found.items():
            if len(env) != len(set(env)):
                twice = sorted({i for i in env if env.count(i) > 1})
                raise ValueError(f"Environment for center atom {centre} contains duplicate atom indices {twice}. "
                                 f"This typically occurs in small unit cells where the same atom "
                                 f"appears as a neighbor in multiple periodic images. "
                                 f"Please use a larger supercell for your analysis.")
        return found

    def _map_environments(self, ref_environments: list[list[int]], target_species=None) -> list[list[int]]:
        if not ref_environments:
            return []
        try:
            if self._index_mapper is None:
                self._index_mapper = IndexMapper()
            return self._index_mapper.map_coordinating_atoms(
                ref_frac_coords=self._effective_ref_coords, target_frac_coords=self._target_frac_coords,
                lattice_matrix=self._ref_lattice_matrix, ref_coordinating=ref_environments,
                target_species=self._target_species, species_filter=target_species)
        except ValueError as e:
            which = f" for {target_species} species" if target_species else ""
            raise ValueError(f"Failed to map coordination environments{which}: {str(e)}") from e

    def _build(self, make: str, center_species, shell_species, cutoff, n, label, labels, target_species, use_reference_centers):
        found = self._environments(center_species, shell_species, cutoff, n)
        centres = None
        if use_reference_centers:
            coords = self._effective_ref_coords
            centres = [coords[i].copy() for i in found]
        mapped = self._map_environments(list(found.values()), target_species)
        if self._site_factory is None:
            self._site_factory = SiteFactory(self.target_structure)
        return getattr(self._site_factory, make)(mapped, reference_centers=centres, label=label, labels=labels)

    def create_polyhedral_sites(self, center_species: str, vertex_species, cutoff: float, n_vertices: int,
                                label: str | None = None, labels: list[str] | None = None, target_species=None,
                                use_reference_centers: bool = True) -> list[PolyhedralSite]:
        return self._build("create_polyhedral_sites", center_species, vertex_species, cutoff, n_vertices, label, labels,
                           target_species, use_reference_centers)

    def create_dynamic_voronoi_sites(self, center_species: str, reference_species, cutoff: float, n_reference: int,
                                     label: str | None = None, labels: list[str] | None = None, target_species=None,
                                     use_reference_centers: bool = True) -> list[DynamicVoronoiSite]:
        return self._build("create_dynamic_voronoi_sites", center_species, reference_species, cutoff, n_reference, label,
                           labels, target_species, use_reference_centers)
