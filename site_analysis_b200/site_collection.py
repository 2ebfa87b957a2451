"""GPU-backed site collections (mirror of the reference's ``SiteCollection`` family).

Public surface of reference ``site_collection.py:185-365`` and the four concrete collections
(``polyhedral_site_collection.py``, ``voronoi_site_collection.py``,
``dynamic_voronoi_site_collection.py``, ``spherical_site_collection.py``): same constructor
checks and exceptions, ``analyse_structure(atoms, structure)``,
``assign_site_occupations(atoms, lattice_matrix)``, ``update_occupation``, ``reset``,
``reset_site_occupations``, ``site_by_index``, ``neighbouring_sites``, ``summaries``.

What is different is *where the work happens*: every call hands whole frames to the CUDA
engine (:mod:`site_analysis_b200.engine`), and :meth:`SiteCollection.analyse_trajectory` is the
array-native entry for entire trajectories.  State the reference keeps in Python objects
(``Atom._recent_sites``, ``Atom.trajectory[-1]``, ``Site.transitions``) is pulled from the
objects before a call and pushed back after it, so mixing per-frame and batch calls, or
editing those attributes by hand, behaves as with the reference.
"""
from __future__ import annotations

from collections import Counter
from typing import Sequence

import numpy as np

from .atom import Atom
from .dynamic_voronoi_site import DynamicVoronoiSite
from .engine import AssignmentEngine
from .polyhedral_site import FaceTopology, PolyhedralSite
from .site import Site
from .spherical_site import SphericalSite
from .voronoi_site import VoronoiSite

POINTS_LIMIT = 2_000_000     # atom-frames per batch above which Site.points is filled lazily (Trajectory) instead of eagerly


class SiteCollection:
    """Parent class for collections of sites (reference site_collection.py:185)."""

    _kind = ""
    _site_type: type = Site

    def __init__(self, sites: Sequence[Site]) -> None:
        self.sites = sites
        self._site_lookup: dict[int, Site] = {}
        for site in sites:
            if site.index in self._site_lookup:
                raise ValueError(f"Duplicate site index detected: {site.index}. Site indices must be unique.")
            self._site_lookup[site.index] = site
        self._pos = {s.index: i for i, s in enumerate(sites)}
        self._index_of_pos = np.array([s.index for s in sites], dtype=np.int32)
        self._identity_index = bool(np.array_equal(self._index_of_pos, np.arange(len(sites))))   # Site.index == list position
        self._engine: AssignmentEngine | None = None
        self._engine_key = None
        self._tr_token = None           # (engine, tag) of the state whose Site.transitions are parked in the engine's store
        self._current = None            # (frac (N,3), lattice (3,3)) of the last analysed structure
        self.strict = True
        self.exact_lazy = None          # polyhedral sites: None = automatic, True = always the exact replay kernel (engine.py)
        self.use_cells = True           # exact spatial pruning (False: exhaustive kernels only; for cross-checks)
        self.rank_table = None          # optional override of the distance-rank table (testing)
        self.rank_dense_max = None      # sites up to which the dense S x (S - 1) rank table is built (None: engine default)
        self.engine_options: dict[int, int] = {}   # sab_set_option(key, value) pairs applied to the engine

    # -- reference API ---------------------------------------------------------------------
    def site_by_index(self, index):
        site = self._site_lookup.get(index)
        if site is None:
            raise ValueError(f"No site with index {index} found")
        return site

    def neighbouring_sites(self, site_index):
        raise NotImplementedError("neighbouring_sites should be implemented in the derived class")

    def update_occupation(self, site, atom):
        """site_collection.py:279-310, verbatim behaviour on the Python objects."""
        previous = atom.trajectory[-1] if atom.trajectory else None
        if previous is not None and previous != site.index:
            self.site_by_index(previous).transitions[site.index] += 1
        site.contains_atoms.append(atom.index)
        site.points.append(atom.frac_coords)
        atom.in_site = site.index
        atom.update_recent_site(site.index)

    def reset(self) -> None:
        for site in self.sites:
            site.reset()
        if self._engine is not None:
            self._engine.reset()
        self._current = None

    def reset_site_occupations(self):
        for s in self.sites:
            s.contains_atoms = []

    def sites_contain_points(self, points, all_frac_coords, lattice_matrix) -> bool:
        raise NotImplementedError("sites_contain_points() should be implemented in the derived class")

    def summaries(self, metrics: list[str] | None = None) -> list[dict]:
        return [site.summary(metrics=metrics) for site in self.sites]

    def analyse_structure(self, atoms, structure) -> None:
        """One frame: assign coordinates to atoms, then assign atoms to sites on the GPU."""
        frac = np.ascontiguousarray(structure.frac_coords, dtype=np.float64)
        lattice = np.ascontiguousarray(structure.lattice.matrix, dtype=np.float64)
        self._current = (frac, lattice)
        self._run(atoms, frac[None], lattice, append=False)

    def assign_site_occupations(self, atoms, lattice_matrix) -> None:
        """Assign atoms to sites for the structure last passed to ``analyse_structure`` (or, for
        centre-based sites, for the atoms' current coordinates)."""
        if not atoms:
            self.reset_site_occupations()
            return
        lattice = np.ascontiguousarray(lattice_matrix, dtype=np.float64)
        if self._current is not None:
            frac = self._current[0]
        elif self._kind in ("spherical", "voronoi"):
            n = max(a.index for a in atoms) + 1
            frac = np.zeros((n, 3))
            for a in atoms:
                frac[a.index] = a.frac_coords
        else:
            raise RuntimeError("assign_site_occupations needs the structure: call analyse_structure first")
        self._run(atoms, frac[None], lattice, append=False)

    # -- array-native entry ------------------------------------------------------------------
    def analyse_trajectory(self, atoms, frac: np.ndarray, lattice: np.ndarray) -> np.ndarray:
        """Assign every atom in every frame of ``frac`` (F, N, 3); ``lattice`` (3, 3) or (F, 3, 3).

        Equivalent to the reference's ``for s in structures: trajectory.append_timestep(s)``
        as far as the collection is concerned.  Returns (F, A) int32 of ``Site.index``
        (-1 = ``None``); the caller appends it to the atoms' / sites' trajectories."""
        return self._run(atoms, np.asarray(frac), np.asarray(lattice), append=True)

    # -- engine plumbing -----------------------------------------------------------------------
    def _engine_kwargs(self) -> dict:
        raise NotImplementedError

    def _get_engine(self, atoms, n_atoms) -> AssignmentEngine:
        key = (n_atoms, tuple(a.index for a in atoms))
        if self._engine is None or self._engine_key != key:
            self._engine = AssignmentEngine(self._kind, n_atoms, [a.index for a in atoms], strict=self.strict,
                                            rank_table=self.rank_table, **self._engine_kwargs())
            self._engine.use_cells = self.use_cells
            self._engine.exact_lazy = self.exact_lazy
            if self.rank_dense_max is not None:
                self._engine.rank_dense_max = self.rank_dense_max
            for k, v in self.engine_options.items():
                self._engine.set_option(k, v)
            self._engine_key = key
        return self._engine

    def _to_pos(self, index, what):
        if index is None:
            return -1
        p = self._pos.get(index)
        if p is None:
            raise ValueError(f"No site with index {index} found ({what})")
        return p

    def _pull_state(self, eng: AssignmentEngine, atoms) -> None:
        eng._pending = []
        for a, atom in enumerate(atoms):
            r = atom._recent_sites
            eng.recent[a, 0] = self._to_pos(r[0], "atom history")
            eng.recent[a, 1] = self._to_pos(r[1], "atom history")
            last = atom._last_trajectory_entry()
            eng.prev[a] = last if last < 0 else self._to_pos(last, "atom trajectory")
        # Site.transitions: untouched since this collection parked them in the engine's store (lazy) -> nothing to pull
        token = self._tr_token
        if token is None or token[0] is not eng or any(getattr(s, "_lazy_transitions", None) is None or s._lazy_transitions[1] is not token
                                                        for s in self.sites):
            eng.transitions = [{self._to_pos(k, "transition"): int(v) for k, v in s.transitions.items()} for s in self.sites]

    def _push_state(self, eng: AssignmentEngine, atoms, last_row: np.ndarray, last_frac: np.ndarray) -> None:
        eng.sync_history()
        idx = self._index_of_pos
        wrapped = last_frac[[a.index for a in atoms]] % 1.0
        self.reset_site_occupations()
        for a, atom in enumerate(atoms):
            atom._frac_coords = wrapped[a]
            p = int(last_row[a])
            atom.in_site = None if p < 0 else int(idx[p])
            if p >= 0:
                self.sites[p].contains_atoms.append(atom.index)
            r0, r1 = int(eng.recent[a, 0]), int(eng.recent[a, 1])
            atom._recent_sites = [None if r0 < 0 else int(idx[r0]), None if r1 < 0 else int(idx[r1])]
        # Site.transitions stay in the engine's store; each site builds its Counter on first access (_site_transitions)
        token = self._tr_token = (eng, object())
        for s, site in enumerate(self.sites):
            if isinstance(site, Site):
                site._lazy_transitions = (self, token)
            else:                                               # a site object of the reference package: plain attribute
                t = eng.transitions[s]
                if t or site.transitions:
                    site.transitions = Counter({int(idx[k]): v for k, v in t.items()})

    def _site_transitions(self, site, token) -> Counter:
        """The Counter of one site from the engine's store (called by the lazy ``Site.transitions``)."""
        eng = token[0]
        t = eng.transitions[self._pos[site.index]]
        idx = self._index_of_pos
        return Counter({int(idx[k]): v for k, v in t.items()})

    def _run(self, atoms, frac, lattice, append) -> np.ndarray:
        frac = np.ascontiguousarray(frac, dtype=np.float64)
        if frac.ndim != 3 or frac.shape[2] != 3:
            raise ValueError(f"expected fractional coordinates of shape (F, N, 3), got {frac.shape}")
        F = frac.shape[0]
        if not atoms or F == 0:
            self.reset_site_occupations()
            return np.empty((F, len(atoms)), dtype=np.int32)
        eng = self._get_engine(atoms, frac.shape[1])
        self._pull_state(eng, atoms)
        rows = eng.assign(frac, lattice, append=append)
        self._push_state(eng, atoms, rows[-1], frac[-1])
        self._after_batch(eng, atoms, frac, rows)
        if F * len(atoms) <= POINTS_LIMIT:                      # Site.points (site_collection.py:308)
            mobile = [a.index for a in atoms]
            for f in range(F):
                w = frac[f, mobile] % 1.0
                for a in np.nonzero(rows[f] >= 0)[0]:
                    self.sites[rows[f, a]].points.append(w[a])
        if self._identity_index:                                # the usual case: nothing to translate
            return rows
        return np.where(rows >= 0, self._index_of_pos[np.clip(rows, 0, None)], -1).astype(np.int32)

    def _after_batch(self, eng, atoms, frac, rows) -> None:
        pass


class PriorityAssignmentMixin:
    """Marker for collections whose sites may overlap (reference site_collection.py:56): the
    reference's priority order -- recent sites, learned transitions, distance rank -- is applied
    on the GPU by the resolver kernel (``csrc/k_resolve.cuh``) to atom-frames in several sites."""


class SphericalSiteCollection(PriorityAssignmentMixin, SiteCollection):
    _kind = "spherical"

    def __init__(self, sites) -> None:
        for s in sites:
            if not isinstance(s, SphericalSite) and type(s).__name__ != "SphericalSite":
                raise TypeError(f"Expected SphericalSite, got {type(s).__name__}")
        super().__init__(sites)

    def _engine_kwargs(self):
        return dict(centres=np.array([s.frac_coords for s in self.sites], dtype=np.float64),
                    rcut=np.array([s.rcut for s in self.sites], dtype=np.float64))


class VoronoiSiteCollection(SiteCollection):
    _kind = "voronoi"

    def __init__(self, sites) -> None:
        for s in sites:
            if not isinstance(s, VoronoiSite) and type(s).__name__ != "VoronoiSite":
                raise TypeError
        super().__init__(sites)

    def _engine_kwargs(self):
        return dict(centres=np.array([s.frac_coords for s in self.sites], dtype=np.float64))


class DynamicVoronoiSiteCollection(SiteCollection):
    _kind = "dynamic_voronoi"

    def __init__(self, sites) -> None:
        for s in sites:
            if not isinstance(s, DynamicVoronoiSite) and type(s).__name__ != "DynamicVoronoiSite":
                raise TypeError("All sites must be DynamicVoronoiSite instances")
        super().__init__(sites)

    def _engine_kwargs(self):
        return dict(slots=[list(s.reference_indices) for s in self.sites],
                    reference_centres=[s.reference_center for s in self.sites])

    def _after_batch(self, eng, atoms, frac, rows) -> None:
        import torch
        d = torch.from_numpy(np.ascontiguousarray(frac[-1:])).to(torch.device("cuda", eng.device))
        centres = eng.dynvor_centres(d)[0].cpu().numpy()
        for s, site in enumerate(self.sites):
            site._centre_coords = centres[s]


class PolyhedralSiteCollection(PriorityAssignmentMixin, SiteCollection):
    _kind = "polyhedral"

    def __init__(self, sites) -> None:
        for s in sites:
            if not isinstance(s, PolyhedralSite) and type(s).__name__ != "PolyhedralSite":
                raise TypeError(f"Expected PolyhedralSite, got {type(s).__name__}")
        super().__init__(sites)
        self._neighbours = None

    def neighbouring_sites(self, index: int):
        """Face-sharing neighbours (>= 3 common vertices), polyhedral_site_collection.py:165-191."""
        if self._neighbours is None:
            from ._tables import face_sharing_neighbours
            off, flat = face_sharing_neighbours([s.vertex_indices for s in self.sites])
            self._neighbours = {s.index: [self.sites[j] for j in flat[off[i]:off[i + 1]]]
                                for i, s in enumerate(self.sites)}
        return self._neighbours[index]

    def _engine_kwargs(self):
        primed, simp = [], []
        for s in self.sites:
            sh, raw = getattr(s, "_pbc_image_shifts", None), getattr(s, "_pbc_cached_raw_frac", None)
            primed.append(None if sh is None or raw is None else (np.asarray(sh), np.asarray(raw)))
            cache = getattr(s, "_face_topology_cache", None)
            simp.append(None if cache is None else np.asarray(cache.face_simplices))
        return dict(slots=[list(s.vertex_indices) for s in self.sites],
                    reference_centres=[s.reference_center for s in self.sites], primed=primed, simplices=simp)

    def _after_batch(self, eng, atoms, frac, rows) -> None:
        open_topo = getattr(eng, "_open_topology", None)
        for s, site in enumerate(self.sites):              # topology is permanent once computed
            if open_topo is not None and len(open_topo) > s and open_topo[s]:
                continue                                   # never queried so far: the reference has hulled nothing yet either
            if getattr(site, "_face_topology_cache", None) is None and eng.simplices[s] is not None:
                site._face_topology_cache = FaceTopology(eng.simplices[s])

    def reset(self) -> None:
        super().reset()
        self._engine = None          # primed caches are gone: rebuild from the (reset) sites

    def sites_contain_points(self, points, all_frac_coords, lattice_matrix) -> bool:
        """True if every point lies inside its corresponding site (polyhedral_site_collection.py:113-141): all sites in
        ONE GPU query (``sab_polyhedra_contain``), after each site's vertex coordinates were brought up to date exactly as
        the reference's lazy ``notify_structure_changed`` -> ``contains_point`` sequence does."""
        if len(points) != len(self.sites):
            raise ValueError(f"Expected {len(self.sites)} points (one per site), got {len(points)}")
        from .distances import polyhedra_contain, x_pbc
        vcs, simps = [], []
        for s in self.sites:
            s.notify_structure_changed(all_frac_coords, lattice_matrix)
            vc, simp = s._query_ready()
            vcs.append(vc); simps.append(simp)
        return bool(polyhedra_contain(vcs, simps, [x_pbc(p) for p in points]).all())
