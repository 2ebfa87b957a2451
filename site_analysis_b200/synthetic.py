"""Synthetic workloads: argyrodite (BASELINE.json configs 2/3/5) and LLZO garnet (config 4); SURVEY.md section 8(d).

Geometry: the F-4 3m Li6PS5Cl reference of the reference's own benchmark scripts
(tests/benchmark_all_site_types.py:29-46 -- P on 4b, anions on 4a/4c/16e, and the five Li site
types 48h/48h/4/16e/48h) in an n x n x n supercell; tetrahedral sites = the four nearest anions
of every Li-site position (cutoff 3.0 A), exactly what the reference workflow derives for this
structure.  Dynamics: framework atoms vibrate around their ideal positions (Gaussian, sigma in
Angstrom), mobile ions hop between neighbouring type-5 / type-2 sites and jitter around the
site centre; coordinates are stored wrapped into [0, 1) so atoms near a cell face flip 0 <-> 1.

Garnet (:func:`garnet_geometry`): cubic Li7La3Zr2O12, Ia-3d, a = 12.97 A, 192 atoms per cell (56 Li, 24 La on 24c,
16 Zr on 16a, 96 O on 96h); dynamic-Voronoi sites from the oxygen coordination of the Li positions: 24d tetrahedra
(4 reference atoms) and 48g octahedra (6), i.e. two centre groups
(reference dynamic_voronoi_site_collection.py:151-164).

Trajectories are generated with torch on the requested device from a seeded generator, so the
same call gives the same data on every run and a host copy of any frame range is exactly what
the kernels saw.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

_SITE_TYPES = [("1", (0.9, 0.9, 0.6)), ("2", (0.77, 0.585, 0.585)), ("3", (0.25, 0.25, 0.25)),
               ("4", (0.15, 0.15, 0.15)), ("5", (0.0, 0.183, 0.183))]
_FRAMEWORK = [("P", (0.5, 0.5, 0.5)), ("S", (0.0, 0.0, 0.0)), ("S", (0.75, 0.25, 0.25)),
              ("S", (0.11824, 0.11824, 0.38176))]
A_CELL = 10.155


def _f43m_orbit(xyz, tol=1e-6):
    """Orbit of a position under F-4 3m: signed permutations with an even number of sign changes
    (T_d) times the four F-centring translations."""
    pts: list[np.ndarray] = []
    centring = [(0, 0, 0), (0, .5, .5), (.5, 0, .5), (.5, .5, 0)]
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            if sum(s < 0 for s in signs) % 2:
                continue
            base = np.array([signs[i] * xyz[perm[i]] for i in range(3)], dtype=float)
            for t in centring:
                p = (base + np.array(t)) % 1.0
                p[np.abs(p - 1.0) < tol] = 0.0
                if not any(np.all(np.abs((p - q) - np.round(p - q)) < tol) for q in pts):
                    pts.append(p)
    return np.array(pts)


def _supercell(frac, n):
    shifts = np.array(list(itertools.product(range(n), repeat=3)), dtype=float)
    return ((frac[:, None, :] + shifts[None, :, :]) / n).reshape(-1, 3)


def _mic_cart(d_frac, lattice):
    d = d_frac - np.round(d_frac)
    return np.linalg.norm(d @ lattice, axis=-1)


@dataclass
class ArgyroditeGeometry:
    n: int
    lattice: np.ndarray            # (3, 3) mean cell
    species: list                  # per atom, mobile first ("Li", then framework)
    ideal_framework: np.ndarray    # (N_fw, 3) fractional
    n_mobile: int
    site_centres: np.ndarray       # (S, 3)
    site_labels: list
    tetrahedra: np.ndarray         # (S, 4) structure indices of the vertex anions
    hop_sites: np.ndarray          # positions of the sites mobile ions live on (types 5 and 2)
    hop_neighbours: np.ndarray     # (len(hop_sites), K) indices into hop_sites, padded with self

    @property
    def n_atoms(self) -> int:
        return self.n_mobile + len(self.ideal_framework)

    @property
    def mobile_idx(self) -> np.ndarray:
        return np.arange(self.n_mobile, dtype=np.int32)


def argyrodite_geometry(n: int = 2, hop_radius: float = 2.6) -> ArgyroditeGeometry:
    lattice = np.eye(3) * (A_CELL * n) + 1e-4 * np.array([[0, 1.9, -1.13], [1.9, 0, -0.33], [-1.13, -0.33, 0]])
    fw_species, fw = [], []
    for sp, xyz in _FRAMEWORK:
        orb = _supercell(_f43m_orbit(xyz), n)
        fw.append(orb); fw_species += [sp] * len(orb)
    fw = np.vstack(fw)
    centres, labels = [], []
    for lab, xyz in _SITE_TYPES:
        orb = _supercell(_f43m_orbit(xyz), n)
        centres.append(orb); labels += [lab] * len(orb)
    centres = np.vstack(centres)
    n_mobile = 24 * n ** 3                                   # Li6PS5Cl: 24 Li per conventional cell
    anion = np.array([i for i, s in enumerate(fw_species) if s == "S"])
    tets = np.zeros((len(centres), 4), dtype=np.int64)
    for i, c in enumerate(centres):
        d = _mic_cart(fw[anion] - c, lattice)
        order = np.argsort(d, kind="stable")[:4]
        assert d[order[-1]] < 3.0, "site without four anions inside the 3.0 A cutoff"
        tets[i] = n_mobile + anion[order]                    # framework atoms follow the mobile ions
    hop = np.array([i for i, l in enumerate(labels) if l in ("5", "2")])
    k = 12
    nb = np.tile(np.arange(len(hop))[:, None], (1, k))
    for j, i in enumerate(hop):
        d = _mic_cart(centres[hop] - centres[i], lattice)
        near = [q for q in np.argsort(d, kind="stable")[1:k + 1] if d[q] < hop_radius]
        nb[j, :len(near)] = near
    return ArgyroditeGeometry(n=n, lattice=lattice, species=["Li"] * n_mobile + fw_species, ideal_framework=fw,
                              n_mobile=n_mobile, site_centres=centres, site_labels=labels,
                              tetrahedra=tets, hop_sites=hop, hop_neighbours=nb)


def argyrodite_trajectory(geom: ArgyroditeGeometry, n_frames: int, *, seed: int = 20260105, device="cpu",
                          sigma_framework: float = 0.10, sigma_mobile: float = 0.25, p_hop: float = 0.01,
                          lattice_fluctuation: float = 1e-3, wrapped: bool = True):
    """Returns (frac (F, N, 3) float64, lattice (F, 3, 3) float64) torch tensors on ``device``."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    F, A = n_frames, geom.n_mobile
    a_len = float(np.linalg.norm(geom.lattice[0]))
    frac = torch.empty((F, geom.n_atoms, 3), dtype=torch.float64, device=dev)
    ideal = torch.as_tensor(geom.ideal_framework, device=dev)
    frac[:, A:, :] = ideal[None] + (sigma_framework / a_len) * torch.randn((F, ideal.shape[0], 3), generator=g,
                                                                           dtype=torch.float64, device=dev)
    # hop process: per ion a path over the hop-site graph, advanced whenever a hop fires
    hops = torch.rand((F, A), generator=g, device=dev) < p_hop
    hops[0] = False
    n_hops = torch.cumsum(hops.to(torch.int64), dim=0)                       # (F, A)
    max_hops = int(n_hops.max().item()) + 1
    nb = torch.as_tensor(geom.hop_neighbours, device=dev)
    start = torch.randperm(nb.shape[0], generator=g, device=dev)[:A]
    path = torch.empty((max_hops, A), dtype=torch.int64, device=dev)
    path[0] = start
    choice = torch.randint(0, nb.shape[1], (max_hops, A), generator=g, device=dev)
    for h in range(1, max_hops):
        path[h] = nb[path[h - 1], choice[h]]
    site = torch.gather(path, 0, n_hops)                                      # (F, A) index into hop_sites
    centres = torch.as_tensor(geom.site_centres[geom.hop_sites], device=dev)
    frac[:, :A, :] = centres[site] + (sigma_mobile / a_len) * torch.randn((F, A, 3), generator=g,
                                                                           dtype=torch.float64, device=dev)
    if wrapped:
        frac.remainder_(1.0)
        frac[frac >= 1.0] = 0.0
    scale = 1.0 + lattice_fluctuation * (2.0 * torch.rand((F, 1, 1), generator=g, dtype=torch.float64, device=dev) - 1.0)
    lattice = torch.as_tensor(geom.lattice, device=dev)[None] * scale
    return frac, lattice


def polyhedral_engine_kwargs(geom: ArgyroditeGeometry, frame0: np.ndarray, lattice0: np.ndarray) -> dict:
    """Site tables as the reference's builder would leave them: one tetrahedron per Li-site
    position, reference centre = that position, PBC-shift cache primed from the build structure
    (= frame 0; site_factory.py:248-257)."""
    from . import _tables as tab
    slots = [list(map(int, t)) for t in geom.tetrahedra]
    primed = []
    for t, c in zip(slots, geom.site_centres):
        raw = frame0[t]
        primed.append((tab.full_pbc_unwrap(raw, c, lattice0), raw.copy()))
    return dict(slots=slots, reference_centres=list(geom.site_centres), primed=primed)


# ------------------------------------------------------------------------------------------------ LLZO garnet
A_GARNET = 12.97
_GARNET_FRAMEWORK = [("La", (0.125, 0.0, 0.25)), ("Zr", (0.0, 0.0, 0.0)), ("O", (-0.0318, 0.0546, 0.1498))]
_GARNET_SITES = [("24d", (0.375, 0.0, 0.25), 4), ("48g", (0.125, 0.6838, 0.5662), 6)]


def _ia3d_operations():
    """The 96 operations (R, t) of Ia-3d, closed from the generators listed in International Tables A, No. 230."""
    def op(R, t):
        return np.array(R, dtype=float), np.array(t, dtype=float) % 1.0
    gens = [op([[-1, 0, 0], [0, -1, 0], [0, 0, 1]], [.5, 0, .5]), op([[-1, 0, 0], [0, 1, 0], [0, 0, -1]], [0, .5, .5]),
            op([[0, 0, 1], [1, 0, 0], [0, 1, 0]], [0, 0, 0]), op([[0, 1, 0], [1, 0, 0], [0, 0, -1]], [.75, .25, .25]),
            op(-np.eye(3), [0, 0, 0]), op(np.eye(3), [.5, .5, .5])]

    def key(o):
        return tuple(o[0].astype(int).ravel()), tuple(np.round(o[1] * 8).astype(int) % 8)
    group = [op(np.eye(3), [0, 0, 0])]
    seen = {key(group[0])}
    frontier = list(group)
    while frontier:
        new = []
        for a in frontier:
            for g in gens:
                c = (g[0] @ a[0], (g[0] @ a[1] + g[1]) % 1.0)
                if key(c) not in seen:
                    seen.add(key(c)); new.append(c); group.append(c)
        frontier = new
    assert len(group) == 96
    return group


def _orbit(ops, xyz, tol=1e-6):
    pts: list[np.ndarray] = []
    for R, t in ops:
        p = (R @ np.asarray(xyz, dtype=float) + t) % 1.0
        p[np.abs(p - 1.0) < tol] = 0.0
        if not any(np.all(np.abs((p - q) - np.round(p - q)) < tol) for q in pts):
            pts.append(p)
    return np.array(pts)


@dataclass
class GarnetGeometry(ArgyroditeGeometry):
    reference_slots: list = None   # per site, the structure indices of its 4 (24d) or 6 (48g) oxygen atoms


def garnet_geometry(n: int = 2, hop_radius: float = 2.6) -> GarnetGeometry:
    """LLZO n x n x n: N = 192 n^3 atoms, A = 56 n^3 Li, S = 72 n^3 dynamic-Voronoi sites in two centre groups."""
    ops = _ia3d_operations()
    lattice = np.eye(3) * (A_GARNET * n) + 1e-4 * np.array([[0, 1.7, -0.9], [1.7, 0, 0.4], [-0.9, 0.4, 0]])
    fw_species, fw = [], []
    for sp, xyz in _GARNET_FRAMEWORK:
        orb = _supercell(_orbit(ops, xyz), n)
        fw.append(orb); fw_species += [sp] * len(orb)
    fw = np.vstack(fw)
    n_mobile = 56 * n ** 3
    oxygen = np.array([i for i, s in enumerate(fw_species) if s == "O"])
    centres, labels, slots = [], [], []
    for lab, xyz, n_ref in _GARNET_SITES:                    # all tetrahedra first, then all octahedra: two groups
        for c in _supercell(_orbit(ops, xyz), n):
            d = _mic_cart(fw[oxygen] - c, lattice)
            order = np.argsort(d, kind="stable")
            assert d[order[n_ref - 1]] < 2.6 < d[order[n_ref]], "unexpected oxygen coordination"
            centres.append(c); labels.append(lab)
            slots.append([int(n_mobile + oxygen[j]) for j in order[:n_ref]])
    centres = np.array(centres)
    k = 8
    nb = np.tile(np.arange(len(centres))[:, None], (1, k))
    for i in range(len(centres)):
        d = _mic_cart(centres - centres[i], lattice)
        near = [q for q in np.argsort(d, kind="stable")[1:k + 1] if d[q] < hop_radius]
        nb[i, :len(near)] = near
    return GarnetGeometry(n=n, lattice=lattice, species=["Li"] * n_mobile + fw_species, ideal_framework=fw,
                          n_mobile=n_mobile, site_centres=centres, site_labels=labels, tetrahedra=None,
                          hop_sites=np.arange(len(centres)), hop_neighbours=nb, reference_slots=slots)


synthetic_trajectory = argyrodite_trajectory      # the dynamics only need the fields both geometries share
