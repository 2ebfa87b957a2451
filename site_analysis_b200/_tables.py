"""One-off host-side tables for the GPU engine (numpy / scipy, run once per analysis).

These are the pieces of site *construction* that the reference also computes once on
the host and whose results are integers (orderings, image shifts, face topology), so
the same library calls give the same answers:

* ``distance_rank_table``  -- reference ``site_collection.py:90-111``
* ``face_sharing_neighbours`` -- reference ``polyhedral_site_collection.py:165-191``
* ``full_pbc_unwrap``      -- reference ``pbc_utils.py:65-145`` (first-frame image shifts)
* ``hull_simplices``       -- reference ``polyhedral_site.py:154-155`` (Qhull face topology)
* ``sqrt_threshold``       -- turns ``sqrt(q) <= rcut`` (``spherical_site.py:145-146``) into
  an exactly equivalent ``q <= T`` for the device.
"""
from __future__ import annotations

import itertools

import numpy as np

_SHIFTS27 = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)],
                     dtype=np.int64)


def distance_rank_table(centres: np.ndarray) -> np.ndarray:
    """(S, S-1) int32: for every anchor site the other sites by fractional minimum-image norm.

    Uses the very numpy calls of the reference (``np.linalg.norm`` then the default unstable
    ``np.argsort``) because ties between symmetry-equivalent sites are resolved by the sort
    implementation itself.
    """
    centres = np.asarray(centres, dtype=np.float64)
    n = len(centres)
    table = np.empty((n, max(n - 1, 0)), dtype=np.int32)
    for i in range(n):
        delta = centres - centres[i]
        delta -= np.round(delta)
        order = np.argsort(np.linalg.norm(delta, axis=1))
        table[i] = order[order != i]
    return table


def face_sharing_neighbours(vertex_lists) -> tuple[np.ndarray, np.ndarray]:
    """CSR (offsets, flat) of sites sharing >= 3 vertices, in site-list order."""
    sets = [frozenset(v) for v in vertex_lists]
    by_vertex: dict[int, list[int]] = {}
    for i, vs in enumerate(sets):
        for v in vs:
            by_vertex.setdefault(v, []).append(i)
    offsets = np.zeros(len(sets) + 1, dtype=np.int32)
    flat: list[int] = []
    for i, vs in enumerate(sets):
        cand = sorted({j for v in vs for j in by_vertex[v] if j != i})
        nb = [j for j in cand if len(vs & sets[j]) >= 3]
        flat.extend(nb)
        offsets[i + 1] = len(flat)
    return offsets, np.asarray(flat if flat else [0], dtype=np.int32)


def full_pbc_unwrap(raw: np.ndarray, reference_centre, lattice: np.ndarray) -> np.ndarray:
    """Integer image shifts (n, 3) of the reference's full PBC correction of one site.

    With a reference centre: nearest of the 27 images to the centre, in Cartesian space, through
    the same ``@`` / ``np.linalg.norm`` / ``argmin`` sequence as the reference.  Without one: the
    legacy spread rule (+1 on coordinates below 0.5 along axes whose spread exceeds 0.5).
    """
    raw = np.asarray(raw, dtype=np.float64)
    if reference_centre is None:
        out = raw.copy()
        for d in range(3):
            if out[:, d].max() - out[:, d].min() > 0.5:
                out[out[:, d] < 0.5, d] += 1.0
        return np.round(out - raw).astype(np.int64)
    n = len(raw)
    images = raw[:, None, :] + _SHIFTS27[None, :, :]
    cart = images.reshape(n * 27, 3) @ lattice
    centre_cart = np.asarray(reference_centre, dtype=np.float64) @ lattice
    dist = np.linalg.norm(cart - centre_cart, axis=1).reshape(n, 27)
    return _SHIFTS27[np.argmin(dist, axis=1)]


def unwrap_margins(raw: np.ndarray, reference_centre, lattice: np.ndarray) -> tuple[float, float]:
    """How robust the image choice of :func:`full_pbc_unwrap` is (guard G2 of the engine).

    With a reference centre: (half the smallest gap between the best and the second-best image distance over the
    vertices, Cartesian; inf).  Legacy rule: (inf; the smallest distance, in fractional units, of the rule's inputs from
    its thresholds -- the spread from 0.5 (halved: both ends may move) and every coordinate of a spread axis from 0.5)."""
    raw = np.asarray(raw, dtype=np.float64)
    if reference_centre is None:
        m = np.inf
        for d in range(3):
            spread = raw[:, d].max() - raw[:, d].min()
            m = min(m, abs(spread - 0.5) / 2.0)
            if spread > 0.5:
                m = min(m, float(np.abs(raw[:, d] - 0.5).min()))
        return np.inf, float(m)
    n = len(raw)
    images = raw[:, None, :] + _SHIFTS27[None, :, :]
    cart = images.reshape(n * 27, 3) @ lattice
    centre_cart = np.asarray(reference_centre, dtype=np.float64) @ lattice
    dist = np.sort(np.linalg.norm(cart - centre_cart, axis=1).reshape(n, 27), axis=1)
    return float((dist[:, 1] - dist[:, 0]).min() / 2.0), np.inf


def unwrapped_vertices(raw: np.ndarray, shifts: np.ndarray, non_negative: bool = True) -> np.ndarray:
    """raw + shifts, then the per-axis non-negative offset (pbc_utils.py:102-104 / :218-229)."""
    out = raw + shifts
    if non_negative:
        out = out + np.maximum(0, np.ceil(-out.min(axis=0)))
    return out


def hull_simplices(vertex_coords: np.ndarray, site_label="?") -> np.ndarray:
    """Qhull triangulated faces (n_faces, 3) of a polyhedron's vertices."""
    from scipy.spatial import ConvexHull, QhullError
    try:
        return np.asarray(ConvexHull(vertex_coords).simplices, dtype=np.int32)
    except (QhullError, ValueError) as exc:
        raise RuntimeError(f"Degenerate vertex geometry for polyhedral_site {site_label}") from exc


def sqrt_threshold(rcut: np.ndarray) -> np.ndarray:
    """Largest float64 q with fl(sqrt(q)) <= rcut, elementwise (sqrt is monotone, so
    ``q <= T`` is exactly the reference's ``sqrt(q) <= rcut``)."""
    rcut = np.asarray(rcut, dtype=np.float64)
    q = rcut * rcut
    for _ in range(8):
        up = np.nextafter(q, np.inf)
        grow = np.sqrt(up) <= rcut
        q = np.where(grow, up, q)
    for _ in range(8):
        shrink = np.sqrt(q) > rcut
        q = np.where(shrink, np.nextafter(q, -np.inf), q)
    assert np.all(np.sqrt(q) <= rcut) and np.all(np.sqrt(np.nextafter(q, np.inf)) > rcut)
    return q


# ---------------------------------------------------------------------------------------------
# exact-pruning tables for the tetrahedral kernel (csrc/k_polyhedral_cells.cuh)
# ---------------------------------------------------------------------------------------------

_TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))
_TET_EDGES = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


def _box_tet_separated(V: np.ndarray, h: float) -> np.ndarray:
    """Separating-axis test between the cube [-h, h]^3 and tetrahedra V (n, 4, 3) given relative to
    the cube centre.  True = certainly disjoint.  Slack makes every verdict conservative
    (a touching or nearly touching pair counts as intersecting)."""
    tol = 1e-12
    sep = (V.min(axis=1) > h + tol).any(axis=1) | (V.max(axis=1) < -h - tol).any(axis=1)

    def axis_test(ax):
        proj = np.einsum("nvk,nk->nv", V, ax)
        r = h * np.abs(ax).sum(axis=1)
        slack = tol * (1.0 + np.abs(proj).max(axis=1) + r)
        return (proj.min(axis=1) > r + slack) | (proj.max(axis=1) < -r - slack)

    for a, b, c in _TET_FACES:
        sep |= axis_test(np.cross(V[:, b] - V[:, a], V[:, c] - V[:, a]))
    eye = np.eye(3)
    for a, b in _TET_EDGES:
        e = V[:, b] - V[:, a]
        for i in range(3):
            sep |= axis_test(np.cross(e, eye[i][None, :]))
    return sep


def tet_cell_lists(anchor_vertices: np.ndarray, delta: float, G: int, eps: float = 1e-9):
    """CSR cell lists over the unit cube: for each of the G^3 cells the tetrahedra whose hull,
    with every vertex free to move by +-delta per axis (=> hull (+) [-delta, delta]^3), can reach
    the cell for some periodic image.  ``anchor_vertices`` (S, 4, 3) are unwrapped anchor
    positions (any integer offset per site).  Returns (offsets int32 [G^3 + 1], entries uint16)."""
    P = np.asarray(anchor_vertices, dtype=np.float64)
    S = len(P)
    grow = delta + eps
    lo = np.floor((P.min(axis=1) - grow) * G).astype(np.int64)
    hi = np.floor((P.max(axis=1) + grow) * G).astype(np.int64)
    span = np.minimum(hi - lo + 1, G)                            # cells touched per axis (periodic)
    nmax = span.max(axis=0)
    off = np.stack(np.meshgrid(*[np.arange(n) for n in nmax], indexing="ij"), axis=-1).reshape(-1, 3)   # (K, 3)
    valid = (off[None, :, :] < span[:, None, :]).all(axis=2)     # (S, K)
    s_idx, k_idx = np.nonzero(valid)
    cell_unwrapped = lo[s_idx] + off[k_idx]                      # (n, 3)
    centre = (cell_unwrapped + 0.5) / G
    keep = np.ones(len(s_idx), dtype=bool)
    h = 0.5 / G + grow
    step = 2_000_000
    for a in range(0, len(s_idx), step):
        sl = slice(a, a + step)
        keep[sl] = ~_box_tet_separated(P[s_idx[sl]] - centre[sl][:, None, :], h)
    s_idx, cell_unwrapped = s_idx[keep], cell_unwrapped[keep]
    c = np.mod(cell_unwrapped, G)
    cell = (c[:, 0] * G + c[:, 1]) * G + c[:, 2]
    pair = np.unique(cell * S + s_idx)                           # dedupe, sorted by (cell, site)
    cell, site = pair // S, pair % S
    offsets = np.zeros(G ** 3 + 1, dtype=np.int32)
    np.add.at(offsets, cell + 1, 1)
    offsets = np.cumsum(offsets).astype(np.int32)
    return offsets, site.astype(np.uint16)


def tet_records(slot_shift: np.ndarray, slot_fw: np.ndarray, simplices, kmin: int, m3: int) -> np.ndarray:
    """(S, 20) uint16 site records of the tetrahedral kernels: [0..11] BYTE offsets into the staged
    coordinate table of the 4 vertices in slot order, offset(v, d) = 8 * ((slot_shift[s, v, d] - kmin) * m3
    + 3 * slot_fw[s, v] + d); [12..17] twelve bytes = local vertex number of (face j, corner k) in Qhull's
    own order; [18..19] padding."""
    S = len(slot_fw)
    off = 8 * ((slot_shift[:, :4, :].astype(np.int64) - kmin) * m3 + 3 * slot_fw[:, :4, None].astype(np.int64) + np.arange(3))
    assert off.min() >= 0 and off.max() < 65536
    rec = np.zeros((S, 20), dtype=np.uint16)
    rec[:, :12] = off.reshape(S, 12)
    simp = np.stack([np.asarray(f, dtype=np.uint8) for f in simplices]).reshape(S, 12)     # (S, 4 faces x 3 corners)
    # the FP32 pre-test (csrc/k_polyhedral_cells.cuh) assumes the hull of 4 points: its four faces are
    # exactly the four vertex triples, in any order and orientation
    want = {(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)}
    for s in range(S):
        got = {tuple(sorted(int(v) for v in simp[s, 3 * j:3 * j + 3])) for j in range(4)}
        if got != want:
            raise ValueError(f"site {s}: face list {simp[s].reshape(4, 3).tolist()} is not the hull of a tetrahedron")
    rec[:, 12:18] = np.ascontiguousarray(simp).view(np.uint16)
    return rec


# ---------------------------------------------------------------------------------------------
# exact-pruning tables for the nearest-centre (Voronoi) kernel (csrc/k_nearest_cells.cuh)
# ---------------------------------------------------------------------------------------------

def _mic_cartesian(d_frac: np.ndarray, lattice: np.ndarray) -> np.ndarray:
    """Minimum over the 27 images around the reduced separation of |D . L| (distances.py:175-189)."""
    d = d_frac - np.round(d_frac)
    best = np.full(d.shape[:-1], np.inf)
    for s in _SHIFTS27:
        best = np.minimum(best, np.linalg.norm((d + s) @ lattice, axis=-1))
    return best


def nearest_cell_lists(centres: np.ndarray, lattice0: np.ndarray, G: int, rho: float = 1.06, move: float = 0.0):
    """For each of the G^3 cells of the unit cube: the sites that can be the nearest centre of some
    point of the cell, with head-room ``rho`` for lattice strain, and the exclusion radius of the list.

    With R = circumradius of a cell (metric of ``lattice0``), p = cell centre: every point x of the cell
    has d(x, s) in [d(p, s) - R, d(p, s) + R].  D = min_s (d(p, s) + R) bounds the nearest distance from
    above; sites with d(p, s) - R > rho * D are left out, and ``excl[c] = rho * D`` is a lower bound on
    the distance (lattice0 metric) from any point of the cell to any site NOT in its list.  The kernel
    accepts the best listed site only if it is provably closer than that bound in the frame's own
    metric, and otherwise searches all sites, so the lists never decide a result on their own.

    ``move`` (dynamic Voronoi): every centre may be up to ``move`` away from the given (anchor)
    position; then d(x, s_t) is within +-move of d(x, anchor_s), D gains +move, the keep test -move.

    Returns (offsets int32 [G^3 + 1], entries uint16 ascending per cell, excl float64 [G^3])."""
    from scipy.spatial import cKDTree
    centres = np.asarray(centres, dtype=np.float64) % 1.0
    S = len(centres)
    idx = np.stack(np.meshgrid(*[np.arange(G)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)
    p = (idx + 0.5) / G
    corners = np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], dtype=float) * (0.5 / G)
    R = float(np.linalg.norm(corners @ lattice0, axis=1).max()) * (1.0 + 1e-12)
    a0, a1, a2 = lattice0
    vol = abs(np.dot(a0, np.cross(a1, a2)))
    h = vol / np.array([np.linalg.norm(np.cross(a1, a2)), np.linalg.norm(np.cross(a2, a0)), np.linalg.norm(np.cross(a0, a1))])
    p_cart = p @ lattice0

    def image_tree(reach, margin):
        """k-d tree over the images c + n, n in {-reach..reach}^3, that lie within ``margin`` (Cartesian) of the cell."""
        n = np.array(list(itertools.product(range(-reach, reach + 1), repeat=3)), dtype=float)
        pts = (centres[None, :, :] + n[:, None, :]).reshape(-1, 3)
        site = np.tile(np.arange(S), len(n))
        m = margin / h                                            # fractional margin per axis
        ok = np.all((pts >= -m) & (pts <= 1.0 + m), axis=1)
        return cKDTree(pts[ok] @ lattice0), site[ok]

    # D: upper bound of the nearest-centre distance anywhere in the cell (27 images around the cell centre's own
    # reduced separations; only the efficiency of the lists depends on it, not their validity)
    tree27, _ = image_tree(1, float(np.linalg.norm(lattice0, axis=1).sum()))
    D = (tree27.query(p_cart)[0] + R) * (1.0 + 1e-12) + move
    excl = rho * D
    # keep every site with min_n |p - c - n| - R - move <= rho D over n in {-2..2}^3, a superset of the 27 images
    # the kernels take around the reduced separation of any wrapped point (distances.py:175-189), so a site left
    # out is farther than excl from every point of the cell for the kernels' own distance as well
    radius = (excl + R + move) * (1.0 + 1e-9)
    tree125, site125 = image_tree(2, float(radius.max()) * 1.01 + 1e-9)
    hits = tree125.query_ball_point(p_cart, radius)
    offsets = np.zeros(G ** 3 + 1, dtype=np.int64)
    entries = []
    for c, hit in enumerate(hits):
        u = np.unique(site125[hit])
        offsets[c + 1] = len(u)
        entries.append(u)
    offsets = np.cumsum(offsets).astype(np.int32)
    return offsets, np.concatenate(entries).astype(np.uint16), excl


def sphere_cell_lists(centres: np.ndarray, rcut: np.ndarray, lattice0: np.ndarray, G: int, rho: float = 1.06):
    """For each of the G^3 cells of the unit cube: the spherical sites that can contain some point of the cell.

    Site s is listed for cell c (centre p, circumradius R in the metric of ``lattice0``) when
    min over n in {-2..2}^3 of |p - c_s - n| <= rho * rcut_s + R.  Hence for every point x of the cell and every
    unlisted site, |x - c_s - n| > rho * rcut_s for all those images -- a superset of the 27 images the reference takes
    around the reduced separation of a wrapped point (distances.py:27-66).  ``rho`` is head-room for lattice strain:
    the kernel uses the lists only in frames with (1 - ||L0^-1 L_t - I||_F) * rho >= 1.

    Returns (offsets int32 [G^3 + 1], entries uint16 ascending per cell)."""
    from scipy.spatial import cKDTree
    centres = np.asarray(centres, dtype=np.float64) % 1.0
    rcut = np.broadcast_to(np.asarray(rcut, dtype=np.float64), (len(centres),))
    S = len(centres)
    idx = np.stack(np.meshgrid(*[np.arange(G)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)
    p = (idx + 0.5) / G
    corners = np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], dtype=float) * (0.5 / G)
    R = float(np.linalg.norm(corners @ lattice0, axis=1).max()) * (1.0 + 1e-12)
    a0, a1, a2 = lattice0
    vol = abs(np.dot(a0, np.cross(a1, a2)))
    h = vol / np.array([np.linalg.norm(np.cross(a1, a2)), np.linalg.norm(np.cross(a2, a0)), np.linalg.norm(np.cross(a0, a1))])
    reach = (rho * rcut + R) * (1.0 + 1e-9) + 1e-12
    n = np.array(list(itertools.product(range(-2, 3), repeat=3)), dtype=float)
    pts = (centres[None, :, :] + n[:, None, :]).reshape(-1, 3)
    site = np.tile(np.arange(S), len(n))
    m = float(reach.max()) * 1.01 / h
    ok = np.all((pts >= -m) & (pts <= 1.0 + m), axis=1)
    pts, site = pts[ok], site[ok]
    tree = cKDTree(p @ lattice0)
    hits = tree.query_ball_point(pts @ lattice0, reach[site])          # per site image: the cells it can reach
    cell = np.concatenate([np.asarray(hh, dtype=np.int64) for hh in hits]) if len(hits) else np.zeros(0, np.int64)
    who = np.repeat(site, [len(hh) for hh in hits])
    key = np.unique(cell * S + who)                                     # ascending cell, then ascending site
    cell_u, site_u = key // S, key % S
    offsets = np.zeros(G ** 3 + 1, dtype=np.int64)
    np.add.at(offsets, cell_u + 1, 1)
    return np.cumsum(offsets).astype(np.int32), site_u.astype(np.uint16)
