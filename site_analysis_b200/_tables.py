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
