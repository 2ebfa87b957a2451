"""CPU restatement of the reference's post-processing of a finished assignment history.

TEST INFRASTRUCTURE ONLY (imported by ``tests/`` -- never by ``site_analysis_b200``).  Pinned against the
unmodified reference by ``tests/golden/folds_*.npz`` (``oracle/gen_golden_folds.py``).

Plain Python loops over nested lists, following the reference line by line:

* ``site_trajectories``  -- what ``Trajectory.append_timestep`` leaves in ``Site.trajectory``
  (reference trajectory.py:357-360; ``contains_atoms`` in atom list order, site_collection.py:304).
* ``residence_times`` / ``filter_gaps`` -- reference site.py:241-325 and :327-352.
* ``average_occupation`` -- reference site.py:225-239.
* ``transitions`` -- the Counter updates of ``SiteCollection.update_occupation`` (reference
  site_collection.py:300-302), frames in order, atoms in list order, Counter insertion order kept.
* ``runs_from_rows`` -- the run-length form the CUDA fold (``csrc/k_fold.cuh``) must produce.
"""
from __future__ import annotations

import numpy as np


def site_trajectories(rows, atom_index, site_index):
    """rows: (F, A) of site INDEX values (-1 / None = no site).  Returns {site index: [[atom index, ...] per frame]}."""
    out = {int(s): [] for s in site_index}
    for row in rows:
        per = {int(s): [] for s in site_index}
        for a, v in zip(atom_index, row):
            if v is not None and v >= 0:
                per[int(v)].append(int(a))
        for s in out:
            out[s].append(per[s])
    return out


def filter_gaps(occupied, filter_length):
    """site.py:327-352: fill interior gaps of at most filter_length frames."""
    n = len(occupied)
    result = list(occupied)
    f = 0
    while f < n:
        if occupied[f]:
            f += 1
            continue
        g0 = f
        while f < n and not occupied[f]:
            f += 1
        g1 = f                                   # gap is [g0, g1)
        if g0 == 0 or g1 == n:                   # touches an edge: never filled
            continue
        if g1 - g0 <= filter_length:
            for k in range(g0, g1):
                result[k] = True
    return result


def residence_times(trajectory, filter_length=0, include_edge_runs=False):
    """site.py:241-325 for one site's ``trajectory`` (list over frames of atom-index lists)."""
    if not isinstance(filter_length, int):
        raise TypeError
    if filter_length < 0:
        raise ValueError
    if not trajectory:
        return ()
    atoms = set()
    for ts in trajectory:
        atoms.update(ts)
    if not atoms:
        return ()
    n = len(trajectory)
    out = []
    for atom in sorted(atoms):
        occ = [atom in ts for ts in trajectory]
        if filter_length > 0:
            occ = filter_gaps(occ, filter_length)
        f = 0
        while f < n:
            if not occ[f]:
                f += 1
                continue
            s = f
            while f < n and occ[f]:
                f += 1
            e = f                                # run is [s, e)
            if not include_edge_runs and (s == 0 or e == n):
                continue
            out.append(e - s)
    return tuple(out)


def average_occupation(trajectory):
    """site.py:225-239."""
    if not trajectory:
        return None
    return sum(1 for ts in trajectory if ts) / len(trajectory)


def transitions(rows, prev=None):
    """{from: {to: count}} in Counter insertion order; rows (F, A) of site values, < 0 / None = no site.
    prev: the atoms' trajectory entries before the first row (None = empty trajectories)."""
    out: dict[int, dict[int, int]] = {}
    A = len(rows[0]) if len(rows) else 0
    last = [None] * A if prev is None else [None if (p is None or p < 0) else int(p) for p in prev]
    for row in rows:
        for a in range(A):
            v = row[a]
            v = None if (v is None or v < 0) else int(v)
            if v is not None and last[a] is not None and last[a] != v:       # site_collection.py:300-302
                d = out.setdefault(last[a], {})
                d[v] = d.get(v, 0) + 1
        last = [None if (v is None or v < 0) else int(v) for v in row]      # trajectory.py:357
    return out


def runs_from_rows(rows):
    """(site, atom column, start, length) of every maximal run, sorted by (site, atom, start)."""
    rows = np.asarray(rows)
    F, A = rows.shape
    out = []
    for a in range(A):
        f = 0
        while f < F:
            s = int(rows[f, a])
            g = f
            while g < F and rows[g, a] == s:
                g += 1
            if s >= 0:
                out.append((s, a, f, g - f))
            f = g
    out.sort()
    arr = np.array(out, dtype=np.int64).reshape(-1, 4)
    return arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32), arr[:, 2], arr[:, 3]


def random_rows(rng, F, A, S, p_stay=0.9, p_none=0.05):
    """Synthetic history with long stretches, brief excursions, None entries and shared sites."""
    rows = np.empty((F, A), dtype=np.int32)
    cur = rng.integers(-1, S, size=A)
    for f in range(F):
        u = rng.random(A)
        move = u > p_stay
        none = u > 1.0 - p_none
        new = rng.integers(0, S, size=A)
        cur = np.where(move, new, cur)
        cur = np.where(none, -1, cur)
        rows[f] = cur
    return rows
