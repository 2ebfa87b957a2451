"""ctypes wrapper around ``oracle/site_oracle.c`` -- the CPU ORACLE.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module.  ``site_analysis_b200`` never does.

The wrapper restates the host-side pieces the reference computes with numpy /
scipy so that they are produced by the *same library calls*:

* distance-rank table: ``site_collection.py:90-111`` (``np.argsort`` of fractional
  minimum-image norms; unstable sort, so it must be the same numpy call);
* face-sharing neighbours: ``polyhedral_site_collection.py:165-191``;
* face topology: ``scipy.spatial.ConvexHull(vertex_coords).simplices`` at the
  site's first query (``polyhedral_site.py:154-155``), served to the C code
  through a callback.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None

KINDS = {"SphericalSite": 0, "VoronoiSite": 1, "PolyhedralSite": 2, "DynamicVoronoiSite": 3,
         "spherical": 0, "voronoi": 1, "polyhedral": 2, "dynamic_voronoi": 3}

_HULL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int,
                       C.POINTER(C.c_int32))


def build(force: bool = False) -> Path:
    so = HERE / "libsite_oracle.so"
    src = HERE / "site_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        P = C.c_void_p
        dp, ip, lp, bp = (np.ctypeslib.ndpointer(dtype=t, flags="C_CONTIGUOUS")
                          for t in (np.float64, np.int32, np.int64, np.uint8))
        L.so_create.restype = P
        L.so_create.argtypes = [C.c_int, C.c_int, C.c_int, ip]
        L.so_set_centres.argtypes = [P, dp, C.c_void_p]
        L.so_set_ranking.argtypes = [P, ip, dp]
        L.so_set_neighbours.argtypes = [P, ip, ip]
        L.so_set_polyhedra.argtypes = [P, C.c_int, ip, ip, bp, dp, bp, lp, dp]
        L.so_set_topology.argtypes = [P, C.c_int, C.c_int, ip]
        L.so_set_hull_callback.argtypes = [P, _HULL_FN, C.c_void_p]
        L.so_set_dynvor.argtypes = [P, C.c_int, ip, ip, bp, dp]
        L.so_analyse.argtypes = [P, dp, dp]
        L.so_analyse.restype = C.c_int
        L.so_append.argtypes = [P]
        L.so_run.argtypes = [P, dp, C.c_int64, C.c_int64, dp, C.c_int, C.c_void_p, C.c_int]
        L.so_run.restype = C.c_int
        L.so_reset.argtypes = [P]
        L.so_error_message.argtypes = [P]
        L.so_error_message.restype = C.c_char_p
        L.so_get_in_site.argtypes = [P, ip]
        L.so_get_recent.argtypes = [P, ip]
        L.so_get_centres.argtypes = [P, dp]
        L.so_transition_count.argtypes = [P, C.c_int]
        L.so_transition_count.restype = C.c_int
        L.so_get_transitions.argtypes = [P, C.c_int, ip, lp]
        L.so_get_topology.argtypes = [P, C.c_int, C.c_void_p]
        L.so_get_topology.restype = C.c_int
        L.so_get_vertex_coords.argtypes = [P, C.c_int, dp, lp]
        L.so_get_vertex_coords.restype = C.c_int
        L.so_get_stats.argtypes = [P, lp]
        L.so_destroy.argtypes = [P]
        L.so_mic_distance.argtypes = [dp, dp, dp]
        L.so_mic_distance.restype = C.c_double
        L.so_all_mic_distances.argtypes = [dp, C.c_int, dp, C.c_int, dp, dp]
        _LIB = L
    return _LIB


# ---------------------------------------------------------------------------
# host-side tables, restated with the reference's own numpy calls
# ---------------------------------------------------------------------------

def rank_table(centres: np.ndarray) -> np.ndarray:
    """site_collection.py:100-108 in list positions: (S, S-1) int32."""
    centres = np.asarray(centres, dtype=np.float64)
    n = len(centres)
    out = np.zeros((n, max(n - 1, 0)), dtype=np.int32)
    for i in range(n):
        diffs = centres - centres[i]
        diffs -= np.round(diffs)
        dists = np.linalg.norm(diffs, axis=1)
        order = np.argsort(dists)
        out[i] = order[order != i]
    return out


def neighbour_lists(vertex_lists) -> list[list[int]]:
    """polyhedral_site_collection.py:165-191: >= 3 shared vertices, list order."""
    sets = [set(v) for v in vertex_lists]
    out = []
    for i, si in enumerate(sets):
        out.append([j for j, sj in enumerate(sets) if j != i and len(si & sj) >= 3])
    return out


def all_mic_distances(c1, c2, lattice):
    c1 = np.ascontiguousarray(c1, dtype=np.float64); c2 = np.ascontiguousarray(c2, dtype=np.float64)
    out = np.empty((len(c1), len(c2)))
    lib().so_all_mic_distances(c1, len(c1), c2, len(c2), np.ascontiguousarray(lattice, dtype=np.float64), out)
    return out


class SiteOracle:
    """One reference ``Trajectory`` worth of state (atoms + one site collection)."""

    def __init__(self, kind, atom_idx, *, centres=None, rcut=None,
                 vert_idx=None, n_vert=None, has_ref_center=None, ref_center=None,
                 primed=None, shift0=None, cached_raw0=None, topology=None,
                 ref_idx=None, n_ref=None, rank=None, site_index=None):
        self.L = lib()
        self.kind = KINDS[str(kind)]
        self.atom_idx = np.ascontiguousarray(atom_idx, dtype=np.int32)
        self.A = len(self.atom_idx)
        if self.kind in (0, 1):
            centres = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 3)
            self.S = len(centres)
        elif self.kind == 2:
            vert_idx = np.ascontiguousarray(vert_idx, dtype=np.int32)
            self.S = len(vert_idx)
        else:
            ref_idx = np.ascontiguousarray(ref_idx, dtype=np.int32)
            self.S = len(ref_idx)
        self.site_index = (np.arange(self.S, dtype=np.int32) if site_index is None
                           else np.asarray(site_index, dtype=np.int32))
        self.h = self.L.so_create(self.kind, self.S, self.A, self.atom_idx)
        self._cb = None
        if self.kind == 0:
            rc = np.ascontiguousarray(np.broadcast_to(np.asarray(rcut, dtype=np.float64), (self.S,)))
            self.L.so_set_centres(self.h, centres, rc.ctypes.data)
            rk = rank_table(centres) if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
            self.L.so_set_ranking(self.h, rk, centres)
        elif self.kind == 1:
            self.L.so_set_centres(self.h, centres, None)
        elif self.kind == 2:
            n_vert = np.ascontiguousarray(n_vert, dtype=np.int32)
            assert n_vert.max() <= 64
            S, vmax = vert_idx.shape
            hrc = (np.zeros(S, dtype=np.uint8) if has_ref_center is None
                   else np.ascontiguousarray(has_ref_center, dtype=np.uint8))
            rcs = (np.zeros((S, 3)) if ref_center is None
                   else np.ascontiguousarray(np.nan_to_num(ref_center), dtype=np.float64))
            pr = np.zeros(S, dtype=np.uint8) if primed is None else np.ascontiguousarray(primed, dtype=np.uint8)
            s0 = (np.zeros((S, vmax, 3), dtype=np.int64) if shift0 is None
                  else np.ascontiguousarray(shift0, dtype=np.int64))
            r0 = (np.zeros((S, vmax, 3)) if cached_raw0 is None
                  else np.ascontiguousarray(cached_raw0, dtype=np.float64))
            self.L.so_set_polyhedra(self.h, vmax, n_vert, vert_idx, hrc, rcs, pr, s0, r0)
            vlists = [list(vert_idx[s, :n_vert[s]]) for s in range(S)]
            if hrc.all():
                rk = rank_table(rcs) if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
                self.L.so_set_ranking(self.h, rk, rcs)
            else:
                nb = neighbour_lists(vlists)
                off = np.zeros(S + 1, dtype=np.int32)
                off[1:] = np.cumsum([len(x) for x in nb])
                flat = np.ascontiguousarray([j for x in nb for j in x] or [0], dtype=np.int32)
                self.L.so_set_neighbours(self.h, off, flat)
            if topology is not None:
                for s, simp in enumerate(topology):
                    if simp is not None and len(simp):
                        simp = np.ascontiguousarray(simp, dtype=np.int32)
                        self.L.so_set_topology(self.h, s, len(simp), simp)
            self._install_hull_callback()
        else:
            n_ref = np.ascontiguousarray(n_ref, dtype=np.int32)
            assert n_ref.max() <= 64
            S, rmax = ref_idx.shape
            hrc = (np.zeros(S, dtype=np.uint8) if has_ref_center is None
                   else np.ascontiguousarray(has_ref_center, dtype=np.uint8))
            rcs = (np.zeros((S, 3)) if ref_center is None
                   else np.ascontiguousarray(np.nan_to_num(ref_center), dtype=np.float64))
            self.L.so_set_dynvor(self.h, rmax, n_ref, ref_idx, hrc, rcs)

    # -- construction from a golden fixture ---------------------------------
    @classmethod
    def from_case(cls, case, use_case_rank=True, topology=None):
        kind = str(case["kind"])
        kw = dict(site_index=case["site_index"])
        for k in ("centres", "rcut", "vert_idx", "n_vert", "has_ref_center", "ref_center", "primed",
                  "shift0", "cached_raw0", "ref_idx", "n_ref"):
            if k in case:
                kw[k] = case[k]
        if use_case_rank and "rank_table" in case:
            kw["rank"] = np.asarray(case["rank_table"], dtype=np.int32)
        return cls(kind, case["atom_idx"], topology=topology, **kw)

    def _install_hull_callback(self):
        from scipy.spatial import ConvexHull, QhullError

        def cb(_ctx, site, nv, vc, max_faces, out):
            try:
                pts = np.ctypeslib.as_array(vc, shape=(nv * 3,)).reshape(nv, 3).copy()
                simp = np.asarray(ConvexHull(pts).simplices, dtype=np.int32)
            except (QhullError, ValueError):
                return -1
            if len(simp) > max_faces:
                return -1
            for i, v in enumerate(simp.ravel()):
                out[i] = int(v)
            return len(simp)
        self._cb = _HULL_FN(cb)
        self.L.so_set_hull_callback(self.h, self._cb, None)

    # -- running ---------------------------------------------------------------
    def run(self, frac, lattice, append=True, positions=False) -> np.ndarray:
        """trajectory_from_structures over arrays -> (F, A) int32 of Site.index (-1 = None)."""
        frac = np.ascontiguousarray(frac, dtype=np.float64)
        F, N, _ = frac.shape
        lattice = np.ascontiguousarray(lattice, dtype=np.float64)
        stride = 9 if lattice.ndim == 3 else 0
        out = np.empty((F, self.A), dtype=np.int32)
        err = self.L.so_run(self.h, frac, F, N, lattice, stride, out.ctypes.data, int(append))
        if err:
            raise RuntimeError(self.L.so_error_message(self.h).decode())
        return out if positions else self.to_index(out)

    def run_timed(self, frac, lattice) -> None:
        """analyse_structure loop without result storage (the reference benchmark's timed region)."""
        frac = np.ascontiguousarray(frac, dtype=np.float64)
        F, N, _ = frac.shape
        lattice = np.ascontiguousarray(lattice, dtype=np.float64)
        stride = 9 if lattice.ndim == 3 else 0
        err = self.L.so_run(self.h, frac, F, N, lattice, stride, None, 1)
        if err:
            raise RuntimeError(self.L.so_error_message(self.h).decode())

    def to_index(self, pos):
        pos = np.asarray(pos)
        return np.where(pos >= 0, self.site_index[np.clip(pos, 0, None)], -1).astype(np.int32)

    def reset(self):
        self.L.so_reset(self.h)

    # -- state inspection ---------------------------------------------------------
    def transitions(self) -> dict:
        """{src Site.index: {dst Site.index: count}} in Counter insertion order."""
        out = {}
        for s in range(self.S):
            n = self.L.so_transition_count(self.h, s)
            if n:
                k = np.empty(n, dtype=np.int32); c = np.empty(n, dtype=np.int64)
                self.L.so_get_transitions(self.h, s, k, c)
                out[int(self.site_index[s])] = {int(self.site_index[a]): int(b) for a, b in zip(k, c)}
        return out

    def recent(self):
        r = np.empty((self.A, 2), dtype=np.int32)
        self.L.so_get_recent(self.h, r)
        return self.to_index(r)

    def centres(self):
        c = np.empty((self.S, 3))
        self.L.so_get_centres(self.h, c)
        return c

    def topology(self, s):
        n = self.L.so_get_topology(self.h, s, None)
        if n < 0:
            return None
        simp = np.empty((n, 3), dtype=np.int32)
        self.L.so_get_topology(self.h, s, simp.ctypes.data)
        return simp

    def vertex_coords(self, s, nv):
        vc = np.empty((nv, 3)); sh = np.empty((nv, 3), dtype=np.int64)
        n = self.L.so_get_vertex_coords(self.h, s, vc, sh)
        return (vc, sh) if n else (None, None)

    def stats(self) -> dict:
        a = np.zeros(8, dtype=np.int64)
        self.L.so_get_stats(self.h, a)
        return dict(zip(["contains", "full_unwrap", "incremental", "invalidations", "face_updates",
                         "hulls", "lookups", "frames"], map(int, a)))

    def __del__(self):
        try:
            self.L.so_destroy(self.h)
        except Exception:
            pass


def load_case(name, golden_dir=None) -> dict:
    """Load a fixture from tests/golden; resolves ``frames_file`` into a float64 ``frac``."""
    golden_dir = Path(golden_dir) if golden_dir else HERE.parent / "tests" / "golden"
    z = np.load(golden_dir / f"{name}.npz", allow_pickle=False)
    case = {k: z[k] for k in z.files}
    if "frac" not in case:
        fz = np.load(golden_dir / f"{str(case['frames_file'])}.npz", allow_pickle=False)
        case["frac"] = fz["frac_i32"] / 1e8
    return case


def transitions_from_case(case) -> dict:
    out: dict = {}
    for s, d, c in zip(case["tr_src"], case["tr_dst"], case["tr_cnt"]):
        out.setdefault(int(s), {})[int(d)] = int(c)
    return out
