"""Array-level host driver of the CUDA backend: one ``AssignmentEngine`` per site collection.

This is the layer between the reference-shaped Python objects (``Trajectory``,
``*SiteCollection``) and the C ABI in ``include/sab200.h``.  It owns

* the static site tables uploaded to the device,
* the PBC image-shift state (reference ``polyhedral_site.py:377-416``,
  ``dynamic_voronoi_site_collection.py:38-105``) in its per-atom prefix-scan form,
* the sequential history the reference keeps in Python objects -- ``Atom._recent_sites``
  (``atom.py:55``), the previous trajectory entry, ``Site.transitions`` (``site.py:71``) --
  which only the priority resolver and the user-visible attributes need.

PyTorch is used for device memory, pinned staging and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

from . import _native as nat
from . import _tables as tab

EXCURSION_LIMIT = 0.3 - 1e-6     # guard of the scan formulation (pbc_utils.py:216)


class GuardError(RuntimeError):
    """The trajectory violates a precondition under which the frame-parallel formulation is
    provably identical to the reference's sequential caches ("Guards")."""


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise nat.NativeError("site_analysis_b200 needs a CUDA device: there is no CPU fallback")
    return torch


def fold_history(recent: np.ndarray, prev: np.ndarray, transitions: list, rows: np.ndarray, append: bool) -> None:
    """Advance (recent, prev, transitions) over finished result rows, in place.

    Vectorised restatement of what the reference does per atom per frame in
    ``SiteCollection.update_occupation`` (site_collection.py:300-310) and
    ``Atom.update_recent_site`` (atom.py:190-192), followed by ``atom.trajectory.append``
    (trajectory.py:357) when ``append``.  ``rows`` is (F, A) of list positions, -1 = None.
    """
    rows = np.asarray(rows)
    F, A = rows.shape
    if F == 0:
        return
    if not append and F != 1:
        raise ValueError("analyse-without-append is defined for one frame at a time")
    before = np.vstack([prev[None, :], rows[:-1]]) if append else prev[None, :]
    moved = (before >= 0) & (rows >= 0) & (before != rows)
    if moved.any():
        f_idx, a_idx = np.nonzero(moved)              # row-major == (frame, atom) order
        src = before[f_idx, a_idx].astype(np.int64)
        dst = rows[f_idx, a_idx].astype(np.int64)
        key = src * (int(rows.max()) + 2) + dst
        _, first, counts = np.unique(key, return_index=True, return_counts=True)
        for i in np.argsort(first, kind="stable"):    # Counter insertion order = first occurrence
            d = transitions[int(src[first[i]])]
            k = int(dst[first[i]])
            d[k] = d.get(k, 0) + int(counts[i])
    # two most recent distinct assigned sites per atom
    assigned = rows >= 0
    any_assigned = assigned.any(axis=0)
    if any_assigned.any():
        last_f = F - 1 - np.argmax(assigned[::-1], axis=0)
        cols = np.arange(A)
        last_v = rows[last_f, cols]
        other = assigned & (rows != last_v[None, :])
        has_other = other.any(axis=0)
        other_f = F - 1 - np.argmax(other[::-1], axis=0)
        other_v = rows[other_f, cols]
        old0, old1 = recent[:, 0].copy(), recent[:, 1].copy()
        new1 = np.where(has_other, other_v, np.where(old0 == last_v, old1, old0))
        recent[:, 0] = np.where(any_assigned, last_v, old0)
        recent[:, 1] = np.where(any_assigned, new1, old1)
    if append:
        prev[:] = rows[-1]


DEVICE_FOLD_MIN = 1 << 15     # atom-frames per batch from which the history is folded on the GPU


def fold_history_device(recent: np.ndarray, prev: np.ndarray, transitions, rows, n_sites: int) -> None:
    """:func:`fold_history` (``append=True``) with the O(F A) part on the GPU: ``sab_fold_result`` (csrc/k_fold.cuh)
    returns the transition pairs in first-occurrence order and each atom's last / last-different site; only O(A + pairs)
    work is left for the host.  ``rows`` (F, A) int32 positions, CUDA tensor or numpy array.  ``transitions``: the list of
    per-site dicts, or an :class:`AssignmentEngine` (then the pairs are merged into its array store without touching
    any dict)."""
    from .folds import fold_rows
    out = fold_rows(rows, n_sites, prev=prev, want_pairs=True, want_recent=True, want_runs=False)
    src, dst, cnt, _ = out["pairs"]
    if hasattr(transitions, "_tr_add_pairs"):
        transitions._tr_add_pairs(src, dst, cnt)
    else:
        for a, b, c in zip(src.tolist(), dst.tolist(), cnt.tolist()):  # sorted by first occurrence: Counter insertion order
            d = transitions[a]
            d[b] = d.get(b, 0) + c
    last_v, other_v = out["recent"]
    any_assigned, has_other = last_v >= 0, other_v >= 0
    old0, old1 = recent[:, 0].copy(), recent[:, 1].copy()
    new1 = np.where(has_other, other_v, np.where(old0 == last_v, old1, old0))
    recent[:, 0] = np.where(any_assigned, last_v, old0)
    recent[:, 1] = np.where(any_assigned, new1, old1)
    last_row = rows[-1]
    prev[:] = last_row if isinstance(last_row, np.ndarray) else last_row.cpu().numpy()


class AssignmentEngine:
    """GPU assignment of every mobile atom in every frame to a site (list positions).

    Args:
        kind: ``"spherical" | "voronoi" | "polyhedral" | "dynamic_voronoi"``.
        n_atoms: atoms per frame N.
        mobile_idx: structure indices of the tracked atoms (A,).
        centres, rcut: spherical / Voronoi site tables.
        slots: per site, the vertex (polyhedral) or reference (dynamic Voronoi) atom indices.
        reference_centres: per site ``None`` or a (3,) array (PBC unwrapping target).
        primed: per polyhedral site ``None`` or ``(image_shifts (V,3), cached_raw (V,3))`` as
            left by the reference's ``SiteFactory`` (site_factory.py:248-257).
        simplices: per polyhedral site ``None`` or an existing (n_faces, 3) face topology.
        rank_table: optional precomputed distance-rank table (S, S-1).
        strict: raise ``GuardError`` when a guard fails (default) instead of warning.
    """

    KINDS = {"spherical": nat.SPHERICAL, "voronoi": nat.VORONOI, "polyhedral": nat.POLYHEDRAL,
             "dynamic_voronoi": nat.DYNAMIC_VORONOI}

    def __init__(self, kind: str, n_atoms: int, mobile_idx, *, centres=None, rcut=None,
                 slots: Sequence[Sequence[int]] | None = None, reference_centres=None, primed=None,
                 simplices=None, rank_table=None, device: int | None = None, strict: bool = True):
        torch = _torch()
        self.L = nat.lib()
        self.kind = kind
        self.k = self.KINDS[kind]
        self.N = int(n_atoms)
        self.mobile = np.ascontiguousarray(mobile_idx, dtype=np.int32)
        self.A = len(self.mobile)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.strict = strict
        self.rank_table = None if rank_table is None else np.ascontiguousarray(rank_table, dtype=np.int32)
        self._open_topology = np.zeros(0, dtype=bool)
        self.rank_dense_max = 2048      # above this many sites the distance ranking is lazy (no S x S tables)
        self._priority_uploaded = False
        self.slots = None
        if self.k in (nat.SPHERICAL, nat.VORONOI):
            self.centres = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 3)
            self.S = len(self.centres)
        else:
            self.slots = [list(map(int, s)) for s in slots]
            self.S = len(self.slots)
            self.ref_centres = ([None] * self.S if reference_centres is None else
                                [None if (c is None or np.any(np.isnan(c))) else np.asarray(c, dtype=np.float64)
                                 for c in reference_centres])
            self.primed = [None] * self.S if primed is None else list(primed)
            self.simplices = [None] * self.S if simplices is None else list(simplices)
            # Non-simplicial polyhedra without a cached face topology: the reference hulls a site at its FIRST QUERY
            # (polyhedral_site.py:468-471) and Qhull's triangulation of e.g. a cube's quads depends on that frame's vertex
            # coordinates.  Such sites are "open": the exact replay kernel reports the coordinates of their first query and
            # the host hulls from exactly those (_assign_device_lazy).  Simplicial sites need none of this (every vertex
            # triple is a face).
            self._open_topology = np.array([self.simplices[s] is None and len(self.slots[s]) > 4 for s in range(self.S)], dtype=bool) \
                if self.k == nat.POLYHEDRAL else np.zeros(self.S, dtype=bool)
        h = C.c_void_p()
        nat.check(self.L.sab_engine_create(C.byref(h), self.k, self.device, self.N, self.A,
                                           nat.ptr(self.mobile), self.S))
        self.h = h
        if self.k == nat.SPHERICAL:
            self.rcut = np.ascontiguousarray(np.broadcast_to(np.asarray(rcut, dtype=np.float64), (self.S,)))
            self.rcut_q = np.ascontiguousarray(tab.sqrt_threshold(self.rcut))
            nat.check(self.L.sab_set_centres(self.h, nat.ptr(self.centres), nat.ptr(self.rcut), nat.ptr(self.rcut_q)))
        elif self.k == nat.VORONOI:
            nat.check(self.L.sab_set_centres(self.h, nat.ptr(self.centres), None, None))
        elif self.k == nat.DYNAMIC_VORONOI:
            n_ref, ref = self._padded(self.slots)
            self._n_slot, self._slot_atom = n_ref, ref
            nat.check(self.L.sab_set_reference_atoms(self.h, ref.shape[1], nat.ptr(n_ref), nat.ptr(ref)))
            self._load_framework()
            # dynamic_voronoi_site_collection.py:151-164: groups by reference count, first seen first
            seen: list[int] = []
            for s in self.slots:
                if len(s) not in seen:
                    seen.append(len(s))
            self.group_of_site = np.array([seen.index(len(s)) for s in self.slots], dtype=np.int32)
        else:
            self._n_slot, self._slot_atom = self._padded(self.slots)
        self._pbc_ready = False
        self.frames_done = 0
        # polyhedral sites: exact replay of the reference's lazily updated per-site caches (csrc/k_polyhedral_lazy.cuh).
        # exact_lazy = True uses it from the first frame on, False never; None (default): it serves engines that are fed
        # frame by frame (first call shorter than LAZY_AUTO_FRAMES: the reference's own append_timestep loop), and it takes
        # over when a guard of the frame-parallel formulation fails in a batch that starts from the initial state
        # (DESIGN.md section 4).
        self.exact_lazy = None
        self._lazy_mode = False
        self._g2_cart_margin = np.inf          # guard G2: how far an unprimed vertex may move before its image choice can change
        self._g2_frac_margin = np.inf
        self.reset_history()
        self.last_info: dict = {}
        # exact-pruning tables of the tetrahedral kernel (built lazily from the first sizeable batch)
        self.use_cells = True
        self.nearest_cell_factor = float(os.environ.get("SAB_NEAREST_CELL_FACTOR", "0.2"))   # cell edge / mean centre spacing
        self.cells_min_frames = 64
        self._cells = None              # dict(G, delta, n_entries) once uploaded
        self._cells_stale = False
        self._seen_exc = None

    # ------------------------------------------------------------------ small helpers
    @staticmethod
    def _padded(lists):
        n = np.array([len(x) for x in lists], dtype=np.int32)
        out = np.zeros((len(lists), int(n.max())), dtype=np.int32)
        for i, x in enumerate(lists):
            out[i, :len(x)] = x
        return n, np.ascontiguousarray(out)

    def _load_framework(self):
        M = self.L.sab_framework_count(self.h)
        self.fw_atoms = np.zeros(M, dtype=np.int32)
        nat.check(self.L.sab_framework_atoms(self.h, nat.ptr(self.fw_atoms)))
        lut = -np.ones(self.N, dtype=np.int64)
        lut[self.fw_atoms] = np.arange(M)
        self.slot_fw = lut[self._slot_atom]           # (S, smax) compact ids (padding maps atom 0)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.sab_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_option(self, key: int, value: int) -> None:
        nat.check(self.L.sab_set_option(self.h, int(key), int(value)))

    # ------------------------------------------------------------------ history (host)
    # Site.transitions of all sites (site.py:71): per-site dicts {destination position: count} in Counter insertion order.
    # Long batches are folded on the GPU into (source, destination, count) arrays; the dicts are built from them only when
    # somebody asks (the priority resolver, the Python objects), so streaming batches never pay for S Python dicts.
    @property
    def transitions(self) -> list:
        if self._tr_dicts is None:
            src, dst, cnt = self._tr_arrays
            order = np.argsort(src, kind="stable")               # per source, insertion order kept
            bounds = np.searchsorted(src[order], np.arange(self.S + 1))
            d_l, c_l = dst[order].tolist(), cnt[order].tolist()
            self._tr_dicts = [dict(zip(d_l[a:b], c_l[a:b])) if b > a else {} for a, b in zip(bounds[:-1].tolist(), bounds[1:].tolist())]
            self._tr_arrays = None                               # the caller may edit the dicts: they are authoritative now
        return self._tr_dicts

    @transitions.setter
    def transitions(self, value) -> None:
        self._tr_dicts, self._tr_arrays = value, None

    def _tr_snapshot(self):
        """Copy of the transition store in whichever form it is held (no array <-> dict conversion)."""
        if self._tr_dicts is None:
            return ("arrays", tuple(x.copy() for x in self._tr_arrays))
        return ("dicts", [dict(t) for t in self._tr_dicts])

    def _tr_restore(self, snap) -> None:
        kind, value = snap
        if kind == "arrays":
            self._tr_arrays, self._tr_dicts = value, None
        else:
            self._tr_dicts, self._tr_arrays = value, None

    def _tr_add_pairs(self, src, dst, cnt) -> None:
        """Merge transition pairs (in first-occurrence order) into the array store."""
        if self._tr_arrays is None:
            s0 = np.fromiter((s for s, t in enumerate(self._tr_dicts) for _ in t), dtype=np.int64)
            d0 = np.fromiter((k for t in self._tr_dicts for k in t), dtype=np.int64)
            c0 = np.fromiter((v for t in self._tr_dicts for v in t.values()), dtype=np.int64)
        else:
            s0, d0, c0 = self._tr_arrays
        s1 = np.concatenate([s0, np.asarray(src, dtype=np.int64)]); d1 = np.concatenate([d0, np.asarray(dst, dtype=np.int64)])
        c1 = np.concatenate([c0, np.asarray(cnt, dtype=np.int64)])
        _, first, inv = np.unique(s1 * self.S + d1, return_index=True, return_inverse=True)
        total = np.zeros(len(first), dtype=np.int64)
        np.add.at(total, inv, c1)
        order = np.argsort(first, kind="stable")                 # insertion order = first occurrence
        self._tr_arrays = (s1[first[order]], d1[first[order]], total[order])
        self._tr_dicts = None

    def reset_history(self):
        self.recent = -np.ones((self.A, 2), dtype=np.int32)
        self.prev = np.full(self.A, -2, dtype=np.int32)
        self.transitions = [dict() for _ in range(self.S)]
        self._pending: list[tuple[np.ndarray, bool]] = []

    def reset(self):
        """Trajectory.reset(): history and PBC caches cleared, face topology kept
        (trajectory.py:364-373, polyhedral_site.py:264-279, dynamic_voronoi_site_collection.py:197-201)."""
        self.reset_history()
        self._pbc_ready = False
        self._lazy_mode = False
        self._initial_pbc = None          # distributed.py: state at the start of the sharded trajectory, device copies of it
        self._shard_dev = None
        self._history_stale = False
        self._sharded_checks = []
        if self.k in (nat.POLYHEDRAL, nat.DYNAMIC_VORONOI):   # anchors belong to the PBC state; Voronoi lists are static
            self._cells = None
            if self.k == nat.DYNAMIC_VORONOI:
                nat.check(self.L.sab_set_dynvor_lists(self.h, 1, None, None, None, 0.0, None, None, None))
        self._seen_exc = None
        if self.slots is not None:
            self.primed = [None] * self.S
        self.frames_done = 0

    def sync_history(self):
        """Fold every finished batch into (recent, prev, transitions), in order; batches whose ambiguous
        atom-frames were left for later (``assign_device(defer_resolve=True)``) go through the resolver here."""
        pending, self._pending = self._pending, []
        for item in pending:
            if isinstance(item[0], str):
                self._resolve_now(*item[1:])
                continue
            rows, append = item
            if append and rows.shape[0] * rows.shape[1] >= DEVICE_FOLD_MIN:
                fold_history_device(self.recent, self.prev, self, rows, self.S)
                continue
            if not isinstance(rows, np.ndarray):
                rows = rows.cpu().numpy()
            fold_history(self.recent, self.prev, self.transitions, rows, append)

    # ------------------------------------------------------------------ PBC state initialisation
    def _init_pbc(self, raw0: np.ndarray, lat0: np.ndarray):
        """Build slot shifts relative to prev_raw = the first frame (W = 0).

        Polyhedral sites primed by the builder carry (shift, cached raw) from the build
        structure; they are advanced to the first frame exactly as ``update_pbc_shifts``
        would (pbc_utils.py:210-217).  Unprimed sites (hand-made, or after reset) get the
        reference's full correction on the first frame (polyhedral_site.py:410-416;
        dynamic_voronoi_site_collection.py:188-195).
        """
        S = self.S
        smax = self._slot_atom.shape[1]
        self._initial_pbc = None          # cached by distributed._initial_state / _carry_by_anchor: stale from here on
        self._shard_dev = None
        slot_shift = np.zeros((S, smax, 3), dtype=np.int32)
        legacy = np.zeros(S, dtype=np.uint8)
        self._g2_cart_margin, self._g2_frac_margin = np.inf, np.inf
        lo = {}
        for s in range(S):
            atoms = self.slots[s]
            r = raw0[atoms]
            pr = self.primed[s] if self.k == nat.POLYHEDRAL else None
            done = False
            if pr is not None:
                sh, cached = np.asarray(pr[0], dtype=np.int64), np.asarray(pr[1], dtype=np.float64)
                diff = r - cached
                w = np.round(diff)
                if np.all(np.abs(diff - w) < 0.3):
                    slot_shift[s, :len(atoms)] = sh - w.astype(np.int64)
                    for a, c in zip(atoms, cached + w):      # cached point in the unwrapped frame
                        lo.setdefault(a, []).append(c)
                    done = True
            if not done:
                rc = self.ref_centres[s]
                slot_shift[s, :len(atoms)] = tab.full_pbc_unwrap(r, rc, lat0)
                if rc is None:
                    legacy[s] = 1
                if self.k == nat.POLYHEDRAL:                    # guard G2 margins of this lazily initialised site
                    cart, frac_m = tab.unwrap_margins(r, rc, lat0)
                    self._g2_cart_margin = min(self._g2_cart_margin, cart)
                    self._g2_frac_margin = min(self._g2_frac_margin, frac_m)
        if self.k == nat.POLYHEDRAL:
            n_face = np.zeros(S, dtype=np.int32)
            faces = []
            for s in range(S):
                if self.simplices[s] is None:
                    atoms = self.slots[s]
                    vc = tab.unwrapped_vertices(raw0[atoms], slot_shift[s, :len(atoms)], non_negative=not legacy[s])
                    self.simplices[s] = tab.hull_simplices(vc, s)
                faces.append(np.asarray(self.simplices[s], dtype=np.int32))
                n_face[s] = len(faces[-1])
            fmax = int(n_face.max())
            simp = np.zeros((S, fmax, 3), dtype=np.int32)
            for s, f in enumerate(faces):
                simp[s, :len(f)] = f
            nat.check(self.L.sab_set_polyhedra(self.h, smax, nat.ptr(self._n_slot), nat.ptr(self._slot_atom),
                                               fmax, nat.ptr(n_face), nat.ptr(np.ascontiguousarray(simp))))
            self._load_framework()
        M = len(self.fw_atoms)
        prev_raw = np.ascontiguousarray(raw0[self.fw_atoms], dtype=np.float64)
        wraps = np.zeros((M, 3), dtype=np.int32)
        exc = np.stack([prev_raw, prev_raw], axis=-1)           # (M, 3, 2) min, max
        pos = {int(a): i for i, a in enumerate(self.fw_atoms)}
        for a, pts in lo.items():
            p = np.array(pts)
            exc[pos[a], :, 0] = np.minimum(exc[pos[a], :, 0], p.min(axis=0))
            exc[pos[a], :, 1] = np.maximum(exc[pos[a], :, 1], p.max(axis=0))
        self.slot_shift = np.ascontiguousarray(slot_shift)
        self._lat0 = np.asarray(lat0, dtype=np.float64)
        nat.check(self.L.sab_set_pbc_state(self.h, nat.ptr(self.slot_shift), nat.ptr(prev_raw), nat.ptr(wraps),
                                           nat.ptr(np.ascontiguousarray(exc))))
        self._legacy_flags = legacy
        if legacy.any():
            nat.check(self.L.sab_set_legacy_init(self.h, nat.ptr(legacy), 0))
        else:
            nat.check(self.L.sab_set_legacy_init(self.h, None, 0))
        self._pbc_ready = True

    # ------------------------------------------------------------------ exact-pruning tables
    def _maybe_build_cells(self, d_frac, nf, d_ws, ws_bytes, stream, margin=0.55):
        """Tetrahedral sites: place anchors at the middle of each framework coordinate's observed
        range, size the validity tube ``delta`` from the widest range, and build the cell lists
        (``_tables.tet_cell_lists``).  Purely an accelerator: frames outside the tube fall back to
        the exhaustive path inside the kernel."""
        if self.k != nat.POLYHEDRAL or not self.use_cells:
            return
        if self._cells is not None and not self._cells_stale:
            return
        if nf < self.cells_min_frames:
            return
        if any(len(s) != 4 for s in self.slots) or any(len(f) != 4 for f in self.simplices) or self.S > 65535:
            return
        M = len(self.fw_atoms)
        exc = np.zeros((M, 3, 2))
        nat.check(self.L.sab_scan_excursion(self.h, d_frac.data_ptr(), nf, d_ws.data_ptr(), ws_bytes, stream, nat.ptr(exc)))
        if self._seen_exc is not None:
            exc[..., 0] = np.minimum(exc[..., 0], self._seen_exc[..., 0])
            exc[..., 1] = np.maximum(exc[..., 1], self._seen_exc[..., 1])
        self._seen_exc = exc
        # anchors travel to the stream kernel's producers as float: keep them exactly representable (the tube's
        # margin dwarfs the 6e-8 relative rounding; every kernel uses the same rounded value)
        anchor = np.ascontiguousarray((0.5 * (exc[..., 0] + exc[..., 1])).astype(np.float32).astype(np.float64))
        delta = float(max(margin * (exc[..., 1] - exc[..., 0]).max() + 1e-6, 5e-3))
        ss = self.slot_shift[:, :4, :].astype(np.int64)
        kmin, kmax = int(ss.min()), int(ss.max())
        nk = kmax - kmin + 1
        P = anchor[self.slot_fw[:, :4]] + ss                        # (S, 4, 3) anchor vertices, unwrapped
        extent = (P.max(axis=1) - P.min(axis=1)).max()
        if nk > 8 or nk * 3 * M * 8 > 65535 or delta > 0.12 or extent + 2 * delta + 1e-6 >= 0.45:
            self._cells, self._cells_stale = None, False
            nat.check(self.L.sab_set_cell_lists(self.h, 1, 1.0, None, 0, 1, None, None, None))
            return
        length = float(np.mean(np.linalg.norm(self._lat0, axis=1)))
        G = int(min(64, max(4, round(length / 0.65 / 4) * 4)))
        off, ent = tab.tet_cell_lists(P, delta, G, eps=2e-6)    # 2e-6: the tile kernel looks cells up in float32
        try:
            rec = np.ascontiguousarray(tab.tet_records(self.slot_shift, self.slot_fw, self.simplices, kmin, 3 * M))
        except ValueError:            # face list is not a tetrahedron's hull: the general kernel handles it
            self._cells, self._cells_stale = None, False
            nat.check(self.L.sab_set_cell_lists(self.h, 1, 1.0, None, 0, 1, None, None, None))
            return
        ent = np.ascontiguousarray(ent if len(ent) else np.zeros(1, dtype=np.uint16))
        nat.check(self.L.sab_set_cell_lists(self.h, G, delta * (1.0 - 1e-9), nat.ptr(anchor), kmin, nk, nat.ptr(rec),
                                            nat.ptr(np.ascontiguousarray(off)), nat.ptr(ent)))
        self._cells = dict(G=G, delta=delta, entries=int(len(ent)), mean_candidates=float(len(ent)) / G ** 3)
        self._cells_stale = False

    def _exhaustive_too_big(self) -> bool:
        """The exhaustive kernels stage every site in shared memory (24 B per Voronoi / dynamic-Voronoi centre, 40 B per
        sphere): beyond ~200 KB they cannot launch, so even single-frame calls (the reference's per-frame API) build the
        cell lists first."""
        per_site = 40 if self.k == nat.SPHERICAL else 24
        return self.S * per_site > 200_000

    def _build_nearest_lists(self, lattice0):
        """Voronoi: cell lists of the centres that can be nearest anywhere in a cell (``_tables.nearest_cell_lists``)."""
        lattice0 = np.asarray(lattice0, dtype=np.float64)
        length = float(np.mean(np.linalg.norm(lattice0, axis=1)))
        spacing = (abs(np.linalg.det(lattice0)) / self.S) ** (1.0 / 3.0)        # mean centre spacing
        G = int(min(64, max(2, round(length / max(self.nearest_cell_factor * spacing, 0.25)))))
        off, ent, excl = tab.nearest_cell_lists(self.centres, lattice0, G)
        inv = np.ascontiguousarray(np.linalg.inv(lattice0))
        nat.check(self.L.sab_set_nearest_lists(self.h, G, nat.ptr(inv), nat.ptr(np.ascontiguousarray(off)),
                                               nat.ptr(np.ascontiguousarray(ent)), nat.ptr(np.ascontiguousarray(excl))))
        self._cells = dict(G=G, entries=int(len(ent)), mean_candidates=float(len(ent)) / G ** 3)

    def _build_sphere_lists(self, lattice0, rho: float = 1.06):
        """Spherical: cell lists of the sites whose (inflated) sphere reaches a cell (``_tables.sphere_cell_lists``)."""
        lattice0 = np.asarray(lattice0, dtype=np.float64)
        length = float(np.mean(np.linalg.norm(lattice0, axis=1)))
        G = int(min(64, max(2, round(length / max(0.7 * float(self.rcut.max()), 0.5)))))
        off, ent = tab.sphere_cell_lists(self.centres, self.rcut, lattice0, G, rho)
        ent = np.ascontiguousarray(ent if len(ent) else np.zeros(1, dtype=np.uint16))
        inv = np.ascontiguousarray(np.linalg.inv(lattice0))
        nat.check(self.L.sab_set_sphere_lists(self.h, G, nat.ptr(inv), float(rho), nat.ptr(np.ascontiguousarray(off)), nat.ptr(ent)))
        self._cells = dict(G=G, entries=int(off[-1]), mean_candidates=float(off[-1]) / G ** 3, rho=rho)

    def _build_dynvor_lists(self, d_frac, lattice0):
        """Dynamic Voronoi: anchors = centres of the batch's first frame, head-room from a strided sample."""
        torch = _torch()
        lattice0 = np.asarray(lattice0, dtype=np.float64)
        F = int(d_frac.shape[0])
        stride = max(1, F // 256)
        sample = d_frac[::stride][:256].contiguous()
        cen = self.dynvor_centres(sample).cpu().numpy()                         # (n, S, 3)
        rel = cen - cen[:1]
        rel -= np.round(rel)                                                     # displacement from the first sample
        anchors = np.ascontiguousarray(cen[0] + rel.mean(axis=0))                # mean position: half the head-room
        d = tab._mic_cartesian(cen - anchors[None], lattice0)
        move = float(d.max() * 1.5 + 0.05)
        length = float(np.mean(np.linalg.norm(lattice0, axis=1)))
        spacing = (abs(np.linalg.det(lattice0)) / self.S) ** (1.0 / 3.0)
        G = int(min(64, max(2, round(length / max(self.nearest_cell_factor * spacing, 0.25)))))
        off, ent, excl = tab.nearest_cell_lists(anchors, lattice0, G, move=move)
        inv = np.ascontiguousarray(np.linalg.inv(lattice0))
        nat.check(self.L.sab_set_dynvor_lists(self.h, G, nat.ptr(np.ascontiguousarray(lattice0)), nat.ptr(inv), nat.ptr(anchors),
                                              move, nat.ptr(np.ascontiguousarray(off)), nat.ptr(np.ascontiguousarray(ent)),
                                              nat.ptr(np.ascontiguousarray(excl))))
        self._cells = dict(G=G, entries=int(len(ent)), mean_candidates=float(len(ent)) / G ** 3, max_move=move)

    def _reinit_groups(self, raw_t, lat_t):
        """Dynamic Voronoi: a step >= 0.3 at this frame re-initialises every group that contains a
        jumping reference slot (dynamic_voronoi_site_collection.py:94-95, 188-195)."""
        M = len(self.fw_atoms)
        sh = np.zeros((self.S,) + self.slot_shift.shape[1:], dtype=np.int32)
        prev_raw = np.zeros((M, 3)); wraps = np.zeros((M, 3), dtype=np.int32); exc = np.zeros((M, 3, 2))
        nat.check(self.L.sab_get_pbc_state(self.h, nat.ptr(sh), nat.ptr(prev_raw), nat.ptr(wraps), nat.ptr(exc)))
        diff = raw_t[self.fw_atoms] - prev_raw
        w = np.round(diff)
        bad_atom = np.any(np.abs(diff - w) >= 0.3, axis=1)                     # (M,)
        w_t = wraps + w.astype(np.int32)                                       # W after this frame
        legacy = np.zeros(self.S, dtype=np.uint8)
        bad_groups = set()
        for s in range(self.S):
            if bad_atom[self.slot_fw[s, :len(self.slots[s])]].any():
                bad_groups.add(int(self.group_of_site[s]))
        for s in range(self.S):
            if int(self.group_of_site[s]) not in bad_groups:
                continue
            atoms = self.slots[s]
            rc = self.ref_centres[s]
            image = tab.full_pbc_unwrap(raw_t[atoms], rc, lat_t)
            sh[s, :len(atoms)] = image + w_t[self.slot_fw[s, :len(atoms)]]
            if rc is None:
                legacy[s] = 1
        # the jumping atoms start a fresh excursion record at this frame
        u_t = raw_t[self.fw_atoms] - w_t
        exc[bad_atom, :, 0] = u_t[bad_atom]
        exc[bad_atom, :, 1] = u_t[bad_atom]
        self.slot_shift = np.ascontiguousarray(sh)
        nat.check(self.L.sab_set_pbc_state(self.h, nat.ptr(self.slot_shift), nat.ptr(np.ascontiguousarray(prev_raw)),
                                           nat.ptr(np.ascontiguousarray(wraps)), nat.ptr(np.ascontiguousarray(exc))))
        nat.check(self.L.sab_set_legacy_init(self.h, nat.ptr(legacy) if legacy.any() else None, 0))
        nat.check(self.L.sab_set_option(self.h, nat.OPT_SKIP_STEP_CHECK_ONCE, 1))

    def _check_excursion(self, span_bits=None):
        if self.k != nat.POLYHEDRAL:
            return
        if span_bits is not None and float(np.array([span_bits], dtype=np.int64).view(np.float64)[0]) < EXCURSION_LIMIT:
            # the device-side maximum already clears guard G1; guard G2: sites without cached shifts took the full unwrap at
            # the engine's first frame, the reference takes it at the site's first QUERY -- the same image choice as long as
            # no vertex moved further than half the margin between its best and second-best image (legacy rule: than its
            # distance from the rule's thresholds)
            span = float(np.array([span_bits], dtype=np.int64).view(np.float64)[0])
            if np.isfinite(self._g2_cart_margin) or np.isfinite(self._g2_frac_margin):
                cart = span * float(np.abs(self._lat0).sum()) * 1.01
                if cart >= self._g2_cart_margin or span >= self._g2_frac_margin:
                    msg = (f"unwrapped framework excursion {span:.3f}: a site without cached PBC shifts could choose another "
                           "periodic image at its first query than at the first frame (guard G2)")
                    if self.strict:
                        raise GuardError(msg)
                    import warnings
                    warnings.warn(msg)
            return
        M = len(self.fw_atoms)
        exc = np.zeros((M, 3, 2))
        nat.check(self.L.sab_get_pbc_state(self.h, None, None, None, nat.ptr(exc)))
        span = exc[:, :, 1] - exc[:, :, 0]
        if span.max() >= EXCURSION_LIMIT:
            m, d = np.unravel_index(np.argmax(span), span.shape)
            msg = (f"framework atom {int(self.fw_atoms[m])} moved over {span.max():.3f} fractional units along "
                   f"axis {d}; beyond {EXCURSION_LIMIT:.3f} the reference's lazily updated PBC caches "
                   "(polyhedral_site.py:398-416) are no longer equivalent to the frame-parallel scan")
            if self.strict:
                raise GuardError(msg)
            import warnings
            warnings.warn(msg)

    # ------------------------------------------------------------------ priority tables
    def _upload_priority_tables(self):
        if self._priority_uploaded:
            return
        rank = lookup = off = flat = None
        if self.k == nat.SPHERICAL:
            lookup = self.centres
        elif self.k == nat.POLYHEDRAL and all(c is not None for c in self.ref_centres):
            lookup = np.ascontiguousarray(np.array(self.ref_centres, dtype=np.float64))
        self._lazy_rank = False
        if lookup is not None:
            # small systems (or a table handed in): the dense S x (S - 1) table of site_collection.py:90-111.  Large ones:
            # lazy ranking -- candidates are ranked by distance on the device, true rows only for anchors with tied candidates
            if self.rank_table is not None or self.S <= self.rank_dense_max:
                rank = self.rank_table if self.rank_table is not None else tab.distance_rank_table(lookup)
                self.rank_table = np.ascontiguousarray(rank, dtype=np.int32)
            else:
                self._lazy_rank = True
                self._lookup = np.ascontiguousarray(lookup, dtype=np.float64)
        elif self.k == nat.POLYHEDRAL:
            off, flat = tab.face_sharing_neighbours(self.slots)
        nat.check(self.L.sab_set_priority_tables(self.h, nat.ptr(self.rank_table) if (lookup is not None and not self._lazy_rank) else None,
                                                 nat.ptr(lookup), nat.ptr(off), nat.ptr(flat)))
        self._priority_uploaded = True

    def _supply_missing_rank_rows(self) -> int:
        """Lazy ranking: rows of the anchors the last resolver / replay run reported (tied candidates), computed with the
        reference's own numpy calls (site_collection.py:100-107) and cached on the device.  Returns how many were added."""
        if not getattr(self, "_lazy_rank", False):
            return 0
        buf = np.zeros(4096, dtype=np.int32)
        n = int(self.L.sab_rank_missing(self.h, nat.ptr(buf), len(buf)))
        if n <= 0:
            return 0
        anchors = np.ascontiguousarray(buf[:min(n, len(buf))])
        rows = np.empty((len(anchors), self.S - 1), dtype=np.int32)
        for r, i in enumerate(anchors):
            delta = self._lookup - self._lookup[i]
            delta -= np.round(delta)
            order = np.argsort(np.linalg.norm(delta, axis=1))
            rows[r] = order[order != i]
        nat.check(self.L.sab_add_rank_rows(self.h, len(anchors), nat.ptr(anchors), nat.ptr(rows)))
        self.rank_rows_cached = getattr(self, "rank_rows_cached", 0) + len(anchors)
        return len(anchors)

    def _resolve(self, d_frac, F, d_result, d_spill, stream, append):
        self.sync_history()
        self._resolve_now(d_frac, F, d_result, d_spill, stream, append)

    def _resolve_now(self, d_frac, F, d_result, d_spill, stream, append):
        self._upload_priority_tables()
        width = max(12, 2 * max((len(t) for t in self.transitions), default=0))
        while True:
            tr_n = np.zeros(self.S, dtype=np.int32)
            tr_key = np.zeros((self.S, width), dtype=np.int32)
            tr_cnt = np.zeros((self.S, width), dtype=np.int64)
            for s, t in enumerate(self.transitions):
                tr_n[s] = len(t)
                if t:
                    tr_key[s, :len(t)] = list(t.keys())
                    tr_cnt[s, :len(t)] = list(t.values())
            recent, prev = self.recent.copy(), self.prev.copy()
            d_backup = d_result.clone()
            code = self.L.sab_resolve(self.h, d_frac.data_ptr(), F, d_result.data_ptr(), d_spill.data_ptr(), stream,
                                      nat.ptr(recent), nat.ptr(prev), nat.ptr(tr_n), nat.ptr(tr_key), nat.ptr(tr_cnt), width)
            if code == 3 and width < 4 * self.S:         # SAB_EOVERFLOW: widen the counter rows, redo
                d_result.copy_(d_backup)
                width *= 2
                continue
            nat.check(code)
            if self._supply_missing_rank_rows():          # lazy ranking: a tie needed an anchor's true row -- redo with it
                d_result.copy_(d_backup)
                continue
            break
        self.recent = recent
        if append:
            self.prev = prev
        self.transitions = [dict(zip(map(int, tr_key[s, :tr_n[s]]), map(int, tr_cnt[s, :tr_n[s]])))
                            for s in range(self.S)]

    # ------------------------------------------------------------------ the hot path
    def set_result_peers(self, peer_slot_ptrs) -> None:
        """Multi-GPU: device addresses (ints) where THIS rank's result rows start inside every other rank's full result
        array (``sab_set_result_peers``); the tetrahedral stream kernel then stores each row there as well.  ``[]`` clears."""
        ptrs = (C.c_void_p * max(len(peer_slot_ptrs), 1))(*[C.c_void_p(int(p)) for p in peer_slot_ptrs])
        nat.check(self.L.sab_set_result_peers(self.h, len(peer_slot_ptrs), ptrs))

    def set_result_multicast(self, multicast_slot_ptr: int) -> None:
        """Multi-GPU with NVLS: ONE multicast address (int; this rank's slot inside the multicast view of the full result
        array, ``sab_set_result_multicast``); each row is stored once and the NVSwitch replicates it to every rank.  0 clears."""
        nat.check(self.L.sab_set_result_multicast(self.h, C.c_void_p(int(multicast_slot_ptr)) if multicast_slot_ptr else None))

    # ------------------------------------------------------------------ exact replay (polyhedral)
    def _lazy_start(self):
        """Hand the per-site caches to the replay kernel as the reference holds them before the first frame."""
        S, smax = self.S, self._slot_atom.shape[1]
        shift = np.zeros((S, smax, 3), dtype=np.int32); cached = np.zeros((S, smax, 3)); has = np.zeros(S, dtype=np.uint8)
        for s, pr in enumerate(self.primed):
            if pr is not None:
                n = len(self.slots[s])
                shift[s, :n] = np.asarray(pr[0]); cached[s, :n] = np.asarray(pr[1]); has[s] = 1
        has_rc = np.array([c is not None for c in self.ref_centres], dtype=np.uint8)
        rc = np.array([c if c is not None else np.zeros(3) for c in self.ref_centres], dtype=np.float64).reshape(S, 3)
        self._upload_priority_tables()
        nat.check(self.L.sab_lazy_set_state(self.h, nat.ptr(shift), nat.ptr(cached), nat.ptr(has), nat.ptr(has_rc),
                                            nat.ptr(np.ascontiguousarray(rc))))
        self._lazy_mode = True

    def _lazy_snapshot(self):
        S, smax = self.S, self._slot_atom.shape[1]
        shift = np.zeros((S, smax, 3), dtype=np.int32); cached = np.zeros((S, smax, 3)); has = np.zeros(S, dtype=np.uint8)
        nat.check(self.L.sab_lazy_get_state(self.h, nat.ptr(shift), nat.ptr(cached), nat.ptr(has), None, None))
        return shift, cached, has

    def lazy_stats(self) -> dict:
        st = np.zeros(4, dtype=np.int64)
        nat.check(self.L.sab_lazy_get_state(self.h, None, None, None, None, nat.ptr(st)))
        return dict(queries=int(st[0]), incremental=int(st[1]), full=int(st[2]), invalidations=int(st[3]))

    def _rehull_first_queries(self, topo_pass: int) -> bool:
        """After a replay attempt: hull every open site that was queried for the first time from the vertex coordinates
        of that query (what ``FaceTopologyCache.__init__`` does, polyhedral_site.py:138-159).  Returns True -- and uploads
        the new triangulations -- when one differs from the triangulation the attempt used."""
        S, smax = self.S, self._slot_atom.shape[1]
        first_frame = np.zeros(S, dtype=np.int64); first_vc = np.zeros((S, smax, 3))
        nat.check(self.L.sab_lazy_get_first_queries(self.h, nat.ptr(first_frame), nat.ptr(first_vc)))
        self._first_queried = (first_frame >= 0) & self._open_topology
        changed = False
        for s in np.nonzero(self._first_queried)[0]:
            n = len(self.slots[s])
            new = tab.hull_simplices(first_vc[s, :n], s)
            if not np.array_equal(new, np.asarray(self.simplices[s])):
                self.simplices[s] = new
                changed = True
        if not changed or topo_pass >= 16:
            return False
        fmax = max(len(f) for f in self.simplices)
        n_face = np.array([len(f) for f in self.simplices], dtype=np.int32)
        simp = np.zeros((S, fmax, 3), dtype=np.int32)
        for s, f in enumerate(self.simplices):
            simp[s, :len(f)] = f
        nat.check(self.L.sab_update_face_topology(self.h, fmax, nat.ptr(n_face), nat.ptr(np.ascontiguousarray(simp))))
        return True

    def _assign_device_lazy(self, d_frac, d_lattice, stride, d_result, append):
        """One batch through k_polyhedral_lazy: final positions and the advanced history, nothing left to resolve or fold."""
        torch = _torch()
        F = int(d_frac.shape[0])
        if not self._pbc_ready:
            lat0 = (d_lattice[0] if stride else d_lattice).cpu().numpy()
            self._init_pbc(d_frac[0].cpu().numpy(), lat0)          # site tables + face topology (hulled from the first frame, G3)
        if not self._lazy_mode:
            self._lazy_start()
        self.sync_history()
        stream = torch.cuda.current_stream(d_frac.device).cuda_stream
        width = max(16, 2 * max((len(t) for t in self.transitions), default=0))
        has_rc = np.array([c is not None for c in self.ref_centres], dtype=np.uint8)
        rc = np.ascontiguousarray(np.array([c if c is not None else np.zeros(3) for c in self.ref_centres], dtype=np.float64).reshape(self.S, 3))
        topo_pass = 0
        while True:
            tr_n = np.zeros(self.S, dtype=np.int32)
            tr_key = np.zeros((self.S, width), dtype=np.int32)
            tr_cnt = np.zeros((self.S, width), dtype=np.int64)
            for s, t in enumerate(self.transitions):
                tr_n[s] = len(t)
                if t:
                    tr_key[s, :len(t)] = list(t.keys()); tr_cnt[s, :len(t)] = list(t.values())
            recent, prev = self.recent.copy(), self.prev.copy()
            snap = self._lazy_snapshot()
            open_now = self._open_topology.any()
            if open_now:                                        # clears the first-query records of the previous attempt
                nat.check(self.L.sab_lazy_set_open_topology(self.h, nat.ptr(np.ascontiguousarray(self._open_topology, dtype=np.uint8))))
            code = self.L.sab_assign_lazy(self.h, d_frac.data_ptr(), d_lattice.data_ptr(), stride, F, d_result.data_ptr(), stream,
                                          nat.ptr(recent), nat.ptr(prev), nat.ptr(tr_n), nat.ptr(tr_key), nat.ptr(tr_cnt), width,
                                          1 if append else 0)
            if code == 3 and width < 4 * self.S:               # SAB_EOVERFLOW: a Counter row outgrew the table; redo wider
                nat.check(self.L.sab_lazy_set_state(self.h, nat.ptr(snap[0]), nat.ptr(snap[1]), nat.ptr(snap[2]), nat.ptr(has_rc), nat.ptr(rc)))
                width *= 2
                continue
            nat.check(code)
            if self._supply_missing_rank_rows():                # lazy ranking: redo the batch with the true rows of the tied anchors
                nat.check(self.L.sab_lazy_set_state(self.h, nat.ptr(snap[0]), nat.ptr(snap[1]), nat.ptr(snap[2]), nat.ptr(has_rc), nat.ptr(rc)))
                continue
            if open_now and self._rehull_first_queries(topo_pass):
                # some site's triangulation differs from the one this attempt used: redo the batch with it.  Sites queried
                # EARLIER than the first wrong one keep their (now confirmed) triangulation, so every pass settles at least
                # the earliest open site -- the loop ends after at most as many passes as there are first-query frames.
                topo_pass += 1
                nat.check(self.L.sab_lazy_set_state(self.h, nat.ptr(snap[0]), nat.ptr(snap[1]), nat.ptr(snap[2]), nat.ptr(has_rc), nat.ptr(rc)))
                continue
            break
        if open_now:                                            # first-queried sites own their topology from now on
            self._open_topology &= ~self._first_queried
        self.recent, self.prev = recent, prev
        self.transitions = [dict(zip(map(int, tr_key[s, :tr_n[s]]), map(int, tr_cnt[s, :tr_n[s]]))) for s in range(self.S)]
        self.frames_done += F
        self.last_info = dict(ambiguous=0, full27=0, launches=1, reinits=0, assign_ns=0, wrap_ns=0, deferred=0, scan_redo=0,
                              peers_written=0, lazy=1, topology_passes=topo_pass)
        self._last_pending = []
        return d_result

    def assign_device(self, d_frac, d_lattice, *, append: bool = True, defer_resolve: bool = False, out=None):
        """Assign a device-resident batch (see :meth:`_assign_device_fast` for the arguments).

        Polyhedral sites: when a guard of the frame-parallel formulation fails (a framework step >= 0.3 invalidates a
        site's cache; the unwrapped excursion reaches 0.3; an unprimed site could pick another image) in a batch that
        starts from the engine's initial state, the batch is redone by the exact replay kernel
        (csrc/k_polyhedral_lazy.cuh) and the engine stays on it.  A violation in a LATER batch cannot be repaired (when each
        site was last queried during the earlier batches is not known) and raises ``GuardError``; set
        ``exact_lazy = True`` before the first frame to stream such trajectories exactly."""
        if self.k != nat.POLYHEDRAL:
            return self._assign_device_fast(d_frac, d_lattice, append=append, defer_resolve=defer_resolve, out=out)
        torch = _torch()
        F = int(d_frac.shape[0])
        auto = self.exact_lazy is None and self.frames_done == 0 and 0 < F < self.LAZY_AUTO_FRAMES and not defer_resolve
        if self.exact_lazy is None and self.frames_done == 0 and self._open_topology.any() and not defer_resolve:
            auto = True                                         # face topology at the first query needs the replay kernel
        if (self.exact_lazy or self._lazy_mode or auto) and F > 0:
            d_lattice = d_lattice.contiguous()
            d_result = out if out is not None else torch.empty((F, self.A), dtype=torch.int32, device=d_frac.device)
            return self._assign_device_lazy(d_frac, d_lattice, 9 if d_lattice.dim() == 3 else 0, d_result, append)
        # snapshot for the redo below -- only a batch that starts from the initial state can be redone, so later batches
        # (the streaming case) pay nothing for it
        start = (self.frames_done, len(self._pending), self.recent.copy(), self.prev.copy(),
                 self._tr_snapshot() if self.frames_done == 0 else None)
        try:
            return self._assign_device_fast(d_frac, d_lattice, append=append, defer_resolve=defer_resolve, out=out)
        except GuardError:
            if start[0] != 0 or not self.strict or defer_resolve or self.exact_lazy is False:
                raise
        # redo from the initial state with the exact replay
        self.frames_done = 0
        del self._pending[start[1]:]
        self.recent, self.prev = start[2], start[3]
        self._tr_restore(start[4])
        d_lattice = d_lattice.contiguous()
        d_result = out if out is not None else torch.empty((F, self.A), dtype=torch.int32, device=d_frac.device)
        return self._assign_device_lazy(d_frac, d_lattice, 9 if d_lattice.dim() == 3 else 0, d_result, append)

    def _assign_device_fast(self, d_frac, d_lattice, *, append: bool = True, defer_resolve: bool = False, out=None):
        """Assign a device-resident batch.  ``d_frac`` (F, N, 3) float64 cuda tensor, ``d_lattice`` (3, 3)
        or (F, 3, 3).  Returns an (F, A) int32 cuda tensor of site list positions (-1 = None).

        ``defer_resolve``: atom-frames in several sites keep their spill codes (<= -2) in the returned tensor
        until :meth:`sync_history` runs the priority resolver over them IN PLACE -- the sharded driver sets the
        history left by the preceding shard first (``distributed.relay_history``).  ``out``: (F, A) int32 cuda tensor
        to write into (default: a new one).  ``last_info["peers_written"]`` tells whether every row also went to the
        peers set by :meth:`set_result_peers`."""
        torch = _torch()
        assert d_frac.is_cuda and d_frac.dtype == torch.float64 and d_frac.is_contiguous()
        F = int(d_frac.shape[0])
        assert tuple(d_frac.shape[1:]) == (self.N, 3), f"expected (F, {self.N}, 3), got {tuple(d_frac.shape)}"
        if not append and F != 1:
            raise ValueError("analyse-without-append is defined for one frame at a time")
        d_lattice = d_lattice.contiguous()
        stride = 9 if d_lattice.dim() == 3 else 0
        if stride:
            assert d_lattice.shape[0] == F
        dev = d_frac.device
        if out is None:
            d_result = torch.empty((F, self.A), dtype=torch.int32, device=dev)
        else:
            assert out.is_cuda and out.dtype == torch.int32 and out.is_contiguous() and tuple(out.shape) == (F, self.A)
            d_result = out
        self._last_pending = []
        if F == 0:
            return d_result
        if self.slots is not None and not self._pbc_ready:
            lat0 = (d_lattice[0] if stride else d_lattice).cpu().numpy()
            self._init_pbc(d_frac[0].cpu().numpy(), lat0)
        if self.k == nat.VORONOI and self.use_cells and self._cells is None and (F >= self.cells_min_frames or self._exhaustive_too_big()) and 64 <= self.S <= 65535:
            self._build_nearest_lists((d_lattice[0] if stride else d_lattice).cpu().numpy())
        if self.k == nat.DYNAMIC_VORONOI and self.use_cells and self._cells is None and (F >= self.cells_min_frames or self._exhaustive_too_big()) and 64 <= self.S <= 65535:
            self._build_dynvor_lists(d_frac, (d_lattice[0] if stride else d_lattice).cpu().numpy())
        if self.k == nat.SPHERICAL and self.use_cells and self._cells is None and (F >= self.cells_min_frames or self._exhaustive_too_big()) and 64 <= self.S <= 65535:
            self._build_sphere_lists((d_lattice[0] if stride else d_lattice).cpu().numpy())
        stream = torch.cuda.current_stream(dev).cuda_stream
        f0 = 0
        total = dict(ambiguous=0, full27=0, launches=0, reinits=0, assign_ns=0, wrap_ns=0, deferred=0, scan_redo=0,
                     peers_written=1)
        span_bits = 0
        while f0 < F:
            nf = F - f0
            sub = d_frac[f0:]
            sub_lat = d_lattice[f0:] if stride else d_lattice
            sub_res = d_result[f0:]
            ws_bytes = int(self.L.sab_workspace_bytes(self.h, nf))
            d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            cap = self._spill_capacity(nf)
            info = np.zeros(12, dtype=np.int64)
            self._maybe_build_cells(sub, nf, d_ws, ws_bytes, stream)
            while True:
                d_spill = torch.empty(cap + 64, dtype=torch.int32, device=dev)   # slack: the resolver prefetches whole records
                code = self.L.sab_assign_device(self.h, sub.data_ptr(), sub_lat.data_ptr(), stride, nf,
                                                sub_res.data_ptr(), d_spill.data_ptr(), cap, d_ws.data_ptr(), ws_bytes,
                                                stream, nat.ptr(info))
                if code == 3 and int(info[1]) > cap:      # spill overflow (nothing committed): grow and redo
                    cap = int(info[1]) + int(info[1]) // 4 + 1024
                    continue
                nat.check(code)
                break
            n_ok = nf if info[3] < 0 else int(info[3])
            if self._cells is not None and info[4] > max(1, nf // 50):
                self._cells_stale = True                 # too many frames left the tube: re-anchor next batch
            total["full27"] += int(info[4]); total["launches"] += int(info[5]); total["deferred"] += int(info[9])
            total["scan_redo"] += int(info[10])
            total["peers_written"] &= int(info[11]) if (info[3] < 0 and f0 == 0) else 0   # one clean pass over the whole batch
            span_bits = max(span_bits, int(info[8]))
            total["assign_ns"] += int(info[6]); total["wrap_ns"] += int(info[7])
            if info[3] >= 0:
                if self.k == nat.POLYHEDRAL:
                    raise GuardError(
                        f"frame {self.frames_done + f0 + n_ok}: a polyhedron vertex moved >= 0.3 fractional units in "
                        "one step (pbc_utils.py:216 invalidates the site's cache and the reference re-derives "
                        "it lazily, which is inherently sequential)")
                nat.check(self.L.sab_commit_pbc(self.h, sub.data_ptr(), nf, n_ok, d_ws.data_ptr(), ws_bytes, stream))
            if n_ok:
                ambiguous = bool(info[0]) and bool((sub_res[:n_ok] <= nat.AMBIGUOUS_BASE).any().item())
                if ambiguous and defer_resolve:
                    self._pending.append(("resolve", sub, n_ok, sub_res, d_spill, stream, append))
                    total["ambiguous"] += int(info[0])
                elif ambiguous:
                    self._resolve(sub, n_ok, sub_res, d_spill, stream, append)
                    total["ambiguous"] += int(info[0])
                else:
                    self._pending.append((sub_res[:n_ok], append))
                    self._last_pending.append((len(self._pending) - 1, f0, n_ok))
            f0 += n_ok
            if info[3] >= 0:
                raw_t = d_frac[f0].cpu().numpy()
                lat_t = (d_lattice[f0] if stride else d_lattice).cpu().numpy()
                self._reinit_groups(raw_t, lat_t)
                total["reinits"] += 1
        self.frames_done += F
        self._check_excursion(span_bits)
        self.last_info = total
        return d_result

    def _spill_capacity(self, nf: int) -> int:
        if self.k == nat.SPHERICAL:
            return int(nf * self.A * 8 + 1024)          # records are multiples of 4 ints (SAB_SPILL_RECORD)
        if self.k == nat.POLYHEDRAL:
            return int(max(nf * self.A // 8, 1) + 1024)
        return 16

    PIPELINE_MIN_FRAMES = 2048      # host batches from this size go through the pipelined C-ABI call (sab_assign_host)
    LAZY_AUTO_FRAMES = 64           # polyhedral engines whose FIRST call is shorter than this stay on the exact replay kernel

    def _snapshot_pbc(self):
        if self.slots is None or not self._pbc_ready:
            return None
        M = len(self.fw_atoms)
        sh = np.zeros_like(self.slot_shift)
        prev_raw = np.zeros((M, 3)); wraps = np.zeros((M, 3), dtype=np.int32); exc = np.zeros((M, 3, 2))
        nat.check(self.L.sab_get_pbc_state(self.h, nat.ptr(sh), nat.ptr(prev_raw), nat.ptr(wraps), nat.ptr(exc)))
        return sh, prev_raw, wraps, exc, self.frames_done, len(self._pending)

    def _restore_pbc(self, snap):
        if snap is None:
            return
        sh, prev_raw, wraps, exc, done, n_pending = snap
        nat.check(self.L.sab_set_pbc_state(self.h, nat.ptr(sh), nat.ptr(prev_raw), nat.ptr(wraps), nat.ptr(exc)))
        if done == 0 and getattr(self, "_legacy_flags", None) is not None and self._legacy_flags.any():
            nat.check(self.L.sab_set_legacy_init(self.h, nat.ptr(self._legacy_flags), 0))     # the first frame is still ahead
        self.frames_done = done
        del self._pending[n_pending:]

    def assign(self, frac: np.ndarray, lattice: np.ndarray, *, append: bool = True, chunk_frames: int = 32768) -> np.ndarray:
        """Assign a host trajectory (F, N, 3): staging -> device -> (F, A) int32 numpy.

        Long batches take the pipelined host-buffer call of the C ABI (``sab_assign_host``: chunked, H2D copies overlapped
        with the kernels; pinned input -- e.g. from ``io.read_xdatcar(pinned=True)`` -- is copied asynchronously).  A batch that
        call cannot finish (atom-frames inside several sites need the priority resolver; a framework step >= 0.3 needs the
        re-initialising path) is redone from the saved PBC state by the chunked path below, with identical results."""
        torch = _torch()
        frac = np.ascontiguousarray(frac, dtype=np.float64)
        lattice = np.ascontiguousarray(lattice, dtype=np.float64)
        F = frac.shape[0]
        call_start = (self.frames_done, len(self._pending), self.recent.copy(), self.prev.copy(),
                      self._tr_snapshot() if self.k == nat.POLYHEDRAL and self.frames_done == 0 else None)
        fast_ok = not (self.k == nat.POLYHEDRAL and (self.exact_lazy or self._lazy_mode or
                                                     (self.exact_lazy is None and self.frames_done == 0 and
                                                      (F < self.LAZY_AUTO_FRAMES or self._open_topology.any()))))
        if fast_ok and append and F >= self.PIPELINE_MIN_FRAMES and self.k != nat.SPHERICAL:
            if self.slots is not None and not self._pbc_ready:
                self._init_pbc(frac[0], lattice if lattice.ndim == 2 else lattice[0])
            snap = self._snapshot_pbc()
            try:
                return self.assign_host_pipelined(frac, lattice)
            except (nat.NativeError, GuardError):
                if self.slots is not None and snap is None:
                    raise
                self._restore_pbc(snap)
        out = np.empty((F, self.A), dtype=np.int32)
        dev = torch.device("cuda", self.device)

        def chunks():
            for f0 in range(0, F, chunk_frames):
                f1 = min(F, f0 + chunk_frames)
                d_frac = torch.from_numpy(frac[f0:f1]).to(dev)
                lat = lattice if lattice.ndim == 2 else lattice[f0:f1]
                d_lat = torch.from_numpy(np.ascontiguousarray(lat)).to(dev)
                out[f0:f1] = self.assign_device(d_frac, d_lat, append=append).cpu().numpy()
                for idx, start, n in self._last_pending:      # history folds lazily from the host copy
                    self._pending[idx] = (out[f0 + start:f0 + start + n], self._pending[idx][1])
        try:
            chunks()
        except GuardError:
            # a guard failed in a later chunk of a call that started from the initial state: the exact replay redoes the call
            if self.k != nat.POLYHEDRAL or call_start[0] != 0 or not self.strict or self._lazy_mode or self.exact_lazy is False:
                raise
            self.frames_done = 0
            del self._pending[call_start[1]:]
            self.recent, self.prev = call_start[2], call_start[3]
            self._tr_restore(call_start[4])
            self._lazy_start()
            chunks()
        return out

    def assign_host_pipelined(self, frac: np.ndarray, lattice: np.ndarray, out: np.ndarray | None = None,
                              chunk_frames: int = 8192) -> np.ndarray:
        """The C-ABI host-buffer call (``sab_assign_host``): chunked, copy/compute overlapped.

        Frames whose atoms sit in several sites at once need the resolver and are not handled by
        this entry point (it raises); use :meth:`assign` for such trajectories."""
        assert frac.dtype == np.float64 and frac.flags["C_CONTIGUOUS"]
        lattice = np.ascontiguousarray(lattice, dtype=np.float64)
        F = frac.shape[0]
        if out is None:
            out = np.empty((F, self.A), dtype=np.int32)
        if self.slots is not None and not self._pbc_ready and F:
            self._init_pbc(frac[0], lattice if lattice.ndim == 2 else lattice[0])
        if self.k == nat.POLYHEDRAL and self.use_cells and (self._cells is None or self._cells_stale) and F >= self.cells_min_frames:
            torch = _torch()                              # anchors from a strided sample spanning the trajectory
            stride = max(1, F // 8192)
            dev = torch.device("cuda", self.device)
            d_s = torch.from_numpy(np.ascontiguousarray(frac[::stride][:8192])).to(dev)
            ws_bytes = int(self.L.sab_workspace_bytes(self.h, d_s.shape[0]))
            d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            self._maybe_build_cells(d_s, int(d_s.shape[0]), d_ws, ws_bytes, torch.cuda.current_stream(dev).cuda_stream,
                                    margin=0.65)
            del d_s, d_ws
        lat0 = lattice if lattice.ndim == 2 else lattice[0]
        if self.use_cells and self._cells is None and (F >= self.cells_min_frames or self._exhaustive_too_big()) and 64 <= self.S <= 65535:
            if self.k == nat.VORONOI:
                self._build_nearest_lists(lat0)
            elif self.k == nat.SPHERICAL:
                self._build_sphere_lists(lat0)
            elif self.k == nat.DYNAMIC_VORONOI:
                torch = _torch()
                stride = max(1, F // 256)
                d_s = torch.from_numpy(np.ascontiguousarray(frac[::stride][:256])).to(torch.device("cuda", self.device))
                self._build_dynvor_lists(d_s, lat0)
                del d_s
        info = np.zeros(12, dtype=np.int64)
        d_mirror = None
        if F * self.A >= DEVICE_FOLD_MIN:                 # the history fold runs on the GPU: keep a device copy of the rows
            torch = _torch()
            d_mirror = torch.empty((F, self.A), dtype=torch.int32, device=torch.device("cuda", self.device))
            nat.check(self.L.sab_set_result_mirror(self.h, d_mirror.data_ptr()))
        try:
            nat.check(self.L.sab_assign_host(self.h, frac.ctypes.data, nat.ptr(lattice), 9 if lattice.ndim == 3 else 0,
                                             F, out.ctypes.data, chunk_frames, nat.ptr(info)))
        finally:
            if d_mirror is not None:
                nat.check(self.L.sab_set_result_mirror(self.h, None))
        if info[3] >= 0:
            raise GuardError(f"frame {int(info[3])}: a framework step >= 0.3 fractional units needs the re-initialising "
                             "path; use AssignmentEngine.assign")
        self._check_excursion(int(info[8]))
        if self._cells is not None and info[4] > max(1, F // 50):
            self._cells_stale = True
        self._pending.append((out if d_mirror is None else d_mirror, True))
        self.frames_done += F
        self.last_info = dict(ambiguous=0, full27=int(info[4]), launches=int(info[5]), reinits=0,
                              assign_ns=int(info[6]), wrap_ns=int(info[7]), deferred=int(info[9]))
        return out

    def dynvor_centres(self, d_frac):
        """(F, S, 3) float64 cuda tensor of dynamic-Voronoi centres for a batch (state not advanced)."""
        torch = _torch()
        F = int(d_frac.shape[0])
        if not self._pbc_ready:
            raise nat.NativeError("dynvor_centres needs an initialised PBC state (assign a frame first)")
        ws_bytes = int(self.L.sab_workspace_bytes(self.h, F))
        d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d_frac.device)
        d_c = torch.empty((F, self.S, 3), dtype=torch.float64, device=d_frac.device)
        nat.check(self.L.sab_dynvor_centres(self.h, d_frac.data_ptr(), F, d_c.data_ptr(), d_ws.data_ptr(), ws_bytes,
                                            torch.cuda.current_stream(d_frac.device).cuda_stream))
        return d_c
