"""Post-processing of finished assignment histories on the GPU.

The reference answers ``Site.trajectory``, ``Site.residence_times`` (reference site.py:241-352),
``Site.average_occupation`` (site.py:225-239) and the transition tables (trajectory.py:188-321) by walking
per-object Python lists: O(frames x sites) work per question.  Here one pass of ``sab_fold_result``
(``csrc/k_fold.cuh``, C ABI in ``include/sab200.h``) over the int32 result array of a batch gives

* the *runs* -- (site, atom, first frame, length) of every maximal stretch an atom spends in a site, sorted by
  (site, atom, first frame) -- i.e. the run-length form of every ``Site.trajectory`` at once, and
* the *pairs* -- (from site, to site, count, first occurrence) of every frame-to-frame change, i.e. the increments
  of ``Site.transitions`` (site_collection.py:300-302) in Counter insertion order,

and everything else is arithmetic on those short arrays.  No CPU fallback: folding needs the CUDA backend.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def _next_pow2(n: int) -> int:
    return 1 << max(int(n) - 1, 1).bit_length()


def fold_rows(rows, n_sites: int, prev=None, want_pairs: bool = True, device: int | None = None,
              want_recent: bool = False, want_runs: bool = True):
    """Fold a result array on the GPU.

    Args:
        rows: (F, A) int32 of site LIST POSITIONS (``-1`` = no site), numpy array or CUDA tensor.
        n_sites: number of sites S (positions are ``0 .. S-1``).
        prev: optional (A,) int32, the atoms' trajectory entries before the first row (-2 none, -1 None, position).
        want_pairs: also return the transition pairs.
        want_recent: also return ``recent`` = (last, other), two (A,) int32 arrays: the site of each atom's last run in
            the batch and of its last run in a different site (-1 = none), cf. ``Atom._recent_sites`` (atom.py:181-192).

    Returns:
        dict with ``runs`` = (site, atom, start, length) numpy arrays sorted by (site, atom, start),
        ``pairs`` = (src, dst, count, first) sorted by first occurrence (or ``None``), ``info`` = dict(kernel_ns, launches).
    """
    import torch
    if not torch.cuda.is_available():
        raise nat.NativeError("site_analysis_b200 needs a CUDA device: there is no CPU fallback")
    L = nat.lib()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    if isinstance(rows, np.ndarray):
        d_rows = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32)).to(dev)
    else:
        d_rows = rows.to(device=dev, dtype=torch.int32).contiguous()
    if d_rows.dim() != 2:
        raise ValueError(f"expected a (frames, atoms) array, got shape {tuple(d_rows.shape)}")
    F, A = int(d_rows.shape[0]), int(d_rows.shape[1])
    empty = (np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.int64), np.empty(0, np.int64))
    if F == 0 or A == 0:
        none = np.full(A, -1, dtype=np.int32)
        return dict(runs=empty, pairs=empty if want_pairs else None, recent=(none, none.copy()) if want_recent else None,
                    info=dict(kernel_ns=0, launches=0))
    d_prev = None
    if prev is not None:
        d_prev = torch.from_numpy(np.ascontiguousarray(prev, dtype=np.int32)).to(dev)
        if d_prev.numel() != A:
            raise ValueError("prev must hold one entry per atom")
    slots = min(_next_pow2(max(1024, 64 * int(n_sites))), 1 << 26) if want_pairs else 0
    run_cap = max(1024, F * A // 16 + 2 * A)
    stream = torch.cuda.current_stream(dev).cuda_stream
    info = np.zeros(6, dtype=np.int64)
    with torch.cuda.device(dev):
        for attempt in range(4):
            ws_bytes = L.sab_fold_workspace_bytes(F, A, int(n_sites), run_cap, slots)
            if ws_bytes < 0:
                raise nat.NativeError(L.sab_last_error().decode())
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            r_site = torch.empty(run_cap, dtype=torch.int32, device=dev)
            r_atom = torch.empty(run_cap, dtype=torch.int32, device=dev)
            r_start = torch.empty(run_cap, dtype=torch.int64, device=dev)
            r_len = torch.empty(run_cap, dtype=torch.int64, device=dev)
            p_src = torch.empty(max(slots, 1), dtype=torch.int32, device=dev)
            p_dst = torch.empty(max(slots, 1), dtype=torch.int32, device=dev)
            p_cnt = torch.empty(max(slots, 1), dtype=torch.int64, device=dev)
            p_first = torch.empty(max(slots, 1), dtype=torch.int64, device=dev)
            rec = torch.empty((2, A), dtype=torch.int64, device=dev) if want_recent else None
            rc = L.sab_fold_result(dev.index, d_rows.data_ptr(), F, A, int(n_sites),
                                   None if d_prev is None else d_prev.data_ptr(), run_cap, r_site.data_ptr(),
                                   r_atom.data_ptr(), r_start.data_ptr(), r_len.data_ptr(), slots, slots,
                                   p_src.data_ptr(), p_dst.data_ptr(), p_cnt.data_ptr(), p_first.data_ptr(),
                                   None if rec is None else rec[0].data_ptr(), None if rec is None else rec[1].data_ptr(),
                                   ws.data_ptr(), ws_bytes, C.c_void_p(stream), nat.ptr(info))
            if rc == 0:
                break
            if rc != 3:                                         # SAB_EOVERFLOW: retry with room
                raise nat.NativeError(L.sab_last_error().decode())
            if int(info[0]) > run_cap:
                run_cap = int(info[0])
            else:
                if slots >= 1 << 30:
                    raise nat.NativeError(L.sab_last_error().decode())
                slots *= 8
        else:
            raise nat.NativeError(L.sab_last_error().decode())
        R, P = int(info[0]), int(info[1])
        runs = tuple(t[:R].cpu().numpy() for t in (r_site, r_atom, r_start, r_len)) if want_runs else None   # the history fold needs no runs on the host
        pairs = None
        if want_pairs:
            src, dst, cnt, first = (t[:P].cpu().numpy() for t in (p_src, p_dst, p_cnt, p_first))
            o = np.argsort(first, kind="stable")
            pairs = (src[o], dst[o], cnt[o], first[o])
        recent = None
        if want_recent:
            k = rec.cpu().numpy()
            recent = tuple(np.where(k[i] > 0, k[i] & 0xffffff, -1).astype(np.int32) for i in range(2))
    return dict(runs=runs, pairs=pairs, recent=recent, info=dict(kernel_ns=int(info[3]), launches=int(info[2])))


class BlockFolds:
    """Lazily folded form of ONE batch of a trajectory (the rows ``trajectory_from_arrays`` appended).

    ``rows_index`` holds ``Site.index`` values (-1 = ``None``); the fold itself works on list positions."""

    def __init__(self, rows_index: np.ndarray, atom_index: np.ndarray, site_index: np.ndarray):
        self.rows_index = rows_index
        self.atom_index = np.asarray(atom_index)
        self.site_index = np.asarray(site_index)
        self.n_frames = int(rows_index.shape[0])
        self._runs = None
        self._site_off = None
        self.kernel_ns = 0
        self.frac = None                             # (F, N, 3) coordinates of the batch, kept only for lazy Site.points

    def _fold(self):
        if self._runs is None:
            S = len(self.site_index)
            top = int(self.site_index.max()) + 1
            inv = np.full(top + 1, -1, dtype=np.int32)          # inv[-1] (the slot after the largest index) stays -1
            inv[self.site_index] = np.arange(S, dtype=np.int32)
            if int(self.rows_index.max(initial=-1)) >= top or int(self.rows_index.min(initial=0)) < -1:
                raise ValueError("the result rows hold values that are neither -1 nor a site index of this trajectory")
            pos = inv[self.rows_index]
            if np.any((pos < 0) & (self.rows_index >= 0)):
                raise ValueError("the result rows hold a site index that is not in this trajectory")
            out = fold_rows(pos, S, want_pairs=False)
            self._runs = out["runs"]
            self.kernel_ns = out["info"]["kernel_ns"]
            self._site_off = np.searchsorted(self._runs[0], np.arange(S + 1))
        return self._runs

    def site_runs(self, site_pos: int):
        """(atom column, start, length) of the runs of one site, sorted by (atom column, start)."""
        site, atom, start, length = self._fold()
        a, b = int(self._site_off[site_pos]), int(self._site_off[site_pos + 1])
        return atom[a:b], start[a:b], length[a:b]


def site_runs_over_blocks(blocks, site_pos_of_block):
    """Concatenate one site's runs over consecutive batches (frame numbers made global, stretches that continue across
    a batch boundary joined).  Returns (atom_index, start, end) sorted by (atom INDEX, start) and the total frame count."""
    atoms, starts, ends = [], [], []
    f0 = 0
    for blk, pos in zip(blocks, site_pos_of_block):
        col, st, ln = blk.site_runs(pos)
        atoms.append(blk.atom_index[col].astype(np.int64))
        starts.append(st + f0)
        ends.append(st + ln + f0)
        f0 += blk.n_frames
    atom = np.concatenate(atoms) if atoms else np.empty(0, np.int64)
    start = np.concatenate(starts) if starts else np.empty(0, np.int64)
    end = np.concatenate(ends) if ends else np.empty(0, np.int64)
    o = np.lexsort((start, atom))
    atom, start, end = atom[o], start[o], end[o]
    if len(atom) > 1:                                           # a stretch continuing across a batch boundary is one run
        joined = np.concatenate(([False], (atom[1:] == atom[:-1]) & (start[1:] == end[:-1])))
        head = ~joined
        idx = np.flatnonzero(head)
        atom, start, end = atom[idx], start[idx], np.maximum.reduceat(end, idx)
    return atom, start, end, f0


def residence_times_from_runs(atom, start, end, n_frames: int, filter_length: int = 0,
                              include_edge_runs: bool = False) -> tuple[int, ...]:
    """Reference site.py:241-352 on the run-length form: runs sorted by (atom index, start); gaps of at most
    ``filter_length`` frames between two runs of the same atom are filled; runs touching frame 0 or the last frame
    are dropped unless ``include_edge_runs``."""
    if len(atom) == 0:
        return ()
    if filter_length > 0 and len(atom) > 1:
        fill = np.concatenate(([False], (atom[1:] == atom[:-1]) & (start[1:] - end[:-1] <= filter_length)))
        idx = np.flatnonzero(~fill)
        start, end = start[idx], np.maximum.reduceat(end, idx)
    if not include_edge_runs:
        keep = (start != 0) & (end != n_frames)
        start, end = start[keep], end[keep]
    return tuple(int(v) for v in (end - start))


def occupied_frames_from_runs(start, end) -> int:
    """Number of frames covered by the union of the intervals [start, end) (site.py:238: frames with >= 1 atom)."""
    if len(start) == 0:
        return 0
    o = np.argsort(start, kind="stable")
    s, e = start[o], end[o]
    reach = np.maximum.accumulate(e)
    before = np.concatenate(([s[0]], reach[:-1]))
    return int(np.clip(e - np.maximum(s, before), 0, None).sum())
