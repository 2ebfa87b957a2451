"""Multi-GPU sharding of a trajectory over the ranks of one node (one process per GPU).

Frames are independent given the PBC image shifts, so the trajectory is cut into contiguous
frame ranges, one per rank.  Two things cross rank boundaries:

* the **wrap carry**: the cumulative wrap count of every framework atom at the start of a shard is
  the sum of the wrap totals of all earlier shards (reference state: ``_pbc_image_shifts`` of
  ``polyhedral_site.py:377-416`` / ``_CentreGroup.pbc_shifts`` of
  ``dynamic_voronoi_site_collection.py:38-105``, here a per-atom prefix scan).  Each rank needs the
  last frame of its predecessor (a halo of ``M x 3`` doubles).  Polyhedral sites: the carry is
  ``rint(halo - anchor0)`` with ``anchor0`` the unwrapped first frame of the trajectory, set on the
  device -- exact under the excursion guard that this path enforces anyway (``carry_from_halo``).
  Dynamic Voronoi (no such guard): one extra pass for the per-shard totals (``M x 3`` int32) and a
  second tiny all-gather;
* the **result**: one all-gather of ``int32[F_local, A]`` reassembles ``atoms_trajectory``
  (``trajectory.py:375-383``) on every rank.

There is no other collective on the data path.  All functions take a ``torch.distributed`` process
group and work with the ``nccl`` backend (GPU tensors) and the ``gloo`` backend (CPU tensors; used by
the CPU tests of this logic).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_frames: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced frame range [f0, f1) of ``rank`` (earlier ranks take the remainder)."""
    base, rem = divmod(int(n_frames), int(world_size))
    f0 = rank * base + min(rank, rem)
    return f0, f0 + base + (1 if rank < rem else 0)


def _all_gather(t, group):
    """List over ranks of every rank's ``t`` (views of ONE gathered tensor: a single collective, no per-rank copies)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out, t, group=group)
    except (RuntimeError, NotImplementedError):            # backends without the flat form (old gloo)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        return parts
    return list(out.unbind(0))


_RESULT_GROUPS: dict = {}


def _result_group(group):
    """A second communicator over the same ranks for the big result all-gather: NCCL runs the collectives of one
    communicator in order, so on the data-path group the next batch's halo exchange (which its kernels wait for) would
    queue behind the previous batch's result all-gather and nothing would overlap.  Collective call."""
    import torch.distributed as dist
    key = id(group)
    if key not in _RESULT_GROUPS:
        ranks = dist.get_process_group_ranks(group)
        _RESULT_GROUPS[key] = dist.new_group(ranks=ranks, backend=dist.get_backend(group))
    return _RESULT_GROUPS[key]


def exchange_halo(last_raw, group, device=None):
    """All-gather the raw framework coordinates of every rank's last frame.

    ``last_raw``: (M, 3) float64 numpy.  Returns the list over ranks; rank r uses entry r - 1 as the
    "previous frame" of its first frame (rank 0 keeps the state it was initialised with)."""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(last_raw), dtype=torch.float64, device=device)
    return [x.cpu().numpy() for x in _all_gather(t, group)]


def exchange_wrap_totals(totals, group, device=None):
    """All-gather the per-shard wrap totals; returns (exclusive prefix for this rank, all totals).

    ``totals``: (M, 3) int32 numpy = sum over this shard's frames of rint(raw_t - raw_{t-1}), with the
    predecessor's last frame as t - 1 of the first frame."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(totals), dtype=torch.int32, device=device)
    parts = [x.cpu().numpy() for x in _all_gather(t, group)]
    rank = dist.get_rank(group)
    base = np.zeros_like(parts[0])
    for q in range(rank):
        base = base + parts[q]
    return base, parts


def gather_assignments(local_rows, group, equal_shards: bool = False, async_op: bool = False):
    """All-gather the per-rank (F_local, A) int32 results into (sum F_local, A) on every rank.
    Shards may differ in length by one frame; they are padded for the collective and trimmed
    (``equal_shards=True`` skips the length exchange when the caller knows all shards are equal).

    ``async_op`` (equal shards only): returns ``(out, work)`` without making the current stream wait for the
    collective, so the next batch's kernels overlap with it; ``work.wait()`` orders the current stream after it."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if equal_shards:
        out = torch.empty((world * local_rows.shape[0], local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
        work = dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    if async_op:
        raise ValueError("async_op needs equal_shards=True")
    n_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    counts = [int(x.item()) for x in _all_gather(n_local, group)]
    fmax = max(counts)
    if local_rows.shape[0] < fmax:
        pad = torch.full((fmax - local_rows.shape[0], local_rows.shape[1]), -1, dtype=local_rows.dtype,
                         device=local_rows.device)
        local_rows = torch.cat([local_rows, pad], dim=0)
    out = torch.empty((world * fmax, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
    if all(c == fmax for c in counts):
        return out
    return torch.cat([out[r * fmax:r * fmax + c] for r, c in enumerate(counts)], dim=0)


class PeerResultBuffers:
    """Two full result arrays int32[world * F_local, A] per rank in symmetric (peer-mapped) memory.

    With these the result all-gather is FUSED into the assignment kernel: the tetrahedral stream kernel stores every row
    it produces into its own slot of its own array and, through NVLink peer addresses, into the same slot of every other
    rank's array (``sab_set_result_peers``), frame by frame while it computes.  The small guard/flag all-gather that
    follows the kernels anyway is the completion barrier.  Two arrays alternate, so that a rank still reading batch k is
    not overwritten by a neighbour that has already started batch k + 1 (nobody starts batch k + 2 before everybody has
    finished the kernels of batch k + 1).  Collective constructor (rendezvous over ``group``)."""

    def __init__(self, n_frames_local: int, n_mobile: int, group, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.F, self.A = int(n_frames_local), int(n_mobile)
        self.buffers, self.ptrs, self.mc_ptrs = [], [], []
        import os
        for _ in range(2):
            t = symm_mem.empty((self.world * self.F, self.A), dtype=torch.int32, device=device)
            h = symm_mem.rendezvous(t, group)
            self.buffers.append(t)
            self.ptrs.append([int(p) for p in h.buffer_ptrs])
            mc = 0
            try:                                  # NVLS multicast view of the same buffers (0 / absent when the fabric has none)
                # Measured (profiles/README.md): on 4 B200s the multimem.st form (a strong system-scope store) makes the stream
                # kernel 10 % slower than three plain unicast stores (0.663 vs 0.604 ms); on 8 the seven unicast stores are
                # NVLink-egress bound (672 MB per rank and step: kernel 1.23 ms against 0.55 ms on one GPU), which is where one
                # multicast store per row pays.  Default: multicast from 8 ranks on; SAB_MULTICAST=1 / 0 forces it on / off.
                want = os.environ.get("SAB_MULTICAST")
                if self.world > 2 and (want == "1" or (want != "0" and self.world >= 8)):
                    mc = int(getattr(h, "multicast_ptr", 0) or 0)
            except Exception:
                mc = 0
            self.mc_ptrs.append(mc)
        self.multicast = all(p != 0 for p in self.mc_ptrs)
        self._handles = None
        self.turn = 0

    def next(self):
        """(full array of this turn, its local slot, peer slot addresses for ``engine.set_result_peers``)."""
        k = self.turn
        self.turn ^= 1
        full = self.buffers[k]
        slot_bytes = self.rank * self.F * self.A * 4
        peers = [self.ptrs[k][q] + slot_bytes for q in range(self.world) if q != self.rank]
        self.mc_slot = self.mc_ptrs[k] + slot_bytes if getattr(self, "multicast", False) else 0
        return full, full[self.rank * self.F:(self.rank + 1) * self.F], peers


def carry_from_halo(halo, anchor0):
    """Wrap counts and unwrapped coordinates at the frame before a shard, WITHOUT the per-shard totals:
    ``W = rint(halo - anchor0)``, ``u = halo - W`` (the host statement of ``k_carry_from_halo``).

    ``anchor0`` = unwrapped coordinates of the trajectory's first frame.  The reference's caches count
    ``rint(raw_t - raw_{t-1})`` frame by frame (pbc_utils.py:210-217); as long as every unwrapped coordinate
    stays within 0.3 of every other over the whole trajectory (guard G1, checked on the gathered
    excursions), ``|u(halo) - anchor0| < 0.5`` and the two are the same integer."""
    halo = np.asarray(halo, dtype=np.float64)
    W = np.rint(halo - np.asarray(anchor0, dtype=np.float64)).astype(np.int32)
    return W, halo - W


def _initial_state(engine, first_frame_global, first_lattice_global):
    from . import _native as nat
    if not engine._pbc_ready:
        engine._init_pbc(np.asarray(first_frame_global, dtype=np.float64), np.asarray(first_lattice_global, dtype=np.float64))
    if getattr(engine, "_initial_pbc", None) is None:            # state at the start of the (global) trajectory
        M = len(engine.fw_atoms)
        sh = np.zeros_like(engine.slot_shift)
        p0 = np.zeros((M, 3)); w0 = np.zeros((M, 3), dtype=np.int32); e0 = np.zeros((M, 3, 2))
        nat.check(engine.L.sab_get_pbc_state(engine.h, nat.ptr(sh), nat.ptr(p0), nat.ptr(w0), nat.ptr(e0)))
        engine._initial_pbc = (sh, p0, w0, e0)
    return engine._initial_pbc


def _carry_by_totals(engine, d_frac_local, group):
    """Exact for any data: halo exchange, one extra pass over the shard for its wrap totals, exclusive prefix."""
    import torch
    import torch.distributed as dist
    from . import _native as nat
    rank = dist.get_rank(group)
    dev = d_frac_local.device
    sh, p0, w0, e0 = engine._initial_pbc
    M = len(engine.fw_atoms)
    fw = torch.as_tensor(engine.fw_atoms.astype(np.int64), device=dev)
    last_raw = d_frac_local[-1].index_select(0, fw).cpu().numpy()
    halos = exchange_halo(last_raw, group, device=dev)
    prev_raw = p0 if rank == 0 else np.ascontiguousarray(halos[rank - 1])
    nat.check(engine.L.sab_set_pbc_state(engine.h, nat.ptr(sh), nat.ptr(prev_raw), nat.ptr(w0), nat.ptr(e0)))
    F = int(d_frac_local.shape[0])
    ws_bytes = int(engine.L.sab_workspace_bytes(engine.h, F))
    d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    totals = np.zeros((M, 3), dtype=np.int32); lr = np.zeros((M, 3))
    nat.check(engine.L.sab_wrap_totals(engine.h, d_frac_local.data_ptr(), F, d_ws.data_ptr(), ws_bytes,
                                       torch.cuda.current_stream(dev).cuda_stream, nat.ptr(totals), nat.ptr(lr)))
    base, _ = exchange_wrap_totals(totals, group, device=dev)
    if rank > 0:
        wraps = np.ascontiguousarray(w0 + base.astype(np.int32))
        u0 = prev_raw - wraps                                # the halo frame in the global unwrapped frame
        exc = np.ascontiguousarray(np.stack([u0, u0], axis=-1))
        nat.check(engine.L.sab_set_pbc_state(engine.h, nat.ptr(sh), nat.ptr(prev_raw), nat.ptr(wraps), nat.ptr(exc)))


def _carry_by_anchor(engine, d_frac_local, group, halo_frame=None):
    """Polyhedral sites (guard G1 makes it exact, see ``carry_from_halo``): one tiny all-gather of the halo frames,
    then the carry is set ON THE DEVICE (sab_set_carry_device) -- no pass over the data, no host round trip.

    ``halo_frame``: the frame that precedes this rank's shard, (N, 3) float64 on the shard's device, when the caller has
    it already (a loader that hands rank r the frames [f0 - 1, f1) gets it for free): then NO collective is issued here.
    Ignored on rank 0, whose predecessor is the state the engine was initialised with."""
    import torch
    import torch.distributed as dist
    from . import _native as nat
    rank = dist.get_rank(group)
    dev = d_frac_local.device
    st = getattr(engine, "_shard_dev", None)
    if st is None:
        sh, p0, w0, e0 = engine._initial_pbc
        st = engine._shard_dev = dict(
            fw=torch.as_tensor(engine.fw_atoms.astype(np.int64), device=dev),
            p0=torch.as_tensor(np.ascontiguousarray(p0), device=dev),
            anchor0=torch.as_tensor(np.ascontiguousarray(p0 - w0), device=dev))
    if halo_frame is not None:
        halo = st["p0"] if rank == 0 else halo_frame.index_select(0, st["fw"]).contiguous()
        st["keep"] = (halo,)
    else:
        last_raw = d_frac_local[-1].index_select(0, st["fw"])
        halos = _all_gather(last_raw, group)
        halo = st["p0"] if rank == 0 else halos[rank - 1]
        st["keep"] = (halos, last_raw)                       # alive until the stream has consumed them
    nat.check(engine.L.sab_set_carry_device(engine.h, halo.data_ptr(), st["anchor0"].data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream))


def relay_history(engine, group) -> None:
    """Sequential hand-over of the history the priority order depends on (``Atom._recent_sites`` atom.py:55, the
    previous trajectory entry, ``Site.transitions`` site.py:71; read by site_collection.py:113-182).

    Rank after rank: take the history the preceding shard ended with, fold / resolve this shard's batches in frame
    order (``engine.sync_history``), pass the result on.  Each shard's frames stay on their GPU; only
    (A x 3 ints + the transition Counters) travel, world_size - 1 times.  After the call every engine holds the
    history at the end of the WHOLE trajectory and every rank's result tensor is final."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    for r in range(world):
        if rank == r:
            engine.sync_history()
            box = [(engine.recent, engine.prev, engine.transitions)]
        else:
            box = [None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, r), group=group)
        if rank != r:
            recent, prev, transitions = box[0]
            if rank > r:                                   # not yet processed: start from the predecessor's state
                engine.recent, engine.prev = np.array(recent, dtype=np.int32), np.array(prev, dtype=np.int32)
                engine.transitions = [dict(t) for t in transitions]
            elif r == world - 1:                           # done earlier: adopt the final state
                engine._pending = []
                engine.recent, engine.prev = np.array(recent, dtype=np.int32), np.array(prev, dtype=np.int32)
                engine.transitions = [dict(t) for t in transitions]


def exchange_halo_frames(d_frac_local, group):
    """One-off helper for callers whose loader did not keep the frame before each shard: all-gather of the shards' last
    frames; returns the (N, 3) frame preceding THIS rank's shard (rank 0: its own last frame, unused).  Input
    distribution, not part of the per-batch data path: pass the result as ``assign_sharded(halo_frame=...)``."""
    import torch.distributed as dist
    last = _all_gather(d_frac_local[-1].contiguous(), group)
    return last[max(dist.get_rank(group) - 1, 0)].clone()


def finish_sharded(engine) -> None:
    """Evaluate the guard / flag records of every ``assign_sharded(..., deferred_check=True)`` call made so far (one
    host synchronisation for all of them).  Raises what the immediate form would have raised: ``GuardError`` for a
    framework excursion >= 0.3 over the sharded trajectory, ``RuntimeError`` when some rank met an atom-frame in several
    sites or could not store its rows into the peers (such a batch must be redone with ``deferred_check=False``, whose
    history relay / all-gather fallback handles it)."""
    from .engine import EXCURSION_LIMIT, GuardError
    checks, engine._sharded_checks = getattr(engine, "_sharded_checks", []), []
    engine._barrier_events = []                # ev.synchronize() below covers them: every peer's rows of every batch have arrived
    pool = getattr(engine, "_pinned_pool", None)
    for ev, host, guarded, M, fused in checks:
        ev.synchronize()
        parts = host.numpy().copy()
        if pool is not None and len(pool) < 64:
            pool.append(host)
        if float(parts[:, -1].max()) > 0:
            raise RuntimeError("assign_sharded(deferred_check=True): some shard holds atom-frames contained in several sites; "
                               "the priority order needs the history relay -- redo the batch with deferred_check=False")
        if fused and float(parts[:, -2].max()) > 0:
            raise RuntimeError("assign_sharded(deferred_check=True): some rank could not store its rows into the peers "
                               "(batch redone by the scan kernels); redo the batch with deferred_check=False")
        if guarded:
            exc = parts[:, :-2].reshape(parts.shape[0], M, 3, 2)
            span = float((exc[..., 1].max(axis=0) - exc[..., 0].min(axis=0)).max())
            if span >= EXCURSION_LIMIT and engine.strict:
                raise GuardError("a framework atom moved over >= 0.3 fractional units across the sharded trajectory")


def assign_sharded(engine, d_frac_local, d_lattice_local, first_frame_global, first_lattice_global, group,
                   equal_shards: bool = False, async_gather: bool = False, peer_buffers: PeerResultBuffers | None = None,
                   halo_frame=None, deferred_check: bool = False, keep_history: bool = False):
    """Assign this rank's contiguous shard with the correct PBC carry and return the FULL
    (F_total, A) int32 result (device tensor) on every rank.

    ``first_frame_global`` / ``first_lattice_global``: the trajectory's very first frame (N, 3) and
    lattice (3, 3) as numpy, identical on every rank -- the structure from which unprimed sites take
    their initial image shifts (polyhedral_site.py:410-416).

    Cross-rank traffic per call: the halo all-gather (before the kernels), ONE small all-gather after them that
    carries the excursion record of guard G1 together with the "some atom-frame sits in several sites" flag (one
    host synchronisation), and the result rows.  The rows travel either

    * fused into the kernel (``peer_buffers``, equal shards, tetrahedral sites): stored straight into every rank's
      full array over NVLink peer memory while the kernel runs; the small all-gather is the completion barrier and no
      result collective is issued at all (unless some rank reports overlapping sites or a redone batch -- then this call
      falls back to the all-gather below), or
    * by one NCCL all-gather of ``int32[F_local, A]``; ``async_gather=True`` (equal shards) returns ``(full, work)``
      with it still in flight on its own communicator.

    ``halo_frame``: the (N, 3) frame preceding this rank's shard, if the caller has it (``exchange_halo_frames``): the halo
    all-gather is then skipped.  ``deferred_check=True``: the guard / flag all-gather is still issued on the stream (it is
    the completion barrier of the peer stores) but its result is copied to pinned host memory asynchronously and examined
    by ``finish_sharded(engine)`` later -- no host synchronisation per call, so consecutive batches queue back to back.
    ``keep_history=True``: every rank folds the gathered full result into its engine's history (recent sites, previous
    entries, transition Counters; device fold) on every call, so that the engines stay usable for a later call that does
    meet overlapping sites.  Without it a call that leaves no ambiguity drops its history and marks the engine's history
    stale; a later ambiguous call on a stale engine raises instead of resolving against the wrong Counters."""
    import torch
    import torch.distributed as dist
    from . import _native as nat
    from .engine import EXCURSION_LIMIT, GuardError
    dev = d_frac_local.device
    if engine.slots is not None:
        _initial_state(engine, first_frame_global, first_lattice_global)
        if engine.k == nat.POLYHEDRAL:
            _carry_by_anchor(engine, d_frac_local, group, halo_frame)
        else:
            _carry_by_totals(engine, d_frac_local, group)
    full = out = None
    if peer_buffers is not None:
        if not equal_shards or int(d_frac_local.shape[0]) != peer_buffers.F:
            raise ValueError("peer_buffers need equal shards of the size they were made for")
        evs = getattr(engine, "_barrier_events", None)
        if evs:                                     # deferred barriers on the side stream: this call reuses the array of the call
            if len(evs) >= 2:                       # before the previous one -- every rank must have finished that batch's kernels
                torch.cuda.current_stream(dev).wait_event(evs[-2])
            engine._barrier_events = evs[-2:]
        full, out, peers = peer_buffers.next()
        if getattr(peer_buffers, "multicast", False):     # one multimem.st per row; the switch replicates it (N > 2, NVLS present)
            engine.set_result_multicast(peer_buffers.mc_slot)
        else:
            engine.set_result_peers(peers)
    try:
        local = engine.assign_device(d_frac_local, d_lattice_local, defer_resolve=True, out=out)
    finally:
        if peer_buffers is not None:
            engine.set_result_peers([])
    fused = peer_buffers is not None and bool(engine.last_info.get("peers_written"))
    guarded = engine.slots is not None and engine.k == nat.POLYHEDRAL
    overlapping = engine.k in (nat.SPHERICAL, nat.POLYHEDRAL)      # sites may overlap: does ANY shard need the priority order?
    any_ambiguous = 0.0
    all_fused = fused
    if guarded or overlapping or peer_buffers is not None:
        M = len(engine.fw_atoms) if guarded else 0
        note = torch.zeros(6 * M + 2, dtype=torch.float64, device=dev)
        if guarded:         # guard G1 holds over the WHOLE trajectory: the shards' excursions travel with the flags
            nat.check(engine.L.sab_get_excursion_device(engine.h, note.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        if engine.last_info.get("ambiguous", 0) > 0:
            note[-1] = 1.0
        if not fused:
            note[-2] = 1.0
        if deferred_check:
            # Examined later (finish_sharded): no host synchronisation here.  SAB_BARRIER_STREAM=1 additionally moves the
            # all-gather (with the fused result exchange it is only the completion barrier of the peer stores) to a side
            # stream behind this batch's kernels, the kernels of batch k + 2 waiting on the device for the barrier of batch k
            # (what keeps the two alternating result arrays safe).  Measured on 4 B200s (profiles/README.md) it does not pay:
            # the persistent stream kernel fills every SM, NCCL's kernel runs behind it either way, and short runs end with
            # a queue of barriers (1.48 vs 1.28 ms per step over 5 steps, 1.19 over 20) -- so it is off by default.
            world_n = dist.get_world_size(group)
            shape = (world_n, int(note.numel()))
            pool = getattr(engine, "_pinned_pool", None)           # page-locked allocation is slow (~0.2 ms): reuse the buffers
            if pool is None:                                       # first deferred call: a batch of buffers at once
                pool = engine._pinned_pool = [torch.empty(shape, dtype=note.dtype, pin_memory=True) for _ in range(16)]
            host = pool.pop() if pool and tuple(pool[-1].shape) == shape else torch.empty(shape, dtype=note.dtype, pin_memory=True)
            main = torch.cuda.current_stream(dev)
            side = main
            import os
            use_side = peer_buffers is not None and os.environ.get("SAB_BARRIER_STREAM") == "1"
            if use_side:
                side = getattr(engine, "_barrier_stream", None)
                if side is None:
                    side = engine._barrier_stream = torch.cuda.Stream(device=dev)
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                note.record_stream(side)
            with torch.cuda.stream(side):
                parts = torch.stack(_all_gather(note, group))      # also orders every peer's row stores before our reads
                host.copy_(parts, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            if not hasattr(engine, "_sharded_checks"):
                engine._sharded_checks = []
            engine._sharded_checks.append((ev, host, guarded, M, fused))
            if use_side:
                engine._barrier_events = (getattr(engine, "_barrier_events", None) or []) + [ev]
            engine._pending = []
            engine._history_stale = True
            if peer_buffers is not None:
                return (full, None) if async_gather else full
            return gather_assignments(local, _result_group(group) if async_gather else group, equal_shards, async_op=async_gather)
        parts = torch.stack(_all_gather(note, group))              # also orders every peer's row stores before our reads
        parts = parts.cpu().numpy()                                # the one host synchronisation after the kernels (a few KB)
        any_ambiguous, any_unfused, span = float(parts[:, -1].max()), float(parts[:, -2].max()), 0.0
        if guarded:
            exc = parts[:, :-2].reshape(parts.shape[0], M, 3, 2)
            span = float((exc[..., 1].max(axis=0) - exc[..., 0].min(axis=0)).max())
        all_fused = fused and not any_unfused and not any_ambiguous
        if any_ambiguous:
            if getattr(engine, "_history_stale", False):
                raise RuntimeError("assign_sharded: this batch needs the priority order (atom-frames in several sites) but the "
                                   "history of earlier sharded batches was dropped; call assign_sharded(keep_history=True) "
                                   "from the first batch on, or engine.reset() before a new trajectory")
            relay_history(engine, group)                           # resolves `local` in place, shard after shard
        if guarded and span >= EXCURSION_LIMIT and engine.strict:
            raise GuardError("a framework atom moved over >= 0.3 fractional units across the sharded trajectory")
    folded = bool(any_ambiguous)                                   # relay_history left every engine with the whole history
    engine._pending = []
    if peer_buffers is not None:
        if not all_fused:
            dist.all_gather_into_tensor(full, local.clone(), group=group)        # some rank could not fuse: plain all-gather
        out_full = full
    elif async_gather and not keep_history:
        if not folded:
            engine._history_stale = True
        return gather_assignments(local, _result_group(group), equal_shards, async_op=True)
    else:
        out_full = gather_assignments(local, group, equal_shards)
    if not folded:
        if keep_history:       # every rank folds the WHOLE gathered batch: all engines end with the same history, as after a relay
            from .engine import fold_history_device
            torch.cuda.current_stream(dev).synchronize()
            fold_history_device(engine.recent, engine.prev, engine, out_full, engine.S)
        else:
            engine._history_stale = True
    return (out_full, None) if async_gather else out_full
