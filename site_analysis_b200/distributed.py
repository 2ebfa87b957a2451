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
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous(), group=group)
    return out


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


def _carry_by_anchor(engine, d_frac_local, group):
    """Polyhedral sites (guard G1 makes it exact, see ``carry_from_halo``): one tiny all-gather of the halo frames,
    then the carry is set ON THE DEVICE (sab_set_carry_device) -- no pass over the data, no host round trip."""
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
    last_raw = d_frac_local[-1].index_select(0, st["fw"])
    halos = _all_gather(last_raw, group)
    halo = st["p0"] if rank == 0 else halos[rank - 1]
    nat.check(engine.L.sab_set_carry_device(engine.h, halo.data_ptr(), st["anchor0"].data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream))
    st["keep"] = (halos, last_raw)                           # alive until the stream has consumed them


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


def assign_sharded(engine, d_frac_local, d_lattice_local, first_frame_global, first_lattice_global, group,
                   equal_shards: bool = False, async_gather: bool = False):
    """Assign this rank's contiguous shard with the correct PBC carry and return the FULL
    (F_total, A) int32 result (device tensor) on every rank.

    ``first_frame_global`` / ``first_lattice_global``: the trajectory's very first frame (N, 3) and
    lattice (3, 3) as numpy, identical on every rank -- the structure from which unprimed sites take
    their initial image shifts (polyhedral_site.py:410-416).

    Cross-rank traffic per call: the halo all-gather (before the kernels), ONE small all-gather after them that
    carries the excursion record of guard G1 together with the "some atom-frame sits in several sites" flag (one
    host synchronisation), and the result all-gather.  ``async_gather=True`` (equal shards) returns
    ``(full, work)`` with the result all-gather still in flight: a caller that streams a long trajectory through
    in batches calls ``work.wait()`` when it reads ``full``, and the next batch's kernels overlap with the gather."""
    import torch
    import torch.distributed as dist
    from . import _native as nat
    from .engine import EXCURSION_LIMIT, GuardError
    dev = d_frac_local.device
    if engine.slots is not None:
        _initial_state(engine, first_frame_global, first_lattice_global)
        if engine.k == nat.POLYHEDRAL:
            _carry_by_anchor(engine, d_frac_local, group)
        else:
            _carry_by_totals(engine, d_frac_local, group)
    local = engine.assign_device(d_frac_local, d_lattice_local, defer_resolve=True)
    guarded = engine.slots is not None and engine.k == nat.POLYHEDRAL
    overlapping = engine.k in (nat.SPHERICAL, nat.POLYHEDRAL)      # sites may overlap: does ANY shard need the priority order?
    if guarded or overlapping:
        M = len(engine.fw_atoms) if guarded else 0
        note = torch.zeros(6 * M + 1, dtype=torch.float64, device=dev)
        if guarded:         # guard G1 holds over the WHOLE trajectory: the shards' excursions travel with the flag
            nat.check(engine.L.sab_get_excursion_device(engine.h, note.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        if engine.last_info.get("ambiguous", 0) > 0:
            note[-1] = 1.0
        parts = torch.stack(_all_gather(note, group))
        summary = torch.zeros(2, dtype=torch.float64, device=parts.device)
        summary[0] = parts[:, -1].max()
        if guarded:
            exc = parts[:, :-1].reshape(parts.shape[0], M, 3, 2)
            summary[1] = (exc[..., 1].amax(dim=0) - exc[..., 0].amin(dim=0)).max()
        any_ambiguous, span = summary.cpu().tolist()                # the one host synchronisation after the kernels
        if any_ambiguous:
            relay_history(engine, group)                           # resolves `local` in place, shard after shard
        if guarded and span >= EXCURSION_LIMIT and engine.strict:
            raise GuardError("a framework atom moved over >= 0.3 fractional units across the sharded trajectory")
    engine._pending = []
    return gather_assignments(local, group, equal_shards, async_op=async_gather)
