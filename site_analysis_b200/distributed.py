"""Multi-GPU sharding of a trajectory over the ranks of one node (one process per GPU).

Frames are independent given the PBC image shifts, so the trajectory is cut into contiguous
frame ranges, one per rank.  Two things cross rank boundaries:

* the **wrap carry**: the cumulative wrap count of every framework atom at the start of a shard is
  the sum of the wrap totals of all earlier shards (reference state: ``_pbc_image_shifts`` of
  ``polyhedral_site.py:377-416`` / ``_CentreGroup.pbc_shifts`` of
  ``dynamic_voronoi_site_collection.py:38-105``, here a per-atom prefix scan).  Each rank needs the
  last frame of its predecessor (a halo of ``M x 3`` doubles) and the per-rank totals
  (``M x 3`` int32): two tiny all-gathers;
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


def gather_assignments(local_rows, group, equal_shards: bool = False):
    """All-gather the per-rank (F_local, A) int32 results into (sum F_local, A) on every rank.
    Shards may differ in length by one frame; they are padded for the collective and trimmed
    (``equal_shards=True`` skips the length exchange when the caller knows all shards are equal)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if equal_shards:
        out = torch.empty((world * local_rows.shape[0], local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
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


def assign_sharded(engine, d_frac_local, d_lattice_local, first_frame_global, first_lattice_global, group,
                   equal_shards: bool = False):
    """Assign this rank's contiguous shard with the correct PBC carry and return the FULL
    (F_total, A) int32 result (device tensor) on every rank.

    ``first_frame_global`` / ``first_lattice_global``: the trajectory's very first frame (N, 3) and
    lattice (3, 3) as numpy, identical on every rank -- the structure from which unprimed sites take
    their initial image shifts (polyhedral_site.py:410-416)."""
    import torch
    import torch.distributed as dist
    from . import _native as nat
    rank = dist.get_rank(group)
    dev = d_frac_local.device
    if engine.slots is not None:
        if not engine._pbc_ready:
            engine._init_pbc(np.asarray(first_frame_global, dtype=np.float64), np.asarray(first_lattice_global, dtype=np.float64))
        M = len(engine.fw_atoms)
        if getattr(engine, "_initial_pbc", None) is None:        # state at the start of the (global) trajectory
            sh = np.zeros_like(engine.slot_shift)
            p0 = np.zeros((M, 3)); w0 = np.zeros((M, 3), dtype=np.int32); e0 = np.zeros((M, 3, 2))
            nat.check(engine.L.sab_get_pbc_state(engine.h, nat.ptr(sh), nat.ptr(p0), nat.ptr(w0), nat.ptr(e0)))
            engine._initial_pbc = (sh, p0, w0, e0)
        sh, p0, w0, e0 = engine._initial_pbc
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
    local = engine.assign_device(d_frac_local, d_lattice_local)
    if engine.last_info.get("ambiguous", 0) > 0:
        raise NotImplementedError("atom-frames contained by several sites in a sharded run need the gathered resolver "
                                  "(run the affected trajectory on one GPU)")
    engine._pending = []
    if engine.slots is not None and engine.k == nat.POLYHEDRAL:
        # the scan formulation's guard holds over the WHOLE trajectory: combine the shards' excursions
        M = len(engine.fw_atoms)
        exc = np.zeros((M, 3, 2))
        nat.check(engine.L.sab_get_pbc_state(engine.h, None, None, None, nat.ptr(exc)))
        parts = _all_gather(torch.as_tensor(exc, device=dev), group)
        lo = torch.stack([x[..., 0] for x in parts]).amin(dim=0)
        hi = torch.stack([x[..., 1] for x in parts]).amax(dim=0)
        from .engine import EXCURSION_LIMIT, GuardError
        if float((hi - lo).max().item()) >= EXCURSION_LIMIT and engine.strict:
            raise GuardError("a framework atom moved over >= 0.3 fractional units across the sharded trajectory")
    return gather_assignments(local, group, equal_shards)
