"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in site_analysis_b200.distributed:
contiguous sharding, the wrap-carry exchange (halo + exclusive scan of per-shard wrap totals) and
the all-gather that reassembles atoms_trajectory."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from site_analysis_b200 import distributed as sd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _wrap_totals(raw, prev):
    """numpy statement of the per-shard wrap total: sum_t rint(raw_t - raw_{t-1}) (pbc_utils.py:210-217)."""
    chain = np.concatenate([prev[None], raw], axis=0)
    return np.rint(chain[1:] - chain[:-1]).sum(axis=0).astype(np.int32)


def _worker(rank, world, port, F, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)                       # same data on every rank
        M, A = 11, 5
        walk = 0.2 + np.cumsum(rng.normal(scale=0.02, size=(F, M, 3)), axis=0)     # drifts across cell faces
        raw = walk % 1.0
        init_prev = raw[0].copy()
        f0, f1 = sd.shard_bounds(F, world, rank)
        mine = raw[f0:f1]
        halos = sd.exchange_halo(mine[-1], dist.group.WORLD)
        prev = init_prev if rank == 0 else halos[rank - 1]
        assert rank == 0 or np.array_equal(prev, raw[f0 - 1])
        totals = _wrap_totals(mine, prev)
        base, parts = sd.exchange_wrap_totals(totals, dist.group.WORLD)
        # the carry equals the wrap count of the full, unsharded scan up to the shard start
        full = np.rint(np.diff(np.concatenate([init_prev[None], raw]), axis=0)).cumsum(axis=0).astype(np.int32)
        want = full[f0 - 1] if f0 > 0 else np.zeros_like(full[0])
        assert np.array_equal(base, want), (rank, base, want)
        # unwrapped coordinates raw - W are continuous across the shard boundary
        W_first = base + np.rint(mine[0] - prev).astype(np.int32)
        assert np.abs((mine[0] - W_first) - (prev - base)).max() < 0.3
        # polyhedral carry without totals (distributed.carry_from_halo / k_carry_from_halo): bounded vibration
        # (excursion < 0.3, guard G1) around centres on and off the cell faces, so coordinates flip 0 <-> 1
        centre = rng.choice([0.0, 0.999, 0.5, 0.27, 1e-4], size=(1, M, 3))
        t = np.arange(F)[:, None, None]
        vib = (centre + 0.12 * np.sin(0.31 * t + rng.uniform(0, 6.28, size=(1, M, 3))) + rng.normal(scale=0.008, size=(F, M, 3))) % 1.0
        full_v = np.rint(np.diff(np.concatenate([vib[:1], vib]), axis=0)).cumsum(axis=0).astype(np.int32)
        assert np.ptp(vib - full_v, axis=0).max() < 0.3
        halos_v = sd.exchange_halo(vib[f0:f1][-1], dist.group.WORLD)
        if rank > 0:
            W, u = sd.carry_from_halo(halos_v[rank - 1], vib[0])        # anchor0 = unwrapped first frame (W = 0 there)
            assert np.array_equal(W, full_v[f0 - 1]), (rank, W, full_v[f0 - 1])
            assert np.array_equal(u, vib[f0 - 1] - full_v[f0 - 1])
        # result all-gather (ragged shards)
        rows = torch.arange(f0, f1, dtype=torch.int32)[:, None].repeat(1, A) * 10 + torch.arange(A, dtype=torch.int32)
        full_rows = sd.gather_assignments(rows, dist.group.WORLD)
        want_rows = torch.arange(F, dtype=torch.int32)[:, None].repeat(1, A) * 10 + torch.arange(A, dtype=torch.int32)
        assert torch.equal(full_rows, want_rows)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


class _HistoryOnly:
    """The part of AssignmentEngine that relay_history touches, with the host fold as the 'resolver'."""

    def __init__(self, A, S):
        self.recent = -np.ones((A, 2), dtype=np.int32)
        self.prev = np.full(A, -2, dtype=np.int32)
        self.transitions = [dict() for _ in range(S)]
        self._pending = []

    def sync_history(self):
        from site_analysis_b200.engine import fold_history
        pending, self._pending = self._pending, []
        for rows, append in pending:
            fold_history(self.recent, self.prev, self.transitions, rows, append)


def _relay_worker(rank, world, port, F, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        A, S = 9, 6
        rows = rng.integers(-1, S, size=(F, A)).astype(np.int32)
        rows[1:][rng.random((F - 1, A)) < 0.6] = -3                 # mostly "stay": copy the frame before
        for f in range(1, F):
            rows[f] = np.where(rows[f] == -3, rows[f - 1], rows[f])
        whole = _HistoryOnly(A, S)
        whole._pending = [(rows, True)]
        whole.sync_history()
        f0, f1 = sd.shard_bounds(F, world, rank)
        mine = _HistoryOnly(A, S)
        mine._pending = [(rows[f0:(f0 + f1) // 2], True), (rows[(f0 + f1) // 2:f1], True)]
        sd.relay_history(mine, dist.group.WORLD)
        assert np.array_equal(mine.recent, whole.recent) and np.array_equal(mine.prev, whole.prev)
        assert [list(t.items()) for t in mine.transitions] == [list(t.items()) for t in whole.transitions]   # insertion order too
        open(os.path.join(out_dir, f"relay{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_history_relay_world2(tmp_path):
    """Shard-after-shard hand-over of (recent sites, previous entry, transition Counters) == one sequential fold."""
    port = _free_port()
    mp.spawn(_relay_worker, args=(2, port, 41, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "relay0").exists() and (tmp_path / "relay1").exists()


@pytest.mark.parametrize("F", [64, 37])
def test_wrap_carry_and_gather_world2(tmp_path, F):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, F, 123, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_bounds_cover_all_frames():
    for F in (0, 1, 7, 8, 125000, 1000003):
        for world in (1, 2, 4, 8):
            spans = [sd.shard_bounds(F, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _halo_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, F = 7, 5
        frames = torch.arange(world * F * N * 3, dtype=torch.float64).reshape(world * F, N, 3)      # same on every rank
        mine = frames[rank * F:(rank + 1) * F]
        halo = sd.exchange_halo_frames(mine, dist.group.WORLD)
        want = frames[rank * F - 1] if rank > 0 else mine[-1]
        assert torch.equal(halo, want), rank
    finally:
        dist.destroy_process_group()


def test_halo_frames_come_from_the_preceding_shard_world2():
    """distributed.exchange_halo_frames: rank r gets the last frame of rank r - 1's shard (input distribution of the sharded
    polyhedral path; with it assign_sharded issues no halo collective per batch)."""
    mp.spawn(_halo_worker, args=(2, _free_port()), nprocs=2, join=True)


class _DoneEvent:
    def synchronize(self):
        pass


class _FakeEngine:
    strict = True


def test_deferred_guard_records_raise_what_the_immediate_check_raises():
    """distributed.finish_sharded: the records assign_sharded(deferred_check=True) leaves behind (per rank: excursion
    min / max of every framework coordinate, 'rows not stored into the peers' flag, 'ambiguous atom-frames' flag)."""
    from site_analysis_b200.engine import GuardError
    M, world = 3, 2

    def record(span=0.1, unfused=0.0, ambiguous=0.0):
        parts = torch.zeros((world, 6 * M + 2), dtype=torch.float64)
        exc = parts[:, :-2].reshape(world, M, 3, 2)
        exc[..., 0] = 0.2
        exc[..., 1] = 0.25
        exc[1, 2, 1, 1] = 0.2 + span                          # one coordinate of one rank carries the span
        parts[0, -2] = unfused
        parts[1, -1] = ambiguous
        return (_DoneEvent(), parts, True, M, True)

    e = _FakeEngine()
    e._sharded_checks = [record(), record(span=0.29)]
    sd.finish_sharded(e)                                      # within the guard: nothing raised, list consumed
    assert e._sharded_checks == []
    e._sharded_checks = [record(span=0.31)]
    with pytest.raises(GuardError):
        sd.finish_sharded(e)
    e._sharded_checks = [record(ambiguous=1.0)]
    with pytest.raises(RuntimeError, match="several sites"):
        sd.finish_sharded(e)
    e._sharded_checks = [record(unfused=1.0)]
    with pytest.raises(RuntimeError, match="peers"):
        sd.finish_sharded(e)
