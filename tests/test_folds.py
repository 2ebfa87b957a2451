"""Post-processing folds: residence times, average occupation, per-site trajectories and
transition counts from a finished int32[F][A] result array.

CPU part: the plain-Python oracle (oracle/folds_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden_folds.py), and the host-side run arithmetic of site_analysis_b200.folds against the same fixtures.
GPU part: sab_fold_result (csrc/k_fold.cuh) through the C ABI against the oracle -- integer work, bit-exact."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

FOLD_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(str(GOLDEN / "folds_*.npz")))


def _load(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


def _golden_rt(case, k, s):
    off = case[f"rt{k}_off"]
    return tuple(int(v) for v in case[f"rt{k}_flat"][off[s]:off[s + 1]])


def _golden_site_traj(case, s):
    F = case["rows"].shape[0]
    off, flat = case["site_traj_off"], case["site_traj_flat"]
    return [flat[off[s * F + f]:off[s * F + f + 1]].tolist() for f in range(F)]


def _golden_transitions(case):
    out = {}
    for src, dst, cnt in case["transitions"].tolist():
        out.setdefault(src, {})[dst] = cnt
    return out


def _pos_rows(case):
    inv = {int(v): i for i, v in enumerate(case["site_index"])}
    return np.vectorize(lambda v: inv.get(int(v), -1))(case["rows"]).astype(np.int32)


def test_fixture_list_is_complete():
    assert len(FOLD_CASES) >= 6


@pytest.mark.parametrize("name", FOLD_CASES)
def test_oracle_reproduces_reference_folds(name):
    import folds_oracle as fo
    case = _load(name)
    st = fo.site_trajectories(case["rows"], case["atom_index"], case["site_index"])
    for s, idx in enumerate(case["site_index"].tolist()):
        assert st[idx] == _golden_site_traj(case, s)
        for k, (fl, edge) in enumerate(case["params"].tolist()):
            assert fo.residence_times(st[idx], fl, bool(edge)) == _golden_rt(case, k, s)
        want = case["average_occupation"][s]
        assert fo.average_occupation(st[idx]) == want
    prev = None if np.all(case["prev"] == -2) else case["prev"]
    got, want = fo.transitions(case["rows"], prev), _golden_transitions(case)
    assert got == want and all(list(got[k]) == list(want[k]) for k in got)


def test_oracle_docstring_examples():
    """reference site.py:268-290."""
    import folds_oracle as fo
    assert fo.residence_times([[], [1], [1], [1], []]) == (3,)
    assert fo.residence_times([[1], [1], [1], []]) == ()
    assert fo.residence_times([[1], [1], [1], []], include_edge_runs=True) == (3,)
    assert fo.residence_times([[], [1], [], [1], []]) == (1, 1)
    assert fo.residence_times([[], [1], [], [1], []], filter_length=1) == (3,)
    with pytest.raises(ValueError):
        fo.residence_times([[1]], filter_length=-1)
    with pytest.raises(TypeError):
        fo.residence_times([[1]], filter_length=1.5)


def _blocks_from_oracle_runs(case, cuts):
    """BlockFolds objects whose device fold is replaced by the oracle's run extraction (host-logic test)."""
    import folds_oracle as fo
    from site_analysis_b200.folds import BlockFolds
    pos = _pos_rows(case)
    S = len(case["site_index"])
    blocks = []
    edges = [0] + list(cuts) + [pos.shape[0]]
    for a, b in zip(edges[:-1], edges[1:]):
        blk = BlockFolds(case["rows"][a:b], case["atom_index"], case["site_index"])
        blk._runs = fo.runs_from_rows(pos[a:b])
        blk._site_off = np.searchsorted(blk._runs[0], np.arange(S + 1))
        blocks.append(blk)
    return blocks


@pytest.mark.parametrize("name", FOLD_CASES)
@pytest.mark.parametrize("n_cuts", [0, 1, 3])
def test_run_arithmetic_matches_reference(name, n_cuts):
    """Site.residence_times / average_occupation / trajectory from the run-length form, also when the history arrived
    in several batches (a stretch continuing across a batch boundary is ONE run)."""
    import site_analysis_b200 as sab
    case = _load(name)
    F = case["rows"].shape[0]
    rng = np.random.default_rng(5)
    cuts = sorted(rng.choice(np.arange(1, F), size=min(n_cuts, F - 1), replace=False).tolist())
    blocks = _blocks_from_oracle_runs(case, cuts)
    for s, idx in enumerate(case["site_index"].tolist()):
        site = sab.VoronoiSite(np.zeros(3))
        site.index = idx
        site._lazy_blocks = [(b, s) for b in blocks]
        for k, (fl, edge) in enumerate(case["params"].tolist()):
            assert site.residence_times(filter_length=fl, include_edge_runs=bool(edge)) == _golden_rt(case, k, s)
        assert site.average_occupation == case["average_occupation"][s]
        assert site._lazy_blocks                     # answered without materialising the lists
        assert site.trajectory == _golden_site_traj(case, s)
        assert not site._lazy_blocks
        for k, (fl, edge) in enumerate(case["params"].tolist()):   # and again from the list form
            assert site.residence_times(filter_length=fl, include_edge_runs=bool(edge)) == _golden_rt(case, k, s)
        assert site.average_occupation == case["average_occupation"][s]


def test_residence_times_argument_checks():
    import site_analysis_b200 as sab
    site = sab.VoronoiSite(np.zeros(3))
    assert site.residence_times() == () and site.average_occupation is None
    with pytest.raises(ValueError):
        site.residence_times(filter_length=-1)
    with pytest.raises(TypeError):
        site.residence_times(filter_length=2.0)
    site.trajectory = [[], [4], [4, 9], [9], []]
    assert site.residence_times() == (2, 2)
    assert site.residence_times(include_edge_runs=True) == (2, 2)
    site.trajectory = [[4], [], [4]]
    assert site.residence_times() == () and site.residence_times(1) == ()
    assert site.residence_times(1, include_edge_runs=True) == (3,)


def test_transition_table_behaves_like_the_reference():
    """reference transition_table.py:18-234 / tests/test_transition_table.py."""
    from site_analysis_b200 import TransitionTable
    t = TransitionTable(keys=(3, 1, 2), matrix=np.array([[0, 5, 1], [2, 0, 0], [0, 7, 0]]))
    assert t.keys == (3, 1, 2) and t.get(3, 1) == 5 and t.get(2, 1) == 7
    assert isinstance(t.get(3, 1), int)
    assert t.to_dict()[1] == {3: 2, 1: 0, 2: 0}
    r = t.reorder([1, 2, 3])
    assert r.keys == (1, 2, 3) and r.get(3, 1) == 5 and r.matrix.tolist() == [[0, 0, 2], [7, 0, 0], [5, 1, 0]]
    f = t.filter([2, 3])
    assert f.matrix.tolist() == [[0, 0], [1, 0]] and t.filter([]).matrix.shape == (0, 0)
    assert t == TransitionTable((3, 1, 2), t.matrix) and t != r
    assert str(TransitionTable((), np.empty((0, 0)))) == ""
    assert str(TransitionTable(("A", "B"), np.array([[1, 20], [3, 4]]))) == "     A   B\n A   1  20\n B   3   4"
    assert "<th>a&lt;b</th>" in TransitionTable(("a<b",), np.array([[0.5]]))._repr_html_()
    assert "0.500" in str(TransitionTable(("x",), np.array([[0.5]])))
    assert repr(t) == "TransitionTable(keys=(3, 1, 2), shape=3x3)"
    with pytest.raises(KeyError):
        t.get(9, 1)
    with pytest.raises(ValueError, match="permutation"):
        t.reorder([1, 2])
    with pytest.raises(ValueError, match="unknown keys"):
        t.filter([1, 8])
    with pytest.raises(ValueError, match="duplicates"):
        t.filter([1, 1])
    with pytest.raises(ValueError):
        TransitionTable((1, 2), np.zeros((2, 3)))
    with pytest.raises(ValueError):
        TransitionTable((1, 1), np.zeros((2, 2)))
    with pytest.raises(ValueError):
        TransitionTable((1,), np.array([["a"]]))
    with pytest.raises(AttributeError):
        t.keys = (1,)
    with pytest.raises(ValueError):
        t.matrix[0, 0] = 3


def test_trajectory_transition_tables():
    """reference trajectory.py:188-321 (known answers in the style of the reference's tests/test_trajectory.py)."""
    import site_analysis_b200 as sab
    sab.Site.reset_index()
    sites = [sab.VoronoiSite(np.zeros(3), label=l) for l in ("A", "B", "A", None)]
    t = sab.Trajectory(sites=sites, atoms=[sab.Atom(0)])
    sites[0].transitions.update({1: 3, 2: 1})
    sites[1].transitions.update({0: 2})
    sites[2].transitions.update({3: 4})
    c = t.transition_counts_by_site()
    assert c.keys == (0, 1, 2, 3) and c.matrix.tolist() == [[0, 3, 1, 0], [2, 0, 0, 0], [0, 0, 0, 4], [0, 0, 0, 0]]
    assert t.transition_counts_by_site(keys=[3, 2, 1, 0]).get(0, 1) == 3
    p = t.transition_probabilities_by_site()
    assert p.matrix[0].tolist() == [0.0, 0.75, 0.25, 0.0] and p.matrix[3].tolist() == [0.0] * 4
    with pytest.warns(UserWarning, match="4 transition"):
        lab = t.transition_counts_by_label()
    assert lab.keys == ("A", "B") and lab.matrix.tolist() == [[1, 3], [2, 0]]
    with pytest.warns(UserWarning):
        assert t.transition_probabilities_by_label().get("A", "B") == 0.75
    sites[1].transitions[99] = 1
    with pytest.raises(ValueError, match="unknown site index 99"):
        t.transition_counts_by_site()


# ------------------------------------------------------------------------------------------------ GPU
gpu = [pytest.mark.gpu, pytest.mark.timeout(180)]


def _check_fold_against_oracle(pos, S, prev=None):
    import folds_oracle as fo
    import site_analysis_b200 as sab
    out = sab.fold_rows(pos, S, prev=prev)
    want = fo.runs_from_rows(pos)
    for g, w in zip(out["runs"], want):
        assert g.dtype == w.dtype and np.array_equal(g, w)
    wt = fo.transitions(pos, prev)
    src, dst, cnt, first = out["pairs"]
    got = {}
    for a, b, c in zip(src.tolist(), dst.tolist(), cnt.tolist()):
        got.setdefault(a, {})[b] = c
    assert got == wt
    # pairs arrive sorted by first occurrence == Counter insertion order (site.py:223 ties)
    assert all(list(got[k]) == list(wt[k]) for k in wt)
    assert np.all(np.diff(first) > 0)
    return out


@pytest.mark.gpu
@pytest.mark.timeout(180)
@pytest.mark.parametrize("name", FOLD_CASES)
def test_gpu_fold_matches_reference_fixture(name):
    case = _load(name)
    prev = None
    if not np.all(case["prev"] == -2):
        inv = {int(v): i for i, v in enumerate(case["site_index"])}
        prev = np.array([inv.get(int(p), int(p)) for p in case["prev"]], dtype=np.int32)
    _check_fold_against_oracle(_pos_rows(case), len(case["site_index"]), prev)


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("F,A,S,stay,none", [(4000, 64, 40, 0.95, 0.02), (3000, 33, 700, 0.5, 0.2), (1, 5, 3, 0.9, 0.1),
                                             (2500, 1, 2, 0.8, 0.1), (700, 192, 1056, 0.98, 0.0)])
def test_gpu_fold_random_histories(F, A, S, stay, none):
    """Includes a history with more runs than the first capacity guess (the EOVERFLOW retry), one frame, one atom."""
    import folds_oracle as fo
    rng = np.random.default_rng(F + A)
    pos = fo.random_rows(rng, F, A, S, stay, none)
    prev = rng.integers(-2, S, size=A).astype(np.int32)
    _check_fold_against_oracle(pos, S, prev)
    _check_fold_against_oracle(pos, S, None)


@pytest.mark.gpu
@pytest.mark.timeout(180)
def test_gpu_fold_rejects_unresolved_entries_and_handles_empty():
    import site_analysis_b200 as sab
    from site_analysis_b200 import _native as nat
    rows = np.array([[0, 1], [-5, 1]], dtype=np.int32)
    with pytest.raises(nat.NativeError, match="unresolved"):
        sab.fold_rows(rows, 3)
    with pytest.raises(nat.NativeError, match="unresolved"):
        sab.fold_rows(np.array([[3]], dtype=np.int32), 3)
    out = sab.fold_rows(np.empty((0, 4), dtype=np.int32), 3)
    assert all(len(a) == 0 for a in out["runs"])
    out = sab.fold_rows(np.full((50, 4), -1, dtype=np.int32), 3)
    assert all(len(a) == 0 for a in out["runs"]) and all(len(a) == 0 for a in out["pairs"])


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["argyrodite_0p_polyhedral", "simple_cubic_spherical", "fuzz_voronoi_1"])
def test_trajectory_post_processing_end_to_end(oracle_mod, name):
    """trajectory_from_arrays (two batches) -> Site.residence_times / average_occupation / trajectory, against the
    plain-Python restatement applied to the reference's own atoms_trajectory of the fixture."""
    import folds_oracle as fo
    from helpers import trajectory_from_case
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case)
    F = case["frac"].shape[0]
    cut = F // 3
    lat = case["lattice"]
    t.trajectory_from_arrays(case["frac"][:cut], lat if lat.ndim == 2 else lat[:cut])
    t.trajectory_from_arrays(case["frac"][cut:], lat if lat.ndim == 2 else lat[cut:])
    want_rows = case["atoms_traj"]
    st = fo.site_trajectories(want_rows, [a.index for a in t.atoms], [s.index for s in t.sites])
    checked = 0
    for site in t.sites:
        traj = st[site.index]
        if not any(traj) and checked > 40:
            continue
        for fl, edge in [(0, False), (2, True), (7, False)]:
            assert site.residence_times(fl, edge) == fo.residence_times(traj, fl, edge)
        assert site.average_occupation == fo.average_occupation(traj)
        checked += 1
    some = [s for s in t.sites if any(st[s.index])][:5] + list(t.sites[:2])
    for site in some:
        assert site.trajectory == st[site.index]
    counts = t.transition_counts_by_site()
    want = fo.transitions(want_rows)
    for src, d in want.items():
        for dst, c in d.items():
            assert counts.get(src, dst) == c
    assert int(counts.matrix.sum()) == sum(sum(d.values()) for d in want.values())


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("F,A,S,stay,none", [(3000, 40, 25, 0.9, 0.05), (800, 192, 1056, 0.97, 0.0), (500, 7, 3, 0.6, 0.5)])
def test_device_history_fold_equals_host_fold(F, A, S, stay, none):
    """engine.fold_history_device (GPU) == engine.fold_history (numpy restatement of site_collection.py:300-310 and
    atom.py:190-192): recent-site pairs, previous entries and transition Counters incl. insertion order, over two
    consecutive batches starting from a non-trivial history."""
    import folds_oracle as fo
    from site_analysis_b200.engine import fold_history, fold_history_device
    rng = np.random.default_rng(F * 7 + A)
    pos = fo.random_rows(rng, F, A, S, stay, none)
    pos[:, 0] = -1                                                   # an atom that is never assigned
    if A > 2:
        pos[:, 1] = 2 % S                                            # an atom that never moves
    state = []
    for fold in (fold_history, fold_history_device):
        recent = np.stack([np.arange(A) % S, (np.arange(A) + 1) % S], axis=1).astype(np.int32)
        recent[::5] = -1
        prev = (np.arange(A) % (S + 2) - 2).astype(np.int32)
        tr = [dict() for _ in range(S)]
        tr[1 % S][0] = 4                                             # existing Counter entries keep their place
        for a, b in ((0, F // 3), (F // 3, F)):
            if fold is fold_history:
                fold(recent, prev, tr, pos[a:b], True)
            else:
                fold(recent, prev, tr, pos[a:b], S)
        state.append((recent, prev, tr))
    (r0, p0, t0), (r1, p1, t1) = state
    assert np.array_equal(r0, r1) and np.array_equal(p0, p1)
    assert t0 == t1 and all(list(x) == list(y) for x, y in zip(t0, t1))
    want = fo.transitions(pos, (np.arange(A) % (S + 2) - 2))
    want.setdefault(1 % S, {})
    got = {s: dict(d) for s, d in enumerate(t1) if d}
    got[1 % S][0] -= 4
    if got[1 % S][0] == 0:
        del got[1 % S][0]
    assert {k: v for k, v in got.items() if v} == {k: v for k, v in want.items() if v}


# Known answers of the reference's own unit tests (reference tests/test_site.py:157-184, 280-470), re-expressed against the
# public API: (site trajectory, residence_times keyword arguments, expected tuple).
_REFERENCE_RESIDENCE_CASES = [
    ([[], [], []], {}, ()),
    ([[1], [1], [1]], dict(include_edge_runs=True), (3,)),
    ([[1], [], [1]], dict(include_edge_runs=True), (1, 1)),
    ([[], [1], [1]], dict(include_edge_runs=True), (2,)),
    ([[1], [1], []], dict(include_edge_runs=True), (2,)),
    ([[1], [1], [2], [2], [2]], dict(include_edge_runs=True), (2, 3)),
    ([[1, 2], [1, 2], [1]], dict(include_edge_runs=True), (3, 2)),
    ([[1]], dict(include_edge_runs=True), (1,)),
    ([[1], [1], [], [1], [1], [1], []], {}, (3,)),
    ([[], [1], [1], [1], [], [1], [1]], {}, (3,)),
    ([[1], [], [1], [1], [1], [], [1]], {}, (3,)),
    ([[1], [1], [1]], {}, ()),
    ([[1], [1], [], [1], [], [1], [1]], {}, (1,)),
    ([[], [1], [], [1], []], dict(filter_length=1, include_edge_runs=True), (3,)),
    ([[], [1], [], [], [1], []], dict(filter_length=1, include_edge_runs=True), (1, 1)),
    ([[], [1], [], [], [1], []], dict(filter_length=2, include_edge_runs=True), (4,)),
    ([[], [1], [1]], dict(filter_length=1, include_edge_runs=True), (2,)),
    ([[1], [1], []], dict(filter_length=1, include_edge_runs=True), (2,)),
    ([[], [1], [], [2], []], dict(filter_length=1, include_edge_runs=True), (1, 1)),
    ([[], [1], [], [1], []], dict(filter_length=0), (1, 1)),
    ([[1], [], [1], [1], [], [], [1]], dict(filter_length=1), ()),
    ([[], [1], [], [1], []], {}, (1, 1)),
    ([[], [1], [], [1], []], dict(filter_length=1), (3,)),
]
_REFERENCE_OCCUPATION_CASES = [([], None), ([[], [], []], 0.0), ([[1], [2], [1, 2]], 1.0), ([[1], [], [2], [], [3]], 0.6),
                               ([[1, 2, 3], [], [4, 5], []], 0.5)]


def _site_from_history(traj, folded):
    """A site whose history is ``traj``, either as the plain list or as one GPU-style folded batch (runs from the oracle)."""
    import folds_oracle as fo
    import site_analysis_b200 as sab
    from site_analysis_b200.folds import BlockFolds
    site = sab.VoronoiSite(np.zeros(3))
    site.index = 7
    if not folded or not traj:
        site.trajectory = [list(ts) for ts in traj]
        return site
    atoms = sorted({a for ts in traj for a in ts}) or [1]
    rows = np.full((len(traj), len(atoms)), -1, dtype=np.int32)
    for f, ts in enumerate(traj):
        for a in ts:
            rows[f, atoms.index(a)] = 0
    blk = BlockFolds(np.where(rows >= 0, 7, -1), np.array(atoms), np.array([7]))
    blk._runs = fo.runs_from_rows(rows)
    blk._site_off = np.searchsorted(blk._runs[0], np.arange(2))
    site._lazy_blocks = [(blk, 0)]
    return site


@pytest.mark.parametrize("folded", [False, True])
def test_reference_unit_test_known_answers(folded):
    for traj, kwargs, want in _REFERENCE_RESIDENCE_CASES:
        assert _site_from_history(traj, folded).residence_times(**kwargs) == want, (traj, kwargs)
    for traj, want in _REFERENCE_OCCUPATION_CASES:
        got = _site_from_history(traj, folded).average_occupation
        assert got == want or (want is not None and abs(got - want) < 1e-12), traj
    site = _site_from_history([[1], [], [1]], folded)
    with pytest.raises(ValueError):
        site.residence_times(filter_length=-1)
    with pytest.raises(TypeError):
        site.residence_times(filter_length=1.5)
    summary = _site_from_history([[1], []], folded).summary()
    assert set(summary) == {"index", "site_type", "average_occupation", "transitions"} and summary["average_occupation"] == 0.5
    empty = _site_from_history([], folded)
    assert "average_occupation" not in empty.summary()
    assert empty.summary(metrics=["index", "average_occupation"])["average_occupation"] is None


def test_device_history_fold_host_logic(monkeypatch):
    """engine.fold_history_device with the device fold answered by the oracle == engine.fold_history, over many random
    histories (the GPU test runs the same comparison through the real kernel on a few)."""
    import folds_oracle as fo
    from site_analysis_b200 import engine, folds

    def fake_fold_rows(rows, n_sites, prev=None, want_pairs=True, device=None, want_recent=False, want_runs=True):
        rows = np.asarray(rows)
        tr = fo.transitions(rows, prev)
        first = {}
        last = np.full(rows.shape[1], -2, dtype=np.int64) if prev is None else np.asarray(prev, dtype=np.int64).copy()
        for f in range(rows.shape[0]):
            for a in range(rows.shape[1]):
                v = int(rows[f, a])
                if v >= 0 and last[a] >= 0 and last[a] != v:
                    first.setdefault((int(last[a]), v), f * rows.shape[1] + a)
            last = rows[f].astype(np.int64)
        items = sorted(first.items(), key=lambda kv: kv[1])
        src = np.array([k[0] for k, _ in items], dtype=np.int32); dst = np.array([k[1] for k, _ in items], dtype=np.int32)
        cnt = np.array([tr[k[0]][k[1]] for k, _ in items], dtype=np.int64)
        lastv, other = np.full(rows.shape[1], -1, dtype=np.int32), np.full(rows.shape[1], -1, dtype=np.int32)
        for a in range(rows.shape[1]):
            col = rows[:, a]
            hit = np.flatnonzero(col >= 0)
            if len(hit):
                lastv[a] = col[hit[-1]]
                diff = np.flatnonzero((col >= 0) & (col != lastv[a]))
                if len(diff):
                    other[a] = col[diff[-1]]
        return dict(runs=fo.runs_from_rows(rows), pairs=(src, dst, cnt, np.array([v for _, v in items], dtype=np.int64)),
                    recent=(lastv, other), info=dict(kernel_ns=0, launches=0))
    monkeypatch.setattr(folds, "fold_rows", fake_fold_rows)
    rng = np.random.default_rng(77)
    for trial in range(25):
        F, A, S = int(rng.integers(1, 60)), int(rng.integers(1, 9)), int(rng.integers(1, 6))
        rows = fo.random_rows(rng, F, A, S, float(rng.uniform(0.3, 0.99)), float(rng.uniform(0, 0.4)))
        state = []
        for device in (False, True):
            recent = rng.integers(-1, S, size=(A, 2)).astype(np.int32) if trial % 2 else np.full((A, 2), -1, dtype=np.int32)
            prev = rng.integers(-2, S, size=A).astype(np.int32)
            tr = [dict() for _ in range(S)]
            if trial % 3 == 0 and S > 1:
                tr[0][S - 1] = 2
            state.append((recent, prev, tr))
        state[1][0][:], state[1][1][:] = state[0][0], state[0][1]
        cut = int(rng.integers(0, F + 1))
        for a, b in ((0, cut), (cut, F)):
            if b > a:
                engine.fold_history(state[0][0], state[0][1], state[0][2], rows[a:b], True)
                engine.fold_history_device(state[1][0], state[1][1], state[1][2], rows[a:b], S)
        assert np.array_equal(state[0][0], state[1][0]) and np.array_equal(state[0][1], state[1][1]), trial
        assert state[0][2] == state[1][2] and all(list(x) == list(y) for x, y in zip(state[0][2], state[1][2])), trial


def test_lazy_site_points_equal_eager_points():
    """Site.points of a long batch (kept as runs + a reference to the coordinates) == the reference's eager appends
    (site_collection.py:308: frames in order, atoms in list order, atom.frac_coords = frac[index] % 1.0)."""
    import folds_oracle as fo
    import site_analysis_b200 as sab
    from site_analysis_b200.folds import BlockFolds
    rng = np.random.default_rng(4)
    F, A, S, N = 40, 5, 3, 9
    pos = fo.random_rows(rng, F, A, S, 0.8, 0.1)
    atom_index = np.array([7, 2, 8, 0, 4])
    frac = rng.random((F, N, 3)) * 3 - 1
    site_index = np.array([11, 12, 13])
    blk = BlockFolds(np.where(pos >= 0, site_index[np.clip(pos, 0, None)], -1), atom_index, site_index)
    blk._runs = fo.runs_from_rows(pos)
    blk._site_off = np.searchsorted(blk._runs[0], np.arange(S + 1))
    blk.frac = frac
    for s in range(S):
        site = sab.VoronoiSite(np.zeros(3)); site.index = int(site_index[s])
        site.points.append(np.array([9.0, 9.0, 9.0]))            # something recorded before the batch stays in front
        site._lazy_points.append((blk, s))
        want = [np.array([9.0, 9.0, 9.0])]
        for f in range(F):
            for a in range(A):
                if pos[f, a] == s:
                    want.append(frac[f, atom_index[a]] % 1.0)
        got = site.points
        assert len(got) == len(want) and all(np.array_equal(g, w) for g, w in zip(got, want))
        site.reset()
        assert len(site.points) == len(want)                      # reset() leaves points alone (reference site.py:86-88)
        site.points = []
        assert site.points == []
