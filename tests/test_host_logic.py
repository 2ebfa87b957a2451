"""CPU-only tests of the host-side logic of site_analysis_b200 (no kernel calls):
C-ABI library loads and exports every declared symbol; the one-off tables; the vectorised
history fold; the reference-shaped object model."""
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names


def test_library_loads_and_exports_every_declared_symbol():
    import ctypes
    from site_analysis_b200 import _native as nat
    header = (ROOT / "include" / "sab200.h").read_text()
    declared = sorted(set(re.findall(r"\b(sab_[a-z0-9_]+)\s*\(", header)) - {"sab_resolve() settles"})
    declared = [d for d in declared if d != "sab_engine"]
    assert sorted(nat.SYMBOLS) == declared
    lib = ctypes.CDLL(str(nat.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert nat.lib().sab_abi_version() == nat.ABI_VERSION


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from site_analysis_b200 import AssignmentEngine, _native as nat
    with pytest.raises(nat.NativeError):
        AssignmentEngine("voronoi", 4, [0], centres=np.zeros((1, 3)))


def test_product_never_imports_the_oracle():
    for p in (ROOT / "site_analysis_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "site_oracle" not in src and "from oracle" not in src, p


@pytest.mark.parametrize("name", ["argyrodite_0p_spherical", "fuzz_spherical_0", "fuzz_polyhedral_0", "readme_spherical"])
def test_distance_rank_table_matches_reference(oracle_mod, name):
    from site_analysis_b200 import _tables as tab
    case = oracle_mod.load_case(name)
    centres = case["centres"] if "centres" in case else case["ref_center"]
    assert np.array_equal(tab.distance_rank_table(centres), case["rank_table"].astype(np.int32))


def test_sqrt_threshold_is_exactly_equivalent():
    from site_analysis_b200 import _tables as tab
    rng = np.random.default_rng(3)
    rcut = np.concatenate([rng.uniform(0.1, 9.0, 2000), [1.5, 2.0, 3.0, 1e-3, 50.0]])
    T = tab.sqrt_threshold(rcut)
    assert np.all(np.sqrt(T) <= rcut)
    assert np.all(np.sqrt(np.nextafter(T, np.inf)) > rcut)


@pytest.mark.parametrize("tag", ["0p", "50p"])
def test_full_pbc_unwrap_reproduces_the_builder_primed_shifts(oracle_mod, tag):
    """SiteFactory primes every argyrodite tetrahedron from the build structure (= frame 0):
    the product's restatement must give the same integer image shifts."""
    from site_analysis_b200 import _tables as tab
    case = oracle_mod.load_case(f"argyrodite_{tag}_polyhedral")
    raw0 = case["frac"][0]
    assert case["primed"].all()
    for s in range(len(case["site_index"])):
        v = case["vert_idx"][s, :case["n_vert"][s]]
        assert np.array_equal(case["cached_raw0"][s], raw0[v])
        got = tab.full_pbc_unwrap(raw0[v], case["ref_center"][s], case["lattice"])
        assert np.array_equal(got, case["shift0"][s])


def test_full_pbc_unwrap_legacy_rule():
    from site_analysis_b200 import _tables as tab
    raw = np.array([[0.98, 0.1, 0.5], [0.02, 0.2, 0.4], [0.95, 0.15, 0.45], [0.05, 0.12, 0.55]])
    sh = tab.full_pbc_unwrap(raw, None, np.eye(3))
    assert sh.tolist() == [[0, 0, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]]


@pytest.mark.parametrize("name", golden_names("fuzz") + ["argyrodite_0p_spherical", "argyrodite_50p_polyhedral"])
def test_history_fold_matches_reference_bookkeeping(oracle_mod, name):
    """fold_history over the golden rows reproduces the reference's transition Counters
    (incl. insertion order) and recent-site pairs, in one batch and in ragged pieces."""
    from site_analysis_b200.engine import fold_history
    case = oracle_mod.load_case(name)
    idx = case["site_index"]
    pos = {int(v): i for i, v in enumerate(idx)}
    rows = np.vectorize(lambda v: -1 if v < 0 else pos[int(v)])(case["atoms_traj"]).astype(np.int32)
    F, A = rows.shape
    want_tr = oracle_mod.transitions_from_case(case)
    for cuts in ([0, F], [0, 1, 2, F // 3, F // 3, F - 1, F]):
        recent = -np.ones((A, 2), dtype=np.int32)
        prev = np.full(A, -2, dtype=np.int32)
        tr = [dict() for _ in idx]
        for a, b in zip(cuts[:-1], cuts[1:]):
            fold_history(recent, prev, tr, rows[a:b], True)
        got_tr = {int(idx[s]): {int(idx[k]): v for k, v in t.items()} for s, t in enumerate(tr) if t}
        assert got_tr == want_tr
        for k in got_tr:
            assert list(got_tr[k]) == list(want_tr[k])
        got_recent = np.where(recent >= 0, idx[np.clip(recent, 0, None)], -1)
        assert np.array_equal(got_recent, case["recent"])
        assert np.array_equal(prev, rows[-1])


def test_trajectory_constructor_errors():
    """trajectory.py:74-98."""
    import site_analysis_b200 as sab
    with pytest.raises(ValueError):
        sab.Trajectory(sites=[], atoms=[sab.Atom(0)])
    with pytest.raises(ValueError):
        sab.Trajectory(sites=[sab.VoronoiSite(np.zeros(3))], atoms=[])
    with pytest.raises(TypeError):
        sab.Trajectory(sites=[sab.VoronoiSite(np.zeros(3)), sab.SphericalSite(np.zeros(3), 1.0)], atoms=[sab.Atom(0)])
    with pytest.raises(TypeError):
        sab.SphericalSiteCollection([sab.VoronoiSite(np.zeros(3))])
    a, b = sab.VoronoiSite(np.zeros(3)), sab.VoronoiSite(np.ones(3) * 0.5)
    b.index = a.index
    with pytest.raises(ValueError):
        sab.VoronoiSiteCollection([a, b])
    c = sab.VoronoiSiteCollection([a])
    with pytest.raises(ValueError):
        c.site_by_index(12345)


def test_builder_validation_and_index_reset():
    """builders.py:734-800: errors, Site index counter restarts at 0, builder resets itself."""
    import site_analysis_b200 as sab
    st = sab.SimpleStructure(np.eye(3) * 8, ["O", "Li"], [[0, 0, 0], [0.5, 0.5, 0.5]])
    with pytest.raises(ValueError):
        sab.TrajectoryBuilder().build()
    with pytest.raises(ValueError):
        sab.TrajectoryBuilder().with_structure(st).build()
    with pytest.raises(ValueError):
        sab.TrajectoryBuilder().with_structure(st).with_voronoi_sites([[0, 0, 0]]).build()
    with pytest.raises(ValueError):
        sab.TrajectoryBuilder().with_structure(st).with_mobile_species("Na").with_voronoi_sites([[0, 0, 0]]).build()
    with pytest.raises(ValueError):
        sab.TrajectoryBuilder().with_spherical_sites([[0, 0, 0]], [1.0, 2.0])
    with pytest.raises(TypeError):
        (sab.TrajectoryBuilder().with_structure(st).with_mobile_species("Li")
         .with_voronoi_sites([[0, 0, 0]]).with_spherical_sites([[0, 0, 0]], 1.0).build())
    sab.Site._newid = 17
    b = sab.TrajectoryBuilder().with_structure(st).with_mobile_species("Li").with_spherical_sites([[0, 0, 0], [.5, .5, .5]], 1.5, "x")
    t = b.build()
    assert [s.index for s in t.sites] == [0, 1] and [s.label for s in t.sites] == ["x", "x"]
    assert [a.index for a in t.atoms] == [1]
    assert b._structure is None and b._site_generators == []


def test_lazy_trajectories_materialise_like_lists():
    import site_analysis_b200 as sab
    atom = sab.Atom(3)
    atom.trajectory.append(5)
    atom._lazy_blocks.append(np.array([5, -1, 7], dtype=np.int32))
    assert atom._last_trajectory_entry() == 7
    assert atom.trajectory == [5, 5, None, 7]
    site = sab.VoronoiSite(np.zeros(3)); site.index = 7
    rows = np.array([[7, 1], [2, 7], [7, 7]], dtype=np.int32)
    site._lazy_blocks.append((rows, np.array([10, 11])))
    assert site.trajectory == [[10], [11], [10, 11]]
    assert site.average_occupation == 1.0


def test_peer_result_buffers_slot_arithmetic():
    """distributed.PeerResultBuffers.next(): this rank's rows start at rank * F_local * A int32 elements inside every
    peer's full array; buffers alternate (host logic only; the symmetric-memory allocation itself needs GPUs)."""
    import torch
    from site_analysis_b200.distributed import PeerResultBuffers
    b = PeerResultBuffers.__new__(PeerResultBuffers)
    b.world, b.rank, b.F, b.A, b.turn = 4, 2, 10, 3, 0
    b.buffers = [torch.zeros((40, 3), dtype=torch.int32), torch.ones((40, 3), dtype=torch.int32)]
    b.ptrs = [[1000, 2000, 3000, 4000], [5000, 6000, 7000, 8000]]
    full, slot, peers = b.next()
    assert full is b.buffers[0] and slot.shape == (10, 3) and slot.data_ptr() == full[20:].data_ptr()
    assert peers == [1000 + 240, 2000 + 240, 4000 + 240]          # 2 * 10 * 3 * 4 bytes into ranks 0, 1 and 3
    full, slot, peers = b.next()
    assert full is b.buffers[1] and peers == [5240, 6240, 8240]
    assert b.next()[0] is b.buffers[0]
    # NVLS: the multicast view of the same arrays, same slot offset
    b.multicast, b.mc_ptrs, b.turn = True, [90000, 91000], 0
    b.next()
    assert b.mc_slot == 90000 + 240
    b.next()
    assert b.mc_slot == 91000 + 240


def test_transition_store_keeps_counter_semantics_in_both_forms():
    """AssignmentEngine's transition store (site.py:71 Counters of all sites): per-site dicts in insertion order <-> the
    (source, destination, count) arrays the device fold merges into.  Merging pairs in first-occurrence order must give what
    ``transitions[src][dst] += n`` gives (counts added, NEW keys appended in order), and a snapshot must survive later
    merges in whichever form it was taken (host logic only: no device is touched)."""
    from site_analysis_b200.engine import AssignmentEngine
    e = AssignmentEngine.__new__(AssignmentEngine)
    e.S = 5
    e.transitions = [dict() for _ in range(5)]
    e.transitions[2][4] = 3                                   # edited through the dict view
    e.transitions[2][0] = 1
    snap_d = e._tr_snapshot()
    assert snap_d[0] == "dicts"
    e._tr_add_pairs(np.array([2, 1, 2, 1]), np.array([0, 3, 1, 0]), np.array([5, 2, 7, 1]))
    snap_a = e._tr_snapshot()
    assert snap_a[0] == "arrays"
    want = [{}, {3: 2, 0: 1}, {4: 3, 0: 6, 1: 7}, {}, {}]
    assert e.transitions == want and [list(t) for t in e.transitions] == [list(t) for t in want]      # insertion order too
    e._tr_add_pairs(np.array([2, 4]), np.array([4, 4]), np.array([1, 9]))
    assert e.transitions[2] == {4: 4, 0: 6, 1: 7} and e.transitions[4] == {4: 9}
    e._tr_restore(snap_a)
    assert e.transitions == want
    e._tr_restore(snap_d)
    assert e.transitions == [{}, {}, {4: 3, 0: 1}, {}, {}]
    e.h = None                                                # nothing to destroy


def test_exhaustive_kernels_are_not_chosen_beyond_their_shared_memory():
    """The exhaustive spherical / Voronoi kernels stage every site in shared memory (40 / 24 B per site); single-frame calls on
    larger site tables must build the cell lists first."""
    from site_analysis_b200 import _native as nat
    from site_analysis_b200.engine import AssignmentEngine
    e = AssignmentEngine.__new__(AssignmentEngine)
    for kind, S, want in ((nat.SPHERICAL, 5000, False), (nat.SPHERICAL, 5001, True), (nat.VORONOI, 8448, True),
                          (nat.VORONOI, 8333, False), (nat.DYNAMIC_VORONOI, 576, False)):
        e.k, e.S = kind, S
        assert e._exhaustive_too_big() == want, (kind, S)
    e.h = None
