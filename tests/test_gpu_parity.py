"""GPU parity: the CUDA path (through the reference-shaped Python API and the C ABI) against
(a) the committed golden outputs of the unmodified reference and (b) the CPU oracle run on
the same inputs.  Bit-exact: every (frame, atom) site index incl. None, transition Counters
incl. insertion order, recent-site history, dynamic-Voronoi centres as float64 bit patterns."""
import numpy as np
import pytest

from conftest import golden_names
from helpers import has_jump, trajectory_from_case, transitions_of

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


def _run_batch(case):
    t = trajectory_from_case(case)
    rows = t.trajectory_from_arrays(case["frac"], case["lattice"])
    return t, rows


@pytest.mark.parametrize("name", golden_names())
def test_batch_matches_reference_golden(oracle_mod, name):
    import site_analysis_b200 as sab
    case = oracle_mod.load_case(name)
    t, rows = _run_batch(case)
    if str(case["kind"]) == "PolyhedralSite" and has_jump(case):
        # a framework step >= 0.3 invalidates site caches (pbc_utils.py:216): the batch was redone by the exact replay of the
        # reference's lazily updated per-site caches (csrc/k_polyhedral_lazy.cuh) -- and must match like any other
        assert t.site_collection._engine._lazy_mode and t.site_collection._engine.lazy_stats()["invalidations"] > 0
    want = case["atoms_traj"]
    assert rows.shape == want.shape
    assert int((rows != want).sum()) == 0
    # the oracle on the same inputs says the same (it is pinned to the reference separately)
    o = oracle_mod.SiteOracle.from_case(case)
    assert np.array_equal(o.run(case["frac"], case["lattice"]), rows)
    want_tr = oracle_mod.transitions_from_case(case)
    got_tr = transitions_of(t)
    assert got_tr == want_tr
    for k in got_tr:
        assert list(got_tr[k]) == list(want_tr[k])
    got_recent = np.array([[-1 if r is None else r for r in a._recent_sites] for a in t.atoms])
    assert np.array_equal(got_recent, case["recent"])
    assert len(t) == rows.shape[0] and t.timesteps[:3] == [1, 2, 3][:len(t)]
    assert [(-1 if v is None else v) for v in t.atoms[0].trajectory] == want[:, 0].tolist()
    assert t.atom_sites == [None if v < 0 else int(v) for v in want[-1]]


@pytest.mark.parametrize("name", ["readme_spherical", "simple_cubic_polyhedral", "fuzz_spherical_1", "fuzz_voronoi_0",
                                  "fuzz_polyhedral_1", "fuzz_polyhedral_6", "fuzz_dynvor_1", "fuzz_dynvor_2"])
def test_per_frame_api_matches_reference_golden(oracle_mod, name):
    """append_timestep one structure at a time == the reference's own loop (trajectory.py:413-432)."""
    import site_analysis_b200 as sab
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case)
    lat = case["lattice"]
    species = [str(s) for s in case["species"]]
    for f in range(case["frac"].shape[0]):
        st = sab.SimpleStructure(lat if lat.ndim == 2 else lat[f], species, case["frac"][f])
        t.append_timestep(st, t=f + 1)
    got = np.array([[(-1 if v is None else v) for v in a.trajectory] for a in t.atoms]).T
    assert np.array_equal(got, case["atoms_traj"])
    assert transitions_of(t) == oracle_mod.transitions_from_case(case)
    # sites_trajectory: atoms listed in atom order per site per frame
    st_last = t.sites_trajectory[-1]
    for s, site in enumerate(t.sites):
        assert st_last[s] == [a.index for a in t.atoms if a.in_site == site.index]


@pytest.mark.parametrize("name", ["fuzz_voronoi_1", "fuzz_polyhedral_0", "fuzz_dynvor_0", "fuzz_spherical_2",
                                  "fuzz_polyhedral_big1", "fuzz_spherical_big3", "fuzz_dynvor_big2"])
def test_ragged_batches_equal_one_batch(oracle_mod, name):
    """State carried between calls (PBC wrap counters, history) makes any split equivalent."""
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case)
    F = case["frac"].shape[0]
    lat = case["lattice"]
    parts = []
    for a, b in [(0, 1), (1, 2), (2, 31), (31, 31), (31, F)]:
        parts.append(t.trajectory_from_arrays(case["frac"][a:b], lat if lat.ndim == 2 else lat[a:b]))
    assert np.array_equal(np.vstack(parts), case["atoms_traj"])
    assert transitions_of(t) == oracle_mod.transitions_from_case(case)


@pytest.mark.parametrize("name", golden_names("fuzz_dynvor") + ["argyrodite_0p_dynvor"])
def test_dynamic_voronoi_centres_bitwise(oracle_mod, name):
    import torch
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case)
    lat = case["lattice"]
    frames = list(case["centre_frames"])
    for f in range(case["frac"].shape[0]):
        t.trajectory_from_arrays(case["frac"][f:f + 1], lat if lat.ndim == 2 else lat[f:f + 1])
        if f in frames:
            got = np.array([s.centre for s in t.sites])
            want = case["centres_at"][frames.index(f)]
            assert np.array_equal(got.view(np.int64), want.view(np.int64)), f
    assert torch.cuda.is_available()


@pytest.mark.parametrize("name", ["argyrodite_0p_polyhedral", "fuzz_polyhedral_1", "fuzz_polyhedral_5", "fuzz_polyhedral_7"])
def test_tetrahedra_need_no_seeded_topology(oracle_mod, name):
    """Simplicial polyhedra: every vertex triple is a face, so the topology hulled from the first
    frame gives the reference's answers without knowing when the reference queried each site."""
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case, seed_topology=False)
    assert np.array_equal(t.trajectory_from_arrays(case["frac"], case["lattice"]), case["atoms_traj"])


@pytest.mark.parametrize("name", ["argyrodite_0p_polyhedral", "argyrodite_100p_polyhedral", "fuzz_polyhedral_1",
                                  "fuzz_polyhedral_3", "fuzz_polyhedral_5", "fuzz_polyhedral_7"])
def test_pruned_kernel_equals_exhaustive_kernel(oracle_mod, name):
    """Cell-list kernel (k_polyhedral_cells.cuh) vs exhaustive kernel (k_polyhedral.cuh), both vs golden."""
    from site_analysis_b200 import _native as nat
    case = oracle_mod.load_case(name)
    outs = []
    for use_cells, opts in ((True, {}), (True, {nat.OPT_SEQ_CHUNK: 7}), (True, {nat.OPT_POLY_KERNEL: 1}),
                            (True, {nat.OPT_POLY_KERNEL: 3}), (True, {nat.OPT_POLY_KERNEL: 4}), (True, {nat.OPT_FP32_FILTER: 0}),
                            (True, {nat.OPT_POLY_KERNEL: 4, nat.OPT_FP32_FILTER: 0}), (True, {nat.OPT_SEQ_CHUNK: 3}),
                            (True, {nat.OPT_POLY_KERNEL: 5}), (True, {nat.OPT_POLY_KERNEL: 5, nat.OPT_SEQ_CHUNK: 7}),
                            (True, {nat.OPT_POLY_KERNEL: 5, nat.OPT_FP32_FILTER: 0}),
                            (True, {nat.OPT_STREAM_PRODUCERS: 1, nat.OPT_SEQ_CHUNK: 5}), (True, {nat.OPT_STREAM_PRODUCERS: 4, nat.OPT_STREAM_REC_SMEM: 0}),
                            # generation 4 of the stream kernel (the default is generation 5) and other ring shapes of generation 5
                            (True, {nat.OPT_STREAM_GEN: 4}), (True, {nat.OPT_STREAM_GEN: 4, nat.OPT_SEQ_CHUNK: 7}),
                            (True, {nat.OPT_STREAM_GEN: 4, nat.OPT_FP32_FILTER: 0}),
                            (True, {nat.OPT_STREAM5_RAW: 2, nat.OPT_STREAM5_ROWS: 2, nat.OPT_SEQ_CHUNK: 5}),
                            (True, {nat.OPT_STREAM5_BLOCKS: 1, nat.OPT_STREAM5_RAW: 4, nat.OPT_STREAM5_ROWS: 8, nat.OPT_STREAM_REC_SMEM: 0}),
                            (False, {})):
        t = trajectory_from_case(case)
        t.site_collection.use_cells = use_cells
        t.site_collection.engine_options = opts          # stream (default) / odd chunk / frame-parallel / tile / ...
        outs.append(t.trajectory_from_arrays(case["frac"], case["lattice"]))
        eng = t.site_collection._engine
        assert (eng._cells is not None) == use_cells, (name, use_cells, eng._cells)
    for o in outs:
        assert np.array_equal(o, case["atoms_traj"])


def test_frames_outside_the_validity_tube_fall_back_exactly(oracle_mod):
    """Anchors from the first 70 frames; the next 70 frames (a different vibration sample, some
    outside the tube) must still come out identical to the oracle."""
    case = oracle_mod.load_case("argyrodite_0p_polyhedral")
    t = trajectory_from_case(case)
    a = t.trajectory_from_arrays(case["frac"][:70], case["lattice"])
    eng = t.site_collection._engine
    assert eng._cells is not None
    b = t.trajectory_from_arrays(case["frac"][70:], case["lattice"])
    assert np.array_equal(np.vstack([a, b]), case["atoms_traj"])
    assert eng.last_info["full27"] >= 0


def test_voronoi_cell_lists_equal_exhaustive_and_oracle(oracle_mod):
    """k_nearest_assign_cells vs k_nearest_assign (all centres) vs the oracle, with a fluctuating and
    then a strongly strained lattice (the latter must take the exhaustive path, not a wrong answer)."""
    import torch
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 400, device="cuda")
    lat[200:] *= 1.09                                   # 9 % strain relative to the lattice the lists were built with
    lat[300:] *= 2.5                                    # absurd strain: the exclusion bound is void -> exhaustive path
    outs = []
    for use_cells in (True, False):
        eng = AssignmentEngine("voronoi", g.n_atoms, g.mobile_idx, centres=g.site_centres)
        eng.use_cells = use_cells
        outs.append(eng.assign_device(frac, lat).cpu().numpy())
        assert (eng._cells is not None) == use_cells
        if use_cells:
            assert eng.last_info["full27"] > 0          # strained frames fell back
    assert np.array_equal(outs[0], outs[1])
    o = oracle_mod.SiteOracle("voronoi", g.mobile_idx, centres=g.site_centres)
    sel = np.r_[0:40, 200:240, 300:320]
    want = o.run(frac.cpu().numpy()[sel], lat.cpu().numpy()[sel], positions=True)
    assert np.array_equal(outs[0][sel], want)
    assert torch.cuda.is_available()


def test_dynamic_voronoi_cell_lists_equal_exhaustive_and_oracle(oracle_mod):
    """k_dynvor_assign_cells vs the exhaustive kernel vs the oracle on the argyrodite geometry."""
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 300, device="cuda", sigma_framework=0.12)
    slots = [list(map(int, t)) for t in g.tetrahedra]
    outs = []
    for use_cells in (True, False):
        eng = AssignmentEngine("dynamic_voronoi", g.n_atoms, g.mobile_idx, slots=slots, reference_centres=list(g.site_centres))
        eng.use_cells = use_cells
        outs.append(eng.assign_device(frac, lat).cpu().numpy())
        assert (eng._cells is not None) == use_cells
    assert np.array_equal(outs[0], outs[1])
    S = len(slots)
    o = oracle_mod.SiteOracle("dynamic_voronoi", g.mobile_idx, ref_idx=np.array(slots, dtype=np.int32), n_ref=np.full(S, 4, np.int32),
                              has_ref_center=np.ones(S, np.uint8), ref_center=g.site_centres)
    want = o.run(frac.cpu().numpy()[:60], lat.cpu().numpy()[:60], positions=True)
    assert np.array_equal(outs[0][:60], want)


def test_spherical_cell_lists_equal_exhaustive_and_oracle(oracle_mod):
    """k_spherical_assign_cells vs k_spherical_assign (all sites) vs the oracle: overlapping spheres (so the spilled
    candidate sets and the resolver are exercised), per-site radii, a fluctuating lattice, then frames strained beyond
    the lists' head-room, which must take the exhaustive loop."""
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 300, device="cuda", seed=77)
    lat[200:] *= 1.09
    rcut = np.where(np.arange(len(g.site_centres)) % 3 == 0, 1.1, 1.6)
    outs, infos = [], []
    for use_cells in (True, False):
        eng = AssignmentEngine("spherical", g.n_atoms, g.mobile_idx, centres=g.site_centres, rcut=rcut)
        eng.use_cells = use_cells
        outs.append(eng.assign_device(frac, lat).cpu().numpy())
        eng.sync_history()
        infos.append((eng.last_info, eng.transitions))
        assert (eng._cells is not None) == use_cells
    assert infos[0][0]["full27"] == 100                  # exactly the strained frames fell back
    assert infos[0][0]["ambiguous"] == infos[1][0]["ambiguous"] > 0
    assert np.array_equal(outs[0], outs[1]) and infos[0][1] == infos[1][1]
    o = oracle_mod.SiteOracle("spherical", g.mobile_idx, centres=g.site_centres, rcut=rcut)
    want = o.run(frac.cpu().numpy(), lat.cpu().numpy(), positions=True)
    assert np.array_equal(outs[0], want)
    assert (want < 0).any() and (want >= 0).any()


def test_baseline_config3_voronoi_4x4x4_supercell(oracle_mod):
    """BASELINE.json config 3: synthetic 4x4x4 argyrodite, N = 3328, A = 1536, S = 8448 Voronoi centres,
    coordinates wrapped into [0, 1), fluctuating slightly triclinic lattice.  Cell-list kernel == exhaustive kernel on
    every frame; both == the oracle on the frames the CPU can afford (0.7 s per frame)."""
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(4)
    assert (g.n_atoms, g.n_mobile, len(g.site_centres)) == (3328, 1536, 8448)
    frac, lat = syn.argyrodite_trajectory(g, 96, device="cuda", seed=20260103)
    outs = []
    for use_cells in (True, False):
        eng = AssignmentEngine("voronoi", g.n_atoms, g.mobile_idx, centres=g.site_centres)
        eng.use_cells = use_cells
        outs.append(eng.assign_device(frac, lat).cpu().numpy())
        assert (eng._cells is not None) == use_cells
    assert np.array_equal(outs[0], outs[1])
    assert outs[0].min() >= 0                                   # Voronoi never yields None
    o = oracle_mod.SiteOracle("voronoi", g.mobile_idx, centres=g.site_centres)
    sel = np.r_[0:3, 93:96]
    want = o.run(frac.cpu().numpy()[sel], lat.cpu().numpy()[sel], positions=True)
    assert np.array_equal(outs[0][sel], want)


def test_baseline_config4_garnet_dynamic_voronoi_two_centre_groups(oracle_mod):
    """BASELINE.json config 4: LLZO garnet 2x2x2, N = 1536, A = 448, S = 576 dynamic-Voronoi sites from the oxygen
    coordination -- 24d tetrahedra (4 reference atoms) and 48g octahedra (6), i.e. two centre groups
    (dynamic_voronoi_site_collection.py:151-164).  Assignments and centres (bit patterns) against the oracle."""
    import torch
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.garnet_geometry(2)
    S = len(g.site_centres)
    assert (g.n_atoms, g.n_mobile, S) == (1536, 448, 576) and {len(s) for s in g.reference_slots} == {4, 6}
    frac, lat = syn.synthetic_trajectory(g, 400, device="cuda", seed=20260104)
    outs = []
    for use_cells in (True, False):
        eng = AssignmentEngine("dynamic_voronoi", g.n_atoms, g.mobile_idx, slots=g.reference_slots,
                               reference_centres=list(g.site_centres))
        eng.use_cells = use_cells
        outs.append(eng.assign_device(frac, lat).cpu().numpy())
        assert (eng._cells is not None) == use_cells
    assert list(np.unique(eng.group_of_site)) == [0, 1]
    assert np.array_equal(outs[0], outs[1])
    ref = np.zeros((S, 6), dtype=np.int32)
    for i, sl in enumerate(g.reference_slots):
        ref[i, :len(sl)] = sl
    o = oracle_mod.SiteOracle("dynamic_voronoi", g.mobile_idx, ref_idx=ref, n_ref=np.array([len(s) for s in g.reference_slots], np.int32),
                              has_ref_center=np.ones(S, np.uint8), ref_center=g.site_centres)
    n = 120
    want = o.run(frac.cpu().numpy()[:n], lat.cpu().numpy()[:n], positions=True)
    assert np.array_equal(outs[0][:n], want)
    eng2 = AssignmentEngine("dynamic_voronoi", g.n_atoms, g.mobile_idx, slots=g.reference_slots, reference_centres=list(g.site_centres))
    eng2.assign_device(frac[:n].contiguous(), lat[:n].contiguous())
    cen = eng2.dynvor_centres(frac[n - 1:n].contiguous())[0].cpu().numpy()
    assert np.array_equal(cen.view(np.int64), np.asarray(o.centres(), dtype=np.float64).view(np.int64))
    assert torch.cuda.is_available()


def test_reset_and_replay(oracle_mod):
    """Trajectory.reset(): history and PBC caches cleared, topology kept, same answers again."""
    case = oracle_mod.load_case("fuzz_polyhedral_1")
    t = trajectory_from_case(case)
    a = t.trajectory_from_arrays(case["frac"], case["lattice"])
    assert all(s._face_topology_cache is not None for s in t.sites)
    t.reset()
    assert len(t) == 0 and all(not s.transitions for s in t.sites) and t.atoms[0].trajectory == []
    assert all(s._pbc_image_shifts is None for s in t.sites)
    b = t.trajectory_from_arrays(case["frac"], case["lattice"])
    o = oracle_mod.SiteOracle.from_case(case)
    o.run(case["frac"], case["lattice"]); o.reset()
    assert np.array_equal(b, o.run(case["frac"], case["lattice"]))
    assert np.array_equal(a, case["atoms_traj"])


def test_empty_and_single_frame_inputs(oracle_mod):
    case = oracle_mod.load_case("fuzz_voronoi_0")
    t = trajectory_from_case(case)
    out = t.trajectory_from_arrays(case["frac"][:0], case["lattice"])
    assert out.shape == (0, len(t.atoms)) and len(t) == 0
    out = t.trajectory_from_arrays(case["frac"][:1], case["lattice"])
    assert np.array_equal(out, case["atoms_traj"][:1])


def test_c_abi_host_call_matches_device_call(oracle_mod):
    """sab_assign_host (pageable host buffers, chunked, overlapped) == sab_assign_device."""
    from site_analysis_b200 import AssignmentEngine
    case = oracle_mod.load_case("argyrodite_0p_voronoi")
    frac = np.ascontiguousarray(np.tile(case["frac"], (3, 1, 1)))
    eng = AssignmentEngine("voronoi", frac.shape[1], case["atom_idx"], centres=case["centres"])
    a = eng.assign_host_pipelined(frac, case["lattice"], chunk_frames=100)
    eng2 = AssignmentEngine("voronoi", frac.shape[1], case["atom_idx"], centres=case["centres"])
    b = eng2.assign(frac, case["lattice"])
    assert np.array_equal(a, b)
    pos = {int(v): i for i, v in enumerate(case["site_index"])}
    want = np.vectorize(lambda v: pos[int(v)])(case["atoms_traj"])
    assert np.array_equal(a[:140], want) and np.array_equal(a[280:], want)
    assert eng.last_info["launches"] > 0


def test_fp32_pretest_never_changes_a_result():
    """The FP32 pre-test (k_polyhedral_cells.cuh: sab_tet_filter32) may only decide points that are
    decisively inside / outside; everything within its error band must reach the float64 test.  Mobile
    ions are put exactly on vertices, edges and faces of the frame's own tetrahedra and 1e-4 .. 1e-16 off
    them; filter on == filter off == oracle, frame by frame."""
    import pathlib
    import sys
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
    import bench
    from site_analysis_b200 import AssignmentEngine, _native as nat, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 260, device="cpu")
    frac, lat = frac.numpy().copy(), lat.numpy().copy()
    rng = np.random.default_rng(77)
    tets = np.asarray(g.tetrahedra)
    mob = g.mobile_idx
    n_adv = 0
    for f in range(20, 260):
        for a in range(len(mob)):
            if rng.random() < 0.5:
                continue
            v = frac[f, tets[rng.integers(len(tets))]]
            rel = v - v[0]
            rel -= np.round(rel)
            kind = rng.integers(4)
            w = rng.random(4)
            if kind == 0:   w[rng.integers(4)] = 0.0                             # on a face
            elif kind == 1: w[rng.choice(4, 2, replace=False)] = 0.0             # on an edge
            elif kind == 2: w = np.eye(4)[rng.integers(4)]                       # on a vertex
            w = w / w.sum()
            p = v[0] + w @ rel
            if rng.random() < 0.7:
                p = p + rng.normal(size=3) * 10.0 ** (-rng.integers(4, 17))
            frac[f, mob[a]] = p % 1.0
            n_adv += 1
    assert n_adv > 10000
    kw = lambda: syn.polyhedral_engine_kwargs(g, frac[0], lat[0])                # noqa: E731
    want = bench.oracle_for(g, frac[0], lat[0])().run(frac, lat, positions=True)
    outs = []
    for kernel, flt in ((0, 1), (0, 0), (5, 1), (5, 0), (4, 1), (4, 0), (-4, 1), (-4, 0)):   # stream gen 5 / tile / chunk-sequential / stream gen 4
        eng = AssignmentEngine("polyhedral", g.n_atoms, mob, device=0, **kw())
        if kernel < 0:
            eng.set_option(nat.OPT_STREAM_GEN, -kernel)
            kernel = 0
        eng.set_option(nat.OPT_POLY_KERNEL, kernel)
        eng.set_option(nat.OPT_FP32_FILTER, flt)
        outs.append(eng.assign(frac, lat))
        assert eng._cells is not None
    for o in outs:
        assert np.array_equal(o, want)


def test_stream_kernel_chain_check_falls_back_to_the_scan_kernels():
    """The stream kernel takes the wrap count as rint(raw - anchor) and checks it against the reference's
    sequential update frame by frame.  A framework atom that drifts more than half a cell away from its
    anchor (outside guard G1, hence strict=False) breaks the chain: the batch must be redone with the scan
    kernels and agree with them exactly; the PBC carry must continue correctly afterwards."""
    import warnings
    from site_analysis_b200 import AssignmentEngine, _native as nat, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 700, device="cpu")
    frac, lat = frac.numpy().copy(), lat.numpy().copy()
    victim = int(np.asarray(g.tetrahedra)[0][0])
    drift = np.zeros(700)
    drift[250:450] = np.linspace(0.0, 0.56, 200)          # slow (0.0028 per frame), then stays displaced
    drift[450:] = 0.56
    frac[:, victim, 0] = (frac[:, victim, 0] + drift) % 1.0
    kw = lambda: syn.polyhedral_engine_kwargs(g, frac[0], lat[0])                # noqa: E731
    outs, redo = [], []
    for kernel in (0, 5, 4, -4):
        eng = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, device=0, strict=False, **kw())
        if kernel < 0:
            eng.set_option(nat.OPT_STREAM_GEN, -kernel)
        eng.set_option(nat.OPT_POLY_KERNEL, max(kernel, 0))
        parts = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for a, b in ((0, 200), (200, 500), (500, 700)):
                parts.append(eng.assign(frac[a:b], lat[a:b]))
                redo.append((kernel, eng.last_info.get("scan_redo", 0)))
        outs.append(np.concatenate(parts))
    assert (0, 1) in redo and (-4, 1) in redo, redo          # both stream kernels did hit their fallback
    assert all(r == 0 for k, r in redo if k > 0), redo
    assert all(np.array_equal(outs[0], o) for o in outs[1:])


def test_fast_path_is_the_one_that_runs():
    """Well-behaved data must not leave the cell lists' validity tube: no frame may take the exhaustive
    generic path (a silent fallback would still be correct, just ~1000x slower)."""
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 2000, device="cuda")
    eng = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, device=0,
                           **syn.polyhedral_engine_kwargs(g, frac[0].cpu().numpy(), lat[0].cpu().numpy()))
    eng.assign_device(frac, lat)
    assert eng._cells is not None
    assert eng.last_info["full27"] == 0, eng.last_info
    assert eng.last_info["scan_redo"] == 0, eng.last_info


@pytest.mark.parametrize("name", ["argyrodite_0p_polyhedral", "simple_cubic_polyhedral", "fuzz_polyhedral_0", "fuzz_polyhedral_2",
                                  "fuzz_polyhedral_3", "fuzz_polyhedral_5", "fuzz_polyhedral_8", "fuzz_polyhedral_9",
                                  "fuzz_polyhedral_big4", "fuzz_polyhedral_big7"])
def test_exact_replay_kernel_equals_reference_golden(oracle_mod, name):
    """exact_lazy = True: every batch through the literal replay of the reference's per-site lazy caches
    (polyhedral_site.py:322-416) -- same rows, Counters (with insertion order) and recent sites as the reference, for
    tetrahedra, cubes, primed and unprimed sites, with and without reference centres, with and without jumps; in two
    ragged batches (state carried in the per-site caches)."""
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case)
    t.site_collection.exact_lazy = True
    F = case["frac"].shape[0]
    cut = F // 3
    lat = case["lattice"]
    rows = np.vstack([t.trajectory_from_arrays(case["frac"][:cut], lat if lat.ndim == 2 else lat[:cut]),
                      t.trajectory_from_arrays(case["frac"][cut:], lat if lat.ndim == 2 else lat[cut:])])
    assert t.site_collection._engine._lazy_mode
    assert np.array_equal(rows, case["atoms_traj"])
    got_tr, want_tr = transitions_of(t), oracle_mod.transitions_from_case(case)
    assert got_tr == want_tr
    for k in got_tr:
        assert list(got_tr[k]) == list(want_tr[k])
    got_recent = np.array([[-1 if r is None else r for r in a._recent_sites] for a in t.atoms])
    assert np.array_equal(got_recent, case["recent"])


def test_slow_drift_beyond_the_excursion_guard_is_replayed_exactly():
    """A framework atom that drifts 0.45 of a cell over 600 frames (0.00075 per frame: no single step invalidates anything)
    breaks guard G1 -- the reference's lazily updated caches (a site sees the atom only when it is queried) are no longer
    a function of the frame.  The engine must notice at the end of the batch, redo it with the exact replay kernel and
    match the sequential oracle (which restates exactly that laziness), instead of raising."""
    import bench
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(2)
    frac, lat = syn.argyrodite_trajectory(g, 600, device="cpu")
    frac, lat = frac.numpy().copy(), lat.numpy().copy()
    for k, victim in enumerate(np.asarray(g.tetrahedra)[[0, 40, 300], 0]):
        frac[:, int(victim), k % 3] = (frac[:, int(victim), k % 3] + np.linspace(0.0, 0.45, 600)) % 1.0
    want = bench.oracle_for(g, frac[0], lat[0])().run(frac, lat, positions=True)
    eng = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, device=0, **syn.polyhedral_engine_kwargs(g, frac[0], lat[0]))
    got = eng.assign(frac, lat)
    assert eng._lazy_mode, "the excursion guard should have sent this batch to the exact replay"
    assert np.array_equal(got, want)
    # the same through the device-resident entry, and with the replay selected from the start
    import torch
    for lazy in (None, True):
        eng = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, device=0, **syn.polyhedral_engine_kwargs(g, frac[0], lat[0]))
        eng.exact_lazy = lazy
        got = eng.assign_device(torch.from_numpy(frac).cuda(), torch.from_numpy(lat).cuda()).cpu().numpy()
        assert eng._lazy_mode and np.array_equal(got, want)


@pytest.mark.parametrize("name", ["argyrodite_0p_spherical", "fuzz_spherical_0", "fuzz_spherical_1", "fuzz_spherical_2", "fuzz_spherical_3",
                                  "simple_cubic_spherical", "fuzz_polyhedral_0", "fuzz_polyhedral_4", "fuzz_polyhedral_9"])
def test_lazy_distance_ranking_equals_the_dense_table(oracle_mod, name):
    """No S x (S - 1) rank table (site_collection.py:90-111) on the device: candidates are ranked by their distance to the
    anchor in the kernel, and only anchors whose candidates TIE get their true row (np.argsort of the reference) from the
    host.  Same rows, Counters (with insertion order) and recent sites as the reference; far fewer rows than sites."""
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case, use_case_rank=False)
    t.site_collection.rank_dense_max = 0
    if "polyhedral" in name:
        t.site_collection.exact_lazy = True        # the replay kernel walks whole ranked rows: by distance, row by row on demand
    rows = t.trajectory_from_arrays(case["frac"], case["lattice"])
    eng = t.site_collection._engine
    assert np.array_equal(rows, case["atoms_traj"])
    got_tr, want_tr = transitions_of(t), oracle_mod.transitions_from_case(case)
    assert got_tr == want_tr
    for k in got_tr:
        assert list(got_tr[k]) == list(want_tr[k])
    if getattr(eng, "_priority_uploaded", False):
        assert eng._lazy_rank and eng.rank_table is None
        assert getattr(eng, "rank_rows_cached", 0) <= len(t.sites)
        print(name, "rows supplied for", getattr(eng, "rank_rows_cached", 0), "of", len(t.sites), "anchors")


def test_large_overlapping_sphere_system_needs_no_dense_rank_table(oracle_mod):
    """8,448 overlapping spheres (4x4x4 argyrodite, r = 1.5 A, A = 1536).  The dense rank table
    and its inverse would take 2 x 285 MB; the lazy ranking supplies rows only for anchors whose candidates tie.  Results
    must equal the sequential oracle (which uses the reference's full table)."""
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    g = syn.argyrodite_geometry(4)
    frac, lat = syn.argyrodite_trajectory(g, 64, device="cuda", seed=5)
    eng = AssignmentEngine("spherical", g.n_atoms, g.mobile_idx, centres=g.site_centres, rcut=1.5)
    got = eng.assign_device(frac, lat).cpu().numpy()
    eng.sync_history()
    assert eng.last_info["ambiguous"] > 50000 and eng._lazy_rank and eng.rank_table is None
    rows_cached = getattr(eng, "rank_rows_cached", 0)
    table_bytes = rows_cached * 2 * 4 * len(g.site_centres) + 24 * len(g.site_centres)
    assert table_bytes < 100e6, (rows_cached, table_bytes)
    o = oracle_mod.SiteOracle("spherical", g.mobile_idx, centres=g.site_centres, rcut=1.5)
    want = o.run(frac.cpu().numpy(), lat.cpu().numpy(), positions=True)
    assert np.array_equal(got, want)
    assert {s: t for s, t in enumerate(eng.transitions) if t} == o.transitions()       # site_index of this oracle = list positions
    print(f"S = 8448: {eng.last_info['ambiguous']} ambiguous atom-frames, {rows_cached} rank rows supplied ({table_bytes / 1e6:.1f} MB of tables)")


def test_speculative_resolver_equals_the_per_frame_resolvers_and_the_oracle(oracle_mod):
    """csrc/k_resolve_spec.cuh settles chunks of frames without a barrier per frame (reads of the shared transition Counters
    are confirmed afterwards, a conflict truncates the chunk).  It must give exactly what the per-frame resolvers
    (k_resolve_fast, k_resolve) and the sequential oracle give: rows, Counters WITH insertion order, recent sites -- on
    overlapping spheres where 98 % of the atom-frames sit in several sites (site_collection.py:113-182), in one batch and
    in ragged batches (history carried across calls)."""
    from site_analysis_b200 import AssignmentEngine, _native as nat, synthetic as syn
    g = syn.argyrodite_geometry(2)
    F = 3000
    frac, lat = syn.argyrodite_trajectory(g, F, device="cuda", seed=11)
    out = {}
    for label, mode, cuts in (("spec", 0, [F]), ("spec ragged", 0, [1, 7, 40, 1000, F]), ("fast", 1, [F]), ("global", 2, [F])):
        eng = AssignmentEngine("spherical", g.n_atoms, g.mobile_idx, centres=g.site_centres, rcut=1.5)
        eng.set_option(nat.OPT_RESOLVER, mode)
        rows, a = [], 0
        for b in cuts:
            rows.append(eng.assign_device(frac[a:b].contiguous(), lat[a:b].contiguous()).cpu().numpy())
            a = b
        eng.sync_history()
        out[label] = (np.concatenate(rows), [list(t.items()) for t in eng.transitions], eng.recent.copy(), eng.prev.copy())
    ref = out["fast"]
    assert (ref[0] >= 0).mean() > 0.9
    for label, got in out.items():
        assert np.array_equal(got[0], ref[0]), label
        assert got[1] == ref[1], label                  # keys in insertion order, counts
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3]), label
    n = 400
    o = oracle_mod.SiteOracle("spherical", g.mobile_idx, centres=g.site_centres, rcut=1.5)
    want = o.run(frac[:n].cpu().numpy(), lat[:n].cpu().numpy(), positions=True)
    assert np.array_equal(out["spec"][0][:n], want)


@pytest.mark.parametrize("name", ["simple_cubic_polyhedral", "fuzz_polyhedral_0", "fuzz_polyhedral_2", "fuzz_polyhedral_4",
                                  "fuzz_polyhedral_6", "fuzz_polyhedral_8", "fuzz_polyhedral_big4"])
def test_unseeded_cubes_take_their_topology_from_the_first_query(oracle_mod, name):
    """Non-simplicial polyhedra WITHOUT a seeded face topology.  The reference hulls a site at its first query
    (FaceTopologyCache, polyhedral_site.py:138-159, 468-471) and Qhull's triangulation of a cube's quads depends on that
    frame's vertex coordinates -- the hull of the engine's first frame differs from the reference's topology for up to 18 of
    the 27 sites of the fuzz fixtures.  The replay kernel reports the coordinates of every site's first query, the host hulls
    from exactly those and redoes the batch until no triangulation changes: rows, Counters and the triangulations
    themselves must be the reference's."""
    case = oracle_mod.load_case(name)
    t = trajectory_from_case(case, seed_topology=False)
    rows = t.trajectory_from_arrays(case["frac"], case["lattice"])
    eng = t.site_collection._engine
    assert eng._lazy_mode
    assert np.array_equal(rows, case["atoms_traj"])
    got_tr, want_tr = transitions_of(t), oracle_mod.transitions_from_case(case)
    assert got_tr == want_tr
    for k in got_tr:
        assert list(got_tr[k]) == list(want_tr[k])
    queried = 0
    for s in range(len(case["n_face"])):
        nf = int(case["n_face"][s])
        if nf > 0:                                             # the reference queried (and hulled) this site
            assert not eng._open_topology[s]
            assert np.array_equal(np.asarray(eng.simplices[s]), case["simplices"][s, :nf]), (name, s)
            queried += 1
        else:
            assert eng._open_topology[s]
    print(name, "sites queried:", queried, "topology passes:", eng.last_info.get("topology_passes"))
