"""The CPU oracle (oracle/site_oracle.c) against every golden fixture generated from the
unmodified reference (oracle/gen_golden.py).  Bit-exact: site indices, None positions,
transition Counters incl. insertion order, recent-site history, Qhull topology,
dynamic-Voronoi centres (float64 bit patterns)."""
import numpy as np
import pytest

from conftest import golden_names


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference(oracle_mod, name):
    case = oracle_mod.load_case(name)
    o = oracle_mod.SiteOracle.from_case(case)
    out = o.run(case["frac"], case["lattice"])
    assert out.shape == case["atoms_traj"].shape
    assert np.array_equal(out, case["atoms_traj"])
    want = oracle_mod.transitions_from_case(case)
    got = o.transitions()
    assert got == want
    for k in got:                                   # Counter insertion order (site.py:223 ties)
        assert list(got[k]) == list(want[k])
    assert np.array_equal(o.recent(), case["recent"])
    if "n_face" in case:                            # polyhedral_site.py:154-155 via the callback
        for s in range(o.S):
            t = o.topology(s)
            nf = int(case["n_face"][s])
            assert (t is None) == (nf < 0)
            if t is not None:
                assert np.array_equal(t, case["simplices"][s, :nf])


@pytest.mark.parametrize("name", golden_names("fuzz_dynvor") + ["argyrodite_0p_dynvor"])
def test_oracle_dynamic_voronoi_centres_bitwise(oracle_mod, name):
    case = oracle_mod.load_case(name)
    o = oracle_mod.SiteOracle.from_case(case)
    lat = case["lattice"]
    frames = list(case["centre_frames"])
    for f in range(case["frac"].shape[0]):
        o.run(case["frac"][f:f + 1], lat if lat.ndim == 2 else lat[f:f + 1])
        if f in frames:
            want = case["centres_at"][frames.index(f)]
            assert np.array_equal(o.centres().view(np.int64), want.view(np.int64))


def test_readme_quickstart_known_answer(oracle_mod):
    """README.md:33-49: [[0],[0],[None],[None],[1]]."""
    case = oracle_mod.load_case("readme_spherical")
    out = oracle_mod.SiteOracle.from_case(case).run(case["frac"], case["lattice"])
    assert out[:, 0].tolist() == [0, 0, -1, -1, 1]


def test_quickstart_simple_cubic_known_answers(oracle_mod):
    """docs/source/quickstart.ipynb cell 15: [21]*6 + [37]*5 for spherical and Voronoi."""
    for name in ("simple_cubic_spherical", "simple_cubic_voronoi"):
        case = oracle_mod.load_case(name)
        out = oracle_mod.SiteOracle.from_case(case).run(case["frac"], case["lattice"])
        assert out[:, 0].tolist() == [21] * 6 + [37] * 5


@pytest.mark.parametrize("tag,want", [
    ("0p", {"5": 80.20, "4": 0.02, "3": 0.00, "2": 19.78, "1": 0.01}),
    ("50p", {"5": 65.92, "4": 2.63, "3": 0.00, "2": 31.43, "1": 0.02}),
    ("100p", {"5": 53.39, "4": 7.33, "3": 0.00, "2": 39.28, "1": 0.00}),
])
def test_argyrodite_notebook_occupations(oracle_mod, tag, want):
    """tutorials/argyrodite_site_analysis.ipynb cells 9/11/13/14 (percent by site type)."""
    case = oracle_mod.load_case(f"argyrodite_{tag}_polyhedral")
    out = oracle_mod.SiteOracle.from_case(case).run(case["frac"], case["lattice"])
    label = {int(i): str(l) for i, l in zip(case["site_index"], case["labels"])}
    flat = out.ravel()
    for lab, pct in want.items():
        got = 100.0 * sum(1 for v in flat if v >= 0 and label[int(v)] == lab) / flat.size
        assert abs(got - pct) < 0.006


def test_reset_keeps_topology_and_replays_identically(oracle_mod):
    """trajectory.py:364-373, polyhedral_site.py:264-279: reset clears PBC caches, keeps faces."""
    case = oracle_mod.load_case("fuzz_polyhedral_1")
    o = oracle_mod.SiteOracle.from_case(case)
    a = o.run(case["frac"], case["lattice"])
    topo = [o.topology(s) for s in range(o.S)]
    o.reset()
    assert o.transitions() == {}
    assert all((o.topology(s) is None) == (topo[s] is None) for s in range(o.S))
    b = o.run(case["frac"], case["lattice"])
    assert np.array_equal(a, b)
