"""Minimum-image distance operator (reference distances.py:27-189): the CPU oracle and
the CUDA operator against float64 BIT PATTERNS produced by the unmodified reference's numba kernels
(oracle/gen_golden_distances.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(str(GOLDEN / "distances_*.npz")))


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


@pytest.mark.parametrize("name", CASES)
def test_oracle_distances_bitwise(oracle_mod, name):
    case = np.load(GOLDEN / f"{name}.npz")
    got = oracle_mod.all_mic_distances(case["c1"], case["c2"], case["lattice"])
    assert np.array_equal(_bits(got), _bits(case["out"]))


def test_cases_present():
    assert len(CASES) == 3


@pytest.mark.gpu
@pytest.mark.timeout(120)
@pytest.mark.parametrize("name", CASES)
def test_gpu_distances_bitwise_against_reference(name):
    from site_analysis_b200 import distances as d
    case = np.load(GOLDEN / f"{name}.npz")
    got = d.all_mic_distances(case["c1"], case["c2"], case["lattice"])
    assert got.shape == case["out"].shape and np.array_equal(_bits(got), _bits(case["out"]))
    for i in range(3):
        assert d.mic_distance(case["c1"][i], case["c2"][i], case["lattice"]) == case["out"][i, i]


@pytest.mark.gpu
@pytest.mark.timeout(180)
def test_gpu_distances_bitwise_against_oracle_at_size(oracle_mod):
    """3328 x 1536 pairs (the 4x4x4 argyrodite supercell of BASELINE config 3), strained triclinic cell."""
    from site_analysis_b200 import distances as d
    rng = np.random.default_rng(11)
    c1, c2 = rng.random((3328, 3)), rng.random((1536, 3)) * 2 - 0.5
    lat = np.array([[40.62, 1e-4, 0.0], [2e-4, 40.7, -1e-4], [0.3, 0.0, 40.5]])
    got = d.all_mic_distances(c1, c2, lat)
    want = oracle_mod.all_mic_distances(c1, c2, lat)
    assert np.array_equal(_bits(got), _bits(want))
    assert d.all_mic_distances(np.zeros((0, 3)), c2, lat).shape == (0, 1536)
