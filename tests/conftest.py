"""Shared fixtures.  ``-m "not gpu"`` runs everywhere; ``-m gpu`` needs a B200."""
from __future__ import annotations

import glob
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names(prefix=""):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(str(GOLDEN / "*.npz")))
    # frames_* are raw trajectories; folds_* / distances_* / workflow_* / queries_* belong to test_folds / test_distances / test_workflow / test_site_queries
    return [n for n in names if not n.startswith(("frames_", "folds_", "distances_", "workflow_", "queries_")) and n.startswith(prefix)]


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle  # oracle/oracle.py -- the CPU checker (test infrastructure)
    oracle.build()
    return oracle
