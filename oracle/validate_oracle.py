"""Pin the C oracle against the LIVE reference (build container only).

TEST INFRASTRUCTURE.  Imports the unmodified reference from /root/reference through
the stand-ins in oracle/refharness and checks, bit for bit:

  1. distances.py `_all_mic_distances_numba` / `_mic_distance_numba` vs so_all_mic_distances
     on random triclinic cells (float64 view equality),
  2. full trajectories of all four site types on freshly generated fuzz inputs
     (seeds disjoint from the committed fixtures), via oracle/gen_golden.py's generators.

Usage: python oracle/validate_oracle.py [n_seeds]
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import gen_golden  # noqa: E402
import oracle  # noqa: E402
import refcases  # noqa: E402


def check_distances(n_cases=200):
    from site_analysis.distances import all_mic_distances, mic_distance
    rng = np.random.default_rng(7)
    bad = 0
    for i in range(n_cases):
        L = gen_golden._triclinic(rng, 3.0, 12.0, 35.0)
        a = rng.uniform(-1.5, 2.5, (17, 3)); b = rng.uniform(0, 1, (13, 3))
        if i % 5 == 0:
            a = np.round(a * 4) / 4; b = np.round(b * 4) / 4     # exact ties between images
        ref = all_mic_distances(a, b, L)
        mine = oracle.all_mic_distances(a, b, L)
        bad += int((ref.view(np.int64) != mine.view(np.int64)).sum())
        one = mic_distance(a[0], b[0], L)
        bad += int(np.float64(one).view(np.int64) != mine[0, 0].view(np.int64))
    print(f"distances: {n_cases} cells x 221 pairs, bitwise mismatches = {bad}")
    return bad


def check_fuzz(n_seeds):
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        gen_golden.GOLDEN = Path(tmp)
        for s in range(100, 100 + n_seeds):
            gen_golden.fuzz_spherical_voronoi(s)
            gen_golden.fuzz_polyhedral(s)
            gen_golden.fuzz_dynvor(s)
        for p in sorted(Path(tmp).glob("*.npz")):
            case = oracle.load_case(p.stem, tmp)
            o = oracle.SiteOracle.from_case(case)
            out = o.run(case["frac"], case["lattice"])
            m = int((out != case["atoms_traj"]).sum())
            t = o.transitions() != oracle.transitions_from_case(case)
            if m or t:
                print(f"  MISMATCH {p.stem}: assignments {m}, transitions differ {t}")
            bad += m + int(t)
    print(f"fuzz: {n_seeds} seeds x 4 site types, mismatching assignments = {bad}")
    return bad


if __name__ == "__main__":
    refcases.import_reference()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    total = check_distances() + check_fuzz(n)
    print("ORACLE PINNED" if total == 0 else f"ORACLE DIFFERS ({total})")
    sys.exit(1 if total else 0)
