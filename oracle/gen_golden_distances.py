"""Golden minimum-image distance matrices from the UNMODIFIED reference (numba path of
``site_analysis.distances.all_mic_distances`` / ``mic_distance``, distances.py:27-189).

TEST INFRASTRUCTURE ONLY.  Run in the build container:  python oracle/gen_golden_distances.py
Writes tests/golden/distances_*.npz: inputs and the reference's float64 outputs (compared as bit patterns).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refcases  # noqa: E402

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    refcases.import_reference()
    from site_analysis import distances as ref
    assert ref.HAS_NUMBA
    rng = np.random.default_rng(20260109)
    cases = {
        "distances_cubic": (np.eye(3) * 10.155, 37, 29),
        "distances_triclinic": (np.array([[7.1, 0.3, -0.2], [2.9, 6.4, 0.5], [-1.7, 2.2, 8.3]]), 41, 53),
        "distances_sheared": (np.array([[5.0, 0.0, 0.0], [4.4, 2.1, 0.0], [3.9, 1.8, 2.5]]), 33, 31),   # needs the 27 images
    }
    for name, (lat, n, m) in cases.items():
        c1 = rng.random((n, 3)) * 3.0 - 1.0                      # unwrapped coordinates too
        c2 = rng.random((m, 3))
        c1[0] = [0.0, 0.5, 1.0]; c2[0] = [0.5, 0.0, 0.0]         # exact half-cell separations (rint ties)
        c1[1] = c2[1]                                            # zero distance
        c2[2] = c1[2] + np.array([1.0, -2.0, 0.0])               # periodic image of the same point
        out = ref.all_mic_distances(c1, c2, lat)
        single = np.array([ref.mic_distance(c1[i], c2[i], lat) for i in range(5)])
        assert np.array_equal(single.view(np.int64), out[np.arange(5), np.arange(5)].view(np.int64))
        np.savez_compressed(GOLDEN / f"{name}.npz", c1=c1, c2=c2, lattice=lat, out=out)
        print(f"  wrote {name}.npz {out.shape}")


if __name__ == "__main__":
    main()
