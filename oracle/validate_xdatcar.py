"""Pin the native XDATCAR parser against the reference's own data files (build container only; TEST INFRASTRUCTURE).

    python oracle/validate_xdatcar.py

Every XDATCAR under /root/reference is read with (a) ``site_analysis_b200.io.read_xdatcar`` and (b) the plain-Python
reader of ``oracle/refharness`` (``float()`` per token, as pymatgen does); coordinates, lattices and species must be
bit-identical."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import refcases  # noqa: E402

refcases.import_reference()
from pymatgen.io.vasp import read_xdatcar_arrays  # noqa: E402  (the refharness stand-in)

from site_analysis_b200.io import read_xdatcar  # noqa: E402

files = sorted(p for p in refcases.REFERENCE_ROOT.rglob("*XDATCAR*") if p.is_file())
assert files, "no XDATCAR files found"
for p in files:
    t0 = time.perf_counter(); sp, lat, frac = read_xdatcar(p); t1 = time.perf_counter()
    sp2, lat2, frac2 = read_xdatcar_arrays(p); t2 = time.perf_counter()
    ok = sp == list(sp2) and np.array_equal(lat.view(np.int64), np.asarray(lat2).view(np.int64)) and \
        frac.shape == frac2.shape and np.array_equal(frac.view(np.int64), frac2.view(np.int64))
    print(f"{'OK ' if ok else 'DIFF'} {p.relative_to(refcases.REFERENCE_ROOT)}: frames {frac.shape} native {1e3 * (t1 - t0):.1f} ms, python {1e3 * (t2 - t1):.1f} ms")
    assert ok
