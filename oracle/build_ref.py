"""Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference's own Python package, so that the real
reference (Python + numba) can be timed on the GPU box, where ``/root/reference`` does not exist.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  ``__graft_entry__.build()`` runs this in the build container; the copy lives only
under ``oracle/_ref/`` (listed in .gitignore, not in .gpurunignore: it travels with the snapshot like a built .so).  Nothing
under ``site_analysis_b200/`` imports it; ``bench.py --impl reference`` and bench.py's ``cpu_baseline`` leg execute it in
worker processes (``oracle/ref_worker.py``) through the pymatgen / monty stand-ins in ``oracle/refharness``.
"""
from __future__ import annotations

import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference")


def build_ref(force: bool = False) -> Path | None:
    dst = HERE / "_ref"
    if not (SRC / "site_analysis").is_dir():
        return dst if (dst / "site_analysis").is_dir() else None
    stamp = dst / "STAMP.json"
    if stamp.exists() and not force:
        return dst
    if dst.exists():
        shutil.rmtree(dst)
    dst.mkdir(parents=True)
    shutil.copytree(SRC / "site_analysis", dst / "site_analysis", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    version = "?"
    for line in (SRC / "site_analysis" / "version.py").read_text().splitlines() if (SRC / "site_analysis" / "version.py").exists() else []:
        if "__version__" in line:
            version = line.split("=")[-1].strip().strip("'\"")
    stamp.write_text(json.dumps({"source": str(SRC / "site_analysis"), "version": version,
                                 "files": sum(1 for _ in (dst / "site_analysis").rglob("*.py"))}) + "\n")
    return dst


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
