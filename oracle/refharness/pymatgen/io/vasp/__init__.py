"""Minimal XDATCAR reader standing in for ``pymatgen.io.vasp.Xdatcar`` (test infrastructure)."""
from __future__ import annotations

import gzip

import numpy as np

from pymatgen.core import Lattice, Structure


def read_xdatcar_arrays(path):
    """Return (species list, lattice (3,3), frac (F,N,3)) from a constant-cell XDATCAR."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        lines = f.read().split("\n")
    scale = float(lines[1].split()[0])
    lattice = np.array([[float(x) for x in lines[i].split()] for i in (2, 3, 4)]) * scale
    symbols = lines[5].split()
    counts = [int(x) for x in lines[6].split()]
    species = [s for s, c in zip(symbols, counts) for _ in range(c)]
    n = sum(counts)
    frames = []
    i = 7
    while i < len(lines):
        if lines[i].strip().startswith("Direct"):
            block = lines[i + 1:i + 1 + n]
            frames.append([[float(x) for x in ln.split()[:3]] for ln in block])
            i += n + 1
        else:
            i += 1
    return species, lattice, np.array(frames, dtype=np.float64)


class Xdatcar:
    def __init__(self, path):
        species, lattice, frac = read_xdatcar_arrays(path)
        self.structures = [Structure(Lattice(lattice), species, fr) for fr in frac]
