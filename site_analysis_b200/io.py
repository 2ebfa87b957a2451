"""Array-native trajectory ingestion.

The reference consumes ``Sequence[Structure]`` (reference trajectory.py:413-432); users build that list with
pymatgen's ``Xdatcar`` -- one Python ``float()`` per coordinate and one ``Structure`` per frame.  ``read_xdatcar`` turns
the file into the arrays ``Trajectory.trajectory_from_arrays`` / ``sab_assign_host`` take, with the text parsed by the
native, multi-threaded parser of ``libsab200.so`` (``csrc/xdatcar_host.h``): numbers are converted exactly as
``float()`` would, so the arrays are bit-identical to ``np.array([s.frac_coords for s in Xdatcar(path).structures])``.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os

import numpy as np

from . import _native as nat


def _pinned_empty(shape, dtype):
    """Page-locked host array when a CUDA device is present (H2D copies then run at full PCIe speed), else pageable."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(shape, dtype=dtype)


def parse_xdatcar_text(text: bytes, n_threads: int | None = None, pinned: bool = False):
    """XDATCAR text (bytes) -> (species list, lattice (3, 3) or (F, 3, 3), frac (F, N, 3) float64).

    VASP 5 layout: comment, scale, three lattice vectors, element symbols, element counts, then per frame a
    ``Direct configuration=`` line followed by N coordinate lines; variable-cell runs repeat the 7-line header in
    front of every frame.  A negative scale is the cell volume.  Raises ``ValueError`` on anything else."""
    L = nat.lib()
    body_off = 0
    for _ in range(7):                                          # end of the 7-line header, without copying the body
        nl = text.find(b"\n", body_off)
        if nl < 0:
            raise ValueError("XDATCAR: fewer than 7 header lines")
        body_off = nl + 1
    head = bytes(text[:body_off]).split(b"\n")
    try:
        scale = float(head[1].split()[0])
        vectors = np.array([[float(x) for x in head[i].split()[:3]] for i in (2, 3, 4)], dtype=np.float64)
        symbols = [s.decode() for s in head[5].split()]
        counts = [int(x) for x in head[6].split()]
    except (ValueError, IndexError) as exc:
        raise ValueError(f"XDATCAR: malformed header ({exc})") from None
    if vectors.shape != (3, 3) or not symbols or len(symbols) != len(counts):
        raise ValueError("XDATCAR: expected three lattice vectors, an element-symbol line and a matching count line (VASP 5 format)")
    if scale < 0.0:
        scale = float(np.cbrt(-scale / abs(np.linalg.det(vectors))))
    lattice0 = vectors * scale
    species = [s for s, c in zip(symbols, counts) for _ in range(c)]
    N = len(species)
    buf = np.frombuffer(text, dtype=np.uint8)
    n_frames = C.c_int64(0)
    cap = len(text) // max(6 * N, 1) + 16
    while True:
        frame_off = np.empty(cap, dtype=np.int64)
        header_off = np.empty(cap, dtype=np.int64)
        rc = L.sab_xdatcar_index(buf.ctypes.data, len(text), body_off, N, nat.ptr(frame_off), nat.ptr(header_off), cap,
                                 C.byref(n_frames))
        if rc == 3:                                             # SAB_EOVERFLOW: more frames than guessed
            cap = int(n_frames.value)
            continue
        if rc != 0:
            raise ValueError(L.sab_last_error().decode())
        break
    F = int(n_frames.value)
    frame_off, header_off = frame_off[:F].copy(), header_off[:F].copy()
    variable = bool(F and (header_off >= 0).any())
    frac = (_pinned_empty if pinned else np.empty)((F, N, 3), np.float64)
    lat = np.empty((F, 3, 3), dtype=np.float64) if variable else None
    threads = n_threads or min(os.cpu_count() or 1, 32)
    rc = L.sab_xdatcar_parse(buf.ctypes.data, len(text), N, F, nat.ptr(frame_off), nat.ptr(header_off), nat.ptr(frac),
                             None if lat is None else nat.ptr(lat), threads)
    if rc != 0:
        raise ValueError(L.sab_last_error().decode())
    if not variable:
        return species, lattice0, frac
    src = np.maximum.accumulate(np.where(header_off >= 0, np.arange(F), -1))   # a frame without its own header keeps the previous cell
    out = np.where((src >= 0)[:, None, None], lat[np.clip(src, 0, None)], lattice0[None])
    return species, out, frac


def read_xdatcar(path, n_threads: int | None = None, pinned: bool = False):
    """Read an XDATCAR (optionally ``.gz``) into ``(species, lattice, frac)``; see :func:`parse_xdatcar_text`."""
    path = os.fspath(path)
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            text = f.read()
    else:
        with open(path, "rb") as f:
            text = f.read()
    return parse_xdatcar_text(text, n_threads=n_threads, pinned=pinned)
