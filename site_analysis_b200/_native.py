"""ctypes binding of ``libsab200.so`` (C ABI declared in ``include/sab200.h``).

The shared library is built in-tree (``site_analysis_b200/csrc/build.sh`` or
``__graft_entry__.build()``).  There is no CPU fallback: if the library is missing
or no CUDA device is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libsab200.so"
ABI_VERSION = 1

SPHERICAL, VORONOI, POLYHEDRAL, DYNAMIC_VORONOI = 0, 1, 2, 3
NONE = -1
AMBIGUOUS_BASE = -2
OPT_SKIP_STEP_CHECK_ONCE = 1
OPT_POLY_KERNEL = 2
OPT_SEQ_CHUNK = 3
OPT_SEQ_BLOCK = 4
OPT_FP32_FILTER = 5
OPT_TILE_FRAMES = 6
OPT_STREAM_PRODUCERS = 7
OPT_STREAM_REC_SMEM = 8
OPT_STREAM_REG_COLS = 9
OPT_STREAM_GEN = 10
OPT_STREAM5_ROWS = 11
OPT_STREAM5_RAW = 12
OPT_STREAM5_BLOCKS = 13
OPT_RESOLVER = 15

# every symbol include/sab200.h declares (tests/test_host_logic.py checks the header against this)
SYMBOLS = [
    "sab_abi_version", "sab_last_error", "sab_engine_create", "sab_engine_destroy", "sab_set_centres",
    "sab_set_polyhedra", "sab_set_reference_atoms", "sab_framework_count", "sab_framework_atoms",
    "sab_set_pbc_state", "sab_get_pbc_state", "sab_set_legacy_init", "sab_set_option",
    "sab_workspace_bytes", "sab_wrap_totals", "sab_assign_device", "sab_commit_pbc", "sab_assign_host", "sab_set_result_mirror",
    "sab_set_priority_tables", "sab_add_rank_rows", "sab_rank_missing", "sab_resolve", "sab_dynvor_centres", "sab_set_cell_lists", "sab_scan_excursion", "sab_set_nearest_lists", "sab_set_dynvor_lists", "sab_set_sphere_lists",
    "sab_set_carry_device", "sab_get_excursion_device", "sab_fold_workspace_bytes", "sab_fold_result", "sab_mic_distances", "sab_polyhedra_contain", "sab_measure_fp64_peak", "sab_lazy_set_state", "sab_lazy_get_state", "sab_assign_lazy", "sab_lazy_set_open_topology", "sab_lazy_get_first_queries", "sab_update_face_topology", "sab_set_result_peers", "sab_set_result_multicast", "sab_xdatcar_index", "sab_xdatcar_parse",
    "sab_parse_double",
]


class NativeError(RuntimeError):
    """A libsab200 call failed (CUDA error, bad argument, overflow ...)."""


def build(verbose: bool = False) -> Path:
    """Compile libsab200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = [str(PKG / "csrc" / "build.sh")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libsab200.so failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout, res.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA backend.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SAB200_LIB", LIB_PATH))       # SAB200_LIB: an alternative BUILD of the same backend (A/B timing)
    if not path.exists():
        raise NativeError(
            f"{path} is missing: the CUDA backend has not been built "
            "(run site_analysis_b200/csrc/build.sh or __graft_entry__.build()). "
            "site_analysis_b200 has no CPU fallback.")
    L = C.CDLL(str(path))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    sig = {
        "sab_abi_version": (i32, []),
        "sab_last_error": (C.c_char_p, []),
        "sab_engine_create": (i32, [C.POINTER(vp), i32, i32, i32, i32, vp, i32]),
        "sab_engine_destroy": (None, [vp]),
        "sab_set_centres": (i32, [vp, vp, vp, vp]),
        "sab_set_polyhedra": (i32, [vp, i32, vp, vp, i32, vp, vp]),
        "sab_set_reference_atoms": (i32, [vp, i32, vp, vp]),
        "sab_framework_count": (i32, [vp]),
        "sab_framework_atoms": (i32, [vp, vp]),
        "sab_set_pbc_state": (i32, [vp, vp, vp, vp, vp]),
        "sab_get_pbc_state": (i32, [vp, vp, vp, vp, vp]),
        "sab_set_legacy_init": (i32, [vp, vp, i64]),
        "sab_set_option": (i32, [vp, i32, i64]),
        "sab_workspace_bytes": (i64, [vp, i64]),
        "sab_wrap_totals": (i32, [vp, vp, i64, vp, i64, vp, vp, vp]),
        "sab_assign_device": (i32, [vp, vp, vp, i32, i64, vp, vp, i64, vp, i64, vp, vp]),
        "sab_commit_pbc": (i32, [vp, vp, i64, i64, vp, i64, vp]),
        "sab_assign_host": (i32, [vp, vp, vp, i32, i64, vp, i64, vp]),
        "sab_set_priority_tables": (i32, [vp, vp, vp, vp, vp]),
        "sab_resolve": (i32, [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32]),
        "sab_dynvor_centres": (i32, [vp, vp, i64, vp, vp, i64, vp]),
        "sab_set_cell_lists": (i32, [vp, i32, C.c_double, vp, i32, i32, vp, vp, vp]),
        "sab_scan_excursion": (i32, [vp, vp, i64, vp, i64, vp, vp]),
        "sab_set_nearest_lists": (i32, [vp, i32, vp, vp, vp, vp]),
        "sab_set_dynvor_lists": (i32, [vp, i32, vp, vp, vp, C.c_double, vp, vp, vp]),
        "sab_set_sphere_lists": (i32, [vp, i32, vp, C.c_double, vp, vp]),
        "sab_set_carry_device": (i32, [vp, vp, vp, vp]),
        "sab_get_excursion_device": (i32, [vp, vp, vp]),
        "sab_set_result_peers": (i32, [vp, i32, vp]),
        "sab_set_result_multicast": (i32, [vp, vp]),
        "sab_lazy_set_open_topology": (i32, [vp, vp]),
        "sab_lazy_get_first_queries": (i32, [vp, vp, vp]),
        "sab_update_face_topology": (i32, [vp, i32, vp, vp]),
        "sab_xdatcar_index": (i32, [vp, i64, i64, i32, vp, vp, i64, vp]),
        "sab_xdatcar_parse": (i32, [vp, i64, i32, i64, vp, vp, vp, vp, i32]),
        "sab_parse_double": (i32, [C.c_char_p, i64, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
        "sab_mic_distances": (i32, [i32, vp, i32, vp, i32, vp, vp, vp]),
        "sab_lazy_set_state": (i32, [vp, vp, vp, vp, vp, vp]),
        "sab_lazy_get_state": (i32, [vp, vp, vp, vp, vp, vp]),
        "sab_assign_lazy": (i32, [vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, i32, i32]),
        "sab_set_result_mirror": (i32, [vp, vp]),
        "sab_add_rank_rows": (i32, [vp, i32, vp, vp]),
        "sab_rank_missing": (i32, [vp, vp, i32]),
        "sab_measure_fp64_peak": (i32, [i32, vp]),
        "sab_polyhedra_contain": (i32, [i32, vp, vp, i32, vp, vp, i32, i32, vp, i32, vp, vp]),
        "sab_fold_workspace_bytes": (i64, [i64, i32, i32, i64, i64]),
        "sab_fold_result": (i32, [i32, vp, i64, i32, i32, vp, i64, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.sab_abi_version() != ABI_VERSION:
        raise NativeError(f"libsab200.so ABI {L.sab_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        raise NativeError(lib().sab_last_error().decode() or f"libsab200 error {code}")


def ptr(a) -> int | None:
    """Host pointer of a C-contiguous numpy array (None passes NULL)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return a.ctypes.data
