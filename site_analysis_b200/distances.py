"""Minimum-image distances on the GPU (mirror of reference ``site_analysis/distances.py``).

``all_mic_distances`` / ``mic_distance`` have the reference's signatures (distances.py:113-189) and return
bit-identical float64 values (27 periodic images, the numba kernel's operation order, no FMA); the work happens in
``sab_mic_distances`` (``csrc/k_distances.cuh``).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def all_mic_distances_device(d_c1, d_c2, d_lattice):
    """CUDA tensors in ((n, 3), (m, 3), (3, 3) float64), CUDA tensor (n, m) out; stream-ordered, no sync."""
    import torch
    L = nat.lib()
    d_c1, d_c2, d_lattice = (t.to(torch.float64).contiguous() for t in (d_c1, d_c2, d_lattice))
    n, m = int(d_c1.shape[0]), int(d_c2.shape[0])
    out = torch.empty((n, m), dtype=torch.float64, device=d_c1.device)
    stream = torch.cuda.current_stream(d_c1.device).cuda_stream
    nat.check(L.sab_mic_distances(d_c1.device.index, d_c1.data_ptr(), n, d_c2.data_ptr(), m, d_lattice.data_ptr(),
                                  out.data_ptr(), C.c_void_p(stream)))
    return out


def all_mic_distances(frac_coords1: np.ndarray, frac_coords2: np.ndarray, lattice_matrix: np.ndarray) -> np.ndarray:
    """(N, M) minimum-image distance matrix (reference distances.py:147-189)."""
    c1 = np.ascontiguousarray(frac_coords1, dtype=np.float64).reshape(-1, 3)
    c2 = np.ascontiguousarray(frac_coords2, dtype=np.float64).reshape(-1, 3)
    if c1.shape[0] == 0 or c2.shape[0] == 0:
        return np.zeros((c1.shape[0], c2.shape[0]))
    import torch
    if not torch.cuda.is_available():
        raise nat.NativeError("site_analysis_b200 needs a CUDA device: there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    lat = np.ascontiguousarray(lattice_matrix, dtype=np.float64).reshape(3, 3)
    out = all_mic_distances_device(torch.from_numpy(c1).to(dev), torch.from_numpy(c2).to(dev), torch.from_numpy(lat).to(dev))
    return out.cpu().numpy()


def mic_distance(frac1: np.ndarray, frac2: np.ndarray, lattice_matrix: np.ndarray) -> float:
    """Minimum-image distance of one pair (reference distances.py:113-144)."""
    return float(all_mic_distances(np.asarray(frac1)[None, :], np.asarray(frac2)[None, :], lattice_matrix)[0, 0])


_SHIFTS8 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], dtype=np.float64)


def x_pbc(x: np.ndarray) -> np.ndarray:
    """The 8 images x + {0,1}^3 in the reference's order (tools.py:246-279)."""
    return _SHIFTS8 + np.asarray(x, dtype=np.float64)


def polyhedra_contain(vertex_coords, simplices, points) -> np.ndarray:
    """For each site: does any of its query points lie inside its polyhedron?  (``sab_polyhedra_contain``.)

    ``vertex_coords``: per site (V, 3) PBC-corrected vertices; ``simplices``: per site (Fc, 3) hull triangles (local vertex
    numbers); ``points``: per site (n_img, 3) query points (same n_img for every site).  Returns a bool array (n_sites,).
    The arithmetic is the reference's ``_numba_update_faces`` + ``_numba_sn_query`` (polyhedral_site.py:26-123), float64, on
    the GPU; there is no CPU fallback."""
    import torch
    if not torch.cuda.is_available():
        raise nat.NativeError("site_analysis_b200 needs a CUDA device: there is no CPU fallback")
    S = len(vertex_coords)
    if S == 0:
        return np.zeros(0, dtype=bool)
    n_vert = np.array([len(v) for v in vertex_coords], dtype=np.int32)
    n_face = np.array([len(f) for f in simplices], dtype=np.int32)
    vmax, fmax = int(n_vert.max()), int(n_face.max())
    if vmax > 16:
        raise ValueError("polyhedra with more than 16 vertices are not supported by the GPU containment query")
    vert = np.zeros((S, vmax, 3))
    simp = np.zeros((S, fmax, 3), dtype=np.int32)
    for s in range(S):
        vert[s, :n_vert[s]] = vertex_coords[s]
        simp[s, :n_face[s]] = simplices[s]
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(S, -1, 3))
    dev = torch.device("cuda", torch.cuda.current_device())
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (vert, n_vert, simp, n_face, pts)]
    out = torch.empty(S, dtype=torch.uint8, device=dev)
    L = nat.lib()
    nat.check(L.sab_polyhedra_contain(dev.index, d[0].data_ptr(), d[1].data_ptr(), vmax, d[2].data_ptr(), d[3].data_ptr(), fmax, S,
                                      d[4].data_ptr(), int(pts.shape[1]), out.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out.cpu().numpy().astype(bool)
