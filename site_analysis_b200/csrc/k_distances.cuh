// k_distances.cuh -- the minimum-image distance matrix as a stand-alone operator.
//
// Drop-in for the reference's all_mic_distances / _all_mic_distances_numba (distances.py:68-110, 147-189): out[i][j] =
// sqrt(min over the 27 images of |((c1[i] - c2[j]) - rint(c1[i] - c2[j]) + s) L|^2), float64, the reference's operation
// order (sab_dbase / sab_mic_q27 in sab_device.cuh), no FMA, IEEE sqrt -- bit-identical to the numba kernel.  The
// assignment kernels never materialise this matrix (they fuse it with the argmin / cut-off test); this operator serves
// the callers that want the matrix itself: coordination search, index mapping and the alignment objective of the
// reference workflow (tools.py:32-111, reference_workflow/structure_aligner.py:105-147).
// FP64-pipe bound: 369 flop per pair against 8 bytes written.  One thread per pair, j fastest (coalesced stores),
// the c1 row and the lattice are warp-uniform loads.
#pragma once
#include "sab_device.cuh"

__global__ void __launch_bounds__(256)
k_mic_distances(const double *__restrict__ c1, int n, const double *__restrict__ c2, int m,
                const double *__restrict__ lattice, double *__restrict__ out)
{
    __shared__ double L[9];
    if (threadIdx.x < 9) L[threadIdx.x] = lattice[threadIdx.x];
    __syncthreads();
    const long long total = (long long)n * m;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const long long i = p / m;
        const int j = (int)(p - i * m);
        double d0, d1, d2;
        sab_dbase(c1 + 3 * i, c2 + 3 * (long long)j, d0, d1, d2);
        out[p] = sqrt(sab_mic_q27(d0, d1, d2, L));
    }
}

// One thread per site: does any of its query points lie inside its polyhedron?  (PolyhedralSite.contains_point,
// polyhedral_site.py:431-520; arithmetic of _numba_update_faces :26-81 and _numba_sn_query :83-123.)
#include "k_polyhedral.cuh"
__global__ void k_polyhedra_contain(const double *__restrict__ vertex, const int32_t *__restrict__ n_vert, int vmax,
                                    const int32_t *__restrict__ simplices, const int32_t *__restrict__ n_face, int fmax,
                                    int n_sites, const double *__restrict__ points, int n_img, uint8_t *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sites) return;
    double vc[SAB_MAX_VERT][3], cen[3];
    const int nv = n_vert[s], nf = n_face[s];
    for (int v = 0; v < nv; v++)
        for (int d = 0; d < 3; d++) vc[v][d] = vertex[((size_t)s * vmax + v) * 3 + d];
    for (int d = 0; d < 3; d++) {
        double sum = 0.0;
        for (int v = 0; v < nv; v++) sum += vc[v][d];          // 0.0 + v0 + v1 + ... (polyhedral_site.py:46-49)
        cen[d] = sum / (double)nv;
    }
    const int32_t *simp = simplices + (size_t)s * fmax * 3;
    bool inside = false;
    for (int i = 0; i < n_img && !inside; i++) {
        const double *p = points + ((size_t)s * n_img + i) * 3;
        bool ok = true;
        for (int j = 0; j < nf && ok; j++)
            ok = sab_face_accepts(vc[simp[3 * j]], vc[simp[3 * j + 1]], vc[simp[3 * j + 2]], cen, p[0], p[1], p[2]);
        inside = ok;
    }
    out[s] = inside ? 1 : 0;
}
