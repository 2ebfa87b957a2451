// k_polyhedral.cuh -- polyhedral containment for every (frame, atom): exhaustive variant.
//
// Replaces, for a whole frame at a time,
//   PolyhedralSite._store_vertex_coords      polyhedral_site.py:377-416 (vertex gather + PBC shifts)
//   pbc_utils._numba_update_pbc_shifts       pbc_utils.py:218-229      (raw + shift, non-negative offset)
//   _numba_update_faces                      polyhedral_site.py:26-81  (centroid, normals, centre signs)
//   _numba_sn_query over x_pbc images        polyhedral_site.py:83-123, tools.py:246-279
// and emits the full set of containing sites per atom-frame (none / one / spilled list), from
// which the reference's first-hit-in-priority-order answer (polyhedral_site_collection.py:
// 99-107) follows: uniquely when the set has <= 1 member, through k_resolve.cuh otherwise.
//
// Layout: one CTA per frame (grid-stride).  SITE-major: each thread owns a polyhedron, builds
// its unwrapped vertices in registers, then walks the wrapped mobile coordinates staged in
// shared memory.  An axis-aligned bounding box of the vertices, widened by SAB_BBOX_EPS,
// rejects almost every (site, atom, image) before any face arithmetic; a point outside the
// widened box is outside the hull by >= 1e-9 fractional units, far beyond the rounding error
// of the face test, so skipping it cannot change the floating-point verdict.
#pragma once
#include "sab_device.cuh"

#define SAB_MAX_VERT 16
#define SAB_POLY_KEEP 4

struct SabPolyTables {
    int smax, fmax;
    const int32_t *n_slot;      // [S]
    const int32_t *slot_atom;   // [S][smax] structure index
    const int32_t *slot_fw;     // [S][smax] compact framework id
    const int32_t *slot_shift;  // [S][smax][3]
    const int32_t *n_face;      // [S]
    const uint8_t *simp;        // [S][fmax][3] local vertex numbers
    const uint8_t *legacy_init; // [S] or null: site initialised by the legacy spread rule
};

// Unwrapped vertices of one polyhedron in one frame + bounding box + centroid.
__device__ __forceinline__ void sab_poly_vertices(const double *__restrict__ frame, const SabPolyTables &T,
                                                  int s, const int32_t *__restrict__ wrow, bool legacy_now,
                                                  double (*vc)[3], int &nv, double *lo, double *hi, double *cen,
                                                  const double *__restrict__ anchor = nullptr)
{
    // wrow == nullptr: the wrap count comes from the anchor formulation W = rint(raw - anchor) of the stream
    // kernel (k_polyhedral_stream.cuh), which equals the scanned count while its chain check holds
    nv = T.n_slot[s];
    const int32_t *sa = T.slot_atom + (size_t)s * T.smax;
    const int32_t *sf = T.slot_fw + (size_t)s * T.smax;
    const int32_t *ss = T.slot_shift + (size_t)s * T.smax * 3;
    for (int v = 0; v < nv; v++)
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            const double raw = frame[(size_t)sa[v] * 3 + d];
            const int W = wrow ? wrow[3 * sf[v] + d] : (int)rint(raw - anchor[3 * sf[v] + d]);
            vc[v][d] = raw + (double)(ss[3 * v + d] - W);
        }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double mn = vc[0][d];
        for (int v = 1; v < nv; v++) if (vc[v][d] < mn) mn = vc[v][d];
        double u = 0.0;
        if (mn < 0.0 && !legacy_now) u = ceil(-mn);
        double mx = -INFINITY; mn = INFINITY;
        double sum = 0.0;
        for (int v = 0; v < nv; v++) {
            vc[v][d] += u;
            mn = fmin(mn, vc[v][d]); mx = fmax(mx, vc[v][d]);
            sum += vc[v][d];                       // 0.0 + v0 + v1 + ... (polyhedral_site.py:46-49)
        }
        lo[d] = mn - SAB_BBOX_EPS; hi[d] = mx + SAB_BBOX_EPS;
        cen[d] = sum / (double)nv;
    }
}

// All faces against one image point, early exit (polyhedral_site.py:109-123).
__device__ __forceinline__ bool sab_poly_contains_point(const double (*vc)[3], const double *cen,
                                                        const uint8_t *__restrict__ simp, int nf,
                                                        double p0, double p1, double p2)
{
    for (int j = 0; j < nf; j++) {
        int i0 = simp[3 * j], i1 = simp[3 * j + 1], i2 = simp[3 * j + 2];
        if (!sab_face_accepts(vc[i0], vc[i1], vc[i2], cen, p0, p1, p2)) return false;
    }
    return true;
}

// Any of the 8 images x + {0,1}^3 (tools.py:246-253) inside; images outside the box are skipped.
__device__ __forceinline__ bool sab_poly_contains_atom(const double (*vc)[3], const double *cen,
                                                       const uint8_t *__restrict__ simp, int nf,
                                                       const double *lo, const double *hi, const double *x)
{
    double p[3][2]; bool ok[3][2];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        p[d][0] = x[d]; p[d][1] = 1.0 + x[d];
        ok[d][0] = p[d][0] >= lo[d] && p[d][0] <= hi[d];
        ok[d][1] = p[d][1] >= lo[d] && p[d][1] <= hi[d];
        if (!(ok[d][0] | ok[d][1])) return false;
    }
    for (int o0 = 0; o0 < 2; o0++) if (ok[0][o0])
        for (int o1 = 0; o1 < 2; o1++) if (ok[1][o1])
            for (int o2 = 0; o2 < 2; o2++) if (ok[2][o2])
                if (sab_poly_contains_point(vc, cen, simp, nf, p[0][o0], p[1][o1], p[2][o2])) return true;
    return false;
}

__global__ void k_polyhedral_assign_exhaustive(const double *__restrict__ frac, int64_t n_frames, int n_atoms,
                                               int n_mobile, const int32_t *__restrict__ mobile, int n_sites,
                                               SabPolyTables T, const int32_t *__restrict__ wrows, int m3,
                                               int64_t legacy_init_frame, int32_t *__restrict__ result,
                                               int32_t *__restrict__ spill, long long spill_capacity,
                                               unsigned long long *__restrict__ counters)
{
    extern __shared__ double smp[];                  // xs[A*3] | cnt[A] | big[A] | hits[A*KEEP]
    double *xs = smp;
    int *cnt = (int *)(smp + 3 * n_mobile);
    int *big = cnt + n_mobile;                       // spill offset of atoms in more than KEEP sites, else -1
    int *hits = big + n_mobile;
    __shared__ int any_big;
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        const int32_t *wrow = wrows + (size_t)f * m3;
        __syncthreads();
        if (threadIdx.x == 0) any_big = 0;
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
#pragma unroll
            for (int d = 0; d < 3; d++) xs[3 * a + d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);
            cnt[a] = 0; big[a] = -1;
        }
        __syncthreads();
        for (int s = threadIdx.x; s < n_sites; s += blockDim.x) {
            double vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
            int nv;
            bool legacy_now = T.legacy_init && T.legacy_init[s] && f == legacy_init_frame;
            sab_poly_vertices(frame, T, s, wrow, legacy_now, vc, nv, lo, hi, cen);
            const uint8_t *simp = T.simp + (size_t)s * T.fmax * 3;
            int nf = T.n_face[s];
            for (int a = 0; a < n_mobile; a++) {
                if (sab_poly_contains_atom(vc, cen, simp, nf, lo, hi, xs + 3 * a)) {
                    int k = atomicAdd(&cnt[a], 1);
                    if (k < SAB_POLY_KEEP) hits[a * SAB_POLY_KEEP + k] = s;
                }
            }
        }
        __syncthreads();
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
            int c = cnt[a], out;
            int *h = hits + a * SAB_POLY_KEEP;
            if (c == 0) out = SAB_NONE;
            else if (c == 1) out = h[0];
            else {
                atomicAdd(&counters[0], 1ULL);
                unsigned long long off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(c));
                if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) {
                    counters[2] = 1ULL;
                    out = SAB_NONE;
                } else {
                    spill[off] = c;
                    if (c <= SAB_POLY_KEEP) {
                        for (int i = 1; i < c; i++) {          // ascending list positions
                            int v = h[i], j = i;
                            while (j > 0 && h[j - 1] > v) { h[j] = h[j - 1]; j--; }
                            h[j] = v;
                        }
                        for (int i = 0; i < c; i++) spill[off + 1 + i] = h[i];
                    } else {                                   // rare: atom on a shared vertex / edge
                        big[a] = (int)off; cnt[a] = 0; any_big = 1;
                    }
                    out = SAB_AMBIGUOUS_BASE - (int)off;
                }
            }
            result[(size_t)f * n_mobile + a] = out;
        }
        __syncthreads();
        if (any_big) {                                         // second sweep fills the long lists
            for (int s = threadIdx.x; s < n_sites; s += blockDim.x) {
                double vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
                int nv;
                bool legacy_now = T.legacy_init && T.legacy_init[s] && f == legacy_init_frame;
                sab_poly_vertices(frame, T, s, wrow, legacy_now, vc, nv, lo, hi, cen);
                const uint8_t *simp = T.simp + (size_t)s * T.fmax * 3;
                int nf = T.n_face[s];
                for (int a = 0; a < n_mobile; a++)
                    if (big[a] >= 0 && sab_poly_contains_atom(vc, cen, simp, nf, lo, hi, xs + 3 * a))
                        spill[big[a] + 1 + atomicAdd(&cnt[a], 1)] = s;
            }
            __syncthreads();
            for (int a = threadIdx.x; a < n_mobile; a += blockDim.x)
                if (big[a] >= 0) {
                    int32_t *l = spill + big[a] + 1;
                    int c = cnt[a];
                    for (int i = 1; i < c; i++) {
                        int v = l[i], j = i;
                        while (j > 0 && l[j - 1] > v) { l[j] = l[j - 1]; j--; }
                        l[j] = v;
                    }
                }
        }
    }
}
