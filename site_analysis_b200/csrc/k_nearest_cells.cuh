// k_nearest_cells.cuh -- nearest-centre (Voronoi) assignment with exact spatial pruning.
//
// Same answers as k_nearest_assign<false> (k_nearest.cuh; reference voronoi_site_collection.py:76-99:
// argmin over sqrt'd 27-image minimum-image distances, first minimum wins), ~50x less arithmetic:
// a static cell list gives each atom the few centres that can be nearest to any point of its cell.
// The list never decides on its own: `excl[c]` is a lower bound (metric of the lattice the list was
// built with, L0) on the distance from any point of cell c to any centre NOT in the list.  In the
// frame's own metric that bound is (1 - eta) excl[c] with eta = ||L0^-1 L_t - I||_F.  The best listed
// centre is accepted only if its squared distance is below ((1 - eta) excl)^2 (1 - 1e-9) -- then no
// unlisted centre can win or tie, sqrt-merging included -- and below the central-image bound
// (0.5 hmin)^2 (k_nearest.cuh).  Otherwise the atom is redone against all centres.
#pragma once
#include "k_nearest.cuh"

#ifndef SAB_NEAREST_BATCH
#define SAB_NEAREST_BATCH 4
#endif
#ifndef SAB_NEAREST_MINBLOCKS
#define SAB_NEAREST_MINBLOCKS 3   // measured (profiles/r2_experiments/ab_nearest.txt): 8.92 -> 8.40 ms on config 3 (2 -> 3 CTAs per SM, 80 registers, cold-path spills)
#endif

struct SabNearestLists {
    int G;
    const int32_t *cell_off;    // [G^3 + 1]
    const uint16_t *cell_sites; // ascending site list positions
    const double *excl;         // [G^3]
    double l0inv[9];            // inverse of the lattice the lists were built with
};

__global__ void __launch_bounds__(256, SAB_NEAREST_MINBLOCKS) k_nearest_assign_cells(const double *__restrict__ frac, const double *__restrict__ lattice,
                                       int lattice_stride, int64_t n_frames, int n_atoms, int n_mobile,
                                       const int32_t *__restrict__ mobile, int n_sites,
                                       const double *__restrict__ centres, SabNearestLists NL,
                                       int32_t *__restrict__ result, unsigned long long *__restrict__ n_exhaustive)
{
    __shared__ SabLattice lat;
    __shared__ double shrink;                            // 1 - eta of this frame (<= 0 forces the exhaustive path)
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        __syncthreads();
        if (threadIdx.x == 0) {
            sab_lattice_load(lattice + (size_t)f * lattice_stride, lat);
            double eta2 = 0.0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double m = NL.l0inv[3 * i] * lat.m[j] + NL.l0inv[3 * i + 1] * lat.m[3 + j] + NL.l0inv[3 * i + 2] * lat.m[6 + j];
                    if (i == j) m -= 1.0;
                    eta2 += m * m;
                }
            shrink = 1.0 - sqrt(eta2) * (1.0 + 1e-9) - 1e-12;
        }
        __syncthreads();
        const double thr = (0.5 * lat.hmin) * (0.5 * lat.hmin) * (1.0 - SAB_REL_MARGIN);
        double L[9];                                     // the frame's lattice in registers
#pragma unroll
        for (int k = 0; k < 9; k++) L[k] = lat.m[k];
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
            double x[3];
#pragma unroll
            for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);   // atom.py:104
            int c0 = min(NL.G - 1, (int)(x[0] * (double)NL.G)), c1 = min(NL.G - 1, (int)(x[1] * (double)NL.G)),
                c2 = min(NL.G - 1, (int)(x[2] * (double)NL.G));
            const int cell = (c0 * NL.G + c1) * NL.G + c2;
            const int b = __ldg(NL.cell_off + cell), e = __ldg(NL.cell_off + cell + 1);
            double best_q = INFINITY, best_s = INFINITY;
            int best_i = -1;
            int i = b;
            for (; i + SAB_NEAREST_BATCH <= e; i += SAB_NEAREST_BATCH) {     // loads of a batch issue together, the argmin stays in list order
                int sb[SAB_NEAREST_BATCH];
                double cb[SAB_NEAREST_BATCH][3], qb[SAB_NEAREST_BATCH];
#pragma unroll
                for (int k = 0; k < SAB_NEAREST_BATCH; k++) sb[k] = __ldg(NL.cell_sites + i + k);
#pragma unroll
                for (int k = 0; k < SAB_NEAREST_BATCH; k++) {
                    cb[k][0] = __ldg(centres + 3 * sb[k]); cb[k][1] = __ldg(centres + 3 * sb[k] + 1); cb[k][2] = __ldg(centres + 3 * sb[k] + 2);
                }
#pragma unroll
                for (int k = 0; k < SAB_NEAREST_BATCH; k++) {
                    double d0, d1, d2;
                    sab_dbase(cb[k], x, d0, d1, d2);
                    qb[k] = sab_image_q(d0, d1, d2, L);
                }
#pragma unroll
                for (int k = 0; k < SAB_NEAREST_BATCH; k++)
                    if (qb[k] < best_q) {
                        double r = sqrt(qb[k]);
                        if (r < best_s) { best_s = r; best_q = qb[k]; best_i = sb[k]; }
                    }
            }
            for (; i < e; i++) {
                const int s = __ldg(NL.cell_sites + i);
                double cs[3] = {__ldg(centres + 3 * s), __ldg(centres + 3 * s + 1), __ldg(centres + 3 * s + 2)};
                double d0, d1, d2;
                sab_dbase(cs, x, d0, d1, d2);
                double q = sab_image_q(d0, d1, d2, L);
                if (q < best_q) {
                    double r = sqrt(q);
                    if (r < best_s) { best_s = r; best_q = q; best_i = s; }
                }
            }
            const double lim = shrink * __ldg(NL.excl + cell);
            const bool safe = best_i >= 0 && shrink > 0.0 && best_q < thr && best_q < lim * lim * (1.0 - SAB_REL_MARGIN);
            if (!safe) {                                 // rare: strained lattice, sparse sites, tiny cell
                atomicAdd(n_exhaustive, 1ULL);
                best_q = INFINITY; best_s = INFINITY; best_i = 0;
                for (int s = 0; s < n_sites; s++) {
                    double cs[3] = {__ldg(centres + 3 * s), __ldg(centres + 3 * s + 1), __ldg(centres + 3 * s + 2)};
                    double d0, d1, d2;
                    sab_dbase(cs, x, d0, d1, d2);
                    double q = sab_mic_q27(d0, d1, d2, lat.m);
                    if (q < best_q) {
                        double r = sqrt(q);
                        if (r < best_s) { best_s = r; best_q = q; best_i = s; }
                    }
                }
            }
            result[(size_t)f * n_mobile + a] = best_i;
        }
    }
}

// Dynamic Voronoi with the same pruning: centres are recomputed per frame into shared memory
// (sab_dyn_centre, k_nearest.cuh); the lists were built around ANCHOR centres with head-room `dmax`
// (lattice-L0 metric) for the centres' motion.  While every centre of the frame is within dmax of its
// anchor the exclusion bound holds for the moving centres too (see _tables.nearest_cell_lists:
// `move`); otherwise the whole frame takes the exhaustive path.
__global__ void __launch_bounds__(256, SAB_NEAREST_MINBLOCKS) k_dynvor_assign_cells(const double *__restrict__ frac, const double *__restrict__ lattice,
                                      int lattice_stride, int64_t n_frames, int n_atoms, int n_mobile,
                                      const int32_t *__restrict__ mobile, int n_sites,
                                      int smax, const int32_t *__restrict__ n_slot,
                                      const int32_t *__restrict__ slot_atom, const int32_t *__restrict__ slot_fw,
                                      const int32_t *__restrict__ slot_shift, const int32_t *__restrict__ wrows, int m3,
                                      const uint8_t *__restrict__ legacy_init_site, int64_t legacy_init_frame,
                                      SabNearestLists NL, const double *__restrict__ anchors, const double *__restrict__ l0,
                                      double dmax, int32_t *__restrict__ result,
                                      unsigned long long *__restrict__ n_exhaustive)
{
    extern __shared__ double sm_centres[];          // [S][3]
    __shared__ SabLattice lat;
    __shared__ double shrink;
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        const int32_t *wrow = wrows + (size_t)f * m3;
        __syncthreads();
        if (threadIdx.x == 0) {
            sab_lattice_load(lattice + (size_t)f * lattice_stride, lat);
            double eta2 = 0.0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double m = NL.l0inv[3 * i] * lat.m[j] + NL.l0inv[3 * i + 1] * lat.m[3 + j] + NL.l0inv[3 * i + 2] * lat.m[6 + j];
                    if (i == j) m -= 1.0;
                    eta2 += m * m;
                }
            shrink = 1.0 - sqrt(eta2) * (1.0 + 1e-9) - 1e-12;
        }
        int moved = 0;
        for (int s = threadIdx.x; s < n_sites; s += blockDim.x) {
            double c[3];
            bool li = legacy_init_site && legacy_init_site[s] && f == legacy_init_frame;
            sab_dyn_centre(frame, n_slot[s], slot_atom + (size_t)s * smax, slot_fw + (size_t)s * smax,
                           slot_shift + (size_t)s * smax * 3, wrow, li, c);
            sm_centres[3 * s] = c[0]; sm_centres[3 * s + 1] = c[1]; sm_centres[3 * s + 2] = c[2];
            double d0, d1, d2;                      // displacement from the anchor, lattice-L0 metric (pruning bound only)
            sab_dbase(c, anchors + 3 * s, d0, d1, d2);
            if (!(sab_image_q(d0, d1, d2, l0) <= dmax * dmax)) moved = 1;   // central image >= true distance: conservative
        }
        moved = __syncthreads_or(moved);
        if (moved && threadIdx.x == 0) atomicAdd(n_exhaustive, (unsigned long long)n_mobile);
        const double thr = (0.5 * lat.hmin) * (0.5 * lat.hmin) * (1.0 - SAB_REL_MARGIN);
        double L[9];
#pragma unroll
        for (int k = 0; k < 9; k++) L[k] = lat.m[k];
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
            double x[3];
#pragma unroll
            for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);
            double best_q = INFINITY, best_s = INFINITY;
            int best_i = -1;
            bool safe = false;
            if (!moved) {
                int c0 = min(NL.G - 1, (int)(x[0] * (double)NL.G)), c1 = min(NL.G - 1, (int)(x[1] * (double)NL.G)),
                    c2 = min(NL.G - 1, (int)(x[2] * (double)NL.G));
                const int cell = (c0 * NL.G + c1) * NL.G + c2;
                const int b = __ldg(NL.cell_off + cell), e = __ldg(NL.cell_off + cell + 1);
                int i = b;
                for (; i + SAB_NEAREST_BATCH <= e; i += SAB_NEAREST_BATCH) {
                    int sb[SAB_NEAREST_BATCH];
                    double qb[SAB_NEAREST_BATCH];
#pragma unroll
                    for (int k = 0; k < SAB_NEAREST_BATCH; k++) sb[k] = __ldg(NL.cell_sites + i + k);
#pragma unroll
                    for (int k = 0; k < SAB_NEAREST_BATCH; k++) {
                        double d0, d1, d2;
                        sab_dbase(sm_centres + 3 * sb[k], x, d0, d1, d2);
                        qb[k] = sab_image_q(d0, d1, d2, L);
                    }
#pragma unroll
                    for (int k = 0; k < SAB_NEAREST_BATCH; k++)
                        if (qb[k] < best_q) {
                            double r = sqrt(qb[k]);
                            if (r < best_s) { best_s = r; best_q = qb[k]; best_i = sb[k]; }
                        }
                }
                for (; i < e; i++) {
                    const int s = __ldg(NL.cell_sites + i);
                    double d0, d1, d2;
                    sab_dbase(sm_centres + 3 * s, x, d0, d1, d2);
                    double q = sab_image_q(d0, d1, d2, L);
                    if (q < best_q) {
                        double r = sqrt(q);
                        if (r < best_s) { best_s = r; best_q = q; best_i = s; }
                    }
                }
                const double lim = shrink * __ldg(NL.excl + cell);
                safe = best_i >= 0 && shrink > 0.0 && best_q < thr && best_q < lim * lim * (1.0 - SAB_REL_MARGIN);
                if (!safe) atomicAdd(n_exhaustive, 1ULL);
            }
            if (!safe) {
                best_q = INFINITY; best_s = INFINITY; best_i = 0;
                for (int s = 0; s < n_sites; s++) {
                    double d0, d1, d2;
                    sab_dbase(sm_centres + 3 * s, x, d0, d1, d2);
                    double q = sab_mic_q27(d0, d1, d2, lat.m);
                    if (q < best_q) {
                        double r = sqrt(q);
                        if (r < best_s) { best_s = r; best_q = q; best_i = s; }
                    }
                }
            }
            result[(size_t)f * n_mobile + a] = best_i;
        }
    }
}
