// k_nearest.cuh -- nearest-centre (Voronoi) assignment, with the dynamic-Voronoi centre
// recomputation fused in front of it.
//
// Replaces VoronoiSiteCollection.assign_site_occupations (voronoi_site_collection.py:76-99)
// and DynamicVoronoiSiteCollection._batch_calculate_centres + assign_site_occupations
// (dynamic_voronoi_site_collection.py:166-195, 222-244):
//     dist = all_mic_distances(site_centres, atom_coords, L)   (distances.py:68-110)
//     site = argmin(dist, axis=0)                               (first minimum of the SQRT'd values)
// The S x A distance matrix is never materialised.
//
// v1 layout: one CTA per frame (grid-stride), centres staged in shared memory once per
// frame (dynamic) or once per CTA (static), one thread per mobile atom walking the sites
// in list order so that "first minimum wins" is the natural tie rule.
//
// Exact pruning of the 27 periodic images: with d = (c - x) - rint(c - x) every non-central
// image has |D_i| >= 0.5 for some i, hence length >= 0.5 * h_i >= 0.5 * hmin.  If the best
// central-image value is below (0.5 hmin)^2 (1 - 1e-9), no non-central image of any site can
// reach or tie the minimum, and central-image values alone reproduce the reference argmin
// bit for bit (the winner's value IS its 27-image minimum).  Otherwise the atom is redone
// with all 27 images.
#pragma once
#include "sab_device.cuh"

// Centre of one dynamic-Voronoi site: dynamic_voronoi_site_collection.py:97-104 (fast path)
// == :191-193 with correct_pbc when the site has a reference centre (pbc_utils.py:102-104).
// `legacy_init` reproduces the first-frame legacy path (no reference centre), which applies
// no non-negative offset (pbc_utils.py:16-41).
#define SAB_DYN_RMAX 6     // reference atoms per site handled with every operand loaded once into registers

__device__ __noinline__ void sab_dyn_centre(const double *__restrict__ frame, int n_ref,
                                               const int32_t *__restrict__ slot_atom,
                                               const int32_t *__restrict__ slot_fw,
                                               const int32_t *__restrict__ slot_shift,
                                               const int32_t *__restrict__ wrow, bool legacy_init,
                                               double *out)
{
    if (n_ref <= SAB_DYN_RMAX) {
        // same operations in the same order as the general loop below; each unwrapped coordinate
        // raw + (shift - W) (pbc_utils.py:220) is formed once and reused by the min and the sum
        double v[SAB_DYN_RMAX][3];
#pragma unroll
        for (int r = 0; r < SAB_DYN_RMAX; r++)
            if (r < n_ref) {
                const double *x = frame + (size_t)__ldg(slot_atom + r) * 3;
                const int32_t *w = wrow + 3 * __ldg(slot_fw + r);
#pragma unroll
                for (int d = 0; d < 3; d++) v[r][d] = x[d] + (double)(__ldg(slot_shift + 3 * r + d) - w[d]);
            }
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double mn = INFINITY;
#pragma unroll
            for (int r = 0; r < SAB_DYN_RMAX; r++) if (r < n_ref) mn = fmin(mn, v[r][d]);
            double nn = ceil(-mn);
            if (!(nn > 0.0) || legacy_init) nn = 0.0;
            double sum = 0.0;
#pragma unroll
            for (int r = 0; r < SAB_DYN_RMAX; r++)
                if (r < n_ref) { const double t = v[r][d] + nn; sum = (r == 0) ? t : sum + t; }
            out[d] = sab_pymod1(sum / (double)n_ref);
        }
        return;
    }
#pragma unroll 1
    for (int d = 0; d < 3; d++) {
        double mn = INFINITY;
        for (int r = 0; r < n_ref; r++) {
            double raw = frame[(size_t)slot_atom[r] * 3 + d];
            double v = raw + (double)(slot_shift[3 * r + d] - wrow[3 * slot_fw[r] + d]);
            mn = fmin(mn, v);
        }
        double nn = ceil(-mn);
        if (!(nn > 0.0) || legacy_init) nn = 0.0;
        double sum = 0.0;
        for (int r = 0; r < n_ref; r++) {
            double raw = frame[(size_t)slot_atom[r] * 3 + d];
            double v = (raw + (double)(slot_shift[3 * r + d] - wrow[3 * slot_fw[r] + d])) + nn;
            sum = (r == 0) ? v : sum + v;
        }
        out[d] = sab_pymod1(sum / (double)n_ref);
    }
}

template <bool DYN>
__global__ void k_nearest_assign(const double *__restrict__ frac, const double *__restrict__ lattice,
                                 int lattice_stride, int64_t n_frames, int n_atoms, int n_mobile,
                                 const int32_t *__restrict__ mobile, int n_sites,
                                 const double *__restrict__ static_centres,
                                 // dynamic only
                                 int smax, const int32_t *__restrict__ n_slot,
                                 const int32_t *__restrict__ slot_atom, const int32_t *__restrict__ slot_fw,
                                 const int32_t *__restrict__ slot_shift, const int32_t *__restrict__ wrows, int m3,
                                 const uint8_t *__restrict__ legacy_init_site, int64_t legacy_init_frame,
                                 double *__restrict__ centres_out,
                                 int32_t *__restrict__ result, unsigned long long *__restrict__ n_full27)
{
    extern __shared__ double sm_centres[];          // [S][3]
    __shared__ SabLattice lat;
    if (!DYN) {
        for (int i = threadIdx.x; i < 3 * n_sites; i += blockDim.x) sm_centres[i] = static_centres[i];
    }
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        __syncthreads();
        if (threadIdx.x == 0) sab_lattice_load(lattice + (size_t)f * lattice_stride, lat);
        if (DYN) {
            const int32_t *wrow = wrows + (size_t)f * m3;
            for (int s = threadIdx.x; s < n_sites; s += blockDim.x) {
                double c[3];
                bool li = legacy_init_site && legacy_init_site[s] && f == legacy_init_frame;
                sab_dyn_centre(frame, n_slot[s], slot_atom + (size_t)s * smax, slot_fw + (size_t)s * smax,
                               slot_shift + (size_t)s * smax * 3, wrow, li, c);
                sm_centres[3 * s] = c[0]; sm_centres[3 * s + 1] = c[1]; sm_centres[3 * s + 2] = c[2];
                if (centres_out) {
                    double *o = centres_out + ((size_t)f * n_sites + s) * 3;
                    o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
                }
            }
        }
        __syncthreads();
        if (result == nullptr) continue;
        const double thr = (0.5 * lat.hmin) * (0.5 * lat.hmin) * (1.0 - SAB_REL_MARGIN);
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
            double x[3];
#pragma unroll
            for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);   // atom.py:104
            double best_q = INFINITY, best_s = INFINITY;
            int best_i = 0;
            for (int s = 0; s < n_sites; s++) {
                double d0, d1, d2;
                sab_dbase(sm_centres + 3 * s, x, d0, d1, d2);
                double q = sab_image_q(d0, d1, d2, lat.m);
                if (q < best_q) {                   // sqrt is monotone: q >= best_q cannot win
                    double r = sqrt(q);
                    if (r < best_s) { best_s = r; best_q = q; best_i = s; }
                }
            }
            if (!(best_q < thr)) {                  // rare: small cell / sparse sites -> all 27 images
                atomicAdd(n_full27, 1ULL);
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
