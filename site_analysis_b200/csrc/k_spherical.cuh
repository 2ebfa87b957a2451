// k_spherical.cuh -- spherical-cutoff containment for every (frame, atom).
//
// Replaces the containment tests of SphericalSiteCollection.assign_site_occupations
// (spherical_site_collection.py:71-92): SphericalSite.contains_point is
//     mic_distance(site.frac_coords, x, L) <= rcut          (spherical_site.py:145-146)
// i.e. sqrt(min over 27 images of |D.L|^2) <= rcut.  The host passes rcut_q[s] = the largest
// double q with sqrt(q) <= rcut[s], so the test is `q <= rcut_q[s]` with no sqrt on the device.
//
// The kernel emits the FULL set of containing sites per atom-frame: none -> SAB_NONE, one ->
// its list position (any priority order finds it, site_collection.py:113-182 yields every
// site), several -> a spilled candidate list for the priority resolver (k_resolve.cuh).
//
// Exact image pruning: if q(central image) <= rcut_q the site contains the atom (the minimum
// can only be smaller); otherwise, when rcut < 0.5 hmin (1 - 1e-9), every non-central image is
// farther than rcut and the site does not contain it.  Sites with larger radii take all 27.
#pragma once
#include "sab_device.cuh"

#define SAB_SPH_KEEP 8

__device__ __forceinline__ bool sab_sphere_contains(const double *c, const double *x, double rcut,
                                                    double rcut_q, const SabLattice &lat)
{
    double d0, d1, d2;
    sab_dbase(c, x, d0, d1, d2);
    if (sab_image_q(d0, d1, d2, lat.m) <= rcut_q) return true;
    if (rcut < 0.5 * lat.hmin * (1.0 - SAB_REL_MARGIN)) return false;
    return sab_mic_q27(d0, d1, d2, lat.m) <= rcut_q;
}

__global__ void k_spherical_assign(const double *__restrict__ frac, const double *__restrict__ lattice,
                                   int lattice_stride, int64_t n_frames, int n_atoms, int n_mobile,
                                   const int32_t *__restrict__ mobile, int n_sites,
                                   const double *__restrict__ centres, const double *__restrict__ rcut,
                                   const double *__restrict__ rcut_q, int32_t *__restrict__ result,
                                   int32_t *__restrict__ spill, long long spill_capacity,
                                   unsigned long long *__restrict__ counters /* [0] ambiguous [1] spill used [2] overflow */)
{
    extern __shared__ double sm[];                   // centres[S*3] | rcut[S] | rcut_q[S]
    double *sc = sm, *sr = sm + 3 * n_sites, *sq = sr + n_sites;
    __shared__ SabLattice lat;
    for (int i = threadIdx.x; i < 3 * n_sites; i += blockDim.x) sc[i] = centres[i];
    for (int i = threadIdx.x; i < n_sites; i += blockDim.x) { sr[i] = rcut[i]; sq[i] = rcut_q[i]; }
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        __syncthreads();
        if (threadIdx.x == 0) sab_lattice_load(lattice + (size_t)f * lattice_stride, lat);
        __syncthreads();
        // every thread takes part in every pass (valid = it has an atom): whole warps share one spill reservation below
        for (int a0 = 0; a0 < n_mobile; a0 += blockDim.x) {
            const int a = a0 + threadIdx.x;
            const bool valid = a < n_mobile;
            double x[3] = {0.0, 0.0, 0.0};
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);
            }
            int cnt = 0;
            int keep[SAB_SPH_KEEP];
            for (int s = 0; s < (valid ? n_sites : 0); s++)
                if (sab_sphere_contains(sc + 3 * s, x, sr[s], sq[s], lat)) {
                    if (cnt < SAB_SPH_KEEP) keep[cnt] = s;
                    cnt++;
                }
            // ONE reservation per warp: 98 % of the atom-frames of the overlapping-sphere benchmark spill, and two atomics
            // each on the same two counters serialise in L2 (ncu: 67 long-scoreboard stall cycles per issued instruction)
            const unsigned FULLM = 0xffffffffu;
            const int lane_w = threadIdx.x & 31;
            const int need = cnt > 1 ? SAB_SPILL_RECORD(cnt) : 0;
            int pre = need;                               // inclusive prefix sum over the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULLM, pre, o); if (lane_w >= o) pre += v; }
            const int warp_total = __shfl_sync(FULLM, pre, 31);
            const unsigned amb_mask = __ballot_sync(FULLM, cnt > 1);
            unsigned long long base = 0ULL;
            if (lane_w == 0 && warp_total > 0) {
                atomicAdd(&counters[0], (unsigned long long)__popc(amb_mask));
                base = atomicAdd(&counters[1], (unsigned long long)warp_total);
            }
            base = __shfl_sync(FULLM, base, 0);
            int out;
            if (cnt == 0) out = SAB_NONE;
            else if (cnt == 1) out = keep[0];
            else {
                const unsigned long long off = base + (unsigned long long)(pre - need);
                if ((long long)(off + cnt + 1) > spill_capacity || off + cnt + 1 >= 0x7ffffff0ULL) {
                    counters[2] = 1ULL;
                    out = SAB_NONE;
                } else {
                    spill[off] = cnt;
                    if (cnt <= SAB_SPH_KEEP) {
                        for (int i = 0; i < cnt; i++) spill[off + 1 + i] = keep[i];
                    } else {
                        int k = 0;
                        for (int s = 0; s < n_sites; s++)
                            if (sab_sphere_contains(sc + 3 * s, x, sr[s], sq[s], lat)) spill[off + 1 + k++] = s;
                    }
                    out = SAB_AMBIGUOUS_BASE - (int)off;
                }
            }
            if (valid) result[(size_t)f * n_mobile + a] = out;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The same containment test with exact spatial pruning.  A static G^3 cell list (host: _tables.sphere_cell_lists)
// gives every cell the sites whose sphere, inflated by the factor rho, can reach the cell in the metric of the
// lattice L0 the list was built with: a site NOT in the list of the atom's cell is farther than rho * rcut from the
// atom for every periodic image n in {-2..2}^3 -- a superset of the images the test above looks at.  In the
// frame's own metric every length is at least (1 - eta) times its L0 value, eta = ||L0^-1 L_t - I||_F, so while
// (1 - eta) rho >= 1 an unlisted site cannot contain the atom and only the listed ones are tested, with the very
// same arithmetic (sab_sphere_contains).  A frame whose lattice is strained beyond that takes the exhaustive
// loop.  Candidates are visited in ascending list position, so spilled sets equal those of k_spherical_assign.
struct SabSphereLists {
    int G;
    const int32_t *cell_off;    // [G^3 + 1]
    const uint16_t *cell_sites; // ascending site list positions
    double rho;                 // inflation the lists were built with
    double l0inv[9];
};

__global__ void k_spherical_assign_cells(const double *__restrict__ frac, const double *__restrict__ lattice,
                                         int lattice_stride, int64_t n_frames, int n_atoms, int n_mobile,
                                         const int32_t *__restrict__ mobile, int n_sites,
                                         const double *__restrict__ centres, const double *__restrict__ rcut,
                                         const double *__restrict__ rcut_q, SabSphereLists SL,
                                         int32_t *__restrict__ result, int32_t *__restrict__ spill, long long spill_capacity,
                                         unsigned long long *__restrict__ counters, unsigned long long *__restrict__ n_exhaustive)
{
    __shared__ SabLattice lat;
    __shared__ int lists_ok;
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        __syncthreads();
        if (threadIdx.x == 0) {
            sab_lattice_load(lattice + (size_t)f * lattice_stride, lat);
            double eta2 = 0.0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) {
                    double m = SL.l0inv[3 * i] * lat.m[j] + SL.l0inv[3 * i + 1] * lat.m[3 + j] + SL.l0inv[3 * i + 2] * lat.m[6 + j];
                    if (i == j) m -= 1.0;
                    eta2 += m * m;
                }
            const double shrink = 1.0 - sqrt(eta2) * (1.0 + 1e-9) - 1e-12;
            lists_ok = shrink * SL.rho >= 1.0 + 1e-9;
            if (!lists_ok) atomicAdd(n_exhaustive, 1ULL);
        }
        __syncthreads();
        const bool use_lists = lists_ok != 0;
        // every thread takes part in every pass (valid = it has an atom) so that whole warps can share one spill reservation
        for (int a0 = 0; a0 < n_mobile; a0 += blockDim.x) {
            const int a = a0 + threadIdx.x;
            const bool valid = a < n_mobile;
            double x[3] = {0.0, 0.0, 0.0};
            if (valid) {
#pragma unroll
                for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);
            }
            int b = 0, e = valid ? n_sites : 0;
            if (use_lists && valid) {
                int c0 = min(SL.G - 1, (int)(x[0] * (double)SL.G)), c1 = min(SL.G - 1, (int)(x[1] * (double)SL.G)),
                    c2 = min(SL.G - 1, (int)(x[2] * (double)SL.G));
                const int cell = (c0 * SL.G + c1) * SL.G + c2;
                b = __ldg(SL.cell_off + cell); e = __ldg(SL.cell_off + cell + 1);
            }
            int cnt = 0;
            int keep[SAB_SPH_KEEP];
            for (int i = b; i < e; i++) {
                const int s = use_lists ? (int)__ldg(SL.cell_sites + i) : i;
                const double cs[3] = {__ldg(centres + 3 * s), __ldg(centres + 3 * s + 1), __ldg(centres + 3 * s + 2)};
                if (sab_sphere_contains(cs, x, __ldg(rcut + s), __ldg(rcut_q + s), lat)) {
                    if (cnt < SAB_SPH_KEEP) keep[cnt] = s;
                    cnt++;
                }
            }
            // ONE reservation per warp: 98 % of the atom-frames of the overlapping-sphere benchmark spill, and two atomics
            // each on the same two counters serialise in L2 (ncu: 67 long-scoreboard stall cycles per issued instruction)
            const unsigned FULLM = 0xffffffffu;
            const int lane_w = threadIdx.x & 31;
            const int need = cnt > 1 ? SAB_SPILL_RECORD(cnt) : 0;
            int pre = need;                               // inclusive prefix sum over the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULLM, pre, o); if (lane_w >= o) pre += v; }
            const int warp_total = __shfl_sync(FULLM, pre, 31);
            const unsigned amb_mask = __ballot_sync(FULLM, cnt > 1);
            unsigned long long base = 0ULL;
            if (lane_w == 0 && warp_total > 0) {
                atomicAdd(&counters[0], (unsigned long long)__popc(amb_mask));
                base = atomicAdd(&counters[1], (unsigned long long)warp_total);
            }
            base = __shfl_sync(FULLM, base, 0);
            int out;
            if (cnt == 0) out = SAB_NONE;
            else if (cnt == 1) out = keep[0];
            else {
                const unsigned long long off = base + (unsigned long long)(pre - need);
                if ((long long)(off + cnt + 1) > spill_capacity || off + cnt + 1 >= 0x7ffffff0ULL) {
                    counters[2] = 1ULL;
                    out = SAB_NONE;
                } else {
                    spill[off] = cnt;
                    if (cnt <= SAB_SPH_KEEP) {
                        for (int i = 0; i < cnt; i++) spill[off + 1 + i] = keep[i];
                    } else {
                        int k = 0;
                        for (int i = b; i < e; i++) {
                            const int s = use_lists ? (int)__ldg(SL.cell_sites + i) : i;
                            const double cs[3] = {__ldg(centres + 3 * s), __ldg(centres + 3 * s + 1), __ldg(centres + 3 * s + 2)};
                            if (sab_sphere_contains(cs, x, __ldg(rcut + s), __ldg(rcut_q + s), lat)) spill[off + 1 + k++] = s;
                        }
                    }
                    out = SAB_AMBIGUOUS_BASE - (int)off;
                }
            }
            if (valid) result[(size_t)f * n_mobile + a] = out;
        }
    }
}
