// k_polyhedral_cells.cuh -- polyhedral (tetrahedral) containment with exact spatial pruning.
//
// Same outputs as k_polyhedral_assign_exhaustive (k_polyhedral.cuh), same arithmetic per
// accepted test (reference polyhedral_site.py:26-123, pbc_utils.py:218-229), ~100x less work:
//
//  * ATOM-major: one thread per mobile atom, one CTA per frame (grid-stride, several CTAs
//    resident per SM).
//  * Per frame the CTA stages, for every referenced framework atom, the nk possible unwrapped
//    coordinates  raw + (double)(k - W)  (k = slot-shift value, W = cumulative wrap count of the
//    atom in this frame) into shared memory: one rounding, exactly the reference's
//    `frac_coords + new_shifts` (pbc_utils.py:220).  A vertex coordinate is then ONE LDS.
//  * A static cell list over the unit cube (built on the host from anchor positions and a
//    margin delta that covers the framework's thermal motion) gives each atom the few
//    tetrahedra whose inflated hull meets its cell.  While staging, the CTA checks that every
//    framework atom is within delta of its anchor; a frame that violates this takes the
//    exhaustive path, so pruning never decides anything by itself.
//  * "Small" sites (extent + 2 delta < 0.45 per axis, checked on the host): only the image
//    o = rint(centroid - x) of the 8 images x + {0,1}^3 (tools.py:246-253) can lie inside the
//    vertices' bounding box, so it is the only one tested (others are outside the hull by
//    > 0.05 fractional units).
//  * Face data are computed lazily, face by face, with early exit (polyhedral_site.py:109-123);
//    each site record holds the shared-memory offsets of every face's vertices in Qhull's own
//    vertex order, so normals and reference points round exactly as in the reference.
#pragma once
#include "k_polyhedral.cuh"

#define SAB_TET_REC 20          // uint16 per site record (40 bytes)

struct SabTetTables {
    int M, m3;                  // referenced framework atoms, 3 M
    int kmin, nk;               // slot-shift values kmin .. kmin + nk - 1
    const int32_t *fw_atoms;    // [M]
    // [S][20] uint16: [0..11] byte offsets into `var` of the 4 vertices in SLOT order (x y z each);
    // [12..17] = 12 bytes: local vertex number of (face j, corner k) in Qhull's order; [18..19] pad.
    // 40 bytes per site keep the whole table L1-resident (1056 sites = 42 KB).
    const uint16_t *rec;
    int G;                      // cells per axis
    const int32_t *cell_off;    // [G^3 + 1]
    const uint16_t *cell_sites; // CSR entries (site list positions, ascending)
    const double *anchor;       // [3 M] anchor of raw - W per framework coordinate
    double delta;               // validity radius of the cell lists
    int use_f32;                // SAB_OPT_FP32_FILTER: run the FP32 pre-test in front of the exact test
};

// record offsets are BYTE offsets into `var` (pre-multiplied by 8; var is < 64 KB)
__device__ __forceinline__ double sab_ldvar(const double *var, unsigned byte_off)
{
    return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(var) + byte_off);
}

// One tetrahedron against one wrapped atom position.  `var` = staged unwrapped coordinates.
__device__ __forceinline__ bool sab_tet_contains(const double *__restrict__ var, const uint16_t *__restrict__ rec,
                                                 const double *x)
{
    double v[4][3];
    {
        const uint2 *r2 = reinterpret_cast<const uint2 *>(rec);      // 24 bytes = 3 x 8
        uint2 a = __ldg(r2), b = __ldg(r2 + 1), c2 = __ldg(r2 + 2);
        v[0][0] = sab_ldvar(var, a.x & 0xffff); v[0][1] = sab_ldvar(var, a.x >> 16); v[0][2] = sab_ldvar(var, a.y & 0xffff);
        v[1][0] = sab_ldvar(var, a.y >> 16);    v[1][1] = sab_ldvar(var, b.x & 0xffff); v[1][2] = sab_ldvar(var, b.x >> 16);
        v[2][0] = sab_ldvar(var, b.y & 0xffff); v[2][1] = sab_ldvar(var, b.y >> 16);    v[2][2] = sab_ldvar(var, c2.x & 0xffff);
        v[3][0] = sab_ldvar(var, c2.x >> 16);   v[3][1] = sab_ldvar(var, c2.y & 0xffff); v[3][2] = sab_ldvar(var, c2.y >> 16);
    }
    double u[3], c[3], p[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        // non-negative offset: u = ceil(-min) when min < 0 (pbc_utils.py:221-229)
        u[d] = 0.0;
        if ((__double2hiint(v[0][d]) | __double2hiint(v[1][d]) | __double2hiint(v[2][d]) | __double2hiint(v[3][d])) < 0) {
            double mn = fmin(fmin(v[0][d], v[1][d]), fmin(v[2][d], v[3][d]));
            if (mn < 0.0) u[d] = ceil(-mn);
        }
        // centroid, vertices in slot order (polyhedral_site.py:46-51): ((((0 + v0) + v1) + v2) + v3) / 4
        double s = (v[0][d] + u[d]) + (v[1][d] + u[d]);
        s += v[2][d] + u[d];
        s += v[3][d] + u[d];
        c[d] = s * 0.25;                                        // == s / 4 exactly
        double o_d = rint(c[d] - x[d]);
        if (o_d != 0.0 && o_d != 1.0) return false;             // no image of x within reach of this site
        p[d] = (o_d == 1.0) ? 1.0 + x[d] : x[d];
    }
    const unsigned fv0 = __ldg(reinterpret_cast<const unsigned *>(rec + 12));      // faces 0, 1(first corner)
    const unsigned fv1 = __ldg(reinterpret_cast<const unsigned *>(rec + 14));
    const unsigned fv2 = __ldg(reinterpret_cast<const unsigned *>(rec + 16));
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        // corners of face j: bytes 3j, 3j+1, 3j+2 of the 12-byte string fv0|fv1|fv2
        unsigned long long lo = ((unsigned long long)fv1 << 32) | fv0;
        unsigned c3;
        if (j < 2) c3 = (unsigned)(lo >> (24 * j));
        else c3 = (unsigned)((((unsigned long long)fv2 << 32) | fv1) >> (24 * j - 32));
        const int i0 = c3 & 0xff, i1 = (c3 >> 8) & 0xff, i2 = (c3 >> 16) & 0xff;
        double v0[3], v1[3], v2[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            v0[d] = sab_ldvar(var, __ldg(rec + 3 * i0 + d)) + u[d];
            v1[d] = sab_ldvar(var, __ldg(rec + 3 * i1 + d)) + u[d];
            v2[d] = sab_ldvar(var, __ldg(rec + 3 * i2 + d)) + u[d];
        }
        if (!sab_face_accepts(v0, v1, v2, c, p[0], p[1], p[2])) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// FP32 pre-test (the "fp32 fast path re-checked in fp64 within a stated epsilon band").
//
// Geometry only: is the image x + k of the atom decisively inside / outside the tetrahedron?
// It uses the four faces of the simplex in a FIXED vertex order (the verdict does not depend on
// Qhull's ordering) and single precision with FMA.  Every float coordinate is within
// delta = 2^-22 of its float64 value (|coordinate| < 4 is checked while staging; the centroid within
// 2 delta), so with E >= every coordinate difference that enters the test (E < 1 enforced), the
// computed plane values are within ~60 E^2 delta of their exact values.  A face
// is DECISIVE only when both |n.(p - v)| and |n.(centroid - v)| exceed T = 256 E^2 delta = 2^-14 E^2
// -- then the float64
// reference arithmetic (relative error ~1e-16) has the same signs.  Anything else -- a point within
// ~1e-4 of a face, a vertex within 1e-5 of a cell face (integer offset u ambiguous), an image choice
// within 0.05 of a half-integer -- returns MAYBE and the caller runs the exact float64 test.
// Return: 0 = certainly not contained, 1 = certainly contained, 2 = undecided.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sab_tet_filter32(const float *__restrict__ varf, const uint16_t *__restrict__ rec,
                                                const double *x)
{
    float v[4][3];
    {
        const uint2 *r2 = reinterpret_cast<const uint2 *>(rec);
        uint2 a = __ldg(r2), b = __ldg(r2 + 1), c2 = __ldg(r2 + 2);
        // record offsets are byte offsets into the float64 table; the float table has the same indexing
        #define SAB_LDF(o) (*reinterpret_cast<const float *>(reinterpret_cast<const char *>(varf) + ((o) >> 1)))
        v[0][0] = SAB_LDF(a.x & 0xffff); v[0][1] = SAB_LDF(a.x >> 16); v[0][2] = SAB_LDF(a.y & 0xffff);
        v[1][0] = SAB_LDF(a.y >> 16);    v[1][1] = SAB_LDF(b.x & 0xffff); v[1][2] = SAB_LDF(b.x >> 16);
        v[2][0] = SAB_LDF(b.y & 0xffff); v[2][1] = SAB_LDF(b.y >> 16);    v[2][2] = SAB_LDF(c2.x & 0xffff);
        v[3][0] = SAB_LDF(c2.x >> 16);   v[3][1] = SAB_LDF(c2.y & 0xffff); v[3][2] = SAB_LDF(c2.y >> 16);
        #undef SAB_LDF
    }
    float p[3], c[3], E = 0.f;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float mn = fminf(fminf(v[0][d], v[1][d]), fminf(v[2][d], v[3][d]));
        const float mx = fmaxf(fmaxf(v[0][d], v[1][d]), fmaxf(v[2][d], v[3][d]));
        if (fabsf(mn - rintf(mn)) < 1e-5f) return 2;           // non-negative offset u = ceil(-min) ambiguous in float
        const float u = (mn < 0.f) ? ceilf(-mn) : 0.f;
        c[d] = 0.25f * ((v[0][d] + v[1][d]) + (v[2][d] + v[3][d]));
        const float xf = (float)x[d];
        const float t = c[d] - xf, k = rintf(t);
        if (fabsf(t - k) > 0.45f) return 2;                    // image choice not clear-cut (or site not small)
        const float o = k + u;                                 // == rint(centroid64 + u - x): the only image in reach
        if (o != 0.f && o != 1.f) return 0;                    // not one of x + {0,1}^3 (tools.py:246-253)
        p[d] = xf + k;
        E = fmaxf(E, fmaxf(mx - mn, fmaxf(fabsf(p[d] - mn), fabsf(p[d] - mx))));
    }
    E += 1e-6f;
    if (!(E < 1.0f)) return 2;                                 // error budget below assumes differences < 1
    const float T = E * E * 6.103515625e-5f;                   // 2^-14 E^2 = 256 E^2 delta
    bool all_in = true;
#pragma unroll
    for (int j = 0; j < 4; j++) {                              // face j = the three vertices other than j
        const int i0 = (j == 0) ? 1 : 0, i1 = (j <= 1) ? 2 : 1, i2 = (j == 3) ? 2 : 3;
        const float e0x = v[i0][0] - v[i2][0], e0y = v[i0][1] - v[i2][1], e0z = v[i0][2] - v[i2][2];
        const float e1x = v[i1][0] - v[i2][0], e1y = v[i1][1] - v[i2][1], e1z = v[i1][2] - v[i2][2];
        // explicit FMAs (the library is built with -fmad=false for the float64 reference arithmetic)
        const float n0 = __fmaf_rn(e0y, e1z, -(e0z * e1y));
        const float n1 = __fmaf_rn(e0z, e1x, -(e0x * e1z));
        const float n2 = __fmaf_rn(e0x, e1y, -(e0y * e1x));
        const float cd = __fmaf_rn(n2, c[2] - v[i0][2], __fmaf_rn(n1, c[1] - v[i0][1], n0 * (c[0] - v[i0][0])));
        const float dt = __fmaf_rn(n2, p[2] - v[i0][2], __fmaf_rn(n1, p[1] - v[i0][1], n0 * (p[0] - v[i0][0])));
        if (!(fabsf(cd) > T)) return 2;                        // degenerate / needle tetrahedron: exact test decides
        if (!(fabsf(dt) > T)) { all_in = false; continue; }    // within the band of this face: undecided so far
        if ((dt > 0.f) != (cd > 0.f)) return 0;                // decisively on the outer side of this face
    }
    return all_in ? 1 : 2;
}

// exact test with the float pre-test in front
#ifdef SAB_COUNT_F32
__device__ unsigned long long sab_f32_stats[3];                // verdict histogram (debug builds only)
#endif
__device__ __forceinline__ bool sab_tet_contains_f(const double *__restrict__ var, const float *__restrict__ varf,
                                                   const uint16_t *__restrict__ rec, const double *x, int use_f32)
{
    const int f = use_f32 ? sab_tet_filter32(varf, rec, x) : 2;
#ifdef SAB_COUNT_F32
    atomicAdd(&sab_f32_stats[f], 1ULL);
#endif
    if (f != 2) return f == 1;
    return sab_tet_contains(var, rec, x);
}

__global__ void __launch_bounds__(256)
k_polyhedral_assign_cells(const double *__restrict__ frac, int64_t n_frames, int n_atoms, int n_mobile,
                          const int32_t *__restrict__ mobile, int n_sites, SabPolyTables T, SabTetTables C,
                          const int32_t *__restrict__ wrows, int64_t legacy_init_frame,
                          int32_t *__restrict__ result, int32_t *__restrict__ spill, long long spill_capacity,
                          unsigned long long *__restrict__ counters, unsigned long long *__restrict__ n_fallback)
{
    extern __shared__ double var[];                  // [nk][3 M]
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        const int32_t *wrow = wrows + (size_t)f * C.m3;
        __syncthreads();                             // previous frame's readers are done
        int bad = 0;
        for (int i = threadIdx.x; i < C.m3; i += blockDim.x) {
            int m = i / 3, d = i - 3 * m;
            double raw = frame[(size_t)C.fw_atoms[m] * 3 + d];
            int W = wrow[i];
            if (fabs((raw - (double)W) - C.anchor[i]) > C.delta) bad = 1;
            for (int k = 0; k < C.nk; k++) var[k * C.m3 + i] = raw + (double)(C.kmin + k - W);
        }
        bad = __syncthreads_or(bad);
        if (T.legacy_init && f == legacy_init_frame) bad = 1;     // first-frame legacy rule: generic path
        if (threadIdx.x == 0 && bad) atomicAdd(n_fallback, 1ULL);
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
            double x[3];
#pragma unroll
            for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);
            int cnt = 0, keep[SAB_POLY_KEEP];
            if (!bad) {
                int ci[3];
#pragma unroll
                for (int d = 0; d < 3; d++) ci[d] = min(C.G - 1, (int)(x[d] * (double)C.G));
                int cell = (ci[0] * C.G + ci[1]) * C.G + ci[2];
                int b = __ldg(C.cell_off + cell), e = __ldg(C.cell_off + cell + 1);
                for (int i = b; i < e; i++) {
                    int s = __ldg(C.cell_sites + i);
                    if (sab_tet_contains(var, C.rec + (size_t)s * SAB_TET_REC, x)) {
                        if (cnt < SAB_POLY_KEEP) keep[cnt] = s;
                        cnt++;
                    }
                }
            } else {                                    // exhaustive, fully general (rare)
                for (int s = 0; s < n_sites; s++) {
                    double vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
                    int nv;
                    bool legacy_now = T.legacy_init && T.legacy_init[s] && f == legacy_init_frame;
                    sab_poly_vertices(frame, T, s, wrow, legacy_now, vc, nv, lo, hi, cen);
                    if (sab_poly_contains_atom(vc, cen, T.simp + (size_t)s * T.fmax * 3, T.n_face[s], lo, hi, x)) {
                        if (cnt < SAB_POLY_KEEP) keep[cnt] = s;
                        cnt++;
                    }
                }
            }
            int out;
            if (cnt == 0) out = SAB_NONE;
            else if (cnt == 1) out = keep[0];
            else {
                atomicAdd(&counters[0], 1ULL);
                unsigned long long off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(cnt));
                if ((long long)(off + cnt + 1) > spill_capacity || off + cnt + 1 >= 0x7ffffff0ULL) {
                    counters[2] = 1ULL;
                    out = SAB_NONE;
                } else {
                    spill[off] = cnt;
                    if (cnt <= SAB_POLY_KEEP) {
                        for (int i = 0; i < cnt; i++) spill[off + 1 + i] = keep[i];   // ascending already
                    } else {                               // atom on a shared vertex / edge: redo, write all
                        int k = 0;
                        if (!bad) {
                            int ci[3];
                            for (int d = 0; d < 3; d++) ci[d] = min(C.G - 1, (int)(x[d] * (double)C.G));
                            int cell = (ci[0] * C.G + ci[1]) * C.G + ci[2];
                            for (int i = C.cell_off[cell]; i < C.cell_off[cell + 1]; i++) {
                                int s = C.cell_sites[i];
                                if (sab_tet_contains(var, C.rec + (size_t)s * SAB_TET_REC, x)) spill[off + 1 + k++] = s;
                            }
                        } else {
                            for (int s = 0; s < n_sites; s++) {
                                double vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
                                int nv;
                                bool legacy_now = T.legacy_init && T.legacy_init[s] && f == legacy_init_frame;
                                sab_poly_vertices(frame, T, s, wrow, legacy_now, vc, nv, lo, hi, cen);
                                if (sab_poly_contains_atom(vc, cen, T.simp + (size_t)s * T.fmax * 3, T.n_face[s], lo, hi, x))
                                    spill[off + 1 + k++] = s;
                            }
                        }
                    }
                    out = SAB_AMBIGUOUS_BASE - (int)off;
                }
            }
            result[(size_t)f * n_mobile + a] = out;
        }
    }
}
