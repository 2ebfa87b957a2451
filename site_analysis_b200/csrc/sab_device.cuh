// sab_device.cuh -- device-side arithmetic shared by every kernel.
//
// Bit-exactness contract: float64 only, NO fused multiply-add (the whole
// library is compiled with -fmad=false), round-half-even rint, IEEE sqrt/div, and the
// exact operation ORDER of the reference's numba kernels.  Every helper cites the
// reference lines whose rounding sequence it reproduces.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define SAB_NONE (-1)
#define SAB_AMBIGUOUS_BASE (-2)
// A candidate list in the spill buffer is {n, c_0 .. c_{n-1}} (ascending); records are allocated in multiples of 4 ints so that
// every list starts on a 16-byte boundary of the (16-byte aligned) buffer: the resolvers fetch them with 128-bit loads.
#define SAB_SPILL_RECORD(n) ((((n) + 1) + 3) & ~3)
#define SAB_BBOX_EPS 1e-9      // fractional slack of the exact-pruning bounding boxes
#define SAB_REL_MARGIN 1e-9    // relative slack of every pruning bound on a distance

// numpy `x % 1.0` (Python-mod semantics), reference atom.py:104 and
// dynamic_voronoi_site_collection.py:104.  fmod(x, 1) == x - trunc(x) exactly.
__device__ __forceinline__ double sab_pymod1(double x)
{
    double r = x - trunc(x);
    if (r != 0.0) { if (r < 0.0) r += 1.0; }
    else r = 0.0;
    return r;
}

struct SabLattice {
    double m[9];      // rows are lattice vectors
    double hmin;      // smallest perpendicular cell height (pruning bound only)
};

// Perpendicular heights h_i = V / |a_j x a_k|: |D . L| >= |D_i| h_i for any fractional D.
// Used only in conservative pruning bounds (with SAB_REL_MARGIN), never in a decision value.
__device__ __forceinline__ void sab_lattice_load(const double *L, SabLattice &lat)
{
#pragma unroll
    for (int i = 0; i < 9; i++) lat.m[i] = L[i];
    const double *a = lat.m, *b = lat.m + 3, *c = lat.m + 6;
    double bcx = b[1] * c[2] - b[2] * c[1], bcy = b[2] * c[0] - b[0] * c[2], bcz = b[0] * c[1] - b[1] * c[0];
    double cax = c[1] * a[2] - c[2] * a[1], cay = c[2] * a[0] - c[0] * a[2], caz = c[0] * a[1] - c[1] * a[0];
    double abx = a[1] * b[2] - a[2] * b[1], aby = a[2] * b[0] - a[0] * b[2], abz = a[0] * b[1] - a[1] * b[0];
    double vol = fabs(a[0] * bcx + a[1] * bcy + a[2] * bcz);
    double h0 = vol / sqrt(bcx * bcx + bcy * bcy + bcz * bcz);
    double h1 = vol / sqrt(cax * cax + cay * cay + caz * caz);
    double h2 = vol / sqrt(abx * abx + aby * aby + abz * abz);
    lat.hmin = fmin(h0, fmin(h1, h2));
}

// Reduced fractional separation d = (a - b) - rint(a - b): distances.py:90-95 (a = site, b = atom).
__device__ __forceinline__ void sab_dbase(const double *a, const double *b, double &d0, double &d1, double &d2)
{
    d0 = a[0] - b[0]; d1 = a[1] - b[1]; d2 = a[2] - b[2];
    d0 -= rint(d0); d1 -= rint(d1); d2 -= rint(d2);
}

// Squared Cartesian length of one image, distances.py:101-105 (left-to-right adds, no FMA).
__device__ __forceinline__ double sab_image_q(double d0, double d1, double d2, const double *L)
{
    double cx = d0 * L[0] + d1 * L[3] + d2 * L[6];
    double cy = d0 * L[1] + d1 * L[4] + d2 * L[7];
    double cz = d0 * L[2] + d1 * L[5] + d2 * L[8];
    return cx * cx + cy * cy + cz * cz;
}

// min over the 27 images, distances.py:96-108.
__device__ __forceinline__ double sab_mic_q27(double d0b, double d1b, double d2b, const double *L)
{
    double best = INFINITY;
#pragma unroll 1
    for (int si = -1; si < 2; si++)
        for (int sj = -1; sj < 2; sj++)
            for (int sk = -1; sk < 2; sk++) {
                double q = sab_image_q(d0b + (double)si, d1b + (double)sj, d2b + (double)sk, L);
                if (q < best) best = q;
            }
    return best;
}

__device__ __forceinline__ double sab_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }

// One face of a polyhedron against one point: polyhedral_site.py:53-79 (normal, reference
// vertex, centroid sign) and :107-121 (query).  Returns true when the face does NOT reject.
__device__ __forceinline__ bool sab_face_accepts(const double *v0, const double *v1, const double *v2,
                                                 const double *c, double p0, double p1, double p2)
{
    double e0x = v0[0] - v2[0], e0y = v0[1] - v2[1], e0z = v0[2] - v2[2];
    double e1x = v1[0] - v2[0], e1y = v1[1] - v2[1], e1z = v1[2] - v2[2];
    double n0 = e0y * e1z - e0z * e1y;
    double n1 = e0z * e1x - e0x * e1z;
    double n2 = e0x * e1y - e0y * e1x;
    double cd = n0 * (c[0] - v0[0]);          // 0.0 + t0 == t0
    cd += n1 * (c[1] - v0[1]);
    cd += n2 * (c[2] - v0[2]);
    double dot = (p0 - v0[0]) * n0;
    dot += (p1 - v0[1]) * n1;
    dot += (p2 - v0[2]) * n2;
    // reject iff dot != 0 and sign(dot) != sign(cd) (polyhedral_site.py:119-121); sign(cd) == 0 never
    // matches a non-zero dot.  Sign bits compared as integers (NaN inputs are outside the contract).
    const bool differ = (__double2hiint(dot) ^ __double2hiint(cd)) < 0;
    return !(dot != 0.0 && (cd == 0.0 || differ));
}
