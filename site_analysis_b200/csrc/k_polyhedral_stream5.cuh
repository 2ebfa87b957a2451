// k_polyhedral_stream5.cuh -- the production tetrahedral kernel, generation 5: the warp-specialised frame
// pipeline of k_polyhedral_stream.cuh with less than half of its instructions and shared-memory traffic.
//
// Same contract and bit-identical results as k_polyhedral_assign_stream (polyhedral_site_collection.py:86-107 with
// the "most recent site first" order of site_collection.py:147-152, vertices from pbc_utils.py:193-230, containment
// polyhedral_site.py:83-123).  What changed, and why (profiles/r1_ncu_full_stream_v4_125kframes.csv: 4,530 warp
// instructions per frame, issue bound; profiles/r2_ncu_stream5_*.csv for the steps in between):
//
//   * A STAGED ROW IS ONE float4 PER ATOM: the wrapped float32 position of every framework atom some site uses, then
//     of every mobile atom.  No per-site image shifts are staged: every vertex difference the pre-test needs is a
//     reduced separation d - rint(d), which equals the reference's unwrapped difference because sites are smaller
//     than 0.45 per axis (see sab_tet5).  A site record is four 16-bit byte offsets; a containment test loads four
//     vertices and the atom with five 128-bit shared loads (was 15 scalar loads + 12 unpack operations).
//   * THE RECORDS OF THE TWO MOST RECENT SITES LIVE IN REGISTERS and are reloaded only when the atom changes site.
//   * CONSTANT ERROR THRESHOLDS.  The error budget of the FP32 pre-test is evaluated once on the host for the largest
//     site of the table; only a test that lands inside that (wider) band recomputes the per-site budget.  Site
//     records are stored positively oriented (host), so the sign flip of the verdict is gone too.
//   * PRODUCERS WORK IN float32: one float64 -> float32 conversion per coordinate, then W = rint(x - anchor) as two
//     additions of 1.5 * 2^23, the validity-tube test and the wrapped value.  Metadata (column, three float anchors)
//     sits in registers, every load of a frame is issued before its first use.  They no longer track the float64
//     excursion of every coordinate: a frame inside the validity tube |u - anchor| <= delta <= 0.12 cannot widen the
//     excursion beyond [anchor - delta, anchor + delta], which is what a chunk reports (guard G1 needs < 0.3;
//     DESIGN.md section 4); frames outside the tube update it explicitly.
//   * One mbarrier arrival per WARP (after __syncwarp), ring positions advanced without integer division.
//   * NO COLD CODE IN THE FRAME LOOP.  Frames outside the validity tube (and the legacy first-frame rule) are only
//     listed; k_polyhedral_flagged settles them with the generic all-sites test after the kernel.  Atom-frames that
//     several sites contain join the ones float32 cannot decide in the exact float64 pass (k_polyhedral_deferred),
//     which writes the candidate list for the priority resolver.  Either way the atom's history restarts as
//     "unknown", which only ever selects the slower path to the same answer.
//   * THE SEARCH READS ONE 64-BYTE CELL RECORD (length + entries), prefetched while the previous site is tested.
//
// Pipeline (unchanged): copy engine (cp.async.bulk of whole raw frames, mbarrier complete_tx) -> producer warps
// (raw frame -> float row) -> consumer warps (one mobile atom per lane, frames in order), coupled by full/empty
// mbarriers on a ring of rows; no CTA-wide barrier in the steady state.  The chain check of the anchor form of the
// wrap count W = rint(raw - anchor) against the reference's sequential update (pbc_utils.py:210-217) is the one of
// k_polyhedral_stream.cuh; in-tube frames never need W itself any more, only the certainty that the reference's
// shift caches are the ones the host tables were built from.
#pragma once
#include "k_polyhedral_stream.cuh"

#ifndef SAB_S5_MAXTHREADS
#define SAB_S5_MAXTHREADS 256
#endif
#ifndef SAB_S5_SLEEP_NS
#define SAB_S5_SLEEP_NS 0         // > 0: nanosleep back-off of a waiting warp between two looks at its mbarrier
#endif
#ifndef SAB_S5_MINBLOCKS
#define SAB_S5_MINBLOCKS 3
#endif
#define SAB_S5_MAXRAW 4
#define SAB_S5_MAXROWS 8
#define SAB_CNT_FLAGGED 11        // counters[]: frames handed to k_polyhedral_flagged
#ifndef SAB_S5_KT
#define SAB_S5_KT 3               // register-resident producer tasks per thread
#endif
#ifndef SAB_S5_PREFETCH_ALL
#define SAB_S5_PREFETCH_ALL 0      // 1: every lane prefetches its cell record every frame (not only after a failed first test)
#endif
#define SAB_S5_DELTA22 2.384185791015625e-7f      // 2^-22: every staged coordinate is within this of its float64 value

struct SabStream5Tables {
    int n_task;                   // framework atoms some site uses (M): one producer task and one row entry each
    int stride4;                  // float4 per row: M + A + 1
    int m3;                       // framework coordinates
    const float4 *task;           // [M] {3 * structure index (int bits), anchor x, y, z (exactly the float64 anchors)}
    const uint2 *rec;             // [S] four uint16 BYTE offsets of the site's vertex atoms in a row, positively oriented
    int G;
    const int32_t *cell_off;      // [G^3 + 1]
    const uint16_t *cell_sites;   // CSR entries, ascending per cell
    const uint16_t *cell_rec;     // [G^3][32] fixed-size form of the same lists: {length or 0xffff, entries...} (or null)
    float delta_t;                // validity tube of the producers' check, shrunk by 1e-6 (float rounding)
    int reg_tasks;                // every producer thread owns <= SAB_S5_KT framework atoms and mobile atoms: metadata in registers
    float T1, T0, TD;             // constant thresholds of the pre-test (largest site of the table, |P| <= 1/2)
    int use_f32;
    long long *defer;
    long long defer_cap;
    long long *flagged;           // frames outside the validity tube (settled by k_polyhedral_flagged after this kernel)
    const double *anchor;         // [m3]
    const int32_t *fw_atoms;      // [M]
    const double *prev_raw;       // [m3] carry: raw coordinates of the frame before this batch
    const int32_t *wraps;         // [m3] carry: wrap counts at that frame
    int skip_check_frame0;
    int n_raw, n_rows, n_prod, use_tma;
    double *chunk_exc;            // out [n_chunks][m3][2]
    double delta_d;               // the validity tube in float64 (>= what delta_t admits)
    int n_peer;
    int32_t *peer[SAB_MAX_PEERS];
    int peer_mc;                  // peer[0] is a multicast address (one store reaches every rank: multimem.st)
};

// ---------------------------------------------------------------------------------------------
// FP32 pre-test, generation 5.  With a = vertex 0, B = b - a, C = c - a, D = d - a and P = p - a:
//   det = D . (B x C),  l_b = C . (D x P),  l_c = -B . (D x P),  l_d = P . (B x C),  l_a = det - l_b - l_c - l_d
// are det times the barycentric coordinates of p; p is inside the closed tetrahedron iff all of them have the
// sign of det.  These are the reference's four face-plane values (polyhedral_site.py:107-121) up to a positive
// factor and float64 rounding.  Records are stored with det > 0 at the anchor positions (host).
//
// Reduced separations.  A row holds WRAPPED positions u = x - rint(x - anchor) of the vertex atoms (any integer
// image) and x - rint(x) of the mobile atom.  The reference's vertices are raw + (shift - W) (pbc_utils.py:220): the
// same positions up to integers, and inside the validity tube (|u - anchor| <= delta for every vertex; the producers
// flag every other frame) they span less than 0.45 per axis (host check on the anchor geometry + 2 delta).  Hence each
// true difference B, C, D is the separation of the wrapped positions reduced by its nearest integer, d - rint(d),
// exact in float (|d - rint(d)| <= 1/2 needs no more bits than d).
// Image choice of the atom.  The reference tests the images x + {0,1}^3 against the vertices shifted by
// u = ceil(-min)^+ (tools.py:246-253, pbc_utils.py:221-229).  A point inside the hull is within 0.45 of vertex a on
// every axis, so the only image in reach is the one with P = t - rint(t), t = x - a.  If it is strictly inside,
// x + o = p + u lies in [min + u, max + u], inside [0, 2) by the host check "anchor vertex + delta < 2", so o is 0
// or 1 and the reference tests exactly this image.  If it is outside, every other image differs from it by >= 1 on
// some axis, i.e. lies >= 1/2 - 2^-20 > 0.45 from a on that axis: outside too.
//
// Error budget (delta = 2^-22).  Staged values are float(x) - W or float(x) - floor with |x| < 4 (the producers flag
// every other frame): within delta / 2 of their float64 values.  A difference of two staged values adds one rounding
// (<= delta / 4, |d| < 4); the reduction is exact.  Every component of B, C, D, P therefore carries an input error
// <= 1.25 delta.  With Ee >= |components of B, C, D| and Ep >= |components of P|, first-order propagation through
// the six products of a 3 x 3 determinant plus the roundings of the FMA evaluation (<= 4.5 Ee^2 Ep delta) give
//   |err(l_b, l_c, l_d)| <= (17.25 Ee Ep + 7.5 Ee^2) delta,   |err(det)| <= 24.75 Ee^2 delta,
// and l_a = det - l_b - l_c - l_d adds them up.  The thresholds below use the LARGER constants 28 / 16 / 44 of the
// earlier generations, twice:  T1 = 2 (28 Ee Ep + 16 Ee^2) delta,  TD = 2 * 44 Ee^2 delta,  T0 = 3 T1 + TD.
// A verdict is returned only when every quantity that enters it is beyond its threshold; then the exact plane
// values -- and the float64 reference arithmetic, whose own error is ~1e-9 of the threshold -- have the same signs.
// The fast path uses the bounds for Ee = Ee_max (largest anchor site + tube), Ep = 1/2 (P is a reduced separation);
// a value inside that wider band is re-judged with the site's own Ee, Ep.
// Return: 0 = certainly not contained, 1 = certainly contained, 2 = undecided (exact float64 pass decides).
// ---------------------------------------------------------------------------------------------
// round-half-even to an integer without the conversion pipe: exact for |x| < 2^51 (double), |x| < 2^22 (float)
__device__ __forceinline__ double sab_rint_magic(double x) { return (x + 6755399441055744.0) - 6755399441055744.0; }
__device__ __forceinline__ float sab_rintf_magic(float x) { return (x + 12582912.f) - 12582912.f; }
__device__ __forceinline__ float sab_mic1(float d) { return d - sab_rintf_magic(d); }

// Wait for an mbarrier phase.  SAB_S5_SLEEP_NS > 0 polls with a nanosleep in between instead of the hardware-suspended
// try_wait (whose suspension ends at every mbarrier update of the CTA); measured equal within 2 % (profiles/README.md).
__device__ __forceinline__ void sab_mbar_wait_sleep(uint64_t *bar, unsigned parity)
{
#if SAB_S5_SLEEP_NS > 0
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    for (;;) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(SAB_S5_SLEEP_NS);
    }
#else
    sab_mbar_wait_backoff(bar, parity);
#endif
}

struct SabTet5Off { unsigned o0, o1, o2, o3; };
__device__ __forceinline__ SabTet5Off sab_tet5_unpack(const uint2 r)
{
    SabTet5Off o; o.o0 = r.x & 0xffffu; o.o1 = r.x >> 16; o.o2 = r.y & 0xffffu; o.o3 = r.y >> 16;
    return o;
}

__device__ __forceinline__ int sab_tet5_slow(float Bx, float By, float Bz, float Cx, float Cy, float Cz, float Dx, float Dy,
                                             float Dz, float Px, float Py, float Pz, float det, float lb, float lc, float ld, float la)
{
    const float Ee = fmaxf(fmaxf(fmaxf(fabsf(Bx), fabsf(By)), fmaxf(fabsf(Bz), fabsf(Cx))),
                           fmaxf(fmaxf(fabsf(Cy), fabsf(Cz)), fmaxf(fmaxf(fabsf(Dx), fabsf(Dy)), fabsf(Dz))));
    const float Ep = fmaxf(fabsf(Px), fmaxf(fabsf(Py), fabsf(Pz)));
    if (!(Ee < 0.45f) || !(Ep < 0.75f)) return 2;                      // not a small site (or NaN): budget void
    const float T1 = Ee * (56.f * Ep + 32.f * Ee) * SAB_S5_DELTA22;    // 2 x bound
    const float TD = Ee * Ee * (88.f * SAB_S5_DELTA22);
    const float T0 = 3.f * T1 + TD;
    if (!(fabsf(det) > TD)) return 2;                                  // flat tetrahedron: float64 decides
    if (det < 0.f) { la = -la; lb = -lb; lc = -lc; ld = -ld; }
    if (lb < -T1 || lc < -T1 || ld < -T1 || la < -T0) return 0;
    if (lb > T1 && lc > T1 && ld > T1 && la > T0) return 1;
    return 2;
}

__device__ __forceinline__ int sab_tet5(const char *__restrict__ rowb, const SabTet5Off &o, const float4 x,
                                        const float T1, const float T0, const float TD)
{
    const float4 a = *reinterpret_cast<const float4 *>(rowb + o.o0), b = *reinterpret_cast<const float4 *>(rowb + o.o1);
    const float4 c = *reinterpret_cast<const float4 *>(rowb + o.o2), d = *reinterpret_cast<const float4 *>(rowb + o.o3);
    const float Px = sab_mic1(x.x - a.x), Py = sab_mic1(x.y - a.y), Pz = sab_mic1(x.z - a.z);
    const float Bx = sab_mic1(b.x - a.x), By = sab_mic1(b.y - a.y), Bz = sab_mic1(b.z - a.z);
    const float Cx = sab_mic1(c.x - a.x), Cy = sab_mic1(c.y - a.y), Cz = sab_mic1(c.z - a.z);
    const float Dx = sab_mic1(d.x - a.x), Dy = sab_mic1(d.y - a.y), Dz = sab_mic1(d.z - a.z);
    // explicit FMAs: the library is compiled with -fmad=false for the float64 reference arithmetic
    const float Qx = __fmaf_rn(Dy, Pz, -(Dz * Py)), Qy = __fmaf_rn(Dz, Px, -(Dx * Pz)), Qz = __fmaf_rn(Dx, Py, -(Dy * Px));   // D x P
    const float Nx = __fmaf_rn(By, Cz, -(Bz * Cy)), Ny = __fmaf_rn(Bz, Cx, -(Bx * Cz)), Nz = __fmaf_rn(Bx, Cy, -(By * Cx));   // B x C
    const float lb = __fmaf_rn(Cz, Qz, __fmaf_rn(Cy, Qy, Cx * Qx));
    const float lc = -__fmaf_rn(Bz, Qz, __fmaf_rn(By, Qy, Bx * Qx));
    const float ld = __fmaf_rn(Pz, Nz, __fmaf_rn(Py, Ny, Px * Nx));
    const float det = __fmaf_rn(Dz, Nz, __fmaf_rn(Dy, Ny, Dx * Nx));
    const float la = (det - lb) - (lc + ld);
    const float m = fminf(fminf(lb, lc), ld);
    if (det > TD) {
        if (m > T1 && la > T0) return 1;
        if (m < -T1 || la < -T0) return 0;
    }
    return sab_tet5_slow(Bx, By, Bz, Cx, Cy, Cz, Dx, Dy, Dz, Px, Py, Pz, det, lb, lc, ld, la);
}

__device__ __forceinline__ int sab_cell_of_f4(const float4 x, int G)      // x in [-1/2, 1/2]: the atom's image in [0, 1) decides
{
    const float g = (float)G;
    const float x0 = x.x < 0.f ? x.x + 1.f : x.x, x1 = x.y < 0.f ? x.y + 1.f : x.y, x2 = x.z < 0.f ? x.z + 1.f : x.z;
    int c0 = min(G - 1, (int)(x0 * g)), c1 = min(G - 1, (int)(x1 * g)), c2 = min(G - 1, (int)(x2 * g));
    return (c0 * G + c1) * G + c2;
}

// one framework atom: wrapped float position, tube flag.  W = rint(raw - anchor) wherever the frame is inside the tube.
__device__ __forceinline__ float4 sab_s5_wrap_fw(double r0, double r1, double r2, float a0, float a1, float a2, float delta_t, int &bad)
{
    const float f0 = (float)r0, f1 = (float)r1, f2 = (float)r2;
    const float t0 = f0 - a0, t1 = f1 - a1, t2 = f2 - a2;
    const float w0 = sab_rintf_magic(t0), w1 = sab_rintf_magic(t1), w2 = sab_rintf_magic(t2);
    const float e = fmaxf(fmaxf(fabsf(t0 - w0), fabsf(t1 - w1)), fabsf(t2 - w2));
    const float big = fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fabsf(f2));
    bad |= !(e <= delta_t) | !(big < 4.f);                             // NaN-safe; |x| >= 4 voids the error budget
    return make_float4(f0 - w0, f1 - w1, f2 - w2, 0.f);
}
// one mobile atom: x - rint(x) in [-1/2, 1/2], i.e. x - floor(x) (atom.py:104) up to an integer, to within 2^-23 (the pre-test
// only takes reduced separations from it; the cell lookup adds 1 to a negative component); `far`: too large for float
__device__ __forceinline__ float4 sab_s5_wrap_mobile(double x0, double x1, double x2, int &far)
{
    const float f0 = (float)x0, f1 = (float)x1, f2 = (float)x2;
    far |= !(fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fabsf(f2)) < 4.f);
    return make_float4(f0 - sab_rintf_magic(f0), f1 - sab_rintf_magic(f1), f2 - sab_rintf_magic(f2), 0.f);
}
__device__ __noinline__ float4 sab_s5_wrap_mobile_far(const double *x)     // far from the cell: reduce in float64 first
{
    return make_float4((float)(x[0] - rint(x[0])), (float)(x[1] - rint(x[1])), (float)(x[2] - rint(x[2])), 0.f);
}

// ---- synchronisation helpers on 32-bit shared addresses (computed once per thread, outside the frame loop) ----------
__device__ __forceinline__ unsigned sab_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sab_mbar_wait_a(unsigned a, unsigned parity)
{
    for (;;) {
        unsigned ok;
#if SAB_S5_SLEEP_NS > 0
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(SAB_S5_SLEEP_NS);
#else
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(SAB_STREAM_SUSPEND_NS) : "memory");
        if (ok) break;
#endif
    }
}
__device__ __forceinline__ void sab_mbar_arrive_a(unsigned a)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ float4 sab_lds128(unsigned a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sab_sts128(unsigned a, const float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Row layout (float4 units): [0] {flag, -, -, -} | [1 .. M] framework atoms | [M + 1 .. M + A] mobile atoms | [M + A + 1] spare.
// Site records hold BYTE offsets of the vertex entries from the start of the row (16 * (1 + framework index)).
template <bool ONE_GROUP, bool REC_SMEM>
__global__ void __launch_bounds__(SAB_S5_MAXTHREADS, SAB_S5_MINBLOCKS)
k_polyhedral_assign_stream5(const double *__restrict__ frac, int64_t n_frames, int64_t chunk, int n_atoms, int n_mobile,
                            const int32_t *__restrict__ mobile, int n_sites, SabPolyTables T, SabStream5Tables C,
                            int64_t legacy_init_frame, int32_t *__restrict__ result,
                            unsigned long long *__restrict__ counters, unsigned long long *__restrict__ n_fallback)
{
    extern __shared__ __align__(128) unsigned char sm_s5[];
    const int n3 = 3 * n_atoms, m3 = C.m3;
    double *rawd = reinterpret_cast<double *>(sm_s5);                                        // [n_raw][n3] raw frames (copy engine)
    float4 *rows = reinterpret_cast<float4 *>(rawd + (((size_t)C.n_raw * n3 + 1) & ~(size_t)1)); // [n_rows][stride4]
    float4 *s_task = rows + (size_t)C.n_rows * C.stride4;                                    // [n_task]
    uint2 *s_rec = reinterpret_cast<uint2 *>(s_task + C.n_task);                             // [S] (REC_SMEM)
    int2 *last = reinterpret_cast<int2 *>(s_rec + (REC_SMEM ? n_sites : 0));                 // [A] (!ONE_GROUP)
    int *s_mobcol = reinterpret_cast<int *>(last + (ONE_GROUP ? 0 : n_mobile));              // [A] 3 * structure index
    __shared__ __align__(8) uint64_t raw_full[SAB_S5_MAXRAW], row_full[SAB_S5_MAXROWS], row_empty[SAB_S5_MAXROWS];
    __shared__ int pbad[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_prod_thr = C.n_prod * 32, n_cons = (blockDim.x >> 5) - C.n_prod, n_cons_thr = n_cons * 32;
    const int RS = C.n_raw, ROWS = C.n_rows;
    const int64_t n_chunks = (n_frames + chunk - 1) / chunk;
    const unsigned row_bytes = 16u * (unsigned)C.stride4;
    const unsigned rows_s = sab_saddr(rows), a_full0 = sab_saddr(&row_full[0]), a_empty0 = sab_saddr(&row_empty[0]);

    // ---- one-time setup (all threads) ----------------------------------------------------------------------
    for (int t = tid; t < C.n_task; t += blockDim.x) s_task[t] = C.task[t];
    for (int q = tid; q < n_mobile; q += blockDim.x) s_mobcol[q] = mobile[q] * 3;
    if (REC_SMEM) for (int i = tid; i < n_sites; i += blockDim.x) s_rec[i] = C.rec[i];
    const uint2 *recs = REC_SMEM ? s_rec : C.rec;
    if (tid == 0) {
        for (int i = 0; i < SAB_S5_MAXRAW; i++) sab_mbar_init(&raw_full[i], 1);
        for (int i = 0; i < SAB_S5_MAXROWS; i++) { sab_mbar_init(&row_full[i], C.n_prod); sab_mbar_init(&row_empty[i], n_cons); }
        for (int i = 0; i < 4; i++) pbad[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < C.n_prod) {
        // =========================================== producers ===========================================
        const int pt = tid;
        const unsigned frame_bytes = (unsigned)((size_t)n3 * 8);
        const unsigned a_raw0 = sab_saddr(&raw_full[0]);
        auto issue = [&](int64_t f, int slot) {                 // thread 0 only: raw frame f -> raw slot
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            sab_mbar_expect_tx(&raw_full[slot], frame_bytes);
            sab_bulk_g2s(rawd + (size_t)slot * n3, frac + (size_t)f * n3, frame_bytes, &raw_full[slot]);
        };
        // register-resident tasks: thread pt owns framework atoms pt + k n_prod_thr and the same mobile atoms.  A task that
        // does not exist repeats task 0 / atom 0 (any tube violation it reports is a real one) and writes the spare float4
        // at the end of the row.  r_col / r_mcol: element offset of the atom's x in a raw frame; r_dst / r_mdst: byte offset in a row.
        int r_col[SAB_S5_KT], r_mcol[SAB_S5_KT];
        unsigned r_dst[SAB_S5_KT], r_mdst[SAB_S5_KT];
        float r_anc[SAB_S5_KT][3];
        const unsigned spare = 16u * (unsigned)(C.stride4 - 1);
#pragma unroll
        for (int k = 0; k < SAB_S5_KT; k++) {
            const int t = pt + k * n_prod_thr;
            const bool ok = C.reg_tasks && t < C.n_task, okm = C.reg_tasks && t < n_mobile;
            const float4 tk = s_task[ok ? t : 0];
            r_col[k] = __float_as_int(tk.x);
            r_anc[k][0] = tk.y; r_anc[k][1] = tk.z; r_anc[k][2] = tk.w;
            r_dst[k] = ok ? 16u * (unsigned)(1 + t) : spare;
            r_mcol[k] = s_mobcol[okm ? t : 0];
            r_mdst[k] = okm ? 16u * (unsigned)(1 + C.n_task + t) : spare;
        }
        unsigned it = 0;                                        // frames processed by this CTA so far
        int rslot = 0, wslot = 0;                               // ring positions of frame `it`
        unsigned rpar = 0, wpar = 0;                            // phase parities: (it / RS) & 1, (it / ROWS) & 1
        unsigned row_s = rows_s;                                // shared address of row slot wslot
        for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
            const int64_t c0 = ch * chunk, c1 = min(c0 + chunk, n_frames);
            sab_bar_producers(n_prod_thr);                      // previous chunk fully converted: raw slots idle
            if (pt == 0 && C.use_tma)
                for (int i = 0; i < RS - 1 && c0 + i < c1; i++) issue(c0 + i, (rslot + i) % RS);
            // excursion of a chunk whose frames stay inside the validity tube (frames outside widen it below)
            for (int j = pt; j < m3; j += n_prod_thr) {
                const double anc = C.anchor[j];
                C.chunk_exc[((size_t)ch * m3 + j) * 2] = anc - C.delta_d;
                C.chunk_exc[((size_t)ch * m3 + j) * 2 + 1] = anc + C.delta_d;
            }
            for (int64_t f = c0; f < c1; f++, it++) {
                double *raw = rawd + (size_t)rslot * n3;
                if (C.use_tma) sab_mbar_wait_a(a_raw0 + 8u * (unsigned)rslot, rpar);
                else {                                          // frames the copy engine cannot take: plain loads
                    const double *g = frac + (size_t)f * n3;
                    for (int q = pt; q < n3; q += n_prod_thr) raw[q] = g[q];
                    sab_bar_producers(n_prod_thr);
                }
                if (it >= (unsigned)ROWS) sab_mbar_wait_a(a_empty0 + 8u * (unsigned)wslot, wpar ^ 1u);
                // ---- framework atoms and mobile atoms -> wrapped float32 positions ---------------------------
                int bad = 0;
                if (C.reg_tasks) {                              // every load of the frame is issued before the first use
                    double rr[SAB_S5_KT][3], xx[SAB_S5_KT][3];
#pragma unroll
                    for (int k = 0; k < SAB_S5_KT; k++) {
                        const double *pr = raw + r_col[k], *px = raw + r_mcol[k];
#pragma unroll
                        for (int d = 0; d < 3; d++) { rr[k][d] = pr[d]; xx[k][d] = px[d]; }
                    }
                    int far = 0;
#pragma unroll
                    for (int k = 0; k < SAB_S5_KT; k++)
                        sab_sts128(row_s + r_dst[k], sab_s5_wrap_fw(rr[k][0], rr[k][1], rr[k][2], r_anc[k][0], r_anc[k][1], r_anc[k][2], C.delta_t, bad));
#pragma unroll
                    for (int k = 0; k < SAB_S5_KT; k++)
                        sab_sts128(row_s + r_mdst[k], sab_s5_wrap_mobile(xx[k][0], xx[k][1], xx[k][2], far));
                    if (far)                                    // unwrapped input far from the cell (rare): float64 reduction
                        for (int k = 0; k < SAB_S5_KT; k++)
                            if (r_mdst[k] != spare) sab_sts128(row_s + r_mdst[k], sab_s5_wrap_mobile_far(raw + r_mcol[k]));
                } else {
                    for (int t = pt; t < C.n_task; t += n_prod_thr) {
                        const float4 tk = s_task[t];
                        const double *pr = raw + __float_as_int(tk.x);
                        sab_sts128(row_s + 16u * (unsigned)(1 + t), sab_s5_wrap_fw(pr[0], pr[1], pr[2], tk.y, tk.z, tk.w, C.delta_t, bad));
                    }
                    for (int q = pt; q < n_mobile; q += n_prod_thr) {
                        const double *px = raw + s_mobcol[q];
                        int far = 0;
                        const float4 v = sab_s5_wrap_mobile(px[0], px[1], px[2], far);
                        sab_sts128(row_s + 16u * (unsigned)(1 + C.n_task + q), far ? sab_s5_wrap_mobile_far(px) : v);
                    }
                }
                if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&pbad[it & 3], 1);
                sab_bar_producers(n_prod_thr);                  // pbad complete; everybody is done with the frame before
                const int tube_bad = pbad[it & 3];
                const bool explicit_check = tube_bad || f == c0 || pbad[(it + 3) & 3];
                const int pslot = rslot == 0 ? RS - 1 : rslot - 1;           // raw slot of frame f - 1 = slot of frame f + RS - 1
                if (pt == 0) {
                    pbad[(it + 2) & 3] = 0;
                    const int flag = tube_bad || (T.legacy_init && legacy_init_frame == f);
                    sab_sts128(row_s, make_float4(__int_as_float(flag), 0.f, 0.f, 0.f));
                    if (flag) {                                 // outside the validity tube (or the legacy first-frame rule):
                        atomicAdd(n_fallback, 1ULL);            // the whole frame goes to the generic pass after this kernel
                        C.flagged[atomicAdd(&counters[SAB_CNT_FLAGGED], 1ULL)] = (long long)f;
                    }
                    if (!explicit_check && C.use_tma && f + RS - 1 < c1) issue(f + RS - 1, pslot);
                }
                if (explicit_check) {                           // chain check against the reference's sequential update
                    const double *praw = rawd + (size_t)pslot * n3;                                   // frame f - 1 (f > c0)
                    for (int j = pt; j < m3; j += n_prod_thr) {
                        const int col = C.fw_atoms[j / 3] * 3 + j % 3;
                        const double r = raw[col], anc = C.anchor[j];
                        double p; int Wp;
                        if (f > c0) { p = praw[col]; Wp = (int)rint(p - anc); }
                        else if (c0 == 0) { p = C.prev_raw[j]; Wp = C.wraps[j]; }
                        else { p = frac[(size_t)(c0 - 1) * n3 + col]; Wp = (int)rint(p - anc); }
                        const double diff = r - p, w = rint(diff), phys = diff - w;          // pbc_utils.py:210-216
                        if ((phys >= 0.3 || phys <= -0.3) && !(C.skip_check_frame0 && f == 0))
                            atomicMin(&counters[3], (unsigned long long)f);
                        if ((int)rint(r - anc) - Wp != (int)w) counters[SAB_CNT_INVALID] = 1ULL;
                        if (tube_bad) {                         // outside the tube: the excursion record is explicit
                            const double u = r - rint(r - anc);
                            double *e2 = C.chunk_exc + ((size_t)ch * m3 + j) * 2;      // column j is owned by this thread
                            if (u < e2[0]) e2[0] = u;
                            if (u > e2[1]) e2[1] = u;
                        }
                    }
                    if (C.use_tma) {
                        sab_bar_producers(n_prod_thr);          // frame f - 1 no longer needed
                        if (pt == 0 && f + RS - 1 < c1) issue(f + RS - 1, pslot);
                    }
                }
                __syncwarp();
                if (lane == 0) sab_mbar_arrive_a(a_full0 + 8u * (unsigned)wslot);
                if (++rslot == RS) { rslot = 0; rpar ^= 1u; }
                row_s += row_bytes;
                if (++wslot == ROWS) { wslot = 0; wpar ^= 1u; row_s = rows_s; }
            }
        }
    } else {
        // =========================================== consumers ===========================================
        const int cw = warp - C.n_prod;
        const unsigned FULL = 0xffffffffu;
        const float T1 = C.T1, T0 = C.T0, TD = C.TD;
        const unsigned x0off = 16u * (unsigned)(1 + C.n_task);      // byte offset of mobile atom 0 in a row
        int wslot = 0; unsigned wpar = 0;
        unsigned row_s = rows_s;
        // One atom-frame.  rcx / rcy: {most recent, previous} site as far as certain (-1 = unknown), ofs / ofs2 their records.
        // Returns the result code; `flag` frames are not written here.
        auto atom_frame = [&](const unsigned rs, const int a, const bool valid, const int g0, const int64_t f,
                              int &rcx, int &rcy, SabTet5Off &ofs, SabTet5Off &ofs2) -> int {
            const char *rowb = reinterpret_cast<const char *>(rows) + (rs - rows_s);
            const float4 x = sab_lds128(rs + x0off + 16u * (unsigned)(valid ? a : 0));
            int out = SAB_NONE;
            bool dirty = false;                                  // rcx changed: its record must be reloaded
            // a definite site h: Atom.update_recent_site (atom.py:189-192)
            #define SAB_RECENT_PUSH(h) do { if ((h) != rcx) { rcy = rcx; ofs2 = ofs; rcx = (h); dirty = true; } } while (0)
            int v = 0;
#if SAB_S5_PREFETCH_ALL
            const int mycell = sab_cell_of_f4(x, C.G);
            if (valid && C.cell_rec) asm volatile("prefetch.global.L1 [%0];" ::"l"(C.cell_rec + (size_t)mycell * 32));
#endif
            if (valid && rcx >= 0 && C.use_f32) v = sab_tet5(rowb, ofs, x, T1, T0, TD);
            if (v == 1) out = rcx;
            unsigned m = __ballot_sync(FULL, valid && v != 1);
            if (m != 0u) {                                       // warp-uniform
#if !SAB_S5_PREFETCH_ALL
                const int mycell = sab_cell_of_f4(x, C.G);
                if (valid && v != 1 && C.cell_rec)               // the search's list arrives while the previous site is tested
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(C.cell_rec + (size_t)mycell * 32));
#endif
                // most recent site certainly failed: the reference tests the previous one next
                if (valid && v == 0 && rcx >= 0 && rcy >= 0) {
                    if (!ONE_GROUP) ofs2 = sab_tet5_unpack(recs[rcy]);
                    if (sab_tet5(rowb, ofs2, x, T1, T0, TD) == 1) {
                        out = rcy; const int t = rcx; rcx = rcy; rcy = t; v = 1;
                        const SabTet5Off to = ofs; ofs = ofs2; ofs2 = to;
                    }
                }
                m = __ballot_sync(FULL, valid && v != 1);
                while (m) {                                      // warp-uniform: one search per atom that needs it
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float4 xs = sab_lds128(rs + x0off + 16u * (unsigned)(g0 + src));
                    const int cell = __shfl_sync(FULL, mycell, src);
                    int c = 0, h0 = -1;
                    unsigned und = C.use_f32 ? 0u : 1u;
                    int b = 0, len = -1;
                    if (C.cell_rec && C.use_f32) {               // fixed-size cell record: one load for length and entries
                        const int e = __ldg(C.cell_rec + (size_t)cell * 32 + lane);
                        const int n = __shfl_sync(FULL, e, 0);
                        if (n <= 31) {
                            len = 0;
                            int vv = 0;
                            if (lane >= 1 && lane <= n) vv = sab_tet5(rowb, sab_tet5_unpack(recs[e]), xs, T1, T0, TD);
                            const unsigned m1 = __ballot_sync(FULL, vv == 1);
                            und |= __ballot_sync(FULL, vv == 2);
                            if (m1) h0 = __shfl_sync(FULL, e, __ffs(m1) - 1);
                            c = __popc(m1);
                        }
                    }
                    if (len < 0) { b = __ldg(C.cell_off + cell); len = __ldg(C.cell_off + cell + 1) - b; }   // long list: CSR form
                    if (C.use_f32)
                        for (int i0 = 0; i0 < len; i0 += 32) {
                            const int i = i0 + lane;
                            int s = 0, vv = 0;
                            if (i < len) {
                                s = __ldg(C.cell_sites + b + i);
                                vv = sab_tet5(rowb, sab_tet5_unpack(recs[s]), xs, T1, T0, TD);
                            }
                            unsigned m1 = __ballot_sync(FULL, vv == 1);
                            und |= __ballot_sync(FULL, vv == 2);
                            if (m1 && c == 0) h0 = __shfl_sync(FULL, s, __ffs(m1) - 1);
                            c += __popc(m1);
                        }
                    if (lane == src) {
                        if (und || c > 1) {                      // float32 could not decide, or several sites: exact pass later
                            out = SAB_DEFERRED; rcx = -1; rcy = -1;
                            const unsigned long long di = atomicAdd(&counters[9], 1ULL);
                            if ((long long)di < C.defer_cap) C.defer[di] = (long long)f * n_mobile + a;
                        } else if (c == 0) out = SAB_NONE;
                        else { out = h0; SAB_RECENT_PUSH(h0); }
                    }
                }
            }
            #undef SAB_RECENT_PUSH
            if (ONE_GROUP) { if (dirty && rcx >= 0) ofs = sab_tet5_unpack(recs[rcx]); }
            return out;
        };
        for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
            const int64_t c0 = ch * chunk, c1 = min(c0 + chunk, n_frames);
            if (!ONE_GROUP)
                for (int g0 = cw * 32; g0 < n_mobile; g0 += n_cons_thr) if (g0 + lane < n_mobile) last[g0 + lane] = make_int2(-1, -1);   // owner-private
            int rcx = -1, rcy = -1;
            SabTet5Off ofs = {0u, 0u, 0u, 0u}, ofs2 = {0u, 0u, 0u, 0u};   // ONE_GROUP: carried from frame to frame in registers
            const int a1 = cw * 32 + lane;                       // ONE_GROUP: this lane's atom
            int32_t *resp = result + (size_t)c0 * n_mobile + a1;
            size_t poff = (size_t)c0 * n_mobile + a1;
            for (int64_t f = c0; f < c1; f++) {
                sab_mbar_wait_a(a_full0 + 8u * (unsigned)wslot, wpar);
                const int flag = __float_as_int(sab_lds128(row_s).x);
                if (ONE_GROUP) {
                    const bool valid = a1 < n_mobile;
                    if (flag) { rcx = -1; rcy = -1; }            // outside the validity tube: k_polyhedral_flagged writes the row;
                    else {                                       // the history restarts (any shortcut needs a certain one)
                        const int out = atom_frame(row_s, a1, valid, cw * 32, f, rcx, rcy, ofs, ofs2);
                        if (valid) {
                            *resp = out;
                            // fused result all-gather: the same row goes to every peer's copy of the full array (posted NVLink stores)
                            if (C.n_peer) {
                                if (C.peer_mc) sab_multicast_store(C.peer[0] + poff, out);      // NVLS: the switch replicates the store
                                else {
#pragma unroll
                                    for (int q = 0; q < SAB_MAX_PEERS; q++) if (q < C.n_peer) C.peer[q][poff] = out;   // static indices: the parameter stays in constant memory
                                }
                            }
                        }
                    }
                    resp += n_mobile; poff += n_mobile;
                } else {
                    for (int g0 = cw * 32; g0 < n_mobile; g0 += n_cons_thr) {
                        const int a = g0 + lane;
                        const bool valid = a < n_mobile;
                        if (flag) { if (valid) last[a] = make_int2(-1, -1); continue; }
                        const int2 rc = valid ? last[a] : make_int2(-1, -1);
                        rcx = rc.x; rcy = rc.y;
                        if (rcx >= 0) ofs = sab_tet5_unpack(recs[rcx]);
                        const int out = atom_frame(row_s, a, valid, g0, f, rcx, rcy, ofs, ofs2);
                        if (valid) {
                            const size_t o = (size_t)f * n_mobile + a;
                            result[o] = out;
                            if (C.n_peer) {
                                if (C.peer_mc) sab_multicast_store(C.peer[0] + o, out);
                                else {
#pragma unroll
                                    for (int q = 0; q < SAB_MAX_PEERS; q++) if (q < C.n_peer) C.peer[q][o] = out;
                                }
                            }
                            last[a] = make_int2(rcx, rcy);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) sab_mbar_arrive_a(a_empty0 + 8u * (unsigned)wslot);
                row_s += row_bytes;
                if (++wslot == ROWS) { wslot = 0; wpar ^= 1u; row_s = rows_s; }
            }
        }
    }
}

// Frames the stream kernel listed as outside the validity tube of its cell lists: every mobile atom against every site
// with the fully general float64 test (any polyhedron, all 8 images, legacy first-frame rule), one thread per atom-frame.
__global__ void k_polyhedral_flagged(const double *__restrict__ frac, int n_atoms, int n_mobile, const int32_t *__restrict__ mobile,
                                     int n_sites, SabPolyTables T, int64_t legacy_init_frame, const long long *__restrict__ flagged,
                                     int32_t *__restrict__ result, int32_t *__restrict__ spill, long long spill_capacity,
                                     unsigned long long *__restrict__ counters, const double *__restrict__ anchor, SabPeers peers)
{
    const long long n_flag = (long long)counters[SAB_CNT_FLAGGED];
    const long long total = n_flag * n_mobile;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = flagged[i / n_mobile];
        const int a = (int)(i % n_mobile);
        int c = 0;
        const int out = sab_exhaustive_atom(frac + (size_t)f * n_atoms * 3, nullptr, f, legacy_init_frame, mobile[a], n_sites, T,
                                            spill, spill_capacity, counters, &c, anchor);
        const size_t o = (size_t)f * n_mobile + a;
        result[o] = out;
#pragma unroll
        if (peers.mc) sab_multicast_store(peers.p[0] + o, out);
        else
            for (int q = 0; q < SAB_MAX_PEERS; q++) if (q < peers.n) peers.p[q][o] = out;
    }
}
