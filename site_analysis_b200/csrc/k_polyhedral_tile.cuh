// k_polyhedral_tile.cuh -- the production tetrahedral kernel: "previous site first" over TILES of frames,
// fed by TMA bulk copies.
//
// Same contract as k_polyhedral_assign_seq1 (k_polyhedral_seq1.cuh) and bit-identical results, but
// organised so that (i) global memory is read ONLY by the copy engine: whole frames (and the frames' wrap
// rows) of the next tile arrive in shared memory through cp.async.bulk + mbarrier while the current tile
// is processed, (ii) a CTA synchronises a few times per tile instead of five times per frame, and
// (iii) the common case never touches float64 arithmetic:
//
//  * A CTA owns a contiguous range of frames (a "chunk"); thread a owns mobile atom a and carries
//    `last` = the atom's most recent site as far as it is certain (unknown at the chunk start and
//    after an atom-frame that several sites contain).
//  * Conversion: from the raw frames in shared memory the CTA writes one ROW per frame:
//    for every (framework coordinate, slot shift) pair that some site uses the unwrapped coordinate
//    raw + (double)(shift - W) (pbc_utils.py:220, one float64 rounding) ROUNDED TO FLOAT, followed by the
//    wrapped mobile coordinates (atom.py:104) rounded to float.  The validity tube of the cell lists
//    is checked on the float64 values.  Then the raw buffer is free and the next tile's copy is issued.
//  * Rounds (repeat until every owner has finished the tile):
//      R1  owners test `last` against all remaining frames of the tile (sab_tet_filter32b, independent
//          per frame => instruction-level parallelism) and accept the leading run of "certainly inside"
//          verdicts: that is the reference's answer (it tests the most recent site first and stops at
//          the first hit, site_collection.py:147-152, polyhedral_site_collection.py:104-107).  The first
//          frame with any other verdict -- outside, undecided, no certain history -- queues the
//          atom-frame's cell-list candidates as (atom, frame, site) pairs (warp-cooperative copy).
//      R2  all threads test the queued pairs: FP32 first, and whenever the point is inside the
//          stated error band of a face, the exact float64 test with the reference's operation order
//          (k_polyhedral.cuh: sab_poly_vertices + sab_poly_contains_atom) from global memory (L2).
//      R3  owners read the verdict: none -> None, one -> that site (`last` certain again),
//          several -> spilled list for the priority resolver (`last` unknown); then continue with R1.
//  * Frames outside the validity tube (and the legacy first-frame rule) take the exhaustive generic
//    path per atom, as in the other kernels.
//
// The FP32 test decides only what float64 would decide identically; see sab_tet_filter32b.
#pragma once
#include "k_polyhedral_seq.cuh"

#define SAB_TILE_MAX 4            // frames per tile (compile-time unroll bound)
#define SAB_TILE_PAIRS 1024       // queued (atom, frame, site) pairs per round
#define SAB_TILE_THREADS 256
#ifndef SAB_TILE_MINBLOCKS
#define SAB_TILE_MINBLOCKS 2
#endif

struct SabTileTables {
    int n_used;                   // staged (framework coordinate, slot shift) pairs
    int stride;                   // floats per row: n_used + 3 A, rounded up to a multiple of 4
    int m3;                       // framework coordinates
    const int32_t *col_goff;      // [m3] position of the coordinate inside a frame: 3 * atom + d
    const int32_t *var_off;       // [m3 + 1] CSR over the variants of a coordinate
    const int2 *vars;             // {slot shift, row position j}
    const uint16_t *rec;          // [S][12] byte offsets into a row, vertices in slot order
    int G;
    const int32_t *cell_off;      // [G^3 + 1]
    const uint16_t *cell_sites;   // CSR entries, ascending per cell; lists conservative by >= 1e-6 (float cell lookup)
    const double *anchor;         // [m3]
    double delta;
    int use_f32;
};

// ---------------------------------------------------------------------------------------------
// FP32 pre-test, barycentric form.  With a = vertex 0, B = b - a, C = c - a, D = d - a, P = p - a:
//   det = B . (C x D),  l_d = (B x C) . P,  l_c = -(B x D) . P,  l_b = (C x D) . P,  l_a = det - l_b - l_c - l_d
// and p is inside the closed tetrahedron iff every l_* has the sign of det (or is 0).  The reference's
// face test (polyhedral_site.py:107-121: reject iff dot != 0 and sign(dot) != sign(centre dot)) evaluates
// the same four plane functions up to a positive factor and float64 rounding (~1e-16 relative).
//
// Error budget.  Every staged coordinate (|value| < 4, checked while staging) and p = xf + k are within
// delta = 2^-22 of their float64 values.  With Ee = max |component| of B, C, D and Ep = max |component|
// of P (both < 1), first-order propagation plus the float roundings of the evaluation give
//   |err(l_b, l_c, l_d)| <= (24 Ee Ep + 20 Ee^2) delta,   |err(det)| <= 44 Ee^2 delta,
//   |err(l_a)| <= 3 (24 Ee Ep + 20 Ee^2) delta + 44 Ee^2 delta.
// The thresholds below are TWICE these bounds.  A verdict is returned only when every quantity that
// enters it is beyond its threshold, so the exact plane values have the same signs and the float64
// test (whose own error is ~1e-9 of the threshold) agrees.  Otherwise: undecided.
//
// Image choice (tools.py:246-253, pbc_utils.py:221-229): the reference shifts the vertices by
// u = ceil(-min)^+ per axis and tests the images x + {0,1}^3; the only image within reach of a small
// site is o = rint(centroid + u - x), i.e. p = x + k with k = rint(centroid - x) in unshifted
// coordinates (u is an integer).  If p is strictly inside the hull then x + o = p + u lies in
// [min + u, max + u] which is inside [0, 2) (host check: every anchor vertex + delta < 2), hence
// o is 0 or 1 and the reference tests exactly this image.  If |frac(centroid - x)| is within 0.05 of
// one half the choice of k is not clear-cut in float: undecided.
// Return: 0 = certainly not contained, 1 = certainly contained, 2 = undecided.
// ---------------------------------------------------------------------------------------------
struct SabTetRec { uint2 q0, q1, q2; };                     // 12 uint16 byte offsets into a staged row
__device__ __forceinline__ SabTetRec sab_tet_rec(const uint16_t *__restrict__ rec)
{
    const uint2 *r2 = reinterpret_cast<const uint2 *>(rec);
    SabTetRec r; r.q0 = __ldg(r2); r.q1 = __ldg(r2 + 1); r.q2 = __ldg(r2 + 2);
    return r;
}
__device__ __forceinline__ int sab_tet_filter32b(const float *__restrict__ tab, const SabTetRec &R,
                                                 const float *__restrict__ xf)
{
    float a[3], B[3], C[3], D[3], P[3];
    {
        const uint2 q0 = R.q0, q1 = R.q1, q2 = R.q2;
        #define SAB_LDT(o) (*reinterpret_cast<const float *>(reinterpret_cast<const char *>(tab) + (o)))
        a[0] = SAB_LDT(q0.x & 0xffff); a[1] = SAB_LDT(q0.x >> 16); a[2] = SAB_LDT(q0.y & 0xffff);
        B[0] = SAB_LDT(q0.y >> 16);    B[1] = SAB_LDT(q1.x & 0xffff); B[2] = SAB_LDT(q1.x >> 16);
        C[0] = SAB_LDT(q1.y & 0xffff); C[1] = SAB_LDT(q1.y >> 16);    C[2] = SAB_LDT(q2.x & 0xffff);
        D[0] = SAB_LDT(q2.x >> 16);    D[1] = SAB_LDT(q2.y & 0xffff); D[2] = SAB_LDT(q2.y >> 16);
        #undef SAB_LDT
    }
    float Ee = 0.f, Ep = 0.f;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float cen = 0.25f * ((a[d] + B[d]) + (C[d] + D[d]));
        const float x = xf[d];
        const float t = cen - x, k = rintf(t);
        if (fabsf(t - k) > 0.45f) return 2;
        P[d] = (x + k) - a[d];
        B[d] -= a[d]; C[d] -= a[d]; D[d] -= a[d];
        Ee = fmaxf(Ee, fmaxf(fabsf(B[d]), fmaxf(fabsf(C[d]), fabsf(D[d]))));
        Ep = fmaxf(Ep, fabsf(P[d]));
    }
    // explicit FMAs: the library is compiled with -fmad=false for the float64 reference arithmetic
    const float cd0 = __fmaf_rn(C[1], D[2], -(C[2] * D[1]));          // C x D
    const float cd1 = __fmaf_rn(C[2], D[0], -(C[0] * D[2]));
    const float cd2 = __fmaf_rn(C[0], D[1], -(C[1] * D[0]));
    const float bc0 = __fmaf_rn(B[1], C[2], -(B[2] * C[1]));          // B x C
    const float bc1 = __fmaf_rn(B[2], C[0], -(B[0] * C[2]));
    const float bc2 = __fmaf_rn(B[0], C[1], -(B[1] * C[0]));
    const float bd0 = __fmaf_rn(B[1], D[2], -(B[2] * D[1]));          // B x D
    const float bd1 = __fmaf_rn(B[2], D[0], -(B[0] * D[2]));
    const float bd2 = __fmaf_rn(B[0], D[1], -(B[1] * D[0]));
    float det = __fmaf_rn(B[2], cd2, __fmaf_rn(B[1], cd1, B[0] * cd0));
    float lb = __fmaf_rn(P[2], cd2, __fmaf_rn(P[1], cd1, P[0] * cd0));
    float ld = __fmaf_rn(P[2], bc2, __fmaf_rn(P[1], bc1, P[0] * bc0));
    float lc = -__fmaf_rn(P[2], bd2, __fmaf_rn(P[1], bd1, P[0] * bd0));
    float la = (det - lb) - (lc + ld);
    if (!(Ee < 0.5f)) return 2;                                        // not a small site: budget void
    const float T1 = Ee * (48.f * Ep + 40.f * Ee) * 2.384185791015625e-7f;          // 2 x bound, delta = 2^-22
    const float TD = Ee * Ee * (88.f * 2.384185791015625e-7f);
    const float T0 = 3.f * T1 + TD;
    if (!(fabsf(det) > TD)) return 2;                                  // flat tetrahedron: float64 decides
    if (det < 0.f) { la = -la; lb = -lb; lc = -lc; ld = -ld; }
    if (lb < -T1 || lc < -T1 || ld < -T1 || la < -T0) return 0;
    if (lb > T1 && lc > T1 && ld > T1 && la > T0) return 1;
    return 2;
}

// exact float64 test of one (site, atom) in one frame, operands from global memory (rare path)
__device__ __noinline__ bool sab_tet_exact_global(const double *__restrict__ frame, const int32_t *__restrict__ wrow,
                                                  const SabPolyTables &T, int s, int atom)
{
    double x[3], vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
    int nv;
#pragma unroll
    for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)atom * 3 + d]);
    sab_poly_vertices(frame, T, s, wrow, false, vc, nv, lo, hi, cen);
    return sab_poly_contains_atom(vc, cen, T.simp + (size_t)s * T.fmax * 3, T.n_face[s], lo, hi, x);
}

__device__ __forceinline__ int sab_cell_of_f(const float *x, int G)
{
    const float g = (float)G;
    int c0 = min(G - 1, (int)(x[0] * g)), c1 = min(G - 1, (int)(x[1] * g)), c2 = min(G - 1, (int)(x[2] * g));
    return (c0 * G + c1) * G + c2;
}

// full test of one (site, atom-frame): FP32 first, float64 when undecided
__device__ __forceinline__ bool sab_tile_test(const float *__restrict__ row, const float *__restrict__ xk, int s,
                                              const double *__restrict__ frame, const int32_t *__restrict__ wrow, int atom,
                                              const SabPolyTables &T, const SabTileTables &C)
{
    int v = 2;
    if (C.use_f32) v = sab_tet_filter32b(row, sab_tet_rec(C.rec + (size_t)s * 12), xk);
    if (v == 2) v = sab_tet_exact_global(frame, wrow, T, s, atom);
    return v != 0;
}

// result of one atom-frame whose containing sites were counted (c) and partly kept (h); `last` follows
// the certainty rules of the file header.  Lists longer than KEEP are rescanned in ascending order.
__device__ __noinline__ int sab_tile_finish(int c, int *h, int &last, const float *__restrict__ row, const float *xk,
                                            const double *__restrict__ frame, const int32_t *__restrict__ wrow, int atom,
                                            const SabPolyTables &T, const SabTileTables &C, int32_t *__restrict__ spill,
                                            long long spill_capacity, unsigned long long *__restrict__ counters)
{
    if (c == 0) return SAB_NONE;
    if (c == 1) { last = h[0]; return h[0]; }
    last = -1;                                        // several sites: the resolver decides; history uncertain
    atomicAdd(&counters[0], 1ULL);
    unsigned long long off = atomicAdd(&counters[1], (unsigned long long)(c + 1));
    if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) { counters[2] = 1ULL; return SAB_NONE; }
    spill[off] = c;
    if (c <= SAB_POLY_KEEP) {                         // ascending list positions
        for (int i = 1; i < c; i++) { int v = h[i], j = i; while (j > 0 && h[j - 1] > v) { h[j] = h[j - 1]; j--; } h[j] = v; }
        for (int i = 0; i < c; i++) spill[off + 1 + i] = h[i];
    } else {                                          // atom on a vertex / edge shared by many sites: rescan
        const int cell = sab_cell_of_f(xk, C.G);
        int k = 0;
        for (int i = C.cell_off[cell]; i < C.cell_off[cell + 1] && k < c; i++) {
            const int s = C.cell_sites[i];
            if (sab_tile_test(row, xk, s, frame, wrow, atom, T, C)) spill[off + 1 + k++] = s;
        }
    }
    return SAB_AMBIGUOUS_BASE - (int)off;
}

// ---- mbarrier / bulk-copy helpers (sm_90+ PTX; SASS: SYNCS / UBLKCP) -------------------------------------
__device__ __forceinline__ void sab_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sab_mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sab_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sab_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__global__ void __launch_bounds__(SAB_TILE_THREADS, SAB_TILE_MINBLOCKS)
k_polyhedral_assign_tile(const double *__restrict__ frac, int64_t n_frames, int chunk, int tile, int n_atoms, int n_mobile,
                         const int32_t *__restrict__ mobile, int n_sites, SabPolyTables T, SabTileTables C,
                         const int32_t *__restrict__ wrows, int64_t legacy_init_frame,
                         int32_t *__restrict__ result, int32_t *__restrict__ spill, long long spill_capacity,
                         unsigned long long *__restrict__ counters, unsigned long long *__restrict__ n_fallback)
{
    extern __shared__ __align__(128) unsigned char sm_tile[];
    const int n3 = 3 * n_atoms;
    double *rawd = reinterpret_cast<double *>(sm_tile);                         // [tile][3 N] raw frames (copy engine)
    int32_t *wbuf = reinterpret_cast<int32_t *>(rawd + (size_t)tile * n3);      // [tile][m3] wrap rows (copy engine)
    float *tab = reinterpret_cast<float *>(wbuf + (((size_t)tile * C.m3 + 3) & ~(size_t)3));   // [tile][stride] rows
    int *pairs = reinterpret_cast<int *>(tab + (size_t)tile * C.stride);        // [SAB_TILE_PAIRS]
    int *cnt = pairs + SAB_TILE_PAIRS;                                          // [A] hits of the pending atom-frame
    int *hits = cnt + n_mobile;                                                 // [A][KEEP]
    int *last = hits + (size_t)n_mobile * SAB_POLY_KEEP;                        // [A]
    int *pos = last + n_mobile;                                                 // [A] next frame of the tile; bit 8 = verdict pending
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int n_pairs[2];
    __shared__ int bad_mask;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int64_t n_chunks = (n_frames + chunk - 1) / chunk;
    unsigned phase = 0;
    if (tid == 0) { sab_mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    // the copy engine needs 16-byte aligned addresses and sizes; a tile that does not qualify is copied by the threads
    auto tma_ok = [&](int64_t f0, int tn) {
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(frac + (size_t)f0 * n3), a1 = reinterpret_cast<uintptr_t>(wrows + (size_t)f0 * C.m3);
        return ((a0 | a1 | reinterpret_cast<uintptr_t>(wbuf)) & 15) == 0 && (((size_t)tn * n3 * 8) & 15) == 0 &&
               (((size_t)tn * C.m3 * 4) & 15) == 0;
    };
    auto issue = [&](int64_t f0, int tn) {           // thread 0 only
        const unsigned b0 = (unsigned)((size_t)tn * n3 * 8), b1 = (unsigned)((size_t)tn * C.m3 * 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sab_mbar_expect_tx(&mbar, b0 + b1);
        sab_bulk_g2s(rawd, frac + (size_t)f0 * n3, b0, &mbar);
        sab_bulk_g2s(wbuf, wrows + (size_t)f0 * C.m3, b1, &mbar);
    };

    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const int64_t c0 = ch * chunk, c1 = min(c0 + (int64_t)chunk, n_frames);
        for (int a = tid; a < n_mobile; a += nthr) last[a] = -1;              // owner-private
        {
            const int tn0 = (int)min((int64_t)tile, c1 - c0);
            if (tid == 0 && tma_ok(c0, tn0)) issue(c0, tn0);                 // buffers are free: see the barrier after conversion
        }
        for (int64_t f0 = c0; f0 < c1; f0 += tile) {
            const int tn = (int)min((int64_t)tile, c1 - f0);
            const double *gframe = frac + (size_t)f0 * n3;
            const int32_t *gwrow = wrows + (size_t)f0 * C.m3;
            int32_t *res = result + (size_t)f0 * n_mobile;
            // every thread is past the final barrier A of the previous tile: nobody reads rows, bad_mask or counters
            if (tid == 0) { bad_mask = 0; n_pairs[0] = 0; n_pairs[1] = 0; }
            if (tma_ok(f0, tn)) {
                sab_mbar_wait(&mbar, phase);
                phase ^= 1;
            } else {
                for (int q = tid; q < tn * n3; q += nthr) rawd[q] = gframe[q];
                for (int q = tid; q < tn * C.m3; q += nthr) wbuf[q] = gwrow[q];
            }
            __syncthreads();
            // ---- conversion: raw frames -> float rows -------------------------------------------------
            {
                int bad = 0;
                for (int i = tid; i < C.m3; i += nthr) {
                    const int goff = __ldg(C.col_goff + i), vb = __ldg(C.var_off + i), ve = __ldg(C.var_off + i + 1);
                    const double anc = __ldg(C.anchor + i);
                    double r[SAB_TILE_MAX]; int w[SAB_TILE_MAX];
#pragma unroll
                    for (int k = 0; k < SAB_TILE_MAX; k++)
                        if (k < tn) {
                            r[k] = rawd[k * n3 + goff]; w[k] = wbuf[k * C.m3 + i];
                            if (fabs((r[k] - (double)w[k]) - anc) > C.delta) bad |= 1 << k;
                        }
                    for (int v = vb; v < ve; v++) {
                        const int2 e = __ldg(C.vars + v);
#pragma unroll
                        for (int k = 0; k < SAB_TILE_MAX; k++)
                            if (k < tn) {
                                const double val = r[k] + (double)(e.x - w[k]);            // pbc_utils.py:220, one rounding
                                if (!(fabs(val) < 3.5)) bad |= 1 << k;                      // FP32 error budget assumes |coordinate| < 4
                                tab[k * C.stride + e.y] = (float)val;
                            }
                    }
                }
                for (int q = tid; q < 3 * n_mobile; q += nthr) {
                    const int a = q / 3, goff = __ldg(mobile + a) * 3 + (q - 3 * a);
#pragma unroll
                    for (int k = 0; k < SAB_TILE_MAX; k++)
                        if (k < tn) tab[k * C.stride + C.n_used + q] = (float)sab_pymod1(rawd[k * n3 + goff]);   // atom.py:104
                }
                if (T.legacy_init && legacy_init_frame >= f0 && legacy_init_frame < f0 + tn) bad |= 1 << (int)(legacy_init_frame - f0);
                if (bad) atomicOr(&bad_mask, bad);
            }
            for (int a = tid; a < n_mobile; a += nthr) pos[a] = 0;
            __syncthreads();                                     // rows complete; raw buffers consumed
            if (tid == 0) {
                if (f0 + tile < c1) {                            // next tile of this chunk: lands during the rounds
                    const int tn2 = (int)min((int64_t)tile, c1 - (f0 + tile));
                    if (tma_ok(f0 + tile, tn2)) issue(f0 + tile, tn2);
                }
                if (bad_mask) atomicAdd(n_fallback, (unsigned long long)__popc(bad_mask));
            }
            const int badm = bad_mask;
            // ---- rounds -----------------------------------------------------------------------------
            for (int round = 0;; round++) {
                int *np = &n_pairs[round & 1];
                int pending = 0;
                for (int a0 = tid - lane; a0 < n_mobile; a0 += nthr) {       // warp-uniform trip count
                    const int a = a0 + lane;
                    int m_len = 0, m_b = 0, m_off = 0, m_tag = 0;            // candidate list to queue (this lane)
                    if (a < n_mobile) {
                        int k = pos[a], la = last[a];
                        const float *xa = tab + C.n_used + 3 * a;
                        if (k & 256) {                               // R3: verdict of the queued atom-frame
                            k &= 255;
                            const int out = sab_tile_finish(cnt[a], hits + a * SAB_POLY_KEEP, la, tab + k * C.stride, xa + k * C.stride,
                                                            gframe + (size_t)k * n3, gwrow + (size_t)k * C.m3, mobile[a], T, C,
                                                            spill, spill_capacity, counters);
                            res[k * n_mobile + a] = out;
                            k++;
                        }
                        while (k < tn) {                             // R1
                            if (la >= 0 && C.use_f32) {
                                const SabTetRec R = sab_tet_rec(C.rec + (size_t)la * 12);
                                int ok = 0;
#pragma unroll
                                for (int kk = 0; kk < SAB_TILE_MAX; kk++)
                                    if (kk >= k && kk < tn && !((badm >> kk) & 1))
                                        ok |= (sab_tet_filter32b(tab + kk * C.stride, R, xa + kk * C.stride) == 1) << kk;
                                while (k < tn && ((ok >> k) & 1)) { res[k * n_mobile + a] = la; k++; }
                                if (k == tn) break;
                            }
                            if ((badm >> k) & 1) {                   // outside the validity tube: generic path
                                int c = 0;
                                const int out = sab_exhaustive_atom(gframe + (size_t)k * n3, gwrow + (size_t)k * C.m3, f0 + k,
                                                                    legacy_init_frame, mobile[a], n_sites, T, spill, spill_capacity,
                                                                    counters, &c);
                                if (c == 1) la = out; else if (c > 1) la = -1;
                                res[k * n_mobile + a] = out;
                                k++;
                                continue;
                            }
                            // full search over the cell list of the atom-frame
                            const float *xk = xa + k * C.stride;
                            const int cell = sab_cell_of_f(xk, C.G);
                            const int b = __ldg(C.cell_off + cell), len = __ldg(C.cell_off + cell + 1) - b;
                            if (len == 0) { res[k * n_mobile + a] = SAB_NONE; k++; continue; }
                            const int off = atomicAdd(np, len);
                            if (off + len <= SAB_TILE_PAIRS) {
                                m_len = len; m_b = b; m_off = off; m_tag = (a << 19) | (k << 16);
                                cnt[a] = 0;
                                k |= 256; pending = 1;
                                break;
                            }
                            for (int i = off; i < SAB_TILE_PAIRS; i++) pairs[i] = -1;      // queue full: test the list here
                            int c = 0, h[SAB_POLY_KEEP];
                            for (int i = 0; i < len; i++) {
                                const int s = __ldg(C.cell_sites + b + i);
                                if (sab_tile_test(tab + k * C.stride, xk, s, gframe + (size_t)k * n3, gwrow + (size_t)k * C.m3, mobile[a], T, C)) {
                                    if (c < SAB_POLY_KEEP) h[c] = s;
                                    c++;
                                }
                            }
                            const int out = sab_tile_finish(c, h, la, tab + k * C.stride, xk, gframe + (size_t)k * n3,
                                                            gwrow + (size_t)k * C.m3, mobile[a], T, C, spill, spill_capacity, counters);
                            res[k * n_mobile + a] = out;
                            k++;
                        }
                        pos[a] = k;
                        last[a] = la;
                    }
                    // warp-cooperative copy of the queued candidate lists
                    unsigned m = __ballot_sync(0xffffffffu, m_len > 0);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int len = __shfl_sync(0xffffffffu, m_len, src), b = __shfl_sync(0xffffffffu, m_b, src);
                        const int off = __shfl_sync(0xffffffffu, m_off, src), tag = __shfl_sync(0xffffffffu, m_tag, src);
                        for (int i = lane; i < len; i += 32) pairs[off + i] = tag | (int)__ldg(C.cell_sites + b + i);
                    }
                }
                if (__syncthreads_or(pending) == 0) break;       // barrier A: queue complete / tile finished
                const int total = min(*np, SAB_TILE_PAIRS);
                if (tid == 0) n_pairs[(round + 1) & 1] = 0;
                // ---- R2: queued pairs over all threads ---------------------------------------------
                for (int p = tid; p < total; p += nthr) {
                    const int e = pairs[p];
                    if (e < 0) continue;
                    const int a = e >> 19, k = (e >> 16) & 7, s = e & 0xffff;
                    if (sab_tile_test(tab + k * C.stride, tab + k * C.stride + C.n_used + 3 * a, s, gframe + (size_t)k * n3,
                                      gwrow + (size_t)k * C.m3, mobile[a], T, C)) {
                        const int c = atomicAdd(&cnt[a], 1);
                        if (c < SAB_POLY_KEEP) hits[a * SAB_POLY_KEEP + c] = s;
                    }
                }
                __syncthreads();                                 // barrier B: verdicts complete
            }
        }
        __syncthreads();                                         // chunk done: raw buffers idle before the next chunk's copy
    }
}
