// k_polyhedral_tile.cuh -- the production tetrahedral kernel: "previous site first" over TILES of frames,
// fed by TMA bulk copies, float32 in the common case, float64 only where float32 cannot decide.
//
// Same contract as k_polyhedral_assign_seq1 (k_polyhedral_seq1.cuh) and bit-identical results:
//
//  * Global memory is read by the copy engine only: the raw frames (and their wrap rows) of the NEXT tile
//    arrive in shared memory through cp.async.bulk + mbarrier while the current tile is processed.
//  * A CTA owns a contiguous range of frames (a "chunk"); thread a owns mobile atom a and carries
//    `last` = the atom's most recent site as far as it is certain (unknown at the chunk start and
//    after an atom-frame that several sites contain or that float32 could not decide).
//  * Conversion: from the raw frames the CTA writes one float ROW per frame: for every (framework
//    coordinate, slot shift) pair that some site uses, float(raw - W) + shift (within 2^-22 of the
//    reference's float64 value raw + (shift - W), pbc_utils.py:220), followed by the wrapped mobile
//    coordinates (atom.py:104).  The validity tube of the cell lists is checked here (conservatively).
//  * Rounds.  Owners emit work items, all threads process them, owners read the verdicts:
//      round 0    every owner tests `last` against all frames of the tile in-thread (independent per frame)
//                 and accepts the leading run of "certainly inside" verdicts: that is the reference's answer
//                 (it tests the most recent site first and stops at the first hit, site_collection.py:147-152,
//                 polyhedral_site_collection.py:104-107).
//      SEARCH     the first frame with any other verdict (outside, undecided, no certain history) queues the
//                 atom-frame's cell list as blocks of <= 16 candidates; verdict none -> None, one -> that site
//                 (`last` certain again), 2..4 -> spilled list for the priority resolver (`last` unknown),
//                 any candidate undecided in float32 (or more than 4 hits) -> the atom-frame is DEFERRED.
//      CONTINUE   after a search the remaining frames of the tile are queued as single (atom, frame, site)
//                 tests of the new `last`, so later rounds stay lane-dense.
//  * DEFERRED atom-frames are appended to a list and settled after the kernel by k_polyhedral_deferred with
//    the exact float64 test in the reference's operation order over the whole cell list
//    (k_polyhedral.cuh: sab_poly_vertices + sab_poly_contains_atom).  The float32 test only ever returns a
//    verdict that float64 would return identically (see sab_tet_filter32b), so results are bit-identical
//    to the all-float64 kernels.
//  * Frames outside the validity tube (and the legacy first-frame rule) take the exhaustive generic
//    path per atom, as in the other kernels.
#pragma once
#include "k_polyhedral_seq.cuh"

#define SAB_TILE_MAX 4            // frames per tile (compile-time unroll bound)
#define SAB_TILE_BLOCKS 256       // queued candidate blocks (<= 16 candidates each) per round
#define SAB_TILE_SINGLES 1024     // queued single tests per round
#define SAB_TILE_THREADS 256
#ifndef SAB_TILE_MINBLOCKS
#define SAB_TILE_MINBLOCKS 2
#endif
#define SAB_DEFERRED INT_MIN      // result marker, replaced by k_polyhedral_deferred before anything reads it

struct SabTileTables {
    int n_used;                   // staged (framework coordinate, slot shift) pairs
    int stride;                   // floats per row: n_used + 3 A, rounded up to a multiple of 4
    int m3;                       // framework coordinates
    const int4 *vars;             // [n_used] {3 * atom + d, framework column i, float bits of the shift, float bits of anchor[i]}
    const uint16_t *rec;          // [S][12] byte offsets into a row, vertices in slot order
    int G;
    const int32_t *cell_off;      // [G^3 + 1]
    const uint16_t *cell_sites;   // CSR entries, ascending per cell; lists conservative by >= 1e-6 (float cell lookup)
    float delta_f;                // validity radius, shrunk by the float error of the check
    int use_f32;
    long long *defer;             // deferred atom-frames: frame * A + atom
    long long defer_cap;
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

// exact float64 test of one (site, atom) in one frame, operands from global memory (deferred pass)
__device__ __noinline__ bool sab_tet_exact_global(const double *__restrict__ frame, const int32_t *__restrict__ wrow,
                                                  const SabPolyTables &T, int s, int atom,
                                                  const double *__restrict__ anchor = nullptr)
{
    double x[3], vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
    int nv;
#pragma unroll
    for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)atom * 3 + d]);
    sab_poly_vertices(frame, T, s, wrow, false, vc, nv, lo, hi, cen, anchor);
    return sab_poly_contains_atom(vc, cen, T.simp + (size_t)s * T.fmax * 3, T.n_face[s], lo, hi, x);
}

__device__ __forceinline__ int sab_cell_of_f(const float *x, int G)
{
    const float g = (float)G;
    int c0 = min(G - 1, (int)(x[0] * g)), c1 = min(G - 1, (int)(x[1] * g)), c2 = min(G - 1, (int)(x[2] * g));
    return (c0 * G + c1) * G + c2;
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

// spill a short list of containing sites (2..KEEP, any order) in ascending list position
__device__ __noinline__ int sab_tile_spill(int c, int *h, int32_t *__restrict__ spill, long long spill_capacity,
                                           unsigned long long *__restrict__ counters)
{
    atomicAdd(&counters[0], 1ULL);
    unsigned long long off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(c));
    if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) { counters[2] = 1ULL; return SAB_NONE; }
    spill[off] = c;
    for (int i = 1; i < c; i++) { int v = h[i], j = i; while (j > 0 && h[j - 1] > v) { h[j] = h[j - 1]; j--; } h[j] = v; }
    for (int i = 0; i < c; i++) spill[off + 1 + i] = h[i];
    return SAB_AMBIGUOUS_BASE - (int)off;
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
    int *blocks = reinterpret_cast<int *>(tab + (size_t)tile * C.stride);       // [SAB_TILE_BLOCKS][3] {tag, first entry, count}
    int *singles = blocks + 3 * SAB_TILE_BLOCKS;                                // [SAB_TILE_SINGLES] (atom, frame, site)
    int *cnt = singles + SAB_TILE_SINGLES;                                      // [A] hits of the pending search; bit 16 = undecided
    int *hits = cnt + n_mobile;                                                 // [A][KEEP]
    int *okb = hits + (size_t)n_mobile * SAB_POLY_KEEP;                         // [A] frames on which `last` certainly holds
    int *last = okb + n_mobile;                                                 // [A]
    int *pos = last + n_mobile;                                                 // [A] next frame | state << 8
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int n_blocks[2], n_singles[2];
    __shared__ int bad_mask;
    const int tid = threadIdx.x, nthr = blockDim.x;
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
            if (tid == 0 && tma_ok(c0, tn0)) issue(c0, tn0);                 // buffers are idle here
        }
        for (int64_t f0 = c0; f0 < c1; f0 += tile) {
            const int tn = (int)min((int64_t)tile, c1 - f0);
            const double *gframe = frac + (size_t)f0 * n3;
            const int32_t *gwrow = wrows + (size_t)f0 * C.m3;
            int32_t *res = result + (size_t)f0 * n_mobile;
            // every thread is past the last barrier of the previous tile: nobody reads rows, bad_mask or the queues
            if (tid == 0) { bad_mask = 0; n_blocks[0] = n_blocks[1] = 0; n_singles[0] = n_singles[1] = 0; }
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
                for (int j = tid; j < C.n_used; j += nthr) {
                    const int4 e = __ldg(C.vars + j);
                    const float sh = __int_as_float(e.z), anc = __int_as_float(e.w);
#pragma unroll
                    for (int k = 0; k < SAB_TILE_MAX; k++)
                        if (k < tn) {
                            const float base = (float)(rawd[k * n3 + e.x] - (double)wbuf[k * C.m3 + e.y]);
                            if (fabsf(base - anc) > C.delta_f) bad |= 1 << k;
                            tab[k * C.stride + j] = base + sh;
                        }
                }
                for (int q = tid; q < 3 * n_mobile; q += nthr) {
                    const int a = q / 3, goff = __ldg(mobile + a) * 3 + (q - 3 * a);
#pragma unroll
                    for (int k = 0; k < SAB_TILE_MAX; k++)
                        if (k < tn) { const double x = rawd[k * n3 + goff]; tab[k * C.stride + C.n_used + q] = (float)(x - floor(x)); }   // atom.py:104
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
                int *nb = &n_blocks[round & 1], *ns = &n_singles[round & 1];
                int pending = 0;
                for (int a = tid; a < n_mobile; a += nthr) {
                    // st: 0 idle, 1 search pending, 2 continuation pending, 3 / 4 search / continuation found its queue full
                    int k = pos[a] & 255, st = pos[a] >> 8, la = last[a];
                    const float *xa = tab + C.n_used + 3 * a;
                    bool walk = la >= 0 && st != 3;
                    if (st == 1) {                                // verdict of the search at frame k
                        const int cv = cnt[a], c = cv & 0xffff;
                        int out;
                        if ((cv >> 16) || c > SAB_POLY_KEEP) {    // float32 could not decide (or a long list): exact pass later
                            out = SAB_DEFERRED; la = -1;
                            const unsigned long long di = atomicAdd(&counters[9], 1ULL);
                            if ((long long)di < C.defer_cap) C.defer[di] = (long long)(f0 + k) * n_mobile + a;
                        } else if (c == 0) out = SAB_NONE;
                        else if (c == 1) { out = hits[a * SAB_POLY_KEEP]; la = out; }
                        else { out = sab_tile_spill(c, hits + a * SAB_POLY_KEEP, spill, spill_capacity, counters); la = -1; }
                        res[k * n_mobile + a] = out;
                        k++;
                        walk = la >= 0;
                    } else if (st == 2) {                         // verdicts of the queued tests of `last`
                        const int ok = okb[a];
                        while (k < tn && ((ok >> k) & 1)) { res[k * n_mobile + a] = la; k++; }
                        walk = false;
                    }
                    st = 0;
                    while (k < tn) {
                        if ((badm >> k) & 1) {                    // outside the validity tube: generic path
                            int c = 0;
                            const int out = sab_exhaustive_atom(gframe + (size_t)k * n3, gwrow + (size_t)k * C.m3, f0 + k,
                                                                legacy_init_frame, mobile[a], n_sites, T, spill, spill_capacity,
                                                                counters, &c);
                            if (c == 1) la = out; else if (c > 1) la = -1;
                            res[k * n_mobile + a] = out;
                            k++;
                            walk = la >= 0;
                            continue;
                        }
                        if (walk && C.use_f32) {
                            if (round == 0) {                     // dense: test in-thread, frames are independent
                                const SabTetRec R = sab_tet_rec(C.rec + (size_t)la * 12);
                                int ok = 0;
#pragma unroll
                                for (int kk = 0; kk < SAB_TILE_MAX; kk++)
                                    if (kk >= k && kk < tn && !((badm >> kk) & 1))
                                        ok |= (sab_tet_filter32b(tab + kk * C.stride, R, xa + kk * C.stride) == 1) << kk;
                                while (k < tn && ((ok >> k) & 1)) { res[k * n_mobile + a] = la; k++; }
                                walk = false;
                                continue;
                            }
                            int n = 0;                            // sparse: queue the tests of the frames up to the next bad one
                            while (k + n < tn && !((badm >> (k + n)) & 1)) n++;
                            const int off = atomicAdd(ns, n);
                            if (off + n <= SAB_TILE_SINGLES) {
                                for (int i = 0; i < n; i++) singles[off + i] = (a << 19) | ((k + i) << 16) | la;
                                okb[a] = 0;
                                st = 2;
                            } else st = 4;                        // queue full: retry next round
                            break;
                        }
                        // search: the cell list of the atom-frame, in blocks of 16 candidates
                        const int cell = sab_cell_of_f(xa + k * C.stride, C.G);
                        const int b = __ldg(C.cell_off + cell), len = __ldg(C.cell_off + cell + 1) - b;
                        if (len == 0) { res[k * n_mobile + a] = SAB_NONE; k++; walk = la >= 0; continue; }
                        const int nbk = (len + 15) >> 4;
                        const int off = atomicAdd(nb, nbk);
                        if (off + nbk <= SAB_TILE_BLOCKS) {
                            for (int i = 0; i < nbk; i++) {
                                blocks[3 * (off + i)] = (a << 19) | (k << 16);
                                blocks[3 * (off + i) + 1] = b + 16 * i;
                                blocks[3 * (off + i) + 2] = min(16, len - 16 * i);
                            }
                            cnt[a] = 0;
                            st = 1;
                        } else {
                            for (int i = off; i < SAB_TILE_BLOCKS; i++) blocks[3 * i + 2] = 0;   // unusable tail of the queue
                            st = 3;                               // retry next round
                        }
                        break;
                    }
                    if (st) pending = 1;
                    pos[a] = k | (st << 8);
                    last[a] = la;
                }
                if (__syncthreads_or(pending) == 0) break;       // barrier A: queues complete / tile finished
                const int nblk = min(*nb, SAB_TILE_BLOCKS), nsing = min(*ns, SAB_TILE_SINGLES);
                if (tid == 0) { n_blocks[(round + 1) & 1] = 0; n_singles[(round + 1) & 1] = 0; }
                // ---- all threads: queued tests ------------------------------------------------------
                for (int p = tid; p < nblk * 16; p += nthr) {
                    const int d = p >> 4, i = p & 15;
                    if (i >= blocks[3 * d + 2]) continue;
                    const int tag = blocks[3 * d], a = tag >> 19, k = (tag >> 16) & 7;
                    const int s = __ldg(C.cell_sites + blocks[3 * d + 1] + i);
                    const int v = C.use_f32 ? sab_tet_filter32b(tab + k * C.stride, sab_tet_rec(C.rec + (size_t)s * 12),
                                                                tab + k * C.stride + C.n_used + 3 * a) : 2;
                    if (v == 1) {
                        const int c = atomicAdd(&cnt[a], 1) & 0xffff;
                        if (c < SAB_POLY_KEEP) hits[a * SAB_POLY_KEEP + c] = s;
                    } else if (v == 2) atomicOr(&cnt[a], 0x10000);
                }
                for (int p = tid; p < nsing; p += nthr) {
                    const int e = singles[p], a = e >> 19, k = (e >> 16) & 7, s = e & 0xffff;
                    if (sab_tet_filter32b(tab + k * C.stride, sab_tet_rec(C.rec + (size_t)s * 12),
                                          tab + k * C.stride + C.n_used + 3 * a) == 1) atomicOr(&okb[a], 1 << k);
                }
                __syncthreads();                                 // barrier B: verdicts complete
            }
        }
        __syncthreads();                                         // chunk done: raw buffers idle before the next chunk's copy
    }
}

// Settles the DEFERRED atom-frames of k_polyhedral_assign_tile with the exact float64 test over the whole cell
// list: one warp per entry, one candidate per lane.  `scan` != 0: the list overflowed -- every warp scans a slice
// of the result array for the marker instead (adversarial inputs only).
#ifndef SAB_MAX_PEERS
#define SAB_MAX_PEERS 7
#endif
struct SabPeers { int n; int32_t *p[SAB_MAX_PEERS]; int mc; };   // other ranks' slots for this rank's result rows (multi-GPU);
                                                                 // mc: p[0] is ONE multicast address that reaches every rank
// One 32-bit store through a multicast mapping (NVLS): the switch replicates it into every rank's copy of the array.
__device__ __forceinline__ void sab_multicast_store(int32_t *mc_addr, int v)
{
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}

__global__ void k_polyhedral_deferred(const double *__restrict__ frac, int64_t n_frames, int n_atoms, int n_mobile,
                                      const int32_t *__restrict__ mobile, SabPolyTables T, int G,
                                      const int32_t *__restrict__ cell_off, const uint16_t *__restrict__ cell_sites,
                                      const int32_t *__restrict__ wrows, int m3, const long long *__restrict__ defer,
                                      long long defer_cap, int32_t *__restrict__ result, int32_t *__restrict__ spill,
                                      long long spill_capacity, unsigned long long *__restrict__ counters,
                                      const double *__restrict__ anchor = nullptr,   // wrows == nullptr: W = rint(raw - anchor)
                                      SabPeers peers = SabPeers{0, {nullptr}, 0})
{
    const long long n_def = (long long)counters[9];
    if (n_def == 0) return;
    const bool scan = n_def > defer_cap;
    const long long total = scan ? (long long)n_frames * n_mobile : n_def;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w0 = warp; w0 < total; w0 += n_warps) {
        const long long idx = scan ? w0 : defer[w0];
        if (result[idx] != SAB_DEFERRED) continue;               // warp-uniform
        const int64_t f = idx / n_mobile; const int a = (int)(idx - f * n_mobile);
        const double *frame = frac + (size_t)f * n_atoms * 3;
        const int32_t *wrow = wrows ? wrows + (size_t)f * m3 : nullptr;
        const int atom = mobile[a];
        double x[3];
#pragma unroll
        for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)atom * 3 + d]);
        const int cell = sab_cell_of(x, G);
        const int b = cell_off[cell], len = cell_off[cell + 1] - b;
        int c = 0, first = -1;
        for (int i0 = 0; i0 < len; i0 += 32) {                   // pass 1: count
            const int i = i0 + lane;
            bool in = false;
            if (i < len) in = sab_tet_exact_global(frame, wrow, T, cell_sites[b + i], atom, anchor);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (m && first < 0) first = cell_sites[b + i0 + __ffs(m) - 1];
            c += __popc(m);
        }
        int out;
        if (c == 0) out = SAB_NONE;
        else if (c == 1) out = first;
        else {
            unsigned long long off = 0;
            if (lane == 0) { atomicAdd(&counters[0], 1ULL); off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(c)); }
            off = __shfl_sync(0xffffffffu, off, 0);
            if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) { if (lane == 0) counters[2] = 1ULL; out = SAB_NONE; }
            else {
                if (lane == 0) spill[off] = c;
                int k = 0;
                for (int i0 = 0; i0 < len; i0 += 32) {           // pass 2: ascending list positions
                    const int i = i0 + lane;
                    bool in = false;
                    if (i < len) in = sab_tet_exact_global(frame, wrow, T, cell_sites[b + i], atom, anchor);
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in) spill[off + 1 + k + __popc(m & ((1u << lane) - 1))] = cell_sites[b + i];
                    k += __popc(m);
                }
                out = SAB_AMBIGUOUS_BASE - (int)off;
            }
        }
        if (lane == 0) {
            result[idx] = out;
#pragma unroll
            if (peers.mc) sab_multicast_store(peers.p[0] + idx, out);
            else
                for (int q = 0; q < SAB_MAX_PEERS; q++) if (q < peers.n) peers.p[q][idx] = out;
        }
    }
}
