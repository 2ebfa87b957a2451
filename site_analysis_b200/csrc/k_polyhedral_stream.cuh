// k_polyhedral_stream.cuh -- the production tetrahedral kernel: a warp-specialised frame pipeline.
//
// Same contract and bit-identical results as k_polyhedral_assign_tile (k_polyhedral_tile.cuh), but
//   * ONE pass over the trajectory: the PBC wrap scan (k_wrap.cuh) is fused in, so the only global traffic is
//     the raw frames in (copy engine) and the int32 site ids out;
//   * no CTA-wide barrier in the steady state: producer warps turn raw frames into float rows, consumer
//     warps own 32 mobile atoms each and walk the frames on their own, coupled only through mbarriers
//     on a ring of row slots.
//
// A CTA owns a contiguous chunk of frames.
//   copy engine  cp.async.bulk of whole raw frames into a ring of raw slots (mbarrier complete_tx).
//   producers    (P warps) per frame, for every referenced framework coordinate j: raw, W = rint(raw - anchor[j]),
//                u = raw - W, excursion min/max of u, validity tube |float(u) - anchor| <= delta of the cell lists,
//                and the float row entries float(u) + shift for every slot shift in use (the unwrapped vertex
//                coordinate raw + (shift - W), pbc_utils.py:220, to within 2^-22); then the wrapped mobile
//                coordinates (atom.py:104).  These are the operands of sab_tet_filter32b.
//       CHAIN CHECK.  The reference counts wraps sequentially, W_t = W_{t-1} + rint(raw_t - raw_{t-1}), and
//                invalidates on a step >= 0.3 (pbc_utils.py:210-217).  W = rint(raw - anchor) equals that count as
//                long as W_t - W_{t-1} == rint(raw_t - raw_{t-1}) for every frame, starting from the engine's
//                carried count (first chunk) or from rint(raw - anchor) of the frame before the chunk -- which is
//                what the CTA of the previous chunk ends with.  Two consecutive frames INSIDE the validity tube
//                (|u - anchor| <= delta <= 0.12 each) satisfy the identity and the step bound automatically:
//                raw_t - raw_{t-1} = (u_t - u_{t-1}) + (W_t - W_{t-1}) with |u_t - u_{t-1}| <= 0.24.  Every other
//                pair (first frame of a chunk, a frame outside the tube or right after one) is checked explicitly.
//                A failed check raises the `invalid` counter and the host redoes the batch with the scan kernels
//                (only reachable outside guard G1).
//   consumers    (C warps) per frame and group of 32 atoms: lane a tests the atom's most recent site and, when that
//                certainly fails, its previous one -- the reference's first two tests (site_collection.py:147-152,
//                atom.py:181-192), as far as both are certain inside the chunk; a certain hit is the reference's
//                answer (first hit in priority order, polyhedral_site_collection.py:104-107).  Every other verdict
//                makes the WARP search that atom's cell list, 32 candidates at a time: none -> None, one -> that
//                site, 2..4 -> spilled list for the priority resolver, anything float32 cannot decide -> deferred
//                to the exact float64 pass (k_polyhedral_deferred).
// Frames outside the validity tube (and the legacy first-frame rule) take the exhaustive generic path per atom.
#pragma once
#include "k_polyhedral_tile.cuh"

#ifndef SAB_STREAM_ROWS
#define SAB_STREAM_ROWS 6         // ring of float rows (producers run at most this far ahead of the slowest consumer)
#endif
#ifndef SAB_STREAM_MAXRAW
#define SAB_STREAM_MAXRAW 4       // ring of raw frames (the copy engine runs raw slots - 2 frames ahead)
#endif
#ifndef SAB_STREAM_MAXTHREADS
#define SAB_STREAM_MAXTHREADS 320
#endif
#ifndef SAB_STREAM_MINBLOCKS
#define SAB_STREAM_MINBLOCKS 2
#endif
#ifndef SAB_STREAM_KMAX
#define SAB_STREAM_KMAX 5         // register-resident producer form: framework / mobile columns owned by one producer thread
#endif
#define SAB_CNT_INVALID 10        // counters[]: the anchor formulation of the wrap count failed its chain check

struct SabStreamTables {
    int n_used;                   // staged (framework coordinate, slot shift) pairs
    int stride;                   // floats per row: n_used + 3 A, rounded up to a multiple of 4
    int m3;                       // framework coordinates
    const int4 *vars;             // [n_used] {3 * atom + d, framework column j, float bits of the shift, -}
    const uint16_t *rec;          // [S][12] byte offsets into a row, vertices in slot order
    int G;
    const int32_t *cell_off;      // [G^3 + 1]
    const uint16_t *cell_sites;   // CSR entries, ascending per cell
    float delta_f;
    int use_f32;
    long long *defer;
    long long defer_cap;
    const double *anchor;         // [m3] anchors of raw - W
    const int32_t *fw_atoms;      // [M] structure index of framework atom m
    const double *prev_raw;       // [m3] carry: raw coordinates of the frame before this batch
    const int32_t *wraps;         // [m3] carry: wrap counts at that frame
    int skip_check_frame0;
    int n_raw;                    // raw slots in use (2..SAB_STREAM_MAXRAW)
    int n_prod;                   // producer warps
    int use_tma;                  // frames are 16-byte aligned and sized
    double *chunk_exc;            // out [n_chunks][m3][2]: excursion (min, max) of u per chunk
    int always_explicit;          // validity tube wider than 0.12: the in-tube shortcut of the chain check does not apply
    int reg_cols;                 // every producer thread owns <= SAB_STREAM_KMAX columns: their metadata and excursions live in registers
    int n_peer;                   // multi-GPU: copies of every result row stored straight into the other ranks' buffers
    int32_t *peer[SAB_MAX_PEERS]; // (NVLink peer memory; slot of THIS rank's frames inside the peer's full result array)
};

__device__ __forceinline__ void sab_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
#ifndef SAB_STREAM_SUSPEND_NS
#define SAB_STREAM_SUSPEND_NS 4000   // try_wait suspends the warp in hardware for up to this long: a waiting warp issues nothing
#endif
__device__ __forceinline__ void sab_mbar_wait_backoff(uint64_t *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    for (;;) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(SAB_STREAM_SUSPEND_NS) : "memory");
        if (ok) break;
    }
}
template <bool SMEM> __device__ __forceinline__ SabTetRec sab_tet_rec_at(const uint16_t *__restrict__ rec)
{
    const uint2 *r2 = reinterpret_cast<const uint2 *>(rec);
    SabTetRec r;
    if (SMEM) { r.q0 = r2[0]; r.q1 = r2[1]; r.q2 = r2[2]; }
    else { r.q0 = __ldg(r2); r.q1 = __ldg(r2 + 1); r.q2 = __ldg(r2 + 2); }
    return r;
}
__device__ __forceinline__ void sab_bar_producers(int n_threads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

template <bool REC_SMEM>
__global__ void __launch_bounds__(SAB_STREAM_MAXTHREADS, SAB_STREAM_MINBLOCKS)
k_polyhedral_assign_stream(const double *__restrict__ frac, int64_t n_frames, int64_t chunk, int n_atoms, int n_mobile,
                           const int32_t *__restrict__ mobile, int n_sites, SabPolyTables T, SabStreamTables C,
                           int64_t legacy_init_frame, int32_t *__restrict__ result, int32_t *__restrict__ spill,
                           long long spill_capacity, unsigned long long *__restrict__ counters,
                           unsigned long long *__restrict__ n_fallback)
{
    extern __shared__ __align__(128) unsigned char sm_stream[];
    const int n3 = 3 * n_atoms, m3 = C.m3, a3 = 3 * n_mobile;
    double *rawd = reinterpret_cast<double *>(sm_stream);                        // [n_raw][n3] raw frames (copy engine)
    double *ancd = rawd + (size_t)C.n_raw * n3;                                   // [m3] anchors
    double *exc = ancd + m3;                                                      // [m3][2] running (min, max) of u
    float *rows = reinterpret_cast<float *>(exc + 2 * (size_t)m3);               // [ROWS][stride]
    float *shf = rows + (size_t)SAB_STREAM_ROWS * C.stride;                      // [n_used] slot shift of every staged pair
    int2 *colvar = reinterpret_cast<int2 *>(shf + ((C.n_used + 1) & ~1));         // [m3] {first, end} staged pair of the column
    int *fwcol = reinterpret_cast<int *>(colvar + m3);                            // [m3] 3 * atom + d
    int *mobcol = fwcol + m3;                                                     // [3 A] 3 * atom + d of the mobile coordinates
    int2 *last = reinterpret_cast<int2 *>(fwcol + ((m3 + a3 + 1) & ~1));          // [A] {most recent, previous} site, -1 = unknown
    uint16_t *srec = reinterpret_cast<uint16_t *>(last + n_mobile);               // [S][12] site records (REC_SMEM)
    __shared__ __align__(8) uint64_t raw_full[SAB_STREAM_MAXRAW], row_full[SAB_STREAM_ROWS], row_empty[SAB_STREAM_ROWS];
    __shared__ int pbad[4], rowflag[SAB_STREAM_ROWS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_prod_thr = C.n_prod * 32, n_cons = (blockDim.x >> 5) - C.n_prod, n_cons_thr = n_cons * 32;
    const int RS = C.n_raw;
    const int64_t n_chunks = (n_frames + chunk - 1) / chunk;

    // ---- one-time setup (all threads) ----------------------------------------------------------------------
    for (int j = tid; j < m3; j += blockDim.x) {
        ancd[j] = C.anchor[j];
        fwcol[j] = C.fw_atoms[j / 3] * 3 + j % 3;
        colvar[j] = make_int2(0, 0);
    }
    for (int q = tid; q < a3; q += blockDim.x) mobcol[q] = mobile[q / 3] * 3 + q % 3;
    __syncthreads();
    for (int v = tid; v < C.n_used; v += blockDim.x) {       // staged pairs are sorted by column
        const int4 e = C.vars[v];
        shf[v] = __int_as_float(e.z);
        if (v == 0 || C.vars[v - 1].y != e.y) colvar[e.y].x = v;
        if (v == C.n_used - 1 || C.vars[v + 1].y != e.y) colvar[e.y].y = v + 1;
    }
    if (REC_SMEM)
        for (int i = tid; i < n_sites * 3; i += blockDim.x)
            reinterpret_cast<uint2 *>(srec)[i] = reinterpret_cast<const uint2 *>(C.rec)[i];
    const uint16_t *recs = REC_SMEM ? srec : C.rec;
    if (tid == 0) {
        for (int i = 0; i < SAB_STREAM_MAXRAW; i++) sab_mbar_init(&raw_full[i], 1);
        for (int i = 0; i < SAB_STREAM_ROWS; i++) { sab_mbar_init(&row_full[i], n_prod_thr); sab_mbar_init(&row_empty[i], n_cons_thr); rowflag[i] = 0; }
        for (int i = 0; i < 4; i++) pbad[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < C.n_prod) {
        // =========================================== producers ===========================================
        const int pt = tid;                                     // 0 .. n_prod_thr - 1
        auto issue = [&](int64_t f, unsigned it) {              // thread 0 only: raw frame f -> slot it % RS
            const unsigned bytes = (unsigned)((size_t)n3 * 8);
            const int slot = (int)(it % (unsigned)RS);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            sab_mbar_expect_tx(&raw_full[slot], bytes);
            sab_bulk_g2s(rawd + (size_t)slot * n3, frac + (size_t)f * n3, bytes, &raw_full[slot]);
        };
        unsigned it = 0;                                        // frames processed by this CTA so far
        // register-resident form: thread pt owns framework columns pt + k n_prod_thr and the same mobile coordinates
        // Branch-free: a column that does not exist reads raw[0] and writes the spare float at the end of the row (its
        // anchor is NaN, so it never raises the tube flag); a column with one slot shift writes its second value there too.
        int r_fcol[SAB_STREAM_KMAX], r_v0[SAB_STREAM_KMAX], r_v1[SAB_STREAM_KMAX], r_mcol[SAB_STREAM_KMAX], r_mdst[SAB_STREAM_KMAX];
        double r_anc[SAB_STREAM_KMAX], r_lo[SAB_STREAM_KMAX], r_hi[SAB_STREAM_KMAX];
        float r_ancf[SAB_STREAM_KMAX], r_s0[SAB_STREAM_KMAX], r_s1[SAB_STREAM_KMAX];
        const int spare = C.stride - 1;                         // >= n_used + 3 A by construction (plan_stream)
        unsigned r_valid = 0, r_extra = 0;                      // bit k: column exists / has more than two slot shifts
#pragma unroll
        for (int k = 0; k < SAB_STREAM_KMAX; k++) {
            const int j = pt + k * n_prod_thr;
            const bool ok = C.reg_cols && j < m3, okm = C.reg_cols && j < a3;
            const int2 cv = ok ? colvar[j] : make_int2(0, 0);
            const int nv = cv.y - cv.x;
            r_fcol[k] = ok ? fwcol[j] : 0;
            r_v0[k] = nv > 0 ? cv.x : spare;
            r_v1[k] = nv > 1 ? cv.x + 1 : spare;
            r_anc[k] = ok ? ancd[j] : 0.0;
            r_ancf[k] = ok ? (float)r_anc[k] : __int_as_float(0x7fc00000);
            r_s0[k] = nv > 0 ? shf[cv.x] : 0.f;
            r_s1[k] = nv > 1 ? shf[cv.x + 1] : 0.f;
            r_mcol[k] = okm ? mobcol[j] : 0;
            r_mdst[k] = okm ? C.n_used + j : spare;
            if (ok) r_valid |= 1u << k;
            if (nv > 2) r_extra |= 1u << k;
            r_lo[k] = INFINITY; r_hi[k] = -INFINITY;
        }
        for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
            const int64_t c0 = ch * chunk, c1 = min(c0 + chunk, n_frames);
            sab_bar_producers(n_prod_thr);                      // previous chunk fully converted: raw slots idle
            if (pt == 0 && C.use_tma)
                for (int i = 0; i < RS - 1 && c0 + i < c1; i++) issue(c0 + i, it + i);
            for (int j = pt; j < m3; j += n_prod_thr) { exc[2 * j] = INFINITY; exc[2 * j + 1] = -INFINITY; }   // owner-private
#pragma unroll
            for (int k = 0; k < SAB_STREAM_KMAX; k++) { r_lo[k] = INFINITY; r_hi[k] = -INFINITY; }
            for (int64_t f = c0; f < c1; f++, it++) {
                const int rslot = (int)(it % (unsigned)RS), wslot = (int)(it % SAB_STREAM_ROWS);
                double *raw = rawd + (size_t)rslot * n3;
                if (C.use_tma) sab_mbar_wait_backoff(&raw_full[rslot], (it / (unsigned)RS) & 1);
                else {                                          // frames the copy engine cannot take: plain loads
                    const double *g = frac + (size_t)f * n3;
                    for (int q = pt; q < n3; q += n_prod_thr) raw[q] = g[q];
                    sab_bar_producers(n_prod_thr);
                }
                if (it >= SAB_STREAM_ROWS) sab_mbar_wait_backoff(&row_empty[wslot], ((it / SAB_STREAM_ROWS) - 1) & 1);
                float *row = rows + (size_t)wslot * C.stride;
                // ---- framework columns -> row entries ------------------------------------------------------
                int bad = 0;
                if (C.reg_cols) {
                    float fbase[SAB_STREAM_KMAX];
#pragma unroll
                    for (int k = 0; k < SAB_STREAM_KMAX; k++) {
                        const double r = raw[r_fcol[k]];
                        const double u = r - rint(r - r_anc[k]);
                        fbase[k] = (float)u;
                        r_lo[k] = u < r_lo[k] ? u : r_lo[k];
                        r_hi[k] = u > r_hi[k] ? u : r_hi[k];
                    }
#pragma unroll
                    for (int k = 0; k < SAB_STREAM_KMAX; k++) {
                        bad |= fabsf(fbase[k] - r_ancf[k]) > C.delta_f;      // false for the NaN anchor of a missing column
                        row[r_v0[k]] = fbase[k] + r_s0[k];
                        row[r_v1[k]] = fbase[k] + r_s1[k];
                    }
                    if (r_extra) {                                           // rare: a coordinate used with > 2 slot shifts
#pragma unroll
                        for (int k = 0; k < SAB_STREAM_KMAX; k++)
                            if (r_extra >> k & 1u) {
                                const int2 cv = colvar[pt + k * n_prod_thr];
                                for (int v = cv.x + 2; v < cv.y; v++) row[v] = fbase[k] + shf[v];
                            }
                    }
#pragma unroll
                    for (int k = 0; k < SAB_STREAM_KMAX; k++) {
                        const double x = raw[r_mcol[k]];
                        row[r_mdst[k]] = (float)(x - floor(x));                  // atom.py:104
                    }
                } else {
                for (int j = pt; j < m3; j += n_prod_thr) {
                    const double r = raw[fwcol[j]], anc = ancd[j];
                    const double u = r - rint(r - anc);
                    const float base = (float)u;
                    if (fabsf(base - (float)anc) > C.delta_f) bad = 1;
                    if (u < exc[2 * j]) exc[2 * j] = u;
                    if (u > exc[2 * j + 1]) exc[2 * j + 1] = u;
                    const int2 cv = colvar[j];
                    for (int v = cv.x; v < cv.y; v++) row[v] = base + shf[v];
                }
                for (int q = pt; q < a3; q += n_prod_thr) {
                    const double x = raw[mobcol[q]];
                    row[C.n_used + q] = (float)(x - floor(x));                               // atom.py:104
                }
                }
                if (bad) atomicOr(&pbad[it & 3], 1);
                sab_bar_producers(n_prod_thr);                  // pbad complete; everybody is done with the frame before
                const int tube_bad = pbad[it & 3];
                const bool explicit_check = tube_bad || f == c0 || pbad[(it + 3) & 3] || C.always_explicit;
                if (pt == 0) {
                    pbad[(it + 2) & 3] = 0;
                    const int flag = tube_bad || (T.legacy_init && legacy_init_frame == f);
                    rowflag[wslot] = flag;
                    if (flag) atomicAdd(n_fallback, 1ULL);
                    if (!explicit_check && C.use_tma && f + RS - 1 < c1) issue(f + RS - 1, it + RS - 1);   // into the slot of frame f - 1
                }
                if (explicit_check) {                           // chain check against the reference's sequential update
                    const double *praw = rawd + (size_t)((it + (unsigned)RS - 1) % (unsigned)RS) * n3;        // frame f - 1 (f > c0)
                    for (int j = pt; j < m3; j += n_prod_thr) {
                        const double r = raw[fwcol[j]], anc = ancd[j];
                        double p; int Wp;
                        if (f > c0) { p = praw[fwcol[j]]; Wp = (int)rint(p - anc); }
                        else if (c0 == 0) { p = C.prev_raw[j]; Wp = C.wraps[j]; }
                        else { p = frac[(size_t)(c0 - 1) * n3 + fwcol[j]]; Wp = (int)rint(p - anc); }
                        const double diff = r - p, w = rint(diff), phys = diff - w;          // pbc_utils.py:210-216
                        if ((phys >= 0.3 || phys <= -0.3) && !(C.skip_check_frame0 && f == 0))
                            atomicMin(&counters[3], (unsigned long long)f);
                        if ((int)rint(r - anc) - Wp != (int)w) counters[SAB_CNT_INVALID] = 1ULL;
                    }
                    if (C.use_tma) {
                        sab_bar_producers(n_prod_thr);          // frame f - 1 no longer needed
                        if (pt == 0 && f + RS - 1 < c1) issue(f + RS - 1, it + RS - 1);
                    }
                }
                sab_mbar_arrive(&row_full[wslot]);
            }
            if (C.reg_cols) {
#pragma unroll
                for (int k = 0; k < SAB_STREAM_KMAX; k++)
                    if (r_valid >> k & 1u) {
                        const int j = pt + k * n_prod_thr;
                        C.chunk_exc[((size_t)ch * m3 + j) * 2] = r_lo[k];
                        C.chunk_exc[((size_t)ch * m3 + j) * 2 + 1] = r_hi[k];
                    }
            } else
            for (int j = pt; j < m3; j += n_prod_thr) {
                C.chunk_exc[((size_t)ch * m3 + j) * 2] = exc[2 * j];
                C.chunk_exc[((size_t)ch * m3 + j) * 2 + 1] = exc[2 * j + 1];
            }
        }
    } else {
        // =========================================== consumers ===========================================
        const int cw = warp - C.n_prod;
        const unsigned FULL = 0xffffffffu;
        const bool one_group = n_mobile <= n_cons_thr;
        unsigned it = 0;
        for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
            const int64_t c0 = ch * chunk, c1 = min(c0 + chunk, n_frames);
            for (int g0 = cw * 32; g0 < n_mobile; g0 += n_cons_thr) if (g0 + lane < n_mobile) last[g0 + lane] = make_int2(-1, -1);   // owner-private
            int2 rc_reg = make_int2(-1, -1);                    // the same history in a register when a lane owns one atom
            for (int64_t f = c0; f < c1; f++, it++) {
                const int wslot = (int)(it % SAB_STREAM_ROWS);
                sab_mbar_wait_backoff(&row_full[wslot], (it / SAB_STREAM_ROWS) & 1);
                const float *row = rows + (size_t)wslot * C.stride;
                const float *xrow = row + C.n_used;
                const int flag = rowflag[wslot];
                int32_t *res = result + (size_t)f * n_mobile;
                for (int g0 = cw * 32; g0 < n_mobile; g0 += n_cons_thr) {
                    const int a = g0 + lane;
                    const bool valid = a < n_mobile;
                    int2 rc = one_group ? rc_reg : (valid ? last[a] : make_int2(-1, -1));   // {most recent, previous} site
                    int out = SAB_NONE;
                    // a definite site h: Atom.update_recent_site (atom.py:189-192)
                    #define SAB_RECENT_PUSH(h) do { if ((h) != rc.x) { rc.y = rc.x; rc.x = (h); } } while (0)
                    if (flag) {                                  // outside the validity tube: generic path
                        if (valid) {
                            int c = 0;
                            out = sab_exhaustive_atom(frac + (size_t)f * n3, nullptr, f, legacy_init_frame, mobile[a], n_sites, T,
                                                      spill, spill_capacity, counters, &c, C.anchor);
                            if (c == 1) SAB_RECENT_PUSH(out); else if (c > 1) rc = make_int2(-1, -1);
                        }
                    } else {
                        int v = 0;
                        if (valid && rc.x >= 0 && C.use_f32)
                            v = sab_tet_filter32b(row, sab_tet_rec_at<REC_SMEM>(recs + (size_t)rc.x * 12), xrow + 3 * a);
                        if (v == 1) out = rc.x;
                        // most recent site certainly failed: the reference tests the previous one next
                        const bool second = valid && v == 0 && rc.x >= 0 && rc.y >= 0;
                        if (__any_sync(FULL, second)) {
                            if (second && sab_tet_filter32b(row, sab_tet_rec_at<REC_SMEM>(recs + (size_t)rc.y * 12), xrow + 3 * a) == 1) {
                                out = rc.y; rc = make_int2(rc.y, rc.x); v = 1;
                            }
                        }
                        unsigned m = __ballot_sync(FULL, valid && v != 1);
                        while (m) {                              // warp-uniform: one search per atom that needs it
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const float *xs = xrow + 3 * (g0 + src);
                            const int cell = sab_cell_of_f(xs, C.G);
                            const int b = __ldg(C.cell_off + cell), len = __ldg(C.cell_off + cell + 1) - b;
                            int c = 0, h0 = -1, h1 = -1, h2 = -1, h3 = -1;
                            unsigned und = C.use_f32 ? 0u : 1u;
                            if (C.use_f32)
                                for (int i0 = 0; i0 < len; i0 += 32) {
                                    const int i = i0 + lane;
                                    int s = 0, vv = 0;
                                    if (i < len) {
                                        s = __ldg(C.cell_sites + b + i);
                                        vv = sab_tet_filter32b(row, sab_tet_rec_at<REC_SMEM>(recs + (size_t)s * 12), xs);
                                    }
                                    unsigned m1 = __ballot_sync(FULL, vv == 1);
                                    und |= __ballot_sync(FULL, vv == 2);
                                    while (m1 && c < SAB_POLY_KEEP) {
                                        const int l = __ffs(m1) - 1;
                                        m1 &= m1 - 1;
                                        const int sv = __shfl_sync(FULL, s, l);
                                        if (c == 0) h0 = sv; else if (c == 1) h1 = sv; else if (c == 2) h2 = sv; else h3 = sv;
                                        c++;
                                    }
                                    c += __popc(m1);
                                }
                            if (lane == src) {
                                if (und || c > SAB_POLY_KEEP) {  // float32 could not decide (or a long list): exact pass later
                                    out = SAB_DEFERRED; rc = make_int2(-1, -1);
                                    const unsigned long long di = atomicAdd(&counters[9], 1ULL);
                                    if ((long long)di < C.defer_cap) C.defer[di] = (long long)f * n_mobile + a;
                                } else if (c == 0) out = SAB_NONE;
                                else if (c == 1) { out = h0; SAB_RECENT_PUSH(h0); }
                                else {
                                    int h[SAB_POLY_KEEP] = {h0, h1, h2, h3};
                                    out = sab_tile_spill(c, h, spill, spill_capacity, counters);
                                    rc = make_int2(-1, -1);
                                }
                            }
                        }
                    }
                    #undef SAB_RECENT_PUSH
                    if (valid) {
                        res[a] = out;
                        // fused result all-gather: the same row goes to every peer's copy of the full array (posted NVLink stores)
                        if (C.n_peer) {
                            const size_t o = (size_t)f * n_mobile + a;
#pragma unroll
                            for (int q = 0; q < SAB_MAX_PEERS; q++) if (q < C.n_peer) C.peer[q][o] = out;   // static indices: the parameter stays in constant memory
                        }
                    }
                    if (one_group) rc_reg = rc; else if (valid) last[a] = rc;
                }
                sab_mbar_arrive(&row_empty[wslot]);
            }
        }
    }
}

// Fused-path commit: advance the carry to the end of the batch from the per-chunk excursions the stream kernel
// wrote.  A CTA of 32 x SAB_COMMIT_GROUPS threads owns 32 framework coordinates; group g folds chunks g, g + G, ...
// (coalesced across the 32 coordinates), shared memory combines the groups.  Skipped by the kernel itself when the
// batch must be redone: spill overflow, an invalidating step, or a failed chain check.
#define SAB_COMMIT_GROUPS 8
__global__ void __launch_bounds__(32 * SAB_COMMIT_GROUPS)
k_wrap_commit_stream(const double *__restrict__ frac, int64_t n_frames, int n_atoms,
                     const int32_t *__restrict__ fw_atoms, int m3, int n_chunks,
                     const double *__restrict__ chunk_exc, const double *__restrict__ anchor,
                     double *__restrict__ prev_raw, int32_t *__restrict__ wraps,
                     double *__restrict__ excursion, unsigned long long *__restrict__ counters)
{
    __shared__ double s_lo[SAB_COMMIT_GROUPS][32], s_hi[SAB_COMMIT_GROUPS][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    if (n_frames <= 0) return;
    if (counters[2] != 0ULL || counters[3] != ~0ULL || counters[SAB_CNT_INVALID] != 0ULL) return;   // uniform over the grid
    double lo = INFINITY, hi = -INFINITY;
    if (col < m3) {
#pragma unroll 4
        for (int c = grp; c < n_chunks; c += SAB_COMMIT_GROUPS) {
            const double2 e2 = *reinterpret_cast<const double2 *>(chunk_exc + ((size_t)c * m3 + col) * 2);
            lo = fmin(lo, e2.x);
            hi = fmax(hi, e2.y);
        }
    }
    s_lo[grp][lane] = lo; s_hi[grp][lane] = hi;
    __syncthreads();
    if (grp != 0 || col >= m3) return;
    lo = excursion[2 * col]; hi = excursion[2 * col + 1];
#pragma unroll
    for (int g = 0; g < SAB_COMMIT_GROUPS; g++) { lo = fmin(lo, s_lo[g][lane]); hi = fmax(hi, s_hi[g][lane]); }
    excursion[2 * col] = lo; excursion[2 * col + 1] = hi;
    atomicMax(&counters[8], (unsigned long long)__double_as_longlong(fmax(hi - lo, 0.0)));        // guard G1, read by the host
    const double r = frac[((size_t)(n_frames - 1) * n_atoms + fw_atoms[col / 3]) * 3 + col % 3];
    wraps[col] = (int)rint(r - anchor[col]);
    prev_raw[col] = r;
}
