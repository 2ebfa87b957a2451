// k_polyhedral_seq1.cuh -- the production chunk-sequential tetrahedral kernel ("previous site first"):
// persistent 256-thread CTAs pull chunks of consecutive frames from a global counter; the next frame's
// raw coordinates are prefetched with cp.async while the current one is processed.  See
// k_polyhedral_seq.cuh for the phase structure (that file keeps an experimental variant, kernel id 3).
#pragma once
#include "k_polyhedral_seq.cuh"
#define SAB_SEQ1_THREADS 256
// SAB_SEQ1_NO_FP32: compile the kernel without the FP32 pre-test (A/B and cross-check builds)
#ifdef SAB_SEQ1_NO_FP32
#define SAB_TET_TEST(var, varf, rec, x) sab_tet_contains(var, rec, x)
#else
#define SAB_TET_TEST(var, varf, rec, x) sab_tet_contains_f(var, varf, rec, x, C.use_f32)
#endif
#ifndef SAB_SEQ1_MINBLOCKS
#define SAB_SEQ1_MINBLOCKS 4      // 64 registers: 4 CTAs (32 warps) per SM measured fastest (profiles/README.md)
#endif
#ifndef SAB_SEQ1_NO_PREFETCH
#define SAB_SEQ1_PREFETCH 1
#endif
#define SAB_SEQ1_MAX_OWN 16

__global__ void __launch_bounds__(SAB_SEQ1_THREADS, SAB_SEQ1_MINBLOCKS)
k_polyhedral_assign_seq1(const double *__restrict__ frac, int64_t n_frames, int chunk, int n_atoms, int n_mobile,
                        const int32_t *__restrict__ mobile, int n_sites, SabPolyTables T, SabTetTables C,
                        const int32_t *__restrict__ wrows, int64_t legacy_init_frame,
                        int32_t *__restrict__ result, int32_t *__restrict__ spill, long long spill_capacity,
                        unsigned long long *__restrict__ counters, unsigned long long *__restrict__ n_fallback)
{
    extern __shared__ double sm_seq[];
    double *var = sm_seq;                                        // [nk][3M]
    double *xs = var + (size_t)C.nk * C.m3;                      // [A][3]
    float *varf = reinterpret_cast<float *>(xs + 3 * (size_t)n_mobile);   // [nk][3M] float copy for the FP32 pre-test
    int *miss_atom = reinterpret_cast<int *>(varf + (size_t)C.nk * C.m3 + ((C.nk * C.m3) & 1));   // [A]
    int *miss_b = miss_atom + n_mobile;                          // [A] first candidate
    int *miss_pref = miss_b + n_mobile;                          // [A+1] exclusive prefix of list lengths
    int *cnt = miss_pref + n_mobile + 1;                         // [A]
    int *hits = cnt + n_mobile;                                  // [A][KEEP]
    int *last = hits + (size_t)n_mobile * SAB_POLY_KEEP;         // [A] certain most-recent site or -1
    int *slot = last + n_mobile;                                 // [A] miss-list slot of the atom in this frame
#ifdef SAB_SEQ1_PREFETCH
    // next frame's raw data, fetched with cp.async while this frame is processed; every thread reads
    // back only the entries it copied itself, so its own wait_group is the only synchronisation
    int *rawW = slot + n_mobile;                                 // [3M]
    double *rawbuf = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(rawW + C.m3) + 7) & ~(uintptr_t)7);   // [3M], 8-byte aligned
    double *xraw = rawbuf + C.m3;                                // [A][3]
#endif
    __shared__ int n_miss;
    __shared__ long long next_chunk;
    const int tid = threadIdx.x;
    const int64_t n_chunks = (n_frames + chunk - 1) / chunk;
    for (;;) {                                                   // persistent CTAs pull chunks from a global counter
        __syncthreads();
        if (tid == 0) next_chunk = (long long)atomicAdd(&counters[7], 1ULL);
        __syncthreads();
        const int64_t ch = next_chunk;
        if (ch >= n_chunks) break;
        const int64_t f0 = ch * chunk, f1 = min(f0 + (int64_t)chunk, n_frames);
        for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS) last[a] = -1;     // owner-private: no sync needed
#ifdef SAB_SEQ1_PREFETCH
        {
            const double *fr0 = frac + (size_t)f0 * n_atoms * 3;
            const int32_t *w0 = wrows + (size_t)f0 * C.m3;
            for (int i = tid; i < C.m3; i += SAB_SEQ1_THREADS) {
                sab_cp_async8(rawbuf + i, fr0 + (size_t)C.fw_atoms[i / 3] * 3 + (i % 3));
                sab_cp_async4(rawW + i, w0 + i);
            }
            for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS)
                for (int d = 0; d < 3; d++) sab_cp_async8(xraw + 3 * a + d, fr0 + (size_t)mobile[a] * 3 + d);
        }
#endif
        for (int64_t f = f0; f < f1; f++) {
            const double *frame = frac + (size_t)f * n_atoms * 3;
            const int32_t *wrow = wrows + (size_t)f * C.m3;
            __syncthreads();                                     // previous frame fully consumed
            if (tid == 0) n_miss = 0;
            int bad = 0;
#ifdef SAB_SEQ1_PREFETCH
            sab_cp_async_wait_all();
#endif
            for (int i = tid; i < C.m3; i += SAB_SEQ1_THREADS) {
                int m = i / 3, d = i - 3 * m;
#ifdef SAB_SEQ1_PREFETCH
                double raw = rawbuf[i];
                int W = rawW[i];
                (void)m; (void)d;
#else
                double raw = frame[(size_t)C.fw_atoms[m] * 3 + d];
                int W = wrow[i];
#endif
                if (fabs((raw - (double)W) - C.anchor[i]) > C.delta) bad = 1;
                for (int k = 0; k < C.nk; k++) {
                    const double val = raw + (double)(C.kmin + k - W);       // pbc_utils.py:220, one rounding
                    var[k * C.m3 + i] = val;
                    varf[k * C.m3 + i] = (float)val;
                    if (!(fabs(val) < 3.5)) bad = 1;                         // the FP32 pre-test's error bound assumes |coordinate| < 4
                }
            }
#ifdef SAB_SEQ1_PREFETCH
            for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS)
                for (int d = 0; d < 3; d++) xs[3 * a + d] = sab_pymod1(xraw[3 * a + d]);       // atom.py:104
#endif
            bad = __syncthreads_or(bad);
#ifdef SAB_SEQ1_PREFETCH
            if (f + 1 < f1) {                                    // lands while this frame is being processed
                const double *fr2 = frame + (size_t)n_atoms * 3;
                const int32_t *w2 = wrow + C.m3;
                for (int i = tid; i < C.m3; i += SAB_SEQ1_THREADS) {
                    sab_cp_async8(rawbuf + i, fr2 + (size_t)C.fw_atoms[i / 3] * 3 + (i % 3));
                    sab_cp_async4(rawW + i, w2 + i);
                }
                for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS)
                    for (int d = 0; d < 3; d++) sab_cp_async8(xraw + 3 * a + d, fr2 + (size_t)mobile[a] * 3 + d);
            }
#endif
            if (T.legacy_init && f == legacy_init_frame) bad = 1;
            if (bad) {
                // ---- exhaustive generic path for this frame (outside the validity tube) -------
                if (tid == 0) atomicAdd(n_fallback, 1ULL);
#pragma unroll 1
                for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS) {
                    int c = 0;
                    int out = sab_exhaustive_atom(frame, wrow, f, legacy_init_frame, mobile[a], n_sites, T, spill,
                                                  spill_capacity, counters, &c);
                    if (c == 1) last[a] = out;
                    else if (c > 1) last[a] = -1;
                    result[(size_t)f * n_mobile + a] = out;
                }
                continue;
            }
            // ---- phase 1: the most recent site first ---------------------------------------------
#pragma unroll 1
            for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS) {
                double x[3];
#pragma unroll
#ifdef SAB_SEQ1_PREFETCH
                for (int d = 0; d < 3; d++) x[d] = xs[3 * a + d];
#else
                for (int d = 0; d < 3; d++) { x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]); xs[3 * a + d] = x[d]; }
#endif
                const int la = last[a];
                slot[a] = -1;
                if (la >= 0 && SAB_TET_TEST(var, varf, C.rec + (size_t)la * SAB_TET_REC, x)) {
                    result[(size_t)f * n_mobile + a] = la;
                } else {
                    int j = atomicAdd(&n_miss, 1);
                    slot[a] = j;
                    miss_atom[j] = a;
                    int cell = sab_cell_of(x, C.G);
                    int b = __ldg(C.cell_off + cell);
                    miss_b[j] = b;
                    miss_pref[j] = __ldg(C.cell_off + cell + 1) - b;     // length for now
                    cnt[j] = 0;
                }
            }
            __syncthreads();
            const int nm = n_miss;
            if (nm == 0) continue;                               // uniform
            if (tid < 32) {                                      // exclusive prefix of the list lengths (warp 0)
                int carry = 0;
                for (int base = 0; base < nm; base += 32) {
                    int j = base + tid;
                    int len = (j < nm) ? miss_pref[j] : 0, inc = len;
                    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
                    if (j < nm) miss_pref[j] = carry + inc - len;
                    carry += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (tid == 0) miss_pref[nm] = carry;
            }
            __syncthreads();
            // ---- phase 2: (atom, candidate) pairs over all threads -------------------------------
            const int total = miss_pref[nm];
            for (int p = tid; p < total; p += SAB_SEQ1_THREADS) {
                int lo_j = 0, hi_j = nm;                         // largest j with miss_pref[j] <= p
                while (hi_j - lo_j > 1) { int mid = (lo_j + hi_j) >> 1; if (miss_pref[mid] <= p) lo_j = mid; else hi_j = mid; }
                int j = lo_j;
                int s = __ldg(C.cell_sites + miss_b[j] + (p - miss_pref[j]));
                if (SAB_TET_TEST(var, varf, C.rec + (size_t)s * SAB_TET_REC, xs + 3 * miss_atom[j])) {
                    int k = atomicAdd(&cnt[j], 1);
                    if (k < SAB_POLY_KEEP) hits[j * SAB_POLY_KEEP + k] = s;
                }
            }
            __syncthreads();
            // ---- phase 3: owners collect ------------------------------------------------------------
#pragma unroll 1
            for (int a = tid; a < n_mobile; a += SAB_SEQ1_THREADS) {
                int j = slot[a];
                if (j < 0) continue;
                int c = cnt[j], out;
                if (c == 0) out = SAB_NONE;
                else if (c == 1) { out = hits[j * SAB_POLY_KEEP]; last[a] = out; }
                else {
                    last[a] = -1;                                // several sites: the resolver decides; history uncertain
                    atomicAdd(&counters[0], 1ULL);
                    unsigned long long off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(c));
                    if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) { counters[2] = 1ULL; out = SAB_NONE; }
                    else {
                        spill[off] = c;
                        int k = 0;                               // ascending list positions: rescan the (sorted) cell list
                        int b = miss_b[j], len = miss_pref[j + 1] - miss_pref[j];
                        for (int i = 0; i < len; i++) {
                            int s = C.cell_sites[b + i];
                            if (sab_tet_contains(var, C.rec + (size_t)s * SAB_TET_REC, xs + 3 * a)) spill[off + 1 + k++] = s;
                        }
                        out = SAB_AMBIGUOUS_BASE - (int)off;
                    }
                }
                result[(size_t)f * n_mobile + a] = out;
            }
        }
    }
}
