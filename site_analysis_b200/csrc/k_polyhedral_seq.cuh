// k_polyhedral_seq.cuh -- tetrahedral containment, chunk-sequential: "previous site first".
//
// The reference tests, for every atom, the site it occupied most recently FIRST
// (PriorityAssignmentMixin._get_priority_sites, site_collection.py:147-152) and stops at the
// first hit (polyhedral_site_collection.py:104-107).  So whenever the most recent site still
// contains the atom, that site IS the reference's answer -- no other site needs to be looked
// at.  In MD data that is ~97% of all atom-frames.
//
// Layout: one CTA walks a chunk of consecutive frames; thread a owns mobile atom a and carries
// `last` = the atom's most recent site *as far as it is certain inside this chunk* (unknown at
// the chunk start and after an atom-frame that several sites contain).
//   phase 1  every thread tests `last` with the exact single-image tetrahedron test
//            (k_polyhedral_cells.cuh); hit -> done.
//   phase 2  the atoms that missed (or have no certain history) are expanded into
//            (atom, candidate) pairs from the static cell list and spread over ALL threads of
//            the CTA, so a few slow atoms do not stall their warps; hits are counted per atom.
//   phase 3  owners read the verdict: none -> SAB_NONE, one -> that site (and `last` becomes
//            certain again), several -> spilled list for the priority resolver, `last` unknown.
// Frames outside the cell lists' validity tube take the exhaustive generic path, as in the
// frame-parallel kernel.  Results are bit-identical to k_polyhedral_assign_exhaustive + resolver.
#pragma once
#include "k_polyhedral_cells.cuh"

#define SAB_SEQ_THREADS 256
#ifndef SAB_SEQ_MINBLOCKS
#define SAB_SEQ_MINBLOCKS 3
#endif
#define SAB_SEQ_MAX_OWN 16       // atoms per thread: A <= 4096 (shared-memory budget permitting)

__device__ __forceinline__ int sab_cell_of(const double *x, int G)
{
    int c0 = min(G - 1, (int)(x[0] * (double)G)), c1 = min(G - 1, (int)(x[1] * (double)G)),
        c2 = min(G - 1, (int)(x[2] * (double)G));
    return (c0 * G + c1) * G + c2;
}

// One atom against every site with the fully general test (any polyhedron, all 8 images):
// the path of frames outside the cell lists' validity tube.  Returns the result code.
__device__ __noinline__ int sab_exhaustive_atom(const double *__restrict__ frame, const int32_t *__restrict__ wrow,
                                                int64_t f, int64_t legacy_init_frame, int atom, int n_sites,
                                                const SabPolyTables &T, int32_t *__restrict__ spill,
                                                long long spill_capacity, unsigned long long *__restrict__ counters,
                                                int *count, const double *__restrict__ anchor = nullptr)
{
    double x[3];
    for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)atom * 3 + d]);
    int c = 0, first = -1;
    unsigned long long off = 0;
    for (int pass = 0; pass < 2; pass++) {
        int k = 0;
        for (int s = 0; s < n_sites; s++) {
            double vc[SAB_MAX_VERT][3], lo[3], hi[3], cen[3];
            int nv;
            bool legacy_now = T.legacy_init && T.legacy_init[s] && f == legacy_init_frame;
            sab_poly_vertices(frame, T, s, wrow, legacy_now, vc, nv, lo, hi, cen, anchor);
            if (sab_poly_contains_atom(vc, cen, T.simp + (size_t)s * T.fmax * 3, T.n_face[s], lo, hi, x)) {
                if (pass == 0) { if (c == 0) first = s; c++; }
                else spill[off + 1 + k++] = s;
            }
        }
        if (pass == 1) return SAB_AMBIGUOUS_BASE - (int)off;
        *count = c;
        if (c == 0) return SAB_NONE;
        if (c == 1) return first;
        atomicAdd(&counters[0], 1ULL);
        off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(c));
        if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) { counters[2] = 1ULL; return SAB_NONE; }
        spill[off] = c;
    }
    return SAB_NONE;
}

// rare paths call the test out of line to keep the hot loop small
__device__ __noinline__ bool sab_tet_contains_cold(const double *var, const uint16_t *rec, const double *x)
{
    return sab_tet_contains(var, rec, x);
}

#define SAB_SEQ_STAGE 4          // framework coordinates staged per thread (3M <= 4 * blockDim)

// Stage one frame: var[k][i] = raw + (double)(kmin + k - W) for every framework coordinate i
// (pbc_utils.py:220: frac_coords + new_shifts, ONE rounding), and test the validity tube.
__device__ __forceinline__ int sab_stage_write(double *__restrict__ var, const SabTetTables &C, int i, double raw, int W)
{
    for (int k = 0; k < C.nk; k++) var[k * C.m3 + i] = raw + (double)(C.kmin + k - W);
    return fabs((raw - (double)W) - C.anchor[i]) > C.delta;
}

__device__ __forceinline__ void sab_cp_async8(void *smem_dst, const void *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sab_cp_async4(void *smem_dst, const void *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sab_cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

// Shared-memory layout (byte offsets from the dynamic base), one atom per thread (A <= blockDim):
//   var[nk][3M] f64 | raw[3M] f64 | xraw[A][3] f64 | xs[A][3] f64 | rawW[3M] i32 | miss_atom/miss_b/miss_len/cnt[A] i32 | hits[A][KEEP] i32
__global__ void __launch_bounds__(SAB_SEQ_THREADS, SAB_SEQ_MINBLOCKS)
k_polyhedral_assign_seq(const double *__restrict__ frac, int64_t n_frames, int chunk, int n_atoms, int n_mobile,
                        const int32_t *__restrict__ mobile, int n_sites, SabPolyTables T, SabTetTables C,
                        const int32_t *__restrict__ wrows, int64_t legacy_init_frame,
                        int32_t *__restrict__ result, int32_t *__restrict__ spill, long long spill_capacity,
                        unsigned long long *__restrict__ counters, unsigned long long *__restrict__ n_fallback)
{
    extern __shared__ double sm_seq[];
    const int nvar = C.nk * C.m3;
    double *const var = sm_seq;
    double *const raw = var + nvar;
    double *const xraw = raw + C.m3;
    double *const xs = xraw + 3 * n_mobile;
    int *const rawW = reinterpret_cast<int *>(xs + 3 * n_mobile);
    int *const miss_atom = rawW + C.m3;
    int *const miss_b = miss_atom + n_mobile;
    int *const miss_len = miss_b + n_mobile;
    int *const cnt = miss_len + n_mobile;
    int *const hits = cnt + n_mobile;
    __shared__ int n_miss;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const bool owner = tid < n_mobile;
    const int my_atom = owner ? mobile[tid] : 0;
    const int64_t n_chunks = (n_frames + chunk - 1) / chunk;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const int64_t f0 = ch * chunk, f1 = min(f0 + (int64_t)chunk, n_frames);
        int la = -1;                                             // certain most-recent site of my atom, or -1
        int prev_nm = 0;
        __syncthreads();                                         // previous chunk fully consumed
        if (tid == 0) n_miss = 0;
        {                                                        // prologue: fetch the chunk's first frame
            const double *frame = frac + (size_t)f0 * n_atoms * 3;
            const int32_t *wrow = wrows + (size_t)f0 * C.m3;
            for (int i = tid; i < C.m3; i += nthr) {
                sab_cp_async8(raw + i, frame + (size_t)C.fw_atoms[i / 3] * 3 + (i % 3));
                sab_cp_async4(rawW + i, wrow + i);
            }
            if (owner) for (int d = 0; d < 3; d++) sab_cp_async8(xraw + 3 * tid + d, frame + (size_t)my_atom * 3 + d);
        }
        for (int64_t f = f0; f < f1; f++) {
            const double *frame = frac + (size_t)f * n_atoms * 3;
            const int32_t *wrow = wrows + (size_t)f * C.m3;
            if (prev_nm) __syncthreads();                        // phase 3 of the previous frame may still read var
            // ---- stage this frame from the prefetched raw values (each thread reads back only what it
            // copied itself, so waiting on its own cp.async group is enough) ------------------------
            sab_cp_async_wait_all();
            int bad = 0;
            for (int i = tid; i < C.m3; i += nthr) bad |= sab_stage_write(var, C, i, raw[i], rawW[i]);
            double x[3] = {0.0, 0.0, 0.0};
            if (owner) {
#pragma unroll
                for (int d = 0; d < 3; d++) { x[d] = sab_pymod1(xraw[3 * tid + d]); xs[3 * tid + d] = x[d]; }   // atom.py:104
            }
            bad = __syncthreads_or(bad);                         // B0: var + xs complete
            if (f + 1 < f1) {                                    // prefetch the next frame (lands during this frame's work)
                const double *fr2 = frame + (size_t)n_atoms * 3;
                const int32_t *w2 = wrow + C.m3;
                for (int i = tid; i < C.m3; i += nthr) {
                    sab_cp_async8(raw + i, fr2 + (size_t)C.fw_atoms[i / 3] * 3 + (i % 3));
                    sab_cp_async4(rawW + i, w2 + i);
                }
                if (owner) for (int d = 0; d < 3; d++) sab_cp_async8(xraw + 3 * tid + d, fr2 + (size_t)my_atom * 3 + d);
            }
            if (T.legacy_init && f == legacy_init_frame) bad = 1;
            if (bad) {
                // ---- exhaustive generic path for this frame (outside the validity tube) -------
                if (tid == 0) atomicAdd(n_fallback, 1ULL);
                if (owner) {
                    int c = 0;
                    int out = sab_exhaustive_atom(frame, wrow, f, legacy_init_frame, my_atom, n_sites, T, spill,
                                                  spill_capacity, counters, &c);
                    if (c == 1) la = out;
                    else if (c > 1) la = -1;
                    result[(size_t)f * n_mobile + tid] = out;
                }
                prev_nm = 0;
                continue;                                        // uniform
            }
            // ---- phase 1: the most recent site first (site_collection.py:147-152) ---------------------
            int myslot = -1;
            if (owner) {
                if (la >= 0 && sab_tet_contains(var, C.rec + (size_t)la * SAB_TET_REC, x)) {
                    result[(size_t)f * n_mobile + tid] = la;
                } else {
                    myslot = atomicAdd(&n_miss, 1);
                    const int cell = sab_cell_of(x, C.G);
                    const int b = __ldg(C.cell_off + cell);
                    miss_atom[myslot] = tid; miss_b[myslot] = b; miss_len[myslot] = __ldg(C.cell_off + cell + 1) - b; cnt[myslot] = 0;
                }
            }
            __syncthreads();                                     // B1: miss list complete
            const int nm = n_miss;
            prev_nm = nm;
            if (nm == 0) continue;                               // uniform
            // ---- phase 2: (atom, candidate) pairs spread over all threads -----------------------------
            {
                int j = 0, acc = 0;
#ifdef SAB_SEQ_STATS
                int ntests = 0;
#endif
#pragma unroll 1
                for (int p = tid;; p += nthr) {
                    while (j < nm && acc + miss_len[j] <= p) { acc += miss_len[j]; j++; }
                    if (j >= nm) break;
#ifdef SAB_SEQ_STATS
                    ntests++;
#endif
                    const int s = __ldg(C.cell_sites + miss_b[j] + (p - acc));
                    if (sab_tet_contains(var, C.rec + (size_t)s * SAB_TET_REC, xs + 3 * miss_atom[j])) {
                        int k = atomicAdd(&cnt[j], 1);
                        if (k < SAB_POLY_KEEP) hits[j * SAB_POLY_KEEP + k] = s;
                    }
                }
#ifdef SAB_SEQ_STATS
                if (ntests) atomicAdd(&counters[6], (unsigned long long)ntests);
#endif
            }
            __syncthreads();                                     // B2: verdicts complete
            if (tid == 0) { n_miss = 0; atomicAdd(&counters[5], (unsigned long long)nm); }
            // ---- phase 3: owners collect -----------------------------------------------------------------
            if (myslot >= 0) {
                const int c = cnt[myslot];
                int out;
                if (c == 0) out = SAB_NONE;                       // recent sites unchanged (atom.py:190-192)
                else if (c == 1) { out = hits[myslot * SAB_POLY_KEEP]; la = out; }
                else {
                    la = -1;                                      // several sites: the resolver decides
                    atomicAdd(&counters[0], 1ULL);
                    unsigned long long off = atomicAdd(&counters[1], (unsigned long long)SAB_SPILL_RECORD(c));
                    if ((long long)(off + c + 1) > spill_capacity || off + c + 1 >= 0x7ffffff0ULL) { counters[2] = 1ULL; out = SAB_NONE; }
                    else {
                        spill[off] = c;
                        int k = 0;                                // ascending: the cell list is sorted
                        for (int i = 0; i < miss_len[myslot]; i++) {
                            int s = C.cell_sites[miss_b[myslot] + i];
                            if (sab_tet_contains_cold(var, C.rec + (size_t)s * SAB_TET_REC, x)) spill[off + 1 + k++] = s;
                        }
                        out = SAB_AMBIGUOUS_BASE - (int)off;
                    }
                }
                result[(size_t)f * n_mobile + tid] = out;
            }
        }
    }
}
