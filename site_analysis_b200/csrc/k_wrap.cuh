// k_wrap.cuh -- PBC wrap bookkeeping as a prefix scan over frames.
//
// The reference keeps, per polyhedron vertex / dynamic-Voronoi reference slot, an integer
// image shift that is decremented by rint(raw_t - raw_prev) whenever the coordinate jumps
// by ~1 (pbc_utils.py:193-230 `_numba_update_pbc_shifts`; dynamic_voronoi_site_collection.py:
// 91-97).  That state depends only on the ATOM, so here it is a per-framework-atom
// cumulative wrap count W[f][m][d] = W_carry + sum_{t<=f} rint(raw_t - raw_{t-1}); the shift
// of slot (s, v) is slot_shift[s][v] - W[f][atom(s, v)].
//
// Three small kernels (chunked scan): per-chunk totals -> exclusive scan of chunk totals
// -> per-frame rows.  They also track the running excursion of raw - W (the validity guard
// of the scan formulation) and the first frame whose step violates |phys| < 0.3
// (pbc_utils.py:216 / dynamic_voronoi_site_collection.py:94).
#pragma once
#include "sab_device.cuh"

#define SAB_WRAP_CHUNK 128

__device__ __forceinline__ double sab_fw_raw(const double *frac, int64_t f, int n_atoms,
                                             const int32_t *fw_atoms, int col)
{
    int m = col / 3, d = col - 3 * m;
    return frac[((size_t)f * n_atoms + fw_atoms[m]) * 3 + d];
}

// grid (n_chunks, ceil(M3/128)), block 128
__global__ void k_wrap_totals(const double *__restrict__ frac, int64_t n_frames, int n_atoms,
                              const int32_t *__restrict__ fw_atoms, int m3,
                              const double *__restrict__ prev_raw, int32_t *__restrict__ totals,
                              unsigned long long *__restrict__ first_bad, int skip_check_frame0)
{
    int col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= m3) return;
    int64_t f0 = (int64_t)blockIdx.x * SAB_WRAP_CHUNK;
    int64_t f1 = min(f0 + (int64_t)SAB_WRAP_CHUNK, n_frames);
    double prev = (f0 == 0) ? prev_raw[col] : sab_fw_raw(frac, f0 - 1, n_atoms, fw_atoms, col);
    int tot = 0;
    for (int64_t f = f0; f < f1; f++) {
        double raw = sab_fw_raw(frac, f, n_atoms, fw_atoms, col);
        double diff = raw - prev;
        double w = rint(diff);
        double phys = diff - w;
        if ((phys >= 0.3 || phys <= -0.3) && !(skip_check_frame0 && f == 0))
            atomicMin(first_bad, (unsigned long long)f);
        tot += (int)w;
        prev = raw;
    }
    totals[(size_t)blockIdx.x * m3 + col] = tot;
}

// 1 thread per column; exclusive scan of the chunk totals starting at the engine's carry.  The
// loads are issued 16 at a time (independent) so the walk is not a chain of L2 round trips.
__global__ void k_wrap_bases(int32_t *__restrict__ totals, int n_chunks, int m3,
                             const int32_t *__restrict__ wraps_carry)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m3) return;
    int run = wraps_carry[col];
    for (int c0 = 0; c0 < n_chunks; c0 += 16) {
        int t[16];
#pragma unroll
        for (int k = 0; k < 16; k++) t[k] = (c0 + k < n_chunks) ? totals[(size_t)(c0 + k) * m3 + col] : 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (c0 + k < n_chunks) totals[(size_t)(c0 + k) * m3 + col] = run;
            run += t[k];
        }
    }
}

// Same grid as k_wrap_totals.  rows[f][col] = inclusive cumulative wrap count at frame f.
__global__ void k_wrap_rows(const double *__restrict__ frac, int64_t n_frames, int n_atoms,
                            const int32_t *__restrict__ fw_atoms, int m3,
                            const double *__restrict__ prev_raw, const int32_t *__restrict__ bases,
                            int32_t *__restrict__ rows, double *__restrict__ chunk_exc)
{
    int col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= m3) return;
    int64_t f0 = (int64_t)blockIdx.x * SAB_WRAP_CHUNK;
    int64_t f1 = min(f0 + (int64_t)SAB_WRAP_CHUNK, n_frames);
    double prev = (f0 == 0) ? prev_raw[col] : sab_fw_raw(frac, f0 - 1, n_atoms, fw_atoms, col);
    int W = bases[(size_t)blockIdx.x * m3 + col];
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t f = f0; f < f1; f++) {
        double raw = sab_fw_raw(frac, f, n_atoms, fw_atoms, col);
        W += (int)rint(raw - prev);
        rows[(size_t)f * m3 + col] = W;
        double u = raw - (double)W;
        lo = fmin(lo, u); hi = fmax(hi, u);
        prev = raw;
    }
    chunk_exc[((size_t)blockIdx.x * m3 + col) * 2] = lo;
    chunk_exc[((size_t)blockIdx.x * m3 + col) * 2 + 1] = hi;
}

// 1 thread per column: advance the carry (wraps, prev_raw, excursion) to the end of frame
// n_commit - 1.  Whole batch: fold the per-chunk excursions; partial batch (rare: the batch is
// split at a cache-invalidating step): walk the committed frames.
__global__ void k_wrap_commit(const double *__restrict__ frac, int64_t n_frames, int64_t n_commit, int n_atoms,
                              const int32_t *__restrict__ fw_atoms, int m3, int n_chunks,
                              const int32_t *__restrict__ rows, const double *__restrict__ chunk_exc,
                              double *__restrict__ prev_raw, int32_t *__restrict__ wraps,
                              double *__restrict__ excursion, const unsigned long long *__restrict__ veto,
                              unsigned long long *__restrict__ max_span)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m3 || n_commit <= 0) return;
    // automatic commit of a whole batch is skipped when the batch must be re-run or split:
    // veto[0] = spill overflow flag, veto[1] = first frame with an invalidating step (~0 = none)
    if (veto && (veto[0] != 0ULL || veto[1] != ~0ULL)) return;
    double lo = excursion[2 * col], hi = excursion[2 * col + 1];
    if (n_commit == n_frames) {
#pragma unroll 8
        for (int c = 0; c < n_chunks; c++) {
            const double2 e2 = *reinterpret_cast<const double2 *>(chunk_exc + ((size_t)c * m3 + col) * 2);
            lo = fmin(lo, e2.x);
            hi = fmax(hi, e2.y);
        }
    } else {
        for (int64_t f = 0; f < n_commit; f++) {
            double u = sab_fw_raw(frac, f, n_atoms, fw_atoms, col) - (double)rows[(size_t)f * m3 + col];
            lo = fmin(lo, u); hi = fmax(hi, u);
        }
    }
    excursion[2 * col] = lo; excursion[2 * col + 1] = hi;
    if (max_span) atomicMax(max_span, (unsigned long long)__double_as_longlong(fmax(hi - lo, 0.0)));   // guard G1, read by the host
    wraps[col] = rows[(size_t)(n_commit - 1) * m3 + col];
    prev_raw[col] = sab_fw_raw(frac, n_commit - 1, n_atoms, fw_atoms, col);
}

// 1 thread per column: total wraps over all chunks and the raw coordinate of the last frame
// (the halo the next GPU shard needs), packed for one small D2H copy.
__global__ void k_wrap_reduce(const double *__restrict__ frac, int64_t n_frames, int n_atoms,
                              const int32_t *__restrict__ fw_atoms, int m3, int n_chunks,
                              const int32_t *__restrict__ totals, int32_t *__restrict__ out_tot,
                              double *__restrict__ out_last)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m3) return;
    int t = 0;
    for (int c = 0; c < n_chunks; c++) t += totals[(size_t)c * m3 + col];
    out_tot[col] = t;
    out_last[col] = sab_fw_raw(frac, n_frames - 1, n_atoms, fw_atoms, col);
}

// Shard carry without a scan (multi-GPU, polyhedral): 1 thread per framework coordinate.  `halo` is the raw frame
// before the shard, anchor0 the unwrapped coordinates u = raw - W of the trajectory's FIRST frame.  While guard G1
// holds over the whole trajectory (every u within 0.3 of every other, checked by the caller from the gathered
// excursions) |u(halo) - anchor0| < 0.5, hence W(halo) = rint(halo - anchor0) exactly -- the sum of the wrap
// totals of all earlier shards, without exchanging them.
__global__ void k_carry_from_halo(const double *__restrict__ halo, const double *__restrict__ anchor0, int m3,
                                  double *__restrict__ prev_raw, int32_t *__restrict__ wraps, double *__restrict__ excursion)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m3) return;
    const double r = halo[col];
    const int W = (int)rint(r - anchor0[col]);
    const double u = r - (double)W;
    prev_raw[col] = r; wraps[col] = W;
    excursion[2 * col] = u; excursion[2 * col + 1] = u;
}
