// k_fold.cuh -- folds over a FINISHED result array int32[F][A].
//
// The reference keeps per-object Python lists (Site.trajectory: per frame the list of atoms in the site,
// trajectory.py:357-360) and derives everything else by walking them: residence times (site.py:241-325), average
// occupation (site.py:225-239), transition Counters (site_collection.py:300-302).  All of these are functions of the
// maximal constant stretches ("runs") of each atom's column of the result array and of the frame-to-frame changes,
// so ONE pass over the array (HBM-bound: 4 bytes per ion-frame read, a few bytes per run written) gives
//   runs   (site, atom, start, length) sorted by (site, atom, start): the run-length form of every Site.trajectory
//   pairs  (from, to, count, first (frame * A + atom)): the increments of Site.transitions with their first
//          occurrence, i.e. the insertion order of the reference's Counters.
// Integer work only, bit-exact by construction.
//
//   k_fold_scan     one thread per (frame, atom): emits one 64-bit key per run start and one per run end,
//                   key = (((site * A + atom) * F + frame) << 1) | is_end   (one reservation per 2048-entry tile),
//                   and inserts (previous site -> site) changes into an open-addressing hash table.
//   cub::DeviceRadixSort::SortKeys over the significant bits: starts and ends of one (site, atom) alternate in time,
//                   so after the sort key 2r is the start and key 2r + 1 the end of run r.
//   k_fold_decode   unpacks the pairs of keys into the run arrays (and verifies the pairing).
//   k_fold_compact  gathers the occupied hash slots.
#pragma once
#include <cstdint>
#include <cub/device/device_radix_sort.cuh>

#define SAB_FOLD_EMPTY 0xffffffffffffffffULL
// flags[]: 0 = keys emitted, 1 = pairs in the table, 2 = error bits, 3 = compacted pairs
#define SAB_FOLD_ERR_UNRESOLVED 1ULL   // an entry <= SAB_AMBIGUOUS_BASE (unresolved) or >= n_sites
#define SAB_FOLD_ERR_TABLE 2ULL        // hash table full
#define SAB_FOLD_ERR_PAIRING 4ULL      // start / end keys do not pair up (cannot happen; checked anyway)

__device__ __forceinline__ unsigned long long sab_fold_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

#define SAB_FOLD_PER_THREAD 8
#define SAB_FOLD_TILE (256 * SAB_FOLD_PER_THREAD)

// One CTA per tile of 2048 consecutive entries (grid-stride).  Thread t owns entries t, t + 256, ... of the tile
// (coalesced); all 24 loads of a thread (entry, row before, row after) are issued before the first use; (frame, atom)
// advance incrementally (no division per entry); the CTA reserves room for all of the tile's keys with ONE atomic
// (block-wide exclusive scan of the per-thread key counts); the rare transition inserts run after the loop, all
// threads that have one at the same time.
__device__ __forceinline__ void sab_fold_insert(unsigned long long key, unsigned long long i, unsigned long long *t_key,
                                                unsigned long long *t_cnt, unsigned long long *t_first,
                                                unsigned long long slot_mask, unsigned long long *flags)
{
    unsigned long long slot = sab_fold_hash(key) & slot_mask;
    for (unsigned long long probe = 0; probe <= slot_mask; probe++) {
        const unsigned long long old = atomicCAS(&t_key[slot], SAB_FOLD_EMPTY, key);
        if (old == SAB_FOLD_EMPTY) atomicAdd(&flags[1], 1ULL);
        if (old == SAB_FOLD_EMPTY || old == key) {
            atomicAdd(&t_cnt[slot], 1ULL);
            atomicMin(&t_first[slot], i);
            return;
        }
        slot = (slot + 1) & slot_mask;
    }
    atomicOr(&flags[2], SAB_FOLD_ERR_TABLE);
}

__global__ void __launch_bounds__(256)
k_fold_scan(const int32_t *__restrict__ result, long long n_frames, int n_mobile, int n_sites,
            const int32_t *__restrict__ prev, unsigned long long *__restrict__ keys, long long key_capacity,
            unsigned long long *__restrict__ t_key, unsigned long long *__restrict__ t_cnt,
            unsigned long long *__restrict__ t_first, unsigned long long slot_mask,
            unsigned long long *__restrict__ flags)
{
    __shared__ unsigned s_warp[8];
    __shared__ unsigned long long s_base;
    const long long total = n_frames * (long long)n_mobile;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n_tiles = (total + SAB_FOLD_TILE - 1) / SAB_FOLD_TILE;
    const unsigned A = (unsigned)n_mobile, step_q = 256u / A, step_r = 256u % A;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long t0 = tile * SAB_FOLD_TILE;
        const long long f0 = t0 / n_mobile;                       // uniform: one 64-bit division per tile
        const unsigned off0 = (unsigned)(t0 - f0 * n_mobile) + (unsigned)tid;
        // 32-bit offsets from the tile's base pointer; the neighbour rows are predicated loads (no clamped 64-bit indices)
        const int32_t *__restrict__ tp = result + t0;
        const long long rem = total - t0;                         // entries from the tile start to the end of the array
        const int n_in = rem < SAB_FOLD_TILE ? (int)rem : SAB_FOLD_TILE;
        const int n_after = rem - n_mobile < SAB_FOLD_TILE ? (int)(rem - n_mobile) : SAB_FOLD_TILE;   // offsets with a row below (may be <= 0)
        const int first_with_before = t0 >= n_mobile ? 0 : (int)(n_mobile - t0);                        // offsets with a row above
        int cur[SAB_FOLD_PER_THREAD], before[SAB_FOLD_PER_THREAD], after[SAB_FOLD_PER_THREAD];
#pragma unroll
        for (int k = 0; k < SAB_FOLD_PER_THREAD; k++) {           // all loads first
            const int o = k * 256 + tid;
            cur[k] = o < n_in ? tp[o] : -1;
            before[k] = (o < n_in && o >= first_with_before) ? tp[o - n_mobile] : -2;
            after[k] = o < n_after ? tp[o + n_mobile] : -3;
        }
        unsigned q = off0 / A, r = off0 - q * A;                  // frame f0 + q, atom r of entry k = 0
        unsigned sbits = 0, ebits = 0, tbits = 0, bad = 0;
#pragma unroll
        for (int k = 0; k < SAB_FOLD_PER_THREAD; k++) {
            const int o = k * 256 + tid;
            if (o < n_in) {
                const int c = cur[k];
                int b = before[k];
                if (o < first_with_before && prev) b = prev[r];  // frame 0: the atoms' previous trajectory entries
                if (c < -1 || c >= n_sites) bad = 1;
                else if (c >= 0) {
                    if (o < first_with_before || b != c) sbits |= 1u << k;
                    if (after[k] != c) ebits |= 1u << k;           // -3 on the last frame: always an end
                    if (b >= 0 && b != c) tbits |= 1u << k;
                }
            }
            r += step_r; q += step_q;
            if (r >= A) { r -= A; q++; }
        }
        if (bad) atomicOr(&flags[2], SAB_FOLD_ERR_UNRESOLVED);
        // Events are rare (a few per cent of the entries): each thread walks only ITS set bits, so a warp spends
        // max-over-lanes(events per thread) rounds on them instead of one round per entry slot.  The entry is re-read
        // (L1 hit) and its (frame, atom) recomputed, which keeps the per-entry state out of dynamically indexed arrays.
        // site_collection.py:300-302: previous trajectory entry is a site and differs from the new one
        if (t_key)
            for (unsigned bits = tbits; bits; bits &= bits - 1) {
                const int k = __ffs(bits) - 1;
                const long long i = t0 + k * 256 + tid;
                const unsigned off = off0 + 256u * k;
                const int c = result[i];
                const int b = i >= n_mobile ? result[i - n_mobile] : prev[off % A];
                sab_fold_insert((unsigned long long)b * (unsigned long long)n_sites + c, (unsigned long long)i, t_key, t_cnt,
                                t_first, slot_mask, flags);
            }
        // block-wide exclusive scan of the key counts
        const unsigned n = __popc(sbits) + __popc(ebits);
        unsigned incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (tid == 0) {
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const unsigned v = s_warp[w]; s_warp[w] = run; run += v; }
            s_base = run ? atomicAdd(&flags[0], (unsigned long long)run) : 0ULL;
        }
        __syncthreads();
        unsigned long long o = s_base + s_warp[warp] + (incl - n);
        for (unsigned bits = sbits | ebits; bits; bits &= bits - 1) {
            const int k = __ffs(bits) - 1;
            const long long i = t0 + k * 256 + tid;
            const unsigned off = off0 + 256u * k;
            const unsigned dq = off / A;
            const unsigned long long body =
                (((unsigned long long)result[i] * A + (off - dq * A)) * (unsigned long long)n_frames + (unsigned long long)(f0 + dq)) << 1;
            if (sbits >> k & 1u) { if ((long long)o < key_capacity) keys[o] = body; o++; }
            if (ebits >> k & 1u) { if ((long long)o < key_capacity) keys[o] = body | 1ULL; o++; }
        }
        __syncthreads();                                          // s_warp / s_base are reused by the next tile
    }
}

__global__ void __launch_bounds__(256)
k_fold_decode(const unsigned long long *__restrict__ keys, long long n_runs, long long n_frames, int n_mobile,
              int32_t *__restrict__ run_site, int32_t *__restrict__ run_atom, int64_t *__restrict__ run_start,
              int64_t *__restrict__ run_length, unsigned long long *__restrict__ flags)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 k2 = reinterpret_cast<const ulonglong2 *>(keys)[r];
        const unsigned long long ks = k2.x >> 1, ke = k2.y >> 1;
        const unsigned long long fs = ks % (unsigned long long)n_frames, ps = ks / (unsigned long long)n_frames;
        const unsigned long long fe = ke % (unsigned long long)n_frames, pe = ke / (unsigned long long)n_frames;
        if ((k2.x & 1ULL) != 0ULL || (k2.y & 1ULL) != 1ULL || ps != pe || fe < fs) atomicOr(&flags[2], SAB_FOLD_ERR_PAIRING);
        run_site[r] = (int32_t)(ps / (unsigned long long)n_mobile);
        run_atom[r] = (int32_t)(ps % (unsigned long long)n_mobile);
        run_start[r] = (int64_t)fs;
        run_length[r] = (int64_t)(fe - fs + 1);
    }
}

__global__ void __launch_bounds__(256)
k_fold_compact(const unsigned long long *__restrict__ t_key, const unsigned long long *__restrict__ t_cnt,
               const unsigned long long *__restrict__ t_first, unsigned long long n_slots, int n_sites,
               long long pair_capacity, int32_t *__restrict__ p_src, int32_t *__restrict__ p_dst,
               int64_t *__restrict__ p_count, int64_t *__restrict__ p_first, unsigned long long *__restrict__ flags)
{
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = t_key[s];
        if (k == SAB_FOLD_EMPTY) continue;
        const unsigned long long o = atomicAdd(&flags[3], 1ULL);
        if ((long long)o >= pair_capacity) continue;
        p_src[o] = (int32_t)(k / (unsigned long long)n_sites);
        p_dst[o] = (int32_t)(k % (unsigned long long)n_sites);
        p_count[o] = (int64_t)t_cnt[s];
        p_first[o] = (int64_t)t_first[s];
    }
}

// Per atom, the site of its last run (key = (end frame << 24) | site, largest wins) and, in a second pass, of its last
// run in a DIFFERENT site: the two entries of Atom._recent_sites after the batch (atom.py:181-192), as far as the
// batch alone determines them.  0 = the atom was never assigned.
__global__ void __launch_bounds__(256)
k_fold_recent(const int32_t *__restrict__ run_site, const int32_t *__restrict__ run_atom,
              const int64_t *__restrict__ run_start, const int64_t *__restrict__ run_length, long long n_runs,
              unsigned long long *__restrict__ last, unsigned long long *__restrict__ other, int pass)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (long long)gridDim.x * blockDim.x) {
        const int site = run_site[r], atom = run_atom[r];
        const unsigned long long key = ((unsigned long long)(run_start[r] + run_length[r]) << 24) | (unsigned long long)site;
        if (pass == 0) atomicMax(&last[atom], key);
        else if (site != (int)(last[atom] & 0xffffffULL)) atomicMax(&other[atom], key);
    }
}
