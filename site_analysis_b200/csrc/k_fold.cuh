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
//                   key = (((site * A + atom) * F + frame) << 1) | is_end   (warp-aggregated append),
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

__global__ void __launch_bounds__(256)
k_fold_scan(const int32_t *__restrict__ result, long long n_frames, int n_mobile, int n_sites,
            const int32_t *__restrict__ prev, unsigned long long *__restrict__ keys, long long key_capacity,
            unsigned long long *__restrict__ t_key, unsigned long long *__restrict__ t_cnt,
            unsigned long long *__restrict__ t_first, unsigned long long slot_mask,
            unsigned long long *__restrict__ flags)
{
    const long long total = n_frames * (long long)n_mobile;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // every warp walks whole 32-entry groups so that the ballots below are warp-uniform
    const long long n_groups = (total + 31) / 32;
    for (long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < n_groups;
         g += (long long)gridDim.x * (blockDim.x >> 5)) {
        const long long i = g * 32 + lane;
        const bool in = i < total;
        int cur = -1, before = -2, after = -3;
        long long f = 0; int a = 0;
        if (in) {
            f = i / n_mobile; a = (int)(i - f * n_mobile);
            cur = result[i];
            before = f > 0 ? result[i - n_mobile] : (prev ? prev[a] : -2);
            after = f + 1 < n_frames ? result[i + n_mobile] : -3;
            if (cur < -1 || cur >= n_sites) { atomicOr(&flags[2], SAB_FOLD_ERR_UNRESOLVED); cur = -1; }
        }
        const bool occupied = in && cur >= 0;
        const bool is_start = occupied && (f == 0 || before != cur);
        const bool is_end = occupied && after != cur;
        const unsigned ms = __ballot_sync(FULL, is_start), me = __ballot_sync(FULL, is_end);
        const int n_s = __popc(ms), n_e = __popc(me);
        if (n_s + n_e) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&flags[0], (unsigned long long)(n_s + n_e));
            base = __shfl_sync(FULL, base, 0);
            const unsigned long long body = (((unsigned long long)cur * n_mobile + a) * (unsigned long long)n_frames + f) << 1;
            const unsigned below = (1u << lane) - 1u;
            if (is_start) {
                const unsigned long long o = base + __popc(ms & below);
                if ((long long)o < key_capacity) keys[o] = body;
            }
            if (is_end) {
                const unsigned long long o = base + n_s + __popc(me & below);
                if ((long long)o < key_capacity) keys[o] = body | 1ULL;
            }
        }
        // site_collection.py:300-302: previous trajectory entry is a site and differs from the new one
        if (occupied && before >= 0 && before != cur && t_key) {
            const unsigned long long k = (unsigned long long)before * (unsigned long long)n_sites + cur;
            unsigned long long slot = sab_fold_hash(k) & slot_mask;
            bool done = false;
            for (unsigned long long probe = 0; probe <= slot_mask; probe++) {
                const unsigned long long old = atomicCAS(&t_key[slot], SAB_FOLD_EMPTY, k);
                if (old == SAB_FOLD_EMPTY) atomicAdd(&flags[1], 1ULL);
                if (old == SAB_FOLD_EMPTY || old == k) {
                    atomicAdd(&t_cnt[slot], 1ULL);
                    atomicMin(&t_first[slot], (unsigned long long)i);
                    done = true;
                    break;
                }
                slot = (slot + 1) & slot_mask;
            }
            if (!done) atomicOr(&flags[2], SAB_FOLD_ERR_TABLE);
        }
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
