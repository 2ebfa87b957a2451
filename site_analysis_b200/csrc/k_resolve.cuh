// k_resolve.cuh -- the priority resolver: settles atom-frames contained by several sites
// exactly as the reference's first-hit-in-priority-order loop would.
//
// Replays, in frame order and (within a frame) atom list order:
//   PriorityAssignmentMixin._get_priority_sites   site_collection.py:113-182
//   _NearestSiteLookup.nearest_site_index          site_collection.py:36-50
//   Site.most_frequent_transitions                  site.py:216-223 (count desc, stable)
//   SiteCollection.update_occupation                site_collection.py:279-310
//   Atom.update_recent_site                         atom.py:181-192
// Given the candidate set C of an atom-frame (all containing sites) the reference's answer is
//   recent[0] if in C, else recent[1] if in C, else (anchor = recent[0], or with no history the
//   fractional-space nearest centre, itself tested first) the anchor's learned destination with
//   the highest count that is in C (ties: Counter insertion order), else the member of C that
//   comes first in the anchor's distance-ranked row (or face-sharing neighbours in list order,
//   then list order, when no ranking exists); with no anchor at all: lowest list position.
//
// One CTA; every thread owns atoms a = tid, tid + T, ...  Decisions that need only per-atom
// state (recent sites) are taken in parallel.  Atoms that need the shared transition
// counters (to read them, or to add a transition) are handled afterwards by warp 0 in
// ascending atom order, so the counters evolve exactly as in the sequential reference.
#pragma once
#include "sab_device.cuh"

struct SabPriorityTables {
    const int32_t *rank;        // [S][S-1] dense distance-ranked rows, or null
    const int32_t *rank_pos;    // [S][S] dense inverse: position of site j in anchor i's row (built at upload), or null
    const double *lookup;       // [S][3] or null
    const int32_t *neigh_off;   // [S+1] or null
    const int32_t *neigh;
    // lazy ranking (large S): no dense table.  The rank of a candidate is its fractional minimum-image distance to the
    // anchor, computed here with numpy's own operation order (site_collection.py:100-107); only when two candidates TIE
    // does the order depend on the sort implementation -- then the anchor's true row (computed by the host with the
    // reference's np.argsort call) is needed: cached inverse rows, and a list of anchors still missing.
    int ranked;                 // a distance ranking exists (reference centres known for every site)
    const int32_t *row_slot;    // [S] slot of the anchor's cached inverse row, -1 = not cached; null with dense tables
    const int32_t *row_pos;     // [slots][S] inverse rows (position of every site; the anchor itself: INT_MAX)
    const int32_t *row_fwd;     // [slots][S] the rows themselves (S - 1 entries used)
    int32_t *missing;           // anchors whose row was needed (the decision taken without it is provisional)
    int *n_missing;
    int missing_cap;
};

// Sort key of candidate c in anchor's ranked row: the exact position when the row is known, else the distance itself.
__device__ __forceinline__ double sab_rank_key(const SabPriorityTables &P, int anchor, int c, int n_sites, bool &exact)
{
    if (P.rank_pos) { exact = true; return (double)P.rank_pos[(size_t)anchor * n_sites + c]; }
    const int slot = P.row_slot ? P.row_slot[anchor] : -1;
    if (slot >= 0) { exact = true; return (double)P.row_pos[(size_t)slot * n_sites + c]; }
    exact = false;
    double d0 = P.lookup[3 * c] - P.lookup[3 * anchor], d1 = P.lookup[3 * c + 1] - P.lookup[3 * anchor + 1], d2 = P.lookup[3 * c + 2] - P.lookup[3 * anchor + 2];
    d0 -= rint(d0); d1 -= rint(d1); d2 -= rint(d2);             // delta -= np.round(delta)
    return sqrt((d0 * d0 + d1 * d1) + d2 * d2);                  // np.linalg.norm(delta, axis=1)
}
__device__ __forceinline__ void sab_rank_missing(const SabPriorityTables &P, int anchor)
{
    if (!P.missing) return;
    const int i = atomicAdd(P.n_missing, 1);
    if (i < P.missing_cap) P.missing[i] = anchor;
}
// The member of c[0..n) that comes first in anchor's ranked row (serial; ties without the row: lowest position, reported).
__device__ __forceinline__ int sab_rank_first(const SabPriorityTables &P, int anchor, const int32_t *c, int n, int n_sites)
{
    double bk = INFINITY; int bs = -1; bool tie = false, exact = true;
    for (int i = 0; i < n; i++) {
        bool ex;
        const double k = sab_rank_key(P, anchor, c[i], n_sites, ex);
        exact = ex;
        if (k < bk) { bk = k; bs = c[i]; tie = false; }
        else if (k == bk) { tie = true; if (c[i] < bs) bs = c[i]; }
    }
    if (tie && !exact) sab_rank_missing(P, anchor);
    return bs;
}

struct SabHistory {
    int32_t *recent;            // [A][2]
    int32_t *prev;              // [A]  -2 empty trajectory, -1 None
    int32_t *tr_n;              // [S]
    int32_t *tr_key;            // [S][K] insertion order
    long long *tr_cnt;          // [S][K]
    int K;
};

__device__ __forceinline__ bool sab_in_list(const int32_t *c, int n, int v)
{
    for (int i = 0; i < n; i++) if (c[i] == v) return true;
    return false;
}

__device__ __forceinline__ void sab_touch_recent(int32_t *recent, int a, int s)
{
    if (s != recent[2 * a]) { recent[2 * a + 1] = recent[2 * a]; recent[2 * a] = s; }
}

// executed by ONE warp; lane-parallel where it matters.  Returns the chosen site.
__device__ int sab_resolve_with_counters(const int32_t *c, int n, int r0, const double *x, int n_sites,
                                         const SabPriorityTables &P, const SabHistory &H, int lane)
{
    int anchor = r0;
    if (anchor < 0 && P.lookup) {          // site_collection.py:153-156
        double bd = INFINITY; int bi = 0;
        for (int s = lane; s < n_sites; s += 32) {
            double d0 = P.lookup[3 * s] - x[0], d1 = P.lookup[3 * s + 1] - x[1], d2 = P.lookup[3 * s + 2] - x[2];
            d0 -= rint(d0); d1 -= rint(d1); d2 -= rint(d2);
            double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            if (nrm < bd) { bd = nrm; bi = s; }
        }
        for (int o = 16; o; o >>= 1) {
            double od = __shfl_xor_sync(0xffffffffu, bd, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        anchor = bi;
        if (sab_in_list(c, n, anchor)) return anchor;
    }
    if (anchor < 0) {                       // :180-182 plain list order
        int best = c[0];
        for (int i = 1; i < n; i++) best = min(best, c[i]);
        return best;
    }
    // learned transitions of the anchor: highest count first, ties by insertion order
    {
        int nk = H.tr_n[anchor];
        const int32_t *key = H.tr_key + (size_t)anchor * H.K;
        const long long *cn = H.tr_cnt + (size_t)anchor * H.K;
        long long bc = -1; int bi = -1;
        for (int i = 0; i < nk; i++)
            if (cn[i] > bc && sab_in_list(c, n, key[i])) { bc = cn[i]; bi = key[i]; }
        if (bi >= 0) return bi;
    }
    if (P.ranked) {                         // site_collection.py:167-171: first member of C in the anchor's ranked row
        const int bests = sab_rank_first(P, anchor, c, n, n_sites);
        if (bests >= 0) return bests;
    }
    if (P.neigh_off && !P.ranked) {         // :172-176 face-sharing neighbours in list order
        for (int i = P.neigh_off[anchor]; i < P.neigh_off[anchor + 1]; i++)
            if (sab_in_list(c, n, P.neigh[i])) return P.neigh[i];
    }
    int best = c[0];                        // :177-179 remaining sites in list order
    for (int i = 1; i < n; i++) best = min(best, c[i]);
    return best;
}

__global__ void k_resolve(const double *__restrict__ frac, int64_t n_frames, int n_atoms, int n_mobile,
                          const int32_t *__restrict__ mobile, int n_sites, int32_t *__restrict__ result,
                          const int32_t *__restrict__ spill, SabPriorityTables P, SabHistory H,
                          int *__restrict__ overflow)
{
    extern __shared__ int sm_flag[];        // [A] 0 = done, 1 = needs ordered handling
    __shared__ int any_flag;
    const int lane = threadIdx.x & 31;
    for (int64_t f = 0; f < n_frames; f++) {
        if (threadIdx.x == 0) any_flag = 0;
        __syncthreads();
        for (int a = threadIdx.x; a < n_mobile; a += blockDim.x) {
            int r = result[(size_t)f * n_mobile + a];
            int cur = r, flag = 0;
            if (r <= SAB_AMBIGUOUS_BASE) {
                const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - r);
                int n = c[0]; c++;
                int r0 = H.recent[2 * a], r1 = H.recent[2 * a + 1];
                if (r0 >= 0 && sab_in_list(c, n, r0)) cur = r0;
                else if (r1 >= 0 && sab_in_list(c, n, r1)) cur = r1;
                else flag = 1;
            }
            if (!flag) {
                int pv = H.prev[a];
                if (cur >= 0 && pv >= 0 && pv != cur) flag = 1;        // adds a transition: ordered
                else {
                    if (cur >= 0) sab_touch_recent(H.recent, a, cur);
                    H.prev[a] = cur;
                    result[(size_t)f * n_mobile + a] = cur;
                }
            }
            if (flag) { result[(size_t)f * n_mobile + a] = (r <= SAB_AMBIGUOUS_BASE && cur >= 0) ? cur : r; any_flag = 1; }
            sm_flag[a] = flag;
        }
        __syncthreads();
        if (any_flag && threadIdx.x < 32) {
            for (int a = 0; a < n_mobile; a++) {
                if (!sm_flag[a]) continue;
                int r = result[(size_t)f * n_mobile + a];
                int cur = r;
                if (r <= SAB_AMBIGUOUS_BASE) {
                    const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - r);
                    int n = c[0]; c++;
                    double x[3];
                    for (int d = 0; d < 3; d++)
                        x[d] = sab_pymod1(frac[((size_t)f * n_atoms + mobile[a]) * 3 + d]);
                    cur = sab_resolve_with_counters(c, n, H.recent[2 * a], x, n_sites, P, H, lane);
                }
                __syncwarp();
                if (lane == 0) {
                    int pv = H.prev[a];
                    if (cur >= 0 && pv >= 0 && pv != cur) {            // site_collection.py:300-306
                        int nk = H.tr_n[pv], i = 0;
                        int32_t *key = H.tr_key + (size_t)pv * H.K;
                        long long *cn = H.tr_cnt + (size_t)pv * H.K;
                        while (i < nk && key[i] != cur) i++;
                        if (i < nk) cn[i]++;
                        else if (nk < H.K) { key[nk] = cur; cn[nk] = 1; H.tr_n[pv] = nk + 1; }
                        else *overflow = 1;
                    }
                    if (cur >= 0) sab_touch_recent(H.recent, a, cur);
                    H.prev[a] = cur;
                    result[(size_t)f * n_mobile + a] = cur;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
}
