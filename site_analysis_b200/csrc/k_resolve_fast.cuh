// k_resolve_fast.cuh -- the priority resolver with all history in shared memory.
//
// Same semantics as k_resolve (k_resolve.cuh): replays PriorityAssignmentMixin._get_priority_sites
// (site_collection.py:113-182), SiteCollection.update_occupation (:279-310) and
// Atom.update_recent_site (atom.py:181-192) over the batch in frame order.  Differences:
//   * recent sites, previous trajectory entry and the per-site transition Counters (insertion-ordered
//     rows of K keys) live in shared memory;
//   * a frame needs the ORDERED section (warp 0, ascending atom order) only if (a) some atom has to
//     consult the Counters (both recent sites miss and several sites contain it), or (b) a transition
//     inserts a NEW Counter key (insertion order decides later ties, site.py:223).  In all other
//     frames -- the vast majority -- transitions only increment existing keys, which commutes, so
//     they are applied with shared-memory atomics in parallel and the frame costs one barrier;
//   * the next frame's result row is prefetched.
// A Counter row overflowing K keys sets *overflow; the host then reruns with the global-memory kernel.
#pragma once
#include "k_resolve.cuh"

struct SabSmemHistory {
    int *recent;     // [A][2]
    int *prev;       // [A]
    int *tr_n;       // [S]
    int *tr_key;     // [S][K]
    int *tr_cnt;     // [S][K]
    int K;
};

__device__ __forceinline__ int sab_row_find(const SabSmemHistory &H, int src, int dst)
{
    const int n = H.tr_n[src];
    const int *key = H.tr_key + (size_t)src * H.K;
    for (int i = 0; i < n; i++) if (key[i] == dst) return i;
    return -1;
}

// One warp; mirrors sab_resolve_with_counters with shared-memory Counters.
__device__ int sab_resolve_with_counters_smem(const int32_t *c, int n, int r0, const double *x, int n_sites,
                                              const SabPriorityTables &P, const SabSmemHistory &H, int lane)
{
    int anchor = r0;
    if (anchor < 0 && P.lookup) {
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
    if (anchor < 0) {
        int best = c[0];
        for (int i = 1; i < n; i++) best = min(best, c[i]);
        return best;
    }
    {
        const int nk = H.tr_n[anchor];
        const int *key = H.tr_key + (size_t)anchor * H.K, *cn = H.tr_cnt + (size_t)anchor * H.K;
        int bc = -1, bi = -1;
        for (int i = 0; i < nk; i++)
            if (cn[i] > bc && sab_in_list(c, n, key[i])) { bc = cn[i]; bi = key[i]; }
        if (bi >= 0) return bi;
    }
    if (P.ranked) {                         // site_collection.py:167-171: first member of C in the anchor's ranked row
        const int bests = sab_rank_first(P, anchor, c, n, n_sites);
        if (bests >= 0) return bests;
    }
    if (P.neigh_off && !P.ranked)
        for (int i = P.neigh_off[anchor]; i < P.neigh_off[anchor + 1]; i++)
            if (sab_in_list(c, n, P.neigh[i])) return P.neigh[i];
    int best = c[0];
    for (int i = 1; i < n; i++) best = min(best, c[i]);
    return best;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT)
k_resolve_fast(const double *__restrict__ frac, int64_t n_frames, int n_atoms, int n_mobile,
               const int32_t *__restrict__ mobile, int n_sites, int32_t *__restrict__ result,
               const int32_t *__restrict__ spill, SabPriorityTables P, SabHistory G /* global in/out */, int K,
               int *__restrict__ overflow)
{
    extern __shared__ int sm_r[];
    SabSmemHistory H;
    H.K = K;
    H.recent = sm_r;
    H.prev = H.recent + 2 * n_mobile;
    H.tr_n = H.prev + n_mobile;
    H.tr_key = H.tr_n + n_sites;
    H.tr_cnt = H.tr_key + (size_t)n_sites * K;
    int *cur_s = H.tr_cnt + (size_t)n_sites * K;     // [A] this frame's (tentative) result
    int *flag = cur_s + n_mobile;                    // [A] 0 done, 1 existing-key adder, 2 new-key adder, 3 reader, 4 reader (published)
    constexpr int PC = 12;                           // candidates kept in registers / published per reader
    int *rd_n = flag + n_mobile;                     // [A] published readers: list length
    int *rd_c = rd_n + n_mobile;                     // [A][PC] candidates
    double *rd_p = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(rd_c + (size_t)n_mobile * PC) + 7) & ~(uintptr_t)7);   // [A][PC] their sort keys in the anchor's ranked row (8-byte aligned)
    int *rd_x = reinterpret_cast<int *>(rd_p + (size_t)n_mobile * PC);   // [A] the keys are exact row positions (not distances)
    int *row_mark = rd_x + n_mobile;                 // [S] frame stamp: a reader of this frame consults this Counter row
    int *ord_list = row_mark + n_sites;              // [A] atoms left for the ordered section (unsorted)
    __shared__ int reader_unknown_row, ord_n;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    // load history (global rows may be wider than K: the host guarantees tr_n <= K on entry)
    for (int a = tid; a < n_mobile; a += nthr) { H.recent[2 * a] = G.recent[2 * a]; H.recent[2 * a + 1] = G.recent[2 * a + 1]; H.prev[a] = G.prev[a]; }
    for (int s = tid; s < n_sites; s += nthr) {
        int n = G.tr_n[s];
        H.tr_n[s] = n;
        for (int i = 0; i < n && i < K; i++) { H.tr_key[(size_t)s * K + i] = G.tr_key[(size_t)s * G.K + i]; H.tr_cnt[(size_t)s * K + i] = (int)G.tr_cnt[(size_t)s * G.K + i]; }
    }
    if (tid == 0) { reader_unknown_row = 0; ord_n = 0; }
    for (int s2 = tid; s2 < n_sites; s2 += nthr) row_mark[s2] = 0;
    __syncthreads();
    // One atom per thread (the usual case): result entries and candidate lists are fetched three / two
    // frames ahead into registers, so their global-memory latency is off the per-frame critical path.
    const bool one_each = n_mobile <= nthr;
    const bool mine = one_each && tid < n_mobile;
    int rq[4] = {SAB_NONE, SAB_NONE, SAB_NONE, SAB_NONE}; // result entries of frames f .. f+3
    int cn[3] = {0, 0, 0};                               // list lengths of frames f .. f+2
    int cc[3][PC];
    auto load_r = [&](int64_t f) -> int { return (mine && f < n_frames) ? result[(size_t)f * n_mobile + tid] : SAB_NONE; };
    // All PC + 1 loads are issued unconditionally (no load depends on the list length), so a warp never
    // waits inside the prefetch; entries are masked by the length when they are used.  Reading a few
    // ints past a short list stays inside the spill buffer (the host allocates >= 1024 ints of slack).
    auto load_c = [&](int r, int &n, int *c) {
        const int32_t *p = spill + ((r <= SAB_AMBIGUOUS_BASE) ? (SAB_AMBIGUOUS_BASE - r) : 0);
        n = p[0];
#pragma unroll
        for (int i = 0; i < PC; i++) c[i] = p[1 + i];
    };
    rq[0] = load_r(0); rq[1] = load_r(1); rq[2] = load_r(2); rq[3] = load_r(3);
    load_c(rq[0], cn[0], cc[0]); load_c(rq[1], cn[1], cc[1]); load_c(rq[2], cn[2], cc[2]);
    for (int64_t f = 0; f < n_frames; f++) {
        // ---- classification (no shared state is modified here except the atom's own recent/prev) ----
        int my_need = 0;                                 // this thread has a reader or a new-key adder
        for (int a = tid; a < n_mobile; a += nthr) {
            const int r = one_each ? rq[0] : result[(size_t)f * n_mobile + a];
            int cur = r, fl = 0;
            const int r0 = H.recent[2 * a], r1 = H.recent[2 * a + 1];
            if (r <= SAB_AMBIGUOUS_BASE) {
                bool in0 = false, in1 = false;
                if (one_each && cn[0] <= PC) {
#pragma unroll
                    for (int i = 0; i < PC; i++) { in0 |= i < cn[0] && cc[0][i] == r0; in1 |= i < cn[0] && cc[0][i] == r1; }
                    in0 &= r0 >= 0; in1 &= r1 >= 0;
                } else {
                    const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - r);
                    const int n = c[0]; c++;
                    in0 = r0 >= 0 && sab_in_list(c, n, r0);
                    in1 = r1 >= 0 && sab_in_list(c, n, r1);
                }
                if (in0) cur = r0;
                else if (in1) cur = r1;
                else {
                    fl = 3; my_need = 1;
                    if (!(one_each && cn[0] <= PC && r0 >= 0 && P.ranked)) reader_unknown_row = 1;
                    if (one_each && cn[0] <= PC && r0 >= 0 && P.ranked) {
                        row_mark[r0] = (int)(f + 1);
                        // anchor = recent[0] is known: publish the candidates and their rank positions now
                        // (one global load each, in parallel over all readers of the frame)
                        fl = 4;
                        rd_n[a] = cn[0];
#pragma unroll
                        bool ex = true;
                        for (int i = 0; i < PC; i++) {
                            rd_c[a * PC + i] = (i < cn[0]) ? cc[0][i] : -1;
                            rd_p[a * PC + i] = (i < cn[0]) ? sab_rank_key(P, r0, cc[0][i], n_sites, ex) : INFINITY;
                        }
                        rd_x[a] = ex ? 1 : 0;
                    }
                }
            }
            if (one_each) {                              // rotate the pipeline and fetch frame f+3's list / f+4's entry
                rq[0] = rq[1]; rq[1] = rq[2]; rq[2] = rq[3];
                cn[0] = cn[1]; cn[1] = cn[2];
#pragma unroll
                for (int i = 0; i < PC; i++) { cc[0][i] = cc[1][i]; cc[1][i] = cc[2][i]; }
                load_c(rq[2], cn[2], cc[2]);
                rq[3] = load_r(f + 4);
            }
            if (fl == 0) {
                const int pv = H.prev[a];
                if (cur >= 0 && pv >= 0 && pv != cur) {
                    fl = (sab_row_find(H, pv, cur) >= 0) ? 1 : 2;
                    if (fl == 2) my_need = 1;
                } else {
                    if (cur >= 0 && cur != r0) { H.recent[2 * a + 1] = r0; H.recent[2 * a] = cur; }
                    H.prev[a] = cur;
                    if (r <= SAB_AMBIGUOUS_BASE) result[(size_t)f * n_mobile + a] = cur;
                }
            }
            cur_s[a] = cur; flag[a] = fl;
        }
        const bool ordered = __syncthreads_or(my_need) != 0;      // B1 (uniform, race-free)
        {
            // increments of EXISTING keys commute with everything except a reader of the same Counter row in
            // this frame: apply those in parallel, leave the rest (and all readers / new keys) to the ordered walk
            const bool all_ordered = reader_unknown_row != 0;
            for (int a = tid; a < n_mobile; a += nthr) {
                const int fl = flag[a];
                if (fl == 0) continue;
                const int pv = H.prev[a];
                if (fl == 1 && !all_ordered && row_mark[pv] != (int)(f + 1)) {
                    const int cur = cur_s[a];
                    atomicAdd(&H.tr_cnt[(size_t)pv * K + sab_row_find(H, pv, cur)], 1);
                    const int r0 = H.recent[2 * a];
                    if (cur != r0) { H.recent[2 * a + 1] = r0; H.recent[2 * a] = cur; }
                    H.prev[a] = cur;
                    result[(size_t)f * n_mobile + a] = cur;
                    flag[a] = 0;
                } else {
                    ord_list[atomicAdd(&ord_n, 1)] = a;
                }
            }
        }
        if (!ordered) continue;                          // uniform: the common case costs one barrier
        __syncthreads();                                 // B2: parallel increments done, ordered list complete
        if (tid < 32) {
            const int n_ord = ord_n;
            // ascending atom order: rank the (few) listed atoms; more than 32 -> walk the flags instead
            int my_a = (lane < n_ord && n_ord <= 32) ? ord_list[lane] : INT_MAX, my_rank = 0;
            if (n_ord <= 32)
                for (int j = 0; j < n_ord; j++) { int aj = __shfl_sync(0xffffffffu, my_a, j); if (aj < my_a) my_rank++; }
            const int n_items = (n_ord <= 32) ? n_ord : n_mobile;
            for (int it = 0; it < n_items; it++) {
                int a;
                if (n_ord <= 32) {
                    unsigned who = __ballot_sync(0xffffffffu, lane < n_ord && my_rank == it);
                    a = __shfl_sync(0xffffffffu, my_a, __ffs(who) - 1);
                } else {
                    a = it;
                    if (flag[a] == 0) continue;
                }
                {
                    int cur = cur_s[a];
                    if (flag[a] == 4) {
                        // lane i <-> candidate i: learned transitions of the anchor first (highest count, ties by
                        // Counter insertion order, site.py:216-223), then the distance-ranked row
                        const int anchor = H.recent[2 * a];
                        const int n = rd_n[a];
                        const int ci = (lane < n) ? rd_c[a * PC + lane] : -1;
                        int cnt = -1, ins = INT_MAX;
                        if (ci >= 0) { int i = sab_row_find(H, anchor, ci); if (i >= 0) { cnt = H.tr_cnt[(size_t)anchor * K + i]; ins = i; } }
                        const double pos = (lane < n) ? rd_p[a * PC + lane] : INFINITY;
                        int bcnt = cnt, bins = ins, bsite = ci, psite = (lane < n) ? ci : INT_MAX;
                        double bpos = pos;
                        for (int o = 16; o; o >>= 1) {
                            int oc = __shfl_xor_sync(0xffffffffu, bcnt, o), oi = __shfl_xor_sync(0xffffffffu, bins, o), os = __shfl_xor_sync(0xffffffffu, bsite, o);
                            if (oc > bcnt || (oc == bcnt && oi < bins)) { bcnt = oc; bins = oi; bsite = os; }
                            const double op = __shfl_xor_sync(0xffffffffu, bpos, o);
                            const int ops = __shfl_xor_sync(0xffffffffu, psite, o);
                            if (op < bpos || (op == bpos && ops < psite)) { bpos = op; psite = ops; }
                        }
                        if (bcnt < 0 && !rd_x[a]) {      // distance keys: two candidates at the same distance need the anchor's true row
                            const unsigned tied = __ballot_sync(0xffffffffu, lane < n && pos == bpos);
                            if (__popc(tied) > 1 && lane == 0) sab_rank_missing(P, anchor);
                        }
                        cur = (bcnt >= 0) ? bsite : psite;
                    } else if (flag[a] == 3) {
                        const int r = result[(size_t)f * n_mobile + a];
                        const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - r);
                        const int n = c[0]; c++;
                        double x[3];
                        for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frac[((size_t)f * n_atoms + mobile[a]) * 3 + d]);
                        cur = sab_resolve_with_counters_smem(c, n, H.recent[2 * a], x, n_sites, P, H, lane);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        const int pv = H.prev[a];
                        if (cur >= 0 && pv >= 0 && pv != cur) {          // site_collection.py:300-306
                            int i = sab_row_find(H, pv, cur);
                            if (i >= 0) H.tr_cnt[(size_t)pv * K + i]++;
                            else {
                                const int nk = H.tr_n[pv];
                                if (nk < K) { H.tr_key[(size_t)pv * K + nk] = cur; H.tr_cnt[(size_t)pv * K + nk] = 1; H.tr_n[pv] = nk + 1; }
                                else *overflow = 1;
                            }
                        }
                        const int r0 = H.recent[2 * a];
                        if (cur >= 0 && cur != r0) { H.recent[2 * a + 1] = r0; H.recent[2 * a] = cur; }
                        H.prev[a] = cur;
                        result[(size_t)f * n_mobile + a] = cur;
                        flag[a] = 0;
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) { reader_unknown_row = 0; ord_n = 0; }
        }
        __syncthreads();                                 // B3
    }
    __syncthreads();
    for (int a = tid; a < n_mobile; a += nthr) { G.recent[2 * a] = H.recent[2 * a]; G.recent[2 * a + 1] = H.recent[2 * a + 1]; G.prev[a] = H.prev[a]; }
    for (int s = tid; s < n_sites; s += nthr) {
        int n = H.tr_n[s];
        G.tr_n[s] = n;
        for (int i = 0; i < n; i++) { G.tr_key[(size_t)s * G.K + i] = H.tr_key[(size_t)s * K + i]; G.tr_cnt[(size_t)s * G.K + i] = H.tr_cnt[(size_t)s * K + i]; }
    }
}
