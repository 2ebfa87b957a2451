// k_resolve_spec.cuh -- the priority resolver without a barrier per frame: speculative chunks of frames.
//
// Same semantics and results as k_resolve / k_resolve_fast (replay of PriorityAssignmentMixin._get_priority_sites
// site_collection.py:113-182, SiteCollection.update_occupation :279-310, Atom.update_recent_site atom.py:181-192,
// Site.most_frequent_transitions site.py:216-223 over the batch in frame order, atoms in list order).
//
// What makes the reference sequential is the shared state: the per-site transition Counters.  An atom READS a Counter row
// only when neither of its two recent sites contains it any more (about 1 % of the atom-frames of the overlapping-sphere
// benchmark) and then only the row of its own most recent site; it WRITES a row when it changes site (3 %).  A read and a
// write by two different atoms meet on the same row only when two atoms occupied the same site -- rare.  Hence:
//
//   phase 1  every thread walks ITS atom through the next T frames on its own, no barrier: decisions that need only the
//            atom's recent sites are final; a read takes the Counter row as it was at the start of the chunk plus the
//            atom's OWN writes of this chunk (exact if nobody else wrote that row); every write is logged, not applied;
//   phase 2  a read is confirmed when no OTHER atom wrote its row earlier in the chunk ((frame, atom) order).  The first
//            unconfirmed read (or an atom without history, which needs the warp-wide nearest-centre search, or a full
//            log) truncates the chunk at its frame t*: frames before t* are exact by construction;
//   phase 3  the accepted frames are committed: results written, increments of existing Counter keys applied with
//            shared-memory atomics (they commute), NEW keys inserted by one thread in (frame, atom) order after a bitonic
//            sort of the (usually very short) list -- insertion order breaks later ties (site.py:223);
//   exact    a truncating frame is then resolved exactly like k_resolve does (parallel classification, ordered walk of
//            the atoms that touch the Counters by warp 0), and the next chunk starts behind it.
//
// One CTA, one atom per thread (the host falls back to k_resolve_fast otherwise), all history in shared memory.  A Counter
// row overflowing K keys sets *overflow (the host widens the rows and redoes the batch, as for k_resolve_fast).
#pragma once
#include "k_resolve_fast.cuh"

#define SAB_SPEC_W 8            // own writes (site changes) an atom may log per chunk
#define SAB_SPEC_R 6            // Counter reads an atom may log per chunk
#define SAB_SPEC_PC 7           // candidates a lane keeps in registers: two 128-bit loads {n, c_0 .. c_6}; longer lists are read from memory

struct SabSpecStats { unsigned long long chunks, truncated, exact_frames, reads, new_keys, cyc_load, cyc_walk, cyc_commit, cyc_exact; };

// Choice among the candidates c[0..n) for an atom whose recent sites both miss: learned destinations of the anchor
// (Counter row at the start of the chunk + the atom's own logged writes, in order), then the distance-ranked row, then
// neighbours / list order -- the tail of sab_resolve_with_counters_smem, executed by ONE thread.
__device__ int sab_spec_decide(const int32_t *c, int n, int anchor, int n_sites, const SabPriorityTables &P, const SabSmemHistory &H,
                               const int *wl_src, const int *wl_dst, int nw)
{
    const int nk = H.tr_n[anchor];
    const int *key = H.tr_key + (size_t)anchor * H.K, *cnt = H.tr_cnt + (size_t)anchor * H.K;
    int bc = 0, bi = INT_MAX, bs = -1;
    for (int i = 0; i < n; i++) {
        const int ci = c[i];
        int k = 0, ins = INT_MAX;
        for (int j = 0; j < nk; j++) if (key[j] == ci) { k = cnt[j]; ins = j; break; }
        // own pending writes to this row, in order: count them; a key they create is inserted behind the existing ones,
        // in the order of the FIRST own write of each new destination
        int created = 0;                                 // new destinations created by own writes so far
        for (int j = 0; j < nw; j++) {
            if (wl_src[j] != anchor) continue;
            const int d = wl_dst[j];
            if (d == ci) { k++; if (ins == INT_MAX) ins = nk + created; }
            bool fresh = true;                           // does write j create a key? (not in the row, not created earlier)
            for (int q = 0; q < nk && fresh; q++) if (key[q] == d) fresh = false;
            for (int q = 0; q < j && fresh; q++) if (wl_src[q] == anchor && wl_dst[q] == d) fresh = false;
            if (fresh) created++;
        }
        if (k > bc || (k == bc && k > 0 && ins < bi)) { bc = k; bi = ins; bs = ci; }
    }
    if (bs >= 0) return bs;
    if (P.ranked) {
        const int r = sab_rank_first(P, anchor, c, n, n_sites);
        if (r >= 0) return r;
    }
    if (P.neigh_off && !P.ranked)
        for (int i = P.neigh_off[anchor]; i < P.neigh_off[anchor + 1]; i++)
            if (sab_in_list(c, n, P.neigh[i])) return P.neigh[i];
    int best = c[0];
    for (int i = 1; i < n; i++) best = min(best, c[i]);
    return best;
}

// The same choice by a whole warp: lane i holds candidate `cand` (< 0: none), n <= 32.  Every lane looks its candidate up in the
// anchor's Counter row (+ the atom's own logged writes, walked by all lanes in lockstep), then one reduction picks the
// learned destination with the highest count (ties: insertion order); without one, the candidates' rank keys are fetched
// in parallel (one load per lane) and the first of the anchor's ranked row wins (site_collection.py:167-171).
__device__ int sab_spec_decide_warp(int cand, int anchor, int n_sites, const SabPriorityTables &P, const SabSmemHistory &H,
                                    const int *wl_src, const int *wl_dst, int nw, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const int nk = H.tr_n[anchor];
    const int *key = H.tr_key + (size_t)anchor * H.K, *cnt = H.tr_cnt + (size_t)anchor * H.K;
    int k = 0, ins = INT_MAX;
    if (cand >= 0) for (int j = 0; j < nk; j++) if (key[j] == cand) { k = cnt[j]; ins = j; break; }
    int created = 0;
    for (int j = 0; j < nw; j++) {                       // uniform over the warp
        if (wl_src[j] != anchor) continue;
        const int d = wl_dst[j];
        bool fresh = true;
        for (int q = 0; q < nk && fresh; q++) if (key[q] == d) fresh = false;
        for (int q = 0; q < j && fresh; q++) if (wl_src[q] == anchor && wl_dst[q] == d) fresh = false;
        if (cand == d) { k++; if (ins == INT_MAX) ins = nk + created; }
        if (fresh) created++;
    }
    int bk = k, bi = ins, bs = cand;
    for (int o = 16; o; o >>= 1) {
        const int ok = __shfl_xor_sync(FULL, bk, o), oi = __shfl_xor_sync(FULL, bi, o), os = __shfl_xor_sync(FULL, bs, o);
        if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; bs = os; }
    }
    if (bk > 0) return bs;
    if (P.ranked) {
        bool exact = true;
        const double pos = cand >= 0 ? sab_rank_key(P, anchor, cand, n_sites, exact) : INFINITY;
        double bp = pos; int ps = cand >= 0 ? cand : INT_MAX;
        for (int o = 16; o; o >>= 1) {
            const double op = __shfl_xor_sync(FULL, bp, o);
            const int os = __shfl_xor_sync(FULL, ps, o);
            if (op < bp || (op == bp && os < ps)) { bp = op; ps = os; }
        }
        const unsigned tied = __ballot_sync(FULL, cand >= 0 && pos == bp);
        const unsigned inexact = __ballot_sync(FULL, cand >= 0 && !exact);
        if (__popc(tied) > 1 && inexact && lane == 0) sab_rank_missing(P, anchor);
        if (ps != INT_MAX) return ps;
    }
    // no ranking (sites without reference centres): face-sharing neighbours in list order, then list order
    int best = INT_MAX;
    if (P.neigh_off && !P.ranked)
        for (int i = P.neigh_off[anchor]; i < P.neigh_off[anchor + 1] && best == INT_MAX; i++)
            if (__ballot_sync(FULL, cand == P.neigh[i])) best = P.neigh[i];
    if (best != INT_MAX) return best;
    int lo = cand >= 0 ? cand : INT_MAX;
    for (int o = 16; o; o >>= 1) lo = min(lo, __shfl_xor_sync(FULL, lo, o));
    return lo;
}

// One transition prev -> cur into the shared-memory Counters (site_collection.py:300-306); single thread.
__device__ __forceinline__ void sab_spec_add(const SabSmemHistory &H, int pv, int cur, int *overflow)
{
    const int i = sab_row_find(H, pv, cur);
    if (i >= 0) { H.tr_cnt[(size_t)pv * H.K + i]++; return; }
    const int nk = H.tr_n[pv];
    if (nk < H.K) { H.tr_key[(size_t)pv * H.K + nk] = cur; H.tr_cnt[(size_t)pv * H.K + nk] = 1; H.tr_n[pv] = nk + 1; }
    else *overflow = 1;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT)
k_resolve_spec(const double *__restrict__ frac, int64_t n_frames, int n_atoms, int n_mobile, const int32_t *__restrict__ mobile,
               int n_sites, int32_t *__restrict__ result, const int32_t *__restrict__ spill, SabPriorityTables P,
               SabHistory G /* global in/out */, int K, int T /* frames per chunk */, int NL /* new-key list capacity (power of two) */,
               int *__restrict__ overflow, SabSpecStats *__restrict__ stats)
{
    extern __shared__ int sm_s[];
    SabSmemHistory H;
    H.K = K;
    H.recent = sm_s;
    H.prev = H.recent + 2 * n_mobile;
    H.tr_n = H.prev + n_mobile;
    H.tr_key = H.tr_n + n_sites;
    H.tr_cnt = H.tr_key + (size_t)n_sites * K;
    int *wcnt = H.tr_cnt + (size_t)n_sites * K;          // [S] writes to the row in this chunk (all atoms)
    int *dec = wcnt + n_sites;                           // [T][A] decisions of the chunk
    int *wl_src = dec + (size_t)T * n_mobile;            // [A][W] own writes: source row
    int *wl_dst = wl_src + (size_t)n_mobile * SAB_SPEC_W;
    int *wl_t = wl_dst + (size_t)n_mobile * SAB_SPEC_W;  // frame within the chunk
    int *rl_row = wl_t + (size_t)n_mobile * SAB_SPEC_W;  // [A][R] rows read
    int *rl_t = rl_row + (size_t)n_mobile * SAB_SPEC_R;
    int *nl_key = rl_t + (size_t)n_mobile * SAB_SPEC_R;  // [NL] new-key candidates: (t * A + a), sorted before they are applied
    int *nl_src = nl_key + NL;
    int *nl_dst = nl_src + NL;
    int *flag = nl_dst + NL;                             // [A] exact frame: atom needs the ordered walk
    int *cur_s = flag + n_mobile;                        // [A]
    int *at_nw = cur_s + n_mobile, *at_nr = at_nw + n_mobile, *at_stop = at_nr + n_mobile, *at_amb = at_stop + n_mobile;   // [A] each: phase 1 -> 2 / 3
    int *at_r0 = at_amb + n_mobile, *at_r1 = at_r0 + n_mobile, *at_pv = at_r1 + n_mobile;                                 // state after the whole chunk
    __shared__ int t_stop, nl_n, any_flag;
    __shared__ unsigned long long st_loc[9];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const bool mine = tid < n_mobile;
    const int a = tid;

    for (int q = tid; q < n_mobile; q += nthr) { H.recent[2 * q] = G.recent[2 * q]; H.recent[2 * q + 1] = G.recent[2 * q + 1]; H.prev[q] = G.prev[q]; }
    for (int s = tid; s < n_sites; s += nthr) {
        const int n = G.tr_n[s];
        H.tr_n[s] = n; wcnt[s] = 0;
        for (int i = 0; i < n && i < K; i++) { H.tr_key[(size_t)s * K + i] = G.tr_key[(size_t)s * G.K + i]; H.tr_cnt[(size_t)s * K + i] = (int)G.tr_cnt[(size_t)s * G.K + i]; }
    }
    if (tid < 9) st_loc[tid] = 0ULL;
    __syncthreads();

    int64_t c0 = 0;
    while (c0 < n_frames) {
        const int Te = (int)min((int64_t)T, n_frames - c0);
        if (tid == 0) { t_stop = Te; nl_n = 0; }
        __syncthreads();
        // ------------------------------------------------------------------ phase 0: the chunk's result entries -> shared memory
        long long tk0 = clock64();
        for (int i = tid; i < Te * n_mobile; i += nthr) dec[i] = result[(size_t)c0 * n_mobile + i];       // coalesced
        __syncthreads();
        long long tk1 = clock64();
        // ------------------------------------------------------------------ phase 1: one WARP per atom, one LANE per frame
        // Lane j holds the entry and the candidate list of frame j.  Whether the atom's most recent site r0 still contains it
        // is tested for ALL frames of the chunk at once (one ballot); the leading run of "stays" is settled in one step, the
        // first frame that is not a stay (3 % of the atom-frames) is handled serially by the whole warp with that lane's list
        // broadcast by shuffles, the state (r0, r1, previous entry) is updated and the remaining frames are tested again.
        for (int q = tid >> 5; q < n_mobile; q += nthr >> 5) {
            const int e = lane < Te ? dec[(size_t)lane * n_mobile + q] : SAB_NONE;
            int n = 0, c[SAB_SPEC_PC];
            {
                const int4 *p = reinterpret_cast<const int4 *>(spill + ((e <= SAB_AMBIGUOUS_BASE) ? (SAB_AMBIGUOUS_BASE - e) : 0));
                int4 q0 = make_int4(0, -3, -3, -3), q1 = make_int4(-3, -3, -3, -3);
                if (e <= SAB_AMBIGUOUS_BASE) { q0 = __ldg(p); q1 = __ldg(p + 1); }
                n = q0.x;
                c[0] = q0.y; c[1] = q0.z; c[2] = q0.w; c[3] = q1.x; c[4] = q1.y; c[5] = q1.z; c[6] = q1.w;
#pragma unroll
                for (int i = 0; i < SAB_SPEC_PC; i++) if (i >= n) c[i] = -3;          // no site id equals -3
            }
            const unsigned ambm = __ballot_sync(0xffffffffu, e <= SAB_AMBIGUOUS_BASE);
            int r0 = H.recent[2 * q], r1 = H.recent[2 * q + 1], pv = H.prev[q];
            int *my_src = wl_src + q * SAB_SPEC_W, *my_dst = wl_dst + q * SAB_SPEC_W, *my_t = wl_t + q * SAB_SPEC_W;
            int nw = 0, nr = 0, stop = INT_MAX, t = 0;
            while (t < Te) {
                // frames in which the atom certainly stays in r0: r0 is in the (short) list, or the entry IS r0
                bool stay = false;
                if (r0 >= 0 && (pv == r0 || pv < 0)) {
                    if (e <= SAB_AMBIGUOUS_BASE) {
                        if (n <= SAB_SPEC_PC) {
#pragma unroll
                            for (int i = 0; i < SAB_SPEC_PC; i++) stay |= c[i] == r0;
                        }
                    } else stay = e == r0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, stay && lane < Te) >> t;
                const int run = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;
                if (run > 0) {                           // no transition (the previous entry is r0 or none), no change of the recent sites
                    if (lane >= t && lane < t + run) dec[(size_t)lane * n_mobile + q] = r0;
                    pv = r0;
                    t += run;
                    if (t >= Te) break;
                }
                // ---- frame t, serially (warp-uniform): its entry and list come from lane t
                const int et = __shfl_sync(0xffffffffu, e, t);
                int cur = et;
                if (et <= SAB_AMBIGUOUS_BASE) {
                    const int nt = __shfl_sync(0xffffffffu, n, t);
                    bool in0 = false, in1 = false;
                    int cand = -1;                       // lane i <-> candidate i of frame t
                    if (nt <= SAB_SPEC_PC) {
#pragma unroll
                        for (int i = 0; i < SAB_SPEC_PC; i++) {
                            const int ci = __shfl_sync(0xffffffffu, c[i], t);
                            in0 |= ci == r0; in1 |= ci == r1;
                            if (lane == i) cand = ci;
                        }
                        in0 &= r0 >= 0; in1 &= r1 >= 0;
                    } else {
                        const int32_t *cl = spill + (SAB_AMBIGUOUS_BASE - et) + 1;
                        in0 = r0 >= 0 && sab_in_list(cl, nt, r0);
                        in1 = r1 >= 0 && sab_in_list(cl, nt, r1);
                        if (lane < nt) cand = cl[lane];
                    }
                    if (in0) cur = r0;
                    else if (in1) cur = r1;
                    else {
                        if (r0 < 0 || nr == SAB_SPEC_R) { stop = t; break; }           // no history (nearest-centre search) / log full
                        if (nt <= 32) cur = sab_spec_decide_warp(cand, r0, n_sites, P, H, my_src, my_dst, nw, lane);
                        else {
                            if (lane == 0) cur = sab_spec_decide(spill + (SAB_AMBIGUOUS_BASE - et) + 1, nt, r0, n_sites, P, H, my_src, my_dst, nw);
                            cur = __shfl_sync(0xffffffffu, cur, 0);
                        }
                        if (lane == 0) { rl_row[q * SAB_SPEC_R + nr] = r0; rl_t[q * SAB_SPEC_R + nr] = t; }
                        nr++;
                    }
                }
                if (cur >= 0 && pv >= 0 && pv != cur) {
                    if (nw == SAB_SPEC_W) { stop = t; break; }
                    if (lane == 0) { my_src[nw] = pv; my_dst[nw] = cur; my_t[nw] = t; atomicAdd(&wcnt[pv], 1); }
                    nw++;
                    __syncwarp();                        // the log is read by lane 0 in a later sab_spec_decide
                }
                if (cur >= 0 && cur != r0) { r1 = r0; r0 = cur; }
                pv = cur;
                if (lane == 0) dec[(size_t)t * n_mobile + q] = cur;
                t++;
            }
            if (lane == 0) { at_nw[q] = nw; at_nr[q] = nr; at_stop[q] = stop; at_amb[q] = (int)ambm; at_r0[q] = r0; at_r1[q] = r1; at_pv[q] = pv; }
        }
        __syncthreads();
        long long tk2 = clock64();
        // ------------------------------------------------------------------ phase 2: confirm the reads
        const int nw = mine ? at_nw[a] : 0, nr = mine ? at_nr[a] : 0, my_stop = mine ? at_stop[a] : INT_MAX;
        const unsigned amb = mine ? (unsigned)at_amb[a] : 0u;
        if (mine) {
            int stop = my_stop;
            const int *my_src = wl_src + a * SAB_SPEC_W;
            for (int i = 0; i < nr; i++) {
                const int row = rl_row[a * SAB_SPEC_R + i];
                int own = 0;
                for (int j = 0; j < nw; j++) own += (my_src[j] == row);
                if (wcnt[row] == own) continue;          // every write of the chunk to this row is the atom's own: confirmed
                // somebody else wrote this row in the chunk: a conflict only if that write comes BEFORE the read in (frame, atom) order
                const int t_read = rl_t[a * SAB_SPEC_R + i], key_read = t_read * n_mobile + a;
                bool conflict = false;
                for (int b = 0; b < n_mobile && !conflict; b++) {
                    if (b == a) continue;
                    const int nb = at_nw[b];
                    for (int j = 0; j < nb; j++)
                        if (wl_src[b * SAB_SPEC_W + j] == row && wl_t[b * SAB_SPEC_W + j] * n_mobile + b < key_read) { conflict = true; break; }
                }
                if (conflict) { stop = min(stop, t_read); break; }
            }
            if (stop < Te) atomicMin(&t_stop, stop);
        }
        __syncthreads();
        const int ts = t_stop;                           // frames [0, ts) of the chunk are exact
        // ------------------------------------------------------------------ phase 3: commit
        if (mine) {
            // the atom's state after the accepted frames: replay its decisions (they need no list)
            int q0 = H.recent[2 * a], q1 = H.recent[2 * a + 1], qp = H.prev[a];
            if (ts == Te && my_stop == INT_MAX) { q0 = at_r0[a]; q1 = at_r1[a]; qp = at_pv[a]; }
            else {
                for (int t = 0; t < ts; t++) {
                    const int cur = dec[(size_t)t * n_mobile + a];
                    if (cur >= 0 && cur != q0) { q1 = q0; q0 = cur; }
                    qp = cur;
                }
            }
            // own writes of the accepted frames: existing keys are incremented in place (commutes), new keys are listed
            for (int j = 0; j < nw; j++) {
                const int t = wl_t[a * SAB_SPEC_W + j];
                if (t >= ts) break;
                const int src = wl_src[a * SAB_SPEC_W + j], dst = wl_dst[a * SAB_SPEC_W + j];
                const int i = sab_row_find(H, src, dst);
                if (i >= 0) atomicAdd(&H.tr_cnt[(size_t)src * K + i], 1);
                else {
                    const int k = atomicAdd(&nl_n, 1);
                    if (k < NL) { nl_key[k] = t * n_mobile + a; nl_src[k] = src; nl_dst[k] = dst; }
                }
            }
            H.recent[2 * a] = q0; H.recent[2 * a + 1] = q1; H.prev[a] = qp;
        }
        for (int s = tid; s < n_sites; s += nthr) wcnt[s] = 0;
        for (int i = tid; i < ts * n_mobile; i += nthr) {   // results of the accepted frames (only ambiguous entries change): all threads, stores only
            const int t = i / n_mobile, q = i - t * n_mobile;
            if ((unsigned)at_amb[q] >> t & 1u) result[(size_t)c0 * n_mobile + i] = dec[i];
        }
        __syncthreads();
        // NOTE: sab_row_find above ran against rows that other threads only INCREMENT meanwhile (keys and tr_n are not
        // modified before the barrier), so the classification existing / new is race-free.
        int n_new = nl_n;
        if (n_new > NL) { if (tid == 0) *overflow = 2; n_new = NL; }     // cannot happen: the host sizes NL >= T * A
        if (n_new > 1) {                                 // bitonic sort by (frame, atom); payload = index into the list
            int m = 2; while (m < n_new) m <<= 1;
            for (int i = tid; i < m; i += nthr) if (i >= n_new) nl_key[i] = INT_MAX;
            __syncthreads();
            for (int k2 = 2; k2 <= m; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < m; i += nthr) {
                        const int l = i ^ j;
                        if (l > i) {
                            const bool up = (i & k2) == 0;
                            const int ki = nl_key[i], kl = nl_key[l];
                            if ((ki > kl) == up) {
                                nl_key[i] = kl; nl_key[l] = ki;
                                int x = nl_src[i]; nl_src[i] = nl_src[l]; nl_src[l] = x;
                                x = nl_dst[i]; nl_dst[i] = nl_dst[l]; nl_dst[l] = x;
                            }
                        }
                    }
                    __syncthreads();
                }
        }
        if (tid == 0) {
            for (int i = 0; i < n_new; i++) sab_spec_add(H, nl_src[i], nl_dst[i], overflow);
            st_loc[0]++; st_loc[4] += (unsigned long long)n_new;
            if (ts < Te) st_loc[1]++;
        }
        __syncthreads();
        c0 += ts;
        long long tk3 = clock64();
        if (tid == 0) { st_loc[5] += (unsigned long long)(tk1 - tk0); st_loc[6] += (unsigned long long)(tk2 - tk1); st_loc[7] += (unsigned long long)(tk3 - tk2); }
        if (ts == Te) continue;
        // ------------------------------------------------------------------ the truncating frame, exactly (as k_resolve)
        {
            const int64_t f = c0;
            if (tid == 0) any_flag = 0;
            __syncthreads();
            if (mine) {
                const int r = result[(size_t)f * n_mobile + a];
                int cur = r, fl = 0;
                if (r <= SAB_AMBIGUOUS_BASE) {
                    const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - r);
                    const int n = c[0]; c++;
                    const int q0 = H.recent[2 * a], q1 = H.recent[2 * a + 1];
                    if (q0 >= 0 && sab_in_list(c, n, q0)) cur = q0;
                    else if (q1 >= 0 && sab_in_list(c, n, q1)) cur = q1;
                    else fl = 1;
                }
                if (!fl) {
                    const int qp = H.prev[a];
                    if (cur >= 0 && qp >= 0 && qp != cur) fl = 1;              // adds a transition: ordered
                    else {
                        const int q0 = H.recent[2 * a];
                        if (cur >= 0 && cur != q0) { H.recent[2 * a + 1] = q0; H.recent[2 * a] = cur; }
                        H.prev[a] = cur;
                        result[(size_t)f * n_mobile + a] = cur;
                    }
                }
                if (fl) any_flag = 1;
                flag[a] = fl; cur_s[a] = cur;
            }
            __syncthreads();
            if (any_flag && tid < 32) {
                for (int q = 0; q < n_mobile; q++) {
                    if (!flag[q]) continue;
                    int cur = cur_s[q];
                    if (cur <= SAB_AMBIGUOUS_BASE) {
                        const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - cur);
                        const int n = c[0]; c++;
                        double x[3];
                        for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frac[((size_t)f * n_atoms + mobile[q]) * 3 + d]);
                        cur = sab_resolve_with_counters_smem(c, n, H.recent[2 * q], x, n_sites, P, H, lane);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        const int qp = H.prev[q];
                        if (cur >= 0 && qp >= 0 && qp != cur) sab_spec_add(H, qp, cur, overflow);
                        const int q0 = H.recent[2 * q];
                        if (cur >= 0 && cur != q0) { H.recent[2 * q + 1] = q0; H.recent[2 * q] = cur; }
                        H.prev[q] = cur;
                        result[(size_t)f * n_mobile + q] = cur;
                    }
                    __syncwarp();
                }
            }
            if (tid == 0) { st_loc[2]++; st_loc[8] += (unsigned long long)(clock64() - tk3); }
            __syncthreads();
            c0 += 1;
        }
    }
    __syncthreads();
    for (int q = tid; q < n_mobile; q += nthr) { G.recent[2 * q] = H.recent[2 * q]; G.recent[2 * q + 1] = H.recent[2 * q + 1]; G.prev[q] = H.prev[q]; }
    for (int s = tid; s < n_sites; s += nthr) {
        const int n = H.tr_n[s];
        G.tr_n[s] = n;
        for (int i = 0; i < n; i++) { G.tr_key[(size_t)s * G.K + i] = H.tr_key[(size_t)s * K + i]; G.tr_cnt[(size_t)s * G.K + i] = H.tr_cnt[(size_t)s * K + i]; }
    }
    if (stats && tid == 0) { stats->chunks = st_loc[0]; stats->truncated = st_loc[1]; stats->exact_frames = st_loc[2]; stats->reads = st_loc[3]; stats->new_keys = st_loc[4]; stats->cyc_load = st_loc[5]; stats->cyc_walk = st_loc[6]; stats->cyc_commit = st_loc[7]; stats->cyc_exact = st_loc[8]; }
}
