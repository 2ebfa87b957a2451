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
//   phase 2  a read is confirmed when all writes of this chunk to its row are the reading atom's own.  The first
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
#define SAB_SPEC_PC 11          // candidates kept in registers per prefetched list: three 128-bit loads {n, c_0 .. c_10}

struct SabSpecStats { unsigned long long chunks, truncated, exact_frames, reads, new_keys; };

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
    __shared__ int t_stop, nl_n, any_flag;
    __shared__ unsigned long long st_loc[5];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const bool mine = tid < n_mobile;
    const int a = tid;

    for (int q = tid; q < n_mobile; q += nthr) { H.recent[2 * q] = G.recent[2 * q]; H.recent[2 * q + 1] = G.recent[2 * q + 1]; H.prev[q] = G.prev[q]; }
    for (int s = tid; s < n_sites; s += nthr) {
        const int n = G.tr_n[s];
        H.tr_n[s] = n; wcnt[s] = 0;
        for (int i = 0; i < n && i < K; i++) { H.tr_key[(size_t)s * K + i] = G.tr_key[(size_t)s * G.K + i]; H.tr_cnt[(size_t)s * K + i] = (int)G.tr_cnt[(size_t)s * G.K + i]; }
    }
    if (tid < 5) st_loc[tid] = 0ULL;
    __syncthreads();

    int64_t c0 = 0;
    while (c0 < n_frames) {
        const int Te = (int)min((int64_t)T, n_frames - c0);
        if (tid == 0) { t_stop = Te; nl_n = 0; }
        __syncthreads();
        // ------------------------------------------------------------------ phase 0: the chunk's result entries -> shared memory
        // (coalesced, independent loads; afterwards the address of every candidate list of the chunk is known at once, so the
        // lists can be fetched several frames ahead -- they are the only global-memory reads left in phase 1)
        unsigned amb = 0u;                               // bit t: this atom's entry of frame t is ambiguous (commit writes only those)
        if (mine)
            for (int t = 0; t < Te; t++) dec[(size_t)t * n_mobile + a] = result[(size_t)(c0 + t) * n_mobile + a];
        // (each thread reads back only what it wrote itself: no barrier needed)
        // ------------------------------------------------------------------ phase 1: every atom on its own
        int nw = 0, nr = 0, my_stop = INT_MAX;
        int r0 = -1, r1 = -1, pv = -2;
        if (mine) {
            r0 = H.recent[2 * a]; r1 = H.recent[2 * a + 1]; pv = H.prev[a];
            int *my_src = wl_src + a * SAB_SPEC_W, *my_dst = wl_dst + a * SAB_SPEC_W, *my_t = wl_t + a * SAB_SPEC_W;
            // candidate lists: into L2 eleven frames ahead, into registers three frames ahead (no load depends on the atom's
            // state).  Three register buffers used round-robin (frame t lives in buffer t mod 3; the loop is unrolled by three
            // so that no buffer is ever copied); slots behind a list's end are filled with -3, which no site id equals.
            int er[3], cn[3], cc[3][SAB_SPEC_PC];
            const int *ent = dec + a;                    // this atom's column of the chunk's entries / decisions
            auto load_r = [&](int t) -> int { return t < Te ? ent[(size_t)t * n_mobile] : SAB_NONE; };
            auto load_c = [&](int r, int &n, int *c) {           // records are 16-byte aligned (SAB_SPILL_RECORD): 3 loads, not 12
                const int4 *p = reinterpret_cast<const int4 *>(spill + ((r <= SAB_AMBIGUOUS_BASE) ? (SAB_AMBIGUOUS_BASE - r) : 0));
                const int4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
                n = q0.x;
                c[0] = q0.y; c[1] = q0.z; c[2] = q0.w; c[3] = q1.x; c[4] = q1.y; c[5] = q1.z; c[6] = q1.w;
                c[7] = q2.x; c[8] = q2.y; c[9] = q2.z; c[10] = q2.w;
            };
            auto touch_c = [&](int t) {
                const int r = load_r(t);
                if (r <= SAB_AMBIGUOUS_BASE) asm volatile("prefetch.global.L2 [%0];" ::"l"(spill + (SAB_AMBIGUOUS_BASE - r)));
            };
            for (int t = 3; t < 11; t++) touch_c(t);
#pragma unroll
            for (int b = 0; b < 3; b++) { er[b] = load_r(b); load_c(er[b], cn[b], cc[b]); }
            bool stop_now = false;
            // one frame, its list in buffer B (a compile-time constant)
#define SAB_SPEC_STEP(B, t_)                                                                                              \
            do {                                                                                                          \
                const int t = (t_);                                                                                       \
                if (t >= Te) break;                                                                                       \
                const int r = er[B];                                                                                      \
                int cur = r;                                                                                              \
                touch_c(t + 11);                                                                                          \
                if (r <= SAB_AMBIGUOUS_BASE) {                                                                            \
                    amb |= 1u << t;                                                                                       \
                    bool in0 = false, in1 = false;                                                                        \
                    const int n = cn[B];                                                                                  \
                    if (n <= SAB_SPEC_PC) {                                                                               \
                        _Pragma("unroll") for (int i = 0; i < SAB_SPEC_PC; i++) { const int ci = i < n ? cc[B][i] : -3; in0 |= ci == r0; in1 |= ci == r1; } \
                    } else {                                                                                              \
                        const int32_t *c = spill + (SAB_AMBIGUOUS_BASE - r) + 1;                                          \
                        in0 = r0 >= 0 && sab_in_list(c, n, r0);                                                           \
                        in1 = r1 >= 0 && sab_in_list(c, n, r1);                                                           \
                    }                                                                                                     \
                    if (in0) cur = r0;                                                                                    \
                    else if (in1) cur = r1;                                                                               \
                    else {                                                                                                \
                        if (r0 < 0 || nr == SAB_SPEC_R) { my_stop = t; stop_now = true; break; }   /* no history (nearest-centre search) / log full */ \
                        cur = sab_spec_decide(spill + (SAB_AMBIGUOUS_BASE - r) + 1, n, r0, n_sites, P, H, my_src, my_dst, nw); \
                        rl_row[a * SAB_SPEC_R + nr] = r0; rl_t[a * SAB_SPEC_R + nr] = t; nr++;                            \
                    }                                                                                                     \
                }                                                                                                         \
                if (cur >= 0 && pv >= 0 && pv != cur) {                                                                   \
                    if (nw == SAB_SPEC_W) { my_stop = t; stop_now = true; break; }                                        \
                    my_src[nw] = pv; my_dst[nw] = cur; my_t[nw] = t; nw++;                                                \
                    atomicAdd(&wcnt[pv], 1);                                                                              \
                }                                                                                                         \
                if (cur >= 0 && cur != r0) { r1 = r0; r0 = cur; }                                                         \
                pv = cur;                                                                                                 \
                er[B] = load_r(t + 3);                       /* before the decision overwrites nothing it needs: t + 3 > t */ \
                load_c(er[B], cn[B], cc[B]);                                                                              \
                dec[(size_t)t * n_mobile + a] = cur;                                                                      \
            } while (0)
            for (int t3 = 0; t3 < Te && !stop_now; t3 += 3) {
                SAB_SPEC_STEP(0, t3);
                if (stop_now) break;
                SAB_SPEC_STEP(1, t3 + 1);
                if (stop_now) break;
                SAB_SPEC_STEP(2, t3 + 2);
            }
#undef SAB_SPEC_STEP
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase 2: confirm the reads
        if (mine) {
            int stop = my_stop;
            const int *my_src = wl_src + a * SAB_SPEC_W;
            for (int i = 0; i < nr; i++) {
                const int row = rl_row[a * SAB_SPEC_R + i];
                int own = 0;
                for (int j = 0; j < nw; j++) own += (my_src[j] == row);
                if (wcnt[row] != own) { stop = min(stop, rl_t[a * SAB_SPEC_R + i]); break; }   // somebody else wrote this row in the chunk
            }
            if (stop < Te) atomicMin(&t_stop, stop);
        }
        __syncthreads();
        const int ts = t_stop;                           // frames [0, ts) of the chunk are exact
        // ------------------------------------------------------------------ phase 3: commit
        if (mine) {
            // the atom's state after the accepted frames: replay its decisions (they need no list)
            int q0 = H.recent[2 * a], q1 = H.recent[2 * a + 1], qp = H.prev[a];
            if (ts == Te && my_stop == INT_MAX) { q0 = r0; q1 = r1; qp = pv; }
            else
                for (int t = 0; t < ts; t++) {
                    const int cur = dec[(size_t)t * n_mobile + a];
                    if (cur >= 0 && cur != q0) { q1 = q0; q0 = cur; }
                    qp = cur;
                }
            for (int t = 0; t < ts; t++)                 // results (only ambiguous entries change): stores only
                if (amb >> t & 1u) result[(size_t)(c0 + t) * n_mobile + a] = dec[(size_t)t * n_mobile + a];
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
            if (tid == 0) st_loc[2]++;
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
    if (stats && tid == 0) { stats->chunks = st_loc[0]; stats->truncated = st_loc[1]; stats->exact_frames = st_loc[2]; stats->reads = st_loc[3]; stats->new_keys = st_loc[4]; }
}
