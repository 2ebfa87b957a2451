// k_polyhedral_lazy.cuh -- the EXACT path for polyhedral sites when the frame-parallel formulation's guards fail.
//
// The reference keeps one PBC-shift cache per SITE and touches it only when the site is QUERIED
// (PolyhedralSite.notify_structure_changed -> contains_point -> _store_vertex_coords, polyhedral_site.py:322-416):
//   incremental  shift -= rint(raw_t - raw_t'), t' = the frame of the site's previous query, valid while every
//                |raw_t - raw_t' - rint| < 0.3 (pbc_utils.py:193-230);
//   otherwise    the full correction (27 images to the reference centre, or the legacy spread rule; :65-145).
// Which sites are queried in a frame depends on the priority walk of every atom (site_collection.py:113-182), i.e. on
// the results themselves, so after a framework step >= 0.3 (cache invalidation) or an unwrapped excursion >= 0.3 over
// the trajectory (guard G1) the caches are no longer a function of the frame alone.  This kernel
// replays the reference literally: frames in order, per-site caches with the frame stamp of their last query, atoms
// in list order wherever they can influence one another (the shared transition Counters), one site query at a time
// along the priority walk so that exactly the reference's set of sites is queried.
//
// One CTA.  Per frame:
//   phase 0  wrapped coordinates (atom.py:104) and the 8 images x + {0,1}^3 (tools.py:246-253) of every atom;
//   phase 1  all warps: every atom with a history queries its most recent site (the first query of its walk in the
//            reference too).  A hit that adds no transition (previous trajectory entry == that site) is final;
//   phase 2  warp 0, ascending atom order: every other atom walks its whole priority list from the start (repeated
//            queries of a frame are idempotent: the cache carries the frame stamp) and updates the Counters.
// A query is warp-cooperative: lanes share the 3 V coordinates of the cache update and the (image, face) pairs of
// the containment test; float64, the reference's operation order (sab_device.cuh), no FMA.
#pragma once
#include "k_polyhedral.cuh"
#include "k_resolve.cuh"

struct SabLazy {
    int smax, fmax;
    const int32_t *n_slot;       // [S]
    const int32_t *slot_atom;    // [S][smax] structure index
    const int32_t *n_face;       // [S]
    const uint8_t *simp;         // [S][fmax][3]
    const uint8_t *has_rc;       // [S] site has a reference centre
    const double *rc;            // [S][3]
    int32_t *shift;              // [S][smax][3] _pbc_image_shifts
    double *cached;              // [S][smax][3] _pbc_cached_raw_frac
    int32_t *has_shift;          // [S]
    long long *stamp;            // [S] global frame number of the last query (-1 = never)
    int32_t *lock;               // [S]
    double *vc;                  // [S][smax][3] vertex_coords of the last query
    unsigned long long *stats;   // [4] queries, incremental updates, full corrections, invalidations
    // Face topology hulled at the site's FIRST query (polyhedral_site.py:468-471): for non-simplicial polyhedra Qhull's
    // triangulation depends on the vertex coordinates of that frame.  Sites whose topology is still provisional are
    // marked open; the kernel records the frame and the vertex coordinates of their first query, the host re-hulls from
    // exactly those and redoes the batch if a triangulation changed (engine._assign_device_lazy).  All three may be null.
    const int32_t *open_topo;    // [S] 1 = topology provisional
    long long *first_frame;      // [S] -1, or the global frame number of the first query in this run
    double *first_vc;            // [S][smax][3] vertex_coords of that query
};

// _store_vertex_coords (polyhedral_site.py:377-416) for site s in the current frame, by one warp.
__device__ void sab_lazy_store(const SabLazy &Z, int s, const double *__restrict__ frame, const double *__restrict__ L, int lane)
{
    const int nv = Z.n_slot[s], n3 = 3 * nv;
    const int32_t *sa = Z.slot_atom + (size_t)s * Z.smax;
    int32_t *shift = Z.shift + (size_t)s * Z.smax * 3;
    double *cached = Z.cached + (size_t)s * Z.smax * 3, *vc = Z.vc + (size_t)s * Z.smax * 3;
    const bool has = Z.has_shift[s] != 0;
    bool bad = false;
    if (has)                                                    // update_pbc_shifts, pbc_utils.py:210-217
        for (int i = lane; i < n3; i += 32) {
            const double raw = frame[(size_t)sa[i / 3] * 3 + i % 3];
            const double diff = raw - cached[i], w = rint(diff), phys = diff - w;
            if (phys >= 0.3 || phys <= -0.3) bad = true;
        }
    bad = __any_sync(0xffffffffu, bad);
    if (has && !bad) {
        for (int i = lane; i < n3; i += 32) {
            const double raw = frame[(size_t)sa[i / 3] * 3 + i % 3];
            const int ns = shift[i] - (int)rint(raw - cached[i]);
            shift[i] = ns; cached[i] = raw;
            vc[i] = raw + (double)ns;                           // pbc_utils.py:220
        }
        __syncwarp();
        double u = 0.0;
        if (lane < 3) {                                         // non-negative offset, pbc_utils.py:221-229
            double mn = vc[lane];
            for (int v = 1; v < nv; v++) if (vc[3 * v + lane] < mn) mn = vc[3 * v + lane];
            if (mn < 0.0) u = ceil(-mn);
        }
        const double u0 = __shfl_sync(0xffffffffu, u, 0), u1 = __shfl_sync(0xffffffffu, u, 1), u2 = __shfl_sync(0xffffffffu, u, 2);
        __syncwarp();
        for (int i = lane; i < n3; i += 32) vc[i] += (i % 3 == 0 ? u0 : (i % 3 == 1 ? u1 : u2));
        if (lane == 0) atomicAdd(&Z.stats[1], 1ULL);
    } else {                                                    // correct_pbc, pbc_utils.py:112-145
        if (lane == 0) { atomicAdd(&Z.stats[2], 1ULL); if (has) atomicAdd(&Z.stats[3], 1ULL); }
        if (Z.has_rc[s]) {                                      // unwrap_vertices_to_reference_center, :65-109
            const double *rc = Z.rc + 3 * (size_t)s;
            const double rcx = rc[0] * L[0] + rc[1] * L[3] + rc[2] * L[6];
            const double rcy = rc[0] * L[1] + rc[1] * L[4] + rc[2] * L[7];
            const double rcz = rc[0] * L[2] + rc[1] * L[5] + rc[2] * L[8];
            for (int v = lane; v < nv; v += 32) {
                double r[3];
                for (int d = 0; d < 3; d++) r[d] = frame[(size_t)sa[v] * 3 + d];
                double best = INFINITY; int bs0 = 0, bs1 = 0, bs2 = 0;
                for (int dx = -1; dx < 2; dx++)
                    for (int dy = -1; dy < 2; dy++)
                        for (int dz = -1; dz < 2; dz++) {
                            const double p0 = r[0] + dx, p1 = r[1] + dy, p2 = r[2] + dz;
                            const double cx = p0 * L[0] + p1 * L[3] + p2 * L[6] - rcx;
                            const double cy = p0 * L[1] + p1 * L[4] + p2 * L[7] - rcy;
                            const double cz = p0 * L[2] + p1 * L[5] + p2 * L[8] - rcz;
                            const double dist = sqrt(cx * cx + cy * cy + cz * cz);
                            if (dist < best) { best = dist; bs0 = dx; bs1 = dy; bs2 = dz; }
                        }
                shift[3 * v] = bs0; shift[3 * v + 1] = bs1; shift[3 * v + 2] = bs2;
                for (int d = 0; d < 3; d++) { cached[3 * v + d] = r[d]; }
                vc[3 * v] = r[0] + (double)bs0; vc[3 * v + 1] = r[1] + (double)bs1; vc[3 * v + 2] = r[2] + (double)bs2;
            }
            __syncwarp();
            double u = 0.0;
            if (lane < 3) {
                double mn = vc[lane];
                for (int v = 1; v < nv; v++) if (vc[3 * v + lane] < mn) mn = vc[3 * v + lane];
                u = ceil(-mn); if (!(u > 0.0)) u = 0.0;         // np.maximum(0, np.ceil(-min))
            }
            const double u0 = __shfl_sync(0xffffffffu, u, 0), u1 = __shfl_sync(0xffffffffu, u, 1), u2 = __shfl_sync(0xffffffffu, u, 2);
            __syncwarp();
            for (int i = lane; i < n3; i += 32) vc[i] = vc[i] + (i % 3 == 0 ? u0 : (i % 3 == 1 ? u1 : u2));
        } else {                                                // apply_legacy_pbc_correction, :16-41 (no offset afterwards)
            for (int i = lane; i < n3; i += 32) { const double raw = frame[(size_t)sa[i / 3] * 3 + i % 3]; cached[i] = raw; vc[i] = raw; }
            __syncwarp();
            bool spread = false;
            if (lane < 3) {
                double mx = vc[lane], mn = vc[lane];
                for (int v = 1; v < nv; v++) { const double x = vc[3 * v + lane]; if (x > mx) mx = x; if (x < mn) mn = x; }
                spread = mx - mn > 0.5;
            }
            const unsigned sm = __ballot_sync(0xffffffffu, spread);
            __syncwarp();
            for (int i = lane; i < n3; i += 32) {
                double x = vc[i];
                if (((sm >> (i % 3)) & 1u) && x < 0.5) x += 1.0;
                shift[i] = (int)rint(x - vc[i]);
                vc[i] = x;
            }
        }
        if (lane == 0) Z.has_shift[s] = 1;
    }
    __syncwarp();
}

// PolyhedralSite.contains_point (polyhedral_site.py:431-485) for site s, frame stamp fs: does any of the 8 images lie
// inside?  Warp-cooperative; all lanes return the answer.
__device__ bool sab_lazy_query(const SabLazy &Z, int s, long long fs, const double *__restrict__ frame, const double *__restrict__ L,
                               const double *__restrict__ img, int lane)
{
    if (lane == 0) atomicAdd(&Z.stats[0], 1ULL);
    // the first query of the frame brings the cache up to date; warps of phase 1 may meet at the same site
    for (;;) {
        const long long st = *((volatile long long *)&Z.stamp[s]);
        if (__shfl_sync(0xffffffffu, st, 0) == fs) break;
        int got = 0;
        if (lane == 0) got = atomicCAS(&Z.lock[s], 0, 1) == 0;
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got) {
            if (*((volatile long long *)&Z.stamp[s]) != fs) {
                sab_lazy_store(Z, s, frame, L, lane);
                if (Z.open_topo && Z.open_topo[s]) {            // provisional topology: keep what the reference hulls from
                    const long long ff = *((volatile long long *)&Z.first_frame[s]);
                    if (__shfl_sync(0xffffffffu, ff, 0) < 0) {
                        const int n3 = 3 * Z.n_slot[s];
                        for (int i = lane; i < n3; i += 32) Z.first_vc[(size_t)s * Z.smax * 3 + i] = Z.vc[(size_t)s * Z.smax * 3 + i];
                        __syncwarp();
                        if (lane == 0) *((volatile long long *)&Z.first_frame[s]) = fs;
                    }
                }
                __threadfence_block();
                if (lane == 0) *((volatile long long *)&Z.stamp[s]) = fs;
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) atomicExch(&Z.lock[s], 0);
            break;
        }
    }
    __syncwarp();
    const int nv = Z.n_slot[s], nf = Z.n_face[s];
    const double *vcg = Z.vc + (size_t)s * Z.smax * 3;
    const uint8_t *simp = Z.simp + (size_t)s * Z.fmax * 3;
    double cen[3];
    for (int d = 0; d < 3; d++) {                               // _numba_update_faces: centroid = sequential sum / n
        double sum = 0.0;
        for (int v = 0; v < nv; v++) sum += __ldcg(vcg + 3 * v + d);
        cen[d] = sum / (double)nv;
    }
    bool rej = false;                                           // lane -> image lane & 7, faces lane >> 3, + 4, ...
    const int im = lane & 7;
    for (int j = lane >> 3; j < nf; j += 4) {
        double v0[3], v1[3], v2[3];
        for (int d = 0; d < 3; d++) {
            v0[d] = __ldcg(vcg + 3 * simp[3 * j] + d); v1[d] = __ldcg(vcg + 3 * simp[3 * j + 1] + d); v2[d] = __ldcg(vcg + 3 * simp[3 * j + 2] + d);
        }
        if (!sab_face_accepts(v0, v1, v2, cen, img[3 * im], img[3 * im + 1], img[3 * im + 2])) rej = true;
    }
    unsigned m = __ballot_sync(0xffffffffu, rej);
    m |= m >> 8; m |= m >> 16;
    return ((~m) & 0xffu) != 0u;
}

__global__ void __launch_bounds__(256)
k_polyhedral_lazy(const double *__restrict__ frac, const double *__restrict__ lattice, int lattice_stride, int64_t n_frames,
                  long long frame0, int n_atoms, int n_mobile, const int32_t *__restrict__ mobile, int n_sites, SabLazy Z,
                  SabPriorityTables P, SabHistory H, int append, int32_t *__restrict__ result, int *__restrict__ overflow)
{
    extern __shared__ double sm_lazy[];                         // img[A][8][3] | flag[A]
    double *img = sm_lazy;
    int *flag = reinterpret_cast<int *>(img + (size_t)n_mobile * 24);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    for (int64_t f = 0; f < n_frames; f++) {
        const double *frame = frac + (size_t)f * n_atoms * 3;
        const double *L = lattice + (size_t)f * lattice_stride;
        const long long fs = frame0 + f;
        for (int a = tid; a < n_mobile; a += blockDim.x) {
            double x[3];
            for (int d = 0; d < 3; d++) x[d] = sab_pymod1(frame[(size_t)mobile[a] * 3 + d]);
            for (int i = 0; i < 8; i++) {                       // tools.py:246-253 image order
                const int o0 = (i == 1 || i == 4 || i == 5 || i == 7), o1 = (i == 2 || i == 4 || i == 6 || i == 7), o2 = (i == 3 || i == 5 || i == 6 || i == 7);
                img[a * 24 + 3 * i] = (double)o0 + x[0]; img[a * 24 + 3 * i + 1] = (double)o1 + x[1]; img[a * 24 + 3 * i + 2] = (double)o2 + x[2];
            }
        }
        __syncthreads();
        // ---- phase 1: the most recent site of every atom, all warps ------------------------------------------------
        for (int a = warp; a < n_mobile; a += n_warps) {
            const int r0 = H.recent[2 * a];
            int fl = 1;
            if (r0 >= 0 && H.prev[a] == r0) {
                if (sab_lazy_query(Z, r0, fs, frame, L, img + a * 24, lane)) { fl = 0; if (lane == 0) result[(size_t)f * n_mobile + a] = r0; }
            }
            if (lane == 0) flag[a] = fl;
        }
        __syncthreads();
        // ---- phase 2: everything that walks on or touches the Counters, warp 0 in atom order -----------------------
        if (warp == 0) {
            for (int a = 0; a < n_mobile; a++) {
                if (!flag[a]) continue;
                const double *im = img + a * 24;
                const int r0 = H.recent[2 * a], r1 = H.recent[2 * a + 1];
                int anchor = -1, cur = -1;
                #define SAB_TRY(s_) (sab_lazy_query(Z, (s_), fs, frame, L, im, lane))
                if (r0 >= 0 || r1 >= 0) {                       // site_collection.py:147-152
                    anchor = r0 >= 0 ? r0 : r1;
                    if (r0 >= 0 && SAB_TRY(r0)) cur = r0;
                    else if (r1 >= 0 && SAB_TRY(r1)) cur = r1;
                } else if (P.lookup) {                          // :153-156 nearest centre (fractional metric), itself tested first
                    double bd = INFINITY; int bi = 0;
                    for (int s = lane; s < n_sites; s += 32) {
                        double d0 = P.lookup[3 * s] - im[0], d1 = P.lookup[3 * s + 1] - im[1], d2 = P.lookup[3 * s + 2] - im[2];
                        d0 -= rint(d0); d1 -= rint(d1); d2 -= rint(d2);
                        const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                        if (nrm < bd) { bd = nrm; bi = s; }
                    }
                    for (int o = 16; o; o >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                    }
                    anchor = bi;
                    if (SAB_TRY(anchor)) cur = anchor;
                }
                if (cur < 0 && anchor >= 0) {
                    // learned transitions of the anchor: count descending, ties by insertion order (site.py:216-223)
                    const int nk = H.tr_n[anchor];
                    const int32_t *key = H.tr_key + (size_t)anchor * H.K;
                    const long long *cn = H.tr_cnt + (size_t)anchor * H.K;
                    long long last_c = LLONG_MAX; int last_i = -1;
                    for (int t = 0; t < nk && cur < 0; t++) {    // the t-th entry of the stable sort
                        long long bc = -1; int bi = -1;
                        for (int i = 0; i < nk; i++) {
                            const bool after = cn[i] < last_c || (cn[i] == last_c && i > last_i);
                            if (after && cn[i] > bc) { bc = cn[i]; bi = i; }
                        }
                        last_c = bc; last_i = bi;
                        const int d = key[bi];
                        if (d == r0 || d == r1) continue;       // already tested
                        if (SAB_TRY(d)) cur = d;
                    }
                    auto tested = [&](int d) {
                        if (d == r0 || d == r1 || (r0 < 0 && r1 < 0 && d == anchor)) return true;
                        for (int i = 0; i < nk; i++) if (key[i] == d) return true;
                        return false;
                    };
                    const int slot = (!P.rank && P.row_slot) ? P.row_slot[anchor] : -1;
                    if (cur < 0 && P.ranked && (P.rank || slot >= 0)) {   // :167-171 the anchor's distance-ranked row
                        const int32_t *rk = P.rank ? P.rank + (size_t)anchor * (n_sites - 1) : P.row_fwd + (size_t)slot * n_sites;
                        for (int i = 0; i < n_sites - 1 && cur < 0; i++) { const int d = rk[i]; if (!tested(d) && SAB_TRY(d)) cur = d; }
                    } else if (cur < 0 && P.ranked) {
                        // no cached row: walk the sites by increasing distance to the anchor (numpy's arithmetic, sab_rank_key);
                        // equal distances are ordered by the sort implementation of the reference's np.argsort -- reported as
                        // missing, the host supplies the true row and the batch is redone
                        double last_k = -1.0; int last_s = -1;
                        for (int step = 0; step < n_sites - 1 && cur < 0; step++) {
                            double bk = INFINITY; int bs = INT_MAX;
                            for (int s2 = lane; s2 < n_sites; s2 += 32) {
                                if (s2 == anchor) continue;
                                bool ex;
                                const double k = sab_rank_key(P, anchor, s2, n_sites, ex);
                                const bool after = k > last_k || (k == last_k && s2 > last_s);
                                if (after && (k < bk || (k == bk && s2 < bs))) { bk = k; bs = s2; }
                            }
                            for (int o = 16; o; o >>= 1) {
                                const double ok = __shfl_xor_sync(0xffffffffu, bk, o);
                                const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                                if (ok < bk || (ok == bk && os < bs)) { bk = ok; bs = os; }
                            }
                            if (bs == INT_MAX) break;
                            if (bk == last_k && lane == 0) sab_rank_missing(P, anchor);
                            last_k = bk; last_s = bs;
                            if (!tested(bs) && SAB_TRY(bs)) cur = bs;
                        }
                    } else if (cur < 0) {                       // :172-179 face-sharing neighbours, then list order
                        const int n0 = P.neigh_off ? P.neigh_off[anchor] : 0, n1 = P.neigh_off ? P.neigh_off[anchor + 1] : 0;
                        for (int i = n0; i < n1 && cur < 0; i++) { const int d = P.neigh[i]; if (!tested(d) && SAB_TRY(d)) cur = d; }
                        for (int s = 0; s < n_sites && cur < 0; s++) {
                            if (tested(s)) continue;
                            bool nb = false;
                            for (int i = n0; i < n1; i++) if (P.neigh[i] == s) nb = true;
                            if (!nb && SAB_TRY(s)) cur = s;
                        }
                    }
                } else if (cur < 0) {                           // :180-182 no anchor at all: list order
                    for (int s = 0; s < n_sites && cur < 0; s++) if (SAB_TRY(s)) cur = s;
                }
                #undef SAB_TRY
                __syncwarp();
                if (lane == 0) {
                    const int pv = H.prev[a];
                    if (cur >= 0 && pv >= 0 && pv != cur) {     // update_occupation, site_collection.py:300-306
                        int nk = H.tr_n[pv], i = 0;
                        int32_t *key = H.tr_key + (size_t)pv * H.K;
                        long long *cn = H.tr_cnt + (size_t)pv * H.K;
                        while (i < nk && key[i] != cur) i++;
                        if (i < nk) cn[i]++;
                        else if (nk < H.K) { key[nk] = cur; cn[nk] = 1; H.tr_n[pv] = nk + 1; }
                        else *overflow = 1;
                    }
                    if (cur >= 0) sab_touch_recent(H.recent, a, cur);
                    result[(size_t)f * n_mobile + a] = cur;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (append)                                             // trajectory.py:357: the previous entry of the next frame
            for (int a = tid; a < n_mobile; a += blockDim.x) H.prev[a] = result[(size_t)f * n_mobile + a];
        __syncthreads();
    }
}
