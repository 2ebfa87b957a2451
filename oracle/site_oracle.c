/*
 * site_oracle.c -- CPU ORACLE for the per-frame site-assignment hot path of
 * bjmorgan/site-analysis (reference v1.8.0 at /root/reference).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a plain-C restatement of
 * the reference's sequential algorithm (float64, no FMA, round-half-even, IEEE
 * sqrt) and is used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the CHECKER and the CPU baseline.  The
 * product package (site_analysis_b200) never imports, links or executes it.
 *
 * Parity status: PINNED.  oracle/validate_oracle.py runs the unmodified reference
 * (numba path) and this oracle on identical inputs in the build container, and
 * tests/test_oracle_golden.py checks it against every committed fixture under
 * tests/golden/ (generated from the reference by oracle/gen_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: the reference's numba/LLVM code emits separate
 * vmulsd/vaddsd, never FMA (SURVEY.md headline fact 3).
 *
 * Each function cites the reference file:line it restates.  Sites are addressed
 * by list position (0..S-1); the Python wrapper maps positions <-> Site.index.
 *
 * Known, documented deviation: the full 27-image unwrap to a reference centre
 * (pbc_utils.py:65-109) converts to Cartesian through OpenBLAS `@` in the
 * reference (FMA rounding); here it is plain mul/add.  Only the integer argmin
 * decision is used, so this matters only for exact near-ties between two
 * periodic images of a vertex (|difference| ~ 1 ulp).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SO_SPHERICAL 0
#define SO_VORONOI 1
#define SO_POLYHEDRAL 2
#define SO_DYNVOR 3

typedef int (*so_hull_fn)(void *ctx, int site, int nv, const double *vc, int max_faces, int32_t *out);

typedef struct { int n, cap; int *key; int64_t *cnt; } Counter;

typedef struct {
    int nv; int *vidx;
    int has_rc; double rc[3];
    int has_shift; int64_t *shift; double *cached_raw;
    int has_vc; double *vc;
    int pending;
    int has_topo; int nf; int32_t *simp; double *normals, *refs, *signs;
    int stale;
} PolySite;

typedef struct {
    int n, R; int *site_pos; int *ref_idx; int64_t *shift; double *cached; int initialised;
    double *batch;            /* scratch (n,R,3) */
} Group;

typedef struct {
    int kind, S, A;
    int *atom_index;
    double *ax;               /* (A,3) wrapped mobile coords: atom.py:104 */
    int *in_site, *prev, *recent;
    int64_t n_traj;
    Counter *tr;
    int has_rank; int32_t *rank; double *look_c;
    int **neigh; int *n_neigh;
    int *mark; int cur_mark;
    double *centres, *rcut;
    PolySite *ps; const double *cur_frac; double lat[9];
    so_hull_fn hull; void *hull_ctx;
    int n_groups; Group *groups; int *dv_has_rc; double *dv_rc; int dv_centres_set;
    int64_t n_contains, n_full, n_incr, n_invalid, n_face_upd, n_hull, n_lookup;
    int error; char msg[256];
    int *order_tmp;
} Oracle;

/* ------------------------------------------------------------------ helpers */

/* numpy float remainder by 1.0 (Python-mod semantics) -- atom.py:104,
 * dynamic_voronoi_site_collection.py:104 */
static double pymod1(double x)
{
    double r = fmod(x, 1.0);
    if (r != 0.0) { if (r < 0.0) r += 1.0; }
    else r = 0.0;
    return r;
}

static double dsign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : ((x == 0.0) ? 0.0 : x)); }

static void counter_inc(Counter *c, int key)
{
    for (int i = 0; i < c->n; i++) if (c->key[i] == key) { c->cnt[i]++; return; }
    if (c->n == c->cap) {
        c->cap = c->cap ? 2 * c->cap : 4;
        c->key = (int *)realloc(c->key, sizeof(int) * c->cap);
        c->cnt = (int64_t *)realloc(c->cnt, sizeof(int64_t) * c->cap);
    }
    c->key[c->n] = key; c->cnt[c->n] = 1; c->n++;
}

/* distances.py:27-66 / :68-110: squared minimum-image distance over 27 images.
 * frac1 = a, frac2 = b; rows of L are lattice vectors. */
static double mic_dist_sq(const double *a, const double *b, const double *L)
{
    double d0b = a[0] - b[0], d1b = a[1] - b[1], d2b = a[2] - b[2];
    d0b -= rint(d0b); d1b -= rint(d1b); d2b -= rint(d2b);
    double best = INFINITY;
    for (int si = -1; si < 2; si++)
        for (int sj = -1; sj < 2; sj++)
            for (int sk = -1; sk < 2; sk++) {
                double d0 = d0b + si, d1 = d1b + sj, d2 = d2b + sk;
                double cx = d0 * L[0] + d1 * L[3] + d2 * L[6];
                double cy = d0 * L[1] + d1 * L[4] + d2 * L[7];
                double cz = d0 * L[2] + d1 * L[5] + d2 * L[8];
                double q = cx * cx + cy * cy + cz * cz;
                if (q < best) best = q;
            }
    return best;
}

double so_mic_distance(const double *a, const double *b, const double *L) { return sqrt(mic_dist_sq(a, b, L)); }

void so_all_mic_distances(const double *c1, int n, const double *c2, int m, const double *L, double *out)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++)
            out[(size_t)i * m + j] = sqrt(mic_dist_sq(c1 + 3 * i, c2 + 3 * j, L));
}

/* pbc_utils.py:16-41 */
static void legacy_pbc(const double *frac, int n, double *out)
{
    memcpy(out, frac, sizeof(double) * 3 * n);
    for (int d = 0; d < 3; d++) {
        double mx = out[d], mn = out[d];
        for (int i = 1; i < n; i++) { double v = out[3 * i + d]; if (v > mx) mx = v; if (v < mn) mn = v; }
        if (mx - mn > 0.5)
            for (int i = 0; i < n; i++) if (out[3 * i + d] < 0.5) out[3 * i + d] += 1.0;
    }
}

/* pbc_utils.py:65-109 (27-image unwrap to a reference centre) and :112-145 (correct_pbc) */
static void correct_pbc(const double *frac, int n, int has_rc, const double *rc, const double *L,
                        double *corrected, int64_t *shifts)
{
    if (!has_rc) {
        legacy_pbc(frac, n, corrected);
        for (int i = 0; i < 3 * n; i++) shifts[i] = (int64_t)rint(corrected[i] - frac[i]);
        return;
    }
    double rcx = rc[0] * L[0] + rc[1] * L[3] + rc[2] * L[6];
    double rcy = rc[0] * L[1] + rc[1] * L[4] + rc[2] * L[7];
    double rcz = rc[0] * L[2] + rc[1] * L[5] + rc[2] * L[8];
    for (int i = 0; i < n; i++) {
        double best = INFINITY; int bs[3] = {0, 0, 0};
        for (int dx = -1; dx < 2; dx++)
            for (int dy = -1; dy < 2; dy++)
                for (int dz = -1; dz < 2; dz++) {
                    double p0 = frac[3 * i] + dx, p1 = frac[3 * i + 1] + dy, p2 = frac[3 * i + 2] + dz;
                    double cx = p0 * L[0] + p1 * L[3] + p2 * L[6] - rcx;
                    double cy = p0 * L[1] + p1 * L[4] + p2 * L[7] - rcy;
                    double cz = p0 * L[2] + p1 * L[5] + p2 * L[8] - rcz;
                    double dist = sqrt(cx * cx + cy * cy + cz * cz);
                    if (dist < best) { best = dist; bs[0] = dx; bs[1] = dy; bs[2] = dz; }
                }
        for (int d = 0; d < 3; d++) { shifts[3 * i + d] = bs[d]; corrected[3 * i + d] = frac[3 * i + d] + (double)bs[d]; }
    }
    for (int d = 0; d < 3; d++) {
        double mn = corrected[d];
        for (int i = 1; i < n; i++) if (corrected[3 * i + d] < mn) mn = corrected[3 * i + d];
        double u = ceil(-mn); if (!(u > 0.0)) u = 0.0;      /* np.maximum(0, np.ceil(-min)) */
        for (int i = 0; i < n; i++) corrected[3 * i + d] = corrected[3 * i + d] + u;
    }
}

/* pbc_utils.py:193-230 (_numba_update_pbc_shifts) */
static int update_pbc_shifts(const double *frac, const double *cached, const int64_t *shifts, int n,
                             double *out, int64_t *new_shifts)
{
    for (int i = 0; i < 3 * n; i++) {
        double diff = frac[i] - cached[i];
        int64_t w = (int64_t)rint(diff);
        double physical = diff - (double)w;
        if (physical >= 0.3 || physical <= -0.3) return 0;
        new_shifts[i] = shifts[i] - w;
    }
    for (int i = 0; i < 3 * n; i++) out[i] = frac[i] + (double)new_shifts[i];
    for (int d = 0; d < 3; d++) {
        double mn = out[d];
        for (int i = 1; i < n; i++) if (out[3 * i + d] < mn) mn = out[3 * i + d];
        double u = 0.0;
        if (mn < 0.0) u = ceil(-mn);
        for (int i = 0; i < n; i++) out[3 * i + d] += u;
    }
    return 1;
}

/* ------------------------------------------------------------------ polyhedral */

/* polyhedral_site.py:377-416 (_store_vertex_coords) */
static void store_vertex_coords(Oracle *o, PolySite *p)
{
    double raw[3 * 64]; double out[3 * 64]; int64_t ns[3 * 64];
    for (int i = 0; i < p->nv; i++)
        for (int d = 0; d < 3; d++) raw[3 * i + d] = o->cur_frac[3 * (size_t)p->vidx[i] + d];
    if (p->has_shift) {
        if (update_pbc_shifts(raw, p->cached_raw, p->shift, p->nv, out, ns)) {
            memcpy(p->shift, ns, sizeof(int64_t) * 3 * p->nv);
            memcpy(p->cached_raw, raw, sizeof(double) * 3 * p->nv);
            memcpy(p->vc, out, sizeof(double) * 3 * p->nv);
            p->has_vc = 1; p->stale = 1; o->n_incr++;
            return;
        }
        o->n_invalid++;
    }
    correct_pbc(raw, p->nv, p->has_rc, p->rc, o->lat, out, ns);
    memcpy(p->shift, ns, sizeof(int64_t) * 3 * p->nv);
    memcpy(p->cached_raw, raw, sizeof(double) * 3 * p->nv);
    memcpy(p->vc, out, sizeof(double) * 3 * p->nv);
    p->has_shift = 1; p->has_vc = 1; p->stale = 1; o->n_full++;
}

/* polyhedral_site.py:26-81 (_numba_update_faces) */
static void update_faces(PolySite *p)
{
    double c[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < p->nv; i++) for (int k = 0; k < 3; k++) c[k] += p->vc[3 * i + k];
    for (int k = 0; k < 3; k++) c[k] /= (double)p->nv;
    for (int j = 0; j < p->nf; j++) {
        const double *v0 = p->vc + 3 * p->simp[3 * j], *v1 = p->vc + 3 * p->simp[3 * j + 1],
                     *v2 = p->vc + 3 * p->simp[3 * j + 2];
        double e0x = v0[0] - v2[0], e0y = v0[1] - v2[1], e0z = v0[2] - v2[2];
        double e1x = v1[0] - v2[0], e1y = v1[1] - v2[1], e1z = v1[2] - v2[2];
        double *n = p->normals + 3 * j, *r = p->refs + 3 * j;
        n[0] = e0y * e1z - e0z * e1y;
        n[1] = e0z * e1x - e0x * e1z;
        n[2] = e0x * e1y - e0y * e1x;
        r[0] = v0[0]; r[1] = v0[1]; r[2] = v0[2];
        double dot = 0.0;
        for (int k = 0; k < 3; k++) dot += n[k] * (c[k] - r[k]);
        p->signs[j] = dsign(dot);
    }
}

/* polyhedral_site.py:83-123 (_numba_sn_query) */
static int sn_query(const PolySite *p, const double *img /* (8,3) */)
{
    for (int i = 0; i < 8; i++) {
        int all_match = 1;
        for (int j = 0; j < p->nf; j++) {
            double dot = 0.0;
            for (int k = 0; k < 3; k++) dot += (img[3 * i + k] - p->refs[3 * j + k]) * p->normals[3 * j + k];
            if (dot != 0.0 && dsign(dot) != p->signs[j]) { all_match = 0; break; }
        }
        if (all_match) return 1;
    }
    return 0;
}

/* polyhedral_site.py:431-481 (contains_point, numba path) */
static int poly_contains(Oracle *o, int s, const double *img)
{
    PolySite *p = &o->ps[s];
    o->n_contains++;
    if (p->pending) { p->pending = 0; store_vertex_coords(o, p); }
    if (!p->has_vc) { o->error = 2; snprintf(o->msg, sizeof o->msg, "no vertex coordinates set for polyhedral_site %d", s); return 0; }
    if (!p->has_topo) {
        int32_t buf[3 * 256];
        int nf = o->hull ? o->hull(o->hull_ctx, s, p->nv, p->vc, 256, buf) : -1;
        if (nf <= 0) { o->error = 3; snprintf(o->msg, sizeof o->msg, "Degenerate vertex geometry for polyhedral_site %d", s); return 0; }
        o->n_hull++;
        p->nf = nf;
        p->simp = (int32_t *)malloc(sizeof(int32_t) * 3 * nf);
        memcpy(p->simp, buf, sizeof(int32_t) * 3 * nf);
        p->normals = (double *)malloc(sizeof(double) * 3 * nf);
        p->refs = (double *)malloc(sizeof(double) * 3 * nf);
        p->signs = (double *)malloc(sizeof(double) * nf);
        p->has_topo = 1;
        update_faces(p); o->n_face_upd++;
        p->stale = 0;
    } else if (p->stale) {
        update_faces(p); o->n_face_upd++;
        p->stale = 0;
    }
    return sn_query(p, img);
}

static int sph_contains(Oracle *o, int s, const double *x)
{
    o->n_contains++;
    /* spherical_site.py:145-146: mic_distance(site, x) <= rcut */
    return sqrt(mic_dist_sq(o->centres + 3 * s, x, o->lat)) <= o->rcut[s];
}

/* ------------------------------------------------------------------ priority order */

/* site_collection.py:36-50 (_NearestSiteLookup.nearest_site_index) */
static int nearest_lookup(Oracle *o, const double *x)
{
    int best = 0; double bd = INFINITY;
    o->n_lookup++;
    for (int s = 0; s < o->S; s++) {
        double d0 = o->look_c[3 * s] - x[0], d1 = o->look_c[3 * s + 1] - x[1], d2 = o->look_c[3 * s + 2] - x[2];
        d0 -= rint(d0); d1 -= rint(d1); d2 -= rint(d2);
        double n = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        if (n < bd) { bd = n; best = s; }
    }
    return best;
}

/* site_collection.py:279-310 (update_occupation) + atom.py:181-192 */
static void update_occupation(Oracle *o, int s, int a)
{
    int prev = o->prev[a];
    if (prev >= 0 && prev != s) counter_inc(&o->tr[prev], s);
    o->in_site[a] = s;
    if (s != o->recent[2 * a]) { o->recent[2 * a + 1] = o->recent[2 * a]; o->recent[2 * a] = s; }
}

static int test_site(Oracle *o, int s, int a, const double *img)
{
    if (o->kind == SO_POLYHEDRAL) return poly_contains(o, s, img);
    return sph_contains(o, s, o->ax + 3 * a);
}

/* site_collection.py:113-182 (_get_priority_sites) driven by
 * polyhedral_site_collection.py:86-107 / spherical_site_collection.py:71-92 */
static void assign_priority(Oracle *o, int a)
{
    double img[24];
    const double *x = o->ax + 3 * a;
    static const int sh[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,1,1},{1,1,1}}; /* tools.py:246-253 */
    for (int i = 0; i < 8; i++) for (int k = 0; k < 3; k++) img[3 * i + k] = (double)sh[i][k] + x[k];
    o->in_site[a] = -1;
    int mk = ++o->cur_mark;
    int anchor = -1;
    int r0 = o->recent[2 * a], r1 = o->recent[2 * a + 1];
    int rec[2], nrec = 0;
    if (r0 >= 0) rec[nrec++] = r0;
    if (r1 >= 0) rec[nrec++] = r1;
#define TRY(s_) do { if (test_site(o, (s_), a, img)) { update_occupation(o, (s_), a); return; } if (o->error) return; } while (0)
    if (nrec) {
        anchor = rec[0];
        for (int i = 0; i < nrec; i++) { TRY(rec[i]); o->mark[rec[i]] = mk; }
    } else if (o->has_rank) {
        anchor = nearest_lookup(o, x);
        TRY(anchor); o->mark[anchor] = mk;
    }
    if (anchor >= 0) {
        /* site.py:216-223: sorted(keys, key=count, reverse=True) -- stable */
        Counter *c = &o->tr[anchor];
        int n = c->n, *ord = o->order_tmp;
        for (int i = 0; i < n; i++) {
            int j = i;
            while (j > 0 && c->cnt[ord[j - 1]] < c->cnt[i]) { ord[j] = ord[j - 1]; j--; }
            ord[j] = i;
        }
        for (int i = 0; i < n; i++) {
            int d = c->key[ord[i]];
            if (o->mark[d] != mk) { TRY(d); o->mark[d] = mk; }
        }
        if (o->has_rank) {
            const int32_t *rk = o->rank + (size_t)anchor * (o->S - 1);
            for (int i = 0; i < o->S - 1; i++) { int d = rk[i]; if (o->mark[d] != mk) { TRY(d); o->mark[d] = mk; } }
        } else {
            for (int i = 0; i < o->n_neigh[anchor]; i++) { int d = o->neigh[anchor][i]; if (o->mark[d] != mk) { TRY(d); o->mark[d] = mk; } }
            for (int s = 0; s < o->S; s++) if (o->mark[s] != mk) TRY(s);
        }
    } else {
        for (int s = 0; s < o->S; s++) TRY(s);
    }
#undef TRY
}

/* ------------------------------------------------------------------ voronoi / dynamic voronoi */

/* voronoi_site_collection.py:76-99: argmin over sqrt'd distances, first minimum wins */
static void assign_nearest(Oracle *o, const double *centres)
{
    for (int a = 0; a < o->A; a++) {
        int best = 0; double bd = INFINITY;
        for (int s = 0; s < o->S; s++) {
            double d = sqrt(mic_dist_sq(centres + 3 * s, o->ax + 3 * a, o->lat));
            if (d < bd) { bd = d; best = s; }
        }
        update_occupation(o, best, a);
    }
}

/* dynamic_voronoi_site_collection.py:72-105 (try_fast_update) and :166-195 */
static void dynvor_centres(Oracle *o)
{
    for (int g = 0; g < o->n_groups; g++) {
        Group *G = &o->groups[g];
        int n = G->n, R = G->R;
        double *b = G->batch;
        for (int i = 0; i < n * R; i++)
            for (int d = 0; d < 3; d++) b[3 * i + d] = o->cur_frac[3 * (size_t)G->ref_idx[i] + d];
        int fast = G->initialised;
        if (fast) {
            for (int i = 0; i < 3 * n * R; i++) {
                double diff = b[i] - G->cached[i];
                double w = rint(diff);
                double physical = diff - (double)(int64_t)w;
                if (!(fabs(physical) < 0.3)) { fast = 0; break; }
            }
            if (!fast) o->n_invalid++;
        }
        if (fast) {
            o->n_incr++;
            for (int i = 0; i < 3 * n * R; i++) {
                double diff = b[i] - G->cached[i];
                G->shift[i] = G->shift[i] - (int64_t)rint(diff);
            }
            for (int s = 0; s < n; s++) {
                double *c = o->centres + 3 * G->site_pos[s];
                for (int d = 0; d < 3; d++) {
                    double mn = INFINITY;
                    for (int r = 0; r < R; r++) {
                        double v = b[3 * (s * R + r) + d] + (double)G->shift[3 * (s * R + r) + d];
                        if (v < mn) mn = v;
                    }
                    double nn = ceil(-mn); if (!(nn > 0.0)) nn = 0.0;
                    double sum = 0.0;
                    for (int r = 0; r < R; r++) {
                        double v = (b[3 * (s * R + r) + d] + (double)G->shift[3 * (s * R + r) + d]) + nn;
                        sum = (r == 0) ? v : sum + v;
                    }
                    c[d] = pymod1(sum / (double)R);
                }
            }
            memcpy(G->cached, b, sizeof(double) * 3 * n * R);
            continue;
        }
        o->n_full++;
        for (int s = 0; s < n; s++) {
            double corr[3 * 64]; int64_t sh[3 * 64];
            int sp = G->site_pos[s];
            correct_pbc(b + 3 * s * R, R, o->dv_has_rc[sp], o->dv_rc + 3 * sp, o->lat, corr, sh);
            for (int d = 0; d < 3; d++) {
                double sum = corr[d];
                for (int r = 1; r < R; r++) sum += corr[3 * r + d];
                o->centres[3 * sp + d] = pymod1(sum / (double)R);
            }
            memcpy(G->shift + 3 * s * R, sh, sizeof(int64_t) * 3 * R);
        }
        memcpy(G->cached, b, sizeof(double) * 3 * n * R);
        G->initialised = 1;
    }
    o->dv_centres_set = 1;
}

/* ------------------------------------------------------------------ public API */

Oracle *so_create(int kind, int S, int A, const int32_t *atom_index)
{
    Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
    o->kind = kind; o->S = S; o->A = A;
    o->atom_index = (int *)malloc(sizeof(int) * A);
    for (int a = 0; a < A; a++) o->atom_index[a] = atom_index[a];
    o->ax = (double *)calloc(3 * (size_t)A, sizeof(double));
    o->in_site = (int *)malloc(sizeof(int) * A);
    o->prev = (int *)malloc(sizeof(int) * A);
    o->recent = (int *)malloc(sizeof(int) * 2 * A);
    o->tr = (Counter *)calloc(S, sizeof(Counter));
    o->mark = (int *)calloc(S, sizeof(int));
    o->order_tmp = (int *)malloc(sizeof(int) * (S + 1));
    for (int a = 0; a < A; a++) { o->in_site[a] = -1; o->prev[a] = -2; o->recent[2 * a] = o->recent[2 * a + 1] = -1; }
    return o;
}

void so_set_centres(Oracle *o, const double *centres, const double *rcut)
{
    o->centres = (double *)malloc(sizeof(double) * 3 * o->S);
    memcpy(o->centres, centres, sizeof(double) * 3 * o->S);
    if (rcut) { o->rcut = (double *)malloc(sizeof(double) * o->S); memcpy(o->rcut, rcut, sizeof(double) * o->S); }
}

/* site_collection.py:90-111: the rank table is the reference's own np.argsort output,
 * computed by the Python wrapper with the same numpy call. */
void so_set_ranking(Oracle *o, const int32_t *rank, const double *lookup_centres)
{
    o->has_rank = 1;
    o->rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)o->S * (o->S > 1 ? o->S - 1 : 1));
    memcpy(o->rank, rank, sizeof(int32_t) * (size_t)o->S * (o->S - 1));
    o->look_c = (double *)malloc(sizeof(double) * 3 * o->S);
    memcpy(o->look_c, lookup_centres, sizeof(double) * 3 * o->S);
}

/* polyhedral_site_collection.py:165-191 result, flattened (CSR) */
void so_set_neighbours(Oracle *o, const int32_t *offsets, const int32_t *flat)
{
    o->neigh = (int **)calloc(o->S, sizeof(int *));
    o->n_neigh = (int *)calloc(o->S, sizeof(int));
    for (int s = 0; s < o->S; s++) {
        int n = offsets[s + 1] - offsets[s];
        o->n_neigh[s] = n;
        o->neigh[s] = (int *)malloc(sizeof(int) * (n ? n : 1));
        for (int i = 0; i < n; i++) o->neigh[s][i] = flat[offsets[s] + i];
    }
}

void so_set_polyhedra(Oracle *o, int vmax, const int32_t *n_vert, const int32_t *vert_idx,
                      const uint8_t *has_rc, const double *rc,
                      const uint8_t *primed, const int64_t *shift0, const double *raw0)
{
    o->ps = (PolySite *)calloc(o->S, sizeof(PolySite));
    for (int s = 0; s < o->S; s++) {
        PolySite *p = &o->ps[s];
        p->nv = n_vert[s];
        p->vidx = (int *)malloc(sizeof(int) * p->nv);
        for (int i = 0; i < p->nv; i++) p->vidx[i] = vert_idx[(size_t)s * vmax + i];
        p->has_rc = has_rc[s];
        for (int d = 0; d < 3; d++) p->rc[d] = rc[3 * s + d];
        p->shift = (int64_t *)calloc(3 * p->nv, sizeof(int64_t));
        p->cached_raw = (double *)calloc(3 * p->nv, sizeof(double));
        p->vc = (double *)calloc(3 * p->nv, sizeof(double));
        p->stale = 1;
        if (primed && primed[s]) {
            p->has_shift = 1;
            for (int i = 0; i < 3 * p->nv; i++) {
                p->shift[i] = shift0[(size_t)s * vmax * 3 + i];
                p->cached_raw[i] = raw0[(size_t)s * vmax * 3 + i];
            }
        }
    }
}

/* a site that already owns a FaceTopologyCache (polyhedral_site.py:468-471 skipped) */
void so_set_topology(Oracle *o, int s, int nf, const int32_t *simp)
{
    PolySite *p = &o->ps[s];
    free(p->simp); free(p->normals); free(p->refs); free(p->signs);
    p->nf = nf;
    p->simp = (int32_t *)malloc(sizeof(int32_t) * 3 * nf);
    memcpy(p->simp, simp, sizeof(int32_t) * 3 * nf);
    p->normals = (double *)malloc(sizeof(double) * 3 * nf);
    p->refs = (double *)malloc(sizeof(double) * 3 * nf);
    p->signs = (double *)malloc(sizeof(double) * nf);
    p->has_topo = 1; p->stale = 1;
}

void so_set_hull_callback(Oracle *o, so_hull_fn fn, void *ctx) { o->hull = fn; o->hull_ctx = ctx; }

/* dynamic_voronoi_site_collection.py:151-164: groups by n_reference in first-seen order */
void so_set_dynvor(Oracle *o, int rmax, const int32_t *n_ref, const int32_t *ref_idx,
                   const uint8_t *has_rc, const double *rc)
{
    int S = o->S;
    o->centres = (double *)calloc(3 * (size_t)S, sizeof(double));
    o->dv_has_rc = (int *)malloc(sizeof(int) * S);
    o->dv_rc = (double *)malloc(sizeof(double) * 3 * S);
    for (int s = 0; s < S; s++) { o->dv_has_rc[s] = has_rc[s]; for (int d = 0; d < 3; d++) o->dv_rc[3 * s + d] = rc[3 * s + d]; }
    int *seen = (int *)malloc(sizeof(int) * S); int ng = 0;
    for (int s = 0; s < S; s++) {
        int f = 0; for (int g = 0; g < ng; g++) if (seen[g] == n_ref[s]) f = 1;
        if (!f) seen[ng++] = n_ref[s];
    }
    o->n_groups = ng; o->groups = (Group *)calloc(ng, sizeof(Group));
    for (int g = 0; g < ng; g++) {
        Group *G = &o->groups[g]; G->R = seen[g];
        for (int s = 0; s < S; s++) if (n_ref[s] == G->R) G->n++;
        G->site_pos = (int *)malloc(sizeof(int) * G->n);
        G->ref_idx = (int *)malloc(sizeof(int) * G->n * G->R);
        G->shift = (int64_t *)calloc(3 * (size_t)G->n * G->R, sizeof(int64_t));
        G->cached = (double *)calloc(3 * (size_t)G->n * G->R, sizeof(double));
        G->batch = (double *)calloc(3 * (size_t)G->n * G->R, sizeof(double));
        int k = 0;
        for (int s = 0; s < S; s++) if (n_ref[s] == G->R) {
            G->site_pos[k] = s;
            for (int r = 0; r < G->R; r++) G->ref_idx[k * G->R + r] = ref_idx[(size_t)s * rmax + r];
            k++;
        }
    }
    free(seen);
}

/* One frame: <collection>.analyse_structure(atoms, structure)
 * (polyhedral_site_collection.py:66-84, voronoi :52-74, dynamic :203-220, spherical :53-69) */
int so_analyse(Oracle *o, const double *frac, const double *lattice)
{
    o->cur_frac = frac;
    memcpy(o->lat, lattice, sizeof(double) * 9);
    for (int a = 0; a < o->A; a++)
        for (int d = 0; d < 3; d++) o->ax[3 * a + d] = pymod1(frac[3 * (size_t)o->atom_index[a] + d]);
    switch (o->kind) {
    case SO_POLYHEDRAL:
        for (int s = 0; s < o->S; s++) o->ps[s].pending = 1;
        for (int a = 0; a < o->A && !o->error; a++) assign_priority(o, a);
        break;
    case SO_SPHERICAL:
        for (int a = 0; a < o->A; a++) assign_priority(o, a);
        break;
    case SO_VORONOI:
        assign_nearest(o, o->centres);
        break;
    case SO_DYNVOR:
        dynvor_centres(o);
        assign_nearest(o, o->centres);
        break;
    }
    return o->error;
}

/* trajectory.py:342-362 (append_timestep bookkeeping after analyse_structure) */
void so_append(Oracle *o) { for (int a = 0; a < o->A; a++) o->prev[a] = o->in_site[a]; o->n_traj++; }

/* trajectory.py:413-432 over an array trajectory; lattice_stride = 0 (one lattice) or 9 */
int so_run(Oracle *o, const double *frac, int64_t n_frames, int64_t n_atoms_total,
           const double *lattice, int lattice_stride, int32_t *out, int append)
{
    for (int64_t f = 0; f < n_frames; f++) {
        int e = so_analyse(o, frac + (size_t)f * n_atoms_total * 3, lattice + (size_t)f * lattice_stride);
        if (e) return e;
        if (append) so_append(o);
        if (out) for (int a = 0; a < o->A; a++) out[(size_t)f * o->A + a] = o->in_site[a];
    }
    o->cur_frac = NULL;
    return 0;
}

/* trajectory.py:364-373 -> atom.py:81-93, site.py:75-88, polyhedral_site.py:264-279,
 * dynamic_voronoi_site_collection.py:197-201.  Face topology survives. */
void so_reset(Oracle *o)
{
    for (int a = 0; a < o->A; a++) { o->in_site[a] = -1; o->prev[a] = -2; o->recent[2 * a] = o->recent[2 * a + 1] = -1; }
    o->n_traj = 0;
    for (int s = 0; s < o->S; s++) o->tr[s].n = 0;
    if (o->kind == SO_POLYHEDRAL)
        for (int s = 0; s < o->S; s++) { PolySite *p = &o->ps[s]; p->has_vc = 0; p->stale = 1; p->pending = 0; p->has_shift = 0; }
    if (o->kind == SO_DYNVOR) { for (int g = 0; g < o->n_groups; g++) o->groups[g].initialised = 0; o->dv_centres_set = 0; }
    o->error = 0;
}

const char *so_error_message(Oracle *o) { return o->msg; }
void so_get_in_site(Oracle *o, int32_t *out) { for (int a = 0; a < o->A; a++) out[a] = o->in_site[a]; }
void so_get_recent(Oracle *o, int32_t *out) { for (int a = 0; a < 2 * o->A; a++) out[a] = o->recent[a]; }
void so_get_centres(Oracle *o, double *out) { memcpy(out, o->centres, sizeof(double) * 3 * o->S); }
int so_transition_count(Oracle *o, int s) { return o->tr[s].n; }
void so_get_transitions(Oracle *o, int s, int32_t *keys, int64_t *counts)
{ for (int i = 0; i < o->tr[s].n; i++) { keys[i] = o->tr[s].key[i]; counts[i] = o->tr[s].cnt[i]; } }
int so_get_topology(Oracle *o, int s, int32_t *simp)
{
    PolySite *p = &o->ps[s];
    if (!p->has_topo) return -1;
    if (simp) memcpy(simp, p->simp, sizeof(int32_t) * 3 * p->nf);
    return p->nf;
}
int so_get_vertex_coords(Oracle *o, int s, double *vc, int64_t *shift)
{
    PolySite *p = &o->ps[s];
    if (!p->has_vc) return 0;
    memcpy(vc, p->vc, sizeof(double) * 3 * p->nv);
    if (shift) memcpy(shift, p->shift, sizeof(int64_t) * 3 * p->nv);
    return p->nv;
}
void so_get_stats(Oracle *o, int64_t *out)
{
    out[0] = o->n_contains; out[1] = o->n_full; out[2] = o->n_incr; out[3] = o->n_invalid;
    out[4] = o->n_face_upd; out[5] = o->n_hull; out[6] = o->n_lookup; out[7] = o->n_traj;
}

void so_destroy(Oracle *o)
{
    if (!o) return;
    for (int s = 0; s < o->S; s++) { free(o->tr[s].key); free(o->tr[s].cnt); }
    if (o->ps) for (int s = 0; s < o->S; s++) {
        PolySite *p = &o->ps[s];
        free(p->vidx); free(p->shift); free(p->cached_raw); free(p->vc);
        free(p->simp); free(p->normals); free(p->refs); free(p->signs);
    }
    if (o->neigh) { for (int s = 0; s < o->S; s++) free(o->neigh[s]); free(o->neigh); free(o->n_neigh); }
    for (int g = 0; g < o->n_groups; g++) {
        Group *G = &o->groups[g];
        free(G->site_pos); free(G->ref_idx); free(G->shift); free(G->cached); free(G->batch);
    }
    free(o->groups); free(o->dv_has_rc); free(o->dv_rc);
    free(o->ps); free(o->atom_index); free(o->ax); free(o->in_site); free(o->prev); free(o->recent);
    free(o->tr); free(o->mark); free(o->order_tmp); free(o->rank); free(o->look_c);
    free(o->centres); free(o->rcut);
    free(o);
}
