/*
 * sab200.h -- C ABI of libsab200.so, the B200 (sm_100a) backend for the per-frame
 * site-assignment hot path of bjmorgan/site-analysis.
 *
 * Plain pointers and sizes only; no torch / pybind types.  Every device pointer is
 * caller-owned (the Python host allocates through torch); `stream` is a cudaStream_t
 * passed as void*.  The library allocates device memory only for the small static
 * site tables uploaded by the sab_set_* calls.  All functions return SAB_OK (0) or
 * an error code; sab_last_error() gives the message of the calling thread's last
 * failure.  There is no CPU fallback anywhere behind this interface.
 *
 * The reference (pure Python + numba) has no FFI of its own: its only seam is the
 * SiteCollection ABC (reference site_analysis/site_collection.py:185-245) selected by
 * the hard-coded map in trajectory.py:86-91.  Each entry point below names the
 * reference code it replaces; INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add at that seam.
 *
 * Conventions
 *   frac      double[F][N][3]  fractional coordinates of every atom in every frame,
 *                              exactly `structure.frac_coords` (NOT wrapped)
 *   lattice   double[3][3] (lattice_stride = 0) or double[F][3][3] (stride = 9),
 *                              rows are lattice vectors (pymatgen `lattice.matrix`)
 *   result    int32[F][A]      site LIST POSITION of mobile atom a in frame f,
 *                              SAB_NONE (-1) = no site, <= SAB_AMBIGUOUS_BASE (-2-k) =
 *                              several sites contain the atom: spill[k] = n followed by
 *                              the n candidate positions; sab_resolve() settles these
 *                              with the reference's priority rules
 */
#ifndef SAB200_H
#define SAB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAB_ABI_VERSION 1

enum { SAB_SPHERICAL = 0, SAB_VORONOI = 1, SAB_POLYHEDRAL = 2, SAB_DYNAMIC_VORONOI = 3 };
enum { SAB_OK = 0, SAB_EINVAL = 1, SAB_ECUDA = 2, SAB_EOVERFLOW = 3, SAB_ESTATE = 4 };

#define SAB_NONE (-1)
#define SAB_AMBIGUOUS_BASE (-2)

typedef struct sab_engine sab_engine;

int sab_abi_version(void);
const char *sab_last_error(void);

/* One engine = one SiteCollection (reference site_collection.py:185) bound to one GPU.
 * mobile_idx[A]: structure indices of the tracked atoms (reference Atom.index, atom.py:25). */
int sab_engine_create(sab_engine **out, int kind, int device, int n_atoms, int n_mobile,
                      const int32_t *mobile_idx, int n_sites);
void sab_engine_destroy(sab_engine *e);

/* ---- static site tables (host pointers; copied to the device) --------------------- */

/* SphericalSite.frac_coords / .rcut (spherical_site.py:40-58), VoronoiSite.frac_coords
 * (voronoi_site.py:36-49).  rcut_sq_max[s] = largest double q with sqrt(q) <= rcut[s]
 * (so `q <= rcut_sq_max` == the reference's `sqrt(q) <= rcut`, spherical_site.py:145-146);
 * both NULL for Voronoi. */
int sab_set_centres(sab_engine *e, const double *centres, const double *rcut,
                    const double *rcut_sq_max);

/* PolyhedralSite.vertex_indices (polyhedral_site.py:218-236) padded to vmax, and the
 * Qhull face topology FaceTopologyCache.face_simplices (polyhedral_site.py:154-155)
 * padded to fmax (local vertex numbers). */
int sab_set_polyhedra(sab_engine *e, int vmax, const int32_t *n_vert, const int32_t *vert_idx,
                      int fmax, const int32_t *n_face, const int32_t *simplices);

/* DynamicVoronoiSite.reference_indices (dynamic_voronoi_site.py:62-78) padded to rmax. */
int sab_set_reference_atoms(sab_engine *e, int rmax, const int32_t *n_ref, const int32_t *ref_idx);

/* Number M and list of the distinct structure atoms referenced by vertices / reference
 * slots ("framework atoms"); compact numbering used by the PBC state below. */
int sab_framework_count(sab_engine *e);
int sab_framework_atoms(sab_engine *e, int32_t *out);

/* ---- PBC image-shift state ---------------------------------------------------------
 * Replaces the per-site caches PolyhedralSite._pbc_image_shifts/_pbc_cached_raw_frac
 * (polyhedral_site.py:377-416) and _CentreGroup.pbc_shifts/cached_raw_frac
 * (dynamic_voronoi_site_collection.py:38-70).  The engine keeps, per framework atom m,
 * the last raw coordinate prev_raw[m] and the cumulative wrap count W[m]
 * (sum of rint(raw_t - raw_{t-1}), pbc_utils.py:210-217); the shift of slot (s, v) at a
 * frame is slot_shift[s][v] - W[atom(s, v)].
 *   slot_shift int32[S][vmax|rmax][3], prev_raw double[M][3], wraps int32[M][3],
 *   excursion double[M][3][2] = running (min, max) of raw - W per atom and dimension. */
int sab_set_pbc_state(sab_engine *e, const int32_t *slot_shift, const double *prev_raw,
                      const int32_t *wraps, const double *excursion);
int sab_get_pbc_state(sab_engine *e, int32_t *slot_shift, double *prev_raw, int32_t *wraps,
                      double *excursion);

/* Multi-GPU shard carry, all on the device and stream-ordered (no host round trip).  d_halo_raw double[M][3]: raw
 * framework coordinates of the frame BEFORE this shard (the previous rank's last frame); d_anchor0 double[M][3]: the
 * unwrapped coordinates raw - W of the trajectory's first frame.  Sets prev_raw = halo, W = rint(halo - anchor0) and
 * excursion = (u, u).  Exact while the excursion guard holds over the whole trajectory: the sum
 * of the earlier shards' wrap totals (pbc_utils.py:210-217 applied frame by frame) equals rint(halo - anchor0).
 * sab_get_excursion_device copies the running excursion double[M][3][2] to a device buffer for the cross-rank guard. */
int sab_set_carry_device(sab_engine *e, const double *d_halo_raw, const double *d_anchor0, void *stream);
int sab_get_excursion_device(sab_engine *e, double *d_out, void *stream);

/* Sites whose PBC state was initialised by the legacy spread rule (reference_center is None,
 * pbc_utils.py:16-41 via correct_pbc :140-145) get NO non-negative offset in the frame of
 * initialisation; site_flags uint8[S] marks them and frames_from_now is that frame's index
 * counted from the next frame the engine will process.  NULL clears the marker. */
int sab_set_legacy_init(sab_engine *e, const uint8_t *site_flags, int64_t frames_from_now);

/* Exact-pruning tables of the tetrahedral kernel (csrc/k_polyhedral_cells.cuh), built once on
 * the host: a G^3 cell list (CSR) of the tetrahedra that can reach each cell while every
 * framework coordinate stays within `delta` of anchor[3M] (anchor of raw - W); per-site records
 * of shared-memory offsets (uint16[S][20]: 12 byte offsets 8 * ((shift - kmin) * 3M + 3 fw + d) of the
 * vertices in slot order, then 12 face-corner bytes in Qhull's order); kmin/nk = range of slot-shift
 * values.  The lists must be conservative by at least 1e-6 fractional units (the default kernel looks
 * the cell of an atom up in float32).  Frames that leave the delta tube take the exhaustive kernel
 * path, so the tables never decide a result on their own.  anchor == NULL clears the tables. */
int sab_set_cell_lists(sab_engine *e, int G, double delta, const double *anchor, int kmin, int nk,
                       const uint16_t *records, const int32_t *cell_off, const uint16_t *cell_sites);

/* Exact-pruning tables of the nearest-centre (Voronoi) kernel (csrc/k_nearest_cells.cuh): a G^3 cell
 * list of the centres that can be nearest to a point of each cell, excl[c] = lower bound on the
 * distance from cell c to any centre not in its list, both in the metric of the lattice whose
 * INVERSE is passed.  Atoms whose best listed centre is not provably inside that bound in the
 * frame's own metric are redone against all centres.  cell_off == NULL clears the tables. */
int sab_set_nearest_lists(sab_engine *e, int G, const double *lattice0_inverse, const int32_t *cell_off,
                          const uint16_t *cell_sites, const double *excl);

/* Exact-pruning table of the spherical kernel (csrc/k_spherical.cuh, k_spherical_assign_cells): a G^3 cell list of the
 * sites whose sphere inflated by `rho` (>= 1) can reach each cell in the metric of the lattice whose INVERSE is passed,
 * over the periodic images {-2..2}^3.  Frames whose lattice is strained by more than the head-room rho allows take the
 * exhaustive loop, so the table never decides a result on its own (replaces the site loop of
 * spherical_site_collection.py:86-91 only as a candidate filter).  cell_off == NULL clears the table. */
int sab_set_sphere_lists(sab_engine *e, int G, const double *lattice0_inverse, double rho, const int32_t *cell_off,
                         const uint16_t *cell_sites);

/* The same for dynamic-Voronoi sites: lists built around ANCHOR centres double[S][3] with head-room
 * max_move (metric of lattice0) for the centres' motion; frames in which a centre is farther than
 * max_move from its anchor take the exhaustive path. */
int sab_set_dynvor_lists(sab_engine *e, int G, const double *lattice0, const double *lattice0_inverse,
                         const double *anchors, double max_move, const int32_t *cell_off,
                         const uint16_t *cell_sites, const double *excl);

/* Range (min, max) of raw - W per framework coordinate over the frames, double[M][3][2] (host),
 * continuing from the engine's wrap state without modifying it; used to place the anchors. */
int sab_scan_excursion(sab_engine *e, const double *d_frac, int64_t n_frames, void *d_workspace,
                       int64_t workspace_bytes, void *stream, double *excursion);

/* Engine options.  SAB_OPT_SKIP_STEP_CHECK_ONCE: the next batch does not test its first frame
 * against the 0.3 step bound (used right after the host re-initialised a group at that frame,
 * dynamic_voronoi_site_collection.py:188-195). */
#define SAB_OPT_SKIP_STEP_CHECK_ONCE 1
/* SAB_OPT_POLY_KERNEL (tetrahedral sites with cell lists set): 0 = stream kernel (default): warp-specialised frame
 * pipeline, "previous site first" with an FP32 pre-test and the PBC wrap scan fused in; 5 = tile kernel (same tests over
 * tiles of frames after the separate wrap-scan kernels); 4 = per-frame chunk-sequential float64 kernel; 1 = frame-parallel
 * cell-list kernel; 3 = experimental prefetch variant.  All give identical results (cross-check twins).
 * SAB_OPT_SEQ_CHUNK: consecutive frames per CTA chunk (default: automatic). */
#define SAB_OPT_POLY_KERNEL 2
#define SAB_OPT_SEQ_CHUNK 3
#define SAB_OPT_SEQ_BLOCK 4   /* threads per CTA of the chunk-sequential kernel (tuning) */
/* 1 (default): tetrahedral tests run an FP32 pre-test first and fall back to the exact float64 test whenever the
 * point is within the stated error band of a face (results identical either way); 0: float64 only. */
#define SAB_OPT_FP32_FILTER 5
#define SAB_OPT_TILE_FRAMES 6 /* frames per tile of the tile kernel, 1..4 (0 = automatic) */
#define SAB_OPT_STREAM_PRODUCERS 7 /* producer warps per CTA of the stream kernel (0 = automatic) */
#define SAB_OPT_STREAM_REC_SMEM 8  /* 1 (default): site records staged in shared memory when they fit; 0: read through L1 */
#define SAB_OPT_STREAM_REG_COLS 9  /* 1 (default): producer threads keep their columns' metadata and excursions in registers when they fit */
#define SAB_OPT_STREAM_GEN 10      /* 0 (default): generation-5 stream kernel (vertex-image rows) when its tables exist; 4: the previous one */
#define SAB_OPT_STREAM5_ROWS 11    /* generation 5: row slots of the ring, 2..8 (0 = automatic) */
#define SAB_OPT_STREAM5_RAW 12     /* generation 5: raw-frame slots of the ring, 2..4 (0 = automatic) */
#define SAB_OPT_STREAM5_BLOCKS 13  /* generation 5: resident CTAs per SM the shared-memory plan aims for, 1..8 (0 = automatic) */
#define SAB_OPT_RESOLVER 15        /* priority resolver (sab_resolve): 0 (default) = speculative chunks of frames (k_resolve_spec.cuh), 1 = one
                                    * barrier per frame (k_resolve_fast.cuh), 2 = global-memory history (k_resolve.cuh); identical results */
int sab_set_option(sab_engine *e, int key, int64_t value);

/* ---- the hot path -------------------------------------------------------------------- */

/* Bytes of device workspace sab_assign_device needs for n_frames frames. */
int64_t sab_workspace_bytes(sab_engine *e, int64_t n_frames);

/* Pass 1 of the PBC state (polyhedral / dynamic Voronoi only): sums the wraps of every
 * framework atom over the frames and returns them (host int32[M][3]) together with the
 * raw coordinates of the last frame (host double[M][3]); engine state is NOT modified.
 * Used by multi-GPU shards to build the exclusive scan of wrap totals across ranks. */
int sab_wrap_totals(sab_engine *e, const double *d_frac, int64_t n_frames, void *d_workspace,
                    int64_t workspace_bytes, void *stream, int32_t *totals, double *last_raw);

/* Assign every mobile atom of every frame; replaces <X>SiteCollection.analyse_structure /
 * assign_site_occupations (polyhedral_site_collection.py:66-107, voronoi_site_collection.py:
 * 52-99, dynamic_voronoi_site_collection.py:203-244, spherical_site_collection.py:53-92)
 * for a whole trajectory (trajectory.py:413-432).  All d_* pointers are device memory.
 *   d_spill / spill_capacity (int32 elements): candidate lists of ambiguous atom-frames.
 *   info[12] (host, written on return): [0] ambiguous atom-frames, [1] spill elements used,
 *   [2] spill overflow (0/1), [3] first frame whose step exceeded the 0.3 cache-validity
 *   bound (pbc_utils.py:216) or -1, [4] frames that took the exhaustive (non-pruned) path,
 *   [5] kernels launched, [6] device time of the assignment kernel in ns, [7] device time of
 *   the PBC wrap-scan kernels in ns (both from CUDA events on `stream`), [8] bit pattern of the
 *   largest excursion span max(raw - W) - min(raw - W) over all framework coordinates so far (double;
 *   guard of the scan formulation), [9] atom-frames that the float32 pre-test of the
 *   tetrahedral kernel could not decide and that were settled by its exact float64 pass, [10] 1 when the stream
 *   kernel's wrap-count chain check failed and the batch was redone with the scan kernels, [11] 1 when every result
 *   row was also stored into the peers registered with sab_set_result_peers.
 * When info[3] < 0 the PBC state advances to the end of the batch.  When info[3] = t >= 0 the
 * results of frames < t are valid, the PBC state is left untouched, and the caller commits
 * the valid prefix with sab_commit_pbc(..., n_commit = t) (same workspace, unmodified). */
int sab_assign_device(sab_engine *e, const double *d_frac, const double *d_lattice,
                      int lattice_stride, int64_t n_frames, int32_t *d_result,
                      int32_t *d_spill, int64_t spill_capacity, void *d_workspace,
                      int64_t workspace_bytes, void *stream, int64_t *info);

int sab_commit_pbc(sab_engine *e, const double *d_frac, int64_t n_frames, int64_t n_commit,
                   void *d_workspace, int64_t workspace_bytes, void *stream);

/* Same, from HOST buffers (pinned or pageable): chunks the trajectory, overlaps H2D of
 * chunk k+1 with the kernels of chunk k on two internal streams and copies the result
 * back.  This is the call a reference-side plugin would make per trajectory. */
int sab_assign_host(sab_engine *e, const double *h_frac, const double *h_lattice,
                    int lattice_stride, int64_t n_frames, int32_t *h_result,
                    int64_t chunk_frames, int64_t *info);

/* Optional device mirror of the rows the next sab_assign_host calls return: row f of the call also lands in
 * d_full[f * n_mobile ...] (device memory of at least n_frames * n_mobile int32), so that a caller that goes on working
 * on the GPU (the history fold, sab_fold_result) need not upload the result again.  NULL switches it off. */
int sab_set_result_mirror(sab_engine *e, int32_t *d_full);

/* Lazy distance ranking (large S): sab_set_priority_tables(e, NULL, lookup, ...) declares that a ranking exists without
 * uploading the dense S x (S - 1) table of site_collection.py:90-111.  The resolver then ranks the few candidates of an
 * ambiguous atom-frame by their fractional minimum-image distance to the anchor (numpy's operation order); only when two
 * candidates are exactly equidistant does the order depend on np.argsort's implementation: the anchor is reported by
 * sab_rank_missing (returns the number of distinct anchors, writes up to `capacity`, clears the list), the caller computes
 * those rows with the reference's own numpy call, hands them over with sab_add_rank_rows (rows int32[n_rows][S-1], ranked
 * order) and redoes the batch.  Cached rows persist for the engine's lifetime. */
int sab_add_rank_rows(sab_engine *e, int n_rows, const int32_t *anchors, const int32_t *rows);
int sab_rank_missing(sab_engine *e, int32_t *anchors, int capacity);

/* ---- priority resolver --------------------------------------------------------------
 * Replays PriorityAssignmentMixin._get_priority_sites (site_collection.py:113-182),
 * SiteCollection.update_occupation (:279-310) and Atom.update_recent_site (atom.py:181-192)
 * over the batch, in frame order and atom order, settling every ambiguous entry of
 * d_result in place.  Tables (host pointers, may be NULL when absent):
 *   rank int32[S][S-1]   reference _distance_ranked_sites as list positions
 *   lookup double[S][3]  centres of _NearestSiteLookup
 *   neigh_off int32[S+1], neigh int32[*]  face-sharing neighbours (polyhedral_site_
 *                        collection.py:165-191), used when rank == NULL
 * History state (host, in/out): recent int32[A][2], prev int32[A] (-2 = empty trajectory,
 *   -1 = None), and the per-site transition Counters (site.py:68-71) as dense rows in Counter
 *   insertion order: tr_n int32[S], tr_key int32[S][tr_width], tr_cnt int64[S][tr_width].
 *   SAB_EOVERFLOW if a site collects more than tr_width distinct destinations. */
int sab_set_priority_tables(sab_engine *e, const int32_t *rank, const double *lookup,
                            const int32_t *neigh_off, const int32_t *neigh);
int sab_resolve(sab_engine *e, const double *d_frac, int64_t n_frames, int32_t *d_result,
                const int32_t *d_spill, void *stream, int32_t *recent, int32_t *prev,
                int32_t *tr_n, int32_t *tr_key, int64_t *tr_cnt, int tr_width);

/* Dynamic-Voronoi centres of one frame range, double[F][S][3] on the device (bitwise
 * comparable with DynamicVoronoiSite.centre, dynamic_voronoi_site_collection.py:104).
 * Does not advance engine state. */
int sab_dynvor_centres(sab_engine *e, const double *d_frac, int64_t n_frames, double *d_centres,
                       void *d_workspace, int64_t workspace_bytes, void *stream);

/* ---- folds over a finished result array -----------------------
 * One pass over d_result int32[F][A] (site list positions, SAB_NONE = no site; ambiguous entries must have been
 * resolved) replaces the per-object Python lists the reference walks afterwards:
 *   runs   maximal stretches of consecutive frames in which one atom stays in one site, sorted by
 *          (site, atom, start): d_run_site int32[R], d_run_atom int32[R] (column of d_result), d_run_start int64[R],
 *          d_run_length int64[R].  This is the run-length form of every Site.trajectory (trajectory.py:357-360),
 *          from which Site.residence_times (site.py:241-325, gap filter :327-352) and Site.average_occupation
 *          (site.py:225-239) follow without touching the frames again.
 *   pairs  every (previous site -> site) change between consecutive frames of one atom, i.e. the increments of
 *          Site.transitions (site_collection.py:300-302): d_pair_src / d_pair_dst int32[P], d_pair_count int64[P],
 *          d_pair_first int64[P] = smallest frame * A + atom of the pair (Counter insertion order), unordered.
 *          d_prev int32[A] (device, may be NULL = empty history): the atoms' previous trajectory entries
 *          (-2 = none, SAB_NONE, or a site position), used for frame 0.  table_slots (power of two, 0 = skip the
 *          pairs) sizes the open-addressing hash table in the workspace.
 *   recent d_recent_last / d_recent_other int64[A] (device, both or neither; need run_capacity >= R): per atom
 *          ((end frame of the run) << 24) | site of its last run and of its last run in a different site, 0 = none --
 *          what the batch contributes to Atom._recent_sites (atom.py:181-192).
 * info[6] (host): [0] runs R, [1] pairs P, [2] kernels launched, [3] device time in ns, [4] error bits, [5] reserved.
 * SAB_EOVERFLOW when R > run_capacity or P > pair_capacity (info[0], info[1] are valid: call again with room) or when
 * the hash table is full; SAB_ESTATE for unresolved entries.  All d_* pointers are caller-owned device memory. */
int64_t sab_fold_workspace_bytes(int64_t n_frames, int n_mobile, int n_sites, int64_t run_capacity, int64_t table_slots);
int sab_fold_result(int device, const int32_t *d_result, int64_t n_frames, int n_mobile, int n_sites,
                    const int32_t *d_prev, int64_t run_capacity, int32_t *d_run_site, int32_t *d_run_atom,
                    int64_t *d_run_start, int64_t *d_run_length, int64_t table_slots, int64_t pair_capacity,
                    int32_t *d_pair_src, int32_t *d_pair_dst, int64_t *d_pair_count, int64_t *d_pair_first,
                    int64_t *d_recent_last, int64_t *d_recent_other,
                    void *d_workspace, int64_t workspace_bytes, void *stream, int64_t *info);

/* ---- minimum-image distance matrix (stand-alone operator) -----------------------------------
 * d_out double[n][m] = all_mic_distances(c1, c2, lattice) of the reference (distances.py:147-189, numba kernel
 * :68-110; one pair = mic_distance :113-144): 27 periodic images, float64, the reference's operation order, no FMA
 * -- bit-identical.  d_c1 double[n][3], d_c2 double[m][3], d_lattice double[3][3] (rows = lattice vectors), all device
 * memory; stream-ordered, no synchronisation.  The assignment kernels fuse this with their argmin / cut-off test and
 * never build the matrix; this entry serves callers that need the matrix (coordination search, index mapping,
 * alignment objective: tools.py:32-111, reference_workflow/structure_aligner.py:105-147). */
/* ---- exact replay of the reference's lazily updated per-site caches (polyhedral sites) -------
 * The frame-parallel kernels are identical to the reference while their guards hold.  When a
 * framework step >= 0.3 invalidates a site's PBC-shift cache (pbc_utils.py:216) or the unwrapped excursion reaches 0.3,
 * the reference's answer depends on WHEN each site was last queried (polyhedral_site.py:322-416); sab_assign_lazy
 * replays that literally on the device: frames in order, one cache per site with the frame stamp of its last query,
 * the priority walk of site_collection.py:113-182 one query at a time, Counters updated in atom order
 * (csrc/k_polyhedral_lazy.cuh).  It returns final site list positions (-1 = None; no spill codes) and advances the
 * history arrays (same layout as sab_resolve) itself.
 * sab_lazy_set_state: per-site caches as the reference holds them BEFORE the first frame: shift int32[S][vmax][3] and
 * cached_raw double[S][vmax][3] where has_shift[s] (sites primed by the builder, site_factory.py:248-257), reference
 * centres double[S][3] where has_ref_centre[s].  sab_lazy_get_state reads the caches back (any pointer may be NULL;
 * stats int64[4] = queries, incremental updates, full corrections, invalidations).
 * On SAB_EOVERFLOW (Counter row wider than tr_width) the caller restores the per-site state and retries wider. */
int sab_lazy_set_state(sab_engine *e, const int32_t *shift, const double *cached_raw, const uint8_t *has_shift,
                       const uint8_t *has_ref_centre, const double *ref_centre);
int sab_lazy_get_state(sab_engine *e, int32_t *shift, double *cached_raw, uint8_t *has_shift, double *vertex_coords,
                       int64_t *stats);
int sab_assign_lazy(sab_engine *e, const double *d_frac, const double *d_lattice, int lattice_stride, int64_t n_frames,
                    int32_t *d_result, void *stream, int32_t *recent, int32_t *prev, int32_t *tr_n, int32_t *tr_key,
                    int64_t *tr_cnt, int tr_width, int append);

/* Face topology at the first query (non-simplicial polyhedra).  The reference hulls a site when it is first QUERIED
 * (FaceTopologyCache, polyhedral_site.py:138-159, 468-471) and Qhull's triangulation of e.g. a cube's quads depends on the
 * vertex coordinates of that frame.  sab_lazy_set_open_topology marks the sites whose uploaded triangulation is provisional
 * (uint8[S], NULL = none) and clears the records; the following sab_assign_lazy calls record, for every open site they query
 * for the first time, the global frame number and vertex_coords of that query; sab_lazy_get_first_queries reads them back
 * (first_frame int64[S], -1 = not queried; first_vc double[S][vmax][3]).  The host re-hulls from exactly those coordinates,
 * uploads changed triangulations with sab_update_face_topology (same layout as sab_set_polyhedra's simplices; nothing else
 * of the engine is touched), restores the per-site state and redoes the batch until no triangulation changes. */
int sab_lazy_set_open_topology(sab_engine *e, const uint8_t *open_site);
int sab_lazy_get_first_queries(sab_engine *e, int64_t *first_frame, double *first_vc);
int sab_update_face_topology(sab_engine *e, int fmax, const int32_t *n_face, const int32_t *simplices);

/* ---- polyhedral containment of given points (stand-alone operator) ----------------------------
 * d_out uint8[n_sites]: 1 when ANY of the site's n_img query points lies inside the site's polyhedron -- the query of
 * PolyhedralSite.contains_point / contains_atom (polyhedral_site.py:431-520: FaceTopologyCache.update +
 * _numba_sn_query, :26-123) and, over many sites at once, of PolyhedralSiteCollection.sites_contain_points
 * (polyhedral_site_collection.py:113-141).  d_vertex double[n_sites][vmax][3] PBC-corrected vertex coordinates
 * (PolyhedralSite.vertex_coords), d_n_vert int32[n_sites]; d_simplices int32[n_sites][fmax][3] local vertex numbers of
 * the hull's triangles (Qhull order), d_n_face int32[n_sites]; d_points double[n_sites][n_img][3] (the caller passes
 * the 8 images x + {0,1}^3 of tools.py:246-279, or its own).  float64, the reference's operation order, no FMA:
 * centroid as the sequential sum over vertices / n, normal (v0 - v2) x (v1 - v2), reject iff dot != 0 and
 * sign(dot) != sign(centre dot).  Device memory, stream-ordered, no synchronisation; vmax <= 16. */
int sab_polyhedra_contain(int device, const double *d_vertex, const int32_t *d_n_vert, int vmax,
                          const int32_t *d_simplices, const int32_t *d_n_face, int fmax, int n_sites,
                          const double *d_points, int n_img, uint8_t *d_out, void *stream);

int sab_mic_distances(int device, const double *d_c1, int n, const double *d_c2, int m,
                      const double *d_lattice, double *d_out, void *stream);

/* Measured FP64 arithmetic peak of the device in TFLOP/s, tflops[4] = {DADD only, DMUL only, DMUL + DADD pairs, DFMA counted
 * as 2}: the roofline denominator of the kernels whose float64 decisions must not use FMA. */
int sab_measure_fp64_peak(int device, double *tflops);

/* Multi-GPU, frames sharded over ranks: the reference's atoms_trajectory (trajectory.py:375-383) is the concatenation
 * of the shards' rows.  Instead of an all-gather AFTER the kernels, the stream kernel stores every result row it
 * produces into the other ranks' copies of the full array as well -- plain stores to NVLink peer memory, overlapped
 * with the computation frame by frame.  d_peer_slots[q] (host array of n_peers <= 7 DEVICE pointers): where this
 * rank's rows start inside peer q's int32[F_total][A] array (peer-mapped address, e.g. from CUDA IPC / torch
 * symmetric memory).  Applies to the following sab_assign_device calls (info[11] tells whether the rows were stored;
 * kernels other than the tetrahedral stream kernel ignore the peers); n_peers = 0 clears.  The caller orders the
 * peers' reads after this rank's kernels (any collective issued after them on the same stream does). */
int sab_set_result_peers(sab_engine *e, int n_peers, int32_t *const *d_peer_slots);
/* The same through ONE multicast mapping (NVLS; e.g. torch symmetric memory's multicast_ptr): d_multicast_slot = where this
 * rank's rows start inside the multicast view of the int32[F_total][A] array.  Every row is then stored once with
 * multimem.st and the NVSwitch replicates it into every rank's copy (this rank's included) -- 1/(N-1) of the NVLink egress of
 * the unicast form.  NULL clears.  Replaces any peers set with sab_set_result_peers. */
int sab_set_result_multicast(sab_engine *e, int32_t *d_multicast_slot);

/* ---- XDATCAR text -> arrays (host code) -----------------------------
 * The reference takes Sequence[Structure] (trajectory.py:413-432), which users build with pymatgen's Xdatcar parser;
 * these two calls turn the text (already in memory, e.g. an mmap or a decompressed .gz) into the frac / lattice arrays
 * of sab_assign_host directly.  Numbers are converted exactly as Python float() would (correctly rounded).
 * sab_xdatcar_index: one pass from body_offset (the byte after the first 7-line header); frame_off[k] = offset of
 *   frame k's first coordinate line, header_off[k] = offset of the header repeated in front of it (variable-cell
 *   files) or -1; *n_frames = frames found; SAB_EOVERFLOW when that exceeds capacity (call again with room).
 * sab_xdatcar_parse: frames split over n_threads host threads; frac double[F][N][3]; lattice (may be NULL)
 *   double[F][3][3], written only for frames with header_off >= 0 (scale applied; negative scale = volume).
 * sab_parse_double: one number (testing aid for the conversion). */
int sab_xdatcar_index(const char *text, int64_t len, int64_t body_offset, int n_atoms, int64_t *frame_off,
                      int64_t *header_off, int64_t capacity, int64_t *n_frames);
int sab_xdatcar_parse(const char *text, int64_t len, int n_atoms, int64_t n_frames, const int64_t *frame_off,
                      const int64_t *header_off, double *frac, double *lattice, int n_threads);
int sab_parse_double(const char *text, int64_t len, double *out, int64_t *consumed);

#ifdef __cplusplus
}
#endif
#endif /* SAB200_H */
