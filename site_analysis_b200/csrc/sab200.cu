// sab200.cu -- engine and C ABI of libsab200.so (see include/sab200.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 -shared -Xcompiler -fPIC
// -fmad=false is part of the numerical contract (no FMA contraction anywhere).
#include "../../include/sab200.h"

#include <algorithm>
#include <array>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "k_nearest.cuh"
#include "k_nearest_cells.cuh"
#include "k_polyhedral.cuh"
#include "k_polyhedral_cells.cuh"
#include "k_polyhedral_seq.cuh"
#include "k_polyhedral_seq1.cuh"
#include "k_polyhedral_tile.cuh"
#include "k_polyhedral_stream.cuh"
#include "k_polyhedral_stream5.cuh"
#include "k_resolve.cuh"
#include "k_resolve_fast.cuh"
#include "k_resolve_spec.cuh"
#include "k_polyhedral_lazy.cuh"
#include "k_spherical.cuh"
#include "k_wrap.cuh"

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(SAB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// Host -> device copy into *dst, (re)allocating only when the size changes (sizes are tracked per
// pointer so that repeated state uploads -- e.g. once per shard pass -- cost one memcpy each).
#include <unordered_map>
static std::unordered_map<void *, size_t> g_alloc_bytes;
template <class T> static cudaError_t upload(T **dst, const T *src, size_t n)
{
    size_t bytes = n * sizeof(T);
    auto it = *dst ? g_alloc_bytes.find((void *)*dst) : g_alloc_bytes.end();
    if (*dst && (it == g_alloc_bytes.end() || it->second != bytes)) {
        if (it != g_alloc_bytes.end()) g_alloc_bytes.erase(it);
        cudaFree(*dst); *dst = nullptr;
    }
    if (n == 0) return cudaSuccess;
    if (!*dst) {
        cudaError_t e = cudaMalloc((void **)dst, bytes);
        if (e != cudaSuccess) return e;
        g_alloc_bytes[(void *)*dst] = bytes;
    }
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
}

#define SAB_MISSING_CAP 4096     // anchors the resolver can report as "row needed" per run
enum { CNT_AMBIGUOUS = 0, CNT_SPILL = 1, CNT_OVERFLOW = 2, CNT_FIRST_BAD = 3, CNT_FULL27 = 4, CNT_SPAN = 8, CNT_N = 12 };

struct sab_engine {
    int kind = 0, device = 0, N = 0, A = 0, S = 0;
    int sm_count = 148;
    int32_t *d_mobile = nullptr;
    double *d_centres = nullptr, *d_rcut = nullptr, *d_rcut_q = nullptr;
    // polyhedron vertices / dynamic-Voronoi reference atoms ("slots")
    int smax = 0, fmax = 0, M = 0;
    std::vector<int32_t> fw_atoms;
    int32_t *d_fw_atoms = nullptr, *d_nslot = nullptr, *d_slot_atom = nullptr, *d_slot_fw = nullptr;
    int32_t *d_nface = nullptr; uint8_t *d_simp = nullptr;
    // PBC state
    int32_t *d_slot_shift = nullptr, *d_wraps = nullptr;
    double *d_prev_raw = nullptr, *d_excursion = nullptr;
    bool pbc_set = false;
    uint8_t *d_legacy_init = nullptr; int64_t legacy_init_frame = -1;
    // priority tables
    int32_t *d_rank = nullptr, *d_rank_pos = nullptr, *d_neigh_off = nullptr, *d_neigh = nullptr; double *d_lookup = nullptr;
    // lazy ranking: cached rows of single anchors (sab_add_rank_rows), anchors reported missing by the last resolver run
    int ranked = 0; std::vector<int32_t> row_slot_h; int n_row_slots = 0, row_slot_cap = 0;
    int32_t *d_row_slot = nullptr, *d_row_pos = nullptr, *d_row_fwd = nullptr, *d_missing = nullptr; int *d_n_missing = nullptr;
    // counters
    unsigned long long *d_counters = nullptr;
    // host-path staging
    void *stage[2] = {nullptr, nullptr}; size_t stage_bytes = 0;
    int32_t *d_mirror = nullptr;           // sab_set_result_mirror: device copy of the rows sab_assign_host returns
    void *h_res_stage[2] = {nullptr, nullptr}; size_t h_res_stage_bytes = 0;   // pinned staging of results bound for pageable memory
    void *ws = nullptr; size_t ws_bytes = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t ev_free[2] = {nullptr, nullptr};
    int64_t launches = 0;
    bool skip_step_check_once = false;     // SAB_OPT_SKIP_STEP_CHECK_ONCE
    int poly_kernel = 0;                   // SAB_OPT_POLY_KERNEL: 0 auto (chunk-sequential), 1 frame-parallel cells
    int seq_chunk = 0;                     // SAB_OPT_SEQ_CHUNK (0 = automatic)
    int fp32_filter = 1;                   // SAB_OPT_FP32_FILTER
    int seq_block = 0;                     // SAB_OPT_SEQ_BLOCK (0 = one warp per 32 mobile atoms)
    cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};   // wrap pass start | assign kernel start | assign kernel end
    // exact-pruning tables of the nearest-centre kernel
    bool nl_set = false; int nl_G = 0; double nl_l0inv[9] = {0};
    int32_t *d_nl_off = nullptr; uint16_t *d_nl_sites = nullptr; double *d_nl_excl = nullptr;
    double *d_nl_anchor = nullptr, *d_nl_l0 = nullptr; double nl_dmax = 0.0;   // dynamic Voronoi: anchor centres, L0, motion head-room
    double nl_rho = 1.0;                   // spherical: inflation factor of the sphere lists
    // exact-pruning tables of the tetrahedral kernel
    bool cells_set = false; int G = 0, kmin = 0, nk = 0; double delta = 0.0;
    uint16_t *d_rec = nullptr, *d_cell_sites = nullptr; int32_t *d_cell_off = nullptr; double *d_anchor = nullptr;
    // tile kernel: compact staging table and float-table site records (derived in sab_set_cell_lists)
    bool tile_ok = false; int n_used = 0; int4 *d_tile_vars = nullptr; uint16_t *d_tile_rec = nullptr;
    int tile_frames = 0;                   // SAB_OPT_TILE_FRAMES (0 = automatic)
    int stream_prod = 0;                   // SAB_OPT_STREAM_PRODUCERS (0 = automatic)
    int stream_rec_smem = 1;               // SAB_OPT_STREAM_REC_SMEM (1 = site records in shared memory when they fit)
    int stream_reg_cols = 1;               // SAB_OPT_STREAM_REG_COLS (1 = producer column metadata in registers when they fit)
    bool force_scan = false;               // the stream kernel's chain check failed: this batch runs on the scan kernels
    // generation-5 stream kernel (k_polyhedral_stream5.cuh): vertex-image rows, derived in sab_set_cell_lists
    bool s5_ok = false; float4 *d_s5_task = nullptr;
    int resolver = 0;                                     // SAB_OPT_RESOLVER: 0 = speculative chunks, 1 = per-frame (shared memory), 2 = global memory
    int64_t resolve_ns = 0; int resolve_kind = 0;         // device time of all sab_resolve kernels so far; kernel of the last call
    uint2 *d_s5_rec = nullptr; uint16_t *d_s5_cellrec = nullptr; float s5_T1 = 0.f, s5_T0 = 0.f, s5_TD = 0.f;
    int stream_gen = 0;                    // SAB_OPT_STREAM_GEN: 0 automatic (5 when its tables exist), 4 = previous generation
    int s5_rows = 0, s5_raw = 0, s5_blocks = 0;   // SAB_OPT_STREAM5_ROWS / _RAW / _BLOCKS (0 = automatic)
    size_t s5_occ_key = 0; int s5_per_sm = 0;
    // exact replay of the reference's lazily updated per-site caches (k_polyhedral_lazy.cuh)
    int32_t *d_lz_shift = nullptr, *d_lz_has = nullptr, *d_lz_lock = nullptr; double *d_lz_cached = nullptr, *d_lz_vc = nullptr, *d_lz_rc = nullptr;
    long long *d_lz_stamp = nullptr; uint8_t *d_lz_has_rc = nullptr; unsigned long long *d_lz_stats = nullptr; bool lazy_set = false;
    int32_t *d_lz_open = nullptr; long long *d_lz_first_frame = nullptr; double *d_lz_first_vc = nullptr;   // sab_lazy_set_open_topology
    long long lazy_frames = 0;
    int n_peer = 0; int32_t *peer[SAB_MAX_PEERS] = {nullptr}; int peer_mc = 0;   // sab_set_result_peers / sab_set_result_multicast
    bool peers_written = false;            // the last sab_assign_device stored its rows into the peers as well
    size_t stream_occ_key = 0; int stream_per_sm = 0;     // cached occupancy query of the stream kernel
    unsigned long long *h_pinned = nullptr;               // pinned [2][CNT_N]: counter initialiser | read-back
};

extern "C" int sab_abi_version(void) { return SAB_ABI_VERSION; }
extern "C" const char *sab_last_error(void) { return g_err.c_str(); }

extern "C" int sab_engine_create(sab_engine **out, int kind, int device, int n_atoms, int n_mobile,
                                 const int32_t *mobile_idx, int n_sites)
{
    if (!out || kind < 0 || kind > 3 || n_atoms <= 0 || n_mobile <= 0 || n_sites <= 0 || !mobile_idx)
        return fail(SAB_EINVAL, "sab_engine_create: bad arguments");
    for (int a = 0; a < n_mobile; a++)
        if (mobile_idx[a] < 0 || mobile_idx[a] >= n_atoms)
            return fail(SAB_EINVAL, "mobile atom index %d out of range [0,%d)", mobile_idx[a], n_atoms);
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SAB_ECUDA, "CUDA device %d not available (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    sab_engine *e = new sab_engine();
    e->kind = kind; e->device = device; e->N = n_atoms; e->A = n_mobile; e->S = n_sites;
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    CU(upload(&e->d_mobile, mobile_idx, (size_t)n_mobile));
    CU(cudaMalloc((void **)&e->d_counters, CNT_N * sizeof(unsigned long long)));
    *out = e;
    return SAB_OK;
}

extern "C" void sab_engine_destroy(sab_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    void *ptrs[] = {e->d_mobile, e->d_centres, e->d_rcut, e->d_rcut_q, e->d_fw_atoms, e->d_nslot, e->d_slot_atom,
                    e->d_slot_fw, e->d_nface, e->d_simp, e->d_slot_shift, e->d_wraps, e->d_prev_raw, e->d_excursion,
                    e->d_legacy_init, e->d_rank, e->d_rank_pos, e->d_neigh_off, e->d_neigh, e->d_lookup, e->d_counters,
                    e->stage[0], e->stage[1], e->ws, e->d_rec, e->d_cell_sites, e->d_cell_off, e->d_anchor,
                    e->d_nl_off, e->d_nl_sites, e->d_nl_excl, e->d_nl_anchor, e->d_nl_l0, e->d_tile_vars, e->d_tile_rec,
                    e->d_s5_task, e->d_s5_rec, e->d_s5_cellrec, e->d_lz_shift, e->d_lz_has, e->d_lz_lock, e->d_lz_cached,
                    e->d_lz_vc, e->d_lz_rc, e->d_lz_stamp, e->d_lz_has_rc, e->d_lz_stats, e->d_lz_open, e->d_lz_first_frame, e->d_lz_first_vc, e->d_row_slot, e->d_row_pos, e->d_row_fwd,
                    e->d_missing, e->d_n_missing};
    for (void *p : ptrs) if (p) { g_alloc_bytes.erase(p); cudaFree(p); }
    if (e->h_pinned) cudaFreeHost(e->h_pinned);
    for (int i = 0; i < 2; i++) if (e->h_res_stage[i]) cudaFreeHost(e->h_res_stage[i]);
    for (int i = 0; i < 3; i++) if (e->ev_t[i]) cudaEventDestroy(e->ev_t[i]);
    for (int i = 0; i < 2; i++) {
        if (e->streams[i]) cudaStreamDestroy(e->streams[i]);
        if (e->ev_free[i]) cudaEventDestroy(e->ev_free[i]);
    }
    delete e;
}

extern "C" int sab_set_centres(sab_engine *e, const double *centres, const double *rcut, const double *rcut_sq_max)
{
    if (!e || !centres) return fail(SAB_EINVAL, "sab_set_centres: null argument");
    if (e->kind != SAB_SPHERICAL && e->kind != SAB_VORONOI) return fail(SAB_EINVAL, "sab_set_centres: wrong engine kind");
    if (e->kind == SAB_SPHERICAL && (!rcut || !rcut_sq_max)) return fail(SAB_EINVAL, "spherical sites need rcut and rcut_sq_max");
    CU(cudaSetDevice(e->device));
    CU(upload(&e->d_centres, centres, (size_t)3 * e->S));
    if (e->kind == SAB_SPHERICAL) {
        CU(upload(&e->d_rcut, rcut, (size_t)e->S));
        CU(upload(&e->d_rcut_q, rcut_sq_max, (size_t)e->S));
    }
    return SAB_OK;
}

static int set_slots(sab_engine *e, int smax, const int32_t *n_slot, const int32_t *slot_atom)
{
    if (smax <= 0) return fail(SAB_EINVAL, "slot width must be positive");
    std::vector<int32_t> uniq;
    for (int s = 0; s < e->S; s++) {
        if (n_slot[s] <= 0 || n_slot[s] > smax) return fail(SAB_EINVAL, "site %d: bad vertex/reference count %d", s, n_slot[s]);
        for (int v = 0; v < n_slot[s]; v++) {
            int32_t at = slot_atom[(size_t)s * smax + v];
            if (at < 0 || at >= e->N) return fail(SAB_EINVAL, "site %d: atom index %d out of range", s, at);
            uniq.push_back(at);
        }
    }
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    e->fw_atoms = uniq; e->M = (int)uniq.size(); e->smax = smax;
    std::vector<int32_t> fw((size_t)e->S * smax, 0), at((size_t)e->S * smax, 0);
    for (int s = 0; s < e->S; s++)
        for (int v = 0; v < n_slot[s]; v++) {
            int32_t a = slot_atom[(size_t)s * smax + v];
            at[(size_t)s * smax + v] = a;
            fw[(size_t)s * smax + v] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), a) - uniq.begin());
        }
    CU(upload(&e->d_fw_atoms, uniq.data(), uniq.size()));
    CU(upload(&e->d_nslot, n_slot, (size_t)e->S));
    CU(upload(&e->d_slot_atom, at.data(), at.size()));
    CU(upload(&e->d_slot_fw, fw.data(), fw.size()));
    e->pbc_set = false;
    e->cells_set = false;
    return SAB_OK;
}

extern "C" int sab_set_polyhedra(sab_engine *e, int vmax, const int32_t *n_vert, const int32_t *vert_idx,
                                 int fmax, const int32_t *n_face, const int32_t *simplices)
{
    if (!e || !n_vert || !vert_idx || !n_face || !simplices) return fail(SAB_EINVAL, "sab_set_polyhedra: null argument");
    if (e->kind != SAB_POLYHEDRAL) return fail(SAB_EINVAL, "sab_set_polyhedra: wrong engine kind");
    if (vmax > SAB_MAX_VERT) return fail(SAB_EINVAL, "polyhedra with more than %d vertices are not supported", SAB_MAX_VERT);
    if (fmax <= 0) return fail(SAB_EINVAL, "fmax must be positive");
    CU(cudaSetDevice(e->device));
    int rc = set_slots(e, vmax, n_vert, vert_idx);
    if (rc) return rc;
    std::vector<uint8_t> simp((size_t)e->S * fmax * 3, 0);
    for (int s = 0; s < e->S; s++) {
        if (n_face[s] < 4 || n_face[s] > fmax) return fail(SAB_EINVAL, "site %d: bad face count %d", s, n_face[s]);
        for (int j = 0; j < 3 * n_face[s]; j++) {
            int v = simplices[(size_t)s * fmax * 3 + j];
            if (v < 0 || v >= n_vert[s]) return fail(SAB_EINVAL, "site %d: face vertex %d out of range", s, v);
            simp[(size_t)s * fmax * 3 + j] = (uint8_t)v;
        }
    }
    e->fmax = fmax;
    CU(upload(&e->d_nface, n_face, (size_t)e->S));
    CU(upload(&e->d_simp, simp.data(), simp.size()));
    return SAB_OK;
}

extern "C" int sab_set_reference_atoms(sab_engine *e, int rmax, const int32_t *n_ref, const int32_t *ref_idx)
{
    if (!e || !n_ref || !ref_idx) return fail(SAB_EINVAL, "sab_set_reference_atoms: null argument");
    if (e->kind != SAB_DYNAMIC_VORONOI) return fail(SAB_EINVAL, "sab_set_reference_atoms: wrong engine kind");
    CU(cudaSetDevice(e->device));
    return set_slots(e, rmax, n_ref, ref_idx);
}

extern "C" int sab_framework_count(sab_engine *e) { return e ? e->M : 0; }
extern "C" int sab_framework_atoms(sab_engine *e, int32_t *out)
{
    if (!e || !out) return fail(SAB_EINVAL, "sab_framework_atoms: null argument");
    std::copy(e->fw_atoms.begin(), e->fw_atoms.end(), out);
    return SAB_OK;
}

extern "C" int sab_set_pbc_state(sab_engine *e, const int32_t *slot_shift, const double *prev_raw,
                                 const int32_t *wraps, const double *excursion)
{
    if (!e || !slot_shift || !prev_raw || !wraps || !excursion) return fail(SAB_EINVAL, "sab_set_pbc_state: null argument");
    if (e->M == 0) return fail(SAB_ESTATE, "sab_set_pbc_state: no vertex/reference tables set");
    CU(cudaSetDevice(e->device));
    CU(upload(&e->d_slot_shift, slot_shift, (size_t)e->S * e->smax * 3));
    CU(upload(&e->d_prev_raw, prev_raw, (size_t)e->M * 3));
    CU(upload(&e->d_wraps, wraps, (size_t)e->M * 3));
    CU(upload(&e->d_excursion, excursion, (size_t)e->M * 6));
    e->pbc_set = true;
    return SAB_OK;
}

extern "C" int sab_get_pbc_state(sab_engine *e, int32_t *slot_shift, double *prev_raw, int32_t *wraps, double *excursion)
{
    if (!e || !e->pbc_set) return fail(SAB_ESTATE, "sab_get_pbc_state: PBC state not set");
    CU(cudaSetDevice(e->device));
    CU(cudaDeviceSynchronize());
    if (slot_shift) CU(cudaMemcpy(slot_shift, e->d_slot_shift, sizeof(int32_t) * (size_t)e->S * e->smax * 3, cudaMemcpyDeviceToHost));
    if (prev_raw) CU(cudaMemcpy(prev_raw, e->d_prev_raw, sizeof(double) * (size_t)e->M * 3, cudaMemcpyDeviceToHost));
    if (wraps) CU(cudaMemcpy(wraps, e->d_wraps, sizeof(int32_t) * (size_t)e->M * 3, cudaMemcpyDeviceToHost));
    if (excursion) CU(cudaMemcpy(excursion, e->d_excursion, sizeof(double) * (size_t)e->M * 6, cudaMemcpyDeviceToHost));
    return SAB_OK;
}

extern "C" int sab_set_carry_device(sab_engine *e, const double *d_halo_raw, const double *d_anchor0, void *stream)
{
    if (!e || !d_halo_raw || !d_anchor0) return fail(SAB_EINVAL, "sab_set_carry_device: null argument");
    if (!e->pbc_set) return fail(SAB_ESTATE, "sab_set_carry_device: PBC state not set");
    CU(cudaSetDevice(e->device));
    const int m3 = 3 * e->M;
    k_carry_from_halo<<<(m3 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_halo_raw, d_anchor0, m3, e->d_prev_raw, e->d_wraps, e->d_excursion);
    e->launches++;
    CU(cudaGetLastError());
    return SAB_OK;
}

extern "C" int sab_get_excursion_device(sab_engine *e, double *d_out, void *stream)
{
    if (!e || !d_out) return fail(SAB_EINVAL, "sab_get_excursion_device: null argument");
    if (!e->pbc_set) return fail(SAB_ESTATE, "sab_get_excursion_device: PBC state not set");
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(d_out, e->d_excursion, sizeof(double) * (size_t)e->M * 6, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SAB_OK;
}

extern "C" int sab_set_legacy_init(sab_engine *e, const uint8_t *site_flags, int64_t frames_from_now)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_legacy_init: null engine");
    CU(cudaSetDevice(e->device));
    if (!site_flags) { if (e->d_legacy_init) { cudaFree(e->d_legacy_init); e->d_legacy_init = nullptr; } e->legacy_init_frame = -1; return SAB_OK; }
    CU(upload(&e->d_legacy_init, site_flags, (size_t)e->S));
    e->legacy_init_frame = frames_from_now;
    return SAB_OK;
}

extern "C" int sab_set_cell_lists(sab_engine *e, int G, double delta, const double *anchor, int kmin, int nk,
                                  const uint16_t *records, const int32_t *cell_off, const uint16_t *cell_sites)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_cell_lists: null engine");
    CU(cudaSetDevice(e->device));
    if (!anchor) { e->cells_set = false; return SAB_OK; }           // NULL clears the tables
    if (e->kind != SAB_POLYHEDRAL || !e->d_simp) return fail(SAB_ESTATE, "sab_set_cell_lists: polyhedra not set");
    if (e->smax != 4 || e->fmax != 4) return fail(SAB_EINVAL, "cell lists are implemented for tetrahedral sites");
    if (G < 1 || G > 256 || nk < 1 || nk > 8 || !records || !cell_off || !cell_sites || !(delta > 0.0))
        return fail(SAB_EINVAL, "sab_set_cell_lists: bad arguments");
    if ((size_t)nk * 3 * e->M * 8 > 65535 || e->S > 65535) return fail(SAB_EINVAL, "cell lists: too many framework atoms or sites for 16-bit offsets");
    size_t ncell = (size_t)G * G * G;
    CU(upload(&e->d_anchor, anchor, (size_t)3 * e->M));
    CU(upload(&e->d_rec, records, (size_t)e->S * SAB_TET_REC));
    CU(upload(&e->d_cell_off, cell_off, ncell + 1));
    CU(upload(&e->d_cell_sites, cell_sites, (size_t)std::max(cell_off[ncell], 1)));
    e->G = G; e->delta = delta; e->kmin = kmin; e->nk = nk; e->cells_set = true;
    // ---- tables of the tile kernel (k_polyhedral_tile.cuh): only the (coordinate, shift) pairs in use are staged
    {
        const int m3 = 3 * e->M;
        std::vector<int> slot((size_t)nk * m3, -1);
        for (int s = 0; s < e->S; s++)
            for (int v = 0; v < 12; v++) {
                const int idx = records[(size_t)s * SAB_TET_REC + v] / 8;
                if (idx >= nk * m3) return fail(SAB_EINVAL, "sab_set_cell_lists: record offset out of range (site %d)", s);
                slot[idx] = 0;
            }
        std::vector<int4> vars;
        bool ok = e->A <= 4096 && (int64_t)e->N * 3 * SAB_TILE_MAX < INT_MAX && delta > 1e-6;
        for (int i = 0; i < m3; i++)
            for (int k = 0; k < nk; k++)
                if (slot[(size_t)k * m3 + i] == 0) {
                    slot[(size_t)k * m3 + i] = (int)vars.size();
                    const float sh = (float)(kmin + k), anc = (float)anchor[i];
                    int4 e4;
                    e4.x = e->fw_atoms[i / 3] * 3 + i % 3; e4.y = i;
                    memcpy(&e4.z, &sh, 4); memcpy(&e4.w, &anc, 4);
                    vars.push_back(e4);
                    // FP32 pre-test preconditions inside the validity tube: unwrapped vertex coordinates stay
                    // below 2 (image rule) and below 3.5 in magnitude (error budget)
                    const double v = anchor[i] + (double)(kmin + k);
                    if (!(v + delta < 2.0 - 1e-6) || !(fabs(v) + delta < 3.5)) ok = false;
                }
        const size_t stride = (vars.size() + 3 * (size_t)e->A + 3) & ~(size_t)3;
        if (stride * 4 > 65535 || vars.empty()) ok = false;          // 16-bit byte offsets into a row
        size_t longest = 0;
        for (size_t c = 0; c < ncell; c++) longest = std::max(longest, (size_t)(cell_off[c + 1] - cell_off[c]));
        if (longest > 16 * (size_t)SAB_TILE_BLOCKS) ok = false;
        e->tile_ok = false;
        if (ok) {
            std::vector<uint16_t> rec2((size_t)e->S * 12);
            for (int s = 0; s < e->S; s++)
                for (int v = 0; v < 12; v++) rec2[(size_t)s * 12 + v] = (uint16_t)(4 * slot[records[(size_t)s * SAB_TET_REC + v] / 8]);
            CU(upload(&e->d_tile_vars, vars.data(), vars.size()));
            CU(upload(&e->d_tile_rec, rec2.data(), rec2.size()));
            e->n_used = (int)vars.size();
            e->tile_ok = true;
        }
    }
    // ---- tables of the generation-5 stream kernel (k_polyhedral_stream5.cuh): rows of wrapped float positions per atom
    e->s5_ok = false;
    if (e->tile_ok && delta <= 0.12) {
        const int m3 = 3 * e->M;
        bool ok = (e->M + 1) * 16 <= 65535;                                                      // 16-bit byte offsets into a row
        for (int i = 0; i < m3; i++) if ((double)(float)anchor[i] != anchor[i]) ok = false;     // anchors travel as float
        std::vector<float4> task(e->M);
        for (int m = 0; m < e->M && ok; m++) {
            const int col = e->fw_atoms[m] * 3;
            float cf; memcpy(&cf, &col, 4);
            task[m] = make_float4(cf, (float)anchor[3 * m], (float)anchor[3 * m + 1], (float)anchor[3 * m + 2]);
        }
        std::vector<uint2> rec5(e->S);
        double ee_max = 0.0;
        for (int s = 0; s < e->S && ok; s++) {
            int pos[4]; double P[4][3];
            for (int v = 0; v < 4; v++)
                for (int d = 0; d < 3; d++) {
                    const int idx = records[(size_t)s * SAB_TET_REC + 3 * v + d] / 8, k = idx / m3, i = idx % m3;
                    if (i % 3 != d || (d > 0 && i / 3 != pos[v])) ok = false;
                    pos[v] = i / 3;
                    P[v][d] = anchor[i] + (double)(kmin + k);                  // anchor vertex, unwrapped
                }
            for (int v = 0; v < 4; v++) for (int u = 0; u < v; u++) for (int d = 0; d < 3; d++)
                ee_max = std::max(ee_max, fabs(P[v][d] - P[u][d]));
            double B[3], Cv[3], D[3];
            for (int d = 0; d < 3; d++) { B[d] = P[1][d] - P[0][d]; Cv[d] = P[2][d] - P[0][d]; D[d] = P[3][d] - P[0][d]; }
            const double det = D[0] * (B[1] * Cv[2] - B[2] * Cv[1]) + D[1] * (B[2] * Cv[0] - B[0] * Cv[2]) + D[2] * (B[0] * Cv[1] - B[1] * Cv[0]);
            if (det < 0.0) std::swap(pos[2], pos[3]);                      // positively oriented records: no sign flip in the kernel
            rec5[s].x = (unsigned)(16 * (1 + pos[0])) | ((unsigned)(16 * (1 + pos[1])) << 16);   // entry 0 of a row is its flag
            rec5[s].y = (unsigned)(16 * (1 + pos[2])) | ((unsigned)(16 * (1 + pos[3])) << 16);
        }
        // reduced separations equal the reference's unwrapped differences only for sites smaller than half a cell
        const double ee = ee_max + 2.0 * delta + 1e-5, ep = 0.5 + 1e-6, d22 = 2.384185791015625e-7;
        if (!(ee < 0.45)) ok = false;
        if (ok) {
            const double t1 = 2.0 * (28.0 * ee * ep + 16.0 * ee * ee) * d22 * (1.0 + 1e-5), td = 2.0 * 44.0 * ee * ee * d22 * (1.0 + 1e-5);
            e->s5_T1 = nextafterf((float)t1, INFINITY); e->s5_TD = nextafterf((float)td, INFINITY);
            e->s5_T0 = nextafterf((float)(3.0 * (double)e->s5_T1 + (double)e->s5_TD), INFINITY);
            CU(upload(&e->d_s5_task, task.data(), task.size()));
            CU(upload(&e->d_s5_rec, rec5.data(), rec5.size()));
            // fixed-size cell records: {length, entries...} in 32 uint16 (one 64-byte load per search); 0xffff = use the CSR form
            std::vector<uint16_t> crec(ncell * 32, 0);
            for (size_t c = 0; c < ncell; c++) {
                const int n = cell_off[c + 1] - cell_off[c];
                crec[c * 32] = n <= 31 ? (uint16_t)n : (uint16_t)0xffff;
                for (int i = 0; i < std::min(n, 31); i++) crec[c * 32 + 1 + i] = cell_sites[cell_off[c] + i];
            }
            CU(upload(&e->d_s5_cellrec, crec.data(), crec.size()));
            e->s5_ok = true;
        }
    }
    return SAB_OK;
}

extern "C" int sab_set_nearest_lists(sab_engine *e, int G, const double *lattice0_inverse, const int32_t *cell_off,
                                     const uint16_t *cell_sites, const double *excl)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_nearest_lists: null engine");
    CU(cudaSetDevice(e->device));
    if (!cell_off) { e->nl_set = false; return SAB_OK; }
    if (e->kind == SAB_DYNAMIC_VORONOI) return fail(SAB_EINVAL, "use sab_set_dynvor_lists for dynamic Voronoi sites");
    if (e->kind != SAB_VORONOI || !e->d_centres) return fail(SAB_ESTATE, "sab_set_nearest_lists: Voronoi centres not set");
    if (G < 1 || G > 256 || !lattice0_inverse || !cell_sites || !excl || e->S > 65535)
        return fail(SAB_EINVAL, "sab_set_nearest_lists: bad arguments");
    size_t ncell = (size_t)G * G * G;
    CU(upload(&e->d_nl_off, cell_off, ncell + 1));
    CU(upload(&e->d_nl_sites, cell_sites, (size_t)std::max(cell_off[ncell], 1)));
    CU(upload(&e->d_nl_excl, excl, ncell));
    for (int i = 0; i < 9; i++) e->nl_l0inv[i] = lattice0_inverse[i];
    e->nl_G = G; e->nl_set = true;
    return SAB_OK;
}

extern "C" int sab_set_sphere_lists(sab_engine *e, int G, const double *lattice0_inverse, double rho, const int32_t *cell_off,
                                    const uint16_t *cell_sites)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_sphere_lists: null engine");
    CU(cudaSetDevice(e->device));
    if (!cell_off) { e->nl_set = false; return SAB_OK; }
    if (e->kind != SAB_SPHERICAL || !e->d_centres) return fail(SAB_ESTATE, "sab_set_sphere_lists: spherical sites not set");
    if (G < 1 || G > 256 || !lattice0_inverse || !cell_sites || e->S > 65535 || !(rho >= 1.0))
        return fail(SAB_EINVAL, "sab_set_sphere_lists: bad arguments");
    size_t ncell = (size_t)G * G * G;
    CU(upload(&e->d_nl_off, cell_off, ncell + 1));
    CU(upload(&e->d_nl_sites, cell_sites, (size_t)std::max(cell_off[ncell], 1)));
    for (int i = 0; i < 9; i++) e->nl_l0inv[i] = lattice0_inverse[i];
    e->nl_G = G; e->nl_rho = rho; e->nl_set = true;
    return SAB_OK;
}

extern "C" int sab_set_dynvor_lists(sab_engine *e, int G, const double *lattice0, const double *lattice0_inverse,
                                    const double *anchors, double max_move, const int32_t *cell_off,
                                    const uint16_t *cell_sites, const double *excl)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_dynvor_lists: null engine");
    CU(cudaSetDevice(e->device));
    if (!cell_off) { e->nl_set = false; return SAB_OK; }
    if (e->kind != SAB_DYNAMIC_VORONOI || !e->d_slot_atom) return fail(SAB_ESTATE, "sab_set_dynvor_lists: reference atoms not set");
    if (G < 1 || G > 256 || !lattice0 || !lattice0_inverse || !anchors || !cell_sites || !excl || e->S > 65535 || !(max_move >= 0.0))
        return fail(SAB_EINVAL, "sab_set_dynvor_lists: bad arguments");
    size_t ncell = (size_t)G * G * G;
    CU(upload(&e->d_nl_off, cell_off, ncell + 1));
    CU(upload(&e->d_nl_sites, cell_sites, (size_t)std::max(cell_off[ncell], 1)));
    CU(upload(&e->d_nl_excl, excl, ncell));
    CU(upload(&e->d_nl_anchor, anchors, (size_t)3 * e->S));
    CU(upload(&e->d_nl_l0, lattice0, (size_t)9));
    for (int i = 0; i < 9; i++) e->nl_l0inv[i] = lattice0_inverse[i];
    e->nl_G = G; e->nl_dmax = max_move; e->nl_set = true;
    return SAB_OK;
}

extern "C" int sab_set_option(sab_engine *e, int key, int64_t value)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_option: null engine");
    if (key == SAB_OPT_SKIP_STEP_CHECK_ONCE) { e->skip_step_check_once = value != 0; return SAB_OK; }
    if (key == SAB_OPT_POLY_KERNEL) { e->poly_kernel = (int)value; return SAB_OK; }
    if (key == SAB_OPT_TILE_FRAMES) { e->tile_frames = (int)value; return SAB_OK; }
    if (key == SAB_OPT_STREAM_REC_SMEM) { e->stream_rec_smem = value != 0; return SAB_OK; }
    if (key == SAB_OPT_STREAM_REG_COLS) { e->stream_reg_cols = value != 0; return SAB_OK; }
    if (key == SAB_OPT_STREAM_PRODUCERS) { if (value < 0 || value > 8) return fail(SAB_EINVAL, "producer warps: 0..8"); e->stream_prod = (int)value; return SAB_OK; }
    if (key == SAB_OPT_FP32_FILTER) { e->fp32_filter = value != 0; return SAB_OK; }
    if (key == SAB_OPT_STREAM_GEN) { e->stream_gen = (int)value; return SAB_OK; }
    if (key == SAB_OPT_STREAM5_ROWS) { if (value < 0 || value > SAB_S5_MAXROWS) return fail(SAB_EINVAL, "row slots: 0..%d", SAB_S5_MAXROWS); e->s5_rows = (int)value; return SAB_OK; }
    if (key == SAB_OPT_STREAM5_RAW) { if (value < 0 || value > SAB_S5_MAXRAW || value == 1) return fail(SAB_EINVAL, "raw slots: 0 or 2..%d", SAB_S5_MAXRAW); e->s5_raw = (int)value; return SAB_OK; }
    if (key == SAB_OPT_RESOLVER) { if (value < 0 || value > 2) return fail(SAB_EINVAL, "resolver: 0..2"); e->resolver = (int)value; return SAB_OK; }
    if (key == SAB_OPT_STREAM5_BLOCKS) { if (value < 0 || value > 8) return fail(SAB_EINVAL, "CTAs per SM: 0..8"); e->s5_blocks = (int)value; return SAB_OK; }
    if (key == SAB_OPT_SEQ_BLOCK) { e->seq_block = (int)value; return SAB_OK; }
    if (key == SAB_OPT_SEQ_CHUNK) { if (value < 1) return fail(SAB_EINVAL, "chunk must be >= 1"); e->seq_chunk = (int)value; return SAB_OK; }
    return fail(SAB_EINVAL, "unknown option %d", key);
}

// ---------------------------------------------------------------- workspace layout
struct Workspace {
    int n_chunks = 0; int m3 = 0;
    int64_t exc_chunks = 0;                                   // capacity of chunk_exc in chunks (>= n_chunks; stream kernel: up to 2048)
    int32_t *totals = nullptr; double *chunk_exc = nullptr; int32_t *rows = nullptr;
    long long *defer = nullptr; long long defer_cap = 0;      // deferred atom-frames of the tile kernel
    long long *flagged = nullptr;                             // [F] frames the generation-5 stream kernel hands to the generic pass
    size_t bytes = 0;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static Workspace layout(const sab_engine *e, int64_t F, void *base)
{
    Workspace w;
    w.m3 = 3 * e->M;
    if (w.m3 == 0 || F == 0) { w.bytes = 256; return w; }
    w.n_chunks = (int)((F + SAB_WRAP_CHUNK - 1) / SAB_WRAP_CHUNK);
    size_t off = 0;
    char *b = (char *)base;
    w.totals = (int32_t *)(b + off); off += align256(sizeof(int32_t) * (size_t)w.n_chunks * w.m3);
    w.exc_chunks = std::max<int64_t>(w.n_chunks, std::min<int64_t>(F, 2048));
    w.chunk_exc = (double *)(b + off); off += align256(sizeof(double) * 2 * (size_t)w.exc_chunks * w.m3);
    w.rows = (int32_t *)(b + off); off += align256(sizeof(int32_t) * (size_t)F * w.m3);
    if (e->kind == SAB_POLYHEDRAL) {
        w.defer_cap = std::max<long long>(4096, (long long)F * e->A / 16);
        w.defer = (long long *)(b + off); off += align256(sizeof(long long) * (size_t)w.defer_cap);
        w.flagged = (long long *)(b + off); off += align256(sizeof(long long) * (size_t)F);
    }
    w.bytes = off + 256;
    return w;
}

extern "C" int64_t sab_workspace_bytes(sab_engine *e, int64_t n_frames)
{
    if (!e || n_frames < 0) return -1;
    return (int64_t)layout(e, n_frames, nullptr).bytes;
}

static int run_wrap_pass(sab_engine *e, const double *d_frac, int64_t F, const Workspace &w, cudaStream_t st, bool rows_and_commit)
{
    dim3 grid(w.n_chunks, (w.m3 + 127) / 128);
    k_wrap_totals<<<grid, 128, 0, st>>>(d_frac, F, e->N, e->d_fw_atoms, w.m3, e->d_prev_raw, w.totals,
                                        e->d_counters + CNT_FIRST_BAD, e->skip_step_check_once ? 1 : 0);
    e->launches++;
    if (!rows_and_commit) return SAB_OK;
    k_wrap_bases<<<(w.m3 + 127) / 128, 128, 0, st>>>(w.totals, w.n_chunks, w.m3, e->d_wraps);
    k_wrap_rows<<<grid, 128, 0, st>>>(d_frac, F, e->N, e->d_fw_atoms, w.m3, e->d_prev_raw, w.totals, w.rows, w.chunk_exc);
    e->launches += 2;
    return SAB_OK;
}

extern "C" int sab_wrap_totals(sab_engine *e, const double *d_frac, int64_t n_frames, void *d_workspace,
                               int64_t workspace_bytes, void *stream, int32_t *totals, double *last_raw)
{
    if (!e || !d_frac || !totals || !last_raw) return fail(SAB_EINVAL, "sab_wrap_totals: null argument");
    if (!e->pbc_set) return fail(SAB_ESTATE, "sab_wrap_totals: PBC state not set");
    if (n_frames <= 0) return fail(SAB_EINVAL, "sab_wrap_totals: no frames");
    CU(cudaSetDevice(e->device));
    Workspace w = layout(e, n_frames, d_workspace);
    if ((int64_t)w.bytes > workspace_bytes) return fail(SAB_EINVAL, "workspace too small: need %zu bytes", w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long init[CNT_N] = {0, 0, 0, ~0ULL, 0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(e->d_counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    run_wrap_pass(e, d_frac, n_frames, w, st, false);
    CU(cudaGetLastError());
    // reduce on the device; the chunk-excursion area of the workspace is free scratch here
    double *d_last = w.chunk_exc;
    int32_t *d_tot = reinterpret_cast<int32_t *>(w.chunk_exc + w.m3);
    k_wrap_reduce<<<(w.m3 + 127) / 128, 128, 0, st>>>(d_frac, n_frames, e->N, e->d_fw_atoms, w.m3, w.n_chunks, w.totals, d_tot, d_last);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(totals, d_tot, sizeof(int32_t) * w.m3, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(last_raw, d_last, sizeof(double) * w.m3, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SAB_OK;
}

extern "C" int sab_scan_excursion(sab_engine *e, const double *d_frac, int64_t n_frames, void *d_workspace,
                                  int64_t workspace_bytes, void *stream, double *excursion)
{
    if (!e || !d_frac || !excursion || n_frames <= 0) return fail(SAB_EINVAL, "sab_scan_excursion: bad arguments");
    if (!e->pbc_set) return fail(SAB_ESTATE, "sab_scan_excursion: PBC state not set");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w = layout(e, n_frames, d_workspace);
    if ((int64_t)w.bytes > workspace_bytes) return fail(SAB_EINVAL, "workspace too small: need %zu bytes", w.bytes);
    unsigned long long init[CNT_N] = {0, 0, 0, ~0ULL, 0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(e->d_counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    bool skip = e->skip_step_check_once;
    run_wrap_pass(e, d_frac, n_frames, w, st, true);
    e->skip_step_check_once = skip;
    CU(cudaGetLastError());
    std::vector<double> h((size_t)w.n_chunks * w.m3 * 2);
    CU(cudaMemcpyAsync(h.data(), w.chunk_exc, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int c = 0; c < w.m3; c++) {
        double lo = INFINITY, hi = -INFINITY;
        for (int k = 0; k < w.n_chunks; k++) {
            lo = std::min(lo, h[((size_t)k * w.m3 + c) * 2]);
            hi = std::max(hi, h[((size_t)k * w.m3 + c) * 2 + 1]);
        }
        excursion[2 * c] = lo; excursion[2 * c + 1] = hi;
    }
    return SAB_OK;
}

// ---------------------------------------------------------------- stream kernel (tetrahedral sites, fused wrap scan)
struct StreamPlan { bool ok = false; int n_prod = 0, n_cons = 0, n_raw = 0, use_tma = 0, rec_smem = 0; size_t smem = 0, stride = 0; int64_t chunk = 0; int grid = 0; };

static StreamPlan plan_stream(sab_engine *e, const double *d_frac, int64_t F, const Workspace &w)
{
    StreamPlan p;
    if (e->kind != SAB_POLYHEDRAL || !e->cells_set || !e->tile_ok || e->poly_kernel != 0 || e->force_scan || !w.defer) return p;
    p.n_cons = std::min(6, (e->A + 31) / 32);
    p.n_prod = e->stream_prod > 0 ? e->stream_prod : std::max(2, (2 * p.n_cons + 2) / 3);   // measured: 4 producers for 6 consumers
    if ((p.n_prod + p.n_cons) * 32 > SAB_STREAM_MAXTHREADS) p.n_prod = SAB_STREAM_MAXTHREADS / 32 - p.n_cons;
    if (p.n_prod < 1) return p;
    const size_t n3 = 3 * (size_t)e->N, m3 = 3 * (size_t)e->M, a3 = 3 * (size_t)e->A;
    p.stride = ((size_t)e->n_used + a3 + 1 + 3) & ~(size_t)3;         // at least one spare float behind the mobile coordinates
    p.use_tma = ((n3 * 8) % 16 == 0 && (reinterpret_cast<uintptr_t>(d_frac) & 15) == 0 && n3 * 8 < (1u << 20)) ? 1 : 0;
    auto smem_for = [&](int n_raw, int rec_smem) {
        return sizeof(double) * ((size_t)n_raw * n3 + 3 * m3) + sizeof(float) * (SAB_STREAM_ROWS * p.stride + (((size_t)e->n_used + 1) & ~(size_t)1)) +
               sizeof(int) * (2 * m3 + ((m3 + a3 + 1) & ~(size_t)1) + 2 * (size_t)e->A) + (rec_smem ? (size_t)e->S * 24 : 0) + 128;
    };
    const size_t budget = (size_t)(226 * 1024 / SAB_STREAM_MINBLOCKS);
    p.n_raw = SAB_STREAM_MAXRAW;
    p.rec_smem = (e->stream_rec_smem && smem_for(3, 1) <= budget) ? 1 : 0;               // site records next to the rows when they fit
    while (p.n_raw > 3 && smem_for(p.n_raw, p.rec_smem) > budget) p.n_raw--;
    p.smem = smem_for(p.n_raw, p.rec_smem);
    if (p.smem > 226 * 1024) return p;
    const size_t key = (p.smem << 13) | (size_t)((p.n_prod + p.n_cons) * 32) << 1 | (size_t)p.rec_smem;
    if (key != e->stream_occ_key) {
        const void *fn = p.rec_smem ? (const void *)k_polyhedral_assign_stream<true> : (const void *)k_polyhedral_assign_stream<false>;
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess) { cudaGetLastError(); return p; }
        int q = 1;
        if (p.rec_smem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_polyhedral_assign_stream<true>, (p.n_prod + p.n_cons) * 32, p.smem);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_polyhedral_assign_stream<false>, (p.n_prod + p.n_cons) * 32, p.smem);
        e->stream_per_sm = q; e->stream_occ_key = key;
    }
    const int per_sm = e->stream_per_sm;
    const int64_t resident = (int64_t)std::max(1, per_sm) * e->sm_count;
    // one contiguous chunk per resident CTA (the history-free first frame of a chunk costs a full search); the
    // per-chunk excursion records go to the workspace's chunk_exc area
    p.chunk = e->seq_chunk > 0 ? e->seq_chunk : std::max<int64_t>(1, (F + resident - 1) / resident);
    const int64_t n_chunks = (F + p.chunk - 1) / p.chunk;
    if (n_chunks > w.exc_chunks) return p;
    p.grid = (int)std::min<int64_t>(n_chunks, resident);
    p.ok = true;
    return p;
}

static int launch_stream(sab_engine *e, const StreamPlan &p, const double *d_frac, int64_t F, int32_t *d_result,
                         int32_t *d_spill, int64_t spill_cap, const Workspace &w, cudaStream_t st)
{
    SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp, e->d_legacy_init};
    SabStreamTables C{e->n_used, (int)p.stride, 3 * e->M, e->d_tile_vars, e->d_tile_rec, e->G, e->d_cell_off, e->d_cell_sites,
                      (float)e->delta - 4e-7f, e->fp32_filter, w.defer, w.defer_cap, e->d_anchor, e->d_fw_atoms, e->d_prev_raw,
                      e->d_wraps, e->skip_step_check_once ? 1 : 0, p.n_raw, p.n_prod, p.use_tma, w.chunk_exc, e->delta > 0.12 ? 1 : 0,
                      (e->stream_reg_cols && 3 * e->M <= SAB_STREAM_KMAX * 32 * p.n_prod && 3 * e->A <= SAB_STREAM_KMAX * 32 * p.n_prod &&
                       e->n_used < 65536) ? 1 : 0, e->peer_mc ? 0 : e->n_peer, {nullptr}};
    const int n_peer4 = e->peer_mc ? 0 : e->n_peer;          // the previous-generation kernel only knows unicast peers
    SabPeers peers{n_peer4, {nullptr}, 0};
    for (int q = 0; q < n_peer4; q++) { C.peer[q] = e->peer[q]; peers.p[q] = e->peer[q]; }
    e->peers_written = n_peer4 > 0;
    const int64_t n_chunks = (F + p.chunk - 1) / p.chunk;
    if (p.rec_smem)
        k_polyhedral_assign_stream<true><<<p.grid, (p.n_prod + p.n_cons) * 32, p.smem, st>>>(d_frac, F, p.chunk, e->N, e->A, e->d_mobile, e->S, T, C,
            e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_counters + CNT_FULL27);
    else
        k_polyhedral_assign_stream<false><<<p.grid, (p.n_prod + p.n_cons) * 32, p.smem, st>>>(d_frac, F, p.chunk, e->N, e->A, e->d_mobile, e->S, T, C,
            e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_counters + CNT_FULL27);
    // exact float64 pass over the atom-frames float32 could not decide (returns at once when there are none)
    k_polyhedral_deferred<<<e->sm_count * 8, 256, 0, st>>>(d_frac, F, e->N, e->A, e->d_mobile, T, e->G, e->d_cell_off,
        e->d_cell_sites, nullptr, 3 * e->M, w.defer, w.defer_cap, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_anchor, peers);
    k_wrap_commit_stream<<<(w.m3 + 31) / 32, 32 * SAB_COMMIT_GROUPS, 0, st>>>(d_frac, F, e->N, e->d_fw_atoms, w.m3, (int)n_chunks, w.chunk_exc, e->d_anchor,
        e->d_prev_raw, e->d_wraps, e->d_excursion, e->d_counters);
    e->launches += 3;
    CU(cudaGetLastError());
    return SAB_OK;
}

// ---------------------------------------------------------------- generation-5 stream kernel (vertex-image rows)
struct StreamPlan5 { bool ok = false; int n_prod = 0, n_cons = 0, n_raw = 0, n_rows = 0, use_tma = 0, rec_smem = 0, one_group = 0;
                     size_t smem = 0; int stride4 = 0; int64_t chunk = 0; int grid = 0; };

template <bool OG, bool RS> static const void *stream5_fn() { return (const void *)k_polyhedral_assign_stream5<OG, RS>; }
static const void *stream5_kernel(int one_group, int rec_smem)
{
    return one_group ? (rec_smem ? stream5_fn<true, true>() : stream5_fn<true, false>())
                     : (rec_smem ? stream5_fn<false, true>() : stream5_fn<false, false>());
}

static StreamPlan5 plan_stream5(sab_engine *e, const double *d_frac, int64_t F, const Workspace &w)
{
    StreamPlan5 p;
    if (e->kind != SAB_POLYHEDRAL || !e->cells_set || !e->s5_ok || e->poly_kernel != 0 || e->stream_gen == 4 || e->force_scan || !w.defer) return p;
    p.n_cons = std::min(6, (e->A + 31) / 32);
    p.n_prod = e->stream_prod > 0 ? e->stream_prod : std::max(1, (p.n_cons + 2) / 3);      // measured: 2 producers for 6 consumers
    if ((p.n_prod + p.n_cons) * 32 > SAB_S5_MAXTHREADS) p.n_prod = SAB_S5_MAXTHREADS / 32 - p.n_cons;
    if (p.n_prod < 1) return p;
    p.one_group = e->A <= p.n_cons * 32 ? 1 : 0;
    const size_t n3 = 3 * (size_t)e->N;
    p.stride4 = e->M + e->A + 2;                                          // flag entry in front, one spare float4 behind the mobile atoms
    p.use_tma = ((n3 * 8) % 16 == 0 && (reinterpret_cast<uintptr_t>(d_frac) & 15) == 0 && n3 * 8 < (1u << 20)) ? 1 : 0;
    auto smem_for = [&](int n_raw, int n_rows, int rec_smem) {
        return sizeof(double) * (((size_t)n_raw * n3 + 1) & ~(size_t)1) + 16 * (size_t)n_rows * p.stride4 + 16 * (size_t)e->M +
               (rec_smem ? 8 * (size_t)e->S : 0) + (p.one_group ? 0 : 8 * (size_t)e->A) + 4 * (size_t)e->A + 128;
    };
    // shared-memory plan: as many resident CTAs per SM as SAB_S5_MINBLOCKS allows, rings as deep as then fit
    static const int shapes[][2] = {{3, 5}, {3, 4}, {3, 3}, {2, 3}, {2, 2}};   // {raw slots, row slots}, preference order (measured)
    static const int shapes_big[][2] = {{4, 6}, {3, 6}, {3, 4}, {3, 3}, {2, 3}, {2, 2}};
    bool found = false;
    for (int blocks = e->s5_blocks > 0 ? e->s5_blocks : SAB_S5_MINBLOCKS; blocks >= 1 && !found; blocks--) {
        const size_t budget = (size_t)(228 * 1024) / blocks - 1024 - 320;   // 1 KB per CTA is reserved by the system; static barriers
        const int (*sh)[2] = blocks >= 3 ? shapes : shapes_big;
        const int n_sh = blocks >= 3 ? 5 : 6;
        for (int i = 0; i < n_sh && !found; i++)
            for (int rec = e->stream_rec_smem ? 1 : 0; rec >= 0 && !found; rec--) {      // deeper rings before records in shared memory
                const int raw = e->s5_raw > 0 ? e->s5_raw : sh[i][0], rows = e->s5_rows > 0 ? e->s5_rows : sh[i][1];
                if (smem_for(raw, rows, rec) <= std::min(budget, (size_t)(227 * 1024))) { p.n_raw = raw; p.n_rows = rows; p.rec_smem = rec; found = true; }
            }
    }
    if (!found) return p;
    p.smem = smem_for(p.n_raw, p.n_rows, p.rec_smem);
    const int threads = (p.n_prod + p.n_cons) * 32;
    const size_t key = (p.smem << 14) | (size_t)threads << 2 | (size_t)p.rec_smem << 1 | (size_t)p.one_group;
    if (key != e->s5_occ_key) {
        const void *fn = stream5_kernel(p.one_group, p.rec_smem);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess) { cudaGetLastError(); return p; }
        int q = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, fn, threads, p.smem) != cudaSuccess) { cudaGetLastError(); return p; }
        e->s5_per_sm = q; e->s5_occ_key = key;
    }
    const int64_t resident = (int64_t)std::max(1, e->s5_per_sm) * e->sm_count;
    p.chunk = e->seq_chunk > 0 ? e->seq_chunk : std::max<int64_t>(1, (F + resident - 1) / resident);
    const int64_t n_chunks = (F + p.chunk - 1) / p.chunk;
    if (n_chunks > w.exc_chunks) return p;
    p.grid = (int)std::min<int64_t>(n_chunks, resident);
    p.ok = true;
    return p;
}

static int launch_stream5(sab_engine *e, const StreamPlan5 &p, const double *d_frac, int64_t F, int32_t *d_result,
                          int32_t *d_spill, int64_t spill_cap, const Workspace &w, cudaStream_t st)
{
    SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp, e->d_legacy_init};
    SabStream5Tables C{};
    C.n_task = e->M; C.stride4 = p.stride4; C.m3 = 3 * e->M;
    C.task = e->d_s5_task; C.rec = e->d_s5_rec;
    C.G = e->G; C.cell_off = e->d_cell_off; C.cell_sites = e->d_cell_sites; C.cell_rec = e->d_s5_cellrec;
    C.delta_t = (float)(e->delta - 1e-6);
    C.reg_tasks = (e->M <= SAB_S5_KT * 32 * p.n_prod && e->A <= SAB_S5_KT * 32 * p.n_prod) ? 1 : 0;
    C.T1 = e->s5_T1; C.T0 = e->s5_T0; C.TD = e->s5_TD;
    C.use_f32 = e->fp32_filter; C.defer = w.defer; C.defer_cap = w.defer_cap; C.flagged = w.flagged;
    C.anchor = e->d_anchor; C.fw_atoms = e->d_fw_atoms; C.prev_raw = e->d_prev_raw; C.wraps = e->d_wraps;
    C.skip_check_frame0 = e->skip_step_check_once ? 1 : 0;
    C.n_raw = p.n_raw; C.n_rows = p.n_rows; C.n_prod = p.n_prod; C.use_tma = p.use_tma;
    C.chunk_exc = w.chunk_exc; C.delta_d = e->delta; C.n_peer = e->n_peer; C.peer_mc = e->peer_mc;
    SabPeers peers{e->n_peer, {nullptr}, e->peer_mc};
    for (int q = 0; q < e->n_peer; q++) { C.peer[q] = e->peer[q]; peers.p[q] = e->peer[q]; }
    e->peers_written = e->n_peer > 0;
    const int64_t n_chunks = (F + p.chunk - 1) / p.chunk;
    const int threads = (p.n_prod + p.n_cons) * 32;
#define SAB_LAUNCH5(OG, RS) k_polyhedral_assign_stream5<OG, RS><<<p.grid, threads, p.smem, st>>>(d_frac, F, p.chunk, e->N, e->A, e->d_mobile, \
        e->S, T, C, e->legacy_init_frame, d_result, e->d_counters, e->d_counters + CNT_FULL27)
    if (p.one_group) { if (p.rec_smem) SAB_LAUNCH5(true, true); else SAB_LAUNCH5(true, false); }
    else { if (p.rec_smem) SAB_LAUNCH5(false, true); else SAB_LAUNCH5(false, false); }
#undef SAB_LAUNCH5
    // frames outside the validity tube: generic all-sites pass (returns at once when there are none)
    k_polyhedral_flagged<<<e->sm_count * 4, 128, 0, st>>>(d_frac, e->N, e->A, e->d_mobile, e->S, T, e->legacy_init_frame, w.flagged,
        d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_anchor, peers);
    // exact float64 pass over the atom-frames float32 could not decide (returns at once when there are none)
    k_polyhedral_deferred<<<e->sm_count * 8, 256, 0, st>>>(d_frac, F, e->N, e->A, e->d_mobile, T, e->G, e->d_cell_off,
        e->d_cell_sites, nullptr, 3 * e->M, w.defer, w.defer_cap, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_anchor, peers);
    k_wrap_commit_stream<<<(w.m3 + 31) / 32, 32 * SAB_COMMIT_GROUPS, 0, st>>>(d_frac, F, e->N, e->d_fw_atoms, w.m3, (int)n_chunks, w.chunk_exc, e->d_anchor,
        e->d_prev_raw, e->d_wraps, e->d_excursion, e->d_counters);
    e->launches += 4;
    CU(cudaGetLastError());
    return SAB_OK;
}

static int launch_assign(sab_engine *e, const double *d_frac, const double *d_lattice, int lattice_stride, int64_t F,
                         int32_t *d_result, int32_t *d_spill, int64_t spill_cap, const Workspace &w, cudaStream_t st,
                         double *d_centres_out)
{
    int grid = (int)std::min<int64_t>(F, (int64_t)e->sm_count * 8);
    switch (e->kind) {
    case SAB_VORONOI: if (e->nl_set) {
        int block = std::min(256, ((e->A + 31) / 32) * 32);
        SabNearestLists NL{e->nl_G, e->d_nl_off, e->d_nl_sites, e->d_nl_excl, {0}};
        for (int i = 0; i < 9; i++) NL.l0inv[i] = e->nl_l0inv[i];
        int grid2 = (int)std::min<int64_t>(F, (int64_t)e->sm_count * 16);
        k_nearest_assign_cells<<<grid2, block, 0, st>>>(d_frac, d_lattice, lattice_stride, F, e->N, e->A, e->d_mobile, e->S,
            e->d_centres, NL, d_result, e->d_counters + CNT_FULL27);
        break;
    } else {
        int block = std::min(512, ((e->A + 31) / 32) * 32);
        size_t smem = sizeof(double) * 3 * (size_t)e->S;
        CU(cudaFuncSetAttribute(k_nearest_assign<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nearest_assign<false><<<grid, block, smem, st>>>(d_frac, d_lattice, lattice_stride, F, e->N, e->A, e->d_mobile,
            e->S, e->d_centres, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, -1, nullptr, d_result,
            e->d_counters + CNT_FULL27);
        break;
    }
    case SAB_DYNAMIC_VORONOI: if (e->nl_set && d_result && !d_centres_out) {
        int block = 256;
        size_t smem = sizeof(double) * 3 * (size_t)e->S;
        CU(cudaFuncSetAttribute(k_dynvor_assign_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SabNearestLists NL{e->nl_G, e->d_nl_off, e->d_nl_sites, e->d_nl_excl, {0}};
        for (int i = 0; i < 9; i++) NL.l0inv[i] = e->nl_l0inv[i];
        int grid2 = (int)std::min<int64_t>(F, (int64_t)e->sm_count * 8);
        k_dynvor_assign_cells<<<grid2, block, smem, st>>>(d_frac, d_lattice, lattice_stride, F, e->N, e->A, e->d_mobile, e->S,
            e->smax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, w.rows, w.m3, e->d_legacy_init,
            e->legacy_init_frame, NL, e->d_nl_anchor, e->d_nl_l0, e->nl_dmax, d_result, e->d_counters + CNT_FULL27);
        break;
    } else {
        int block = std::min(256, ((std::max(e->A, 128) + 31) / 32) * 32);   // 256 x <= 255 registers always fits
        size_t smem = sizeof(double) * 3 * (size_t)e->S;
        CU(cudaFuncSetAttribute(k_nearest_assign<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nearest_assign<true><<<grid, block, smem, st>>>(d_frac, d_lattice, lattice_stride, F, e->N, e->A, e->d_mobile,
            e->S, nullptr, e->smax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, w.rows, w.m3,
            e->d_legacy_init, e->legacy_init_frame, d_centres_out, d_result, e->d_counters + CNT_FULL27);
        break;
    }
    case SAB_SPHERICAL: if (e->nl_set) {
        int block = std::min(256, ((e->A + 31) / 32) * 32);
        SabSphereLists SL{e->nl_G, e->d_nl_off, e->d_nl_sites, e->nl_rho, {0}};
        for (int i = 0; i < 9; i++) SL.l0inv[i] = e->nl_l0inv[i];
        int grid2 = (int)std::min<int64_t>(F, (int64_t)e->sm_count * 16);
        k_spherical_assign_cells<<<grid2, block, 0, st>>>(d_frac, d_lattice, lattice_stride, F, e->N, e->A, e->d_mobile, e->S,
            e->d_centres, e->d_rcut, e->d_rcut_q, SL, d_result, d_spill, (long long)spill_cap, e->d_counters,
            e->d_counters + CNT_FULL27);
        break;
    } else {
        int block = std::min(512, ((e->A + 31) / 32) * 32);
        size_t smem = sizeof(double) * 5 * (size_t)e->S;
        CU(cudaFuncSetAttribute(k_spherical_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_spherical_assign<<<grid, block, smem, st>>>(d_frac, d_lattice, lattice_stride, F, e->N, e->A, e->d_mobile, e->S,
            e->d_centres, e->d_rcut, e->d_rcut_q, d_result, d_spill, (long long)spill_cap, e->d_counters);
        break;
    }
    case SAB_POLYHEDRAL: if (e->cells_set && (e->poly_kernel == 0 || e->poly_kernel == 5) && e->tile_ok) {
        // tile kernel: frames per tile chosen so that SAB_TILE_MINBLOCKS CTAs stay resident per SM
        const int block = SAB_TILE_THREADS;
        const size_t stride = ((size_t)e->n_used + 3 * (size_t)e->A + 3) & ~(size_t)3;
        auto smem_for = [&](int t) {
            return sizeof(double) * (size_t)t * 3 * e->N + sizeof(int32_t) * (((size_t)t * 3 * e->M + 3) & ~(size_t)3) +
                   sizeof(float) * (size_t)t * stride +
                   sizeof(int) * ((size_t)3 * SAB_TILE_BLOCKS + SAB_TILE_SINGLES + (size_t)e->A * (4 + SAB_POLY_KEEP)) + 128;
        };
        int tile = (e->tile_frames >= 1 && e->tile_frames <= SAB_TILE_MAX) ? e->tile_frames : SAB_TILE_MAX;
        if (e->tile_frames <= 0)
            while (tile > 1 && smem_for(tile) > (size_t)(226 * 1024 / SAB_TILE_MINBLOCKS)) tile--;
        if (smem_for(tile) <= 226 * 1024 && (size_t)e->A * (tile - 1) <= SAB_TILE_SINGLES && w.defer) {
            const size_t smem = smem_for(tile);
            CU(cudaFuncSetAttribute(k_polyhedral_assign_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp,
                            e->d_legacy_init};
            SabTileTables C{e->n_used, (int)stride, 3 * e->M, e->d_tile_vars, e->d_tile_rec, e->G, e->d_cell_off, e->d_cell_sites,
                            (float)e->delta - 4e-7f, e->fp32_filter, w.defer, w.defer_cap};
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_polyhedral_assign_tile, block, smem);
            const int64_t resident = (int64_t)std::max(1, per_sm) * e->sm_count;
            // one contiguous chunk per resident CTA (the history-free first frame of a chunk costs a full search)
            int64_t chunk = e->seq_chunk > 0 ? e->seq_chunk : std::max<int64_t>(tile, (F + resident - 1) / resident);
            if (e->seq_chunk <= 0) chunk = (chunk + tile - 1) / tile * tile;
            const int64_t n_chunks = (F + chunk - 1) / chunk;
            const int grid2 = (int)std::min<int64_t>(n_chunks, resident);
            k_polyhedral_assign_tile<<<grid2, block, smem, st>>>(d_frac, F, (int)chunk, tile, e->N, e->A, e->d_mobile, e->S, T, C,
                w.rows, e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_counters + CNT_FULL27);
            // exact float64 pass over the atom-frames float32 could not decide (returns at once when there are none)
            k_polyhedral_deferred<<<e->sm_count * 8, 256, 0, st>>>(d_frac, F, e->N, e->A, e->d_mobile, T, e->G, e->d_cell_off,
                e->d_cell_sites, w.rows, 3 * e->M, w.defer, w.defer_cap, d_result, d_spill, (long long)spill_cap, e->d_counters);
            e->launches++;
            break;
        }
    }
    if (e->cells_set && (e->poly_kernel == 0 || e->poly_kernel == 4) && e->A <= SAB_SEQ1_THREADS * SAB_SEQ1_MAX_OWN) {
        size_t smem = sizeof(double) * ((size_t)e->nk * 3 * e->M + 3 * (size_t)e->A) +
                      sizeof(float) * ((size_t)e->nk * 3 * e->M + 2) +
                      sizeof(int) * ((size_t)e->A * (6 + SAB_POLY_KEEP) + 1);
#ifdef SAB_SEQ1_PREFETCH
        smem += sizeof(int) * (3 * (size_t)e->M + 4) + sizeof(double) * (3 * (size_t)e->M + 3 * (size_t)e->A) + 16;
#endif
        CU(cudaFuncSetAttribute(k_polyhedral_assign_seq1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp,
                        e->d_legacy_init};
        SabTetTables C{e->M, 3 * e->M, e->kmin, e->nk, e->d_fw_atoms, e->d_rec, e->G, e->d_cell_off, e->d_cell_sites,
                       e->d_anchor, e->delta, e->fp32_filter};
        int per_sm = 3;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_polyhedral_assign_seq1, SAB_SEQ1_THREADS, smem);
        int resident = std::max(1, per_sm) * e->sm_count;
        // chunk: long enough to amortise the history-free first frame, short enough to balance the CTAs
        int chunk = (int)std::max<int64_t>(8, std::min<int64_t>(e->seq_chunk > 0 ? e->seq_chunk : 32, F / ((int64_t)resident * 4)));
        int64_t n_chunks = (F + chunk - 1) / chunk;
        int grid2 = (int)std::min<int64_t>(n_chunks, resident);
        k_polyhedral_assign_seq1<<<grid2, SAB_SEQ1_THREADS, smem, st>>>(d_frac, F, chunk, e->N, e->A, e->d_mobile, e->S, T, C,
            w.rows, e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_counters + CNT_FULL27);
        break;
    } else if (e->cells_set && e->poly_kernel == 3 && e->A <= SAB_SEQ_THREADS) {
        // chunk-sequential kernel: one thread per mobile atom
        int block = std::max(64, ((e->A + 31) / 32) * 32);
        if (e->seq_block >= block && e->seq_block <= SAB_SEQ_THREADS) block = e->seq_block;
        size_t m3 = 3 * (size_t)e->M;
        size_t smem = sizeof(double) * ((size_t)e->nk * m3 + m3 + 6 * (size_t)e->A) +
                      sizeof(int) * (m3 + (size_t)e->A * (4 + SAB_POLY_KEEP));
        CU(cudaFuncSetAttribute(k_polyhedral_assign_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp,
                        e->d_legacy_init};
        SabTetTables C{e->M, 3 * e->M, e->kmin, e->nk, e->d_fw_atoms, e->d_rec, e->G, e->d_cell_off, e->d_cell_sites,
                       e->d_anchor, e->delta, e->fp32_filter};
        // chunk: long enough to amortise the history-free first frame, short enough to fill the GPU
        int chunk = (int)std::max<int64_t>(8, std::min<int64_t>(e->seq_chunk > 0 ? e->seq_chunk : 32, F / ((int64_t)e->sm_count * 8)));
        int64_t n_chunks = (F + chunk - 1) / chunk;
        int grid2 = (int)std::min<int64_t>(n_chunks, (int64_t)e->sm_count * 8);
        k_polyhedral_assign_seq<<<grid2, block, smem, st>>>(d_frac, F, chunk, e->N, e->A, e->d_mobile, e->S, T, C,
            w.rows, e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_counters + CNT_FULL27);
        break;
    } else if (e->cells_set) {
        int block = std::min(256, ((e->A + 31) / 32) * 32);
        size_t smem = sizeof(double) * (size_t)e->nk * 3 * e->M;
        CU(cudaFuncSetAttribute(k_polyhedral_assign_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp,
                        e->d_legacy_init};
        SabTetTables C{e->M, 3 * e->M, e->kmin, e->nk, e->d_fw_atoms, e->d_rec, e->G, e->d_cell_off, e->d_cell_sites,
                       e->d_anchor, e->delta, e->fp32_filter};
        int grid2 = (int)std::min<int64_t>(F, (int64_t)e->sm_count * 16);
        k_polyhedral_assign_cells<<<grid2, block, smem, st>>>(d_frac, F, e->N, e->A, e->d_mobile, e->S, T, C, w.rows,
            e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters, e->d_counters + CNT_FULL27);
        break;
    } else {
        int block = 256;
        size_t smem = sizeof(double) * 3 * (size_t)e->A + sizeof(int) * (size_t)e->A * (2 + SAB_POLY_KEEP);
        CU(cudaFuncSetAttribute(k_polyhedral_assign_exhaustive, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SabPolyTables T{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_slot_fw, e->d_slot_shift, e->d_nface, e->d_simp,
                        e->d_legacy_init};
        k_polyhedral_assign_exhaustive<<<grid, block, smem, st>>>(d_frac, F, e->N, e->A, e->d_mobile, e->S, T, w.rows, w.m3,
            e->legacy_init_frame, d_result, d_spill, (long long)spill_cap, e->d_counters);
        break;
    }
    }
    e->launches++;
    CU(cudaGetLastError());
    return SAB_OK;
}

static int check_ready(sab_engine *e)
{
    if ((e->kind == SAB_SPHERICAL || e->kind == SAB_VORONOI) && !e->d_centres) return fail(SAB_ESTATE, "site centres not set");
    if (e->kind == SAB_POLYHEDRAL && !e->d_simp) return fail(SAB_ESTATE, "polyhedra not set");
    if (e->kind == SAB_DYNAMIC_VORONOI && !e->d_slot_atom) return fail(SAB_ESTATE, "reference atoms not set");
    if ((e->kind == SAB_POLYHEDRAL || e->kind == SAB_DYNAMIC_VORONOI) && !e->pbc_set) return fail(SAB_ESTATE, "PBC state not set");
    return SAB_OK;
}

extern "C" int sab_assign_device(sab_engine *e, const double *d_frac, const double *d_lattice, int lattice_stride,
                                 int64_t n_frames, int32_t *d_result, int32_t *d_spill, int64_t spill_capacity,
                                 void *d_workspace, int64_t workspace_bytes, void *stream, int64_t *info)
{
    if (!e || !d_frac || !d_lattice || !d_result || !info) return fail(SAB_EINVAL, "sab_assign_device: null argument");
    if (lattice_stride != 0 && lattice_stride != 9) return fail(SAB_EINVAL, "lattice_stride must be 0 or 9");
    if (n_frames < 0) return fail(SAB_EINVAL, "negative frame count");
    int rc = check_ready(e);
    if (rc) return rc;
    for (int i = 0; i < 12; i++) info[i] = 0;
    info[3] = -1;
    if (n_frames == 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w = layout(e, n_frames, d_workspace);
    if (w.m3 && (!d_workspace || (int64_t)w.bytes > workspace_bytes))
        return fail(SAB_EINVAL, "workspace too small: need %zu bytes, got %lld", w.bytes, (long long)workspace_bytes);
    int64_t l0 = e->launches;
    e->peers_written = false;
    if (!e->h_pinned) {
        CU(cudaMallocHost((void **)&e->h_pinned, 2 * CNT_N * sizeof(unsigned long long)));
        for (int i = 0; i < 2 * CNT_N; i++) e->h_pinned[i] = 0ULL;
        e->h_pinned[CNT_FIRST_BAD] = ~0ULL;
    }
    CU(cudaMemcpyAsync(e->d_counters, e->h_pinned, CNT_N * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    for (int i = 0; i < 3; i++) if (!e->ev_t[i]) CU(cudaEventCreate(&e->ev_t[i]));
    CU(cudaEventRecord(e->ev_t[0], st));
    // tetrahedral sites: one fused kernel (wrap scan + assignment + commit); everything else: scan kernels, then assignment
    const StreamPlan5 sp5 = plan_stream5(e, d_frac, n_frames, w);
    StreamPlan sp;
    if (sp5.ok) sp.ok = true; else sp = plan_stream(e, d_frac, n_frames, w);
    if (!sp.ok && w.m3) { run_wrap_pass(e, d_frac, n_frames, w, st, true); CU(cudaGetLastError()); }
    CU(cudaEventRecord(e->ev_t[1], st));
    rc = sp5.ok ? launch_stream5(e, sp5, d_frac, n_frames, d_result, d_spill, spill_capacity, w, st)
       : sp.ok ? launch_stream(e, sp, d_frac, n_frames, d_result, d_spill, spill_capacity, w, st)
               : launch_assign(e, d_frac, d_lattice, lattice_stride, n_frames, d_result, d_spill, spill_capacity, w, st, nullptr);
    if (rc) return rc;
    CU(cudaEventRecord(e->ev_t[2], st));
    if (w.m3 && !sp.ok) {      // device-side conditional commit: skipped by the kernel itself on overflow / invalidating step
        k_wrap_commit<<<(w.m3 + 127) / 128, 128, 0, st>>>(d_frac, n_frames, n_frames, e->N, e->d_fw_atoms, w.m3, w.n_chunks,
                                                          w.rows, w.chunk_exc, e->d_prev_raw, e->d_wraps, e->d_excursion,
                                                          e->d_counters + CNT_OVERFLOW, e->d_counters + CNT_SPAN);
        e->launches++;
        CU(cudaGetLastError());
    }
    unsigned long long *h = e->h_pinned + CNT_N;
    CU(cudaMemcpyAsync(h, e->d_counters, CNT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (sp.ok && h[SAB_CNT_INVALID]) {       // anchor formulation of the wrap count failed its chain check: nothing was
        e->force_scan = true;                // committed; redo this batch with the scan kernels (identical results)
        rc = sab_assign_device(e, d_frac, d_lattice, lattice_stride, n_frames, d_result, d_spill, spill_capacity, d_workspace,
                               workspace_bytes, stream, info);
        e->force_scan = false;
        info[10] = 1;
        info[11] = 0;                        // the scan kernels do not store into peers: the caller gathers this batch itself
        return rc;
    }
    e->skip_step_check_once = false;
    info[0] = (int64_t)h[CNT_AMBIGUOUS]; info[1] = (int64_t)h[CNT_SPILL]; info[2] = (int64_t)h[CNT_OVERFLOW];
    info[3] = (h[CNT_FIRST_BAD] == ~0ULL) ? -1 : (int64_t)h[CNT_FIRST_BAD];
    info[4] = (int64_t)h[CNT_FULL27];
    info[8] = (int64_t)h[CNT_SPAN];
    info[9] = (int64_t)h[9];                 // atom-frames settled by the exact float64 pass of the tile kernel
    info[11] = e->peers_written ? 1 : 0;
    {
        float ms_wrap = 0.f, ms_main = 0.f;          // CUDA events on the launching stream
        cudaEventElapsedTime(&ms_wrap, e->ev_t[0], e->ev_t[1]);
        cudaEventElapsedTime(&ms_main, e->ev_t[1], e->ev_t[2]);
        info[6] = (int64_t)((double)ms_main * 1e6);
        info[7] = (int64_t)((double)ms_wrap * 1e6);
        if (const char *dbg = getenv("SAB_DEBUG")) if (dbg[0] == '1')
            fprintf(stderr, "[sab] frames=%lld full-search atom-frames=%llu (%.3f per frame) pair tests=%llu (%.1f per frame) fallback frames=%llu assign=%.3f ms wrap=%.3f ms\n",
                    (long long)n_frames, h[5], (double)h[5] / (double)n_frames, h[6], (double)h[6] / (double)n_frames, h[CNT_FULL27], ms_main, ms_wrap);
    }
    if (h[CNT_OVERFLOW]) {                   // nothing committed: the caller retries with a larger spill buffer
        info[5] = e->launches - l0;
        return fail(SAB_EOVERFLOW, "candidate spill buffer overflow (%llu elements needed, capacity %lld)",
                    h[CNT_SPILL], (long long)spill_capacity);
    }
    if (w.m3 && info[3] < 0 && e->legacy_init_frame >= 0) e->legacy_init_frame -= n_frames;      // now in the past
    info[5] = e->launches - l0;
    return SAB_OK;
}

extern "C" int sab_commit_pbc(sab_engine *e, const double *d_frac, int64_t n_frames, int64_t n_commit,
                              void *d_workspace, int64_t workspace_bytes, void *stream)
{
    if (!e || !d_frac || n_commit < 0 || n_commit > n_frames) return fail(SAB_EINVAL, "sab_commit_pbc: bad arguments");
    if (!e->pbc_set) return fail(SAB_ESTATE, "sab_commit_pbc: PBC state not set");
    if (n_commit == 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w = layout(e, n_frames, d_workspace);
    if ((int64_t)w.bytes > workspace_bytes) return fail(SAB_EINVAL, "workspace too small: need %zu bytes", w.bytes);
    k_wrap_commit<<<(w.m3 + 127) / 128, 128, 0, st>>>(d_frac, n_frames, n_commit, e->N, e->d_fw_atoms, w.m3, w.n_chunks,
                                                      w.rows, w.chunk_exc, e->d_prev_raw, e->d_wraps, e->d_excursion, nullptr, nullptr);
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    if (e->legacy_init_frame >= 0) e->legacy_init_frame -= n_commit;
    return SAB_OK;
}

extern "C" int sab_dynvor_centres(sab_engine *e, const double *d_frac, int64_t n_frames, double *d_centres,
                                  void *d_workspace, int64_t workspace_bytes, void *stream)
{
    if (!e || e->kind != SAB_DYNAMIC_VORONOI || !d_frac || !d_centres) return fail(SAB_EINVAL, "sab_dynvor_centres: bad arguments");
    int rc = check_ready(e);
    if (rc) return rc;
    if (n_frames <= 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w = layout(e, n_frames, d_workspace);
    if ((int64_t)w.bytes > workspace_bytes) return fail(SAB_EINVAL, "workspace too small: need %zu bytes", w.bytes);
    unsigned long long init[CNT_N] = {0, 0, 0, ~0ULL, 0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(e->d_counters, init, sizeof init, cudaMemcpyHostToDevice, st));
    run_wrap_pass(e, d_frac, n_frames, w, st, true);
    CU(cudaGetLastError());
    static const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    struct Tmp { double *p = nullptr; ~Tmp() { if (p) cudaFree(p); } } lat;      // released on every exit path
    CU(cudaMalloc((void **)&lat.p, sizeof ident));
    CU(cudaMemcpyAsync(lat.p, ident, sizeof ident, cudaMemcpyHostToDevice, st));
    rc = launch_assign(e, d_frac, lat.p, 0, n_frames, nullptr, nullptr, 0, w, st, d_centres);
    cudaStreamSynchronize(st);
    return rc;
}

// ---------------------------------------------------------------- host-buffer path
extern "C" int sab_assign_host(sab_engine *e, const double *h_frac, const double *h_lattice, int lattice_stride,
                               int64_t n_frames, int32_t *h_result, int64_t chunk_frames, int64_t *info)
{
    if (!e || !h_frac || !h_lattice || !h_result || !info) return fail(SAB_EINVAL, "sab_assign_host: null argument");
    if (lattice_stride != 0 && lattice_stride != 9) return fail(SAB_EINVAL, "lattice_stride must be 0 or 9");
    int rc = check_ready(e);
    if (rc) return rc;
    for (int i = 0; i < 12; i++) info[i] = 0;
    info[3] = -1;
    if (n_frames <= 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    if (chunk_frames <= 0) chunk_frames = 8192;
    chunk_frames = std::min(chunk_frames, n_frames);
    const size_t frame_b = sizeof(double) * 3 * (size_t)e->N, res_b = sizeof(int32_t) * (size_t)e->A;
    const size_t spill_elems = (size_t)chunk_frames * e->A * 2 + 1024;
    const size_t ws_b = layout(e, chunk_frames, nullptr).bytes;
    // per-buffer staging: frames | lattices | result | spill | workspace
    size_t o_lat = align256(frame_b * chunk_frames), o_res = o_lat + align256(sizeof(double) * 9 * chunk_frames);
    size_t o_spill = o_res + align256(res_b * chunk_frames), o_ws = o_spill + align256(sizeof(int32_t) * spill_elems);
    size_t need = o_ws + align256(ws_b);
    if (need > e->stage_bytes) {
        for (int i = 0; i < 2; i++) { if (e->stage[i]) cudaFree(e->stage[i]); e->stage[i] = nullptr; }
        for (int i = 0; i < 2; i++) CU(cudaMalloc(&e->stage[i], need));
        e->stage_bytes = need;
    }
    for (int i = 0; i < 2; i++) {
        if (!e->streams[i]) CU(cudaStreamCreateWithFlags(&e->streams[i], cudaStreamNonBlocking));
        if (!e->ev_free[i]) CU(cudaEventCreateWithFlags(&e->ev_free[i], cudaEventDisableTiming));
    }
    // The wrap scan carries state from chunk to chunk through the engine (stream-ordered):
    // chunk k+1's kernels must follow chunk k's commit, so compute is serialised on stream 0
    // while the H2D copy of the next chunk runs on stream 1 (and vice versa).
    struct Events {                                       // destroyed on every exit path (the CU() early returns included)
        cudaEvent_t copied[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
        ~Events() { for (int i = 0; i < 2; i++) { if (copied[i]) cudaEventDestroy(copied[i]); if (done[i]) cudaEventDestroy(done[i]); } }
    } evs;
    cudaEvent_t *ev_copied = evs.copied, *ev_done = evs.done;
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    }
    cudaStream_t s_copy = e->streams[1], s_comp = e->streams[0];
    int64_t tot_info[12] = {0, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0};
    int status = SAB_OK;
    int64_t n_chunks = (n_frames + chunk_frames - 1) / chunk_frames;
    // Results bound for PAGEABLE host memory (a fresh numpy array): cudaMemcpyAsync would stage them synchronously and stall
    // the pipeline, so they go to pinned staging first and a helper thread copies them out while the next chunks compute.
    bool out_pinned = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, h_result) == cudaSuccess) out_pinned = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
        cudaGetLastError();
    }
    if (!out_pinned && e->h_res_stage_bytes < res_b * (size_t)chunk_frames) {
        for (int i = 0; i < 2; i++) { if (e->h_res_stage[i]) cudaFreeHost(e->h_res_stage[i]); e->h_res_stage[i] = nullptr; }
        for (int i = 0; i < 2; i++) CU(cudaMallocHost(&e->h_res_stage[i], res_b * (size_t)chunk_frames));
        e->h_res_stage_bytes = res_b * (size_t)chunk_frames;
    }
    struct Job { int b; int64_t f0, nf; };
    std::mutex mu; std::condition_variable cv; std::deque<Job> jobs; bool quit = false; int busy[2] = {0, 0};
    cudaEvent_t ev_res[2] = {nullptr, nullptr};
    std::thread copier;
    if (!out_pinned) {
        for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&ev_res[i], cudaEventDisableTiming));
        const int dev = e->device;
        copier = std::thread([&, dev]() {
            cudaSetDevice(dev);
            for (;;) {
                Job j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return quit || !jobs.empty(); });
                    if (jobs.empty()) return;
                    j = jobs.front(); jobs.pop_front();
                }
                cudaEventSynchronize(ev_res[j.b]);
                memcpy(h_result + (size_t)j.f0 * e->A, e->h_res_stage[j.b], res_b * (size_t)j.nf);
                { std::lock_guard<std::mutex> lk(mu); busy[j.b] = 0; }
                cv.notify_all();
            }
        });
    }
    auto finish_copier = [&]() {
        if (!copier.joinable()) return;
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv.notify_all();
        copier.join();
        for (int i = 0; i < 2; i++) if (ev_res[i]) cudaEventDestroy(ev_res[i]);
    };
    auto issue_copy = [&](int64_t k) -> cudaError_t {
        int b = (int)(k & 1);
        int64_t f0 = k * chunk_frames, nf = std::min(chunk_frames, n_frames - f0);
        char *base = (char *)e->stage[b];
        cudaError_t er = cudaStreamWaitEvent(s_copy, ev_done[b], 0);     // buffer b free again
        if (er != cudaSuccess) return er;
        er = cudaMemcpyAsync(base, h_frac + (size_t)f0 * e->N * 3, frame_b * nf, cudaMemcpyHostToDevice, s_copy);
        if (er != cudaSuccess) return er;
        if (lattice_stride)
            er = cudaMemcpyAsync(base + o_lat, h_lattice + (size_t)f0 * 9, sizeof(double) * 9 * nf, cudaMemcpyHostToDevice, s_copy);
        else
            er = cudaMemcpyAsync(base + o_lat, h_lattice, sizeof(double) * 9, cudaMemcpyHostToDevice, s_copy);
        if (er != cudaSuccess) return er;
        return cudaEventRecord(ev_copied[b], s_copy);
    };
    cudaError_t cerr = issue_copy(0);
    for (int64_t k = 0; k < n_chunks && status == SAB_OK && cerr == cudaSuccess; k++) {
        int b = (int)(k & 1);
        int64_t f0 = k * chunk_frames, nf = std::min(chunk_frames, n_frames - f0);
        char *base = (char *)e->stage[b];
        if (k + 1 < n_chunks && (cerr = issue_copy(k + 1)) != cudaSuccess) break;
        if ((cerr = cudaStreamWaitEvent(s_comp, ev_copied[b], 0)) != cudaSuccess) break;
        int64_t ci[12];
        status = sab_assign_device(e, (const double *)base, (const double *)(base + o_lat), lattice_stride, nf,
                                   (int32_t *)(base + o_res), (int32_t *)(base + o_spill), (int64_t)spill_elems,
                                   base + o_ws, (int64_t)align256(ws_b), s_comp, ci);
        if (status != SAB_OK) break;
        if (ci[0] > 0) { status = fail(SAB_ESTATE, "sab_assign_host: %lld atom-frames are contained by several sites; "
                                       "use sab_assign_device + sab_resolve", (long long)ci[0]); break; }
        if (e->d_mirror && (cerr = cudaMemcpyAsync(e->d_mirror + (size_t)f0 * e->A, base + o_res, res_b * nf, cudaMemcpyDeviceToDevice, s_comp)) != cudaSuccess) break;
        if (out_pinned) {
            if ((cerr = cudaMemcpyAsync(h_result + (size_t)f0 * e->A, base + o_res, res_b * nf, cudaMemcpyDeviceToHost, s_comp)) != cudaSuccess) break;
        } else {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return busy[b] == 0; }); busy[b] = 1; }   // staging b copied out
            if ((cerr = cudaMemcpyAsync(e->h_res_stage[b], base + o_res, res_b * nf, cudaMemcpyDeviceToHost, s_comp)) != cudaSuccess) break;
            if ((cerr = cudaEventRecord(ev_res[b], s_comp)) != cudaSuccess) break;
            { std::lock_guard<std::mutex> lk(mu); jobs.push_back(Job{b, f0, nf}); }
            cv.notify_all();
        }
        if ((cerr = cudaEventRecord(ev_done[b], s_comp)) != cudaSuccess) break;
        tot_info[0] += ci[0]; tot_info[1] += ci[1]; tot_info[4] += ci[4]; tot_info[5] += ci[5];
        tot_info[6] += ci[6]; tot_info[7] += ci[7]; tot_info[9] += ci[9];
        tot_info[8] = std::max(tot_info[8], ci[8]);
        if (ci[3] >= 0 && tot_info[3] < 0) tot_info[3] = f0 + ci[3];
    }
    cudaStreamSynchronize(s_comp); cudaStreamSynchronize(s_copy);
    finish_copier();
    for (int i = 0; i < 12; i++) info[i] = tot_info[i];
    if (status == SAB_OK && cerr != cudaSuccess)
        return fail(SAB_ECUDA, "sab_assign_host: %s", cudaGetErrorString(cerr));
    return status;
}

extern "C" int sab_set_result_mirror(sab_engine *e, int32_t *d_full)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_result_mirror: null engine");
    e->d_mirror = d_full;
    return SAB_OK;
}

// ---------------------------------------------------------------- priority resolver
extern "C" int sab_set_priority_tables(sab_engine *e, const int32_t *rank, const double *lookup,
                                       const int32_t *neigh_off, const int32_t *neigh)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_priority_tables: null engine");
    CU(cudaSetDevice(e->device));
    if (rank) {
        CU(upload(&e->d_rank, rank, (size_t)e->S * (size_t)std::max(e->S - 1, 1)));
        std::vector<int32_t> pos((size_t)e->S * e->S, INT_MAX);
        for (int i = 0; i < e->S; i++)
            for (int j = 0; j < e->S - 1; j++) {
                int32_t t = rank[(size_t)i * (e->S - 1) + j];
                if (t < 0 || t >= e->S) return fail(SAB_EINVAL, "rank table entry out of range");
                pos[(size_t)i * e->S + t] = j;
            }
        CU(upload(&e->d_rank_pos, pos.data(), pos.size()));
    }
    if (lookup) CU(upload(&e->d_lookup, lookup, (size_t)3 * e->S));
    e->ranked = lookup ? 1 : 0;                 // a distance ranking exists; without `rank` its rows are supplied lazily
    if (!rank) {
        if (e->d_rank) { g_alloc_bytes.erase(e->d_rank); cudaFree(e->d_rank); e->d_rank = nullptr; }
        if (e->d_rank_pos) { g_alloc_bytes.erase(e->d_rank_pos); cudaFree(e->d_rank_pos); e->d_rank_pos = nullptr; }
    }
    if (lookup && !e->d_missing) {
        CU(cudaMalloc((void **)&e->d_missing, sizeof(int32_t) * SAB_MISSING_CAP));
        CU(cudaMalloc((void **)&e->d_n_missing, sizeof(int)));
        CU(cudaMemset(e->d_n_missing, 0, sizeof(int)));
        e->row_slot_h.assign(e->S, -1);
        CU(upload(&e->d_row_slot, e->row_slot_h.data(), e->row_slot_h.size()));
    }
    if (neigh_off) {
        CU(upload(&e->d_neigh_off, neigh_off, (size_t)e->S + 1));
        CU(upload(&e->d_neigh, neigh, (size_t)std::max(neigh_off[e->S], 1)));
    }
    return SAB_OK;
}

static SabPriorityTables priority_tables(sab_engine *e)
{
    SabPriorityTables P{};
    P.rank = e->d_rank; P.rank_pos = e->d_rank_pos; P.lookup = e->d_lookup;
    P.neigh_off = e->ranked ? nullptr : e->d_neigh_off; P.neigh = e->d_neigh;
    P.ranked = e->ranked;
    P.row_slot = e->d_rank_pos ? nullptr : e->d_row_slot; P.row_pos = e->d_row_pos; P.row_fwd = e->d_row_fwd;
    P.missing = e->d_missing; P.n_missing = e->d_n_missing; P.missing_cap = SAB_MISSING_CAP;
    return P;
}

extern "C" int sab_add_rank_rows(sab_engine *e, int n_rows, const int32_t *anchors, const int32_t *rows)
{
    if (!e || n_rows < 0 || (n_rows > 0 && (!anchors || !rows))) return fail(SAB_EINVAL, "sab_add_rank_rows: bad arguments");
    if (!e->ranked || !e->d_row_slot) return fail(SAB_ESTATE, "sab_add_rank_rows: no lazy ranking (sab_set_priority_tables with lookup, without rank)");
    if (n_rows == 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    const size_t S = (size_t)e->S;
    if (e->n_row_slots + n_rows > e->row_slot_cap) {             // grow the row cache
        const int cap = std::max(2 * e->row_slot_cap, e->n_row_slots + n_rows + 16);
        int32_t *np = nullptr, *nf = nullptr;
        CU(cudaMalloc((void **)&np, sizeof(int32_t) * S * cap));
        CU(cudaMalloc((void **)&nf, sizeof(int32_t) * S * cap));
        if (e->n_row_slots) {
            CU(cudaMemcpy(np, e->d_row_pos, sizeof(int32_t) * S * e->n_row_slots, cudaMemcpyDeviceToDevice));
            CU(cudaMemcpy(nf, e->d_row_fwd, sizeof(int32_t) * S * e->n_row_slots, cudaMemcpyDeviceToDevice));
        }
        if (e->d_row_pos) cudaFree(e->d_row_pos);
        if (e->d_row_fwd) cudaFree(e->d_row_fwd);
        e->d_row_pos = np; e->d_row_fwd = nf; e->row_slot_cap = cap;
    }
    std::vector<int32_t> pos(S), fwd(S, 0);
    for (int r = 0; r < n_rows; r++) {
        const int a = anchors[r];
        if (a < 0 || a >= e->S) return fail(SAB_EINVAL, "sab_add_rank_rows: anchor out of range");
        if (e->row_slot_h[a] >= 0) continue;
        std::fill(pos.begin(), pos.end(), INT_MAX);
        for (size_t j = 0; j + 1 < S; j++) {
            const int32_t t = rows[(size_t)r * (S - 1) + j];
            if (t < 0 || t >= e->S) return fail(SAB_EINVAL, "rank row entry out of range");
            pos[t] = (int32_t)j; fwd[j] = t;
        }
        const int slot = e->n_row_slots++;
        CU(cudaMemcpy(e->d_row_pos + S * slot, pos.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(e->d_row_fwd + S * slot, fwd.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice));   // stride S, S - 1 used
        e->row_slot_h[a] = slot;
    }
    CU(cudaMemcpy(e->d_row_slot, e->row_slot_h.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice));
    return SAB_OK;
}

extern "C" int sab_rank_missing(sab_engine *e, int32_t *anchors, int capacity)
{
    if (!e || capacity < 0 || (capacity > 0 && !anchors)) return -1;
    if (!e->d_n_missing) return 0;
    cudaSetDevice(e->device);
    int n = 0;
    if (cudaMemcpy(&n, e->d_n_missing, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    n = std::min(n, SAB_MISSING_CAP);
    std::vector<int32_t> h(std::max(n, 1));
    if (n && cudaMemcpy(h.data(), e->d_missing, sizeof(int32_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    std::sort(h.begin(), h.begin() + n);
    n = (int)(std::unique(h.begin(), h.begin() + n) - h.begin());
    for (int i = 0; i < std::min(n, capacity); i++) anchors[i] = h[i];
    cudaMemset(e->d_n_missing, 0, sizeof(int));
    return n;
}

extern "C" int sab_resolve(sab_engine *e, const double *d_frac, int64_t n_frames, int32_t *d_result,
                           const int32_t *d_spill, void *stream, int32_t *recent, int32_t *prev,
                           int32_t *tr_n, int32_t *tr_key, int64_t *tr_cnt, int tr_width)
{
    if (!e || !d_frac || !d_result || !recent || !prev || !tr_n || !tr_key || !tr_cnt || tr_width <= 0)
        return fail(SAB_EINVAL, "sab_resolve: null argument");
    if (n_frames <= 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    struct Tmp {                         // device copies of the history, released on every exit path
        void *p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Tmp() { for (void *q : p) if (q) cudaFree(q); }
    } tmp;
    SabHistory H{};
    H.K = tr_width;
    size_t nk = (size_t)e->S * tr_width;
    CU(cudaMalloc(&tmp.p[0], sizeof(int32_t) * 2 * e->A)); CU(cudaMalloc(&tmp.p[1], sizeof(int32_t) * e->A));
    CU(cudaMalloc(&tmp.p[2], sizeof(int32_t) * e->S)); CU(cudaMalloc(&tmp.p[3], sizeof(int32_t) * nk));
    CU(cudaMalloc(&tmp.p[4], sizeof(long long) * nk)); CU(cudaMalloc(&tmp.p[5], sizeof(int)));
    CU(cudaMalloc(&tmp.p[6], sizeof(SabSpecStats)));
    H.recent = (int32_t *)tmp.p[0]; H.prev = (int32_t *)tmp.p[1]; H.tr_n = (int32_t *)tmp.p[2];
    H.tr_key = (int32_t *)tmp.p[3]; H.tr_cnt = (long long *)tmp.p[4];
    int *d_over = (int *)tmp.p[5];
    SabSpecStats *d_stats = (SabSpecStats *)tmp.p[6];
    CU(cudaMemcpyAsync(H.recent, recent, sizeof(int32_t) * 2 * e->A, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.prev, prev, sizeof(int32_t) * e->A, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.tr_n, tr_n, sizeof(int32_t) * e->S, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.tr_key, tr_key, sizeof(int32_t) * nk, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.tr_cnt, tr_cnt, sizeof(long long) * nk, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_over, 0, sizeof(int), st));
    CU(cudaMemsetAsync(d_stats, 0, sizeof(SabSpecStats), st));
    SabPriorityTables P = priority_tables(e);
    int block = std::min(1024, ((e->A + 31) / 32) * 32);
    for (int i = 0; i < 3; i++) if (!e->ev_t[i]) CU(cudaEventCreate(&e->ev_t[i]));
    CU(cudaEventRecord(e->ev_t[0], st));
    // shared-memory resolvers when the Counter rows fit (K keys per site); else the global-memory one
    int max_n = 0;
    for (int s2 = 0; s2 < e->S; s2++) max_n = std::max(max_n, (int)tr_n[s2]);
    int smem_cap = 0;
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
    long long fixed = sizeof(int) * ((long long)e->A * (9 + 3 * 12) + 2LL * e->S) + 1024;
    // rows of tr_width keys must fit in shared memory untruncated; the host starts narrow and widens on
    // overflow, which eventually selects the global-memory kernel
    long long fit = ((long long)smem_cap - fixed) / (8LL * e->S);
    int k_fit = tr_width;
    const bool smem_ok = max_n <= tr_width && getenv("SAB_RESOLVE_GLOBAL") == nullptr && e->resolver != 2;
    // speculative chunks (k_resolve_spec.cuh): one atom per thread, decisions of T frames + logs next to the history
    int spec_T = 0, spec_NL = 0; size_t spec_smem = 0;
    if (smem_ok && e->A <= 1024 && e->resolver == 0 && getenv("SAB_RESOLVE_FAST") == nullptr) {
        int nl = 64; while (nl < e->A * SAB_SPEC_W) nl <<= 1;
        for (int T = 32; T >= 32 && !spec_T; T >>= 1) {      // one lane per frame of a chunk
            const size_t ints = (size_t)e->A * (3 + (size_t)T + 3 * SAB_SPEC_W + 2 * SAB_SPEC_R + 2 + 7) + 2 * (size_t)e->S +
                                2 * (size_t)e->S * tr_width + 3 * (size_t)nl;
            if (sizeof(int) * ints + 64 <= (size_t)smem_cap) { spec_T = T; spec_NL = nl; spec_smem = sizeof(int) * ints + 64; }
        }
    }
    int used = 0;                        // 1 speculative, 2 per-frame shared memory, 3 global memory
    if (spec_T) {                        // one warp per atom at a time: 32 warps share the atoms (and thread a < A commits atom a)
        CU(cudaFuncSetAttribute(k_resolve_spec<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec_smem));
        k_resolve_spec<1024><<<1, 1024, spec_smem, st>>>(d_frac, n_frames, e->N, e->A, e->d_mobile, e->S, d_result, d_spill, P, H, tr_width, spec_T, spec_NL, d_over, d_stats);
        used = 1;
    } else if (smem_ok && tr_width <= fit) {
        size_t smem = sizeof(int) * ((size_t)e->A * (9 + 3 * 12) + 2 * (size_t)e->S + 2 * (size_t)e->S * k_fit) + 16;
        if (block <= 256) {
            CU(cudaFuncSetAttribute(k_resolve_fast<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_resolve_fast<256><<<1, block, smem, st>>>(d_frac, n_frames, e->N, e->A, e->d_mobile, e->S, d_result, d_spill, P, H, k_fit, d_over);
        } else {
            CU(cudaFuncSetAttribute(k_resolve_fast<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_resolve_fast<1024><<<1, block, smem, st>>>(d_frac, n_frames, e->N, e->A, e->d_mobile, e->S, d_result, d_spill, P, H, k_fit, d_over);
        }
        used = 2;
    } else {
        size_t smem = sizeof(int) * (size_t)e->A;
        CU(cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        k_resolve<<<1, block, smem, st>>>(d_frac, n_frames, e->N, e->A, e->d_mobile, e->S, d_result, d_spill, P, H, d_over);
        used = 3;
    }
    e->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->ev_t[1], st));
    int over = 0;
    SabSpecStats hs{};
    CU(cudaMemcpyAsync(recent, H.recent, sizeof(int32_t) * 2 * e->A, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(prev, H.prev, sizeof(int32_t) * e->A, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tr_n, H.tr_n, sizeof(int32_t) * e->S, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tr_key, H.tr_key, sizeof(int32_t) * nk, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tr_cnt, H.tr_cnt, sizeof(long long) * nk, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&hs, d_stats, sizeof(SabSpecStats), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    {
        float ms = 0.f; cudaEventElapsedTime(&ms, e->ev_t[0], e->ev_t[1]);
        e->resolve_ns += (int64_t)((double)ms * 1e6);
        e->resolve_kind = used;
        if (const char *dbg = getenv("SAB_DEBUG")) if (dbg[0] == '1')
            fprintf(stderr, "[sab] resolve: kernel %d frames=%lld %.3f ms (%.3f us/frame) width=%d chunk=%d chunks=%llu truncated=%llu exact frames=%llu new keys=%llu Mcycles: load %.2f walk %.2f commit %.2f exact %.2f\n",
                    used, (long long)n_frames, ms, 1e3 * ms / (double)n_frames, tr_width, spec_T, hs.chunks, hs.truncated, hs.exact_frames, hs.new_keys,
                    hs.cyc_load * 1e-6, hs.cyc_walk * 1e-6, hs.cyc_commit * 1e-6, hs.cyc_exact * 1e-6);
    }
    if (over) return fail(SAB_EOVERFLOW, "transition table width %d exceeded", tr_width);
    return SAB_OK;
}

// ---------------------------------------------------------------- exact replay of the lazily updated per-site caches
extern "C" int sab_lazy_set_state(sab_engine *e, const int32_t *shift, const double *cached_raw, const uint8_t *has_shift,
                                  const uint8_t *has_ref_centre, const double *ref_centre)
{
    if (!e || !shift || !cached_raw || !has_shift || !has_ref_centre || !ref_centre) return fail(SAB_EINVAL, "sab_lazy_set_state: null argument");
    if (e->kind != SAB_POLYHEDRAL || !e->d_simp) return fail(SAB_ESTATE, "sab_lazy_set_state: polyhedra not set");
    CU(cudaSetDevice(e->device));
    const size_t n = (size_t)e->S * e->smax * 3;
    std::vector<int32_t> has(e->S);
    for (int s = 0; s < e->S; s++) has[s] = has_shift[s] ? 1 : 0;
    std::vector<long long> stamp(e->S, -1LL);
    std::vector<int32_t> zero(e->S, 0);
    std::vector<double> vc(n, 0.0);
    unsigned long long st0[4] = {0, 0, 0, 0};
    CU(upload(&e->d_lz_shift, shift, n)); CU(upload(&e->d_lz_cached, cached_raw, n)); CU(upload(&e->d_lz_has, has.data(), has.size()));
    CU(upload(&e->d_lz_stamp, stamp.data(), stamp.size())); CU(upload(&e->d_lz_lock, zero.data(), zero.size()));
    CU(upload(&e->d_lz_vc, vc.data(), n)); CU(upload(&e->d_lz_has_rc, has_ref_centre, (size_t)e->S));
    CU(upload(&e->d_lz_rc, ref_centre, (size_t)3 * e->S)); CU(upload(&e->d_lz_stats, st0, (size_t)4));
    e->lazy_set = true; e->lazy_frames = 0;
    return SAB_OK;
}

extern "C" int sab_lazy_get_state(sab_engine *e, int32_t *shift, double *cached_raw, uint8_t *has_shift, double *vertex_coords, int64_t *stats)
{
    if (!e || !e->lazy_set) return fail(SAB_ESTATE, "sab_lazy_get_state: lazy state not set");
    CU(cudaSetDevice(e->device));
    CU(cudaDeviceSynchronize());
    const size_t n = (size_t)e->S * e->smax * 3;
    if (shift) CU(cudaMemcpy(shift, e->d_lz_shift, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    if (cached_raw) CU(cudaMemcpy(cached_raw, e->d_lz_cached, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (vertex_coords) CU(cudaMemcpy(vertex_coords, e->d_lz_vc, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (has_shift) {
        std::vector<int32_t> has(e->S);
        CU(cudaMemcpy(has.data(), e->d_lz_has, sizeof(int32_t) * e->S, cudaMemcpyDeviceToHost));
        for (int s = 0; s < e->S; s++) has_shift[s] = (uint8_t)(has[s] != 0);
    }
    if (stats) {
        unsigned long long st[4];
        CU(cudaMemcpy(st, e->d_lz_stats, sizeof st, cudaMemcpyDeviceToHost));
        for (int i = 0; i < 4; i++) stats[i] = (int64_t)st[i];
    }
    return SAB_OK;
}

extern "C" int sab_lazy_set_open_topology(sab_engine *e, const uint8_t *open_site)
{
    if (!e || e->kind != SAB_POLYHEDRAL || !e->d_simp) return fail(SAB_ESTATE, "sab_lazy_set_open_topology: polyhedra not set");
    CU(cudaSetDevice(e->device));
    std::vector<int32_t> open(e->S, 0);
    if (open_site) for (int s = 0; s < e->S; s++) open[s] = open_site[s] ? 1 : 0;
    std::vector<long long> first(e->S, -1LL);
    std::vector<double> vc((size_t)e->S * e->smax * 3, 0.0);
    CU(upload(&e->d_lz_open, open.data(), open.size()));
    CU(upload(&e->d_lz_first_frame, first.data(), first.size()));
    CU(upload(&e->d_lz_first_vc, vc.data(), vc.size()));
    return SAB_OK;
}

extern "C" int sab_lazy_get_first_queries(sab_engine *e, int64_t *first_frame, double *first_vc)
{
    if (!e || !first_frame || !first_vc) return fail(SAB_EINVAL, "sab_lazy_get_first_queries: null argument");
    if (!e->d_lz_first_frame || !e->d_lz_first_vc) return fail(SAB_ESTATE, "sab_lazy_get_first_queries: call sab_lazy_set_open_topology first");
    CU(cudaSetDevice(e->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(first_frame, e->d_lz_first_frame, sizeof(long long) * (size_t)e->S, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(first_vc, e->d_lz_first_vc, sizeof(double) * (size_t)e->S * e->smax * 3, cudaMemcpyDeviceToHost));
    return SAB_OK;
}

extern "C" int sab_update_face_topology(sab_engine *e, int fmax, const int32_t *n_face, const int32_t *simplices)
{
    if (!e || !n_face || !simplices) return fail(SAB_EINVAL, "sab_update_face_topology: null argument");
    if (e->kind != SAB_POLYHEDRAL || !e->d_simp) return fail(SAB_ESTATE, "sab_update_face_topology: polyhedra not set");
    if (fmax <= 0) return fail(SAB_EINVAL, "fmax must be positive");
    CU(cudaSetDevice(e->device));
    std::vector<int32_t> nv(e->S);
    CU(cudaMemcpy(nv.data(), e->d_nslot, sizeof(int32_t) * (size_t)e->S, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> simp((size_t)e->S * fmax * 3, 0);
    for (int s = 0; s < e->S; s++) {
        if (n_face[s] < 4 || n_face[s] > fmax) return fail(SAB_EINVAL, "site %d: bad face count %d", s, n_face[s]);
        for (int j = 0; j < 3 * n_face[s]; j++) {
            const int v = simplices[(size_t)s * fmax * 3 + j];
            if (v < 0 || v >= nv[s]) return fail(SAB_EINVAL, "site %d: face vertex %d out of range", s, v);
            simp[(size_t)s * fmax * 3 + j] = (uint8_t)v;
        }
    }
    e->fmax = fmax;
    CU(upload(&e->d_nface, n_face, (size_t)e->S));
    CU(upload(&e->d_simp, simp.data(), simp.size()));
    return SAB_OK;
}

extern "C" int sab_assign_lazy(sab_engine *e, const double *d_frac, const double *d_lattice, int lattice_stride, int64_t n_frames,
                               int32_t *d_result, void *stream, int32_t *recent, int32_t *prev, int32_t *tr_n, int32_t *tr_key,
                               int64_t *tr_cnt, int tr_width, int append)
{
    if (!e || !d_frac || !d_lattice || !d_result || !recent || !prev || !tr_n || !tr_key || !tr_cnt || tr_width <= 0)
        return fail(SAB_EINVAL, "sab_assign_lazy: null argument");
    if (!e->lazy_set) return fail(SAB_ESTATE, "sab_assign_lazy: call sab_lazy_set_state first");
    if (lattice_stride != 0 && lattice_stride != 9) return fail(SAB_EINVAL, "lattice_stride must be 0 or 9");
    if (n_frames <= 0) return SAB_OK;
    CU(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    struct Tmp {                         // device copies of the history, released on every exit path
        void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Tmp() { for (void *q : p) if (q) cudaFree(q); }
    } tmp;
    SabHistory H{};
    H.K = tr_width;
    const size_t nk = (size_t)e->S * tr_width;
    CU(cudaMalloc(&tmp.p[0], sizeof(int32_t) * 2 * e->A)); CU(cudaMalloc(&tmp.p[1], sizeof(int32_t) * e->A));
    CU(cudaMalloc(&tmp.p[2], sizeof(int32_t) * e->S)); CU(cudaMalloc(&tmp.p[3], sizeof(int32_t) * nk));
    CU(cudaMalloc(&tmp.p[4], sizeof(long long) * nk)); CU(cudaMalloc(&tmp.p[5], sizeof(int)));
    H.recent = (int32_t *)tmp.p[0]; H.prev = (int32_t *)tmp.p[1]; H.tr_n = (int32_t *)tmp.p[2]; H.tr_key = (int32_t *)tmp.p[3];
    H.tr_cnt = (long long *)tmp.p[4];
    int *d_over = (int *)tmp.p[5];
    CU(cudaMemcpyAsync(H.recent, recent, sizeof(int32_t) * 2 * e->A, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.prev, prev, sizeof(int32_t) * e->A, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.tr_n, tr_n, sizeof(int32_t) * e->S, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.tr_key, tr_key, sizeof(int32_t) * nk, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(H.tr_cnt, tr_cnt, sizeof(long long) * nk, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_over, 0, sizeof(int), st));
    std::vector<int32_t> simp32;          // the kernel reads the uint8 simplices of sab_set_polyhedra
    SabLazy Z{e->smax, e->fmax, e->d_nslot, e->d_slot_atom, e->d_nface, e->d_simp, e->d_lz_has_rc, e->d_lz_rc, e->d_lz_shift,
              e->d_lz_cached, e->d_lz_has, e->d_lz_stamp, e->d_lz_lock, e->d_lz_vc, e->d_lz_stats,
              e->d_lz_open, e->d_lz_first_frame, e->d_lz_first_vc};
    SabPriorityTables P = priority_tables(e);
    const size_t smem = sizeof(double) * 24 * (size_t)e->A + sizeof(int) * (size_t)e->A + 16;
    CU(cudaFuncSetAttribute(k_polyhedral_lazy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    k_polyhedral_lazy<<<1, 256, smem, st>>>(d_frac, d_lattice, lattice_stride, n_frames, e->lazy_frames, e->N, e->A, e->d_mobile, e->S,
                                            Z, P, H, append ? 1 : 0, d_result, d_over);
    e->launches++;
    CU(cudaGetLastError());
    int over = 0;
    CU(cudaMemcpyAsync(recent, H.recent, sizeof(int32_t) * 2 * e->A, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(prev, H.prev, sizeof(int32_t) * e->A, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tr_n, H.tr_n, sizeof(int32_t) * e->S, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tr_key, H.tr_key, sizeof(int32_t) * nk, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tr_cnt, H.tr_cnt, sizeof(long long) * nk, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (over) return fail(SAB_EOVERFLOW, "transition table width %d exceeded", tr_width);     // state untouched by the caller: retry wider
    e->lazy_frames += n_frames;
    return SAB_OK;
}

// ---------------------------------------------------------------- folds over a finished result array
#include "k_fold.cuh"

struct FoldLayout { size_t keys_a, keys_b, cub_tmp, cub_bytes, t_key, t_cnt, t_first, flags, bytes; };

static int fold_end_bit(int64_t n_frames, int n_mobile, int n_sites)
{
    // keys are < n_sites * n_mobile * n_frames * 2
    long double top = (long double)n_sites * n_mobile * (long double)n_frames * 2.0L;
    int bits = 1;
    while (bits < 64 && (long double)(1ULL << bits) < top) bits++;
    return bits;
}

static int fold_layout(int64_t n_frames, int n_mobile, int n_sites, int64_t key_capacity, int64_t table_slots, FoldLayout *L)
{
    if (n_frames < 0 || n_mobile <= 0 || n_sites <= 0 || key_capacity < 0 || table_slots < 0)
        return fail(SAB_EINVAL, "sab_fold: negative size");
    if (table_slots & (table_slots - 1)) return fail(SAB_EINVAL, "sab_fold: table_slots must be a power of two (or 0)");
    if ((long double)n_sites * n_mobile * (long double)std::max<int64_t>(n_frames, 1) * 2.0L >= 9.0e18L)
        return fail(SAB_EINVAL, "sab_fold: n_sites * n_mobile * n_frames does not fit the 63-bit run keys");
    size_t cub_bytes = 0;
    if (key_capacity > 0) {
        unsigned long long *nul = nullptr;
        cudaError_t ce = cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, nul, nul, (long long)key_capacity, 0,
                                                        fold_end_bit(n_frames, n_mobile, n_sites));
        if (ce != cudaSuccess) return fail(SAB_ECUDA, "cub::DeviceRadixSort size query failed: %s", cudaGetErrorString(ce));
    }
    size_t off = 0;
    L->keys_a = off; off += align256(sizeof(unsigned long long) * (size_t)key_capacity);
    L->keys_b = off; off += align256(sizeof(unsigned long long) * (size_t)key_capacity);
    L->cub_tmp = off; L->cub_bytes = cub_bytes; off += align256(cub_bytes);
    L->t_key = off; off += align256(sizeof(unsigned long long) * (size_t)table_slots);
    L->t_cnt = off; off += align256(sizeof(unsigned long long) * (size_t)table_slots);
    L->t_first = off; off += align256(sizeof(unsigned long long) * (size_t)table_slots);
    L->flags = off; off += 256;
    L->bytes = off;
    return SAB_OK;
}

extern "C" int64_t sab_fold_workspace_bytes(int64_t n_frames, int n_mobile, int n_sites, int64_t run_capacity, int64_t table_slots)
{
    FoldLayout L;
    if (fold_layout(n_frames, n_mobile, n_sites, 2 * run_capacity, table_slots, &L)) return -1;
    return (int64_t)L.bytes;
}

extern "C" int sab_fold_result(int device, const int32_t *d_result, int64_t n_frames, int n_mobile, int n_sites,
                               const int32_t *d_prev, int64_t run_capacity, int32_t *d_run_site, int32_t *d_run_atom,
                               int64_t *d_run_start, int64_t *d_run_length, int64_t table_slots, int64_t pair_capacity,
                               int32_t *d_pair_src, int32_t *d_pair_dst, int64_t *d_pair_count, int64_t *d_pair_first,
                               int64_t *d_recent_last, int64_t *d_recent_other,
                               void *d_workspace, int64_t workspace_bytes, void *stream, int64_t *info)
{
    if (!info) return fail(SAB_EINVAL, "sab_fold_result: info is NULL");
    for (int i = 0; i < 6; i++) info[i] = 0;
    if (n_frames == 0) return SAB_OK;
    if (!d_result || !d_workspace) return fail(SAB_EINVAL, "sab_fold_result: null argument");
    if (run_capacity > 0 && (!d_run_site || !d_run_atom || !d_run_start || !d_run_length))
        return fail(SAB_EINVAL, "sab_fold_result: run arrays missing");
    if (table_slots > 0 && pair_capacity > 0 && (!d_pair_src || !d_pair_dst || !d_pair_count || !d_pair_first))
        return fail(SAB_EINVAL, "sab_fold_result: pair arrays missing");
    FoldLayout L;
    int rc = fold_layout(n_frames, n_mobile, n_sites, 2 * run_capacity, table_slots, &L);
    if (rc) return rc;
    if ((int64_t)L.bytes > workspace_bytes) return fail(SAB_EINVAL, "sab_fold_result: workspace too small: need %zu bytes", L.bytes);
    if ((d_recent_last == nullptr) != (d_recent_other == nullptr)) return fail(SAB_EINVAL, "sab_fold_result: pass both recent arrays or neither");
    if (d_recent_last && n_sites > (1 << 24)) return fail(SAB_EINVAL, "sab_fold_result: recent-site keys hold 24-bit site numbers");
    CU(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    char *b = (char *)d_workspace;
    unsigned long long *keys_a = (unsigned long long *)(b + L.keys_a), *keys_b = (unsigned long long *)(b + L.keys_b);
    unsigned long long *t_key = table_slots ? (unsigned long long *)(b + L.t_key) : nullptr;
    unsigned long long *t_cnt = (unsigned long long *)(b + L.t_cnt), *t_first = (unsigned long long *)(b + L.t_first);
    unsigned long long *flags = (unsigned long long *)(b + L.flags);
    cudaEvent_t ev0, ev1;
    CU(cudaEventCreate(&ev0)); CU(cudaEventCreate(&ev1));
    CU(cudaMemsetAsync(flags, 0, 256, st));
    if (table_slots) {
        CU(cudaMemsetAsync(t_key, 0xff, sizeof(unsigned long long) * (size_t)table_slots, st));
        CU(cudaMemsetAsync(t_cnt, 0, sizeof(unsigned long long) * (size_t)table_slots, st));
        CU(cudaMemsetAsync(t_first, 0xff, sizeof(unsigned long long) * (size_t)table_slots, st));
    }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    const long long total = (long long)n_frames * n_mobile;
    const long long tiles = (total + SAB_FOLD_TILE - 1) / SAB_FOLD_TILE;
    const int grid = (int)std::max<long long>(1, std::min<long long>(tiles, (long long)sm * 8));   // 8 resident CTAs of 256 per SM
    CU(cudaEventRecord(ev0, st));
    k_fold_scan<<<grid, 256, 0, st>>>(d_result, (long long)n_frames, n_mobile, n_sites, d_prev, keys_a, (long long)(2 * run_capacity),
                                      t_key, t_cnt, t_first, table_slots ? (unsigned long long)table_slots - 1 : 0ULL, flags);
    CU(cudaGetLastError());
    int launches = 1;
    unsigned long long h[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(h, flags, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const long long n_keys = (long long)h[0];
    info[0] = n_keys / 2; info[1] = (int64_t)h[1]; info[4] = (int64_t)h[2];
    auto finish = [&](int code, const char *msg) {
        cudaEventRecord(ev1, st); cudaEventSynchronize(ev1);
        float ms = 0.f; cudaEventElapsedTime(&ms, ev0, ev1);
        info[2] = launches; info[3] = (int64_t)(ms * 1e6);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
        return code == SAB_OK ? SAB_OK : fail(code, "%s", msg);
    };
    if (h[2] & SAB_FOLD_ERR_UNRESOLVED) return finish(SAB_ESTATE, "sab_fold_result: the result array holds unresolved (ambiguous) or out-of-range entries");
    if (h[2] & SAB_FOLD_ERR_TABLE) return finish(SAB_EOVERFLOW, "sab_fold_result: transition hash table full; pass more table_slots");
    if (n_keys & 1) return finish(SAB_ESTATE, "sab_fold_result: odd number of run boundaries");
    if (n_keys / 2 > run_capacity) return finish(SAB_EOVERFLOW, "sab_fold_result: run_capacity too small (info[0] holds the number of runs)");
    if (table_slots && (int64_t)h[1] > pair_capacity) return finish(SAB_EOVERFLOW, "sab_fold_result: pair_capacity too small (info[1] holds the number of pairs)");
    if (n_keys > 0) {
        size_t tmp = L.cub_bytes;
        cudaError_t ce = cub::DeviceRadixSort::SortKeys((void *)(b + L.cub_tmp), tmp, keys_a, keys_b, n_keys, 0,
                                                        fold_end_bit(n_frames, n_mobile, n_sites), st);
        if (ce != cudaSuccess) return finish(SAB_ECUDA, cudaGetErrorString(ce));
        launches += 4;                          // CUB's histogram + onesweep passes (library code, not counted as ours elsewhere)
        const long long n_runs = n_keys / 2;
        const int g2 = (int)std::max<long long>(1, std::min<long long>((n_runs + 255) / 256, (long long)sm * 8));
        k_fold_decode<<<g2, 256, 0, st>>>(keys_b, n_runs, (long long)n_frames, n_mobile, d_run_site, d_run_atom, d_run_start, d_run_length, flags);
        CU(cudaGetLastError());
        launches++;
        if (d_recent_last) {
            CU(cudaMemsetAsync(d_recent_last, 0, sizeof(int64_t) * (size_t)n_mobile, st));
            CU(cudaMemsetAsync(d_recent_other, 0, sizeof(int64_t) * (size_t)n_mobile, st));
            for (int pass = 0; pass < 2; pass++)
                k_fold_recent<<<g2, 256, 0, st>>>(d_run_site, d_run_atom, d_run_start, d_run_length, n_runs,
                                                  (unsigned long long *)d_recent_last, (unsigned long long *)d_recent_other, pass);
            CU(cudaGetLastError());
            launches += 2;
        }
    } else if (d_recent_last) {
        CU(cudaMemsetAsync(d_recent_last, 0, sizeof(int64_t) * (size_t)n_mobile, st));
        CU(cudaMemsetAsync(d_recent_other, 0, sizeof(int64_t) * (size_t)n_mobile, st));
    }
    if (table_slots && h[1] > 0) {
        const int g3 = (int)std::max<long long>(1, std::min<long long>((table_slots + 255) / 256, (long long)sm * 8));
        k_fold_compact<<<g3, 256, 0, st>>>(t_key, t_cnt, t_first, (unsigned long long)table_slots, n_sites, (long long)pair_capacity,
                                           d_pair_src, d_pair_dst, d_pair_count, d_pair_first, flags);
        CU(cudaGetLastError());
        launches++;
    }
    CU(cudaMemcpyAsync(h, flags, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h[2] & SAB_FOLD_ERR_PAIRING) return finish(SAB_ESTATE, "sab_fold_result: run boundaries do not pair up");
    return finish(SAB_OK, "");
}

// ---------------------------------------------------------------- minimum-image distance matrix (stand-alone operator)
#include "k_distances.cuh"

extern "C" int sab_mic_distances(int device, const double *d_c1, int n, const double *d_c2, int m,
                                 const double *d_lattice, double *d_out, void *stream)
{
    if (n < 0 || m < 0) return fail(SAB_EINVAL, "sab_mic_distances: negative size");
    if (n == 0 || m == 0) return SAB_OK;
    if (!d_c1 || !d_c2 || !d_lattice || !d_out) return fail(SAB_EINVAL, "sab_mic_distances: null argument");
    CU(cudaSetDevice(device));
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    const long long total = (long long)n * m;
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sm * 8));
    k_mic_distances<<<grid, 256, 0, (cudaStream_t)stream>>>(d_c1, n, d_c2, m, d_lattice, d_out);
    CU(cudaGetLastError());
    return SAB_OK;
}

extern "C" int sab_polyhedra_contain(int device, const double *d_vertex, const int32_t *d_n_vert, int vmax,
                                     const int32_t *d_simplices, const int32_t *d_n_face, int fmax, int n_sites,
                                     const double *d_points, int n_img, uint8_t *d_out, void *stream)
{
    if (n_sites < 0 || n_img < 0) return fail(SAB_EINVAL, "sab_polyhedra_contain: negative size");
    if (n_sites == 0) return SAB_OK;
    if (!d_vertex || !d_n_vert || !d_simplices || !d_n_face || !d_points || !d_out) return fail(SAB_EINVAL, "sab_polyhedra_contain: null argument");
    if (vmax < 1 || vmax > SAB_MAX_VERT || fmax < 1) return fail(SAB_EINVAL, "sab_polyhedra_contain: 1 <= vmax <= %d and fmax >= 1", SAB_MAX_VERT);
    CU(cudaSetDevice(device));
    k_polyhedra_contain<<<(n_sites + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_vertex, d_n_vert, vmax, d_simplices, d_n_face, fmax,
                                                                                n_sites, d_points, n_img, d_out);
    CU(cudaGetLastError());
    return SAB_OK;
}

// ---------------------------------------------------------------- measured FP64 issue peak without FMA (roofline denominator)
// mode 0: dependent-free DADD streams, 1: DMUL streams, 2: DMUL + DADD pairs (what -fmad=false code issues), 3: DFMA (context)
template <int MODE> __global__ void k_fp64_peak(double *out, int iters, double seed)
{
    double a[8], b = seed + 1e-9 * threadIdx.x, c = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __dadd_rn(a[i], b);
            else if (MODE == 1) a[i] = __dmul_rn(a[i], c);
            else if (MODE == 2) a[i] = __dadd_rn(__dmul_rn(a[i], c), b);
            else a[i] = __fma_rn(a[i], c, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;            // never true: keeps the arithmetic alive
}

extern "C" int sab_measure_fp64_peak(int device, double *tflops /* [4]: DADD, DMUL, DMUL+DADD, DFMA (2 flop) */)
{
    if (!tflops) return fail(SAB_EINVAL, "sab_measure_fp64_peak: null argument");
    CU(cudaSetDevice(device));
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    double *d = nullptr;
    CU(cudaMalloc((void **)&d, 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int iters = 20000, grid = sm * 8, block = 256;
    for (int mode = 0; mode < 4; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k_fp64_peak<0><<<grid, block>>>(d, iters, 1.0);
            else if (mode == 1) k_fp64_peak<1><<<grid, block>>>(d, iters, 1.0);
            else if (mode == 2) k_fp64_peak<2><<<grid, block>>>(d, iters, 1.0);
            else k_fp64_peak<3><<<grid, block>>>(d, iters, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double ops = (double)grid * block * 8.0 * iters * (mode >= 2 ? 2.0 : 1.0);
        tflops[mode] = ops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    CU(cudaGetLastError());
    return SAB_OK;
}

// ---------------------------------------------------------------- multi-GPU: result rows stored straight into peer memory
extern "C" int sab_set_result_peers(sab_engine *e, int n_peers, int32_t *const *d_peer_slots)
{
    if (!e || n_peers < 0 || n_peers > SAB_MAX_PEERS || (n_peers > 0 && !d_peer_slots))
        return fail(SAB_EINVAL, "sab_set_result_peers: between 0 and %d peers", SAB_MAX_PEERS);
    e->n_peer = n_peers;
    e->peer_mc = 0;
    for (int q = 0; q < SAB_MAX_PEERS; q++) e->peer[q] = q < n_peers ? d_peer_slots[q] : nullptr;
    return SAB_OK;
}

extern "C" int sab_set_result_multicast(sab_engine *e, int32_t *d_multicast_slot)
{
    if (!e) return fail(SAB_EINVAL, "sab_set_result_multicast: null engine");
    e->n_peer = d_multicast_slot ? 1 : 0;
    e->peer_mc = d_multicast_slot ? 1 : 0;
    for (int q = 0; q < SAB_MAX_PEERS; q++) e->peer[q] = nullptr;
    e->peer[0] = d_multicast_slot;
    return SAB_OK;
}

// ---------------------------------------------------------------- XDATCAR text -> arrays (host code)
#include "xdatcar_host.h"

extern "C" int sab_xdatcar_index(const char *text, int64_t len, int64_t body_offset, int n_atoms, int64_t *frame_off,
                                 int64_t *header_off, int64_t capacity, int64_t *n_frames)
{
    if (!text || len < 0 || body_offset < 0 || body_offset > len || n_atoms <= 0 || !n_frames || capacity < 0 ||
        (capacity > 0 && (!frame_off || !header_off)))
        return fail(SAB_EINVAL, "sab_xdatcar_index: bad arguments");
    const int64_t n = sabx_index(text, len, body_offset, n_atoms, frame_off, header_off, capacity);
    if (n < 0) return fail(SAB_EINVAL, "XDATCAR: truncated coordinate block or malformed repeated header");
    *n_frames = n;
    return n > capacity ? SAB_EOVERFLOW : SAB_OK;
}

extern "C" int sab_xdatcar_parse(const char *text, int64_t len, int n_atoms, int64_t n_frames, const int64_t *frame_off,
                                 const int64_t *header_off, double *frac, double *lattice, int n_threads)
{
    if (!text || len < 0 || n_atoms <= 0 || n_frames < 0 || (n_frames > 0 && (!frame_off || !frac)) || (lattice && !header_off))
        return fail(SAB_EINVAL, "sab_xdatcar_parse: bad arguments");
    for (int64_t k = 0; k < n_frames; k++)
        if (frame_off[k] < 0 || frame_off[k] > len || (lattice && header_off[k] > len))
            return fail(SAB_EINVAL, "sab_xdatcar_parse: frame offset outside the text");
    const int64_t bad = sabx_parse(text, len, n_atoms, n_frames, frame_off, header_off, frac, lattice, n_threads);
    if (bad >= 0) return fail(SAB_EINVAL, "XDATCAR: frame %lld holds a line that is not three numbers", (long long)bad);
    return SAB_OK;
}

extern "C" int sab_parse_double(const char *text, int64_t len, double *out, int64_t *consumed)
{
    if (!text || !out || len < 0) return fail(SAB_EINVAL, "sab_parse_double: bad arguments");
    const char *p = text;
    if (!sabx_parse_double(p, text + len, *out)) return fail(SAB_EINVAL, "not a number");
    if (consumed) *consumed = (int64_t)(p - text);
    return SAB_OK;
}
