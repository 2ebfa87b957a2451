// xdatcar_host.h -- host-side ingestion of VASP XDATCAR text into the arrays the kernels consume
// (the reference only accepts Sequence[Structure], trajectory.py:413-432, which users
// obtain from pymatgen's Xdatcar parser -- one Python float() per coordinate and one Structure object per frame).
//
//   sab_xdatcar_index   one sequential pass over the text (memchr speed): where every frame's coordinate block starts
//                       and, for variable-cell files, where the header repeated in front of it starts.
//   sab_xdatcar_parse   frames in parallel over host threads: text -> double[F][N][3] (+ double[F][3][3] lattices).
//
// Numbers are converted EXACTLY as Python's float() / C strtod would (correctly rounded): up to 15 significant digits
// and a decimal exponent within +-22 take Clinger's fast path (both the digit string as an integer and the power of ten
// are exact doubles, and IEEE multiplication / division round correctly); everything else goes to strtod.
// Host code only; compiled into libsab200.so by the host compiler behind nvcc.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static inline bool sabx_is_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

static const double SABX_POW10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Parses one number starting at p (after blanks); on success advances p past it.  Never reads at or beyond end.
static inline bool sabx_parse_double(const char *&p, const char *end, double &out)
{
    while (p < end && sabx_is_blank(*p)) p++;
    const char *s = p;
    if (s >= end) return false;
    bool neg = false;
    if (*s == '-' || *s == '+') { neg = *s == '-'; s++; }
    uint64_t m = 0;
    int digits = 0, sig = 0, exp10 = 0;
    bool seen_point = false, fast = true;
    const char *q = s;
    for (; q < end; q++) {
        const char c = *q;
        if (c >= '0' && c <= '9') {
            digits++;
            if (sig > 0 || c != '0') {
                if (sig < 19) { m = m * 10 + (uint64_t)(c - '0'); sig++; if (seen_point) exp10--; }
                else { fast = false; if (!seen_point) exp10++; }      // more digits than the mantissa holds
            } else if (seen_point) exp10--;                            // leading zeros after the point
        } else if (c == '.' && !seen_point) seen_point = true;
        else break;
    }
    if (digits == 0) return false;
    if (q < end && (*q == 'e' || *q == 'E' || *q == 'd' || *q == 'D')) {  // 'd': Fortran double-precision exponent letter
        const char *r = q + 1;
        bool eneg = false;
        if (r < end && (*r == '-' || *r == '+')) { eneg = *r == '-'; r++; }
        int e = 0, ed = 0;
        for (; r < end && *r >= '0' && *r <= '9'; r++) { if (e < 100000) e = e * 10 + (*r - '0'); ed++; }
        if (ed) { exp10 += eneg ? -e : e; q = r; }                          // "1e" without digits: the number ends before the letter
    }
    if (q < end && !(sabx_is_blank(*q) || *q == '\n')) return false;    // trailing garbage glued to the number
    if (fast && sig <= 15 && exp10 >= -22 && exp10 <= 22) {
        double v = (double)m;                                          // exact: m < 10^15 < 2^53
        v = exp10 < 0 ? v / SABX_POW10[-exp10] : v * SABX_POW10[exp10];
        out = neg ? -v : v;
    } else {
        char buf[128];
        size_t n = (size_t)(q - p);
        if (n >= sizeof buf) return false;
        memcpy(buf, p, n); buf[n] = 0;
        for (size_t i = 0; i < n; i++) if (buf[i] == 'd' || buf[i] == 'D') buf[i] = 'e';   // Fortran exponent letter
        char *stop = nullptr;
        out = strtod(buf, &stop);
        if (stop != buf + n) return false;
    }
    p = q;
    return true;
}

static inline const char *sabx_next_line(const char *p, const char *end)
{
    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl + 1 : end;
}

static inline bool sabx_starts_with_direct(const char *p, const char *end)
{
    while (p < end && sabx_is_blank(*p)) p++;
    return end - p >= 6 && (p[0] == 'D' || p[0] == 'd') && p[1] == 'i' && p[2] == 'r' && p[3] == 'e' && p[4] == 'c' && p[5] == 't';
}

static inline bool sabx_line_is_empty(const char *p, const char *end)
{
    while (p < end && sabx_is_blank(*p)) p++;
    return p >= end || *p == '\n';
}

// Walks the text after the first header.  frame_off[k]: offset of the first coordinate line of frame k;
// header_off[k]: offset of the 7-line header repeated in front of frame k's "Direct" line, or -1.
// Returns the number of frames found (may exceed capacity: call again with room), or -1 on a malformed file.
static int64_t sabx_index(const char *text, int64_t len, int64_t body_off, int n_atoms, int64_t *frame_off,
                          int64_t *header_off, int64_t capacity)
{
    const char *end = text + len, *p = text + body_off;
    int64_t n = 0, pending_header = -1;
    while (p < end) {
        if (sabx_line_is_empty(p, end)) { p = sabx_next_line(p, end); continue; }
        if (sabx_starts_with_direct(p, end)) {
            p = sabx_next_line(p, end);
            if (n < capacity) { frame_off[n] = (int64_t)(p - text); header_off[n] = pending_header; }
            n++;
            pending_header = -1;
            for (int a = 0; a < n_atoms; a++) {
                if (p >= end) return -1;                           // truncated coordinate block
                p = sabx_next_line(p, end);
            }
        } else {                                                   // a repeated header: comment, scale, 3 vectors, symbols, counts
            pending_header = (int64_t)(p - text);
            for (int l = 0; l < 7; l++) {
                if (p >= end) return -1;
                p = sabx_next_line(p, end);
            }
            if (!sabx_starts_with_direct(p, end)) return -1;
        }
    }
    return n;
}

// scale line + 3 lattice vectors starting at `p` (the line AFTER the comment line) -> L[9] = scale * vectors
// (a negative scale is the cell volume, as in POSCAR).
static bool sabx_parse_lattice(const char *p, const char *end, double *L)
{
    double scale;
    if (!sabx_parse_double(p, end, scale)) return false;
    p = sabx_next_line(p, end);
    for (int i = 0; i < 3; i++) {
        for (int d = 0; d < 3; d++) if (!sabx_parse_double(p, end, L[3 * i + d])) return false;
        p = sabx_next_line(p, end);
    }
    if (scale < 0.0) {
        const double vol = L[0] * (L[4] * L[8] - L[5] * L[7]) - L[1] * (L[3] * L[8] - L[5] * L[6]) + L[2] * (L[3] * L[7] - L[4] * L[6]);
        scale = cbrt(-scale / fabs(vol));
    }
    for (int i = 0; i < 9; i++) L[i] *= scale;
    return true;
}

// Returns -1 on success, else the index of the first frame that failed to parse.
static int64_t sabx_parse(const char *text, int64_t len, int n_atoms, int64_t n_frames, const int64_t *frame_off,
                          const int64_t *header_off, double *frac, double *lattice, int n_threads)
{
    const char *end = text + len;
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_frames) n_threads = (int)(n_frames > 0 ? n_frames : 1);
    std::vector<int64_t> bad((size_t)n_threads, -1);
    auto work = [&](int t) {
        const int64_t k0 = n_frames * t / n_threads, k1 = n_frames * (t + 1) / n_threads;
        for (int64_t k = k0; k < k1; k++) {
            const char *p = text + frame_off[k];
            double *out = frac + (size_t)k * n_atoms * 3;
            for (int a = 0; a < n_atoms; a++) {
                for (int d = 0; d < 3; d++)
                    if (!sabx_parse_double(p, end, out[3 * a + d])) { bad[(size_t)t] = k; return; }
                p = sabx_next_line(p, end);
            }
            if (lattice && header_off[k] >= 0) {
                const char *h = sabx_next_line(text + header_off[k], end);      // skip the comment line
                if (!sabx_parse_lattice(h, end, lattice + (size_t)k * 9)) { bad[(size_t)t] = k; return; }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (int t = 0; t < n_threads; t++) if (bad[(size_t)t] >= 0) return bad[(size_t)t];
    return -1;
}
