// mtx_io.cu -- Matrix-Market front end (host only).
//
// Same observable behaviour as the loader inlined in the reference's main (CPU/main.cpp:143-458,
// GPU/main.cu:78-234; banner rules of CPU/mmio.h:254-337, size line mmio.h:339-367):
//   coordinate files; `complex` rejected; real via %lg, integer via %d, pattern -> 1.0;
//   1-based -> 0-based; symmetric / hermitian entries mirrored for i != j (skew-symmetric is not);
//   rows filled in file order (stable counting sort by row): columns are NOT sorted and duplicate
//   (i,j) are NOT merged.  Return codes follow main(): -1 open, -2 banner, -3 complex, -4 size line;
//   -5 (ours) when the matrix does not fit the int32 layout or host memory.  Entries whose row or
//   column lies outside the declared shape are skipped (the reference writes out of bounds there).
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <limits.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

// The entries are not read with fscanf (the reference's loop, ~80 MB/s: 8.6 s for a 20 M-entry file) but parsed from
// the file in memory.  fscanf("%d %d %lg") is token based, not line based, and stops at the first token that does not
// convert; `parse_tokens` restates exactly that, sequentially.  A regular file -- every non-blank line holds exactly one
// entry and nothing else, which is what every writer produces -- gives the same entries whichever way it is cut at line
// ends, so `parse_lines` reads it with several threads; the first irregular line anywhere sends the whole file through
// the sequential tokenizer instead.  Numbers convert as scanf converts them: integers like strtol stored to an int,
// reals correctly rounded (std::from_chars; strtod for the spellings from_chars does not take: a leading '+', hex
// floats, inf/nan, out-of-range magnitudes).  IAS_MTX_LOADER=fscanf keeps the reference's loop for A/B tests.
namespace {

std::string lower(const char *s)
{
    std::string r(s);
    for (char &ch : r) ch = (char)tolower((unsigned char)ch);
    return r;
}

struct Entry { int i, j; double v; };
enum Field { F_REAL, F_INTEGER, F_PATTERN };

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// %d: optional sign, decimal digits; the value as scanf stores it (strtol, then narrowed to int)
inline bool parse_int(const char *&p, const char *end, int &out)
{
    const char *q = p;
    bool neg = false;
    if (q < end && (*q == '+' || *q == '-')) { neg = *q == '-'; ++q; }
    if (q >= end || *q < '0' || *q > '9') return false;
    unsigned long long acc = 0;
    bool over = false;
    while (q < end && *q >= '0' && *q <= '9') {
        if (acc > (ULLONG_MAX - 9) / 10) over = true; else acc = acc * 10 + (unsigned)(*q - '0');
        ++q;
    }
    long val;
    if (over || acc > (neg ? (unsigned long long)LONG_MAX + 1ULL : (unsigned long long)LONG_MAX)) val = neg ? LONG_MIN : LONG_MAX;
    else val = neg ? (long)(0ULL - acc) : (long)acc;
    out = (int)val;
    p = q;
    return true;
}

// %lg: whatever strtod takes; the buffer is NUL-terminated, so strtod cannot run past it
inline bool parse_real(const char *&p, const char *end, double &out)
{
    const char *q = p;
    if (q < end && *q == '-') ++q;
    if (q < end && ((*q >= '0' && *q <= '9') || *q == '.') && !(q + 1 < end && *q == '0' && (q[1] == 'x' || q[1] == 'X'))) {
        double v;
        std::from_chars_result r = std::from_chars(p, end, v, std::chars_format::general);
        if (r.ec == std::errc() && r.ptr != p) { out = v; p = r.ptr; return true; }
    }
    char *stop = nullptr;
    double v = strtod(p, &stop);
    if (stop == p) return false;
    out = v;
    p = stop;
    return true;
}

// one entry, tokens separated by any white space (also line ends): false = a token did not convert
inline bool parse_entry(const char *&p, const char *end, Field field, Entry &x)
{
    while (p < end && is_space(*p)) ++p;
    if (!parse_int(p, end, x.i)) return false;
    while (p < end && is_space(*p)) ++p;
    if (!parse_int(p, end, x.j)) return false;
    x.v = 1.0;
    if (field == F_PATTERN) return true;
    while (p < end && is_space(*p)) ++p;
    if (field == F_REAL) return parse_real(p, end, x.v);
    int iv = 0;
    if (!parse_int(p, end, iv)) return false;
    x.v = iv;
    return true;
}

// the reference's loop: up to nz entries, stop at the first failed conversion
void parse_tokens(const char *buf, size_t len, Field field, long long nz, std::vector<Entry> &e)
{
    const char *p = buf, *end = buf + len;
    for (long long t = 0; t < nz; ++t) {
        Entry x;
        if (!parse_entry(p, end, field, x)) break;
        e.push_back(x);
    }
}

// lines [begin, end) of a regular file; false at the first line that is not "one entry and nothing else"
bool parse_line_range(const char *begin, const char *end, Field field, std::vector<Entry> &e, const std::atomic<bool> &give_up)
{
    const char *p = begin;
    size_t since_check = 0;
    while (p < end) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = eol ? eol : end;
        const char *q = p;
        while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        if (q < le) {
            Entry x;
            const char *r = q;
            if (!parse_entry(r, le, field, x)) return false;
            while (r < le && (*r == ' ' || *r == '\t' || *r == '\r')) ++r;
            if (r != le) return false;                          // something else on the line (or \v, \f, a NUL): not regular
            e.push_back(x);
        }
        p = eol ? eol + 1 : end;
        if (++since_check == 4096) { since_check = 0; if (give_up.load(std::memory_order_relaxed)) return false; }
    }
    return true;
}

// several threads over a regular file; parts[k] = the entries of the k-th piece, in file order.  false: irregular.
bool parse_lines(const char *buf, size_t len, Field field, std::vector<std::vector<Entry>> &parts)
{
    unsigned hw = std::thread::hardware_concurrency();
    size_t threads = hw ? hw : 4;
    if (threads > 32) threads = 32;
    const size_t by_size = len / ((size_t)2 << 20) + 1;          // at least 2 MB of text per thread
    if (threads > by_size) threads = by_size;
    std::vector<const char *> cut(threads + 1);
    cut[0] = buf;
    cut[threads] = buf + len;
    for (size_t k = 1; k < threads; ++k) {
        const char *at = buf + len / threads * k;
        if (at < cut[k - 1]) at = cut[k - 1];
        const char *nl = (const char *)memchr(at, '\n', (size_t)(buf + len - at));
        cut[k] = nl ? nl + 1 : buf + len;
    }
    parts.assign(threads, std::vector<Entry>());
    std::atomic<bool> give_up(false);
    const size_t guess = len / threads / 16 + 16;
    auto work = [&](size_t k) {
        try {
            parts[k].reserve(guess);
            if (!parse_line_range(cut[k], cut[k + 1], field, parts[k], give_up)) give_up.store(true);
        } catch (...) {
            give_up.store(true);                                 // out of memory in one piece: the sequential path reports it
        }
    };
    std::vector<std::thread> pool;
    for (size_t k = 1; k < threads; ++k) {
        try { pool.emplace_back(work, k); } catch (...) { give_up.store(true); break; }
    }
    work(0);
    for (std::thread &t : pool) t.join();
    return !give_up.load();
}

}  // namespace

static int mtx_load_impl(const char *path, IasCsrMatrix *out);

extern "C" {

int ias_mtx_load(const char *path, IasCsrMatrix *out)
{
    try {
        return mtx_load_impl(path, out);
    } catch (const std::bad_alloc &) {        // nothing may unwind through the C boundary
        if (out) memset(out, 0, sizeof *out);
        return -5;
    }
}

}  // extern "C"

static int mtx_load_impl(const char *path, IasCsrMatrix *out)
{
    if (!path || !out) return -1;
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[1100];
    char w0[64] = "", w1[64] = "", w2[64] = "", w3[64] = "", w4[64] = "";
    if (!fgets(line, sizeof line, f) || sscanf(line, "%63s %63s %63s %63s %63s", w0, w1, w2, w3, w4) != 5) { fclose(f); return -2; }
    std::string object = lower(w1), format = lower(w2), field = lower(w3), symmetry = lower(w4);
    bool banner_ok = strcmp(w0, "%%MatrixMarket") == 0 && object == "matrix" && (format == "coordinate" || format == "array") &&
                     (field == "real" || field == "integer" || field == "pattern" || field == "complex") &&
                     (symmetry == "general" || symmetry == "symmetric" || symmetry == "hermitian" || symmetry == "skew-symmetric");
    if (!banner_ok) { fclose(f); return -2; }
    if (field == "complex") { fclose(f); return -3; }
    const bool mirror = symmetry == "symmetric" || symmetry == "hermitian";
    const Field fk = field == "real" ? F_REAL : field == "integer" ? F_INTEGER : F_PATTERN;

    int m = 0, n = 0, nz = 0;
    bool have_size = false;
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '%') continue;
        if (sscanf(line, "%d %d %d", &m, &n, &nz) == 3) { have_size = true; break; }
    }
    if (!have_size || m < 0 || n < 0 || nz < 0) { fclose(f); return -4; }

    // raw entries in file order: parts[k] one after the other, at most nz of them
    std::vector<std::vector<Entry>> parts;
    const bool trace = ias::HostTrace::enabled();              // IAS_HOST_TRACE=1: stage times on stderr
    auto clock_now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(clock_now() - t).count(); };
    auto t_stage = clock_now();
    double ms_read = 0.0, ms_parse = 0.0;
    size_t pieces = 1;
    const char *mode = getenv("IAS_MTX_LOADER");
    if (mode && !strcmp(mode, "fscanf")) {
        parts.assign(1, std::vector<Entry>());
        std::vector<Entry> &e = parts[0];
        e.reserve((size_t)(nz < (1 << 24) ? nz : (1 << 24)));      // the size line is input: do not trust it with memory
        for (int t = 0; t < nz; ++t) {
            Entry x{0, 0, 1.0};
            int iv = 0;
            int got = fk == F_REAL ? fscanf(f, "%d %d %lg", &x.i, &x.j, &x.v)
                    : fk == F_INTEGER ? fscanf(f, "%d %d %d", &x.i, &x.j, &iv)
                                      : fscanf(f, "%d %d", &x.i, &x.j);
            if (got != (fk == F_PATTERN ? 2 : 3)) break;
            if (fk == F_INTEGER) x.v = iv;
            e.push_back(x);
        }
        fclose(f);
    } else {
        // the rest of the file in memory (a stream that cannot seek is read in pieces)
        std::unique_ptr<char[]> text;                            // not value-initialised: a GB of zeros would be written first
        std::vector<char> grown;
        size_t len = 0;
        long here = ftell(f);
        long size = -1;
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0) { size = ftell(f); fseek(f, here, SEEK_SET); }
        if (size >= here && here >= 0) {
            text.reset(new char[(size_t)(size - here) + 1]);
            len = fread(text.get(), 1, (size_t)(size - here), f);
        } else {
            char chunk[1 << 16];
            size_t got;
            while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) grown.insert(grown.end(), chunk, chunk + got);
            len = grown.size();
            grown.push_back(0);
        }
        fclose(f);
        char *buf = text ? text.get() : grown.data();
        buf[len] = 0;                                           // strtod must find an end
        ms_read = ms_since(t_stage);
        t_stage = clock_now();
        if (!parse_lines(buf, len, fk, parts)) {
            parts.assign(1, std::vector<Entry>());
            parts[0].reserve((size_t)(nz < (1 << 24) ? nz : (1 << 24)));
            parse_tokens(buf, len, fk, nz, parts[0]);
        }
        pieces = parts.size();
        ms_parse = ms_since(t_stage);
        t_stage = clock_now();
    }

    // 1-based -> 0-based; entries outside the declared shape are skipped (the reference would write / index out of
    // bounds there); only the first nz entries of the file count.  Rows are filled in file order (a stable counting
    // sort by row).  Each thread owns a range of rows and walks ALL entries in file order, counting / placing those
    // that land in its rows: sequential reads for everyone, scattered writes only inside the thread's own rows, and
    // the order inside a row is the file's whatever the number of threads.
    size_t raw_total = 0;
    for (const std::vector<Entry> &part : parts) raw_total += part.size();
    unsigned hw = std::thread::hardware_concurrency();
    size_t workers = hw ? hw : 4;
    if (workers > 16) workers = 16;
    if (workers > raw_total / ((size_t)1 << 18) + 1) workers = raw_total / ((size_t)1 << 18) + 1;   // at least 256 Ki entries per thread
    if (workers > (size_t)m) workers = m > 0 ? (size_t)m : 1;
    auto scan_rows = [&](int lo, int hi, auto &&fn) {            // fn(row, col, value) for every CSR entry of rows [lo, hi)
        long long seen = 0;
        for (const std::vector<Entry> &part : parts)
            for (const Entry &raw : part) {
                if (seen++ >= (long long)nz) return;
                const int i = raw.i - 1, j = raw.j - 1;
                if (i < 0 || i >= m || j < 0 || j >= n) continue;
                if (i >= lo && i < hi) fn(i, j, raw.v);
                if (mirror && i != j && j < m && i < n && j >= lo && j < hi) fn(j, i, raw.v);
            }
    };
    auto in_parallel = [&](auto &&body) {                        // body(lo, hi) over the row ranges; false if a thread could not start
        std::vector<std::thread> pool;
        bool ok = true;
        for (size_t k = 1; k < workers && ok; ++k) {
            const int lo = (int)((long long)m * (long long)k / (long long)workers), hi = (int)((long long)m * (long long)(k + 1) / (long long)workers);
            try { pool.emplace_back([&body, lo, hi] { body(lo, hi); }); } catch (...) { ok = false; }
        }
        if (ok) body(0, (int)((long long)m / (long long)workers));
        for (std::thread &t : pool) t.join();
        return ok;
    };
    std::vector<long long> fill((size_t)m + 1, 0);
    auto count_rows = [&](int lo, int hi) { scan_rows(lo, hi, [&](int r, int, double) { fill[r]++; }); };
    if (!in_parallel(count_rows)) { workers = 1; std::fill(fill.begin(), fill.end(), 0); count_rows(0, m); }
    long long all = 0;
    for (int i = 0; i < m; ++i) all += fill[i];
    if (all > (long long)INT_MAX) return -5;
    int *rp = (int *)malloc(sizeof(int) * ((size_t)m + 1));
    if (!rp) return -5;
    long long run = 0;
    for (int i = 0; i < m; ++i) { rp[i] = (int)run; run += fill[i]; fill[i] = 0; }
    rp[m] = (int)run;
    size_t total = (size_t)run;
    int *ci = (int *)malloc(sizeof(int) * (total ? total : 1));
    double *v = (double *)malloc(sizeof(double) * (total ? total : 1));
    if (!ci || !v) { free(rp); free(ci); free(v); return -5; }
    auto place_rows = [&](int lo, int hi) {
        scan_rows(lo, hi, [&](int r, int c, double val) {
            const size_t p = (size_t)rp[r] + (size_t)fill[r]++;
            ci[p] = c; v[p] = val;
        });
    };
    if (!in_parallel(place_rows)) {                              // (threads ran out between the two passes: redo the pass alone)
        workers = 1;
        std::fill(fill.begin(), fill.end(), 0);
        place_rows(0, m);
    }
    out->choice = true; out->row = m; out->col = n; out->nnz = (int)total;
    out->row_ind = rp; out->col_ind = ci; out->values = v;
    if (trace)
        fprintf(stderr, "[ias host trace] ias_mtx_load %s: read %.1f ms, parse %.1f ms (%zu piece%s), CSR build %.1f ms, %zu entries\n",
                path, ms_read, ms_parse, pieces, pieces == 1 ? "" : "s", ms_since(t_stage), total);
    return 0;
}

extern "C" {

void ias_free_host_csr(IasCsrMatrix *m)
{
    if (!m) return;
    free(m->row_ind); free(m->col_ind); free(m->values);
    m->row_ind = nullptr; m->col_ind = nullptr; m->values = nullptr;
}

}  // extern "C"
