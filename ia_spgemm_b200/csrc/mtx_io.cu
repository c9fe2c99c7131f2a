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

#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

std::string lower(const char *s)
{
    std::string r(s);
    for (char &ch : r) ch = (char)tolower((unsigned char)ch);
    return r;
}

struct Entry { int i, j; double v; };

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

    int m = 0, n = 0, nz = 0;
    bool have_size = false;
    while (fgets(line, sizeof line, f)) {
        if (line[0] == '%') continue;
        if (sscanf(line, "%d %d %d", &m, &n, &nz) == 3) { have_size = true; break; }
    }
    if (!have_size || m < 0 || n < 0 || nz < 0) { fclose(f); return -4; }

    std::vector<Entry> e;
    e.reserve((size_t)(nz < (1 << 24) ? nz : (1 << 24)));      // the size line is input: do not trust it with memory
    for (int t = 0; t < nz; ++t) {
        Entry x{0, 0, 1.0};
        int iv = 0;
        int got = field == "real" ? fscanf(f, "%d %d %lg", &x.i, &x.j, &x.v)
                : field == "integer" ? fscanf(f, "%d %d %d", &x.i, &x.j, &iv)
                                     : fscanf(f, "%d %d", &x.i, &x.j);
        if (got != (field == "pattern" ? 2 : 3)) break;
        if (field == "integer") x.v = iv;
        --x.i; --x.j;
        if (x.i < 0 || x.i >= m || x.j < 0 || x.j >= n) continue;   // the reference would write / index out of bounds here
        e.push_back(x);
    }
    fclose(f);

    std::vector<long long> fill((size_t)m + 1, 0);
    for (const Entry &x : e) {
        fill[x.i]++;
        if (mirror && x.i != x.j && x.j < m && x.i < n) fill[x.j]++;
    }
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
    for (const Entry &x : e) {
        size_t p = (size_t)rp[x.i] + (size_t)fill[x.i]++;
        ci[p] = x.j; v[p] = x.v;
        if (mirror && x.i != x.j && x.j < m && x.i < n) {
            p = (size_t)rp[x.j] + (size_t)fill[x.j]++;
            ci[p] = x.i; v[p] = x.v;
        }
    }
    out->choice = true; out->row = m; out->col = n; out->nnz = (int)total;
    out->row_ind = rp; out->col_ind = ci; out->values = v;
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
