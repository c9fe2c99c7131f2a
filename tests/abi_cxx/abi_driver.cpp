// abi_driver.cpp -- a compiled C++ caller of the C ABI (no ctypes, no Python): what a reference-side `main`
// does through include/iaspgemm.h.  Operands are declared with the reference's own struct shapes (the layout
// is proven against the reference headers by layout_check.cpp); results are checked against known answers:
//   dia.mtx  A^2  (SURVEY appendix B / CPU/1.jpg): row_ptr 0 3 6 8 9, columns, values, DIA offsets 0 1 2,
//   Poisson 64x64 A^2: nnz(C) = 13 N^2 - 20 N + 4, sum(C) = 16 + 4 (N - 2).
// Built and run by tests/test_abi_cxx_gpu.py (g++ on the GPU box; the same image has gcc).  Prints "ABI_CXX_OK".
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/iaspgemm.h"

#define CHECK(x)                                                                              \
    do {                                                                                      \
        if (!(x)) { printf("FAILED %s:%d: %s   [%s]\n", __FILE__, __LINE__, #x, ias_last_error()); return 1; } \
    } while (0)

struct Seen { long long entries; double sum; int batches; long long nnz_total; int last_row; };

static int consume(const IasStreamBatch *b, void *user)
{
    Seen *s = (Seen *)user;
    std::vector<double> v((size_t)b->batch_nnz);
    std::vector<long long> rp((size_t)(b->row_end - b->row_begin) + 1);
    if (ias_copy(rp.data(), b->row_ptr_dev, sizeof(long long) * rp.size(), 1)) return IAS_E_CUDA;
    if (b->batch_nnz && ias_copy(v.data(), b->values_dev, sizeof(double) * v.size(), 1)) return IAS_E_CUDA;
    if (rp.front() != b->entry_base || rp.back() - rp.front() != b->batch_nnz) return 77;
    if (b->row_begin != s->last_row) return 78;            // batches arrive in row order, without gaps
    s->last_row = b->row_end;
    for (double x : v) s->sum += x;
    s->entries += b->batch_nnz;
    s->batches++;
    s->nnz_total = b->nnz_total;
    return 0;
}

int main()
{
    CHECK(ias_init(0) == IAS_OK);

    // ---- dia.mtx: 4x4 pattern, entries (1,1) (1,2) (2,2) (2,3) (3,3) (3,4) (4,4) -> 1.0
    int a_rp[5] = {0, 2, 4, 6, 7}, a_ci[7] = {0, 1, 1, 2, 2, 3, 3};
    double a_v[7] = {1, 1, 1, 1, 1, 1, 1};
    IasCsrMatrix A;
    A.choice = true; A.row = 4; A.col = 4; A.nnz = 7; A.row_ind = a_rp; A.col_ind = a_ci; A.values = a_v;

    // CSR_MUL_CSR(A, A, C) on host operands
    long long *c_rp = nullptr, c_nnz = 0;
    int *c_ci = nullptr;
    double *c_v = nullptr, h2d = 0, d2h = 0;
    IasSpgemmStats st;
    CHECK(ias_csr_mul_csr_host(&A, &A, &c_rp, &c_ci, &c_v, &c_nnz, &st, &h2d, &d2h) == IAS_OK);
    const long long want_rp[5] = {0, 3, 6, 8, 9};
    const int want_ci[9] = {0, 1, 2, 1, 2, 3, 2, 3, 3};
    const double want_v[9] = {1, 2, 1, 1, 2, 1, 1, 2, 1};
    CHECK(c_nnz == 9 && st.products == 12 && st.kernel_launches > 0);
    CHECK(memcmp(c_rp, want_rp, sizeof want_rp) == 0);
    CHECK(memcmp(c_ci, want_ci, sizeof want_ci) == 0);
    CHECK(memcmp(c_v, want_v, sizeof want_v) == 0);

    // device operands: int32 reference layout, DIA, ELL, COO
    IasCsrMatrixDev dA;
    CHECK(ias_upload_csr(&A, &dA) == IAS_OK);
    IasCsrMatrixDev dC;
    double ms = 0;
    CHECK(ias_csr_mul_csr_dev(&dA, &dA, &dC, &ms) == IAS_OK && dC.nnz == 9 && ms > 0);
    int rp32[5];
    CHECK(ias_download_csr(&dC, rp32, nullptr, nullptr) == IAS_OK && rp32[4] == 9 && rp32[1] == 3);
    ias_free_csr_dev(&dC);

    IasDiaDev a_dia, c_dia, c_rm, back;
    CHECK(ias_csr_to_dia(&dA, 20.0, &a_dia) == IAS_OK && a_dia.choice && a_dia.num_diagonals == 2);
    CHECK(ias_dia_mul_dia_dev(&a_dia, &a_dia, &c_dia, &ms) == IAS_OK && c_dia.num_diagonals == 3);
    int off[3];
    double dv[12];
    CHECK(ias_download_dia(&c_dia, nullptr, off, dv) == IAS_OK);
    CHECK(off[0] == 0 && off[1] == 1 && off[2] == 2);
    const double want_dia[12] = {1, 2, 1, 1, 2, 1, 1, 2, 0, 1, 0, 0};        // row-major [row][slot], dia:162-193
    CHECK(memcmp(dv, want_dia, sizeof want_dia) == 0);
    // a reference-built (row-major) DiaMatrixDev goes through ias_dia_relayout
    CHECK(ias_dia_relayout(&c_dia, 1, &c_rm) == IAS_OK);
    double raw[12];
    CHECK(ias_copy(raw, c_rm.values_dev, sizeof raw, 1) == IAS_OK && memcmp(raw, want_dia, sizeof raw) == 0);
    CHECK(ias_dia_relayout(&c_rm, 0, &back) == IAS_OK);
    CHECK(ias_download_dia(&back, nullptr, nullptr, dv) == IAS_OK && memcmp(dv, want_dia, sizeof want_dia) == 0);
    ias_free_dia_dev(&a_dia); ias_free_dia_dev(&c_dia); ias_free_dia_dev(&c_rm); ias_free_dia_dev(&back);

    IasEllDev a_ell, c_ell;
    CHECK(ias_csr_to_ell(&dA, 20.0, &a_ell) == IAS_OK && a_ell.choice && a_ell.max_nnz_per_row == 2 && a_ell.nnz == 7);
    CHECK(ias_ell_mul_ell_dev(&a_ell, &a_ell, &c_ell, &ms) == IAS_OK && c_ell.max_nnz_per_row == 3 && c_ell.nnz == 9);
    int nr[4], eci[12];
    double ev[12];
    CHECK(ias_download_ell(&c_ell, nr, eci, ev) == IAS_OK);
    CHECK(nr[0] == 3 && nr[1] == 3 && nr[2] == 2 && nr[3] == 1 && eci[3] == 1 && ev[4] == 2.0 && ev[11] == 0.0);
    ias_free_ell_dev(&a_ell); ias_free_ell_dev(&c_ell);

    IasCooDev a_coo, c_coo;
    CHECK(ias_csr_to_coo(&dA, &a_coo) == IAS_OK && a_coo.nnz == 7);
    CHECK(ias_coo_mul_coo_dev(&a_coo, &a_coo, &c_coo, &ms) == IAS_OK && c_coo.nnz == 9);
    int ro[5], ri[9], ci2[9];
    double cv[9];
    CHECK(ias_download_coo(&c_coo, ro, ri, ci2, cv) == IAS_OK);
    CHECK(ro[4] == 9 && ri[0] == 0 && ri[8] == 3 && memcmp(ci2, want_ci, sizeof want_ci) == 0 && memcmp(cv, want_v, sizeof want_v) == 0);
    ias_free_coo_dev(&a_coo); ias_free_coo_dev(&c_coo);

    long long flop = 0;
    double f26[26];
    CHECK(ias_getflop(&dA, &dA, &flop) == IAS_OK && flop == 12);
    CHECK(ias_features26(&dA, &dA, f26) == IAS_OK && f26[0] == 4 && f26[2] == 7 && f26[3] == 0.4375 && f26[18] == 2 && f26[24] == 0.875);
    CHECK(ias_sizeof_csr(4, 9) == 140.0);
    ias_free_csr_dev(&dA);

    // ---- Poisson 64 x 64 through the streaming entry with a consumer callback
    const int N = 64;
    IasCsrMatrixDev P;
    CHECK(ias_gen_poisson2d(N, N, &P) == IAS_OK);
    Seen seen = {0, 0.0, 0, 0, 0};
    CHECK(ias_csr_mul_csr_stream_cb(&P, &P, 0, P.row, 12 * 9000, nullptr, consume, &seen, &st) == IAS_OK);
    const long long want_nnz = 13LL * N * N - 20 * N + 4;
    CHECK(st.nnz == want_nnz && seen.nnz_total == want_nnz && seen.entries == want_nnz);
    CHECK(seen.batches == st.batches && st.batches > 1 && seen.last_row == P.row);
    CHECK(fabs(seen.sum - (16.0 + 4.0 * (N - 2))) < 1e-9 && fabs(st.checksum - seen.sum) < 1e-9);
    // a consumer that refuses stops the multiply with its own status
    struct Refuse { static int f(const IasStreamBatch *, void *) { return 42; } };
    CHECK(ias_csr_mul_csr_stream_cb(&P, &P, 0, P.row, 12 * 9000, nullptr, Refuse::f, nullptr, &st) == 42);
    ias_free_csr_dev(&P);
    ias_release_host();
    printf("ABI_CXX_OK\n");
    return 0;
}
