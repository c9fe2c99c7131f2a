/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" wrappers around the reference's OWN, UNMODIFIED CPU headers, which are
 * #included from where they lie (-I/root/reference/IA-SPGEMM-CPU_release); nothing of the
 * reference is copied into this repository.  The result, oracle/_ref/libiaref.so, is used to
 *   (a) pin oracle/ia_oracle.c (tests/test_oracle_pinning.py),
 *   (b) generate tests/golden/ fixtures (tests/golden/make_golden.py),
 *   (c) time the reference's CPU path beside the GPU engine (bench.py cpu_baseline.kind="reference").
 * The reference kernels take 2-D arrays (malloc2d) for DIA/ELL; the wrappers flatten to
 * row-major so results can be compared with flat arrays.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <vector>

#ifndef VALUE_TYPE
#define VALUE_TYPE double
#endif

#include "detail/dense/common_dense.h"
#include "detail/csr/common_csr.h"
#include "detail/coo/common_coo.h"
#include "detail/dia/common_dia.h"
#include "detail/ell/common_ell.h"

static double now_ms()
{
    timeval t; gettimeofday(&t, NULL);
    return t.tv_sec * 1000.0 + t.tv_usec / 1000.0;
}

static void wrap_csr(CsrMatrix *M, int rows, int cols, int nnz, int *rp, int *ci, double *v)
{
    M->row = rows; M->col = cols; M->nnz = nnz; M->row_ind = rp; M->col_ind = ci; M->values = v;
}

extern "C" {

void ref_free(void *p) { free(p); }
int ref_mkl_threads() { return mkl_get_max_threads(); }
int ref_omp_threads() { return omp_get_max_threads(); }
/* launchers such as torchrun export OMP_NUM_THREADS=1: the timed baseline sets its thread counts explicitly */
extern "C" void mkl_serv_set_num_threads(int);
void ref_set_threads(int n)
{
    if (n < 1) n = 1;
    omp_set_num_threads(n);
    mkl_serv_set_num_threads(n);
}
void ref_mkl_version(char *buf, int len) { MKL_Get_Version_String(buf, len); }

/* CSR_MUL_CSR (common_csr.h:85-193); returns elapsed ms of the call */
double ref_csr_mul_csr(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                       int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v,
                       int *c_nnz, int **c_rp, int **c_ci, double **c_v)
{
    CsrMatrix A, B, C;
    wrap_csr(&A, a_rows, a_cols, a_nnz, a_rp, a_ci, a_v);
    wrap_csr(&B, b_rows, b_cols, b_nnz, b_rp, b_ci, b_v);
    double t0 = now_ms();
    CSR_MUL_CSR(&A, &B, &C);
    double t1 = now_ms();
    *c_nnz = C.nnz; *c_rp = C.row_ind; *c_ci = C.col_ind; *c_v = C.values;
    return t1 - t0;
}

/* MKL_MUL_MKL (common_csr.h:18-47).  keep=0 drops the result right after timing. */
double ref_mkl_mul_mkl(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                       int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v,
                       int keep, int *c_nnz, int **c_rp, int **c_ci, double **c_v)
{
    MKLMatrix A, B, C;
    A.row = a_rows; A.col = a_cols; A.nnz = a_nnz; A.row_ind = a_rp; A.col_ind = a_ci; A.values = a_v;
    B.row = b_rows; B.col = b_cols; B.nnz = b_nnz; B.row_ind = b_rp; B.col_ind = b_ci; B.values = b_v;
    double t0 = now_ms();
    MKL_MUL_MKL(&A, &B, &C);
    double t1 = now_ms();
    *c_nnz = C.nnz;
    if (keep) {
        /* MKL owns the exported arrays; hand the caller malloc'd copies */
        int *rp = (int *)malloc(sizeof(int) * ((size_t)C.row + 1));
        int *ci = (int *)malloc(sizeof(int) * (size_t)(C.nnz > 0 ? C.nnz : 1));
        double *v = (double *)malloc(sizeof(double) * (size_t)(C.nnz > 0 ? C.nnz : 1));
        memcpy(rp, C.row_ind, sizeof(int) * ((size_t)C.row + 1));
        memcpy(ci, C.col_ind, sizeof(int) * (size_t)C.nnz);
        memcpy(v, C.values, sizeof(double) * (size_t)C.nnz);
        *c_rp = rp; *c_ci = ci; *c_v = v;
    }
    return t1 - t0;
}

long long ref_getflop(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                      int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v)
{
    CsrMatrix A, B;
    wrap_csr(&A, a_rows, a_cols, a_nnz, a_rp, a_ci, a_v);
    wrap_csr(&B, b_rows, b_cols, b_nnz, b_rp, b_ci, b_v);
    return GetFlop(&A, &B);
}

double ref_sizeof_csr(int rows, int cols, int nnz)
{
    CsrMatrix A; A.row = rows; A.col = cols; A.nnz = nnz;
    return sizeofcsr(&A);
}

/* 26 features in the order main.cpp:655-679 fills them (gate 50x as on the CPU side) */
void ref_features26(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                    int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v, double *f)
{
    CsrMatrix A, B;
    wrap_csr(&A, a_rows, a_cols, a_nnz, a_rp, a_ci, a_v);
    wrap_csr(&B, b_rows, b_cols, b_nnz, b_rp, b_ci, b_v);
    for (int i = 0; i < 26; i++) f[i] = 0.0;
    GetInfo1(&A, f);
    GetInfo1(&B, f + 9);
    DiaMatrix Ad, Bd; EllMatrix Ae, Be;
    CSRtoDIA(&A, &Ad); CSRtoDIA(&B, &Bd);
    GetInfo2(&Ad, f + 18); GetInfo2(&Bd, f + 21);
    CSRtoELL(&A, &Ae); CSRtoELL(&B, &Be);
    GetInfo3(&Ae, f + 24); GetInfo3(&Be, f + 25);
}

/* CSRtoDIA (common_dia.h:29-96), flattened */
int ref_csr_to_dia(int rows, int cols, int nnz, int *rp, int *ci, double *v,
                   int *choice, int *nd, int **diag_ind, int **offsets, double **values)
{
    CsrMatrix A; wrap_csr(&A, rows, cols, nnz, rp, ci, v);
    DiaMatrix D;
    CSRtoDIA(&A, &D);
    *choice = D.choice ? 1 : 0; *nd = D.num_diagonals;
    *diag_ind = NULL; *offsets = NULL; *values = NULL;
    if (!D.choice) return 0;
    *diag_ind = D.diagonal_ind; *offsets = D.diagonal_offsets;
    double *flat = (double *)calloc((size_t)rows * D.num_diagonals + 1, sizeof(double));
    for (int i = 0; i < rows; i++)
        for (int d = 0; d < D.num_diagonals; d++) flat[(size_t)i * D.num_diagonals + d] = D.values[i][d];
    free2d(D.values);
    *values = flat;
    return 0;
}

/* DIA_mul_DIA (common_dia.h:101-195) on operands converted by the reference's CSRtoDIA */
double ref_dia_mul_dia_from_csr(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                                int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v,
                                int *ok, int *c_nd, int **c_diag_ind, int **c_off, double **c_val)
{
    CsrMatrix A, B;
    wrap_csr(&A, a_rows, a_cols, a_nnz, a_rp, a_ci, a_v);
    wrap_csr(&B, b_rows, b_cols, b_nnz, b_rp, b_ci, b_v);
    DiaMatrix Ad, Bd, Cd;
    CSRtoDIA(&A, &Ad); CSRtoDIA(&B, &Bd);
    *ok = (Ad.choice && Bd.choice) ? 1 : 0;
    if (!*ok) return 0.0;
    double t0 = now_ms();
    DIA_mul_DIA(&Ad, &Bd, &Cd);
    double t1 = now_ms();
    *c_nd = Cd.num_diagonals; *c_diag_ind = Cd.diagonal_ind; *c_off = Cd.diagonal_offsets;
    double *flat = (double *)calloc((size_t)a_rows * Cd.num_diagonals + 1, sizeof(double));
    for (int i = 0; i < a_rows; i++)
        for (int d = 0; d < Cd.num_diagonals; d++) flat[(size_t)i * Cd.num_diagonals + d] = Cd.values[i][d];
    *c_val = flat;
    free2d(Cd.values); FreeDiaMatrix(&Ad); FreeDiaMatrix(&Bd);
    return t1 - t0;
}

/* CSRtoELL (common_ell.h:30-77), flattened */
int ref_csr_to_ell(int rows, int cols, int nnz, int *rp, int *ci, double *v,
                   int *choice, int *width, int **nnz_row, int **col_ind, double **values)
{
    CsrMatrix A; wrap_csr(&A, rows, cols, nnz, rp, ci, v);
    EllMatrix E;
    CSRtoELL(&A, &E);
    *choice = E.choice ? 1 : 0; *width = E.max_nnz_per_row;
    *nnz_row = NULL; *col_ind = NULL; *values = NULL;
    if (!E.choice) return 0;
    int w = E.max_nnz_per_row;
    int *fc = (int *)calloc((size_t)rows * w + 1, sizeof(int));
    double *fv = (double *)calloc((size_t)rows * w + 1, sizeof(double));
    for (int i = 0; i < rows; i++)
        for (int k = 0; k < w; k++) { fc[(size_t)i * w + k] = E.col_ind[i][k]; fv[(size_t)i * w + k] = E.values[i][k]; }
    *nnz_row = E.nnz_row; *col_ind = fc; *values = fv;
    free2d(E.col_ind); free2d(E.values);
    return 0;
}

/* ELL_MUL_ELL (common_ell.h:80-189) on operands converted by the reference's CSRtoELL */
double ref_ell_mul_ell_from_csr(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                                int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v,
                                int *ok, int *c_w, int *c_nnz, int **c_nr, int **c_ci, double **c_v)
{
    CsrMatrix A, B;
    wrap_csr(&A, a_rows, a_cols, a_nnz, a_rp, a_ci, a_v);
    wrap_csr(&B, b_rows, b_cols, b_nnz, b_rp, b_ci, b_v);
    EllMatrix Ae, Be, Ce;
    CSRtoELL(&A, &Ae); CSRtoELL(&B, &Be);
    *ok = (Ae.choice && Be.choice) ? 1 : 0;
    if (!*ok) return 0.0;
    double t0 = now_ms();
    ELL_MUL_ELL(&Ae, &Be, &Ce);
    double t1 = now_ms();
    int w = Ce.max_nnz_per_row;
    *c_w = w; *c_nnz = Ce.nnz; *c_nr = Ce.nnz_row;
    int *fc = (int *)calloc((size_t)a_rows * w + 1, sizeof(int));
    double *fv = (double *)calloc((size_t)a_rows * w + 1, sizeof(double));
    for (int i = 0; i < a_rows; i++)
        for (int k = 0; k < w; k++) { fc[(size_t)i * w + k] = Ce.col_ind[i][k]; fv[(size_t)i * w + k] = Ce.values[i][k]; }
    *c_ci = fc; *c_v = fv;
    free2d(Ce.col_ind); free2d(Ce.values);
    return t1 - t0;
}

/* COO_MUL_COO (common_coo.h:72-161) on operands converted by the reference's CSRtoCOO */
double ref_coo_mul_coo_from_csr(int a_rows, int a_cols, int a_nnz, int *a_rp, int *a_ci, double *a_v,
                                int b_rows, int b_cols, int b_nnz, int *b_rp, int *b_ci, double *b_v,
                                int *c_nnz, int **c_ro, int **c_ri, int **c_ci, double **c_v)
{
    CsrMatrix A, B;
    wrap_csr(&A, a_rows, a_cols, a_nnz, a_rp, a_ci, a_v);
    wrap_csr(&B, b_rows, b_cols, b_nnz, b_rp, b_ci, b_v);
    CooMatrix Ac, Bc, Cc;
    CSRtoCOO(&A, &Ac); CSRtoCOO(&B, &Bc);
    double t0 = now_ms();
    COO_MUL_COO(&Ac, &Bc, &Cc);
    double t1 = now_ms();
    *c_nnz = Cc.nnz; *c_ro = Cc.row_offset; *c_ri = Cc.row_ind; *c_ci = Cc.col_ind; *c_v = Cc.values;
    return t1 - t0;
}

} /* extern "C" */
