/*
 * Shim "mkl.h" -- TEST INFRASTRUCTURE ONLY (used when building oracle/_ref/libiaref.so).
 *
 * Intel MKL's development package is not in this image, but torch's libtorch_cpu.so
 * statically embeds oneMKL 2024.2 and exports the inspector-executor sparse entry points.
 * This header declares exactly the symbols the reference's MKL_MUL_MKL
 * (IA-SPGEMM-CPU_release/detail/csr/common_csr.h:18-47) needs, and maps
 * mkl_sparse_sp2m(N, descr, A, N, descr, B, FULL_MULT, &C) onto mkl_sparse_spmm(N, A, B, &C),
 * the equivalent call the reference keeps commented out at common_csr.h:39.
 */
#ifndef IAS_SHIM_MKL_H
#define IAS_SHIM_MKL_H
#include <stddef.h>
#include <sys/time.h>   /* the reference's utime.h uses timeval without including it */

typedef int MKL_INT;
struct sparse_matrix;
typedef struct sparse_matrix *sparse_matrix_t;

typedef enum { SPARSE_STATUS_SUCCESS = 0 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50 } sparse_diag_type_t;
typedef enum { SPARSE_STAGE_FULL_MULT = 90 } sparse_request_t;
struct matrix_descr { sparse_matrix_type_t type; sparse_fill_mode_t mode; sparse_diag_type_t diag; };

#ifdef __cplusplus
extern "C" {
#endif
sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols,
                                        MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values);
sparse_status_t mkl_sparse_spmm(sparse_operation_t op, const sparse_matrix_t A, const sparse_matrix_t B, sparse_matrix_t *C);
sparse_status_t mkl_sparse_d_export_csr(const sparse_matrix_t src, sparse_index_base_t *indexing, MKL_INT *rows, MKL_INT *cols,
                                        MKL_INT **rows_start, MKL_INT **rows_end, MKL_INT **col_indx, double **values);
sparse_status_t mkl_sparse_destroy(sparse_matrix_t A);
void *mkl_serv_malloc(size_t size, int align);
void mkl_serv_free(void *p);
int mkl_get_max_threads(void);
void MKL_Get_Version_String(char *buf, int len);
#ifdef __cplusplus
}
#endif

static inline void *mkl_malloc(size_t size, int align) { return mkl_serv_malloc(size, align); }
static inline void mkl_free(void *p) { mkl_serv_free(p); }
static inline sparse_status_t mkl_sparse_sp2m(sparse_operation_t opA, struct matrix_descr dA, const sparse_matrix_t A,
                                              sparse_operation_t opB, struct matrix_descr dB, const sparse_matrix_t B,
                                              sparse_request_t req, sparse_matrix_t *C)
{
    (void)dA; (void)opB; (void)dB; (void)req;
    return mkl_sparse_spmm(opA, A, B, C);
}
#endif
