/*
 * iaspgemm.h -- C ABI of the B200-native IA-SpGEMM engine (libiaspgemm.so).
 *
 * Drop-in boundary for the reference's SpGEMM hot path.  The reference has no FFI layer: its
 * `main` calls `#include`d free functions on POD structs.  Every entry point below names the
 * reference function it replaces (paths relative to the reference checkout,
 * GPU = IA-SPGEMM-GPU_release, CPU = IA-SPGEMM-CPU_release).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types;
 *   - every function returns an int status (IAS_OK == 0); nothing calls exit() (the reference's
 *     DevMalloc/DevUpload exit(1), GPU/detail/common.h:62-97);
 *   - the struct layouts are byte-identical to GPU/detail/format.h so a reference `main` can
 *     pass its own CsrMatrix / CsrMatrixDev / DiaMatrixDev / EllMatrixDev / CooMatrixDev objects;
 *   - callee allocates every array of C (as the reference kernels do) from the engine's
 *     stream-ordered device pool; the caller releases with ias_free_*;
 *   - index type is int32 (reference layout); counts that overflow int32 at the BASELINE sizes
 *     (products, nnz(C), row pointers of C) are 64-bit: IasCsr64Dev is the native result type
 *     and the int32 CsrMatrixDev result is offered where nnz(C) < 2^31;
 *   - not re-entrant: one engine context per process/GPU (the reference passes operands through
 *     file-scope globals, CPU/main.cpp:25-41).
 */
#ifndef IASPGEMM_H
#define IASPGEMM_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- status codes */
enum {
    IAS_OK = 0,
    IAS_E_CUDA = 1,        /* a CUDA call failed; see ias_last_error() */
    IAS_E_ARG = 2,         /* bad argument (NULL, negative size, shape mismatch) */
    IAS_E_OVERFLOW = 3,    /* result does not fit the int32 reference layout */
    IAS_E_NOMEM = 4,       /* device or host allocation failed */
    IAS_E_GATE = 5,        /* format rejected by the size gate (choice == false) */
    IAS_E_IO = 6           /* file error; ias_mtx_load returns the reference's -1..-4 instead */
};

/* ---------------------------------------------------------------- formats (GPU/detail/format.h) */
typedef struct {           /* CsrMatrix, GPU/detail/format.h:47-57 (CPU/detail/format.h:29-39) */
    bool choice;
    int row, col, nnz;
    int *row_ind;          /* row pointers, row+1 entries (the reference's name) */
    int *col_ind;
    double *values;
} IasCsrMatrix;

typedef struct {           /* CsrMatrixDev, GPU/detail/format.h:59-69 */
    bool choice;
    int row, col, nnz;
    int *row_ind_dev;
    int *col_ind_dev;
    double *values_dev;
} IasCsrMatrixDev;

typedef struct {           /* native result: 64-bit row pointers, column-sorted rows */
    int row, col;
    long long nnz;
    long long *row_ptr_dev;   /* row+1 entries */
    int *col_ind_dev;
    double *values_dev;
} IasCsr64Dev;

typedef struct {           /* CooMatrixDev, GPU/detail/format.h:29-40 (+ 64-bit offsets) */
    bool choice;
    int row, col;
    long long nnz;
    long long *row_offset_dev;  /* row+1 entries (the reference keeps a CSR-like row_offset) */
    int *row_ind_dev;
    int *col_ind_dev;
    double *values_dev;
} IasCooDev;

typedef struct {           /* DiaMatrixDev, GPU/detail/format.h:82-92 */
    bool choice;
    int row, col, num_diagonals;
    int *diagonal_ind_dev;      /* row+col-1 entries: map index -> slot (0 for absent, as the reference) */
    int *diagonal_offsets_dev;  /* num_diagonals entries, ascending */
    double *values_dev;         /* DIAGONAL-MAJOR on device: values[slot*row + i] (coalesced);
                                   ias_download_dia returns the reference's row-major [i*nd + slot] */
} IasDiaDev;

typedef struct {           /* EllMatrixDev, GPU/detail/format.h:108-119 */
    bool choice;
    int row, col;
    long long nnz;
    int max_nnz_per_row;
    int *nnz_row_dev;
    int *col_ind_dev;           /* row-major [i*width + k] as the reference; padding 0 / 0.0 */
    double *values_dev;
} IasEllDev;

/* per-call statistics of the CSR pipeline (all times in ms, CUDA events on the engine stream) */
typedef struct {
    long long products;         /* GetFlop(A,B): intermediate products in the processed rows */
    long long nnz;              /* nnz(C) of the processed rows */
    double ms_total;            /* symbolic + allocation + numeric (+ consumer in streaming mode) */
    double ms_analyze, ms_symbolic, ms_scan, ms_numeric, ms_consume;
    double ms_bin_sym[8];       /* device time of each symbolic bin kernel (same bin order as below) */
    double ms_bin_num[8];       /* device time of each numeric bin kernel, summed over batches */
    long long sym_bin_rows[8];  /* rows per symbolic bin: empty, tiny, warp, cta-s, cta-l, global, -, - */
    long long num_bin_rows[8];  /* rows per numeric bin */
    int batches;                /* row batches used (1 unless streaming) */
    int kernel_launches;        /* engine kernels launched by this call */
    double checksum;            /* streaming only: sum of all C values */
    unsigned long long structure_hash; /* streaming only: order-independent hash of (row, col) pairs */
} IasSpgemmStats;

/* ---------------------------------------------------------------- context */
int ias_init(int device);                       /* binds the engine to a CUDA device; idempotent */
int ias_set_stream(void *cuda_stream);          /* run on the caller's cudaStream_t (NULL: engine stream) */
int ias_sync(void);
const char *ias_last_error(void);
const char *ias_version(void);
int ias_device_info(int *sm_count, size_t *smem_optin, size_t *free_bytes, size_t *total_bytes);
long long ias_kernel_launches(void);            /* engine kernels launched since ias_init */
/* Kernel-selection knobs (the reference has none: it hard-codes its library calls, GPU/main.cu:470-521).
 * Names: "global_rows_smem" (1 = windowed shared-memory kernels for rows beyond the CTA hash, 0 = L2 bitmap kernels),
 * "gwin_swords", "gwin_win", "gwin_sym_swords" (window sizes, 0 = automatic), "gwin_smem_kb", "gwin_max_sw" (numeric windowed
 * kernel only up to this many column super-windows per row, 0 = no limit), "g_win" (accumulate window of the L2 kernel), "g_coop", "gwin_takes_b2" (0/1 switches kept for A/B runs).  Also read from
 * IAS_OPT_<NAME> in the environment by ias_init.  Results do not depend on any of them. */
int ias_set_option(const char *name, long long value);
int ias_get_option(const char *name, long long *value);

/* ---------------------------------------------------------------- transfers */
/* UploadCsrMatrix, GPU/detail/csr_dev/common_csr_dev.h:111-125 */
int ias_upload_csr(const IasCsrMatrix *host, IasCsrMatrixDev *dev);
/* FreeCsrMatrixDev, csr_dev:285-297 */
int ias_free_csr_dev(IasCsrMatrixDev *m);
int ias_free_csr64_dev(IasCsr64Dev *m);
/* device -> caller-provided host arrays (row+1, nnz, nnz entries) */
int ias_download_csr64(const IasCsr64Dev *dev, long long *row_ptr, int *col_ind, double *values);
int ias_download_csr(const IasCsrMatrixDev *dev, int *row_ptr, int *col_ind, double *values);
/* raw copies on the engine stream, synchronous: kind 0 = host->device, 1 = device->host, 2 = device->device
 * (DevUpload / DevDownload, GPU/detail/common.h:79-97, without the exit(1)) */
int ias_copy(void *dst, const void *src, size_t bytes, int kind);
/* The engine remembers, per B operand (pointer + shape), whether its rows are canonical.  A caller that
 * rewrites an operand's device arrays IN PLACE must call this before the next multiply (ias_free_csr_dev
 * does it for engine-owned operands; NULL forgets everything). */
int ias_forget_operand(const IasCsrMatrixDev *m);
/* canonical = every row strictly increasing in column (sorted, duplicate free) */
int ias_csr_is_canonical(const IasCsrMatrixDev *m, int *canonical);

/* ---------------------------------------------------------------- Algorithm 2 / the CSR hot path */
/* CSR_MUL_CSR_DEV, GPU/detail/csr_dev/common_csr_dev.h:134-254 (and CUSPARSE_MUL_CUSPARSE,
 * GPU/detail/cusparse/common_cusparse.h:29-96; cusp::multiply, GPU/main.cu:482): C = A*B on device
 * operands.  Timed region as CUSPARSE_MUL_CUSPARSE: symbolic, allocation of C, numeric, sort. */
int ias_csr_mul_csr_dev64(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, IasCsr64Dev *C,
                          IasSpgemmStats *stats);
/* same, int32 reference layout; IAS_E_OVERFLOW when nnz(C) >= 2^31 */
int ias_csr_mul_csr_dev(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, IasCsrMatrixDev *C,
                        double *elapsed_ms);
/* rows [row_begin,row_end) of C only (multi-GPU row blocks); C->row = row_end-row_begin */
int ias_csr_mul_csr_rows_dev64(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B,
                               int row_begin, int row_end, IasCsr64Dev *C, IasSpgemmStats *stats);
/* streaming: C of rows [row_begin,row_end) is produced in HBM-budgeted row batches and reduced
 * on device (nnz, checksum, structure hash, optional per-row nnz) -- for results that do not
 * fit (R-MAT scale >= 20).  row_nnz_dev may be NULL; budget_bytes 0 = 60% of free memory. */
int ias_csr_mul_csr_stream(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B,
                           int row_begin, int row_end, size_t budget_bytes,
                           int *row_nnz_dev, IasSpgemmStats *stats);
/* host operands, host result: CSR_MUL_CSR(A,B,C), CPU/detail/csr/common_csr.h:85 -- uploads
 * A (and B unless it aliases A), multiplies on the GPU, downloads C into pinned host memory owned
 * by the engine (valid until the next host call or ias_release_host).  nnz(C) via C_nnz. */
int ias_csr_mul_csr_host(const IasCsrMatrix *A, const IasCsrMatrix *B,
                         long long **c_row_ptr, int **c_col_ind, double **c_values,
                         long long *c_nnz, IasSpgemmStats *stats, double *ms_h2d, double *ms_d2h);
int ias_release_host(void);

/* GetFlop, CPU/detail/csr/common_csr.h:290-304 */
int ias_getflop(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, long long *products);
/* bytes of the B rows referenced at least once by rows [row_begin,row_end) of A: 4*|T| + 12*sum len(B_j), j in T
 * (the "touched B rows" term of the algorithmic-bytes model) */
int ias_touched_b_bytes(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int row_begin, int row_end, long long *bytes);
/* work-balanced contiguous row blocks: bounds[0..parts] with equal shares of products */
int ias_partition_rows(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int parts, int *bounds);
/* getsum_csr, csr_dev:264-273 (verified_sum) */
int ias_checksum(const double *values_dev, long long n, double *sum);
/* order-independent structure hash of a CSR result (same function as the streaming consumer) */
int ias_structure_hash(const IasCsr64Dev *C, int row_base, unsigned long long *hash);

/* ---------------------------------------------------------------- DIA (Algorithm 3) */
/* CSRtoDIA, CPU/detail/dia/common_dia.h:29-96 (gate: GPU/detail/dia/common_dia.h:51 uses 20x) */
int ias_csr_to_dia(const IasCsrMatrixDev *A, double gate, IasDiaDev *out);
/* DIA_MUL_DIA_DEV, GPU/detail/dia_dev/common_dia_dev.h:138-182 */
int ias_dia_mul_dia_dev(const IasDiaDev *A, const IasDiaDev *B, IasDiaDev *C, double *elapsed_ms);
int ias_download_dia(const IasDiaDev *dev, int *diagonal_ind, int *diagonal_offsets,
                     double *values_row_major);
int ias_free_dia_dev(IasDiaDev *m);

/* ---------------------------------------------------------------- ELL (Algorithm 4) */
/* CSRtoELL, CPU/detail/ell/common_ell.h:30-77 */
int ias_csr_to_ell(const IasCsrMatrixDev *A, double gate, IasEllDev *out);
/* ELL_MUL_ELL_DEV, GPU/detail/ell_dev/common_ell_dev.h:310-382 */
int ias_ell_mul_ell_dev(const IasEllDev *A, const IasEllDev *B, IasEllDev *C, double *elapsed_ms);
int ias_download_ell(const IasEllDev *dev, int *nnz_row, int *col_ind, double *values);
int ias_free_ell_dev(IasEllDev *m);

/* ---------------------------------------------------------------- COO (Algorithm 5) */
/* CSRtoCOO, CPU/detail/coo/common_coo.h:29-66 */
int ias_csr_to_coo(const IasCsrMatrixDev *A, IasCooDev *out);
/* COO_MUL_COO_DEV, GPU/detail/coo_dev/common_coo_dev.h:279-602 */
int ias_coo_mul_coo_dev(const IasCooDev *A, const IasCooDev *B, IasCooDev *C, double *elapsed_ms);
int ias_download_coo(const IasCooDev *dev, long long *row_offset, int *row_ind, int *col_ind,
                     double *values);
int ias_free_coo_dev(IasCooDev *m);

/* ---------------------------------------------------------------- features / density */
/* density representation, CPU/main.cpp:516-577: 128x128 int64 image, row-major */
int ias_density_image(const IasCsrMatrixDev *A, long long *img16384_host);
/* GetInfo1, CPU/detail/csr/common_csr.h:257-287 */
int ias_getinfo1(const IasCsrMatrixDev *A, double *f9);
/* GetInfo2, CPU/detail/dia/common_dia.h:222-233;  GetInfo3, CPU/detail/ell/common_ell.h:222-229 */
int ias_getinfo2(int rows, int cols, int num_diagonals, double *f3);
int ias_getinfo3(int rows, long long nnz, int width, double *f1);
/* number of populated diagonals / max row length without building the format */
int ias_count_diagonals(const IasCsrMatrixDev *A, int *num_diagonals);
int ias_max_row_nnz(const IasCsrMatrixDev *A, int *width);
/* the 26-feature vector in the order of CPU/main.cpp:655-679 */
int ias_features26(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, double *f26);
/* sizeofcsr / sizeofdia / sizeofell / sizeofcoo (csr:196-202, dia:20-26, ell:21-27, coo:20-26) */
double ias_sizeof_csr(int rows, long long nnz);
double ias_sizeof_dia(int rows, int cols, int num_diagonals);
double ias_sizeof_ell(int rows, int width);
double ias_sizeof_coo(int rows, long long nnz);

/* ---------------------------------------------------------------- transpose (A * A^T mode of GPU/main.cu:261-269) */
/* B := A^T on device (the reference calls mkl_dcsrcsc on the host); the result is canonical */
int ias_csr_transpose(const IasCsrMatrixDev *A, IasCsrMatrixDev *At);

/* ---------------------------------------------------------------- Matrix-Market front end */
/* loader of CPU/main.cpp:143-458 (banner: CPU/mmio.h:254-337): returns 0 or the reference's
 * exit codes -1 (open) -2 (banner) -3 (complex) -4 (size line); arrays are malloc'd. */
int ias_mtx_load(const char *path, IasCsrMatrix *out);
void ias_free_host_csr(IasCsrMatrix *m);
/* writes a result as "coordinate real general" (mm_write_mtx_crd, CPU/mmio.h:445-486, which the reference
 * never calls); row_base shifts the row indices of a row block */
int ias_mtx_write_csr64(const char *path, const IasCsr64Dev *C, int row_base);

/* ---------------------------------------------------------------- MatNet format selector (host) */
/* MatNet.Pred, CPU/MatNet.py:24-96 (called through embedded CPython at CPU/main.cpp:682-704, GPU/main.cu:446-460):
 * weights from a Keras 2.1 HDF5 file (./NetWeights/{Intel,Amd,P100}_weights.h5), two 128x128 density images and the
 * feature vector (26 values for the CPU nets, 18 for the GPU net) -> class index (CPU: 0 MKL 1 CSR 2 DIA 3 ELL
 * 4 COO; GPU: 0 CUSP 1 cuSPARSE 2 NSPARSE). */
int ias_matnet_load(const char *h5_path, void **net);
int ias_matnet_create(void **net);                /* empty net, to be filled with ias_matnet_set_tensor */
int ias_matnet_set_tensor(void *net, const char *name, const float *data, const long long *dims, int rank);
int ias_matnet_get_tensor(void *net, const char *name, float *out, long long *dims, int *rank);
int ias_matnet_shape(void *net, int *n_features, int *n_classes, long long *n_params);
int ias_matnet_predict(void *net, const long long *img1_16384, const long long *img2_16384, const double *features,
                       int *cls, double *probs);
void ias_matnet_free(void *net);

/* ---------------------------------------------------------------- synthetic operands (device) */
/* BASELINE.json configs, bit-identical to ia_spgemm_b200/workloads.py */
int ias_gen_poisson2d(int nx, int ny, IasCsrMatrixDev *out);   /* nx*ny nodes, row-major node order */
int ias_gen_uniform(int n, int per_row, int seed, IasCsrMatrixDev *out);
int ias_gen_rmat(int scale, int edge_factor, int seed, double a, double b, double c,
                 IasCsrMatrixDev *out);

#ifdef __cplusplus
}
#endif
#endif /* IASPGEMM_H */
