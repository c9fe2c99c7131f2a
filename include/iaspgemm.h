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
 *   - IasCsrMatrix / IasCsrMatrixDev / IasCooDev / IasDiaDev / IasEllDev have the field order, types and
 *     offsets of CsrMatrix / CsrMatrixDev / CooMatrixDev / DiaMatrixDev / EllMatrixDev in GPU/detail/format.h
 *     (tests/abi_cxx/layout_check.cpp static_asserts it against the reference header itself), so a reference
 *     `main` can pass its own objects; the one semantic difference is the DIA value order (see IasDiaDev);
 *     the *64 structs are the engine's own, for results beyond int32;
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

typedef struct {           /* CooMatrixDev, GPU/detail/format.h:29-40 -- byte-identical */
    bool choice;
    int row, col, nnz;
    int *row_offset_dev;        /* row+1 entries (the reference keeps a CSR-like row_offset) */
    int *row_ind_dev;
    int *col_ind_dev;
    double *values_dev;
} IasCooDev;

typedef struct {           /* COO result whose nnz does not fit int32 (no reference counterpart: it is `int` everywhere) */
    bool choice;
    int row, col;
    long long nnz;
    long long *row_offset_dev;
    int *row_ind_dev;
    int *col_ind_dev;
    double *values_dev;
} IasCoo64Dev;

typedef struct {           /* DiaMatrixDev, GPU/detail/format.h:82-92 -- same fields, same offsets */
    bool choice;
    int row, col, num_diagonals;
    int *diagonal_ind_dev;      /* row+col-1 entries: map index -> slot (0 for absent, as the reference) */
    int *diagonal_offsets_dev;  /* num_diagonals entries, ascending */
    double *values_dev;         /* The engine's kernels want DIAGONAL-MAJOR values[slot*row + i] (coalesced); the reference
                                   stores row-major values[i*num_diagonals + slot] (GPU/detail/dia/common_dia.h:70-90).
                                   ias_csr_to_dia / ias_dia_mul_dia_dev produce and consume diagonal-major;
                                   ias_dia_relayout converts a reference-built DiaMatrixDev in either direction and
                                   ias_download_dia returns the reference's row-major array. */
} IasDiaDev;

typedef struct {           /* EllMatrixDev, GPU/detail/format.h:108-119 -- byte-identical */
    bool choice;
    int row, col, nnz;
    int max_nnz_per_row;
    int *nnz_row_dev;
    int *col_ind_dev;           /* row-major [i*width + k] as the reference; padding 0 / 0.0 */
    double *values_dev;
} IasEllDev;

typedef struct {           /* ELL result whose nnz does not fit int32 */
    bool choice;
    int row, col;
    long long nnz;
    int max_nnz_per_row;
    int *nnz_row_dev;
    int *col_ind_dev;
    double *values_dev;
} IasEll64Dev;

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
int ias_set_stream(void *cuda_stream);          /* run on the caller's cudaStream_t; NULL is the legacy default stream */
int ias_use_own_stream(void);                   /* back to the engine's own (non-blocking) stream, the default after ias_init */
int ias_sync(void);
int ias_trim_pool(void);                        /* returns the engine pool's cached (free) device memory to the driver */
const char *ias_last_error(void);
const char *ias_version(void);
int ias_device_info(int *sm_count, size_t *smem_optin, size_t *free_bytes, size_t *total_bytes);
long long ias_kernel_launches(void);            /* engine kernels launched since ias_init */
/* Kernel-selection knobs (the reference has none: it hard-codes its library calls, GPU/main.cu:470-521).
 * Names: "global_rows_smem" (1 = windowed shared-memory kernels for rows beyond the CTA hash, 0 = L2 bitmap kernels),
 * "gwin_swords", "gwin_win", "gwin_sym_swords" (window sizes, 0 = automatic), "gwin_smem_kb", "gwin_max_sw" (numeric windowed
 * kernel only up to this many column super-windows per row, 0 = no limit), "g_win" (accumulate window of the L2 kernel), "g_coop", "gwin_takes_b2" (0/1 switches kept for A/B runs),
 * "trust_operand_cache" (see ias_forget_operand), "ell_onepass" (1 = one-pass ELL x ELL kernel where a row's products fit a
 * warp's register sort, 0 = always the pipeline), "g_block" (1024 / 512 threads per CTA of the L2 kernel), "g_ldca", "g_v2" (1 = second generation of the L2 kernel: rank + emit
 * from shared memory, split tables), "g_tbl" (its split-table capacity), "g_lpt" (1 = global rows in order of decreasing work), "g_scr" (per-CTA global scratch, in ints, for the split
 * tables of rows with more than 1024 A entries), "g2_takes_b2" (rows of the large CTA hash go to that kernel), "e2e_pipeline" (see ias_spgemm_auto_host), "bulk_store" (1 = cp.async.bulk copy-out of staged tiles),
 * "dia_vec" (1 = 128-bit DIA kernel), "g_split" (1 = global rows with very many products are cut into column-range parts, one CTA
 * each), "g_split_ub" (products from which a row is cut; 0 = automatic: a quarter of one CTA's even share of the launch, at least
 * 256 Ki), "g_split_parts" (parts per cut row, default 128), "block_cache" (1 = freed device blocks are kept per size class and reused
 * without a driver call; ias_trim_pool returns them).  Also read from
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
/* device scratch for the caller (DevMalloc, GPU/detail/common.h:62-77, without the memset and the exit(1)): a block of
 * the engine's pool on the engine stream, e.g. the row list ias_row_share fills; released with ias_device_free */
int ias_device_alloc(void **ptr, size_t bytes);
int ias_device_free(void *ptr);
/* With ias_set_option("trust_operand_cache", 1) the engine remembers, per B operand (pointers + shape), whether its
 * rows are canonical, so that repeated row-block multiplies against one B pay the 4 B/entry check once.  Off by
 * default: a caller that turns it on promises to call this before the next multiply whenever it rewrites an
 * operand's device arrays in place OR frees them (a caching allocator can hand the same address to a new operand of
 * the same shape).  ias_free_csr_dev does it for engine-owned operands; NULL forgets everything. */
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
/* Same with a consumer: after each batch's numeric kernels are enqueued the engine calls consumer(batch, user) on the
 * host.  The batch arrays are complete in stream order on batch->cuda_stream and are recycled when the callback
 * returns, so the consumer copies them out (ias_copy, or its own work enqueued on that stream and synchronised) or
 * reduces them before returning; a non-zero return aborts the multiply with that status.  consumer == NULL behaves
 * like ias_csr_mul_csr_stream.  This is how a streamed C leaves the device (download, .mtx append, hand-over to a
 * downstream operator) -- the role CUSP's workspace slices play in COO_MUL_COO_DEV
 * (GPU/detail/coo_dev/common_coo_dev.h:326-337,388-450). */
typedef struct {
    int row_begin, row_end;         /* absolute rows of C in this batch */
    int batch_index, batch_count;
    long long nnz_total;            /* nnz(C) of the whole requested row range (known once the symbolic pass is done) */
    long long entry_base;           /* offset of the batch's first entry within the range's C: entry e of row i sits at
                                       col_ind_dev[row_ptr_dev[i - row_begin] - entry_base + e] */
    long long batch_nnz;
    const long long *row_ptr_dev;   /* row_end - row_begin + 1 offsets into the range's C */
    const int *col_ind_dev;         /* batch_nnz entries, column-sorted inside each row */
    const double *values_dev;
    void *cuda_stream;
} IasStreamBatch;
typedef int (*ias_stream_consumer)(const IasStreamBatch *batch, void *user);
int ias_csr_mul_csr_stream_cb(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B,
                              int row_begin, int row_end, size_t budget_bytes, int *row_nnz_dev,
                              ias_stream_consumer consumer, void *user, IasSpgemmStats *stats);
/* The same for an arbitrary list of rows of A (device array of `nrows` row indices, any order, duplicates allowed): row k
 * of the streamed result is row rows_dev[k] of A*B.  This is the multi-GPU entry for skewed operands: rank r takes the
 * rows r, r + N, r + 2N, ... so that every rank gets its share of hub rows and of tail rows in ONE pass of the pipeline
 * (contiguous blocks concentrate the hubs; many small blocks pay a launch tail each).  Batch row numbers seen by the
 * consumer, the structure hash and row_nnz_dev are in list order (0 .. nrows-1). */
int ias_csr_mul_csr_rowlist_stream(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, const int *rows_dev, int nrows,
                                   size_t budget_bytes, int *row_nnz_dev, ias_stream_consumer consumer, void *user,
                                   IasSpgemmStats *stats);
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
/* work-balanced NON-contiguous share for ias_csr_mul_csr_rowlist_stream: the rows sorted by decreasing products are dealt
 * to the parts in snake order; rows_dev (capacity ceil(rows/parts)) receives the rows of `part`, *count their number.
 * (Un-permuted R-MAT: rows r, r+N, ... would not do -- the even rows hold three quarters of the entries.) */
int ias_row_share(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int parts, int part, int *rows_dev, int *count);
/* getsum_csr, csr_dev:264-273 (verified_sum) */
int ias_checksum(const double *values_dev, long long n, double *sum);
/* order-independent structure hash of a CSR result (same function as the streaming consumer) */
int ias_structure_hash(const IasCsr64Dev *C, int row_base, unsigned long long *hash);

/* ---------------------------------------------------------------- DIA (Algorithm 3) */
/* CSRtoDIA, CPU/detail/dia/common_dia.h:29-96 (gate: GPU/detail/dia/common_dia.h:51 uses 20x) */
int ias_csr_to_dia(const IasCsrMatrixDev *A, double gate, IasDiaDev *out);
/* DIA_MUL_DIA_DEV, GPU/detail/dia_dev/common_dia_dev.h:138-182 */
int ias_dia_mul_dia_dev(const IasDiaDev *A, const IasDiaDev *B, IasDiaDev *C, double *elapsed_ms);
/* rows [row_begin,row_end) of C only (multi-GPU row blocks): C->row = row_end-row_begin, values[slot*C->row + (i-row_begin)];
 * offsets are those of the whole product; diagonal_ind_dev is NULL unless the block is the whole matrix */
int ias_dia_mul_dia_rows_dev(const IasDiaDev *A, const IasDiaDev *B, int row_begin, int row_end, IasDiaDev *C,
                             double *elapsed_ms);
int ias_download_dia(const IasDiaDev *dev, int *diagonal_ind, int *diagonal_offsets,
                     double *values_row_major);
/* value order of a DIA matrix: to_row_major = 0 turns the reference's row-major values[i*nd + slot] (what
 * UploadDiaMatrix, GPU/detail/dia_dev/common_dia_dev.h:10-24, puts on the device) into the engine's diagonal-major
 * order; 1 goes back.  `out` gets its own arrays (release with ias_free_dia_dev). */
int ias_dia_relayout(const IasDiaDev *in, int to_row_major, IasDiaDev *out);
int ias_free_dia_dev(IasDiaDev *m);

/* ---------------------------------------------------------------- ELL (Algorithm 4) */
/* CSRtoELL, CPU/detail/ell/common_ell.h:30-77 */
int ias_csr_to_ell(const IasCsrMatrixDev *A, double gate, IasEllDev *out);
/* ELL_MUL_ELL_DEV, GPU/detail/ell_dev/common_ell_dev.h:310-382 */
int ias_ell_mul_ell_dev(const IasEllDev *A, const IasEllDev *B, IasEllDev *C, double *elapsed_ms);   /* IAS_E_OVERFLOW when nnz(C) >= 2^31 */
int ias_ell_mul_ell_dev64(const IasEllDev *A, const IasEllDev *B, IasEll64Dev *C, double *elapsed_ms);
int ias_download_ell(const IasEllDev *dev, int *nnz_row, int *col_ind, double *values);
int ias_download_ell64(const IasEll64Dev *dev, int *nnz_row, int *col_ind, double *values);
int ias_free_ell_dev(IasEllDev *m);
int ias_free_ell64_dev(IasEll64Dev *m);

/* ---------------------------------------------------------------- COO (Algorithm 5) */
/* CSRtoCOO, CPU/detail/coo/common_coo.h:29-66 */
int ias_csr_to_coo(const IasCsrMatrixDev *A, IasCooDev *out);
/* COO_MUL_COO_DEV, GPU/detail/coo_dev/common_coo_dev.h:279-602 */
int ias_coo_mul_coo_dev(const IasCooDev *A, const IasCooDev *B, IasCooDev *C, double *elapsed_ms);   /* IAS_E_OVERFLOW when nnz(C) >= 2^31 */
int ias_coo_mul_coo_dev64(const IasCooDev *A, const IasCooDev *B, IasCoo64Dev *C, double *elapsed_ms);
int ias_download_coo(const IasCooDev *dev, int *row_offset, int *row_ind, int *col_ind, double *values);
int ias_download_coo64(const IasCoo64Dev *dev, long long *row_offset, int *row_ind, int *col_ind, double *values);
int ias_free_coo_dev(IasCooDev *m);
int ias_free_coo64_dev(IasCoo64Dev *m);

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

/* ---------------------------------------------------------------- the front end's path in one call */
/* Fallback rule of the selector when no MatNet weights are at hand.  f = the 26 features; returns the class in the
 * CPU numbering of MatNet.Pred (CPU/MatNet.py:92; 1 = CSR, 2 = DIA, 3 = ELL). */
int ias_select_format(const double *features26, int dia_ok, int ell_ok);
/* features -> selection -> conversion -> multiply -> host result, i.e. CPU/main.cpp:655-935 with the selected
 * algorithm being the one that runs.  Operands are host CSR (as the loader leaves them); the result comes back in the
 * selected format's own layout, in pinned host memory owned by the engine (valid until the next host call or
 * ias_release_host):  format 1 CSR   row_ptr[row+1], col_ind[nnz], values[nnz]            (CSR_MUL_CSR, csr:85)
 *                     format 2 DIA   diagonal_ind[row+col-1], diagonal_offsets[num_diagonals],
 *                                    values[row][num_diagonals] row-major                  (DIA_mul_DIA, dia:101)
 *                     format 3 ELL   nnz_row[row], col_ind / values [row][max_nnz_per_row] (ELL_MUL_ELL, ell:80)
 * matnet: handle from ias_matnet_load (a 5-class net drives the dispatch) or NULL for the rule; gate: 20 (GPU release).
 * For A*A of a banded operand (B aliases A, rule selection) the call overlaps the PCIe upload of A's row chunks, the DIA
 * multiply of the blocks already there and the download of finished blocks of C: the diagonal set is taken from the
 * first chunk and verified against the whole operand afterwards, with a fall-back to the plain sequence
 * (option "e2e_pipeline" = 0 disables it). */
typedef struct {
    int format;
    int row, col;
    long long nnz;                 /* CSR: entries; DIA: row*num_diagonals cells; ELL: sum of the row lengths */
    long long *row_ptr;            /* CSR */
    int *col_ind;                  /* CSR, ELL */
    double *values;                /* all formats */
    int num_diagonals;             /* DIA */
    int *diagonal_ind, *diagonal_offsets;
    int max_nnz_per_row;           /* ELL */
    int *nnz_row;
    double features[26];
    double ms_h2d, ms_select, ms_convert, ms_multiply, ms_d2h;     /* CUDA events on the engine stream */
    long long h2d_bytes, d2h_bytes;
    double ms_wall;                /* host wall clock of the whole call */
    double ms_host[6];             /* host wall clock at: upload done, selection done, conversion done, multiply done, download done, buffers released */
    int pipelined;                 /* 1: upload, multiply and download were overlapped chunk by chunk (banded A^2; the ms_* phases are then 0) */
} IasAutoResult;
int ias_spgemm_auto_host(const IasCsrMatrix *A, const IasCsrMatrix *B, double gate, void *matnet, IasAutoResult *out);

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
