// formats.cu -- ELL and COO: converters and multiplies (Algorithms 4 and 5).
//
// ELL replaces CSRtoELL (CPU/detail/ell/common_ell.h:30-77), ELL_MUL_ELL (ell:80-189) and the
// never-called ELL_MUL_ELL_DEV chain (GPU/detail/ell_dev/common_ell_dev.h:170-382: expand every
// product into a width-`max_upper` array, O(w^2) in-row dedup, serial <<<1,1>>> scans).  The
// multiply is the same Gustavson pipeline as CSR run on fixed-width rows (EllView: no row-pointer
// gathers, row j of B starts at j*w), writing a row-major ELL result of width max nnz(C_i) with
// column-sorted rows and 0 / 0.0 padding (the reference pads with 0, ell:54-56).
//
// COO replaces CSRtoCOO (CPU/detail/coo/common_coo.h:29-66), COO_MUL_COO (coo:72-161) and
// COO_MUL_COO_DEV (GPU/detail/coo_dev/common_coo_dev.h:279-602, CUSP's sliced ESC): the reference
// COO carries a CSR-like row_offset, so the multiply is the CSR pipeline on (row_offset, col, val)
// followed by a row-index expansion of the result.
#include <algorithm>

#include "spgemm_host.cuh"

using namespace ias;

namespace {

__global__ void __launch_bounds__(256) k_fill_ell(int rows, int w, const int *__restrict__ rp, const int *__restrict__ ci,
                                                  const double *__restrict__ v, int *__restrict__ nnz_row,
                                                  int *__restrict__ e_ci, double *__restrict__ e_v)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)rows * w;
    if (t >= n) return;
    int i = (int)(t / w), k = (int)(t % w);
    int b = rp[i], len = rp[i + 1] - b;
    if (k == 0) nnz_row[i] = len;
    e_ci[t] = k < len ? ci[b + k] : 0;
    e_v[t] = k < len ? v[b + k] : 0.0;
}

__global__ void __launch_bounds__(256) k_rows_only(int rows, const int *__restrict__ rp, int *__restrict__ nnz_row)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) nnz_row[i] = rp[i + 1] - rp[i];
}

__global__ void __launch_bounds__(256) k_widen(int n, const int *__restrict__ in, long long *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void __launch_bounds__(256) k_narrow(int n, const long long *__restrict__ in, int *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

// row index of every entry: each thread owns 16 consecutive entries (one binary search per thread)
__global__ void __launch_bounds__(256) k_expand_rows(int nrows, const long long *__restrict__ rp, int *__restrict__ row_ind)
{
    constexpr int PER = 16;
    long long n = rp[nrows];
    long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PER;
    if (e0 >= n) return;
    long long e1 = min(e0 + PER, n);
    int lo = 0, hi = nrows;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (rp[mid] <= e0) lo = mid; else hi = mid; }
    int row = lo;
    for (long long e = e0; e < e1; ++e) {
        while (rp[row + 1] <= e) ++row;
        row_ind[e] = row;
    }
}

int max_of_counts(const int *counts, int n, int *out)
{
    Ctx &c = ctx();
    *out = 0;
    if (n == 0) return IAS_OK;
    DBuf<int> d;
    IAS_TRY(d.alloc(1));
    size_t tb = 0;
    IAS_CUDA(cub::DeviceReduce::Max(nullptr, tb, counts, d.p, n, c.stream));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceReduce::Max(tmp.p, tb, counts, d.p, n, c.stream));
    c.launches += 2;
    IAS_CUDA(cudaMemcpyAsync(out, d.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- ELL
int ias_max_row_nnz(const IasCsrMatrixDev *A, int *width)
{
    IAS_TRY(ensure_init());
    if (!A || !width) return fail(IAS_E_ARG, "NULL");
    *width = 0;
    if (A->row == 0) return IAS_OK;
    DBuf<int> len;
    IAS_TRY(len.alloc(A->row));
    IAS_LAUNCH(k_rows_only, grid_for(A->row, 256), 256, 0, A->row, A->row_ind_dev, len.p);
    return max_of_counts(len.p, A->row, width);
}

int ias_csr_to_ell(const IasCsrMatrixDev *A, double gate, IasEllDev *out)
{
    IAS_TRY(ensure_init());
    if (!A || !out) return fail(IAS_E_ARG, "NULL");
    memset(out, 0, sizeof *out);
    out->row = A->row; out->col = A->col; out->nnz = A->nnz;
    int w = 0;
    IAS_TRY(ias_max_row_nnz(A, &w));
    out->max_nnz_per_row = w;
    // size gate (ell:47 uses 50x on the CPU, GPU/detail/ell/common_ell.h:46 uses 20x)
    if (!(ias_sizeof_ell(A->row, w) < gate * ias_sizeof_csr(A->row, A->nnz))) { out->choice = false; return IAS_OK; }
    out->choice = true;
    size_t cells = (size_t)A->row * w;
    DBuf<int> nr, ci;
    DBuf<double> v;
    IAS_TRY(nr.alloc((size_t)std::max(A->row, 1)));
    IAS_TRY(ci.alloc(cells));
    IAS_TRY(v.alloc(cells));
    if (cells) IAS_LAUNCH(k_fill_ell, grid_for((long long)cells, 256), 256, 0, A->row, w, A->row_ind_dev, A->col_ind_dev, A->values_dev, nr.p, ci.p, v.p);
    else if (A->row) IAS_CUDA(cudaMemsetAsync(nr.p, 0, sizeof(int) * A->row, ctx().stream));
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    out->nnz_row_dev = nr.release(); out->col_ind_dev = ci.release(); out->values_dev = v.release();
    return IAS_OK;
}

int ias_free_ell_dev(IasEllDev *m)
{
    if (!m) return IAS_OK;
    dfree(m->nnz_row_dev); dfree(m->col_ind_dev); dfree(m->values_dev);
    m->nnz_row_dev = nullptr; m->col_ind_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

static int ell_mul(const IasEllDev *A, const IasEllDev *B, IasEll64Dev *C, double *elapsed_ms)
{
    IAS_TRY(ensure_init());
    if (!A || !B || !C) return fail(IAS_E_ARG, "NULL");
    if (!A->choice || !B->choice) return fail(IAS_E_GATE, "ELL operand was rejected by the size gate (choice == false)");
    if (A->col > B->row) return fail(IAS_E_ARG, "shape mismatch: A is %dx%d, B is %dx%d", A->row, A->col, B->row, B->col);
    Ctx &c = ctx();
    memset(C, 0, sizeof *C);
    C->row = A->row; C->col = B->col; C->choice = true;
    EllView av{A->nnz_row_dev, A->col_ind_dev, A->values_dev, A->max_nnz_per_row};
    EllView bv{B->nnz_row_dev, B->col_ind_dev, B->values_dev, B->max_nnz_per_row};
    int nrows = A->row;
    IAS_CUDA(cudaEventRecord(c.ev[0], c.stream));
    IAS_CUDA(cudaEventRecord(c.ev[1], c.stream));
    RangeWork rw;
    double avg = A->row ? (double)A->nnz / A->row : 0.0;
    bool same = A->col_ind_dev == B->col_ind_dev && A->nnz_row_dev == B->nnz_row_dev;
    IAS_TRY(symbolic_range(av, bv, 0, nrows, B->col, avg, rw, nullptr, same, B->row));
    int w = 0;
    IAS_TRY(max_of_counts(rw.nnz_row.p, nrows, &w));          // C width = max nnz(C_i), ell:117-128
    // total nnz
    DBuf<long long> rp;
    IAS_TRY(rp.alloc((size_t)nrows + 1));
    IAS_TRY(scan_row_ptr(rw.nnz_row.p, nrows, rp.p));
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 32, rp.p + nrows, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    C->nnz = c.h_scalars[32];
    C->max_nnz_per_row = w;
    size_t cells = (size_t)nrows * w;
    DBuf<int> ci;
    DBuf<double> cv;
    IAS_TRY(ci.alloc(cells));
    IAS_TRY(cv.alloc(cells));
    if (cells) {
        IAS_CUDA(cudaMemsetAsync(ci.p, 0, sizeof(int) * cells, c.stream));      // padding = 0 / 0.0
        IAS_CUDA(cudaMemsetAsync(cv.p, 0, sizeof(double) * cells, c.stream));
    }
    OutMap out{nullptr, 0, rw.nnz_row.p, (long long)w};
    IAS_TRY(numeric_rows(av, bv, rw, 0, nrows, B->col, out, ci.p, cv.p, nullptr));
    IAS_CUDA(cudaEventRecord(c.ev[4], c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    if (elapsed_ms) *elapsed_ms = ev_ms(0, 4);
    C->nnz_row_dev = rw.nnz_row.release(); C->col_ind_dev = ci.release(); C->values_dev = cv.release();
    return IAS_OK;
}

int ias_ell_mul_ell_dev64(const IasEllDev *A, const IasEllDev *B, IasEll64Dev *C, double *elapsed_ms)
{
    return ell_mul(A, B, C, elapsed_ms);
}

int ias_free_ell64_dev(IasEll64Dev *m)
{
    if (!m) return IAS_OK;
    dfree(m->nnz_row_dev); dfree(m->col_ind_dev); dfree(m->values_dev);
    m->nnz_row_dev = nullptr; m->col_ind_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

int ias_ell_mul_ell_dev(const IasEllDev *A, const IasEllDev *B, IasEllDev *C, double *elapsed_ms)
{
    if (!C) return fail(IAS_E_ARG, "NULL");
    IasEll64Dev c64;
    IAS_TRY(ell_mul(A, B, &c64, elapsed_ms));
    if (c64.nnz > 0x7fffffffLL) {
        ias_free_ell64_dev(&c64);
        return fail(IAS_E_OVERFLOW, "nnz(C) = %lld does not fit the int32 EllMatrixDev layout; use ias_ell_mul_ell_dev64", c64.nnz);
    }
    C->choice = true; C->row = c64.row; C->col = c64.col; C->nnz = (int)c64.nnz; C->max_nnz_per_row = c64.max_nnz_per_row;
    C->nnz_row_dev = c64.nnz_row_dev; C->col_ind_dev = c64.col_ind_dev; C->values_dev = c64.values_dev;
    return IAS_OK;
}

int ias_download_ell(const IasEllDev *d, int *nnz_row, int *col_ind, double *values)
{
    IAS_TRY(ensure_init());
    if (!d) return fail(IAS_E_ARG, "NULL");
    if (!d->choice) return fail(IAS_E_GATE, "ELL matrix was rejected by the size gate");
    cudaStream_t s = ctx().stream;
    size_t cells = (size_t)d->row * d->max_nnz_per_row;
    if (nnz_row && d->row) IAS_CUDA(cudaMemcpyAsync(nnz_row, d->nnz_row_dev, sizeof(int) * d->row, cudaMemcpyDeviceToHost, s));
    if (col_ind && cells) IAS_CUDA(cudaMemcpyAsync(col_ind, d->col_ind_dev, sizeof(int) * cells, cudaMemcpyDeviceToHost, s));
    if (values && cells) IAS_CUDA(cudaMemcpyAsync(values, d->values_dev, sizeof(double) * cells, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    return IAS_OK;
}

// ---------------------------------------------------------------- COO
int ias_csr_to_coo(const IasCsrMatrixDev *A, IasCooDev *out)
{
    IAS_TRY(ensure_init());
    if (!A || !out) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    memset(out, 0, sizeof *out);
    out->choice = true; out->row = A->row; out->col = A->col; out->nnz = A->nnz;
    DBuf<long long> ro64;
    DBuf<int> ro, ri, ci;
    DBuf<double> v;
    IAS_TRY(ro.alloc((size_t)A->row + 1));
    IAS_TRY(ro64.alloc((size_t)A->row + 1));
    IAS_TRY(ri.alloc((size_t)A->nnz));
    IAS_TRY(ci.alloc((size_t)A->nnz));
    IAS_TRY(v.alloc((size_t)A->nnz));
    IAS_CUDA(cudaMemcpyAsync(ro.p, A->row_ind_dev, sizeof(int) * ((size_t)A->row + 1), cudaMemcpyDeviceToDevice, c.stream));
    if (A->nnz) {
        IAS_LAUNCH(k_widen, grid_for(A->row + 1, 256), 256, 0, A->row + 1, A->row_ind_dev, ro64.p);
        IAS_LAUNCH(k_expand_rows, grid_for(((long long)A->nnz + 15) / 16, 256), 256, 0, A->row, ro64.p, ri.p);
        IAS_CUDA(cudaMemcpyAsync(ci.p, A->col_ind_dev, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToDevice, c.stream));
        IAS_CUDA(cudaMemcpyAsync(v.p, A->values_dev, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToDevice, c.stream));
    }
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    out->row_offset_dev = ro.release(); out->row_ind_dev = ri.release(); out->col_ind_dev = ci.release(); out->values_dev = v.release();
    return IAS_OK;
}

int ias_free_coo_dev(IasCooDev *m)
{
    if (!m) return IAS_OK;
    dfree(m->row_offset_dev); dfree(m->row_ind_dev); dfree(m->col_ind_dev); dfree(m->values_dev);
    m->row_offset_dev = nullptr; m->row_ind_dev = nullptr; m->col_ind_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

int ias_free_coo64_dev(IasCoo64Dev *m)
{
    if (!m) return IAS_OK;
    dfree(m->row_offset_dev); dfree(m->row_ind_dev); dfree(m->col_ind_dev); dfree(m->values_dev);
    m->row_offset_dev = nullptr; m->row_ind_dev = nullptr; m->col_ind_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

int ias_coo_mul_coo_dev64(const IasCooDev *A, const IasCooDev *B, IasCoo64Dev *C, double *elapsed_ms)
{
    IAS_TRY(ensure_init());
    if (!A || !B || !C) return fail(IAS_E_ARG, "NULL");
    if (A->col > B->row) return fail(IAS_E_ARG, "shape mismatch: A is %dx%d, B is %dx%d", A->row, A->col, B->row, B->col);
    Ctx &c = ctx();
    memset(C, 0, sizeof *C);
    // the reference COO carries a CSR-like row_offset (coo:29-66): the multiply is the CSR pipeline on it
    CsrView av{A->row_offset_dev, A->col_ind_dev, A->values_dev};
    CsrView bv{B->row_offset_dev, B->col_ind_dev, B->values_dev};
    IasCsr64Dev c64;
    IasSpgemmStats st;
    double avg = A->row ? (double)A->nnz / A->row : 0.0;
    bool same = A->col_ind_dev == B->col_ind_dev && A->row_offset_dev == B->row_offset_dev;
    IAS_TRY(spgemm_materialise(av, bv, avg, B->col, 0, A->row, &c64, &st, same, B->row));
    DBuf<int> ri;
    IAS_TRY(ri.alloc((size_t)c64.nnz));
    IAS_CUDA(cudaEventRecord(c.ev[5], c.stream));
    if (c64.nnz) IAS_LAUNCH(k_expand_rows, grid_for((c64.nnz + 15) / 16, 256), 256, 0, c64.row, c64.row_ptr_dev, ri.p);
    IAS_CUDA(cudaEventRecord(c.ev[6], c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    C->choice = true; C->row = c64.row; C->col = c64.col; C->nnz = c64.nnz;
    C->row_offset_dev = c64.row_ptr_dev; C->row_ind_dev = ri.release(); C->col_ind_dev = c64.col_ind_dev; C->values_dev = c64.values_dev;
    if (elapsed_ms) *elapsed_ms = st.ms_total + ev_ms(5, 6);
    return IAS_OK;
}

int ias_coo_mul_coo_dev(const IasCooDev *A, const IasCooDev *B, IasCooDev *C, double *elapsed_ms)
{
    if (!C) return fail(IAS_E_ARG, "NULL");
    IasCoo64Dev c64;
    IAS_TRY(ias_coo_mul_coo_dev64(A, B, &c64, elapsed_ms));
    if (c64.nnz > 0x7fffffffLL) {
        ias_free_coo64_dev(&c64);
        return fail(IAS_E_OVERFLOW, "nnz(C) = %lld does not fit the int32 CooMatrixDev layout; use ias_coo_mul_coo_dev64", c64.nnz);
    }
    memset(C, 0, sizeof *C);
    DBuf<int> ro;
    IAS_TRY(ro.alloc((size_t)c64.row + 1));
    IAS_LAUNCH(k_narrow, grid_for(c64.row + 1, 256), 256, 0, c64.row + 1, c64.row_offset_dev, ro.p);
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    dfree(c64.row_offset_dev);
    C->choice = true; C->row = c64.row; C->col = c64.col; C->nnz = (int)c64.nnz;
    C->row_offset_dev = ro.release(); C->row_ind_dev = c64.row_ind_dev; C->col_ind_dev = c64.col_ind_dev; C->values_dev = c64.values_dev;
    return IAS_OK;
}

int ias_download_coo(const IasCooDev *d, int *row_offset, int *row_ind, int *col_ind, double *values)
{
    IAS_TRY(ensure_init());
    if (!d) return fail(IAS_E_ARG, "NULL");
    cudaStream_t s = ctx().stream;
    size_t n = (size_t)d->nnz;
    if (row_offset) IAS_CUDA(cudaMemcpyAsync(row_offset, d->row_offset_dev, sizeof(int) * ((size_t)d->row + 1), cudaMemcpyDeviceToHost, s));
    if (row_ind && n) IAS_CUDA(cudaMemcpyAsync(row_ind, d->row_ind_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (col_ind && n) IAS_CUDA(cudaMemcpyAsync(col_ind, d->col_ind_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (values && n) IAS_CUDA(cudaMemcpyAsync(values, d->values_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    return IAS_OK;
}

int ias_download_coo64(const IasCoo64Dev *d, long long *row_offset, int *row_ind, int *col_ind, double *values)
{
    IAS_TRY(ensure_init());
    if (!d) return fail(IAS_E_ARG, "NULL");
    cudaStream_t s = ctx().stream;
    size_t n = (size_t)d->nnz;
    if (row_offset) IAS_CUDA(cudaMemcpyAsync(row_offset, d->row_offset_dev, sizeof(long long) * ((size_t)d->row + 1), cudaMemcpyDeviceToHost, s));
    if (row_ind && n) IAS_CUDA(cudaMemcpyAsync(row_ind, d->row_ind_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (col_ind && n) IAS_CUDA(cudaMemcpyAsync(col_ind, d->col_ind_dev, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (values && n) IAS_CUDA(cudaMemcpyAsync(values, d->values_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    return IAS_OK;
}

}  // extern "C"
