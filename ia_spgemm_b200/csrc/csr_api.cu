// csr_api.cu -- C ABI of the CSR hot path (Algorithm 2 and the library calls it replaces).
#include <stdlib.h>

#include <vector>

#define IAS_TU tu_csr
#include "spgemm_host.cuh"

using namespace ias;

namespace {

int check_operands(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int r0, int r1)
{
    if (!A || !B) return fail(IAS_E_ARG, "NULL operand");
    if (A->row < 0 || A->col < 0 || B->row < 0 || B->col < 0) return fail(IAS_E_ARG, "negative dimension");
    if (A->col > B->row) return fail(IAS_E_ARG, "shape mismatch: A is %dx%d, B is %dx%d", A->row, A->col, B->row, B->col);
    if (r0 < 0 || r1 > A->row || r0 > r1) return fail(IAS_E_ARG, "row range [%d,%d) outside A (%d rows)", r0, r1, A->row);
    return IAS_OK;
}

CsrView view(const IasCsrMatrixDev *M) { return CsrView{M->row_ind_dev, M->col_ind_dev, M->values_dev}; }

// rows [r0,r1) of C = A*B, materialised
int mul_rows(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int r0, int r1, IasCsr64Dev *C, IasSpgemmStats *st)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, r0, r1));
    if (!C) return fail(IAS_E_ARG, "NULL result");
    double avg = A->row ? (double)A->nnz / A->row : 0.0;
    bool same = A->row_ind_dev == B->row_ind_dev && A->col_ind_dev == B->col_ind_dev;
    return spgemm_materialise(view(A), view(B), avg, B->col, r0, r1, C, st, same, B->row, B->nnz);
}

__global__ void k_narrow_rp(int n, const long long *in, int *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

}  // namespace

// rows [r0, r1) of the view `av` (a CsrView of A, or a CsrRowsView = an arbitrary list of A's rows) times B, streamed
template <class AV>
static int stream_impl(const AV &av, double avg, bool same, const IasCsrMatrixDev *B, int r0, int r1, size_t budget_bytes,
                       int *row_nnz_dev, ias_stream_consumer consumer, void *user, IasSpgemmStats *st)
{
    Ctx &c = ctx();
    long long l0 = c.launches;
    IasSpgemmStats local;
    memset(&local, 0, sizeof local);
    int nrows = r1 - r0;

    IAS_CUDA(cudaEventRecord(c.ev[0], c.stream));
    IAS_CUDA(cudaEventRecord(c.ev[1], c.stream));
    RangeWork rw;
    CsrView bv = view(B);
    IAS_TRY(symbolic_range(av, bv, r0, r1, B->col, avg, rw, &local, same, B->row, B->nnz));
    IAS_CUDA(cudaEventRecord(c.ev[2], c.stream));
    DBuf<long long> rp;
    IAS_TRY(rp.alloc((size_t)nrows + 1));
    IAS_TRY(scan_row_ptr(rw.nnz_row.p, nrows, rp.p));
    if (row_nnz_dev && nrows) IAS_LAUNCH(k_copy_counts, grid_for(nrows, 256), 256, 0, nrows, rw.nnz_row.p, row_nnz_dev);
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 32, rp.p + nrows, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    long long nnz = c.h_scalars[32];

    // batch capacity in entries (12 B each): the budget, but never less than the largest row
    if (budget_bytes == 0) {
        budget_bytes = (size_t)(0.6 * (double)free_device_bytes());
    }
    long long cap = (long long)(budget_bytes / 12);
    cap = std::max<long long>(cap, (long long)B->col);       // nnz(C_i) <= cols: one row always fits
    cap = std::min<long long>(cap, std::max<long long>(nnz, 1));
    const int MAXB = 4096;
    DBuf<int> bounds, nb;
    IAS_TRY(bounds.alloc(MAXB + 1));
    IAS_TRY(nb.alloc(1));
    IAS_LAUNCH(k_batch_bounds, 1, 1, 0, nrows, rp.p, cap, MAXB, bounds.p, nb.p);
    int h_nb = 0;
    std::vector<int> h_bounds(MAXB + 1, 0);
    IAS_CUDA(cudaMemcpyAsync(&h_nb, nb.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaMemcpyAsync(h_bounds.data(), bounds.p, sizeof(int) * (MAXB + 1), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    if (h_nb < 0) return fail(IAS_E_NOMEM, "streaming budget of %zu bytes needs more than %d batches", budget_bytes, MAXB);

    DBuf<int> ci;
    DBuf<double> cv;
    IAS_TRY(ci.alloc((size_t)cap));
    IAS_TRY(cv.alloc((size_t)cap));
    DBuf<unsigned long long> d_hash;
    DBuf<double> d_sum;
    IAS_TRY(d_hash.alloc(1));
    IAS_TRY(d_sum.alloc(1));
    IAS_CUDA(cudaMemsetAsync(d_hash.p, 0, sizeof(unsigned long long), c.stream));
    IAS_CUDA(cudaMemsetAsync(d_sum.p, 0, sizeof(double), c.stream));
    IAS_CUDA(cudaEventRecord(c.ev[3], c.stream));

    for (int b = 0; b < h_nb; ++b) {
        int b0 = h_bounds[b], b1 = h_bounds[b + 1];
        IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 40, rp.p + b0, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
        IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 41, rp.p + b1, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
        IAS_CUDA(cudaStreamSynchronize(c.stream));
        collect_bin_times(rw);                   // previous batch (and the symbolic bins) are complete here
        long long e0 = c.h_scalars[40], e1 = c.h_scalars[41];
        if (e1 == e0) continue;
        OutMap out{rp.p, e0, nullptr, 0};
        IAS_TRY(numeric_rows(av, bv, rw, b0, b1, B->col, out, ci.p, cv.p, &local));
        long long threads = (e1 - e0 + 15) / 16;
        IAS_LAUNCH(k_consume, grid_for(threads, 256), 256, 0, b1 - b0, r0 + b0, rp.p + b0, e0, ci.p, cv.p, d_hash.p, d_sum.p);
        if (consumer) {
            IasStreamBatch sb;
            sb.row_begin = r0 + b0; sb.row_end = r0 + b1; sb.batch_index = b; sb.batch_count = h_nb;
            sb.nnz_total = nnz; sb.entry_base = e0; sb.batch_nnz = e1 - e0;
            sb.row_ptr_dev = rp.p + b0; sb.col_ind_dev = ci.p; sb.values_dev = cv.p; sb.cuda_stream = (void *)c.stream;
            int crc = consumer(&sb, user);
            if (crc != 0) {
                cudaStreamSynchronize(c.stream);
                return fail(crc, "stream consumer returned %d at batch %d of %d (rows [%d,%d))", crc, b, h_nb, sb.row_begin, sb.row_end);
            }
        }
    }
    IAS_CUDA(cudaEventRecord(c.ev[4], c.stream));
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 42, d_hash.p, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 43, d_sum.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));

    collect_bin_times(rw);
    for (int b = 0; b < 8; ++b) { local.ms_bin_sym[b] = rw.ms_bin_sym[b]; local.ms_bin_num[b] = rw.ms_bin_num[b]; }
    local.nnz = nnz;
    local.batches = h_nb;
    local.structure_hash = (unsigned long long)c.h_scalars[42];
    memcpy(&local.checksum, c.h_scalars + 43, sizeof(double));
    local.ms_analyze = ias::ev_ms(0, 1); local.ms_symbolic = ias::ev_ms(1, 2); local.ms_scan = ias::ev_ms(2, 3);
    local.ms_numeric = ias::ev_ms(3, 4); local.ms_total = ias::ev_ms(0, 4);
    for (int b = 0; b < 8; ++b) local.num_bin_rows[b] = b < NBINS ? rw.num_hist[b] : 0;
    local.kernel_launches = (int)(c.launches - l0);
    if (st) *st = local;
    return IAS_OK;
}


extern "C" {

int ias_csr_mul_csr_rows_dev64(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int r0, int r1, IasCsr64Dev *C,
                               IasSpgemmStats *st)
{
    return mul_rows(A, B, r0, r1, C, st);
}

int ias_csr_mul_csr_dev64(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, IasCsr64Dev *C, IasSpgemmStats *st)
{
    if (!A) return fail(IAS_E_ARG, "NULL operand");
    return mul_rows(A, B, 0, A->row, C, st);
}

int ias_csr_mul_csr_dev(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, IasCsrMatrixDev *C, double *elapsed_ms)
{
    if (!A || !C) return fail(IAS_E_ARG, "NULL operand");
    IasCsr64Dev c64;
    IasSpgemmStats st;
    IAS_TRY(mul_rows(A, B, 0, A->row, &c64, &st));
    if (c64.nnz >= 0x7fffffffLL) {
        ias_free_csr64_dev(&c64);
        return fail(IAS_E_OVERFLOW, "nnz(C) = %lld does not fit the int32 CsrMatrixDev layout; use ias_csr_mul_csr_dev64", c64.nnz);
    }
    C->choice = true; C->row = c64.row; C->col = c64.col; C->nnz = (int)c64.nnz;
    IAS_TRY(dalloc(&C->row_ind_dev, (size_t)c64.row + 1));
    IAS_LAUNCH(k_narrow_rp, grid_for(c64.row + 1, 256), 256, 0, c64.row + 1, c64.row_ptr_dev, C->row_ind_dev);
    C->col_ind_dev = c64.col_ind_dev; C->values_dev = c64.values_dev;
    dfree(c64.row_ptr_dev);
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    if (elapsed_ms) *elapsed_ms = st.ms_total;
    return IAS_OK;
}

int ias_csr_mul_csr_stream(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int r0, int r1, size_t budget_bytes,
                           int *row_nnz_dev, IasSpgemmStats *st)
{
    return ias_csr_mul_csr_stream_cb(A, B, r0, r1, budget_bytes, row_nnz_dev, nullptr, nullptr, st);
}

int ias_csr_mul_csr_stream_cb(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int r0, int r1, size_t budget_bytes,
                              int *row_nnz_dev, ias_stream_consumer consumer, void *user, IasSpgemmStats *st)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, r0, r1));
    double avg = A->row ? (double)A->nnz / A->row : 0.0;
    bool same = A->row_ind_dev == B->row_ind_dev && A->col_ind_dev == B->col_ind_dev;
    return stream_impl(view(A), avg, same, B, r0, r1, budget_bytes, row_nnz_dev, consumer, user, st);
}

namespace {
__global__ void __launch_bounds__(256) k_gather_rows(int n, int a_rows, const int *__restrict__ rows, const int *__restrict__ rp,
                                                     int *__restrict__ rb, int *__restrict__ re, int *__restrict__ bad)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = rows[t];
    if (i < 0 || i >= a_rows) { *bad = 1; rb[t] = 0; re[t] = 0; return; }
    rb[t] = rp[i]; re[t] = rp[i + 1];
}
}  // namespace

int ias_csr_mul_csr_rowlist_stream(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, const int *rows_dev, int nrows,
                                   size_t budget_bytes, int *row_nnz_dev, ias_stream_consumer consumer, void *user,
                                   IasSpgemmStats *st)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, 0, A ? A->row : 0));
    if (nrows < 0 || (nrows > 0 && !rows_dev)) return fail(IAS_E_ARG, "bad row list");
    Ctx &c = ctx();
    DBuf<int> rb, re, bad;
    IAS_TRY(rb.alloc((size_t)std::max(nrows, 1)));
    IAS_TRY(re.alloc((size_t)std::max(nrows, 1)));
    IAS_TRY(bad.alloc(1));
    IAS_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), c.stream));
    if (nrows) IAS_LAUNCH(k_gather_rows, grid_for(nrows, 256), 256, 0, nrows, A->row, rows_dev, A->row_ind_dev, rb.p, re.p, bad.p);
    int h_bad = 0;
    IAS_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    if (h_bad) return fail(IAS_E_ARG, "row list names a row outside A (%d rows)", A->row);
    CsrRowsView av{rb.p, re.p, A->col_ind_dev, A->values_dev};
    double avg = A->row ? (double)A->nnz / A->row : 0.0;
    return stream_impl(av, avg, false, B, 0, nrows, budget_bytes, row_nnz_dev, consumer, user, st);
}

int ias_structure_hash(const IasCsr64Dev *C, int row_base, unsigned long long *hash)
{
    IAS_TRY(ensure_init());
    if (!C || !hash) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    *hash = 0;
    if (C->nnz == 0) return IAS_OK;
    DBuf<unsigned long long> d_hash;
    DBuf<double> d_sum;
    IAS_TRY(d_hash.alloc(1));
    IAS_TRY(d_sum.alloc(1));
    IAS_CUDA(cudaMemsetAsync(d_hash.p, 0, sizeof(unsigned long long), c.stream));
    IAS_CUDA(cudaMemsetAsync(d_sum.p, 0, sizeof(double), c.stream));
    long long threads = (C->nnz + 15) / 16;
    IAS_LAUNCH(k_consume, grid_for(threads, 256), 256, 0, C->row, row_base, C->row_ptr_dev, 0LL, C->col_ind_dev, C->values_dev, d_hash.p, d_sum.p);
    IAS_CUDA(cudaMemcpyAsync(hash, d_hash.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

// ---------------------------------------------------------------- host operands (the e2e path)
int ias_release_host(void)
{
    Ctx &c = ctx();
    if (c.h_arena) cudaFreeHost(c.h_arena);
    c.h_arena = nullptr; c.h_arena_bytes = 0;
    return IAS_OK;
}

int ias_csr_mul_csr_host(const IasCsrMatrix *A, const IasCsrMatrix *B, long long **c_rp, int **c_ci, double **c_v,
                         long long *c_nnz, IasSpgemmStats *st, double *ms_h2d, double *ms_d2h)
{
    IAS_TRY(ensure_init());
    if (!A || !B || !c_rp || !c_ci || !c_v || !c_nnz) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    bool alias = (A == B) || (A->row_ind == B->row_ind && A->col_ind == B->col_ind && A->values == B->values &&
                              A->row == B->row && A->col == B->col);
    IasCsrMatrixDev dA = {}, dB = {};
    IAS_CUDA(cudaEventRecord(c.ev[5], s));
    IAS_TRY(ias_upload_csr(A, &dA));
    if (alias) dB = dA; else IAS_TRY(ias_upload_csr(B, &dB));
    IAS_CUDA(cudaEventRecord(c.ev[6], s));
    IasCsr64Dev dC = {};
    int rc = mul_rows(&dA, &dB, 0, dA.row, &dC, st);
    if (rc == IAS_OK) {
        IAS_CUDA(cudaEventRecord(c.ev[7], s));
        size_t b_rp = sizeof(long long) * ((size_t)dC.row + 1), b_ci = sizeof(int) * (size_t)dC.nnz, b_v = sizeof(double) * (size_t)dC.nnz;
        size_t o_v = (b_rp + 255) / 256 * 256, o_ci = o_v + (b_v + 255) / 256 * 256;
        void *base = nullptr;
        rc = host_arena(o_ci + b_ci + 256, &base);
        if (rc == IAS_OK) {
            *c_rp = (long long *)base;
            *c_v = (double *)((char *)base + o_v);
            *c_ci = (int *)((char *)base + o_ci);
            rc = ias_download_csr64(&dC, *c_rp, *c_ci, *c_v);
            *c_nnz = dC.nnz;
        }
        cudaEventRecord(c.ev[0], s);           // mul_rows is done with ev[0]: reuse it as end of download
        cudaStreamSynchronize(s);
        if (ms_h2d) { float ms = 0; cudaEventElapsedTime(&ms, c.ev[5], c.ev[6]); *ms_h2d = ms; }
        if (ms_d2h) { float ms = 0; cudaEventElapsedTime(&ms, c.ev[7], c.ev[0]); *ms_d2h = ms; }
    }
    ias_free_csr64_dev(&dC);
    ias_free_csr_dev(&dA);
    if (!alias) ias_free_csr_dev(&dB);
    return rc;
}

// ---------------------------------------------------------------- GetFlop and work-balanced row blocks
namespace {
__global__ void __launch_bounds__(256) k_row_products(int nrows, CsrView A, CsrView B, long long *__restrict__ out)
{
    int lane = threadIdx.x & 31;
    int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nrows) return;
    long long ub = 0;
    int pe = A.end(i);
    for (int p = A.begin(i) + lane; p < pe; p += 32) ub += B.len(__ldg(A.ci + p));
    ub = warp_sum(ub);
    if (lane == 0) out[i] = ub;
}
__global__ void k_split(int nrows, const long long *__restrict__ incl, int parts, int *__restrict__ bounds)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > parts) return;
    if (p == 0) { bounds[0] = 0; return; }
    if (p == parts) { bounds[parts] = nrows; return; }
    long long total = nrows ? incl[nrows - 1] : 0;
    long long target = (long long)((double)total * p / parts);
    int lo = 0, hi = nrows;                     // first row whose inclusive prefix exceeds the target
    while (lo < hi) { int mid = (lo + hi) >> 1; if (incl[mid] <= target) lo = mid + 1; else hi = mid; }
    bounds[p] = lo;
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_mark_columns(long long n, const int *__restrict__ ci, unsigned *__restrict__ bitmap)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int j = ci[p];
    unsigned bit = 1u << (j & 31);
    if (!(bitmap[j >> 5] & bit)) atomicOr(bitmap + (j >> 5), bit);
}
__global__ void __launch_bounds__(256) k_touched_bytes(int b_rows, const unsigned *__restrict__ bitmap, const int *__restrict__ b_rp,
                                                       unsigned long long *__restrict__ out)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    long long mine = 0;
    if (j < b_rows && ((bitmap[j >> 5] >> (j & 31)) & 1u)) mine = 4 + 12LL * (b_rp[j + 1] - b_rp[j]);
    mine = warp_sum(mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, (unsigned long long)mine);
}
}  // namespace

int ias_touched_b_bytes(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int r0, int r1, long long *bytes)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, r0, r1));
    if (!bytes) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    *bytes = 0;
    int h_rp[2] = {0, 0};
    IAS_CUDA(cudaMemcpyAsync(&h_rp[0], A->row_ind_dev + r0, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaMemcpyAsync(&h_rp[1], A->row_ind_dev + r1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    long long n = (long long)h_rp[1] - h_rp[0];
    if (n <= 0 || B->row == 0) return IAS_OK;
    size_t words = ((size_t)B->row + 31) / 32;
    DBuf<unsigned> bitmap;
    DBuf<unsigned long long> out;
    IAS_TRY(bitmap.alloc(words));
    IAS_TRY(out.alloc(1));
    IAS_CUDA(cudaMemsetAsync(bitmap.p, 0, words * sizeof(unsigned), c.stream));
    IAS_CUDA(cudaMemsetAsync(out.p, 0, sizeof(unsigned long long), c.stream));
    IAS_LAUNCH(k_mark_columns, grid_for(n, 256), 256, 0, n, A->col_ind_dev + h_rp[0], bitmap.p);
    IAS_LAUNCH(k_touched_bytes, grid_for(B->row, 256), 256, 0, B->row, bitmap.p, B->row_ind_dev, out.p);
    IAS_CUDA(cudaMemcpyAsync(bytes, out.p, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

static int row_products_scan(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, DBuf<long long> &incl)
{
    IAS_TRY(incl.alloc((size_t)A->row + 1));
    if (A->row == 0) return IAS_OK;
    IAS_LAUNCH(k_row_products, grid_for((long long)A->row * 32, 256), 256, 0, A->row, view(A), view(B), incl.p);
    size_t tb = 0;
    IAS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, incl.p, incl.p, A->row, ctx().stream));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, incl.p, incl.p, A->row, ctx().stream));
    ctx().launches += 2;
    return IAS_OK;
}

namespace {
__global__ void __launch_bounds__(256) k_iota_ll(int n, int *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
// rows sorted by decreasing work are dealt in snake order (0 1 .. P-1 P-1 .. 1 0 0 1 ..): part `part` keeps its positions
__global__ void __launch_bounds__(256) k_take_share(int n, int parts, int part, const int *__restrict__ sorted_rows, int *__restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;              // k-th row of this part = round k
    int pos = (k & 1) ? k * parts + (parts - 1 - part) : k * parts + part;
    if (pos < n) out[k] = sorted_rows[pos];
}
}  // namespace

int ias_row_share(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int parts, int part, int *rows_dev, int *count)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, 0, A ? A->row : 0));
    if (parts < 1 || part < 0 || part >= parts || !rows_dev || !count) return fail(IAS_E_ARG, "bad share request");
    Ctx &c = ctx();
    const int n = A->row;
    *count = 0;
    if (n == 0) return IAS_OK;
    DBuf<long long> work, work_sorted;
    DBuf<int> ids, ids_sorted;
    IAS_TRY(work.alloc(n));
    IAS_TRY(work_sorted.alloc(n));
    IAS_TRY(ids.alloc(n));
    IAS_TRY(ids_sorted.alloc(n));
    IAS_LAUNCH(k_row_products, grid_for((long long)n * 32, 256), 256, 0, n, view(A), view(B), work.p);
    IAS_LAUNCH(k_iota_ll, grid_for(n, 256), 256, 0, n, ids.p);
    size_t tb = 0;
    IAS_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, work.p, work_sorted.p, ids.p, ids_sorted.p, n, 0, 48, c.stream));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tb, work.p, work_sorted.p, ids.p, ids_sorted.p, n, 0, 48, c.stream));
    c.launches += 2;
    int mine = 0;                                   // rounds k with a valid position
    for (int k = 0;; ++k) {
        long long pos = (k & 1) ? (long long)k * parts + (parts - 1 - part) : (long long)k * parts + part;
        if ((long long)k * parts >= n) break;
        if (pos < n) mine = k + 1;
    }
    // (a round whose position falls beyond n can only be the last one: the share is rounds 0 .. mine-1, all valid)
    if (mine) IAS_LAUNCH(k_take_share, grid_for(mine, 256), 256, 0, n, parts, part, ids_sorted.p, rows_dev);
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    *count = mine;
    return IAS_OK;
}

int ias_getflop(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, long long *products)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, 0, A ? A->row : 0));
    if (!products) return fail(IAS_E_ARG, "NULL");
    *products = 0;
    if (A->row == 0) return IAS_OK;
    DBuf<long long> incl;
    IAS_TRY(row_products_scan(A, B, incl));
    IAS_CUDA(cudaMemcpyAsync(products, incl.p + A->row - 1, sizeof(long long), cudaMemcpyDeviceToHost, ctx().stream));
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

int ias_partition_rows(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, int parts, int *bounds)
{
    IAS_TRY(ensure_init());
    IAS_TRY(check_operands(A, B, 0, A ? A->row : 0));
    if (parts < 1 || !bounds) return fail(IAS_E_ARG, "bad partition request");
    DBuf<long long> incl;
    IAS_TRY(row_products_scan(A, B, incl));
    DBuf<int> d_bounds;
    IAS_TRY(d_bounds.alloc((size_t)parts + 1));
    IAS_LAUNCH(k_split, grid_for(parts + 1, 64), 64, 0, A->row, incl.p, parts, d_bounds.p);
    IAS_CUDA(cudaMemcpyAsync(bounds, d_bounds.p, sizeof(int) * ((size_t)parts + 1), cudaMemcpyDeviceToHost, ctx().stream));
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

}  // extern "C"
