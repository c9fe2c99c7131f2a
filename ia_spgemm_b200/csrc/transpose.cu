// transpose.cu -- device CSR transpose and Matrix-Market output of a result.
//
// The reference's GPU driver multiplies A by its transpose: B := A^T through mkl_dcsrcsc on the host
// (GPU/main.cu:261-269); its CPU header also carries an O(cols*nnz) Transpose_CSR that is never called
// (CPU/detail/csr/common_csr.h:52-82).  Here the transpose is one 64-bit radix sort of (column, row)
// keys with the values as payload -- rows of the result come out column sorted (canonical), which is what
// the merge kernels want from a B operand.
// ias_mtx_write_csr64 is the writer the reference never calls (mm_write_mtx_crd, CPU/mmio.h:445-486):
// coordinate real general, 1-based, rows in order, columns ascending.
#include <stdio.h>

#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"

using namespace ias;

namespace {

__global__ void __launch_bounds__(256) k_transpose_keys(int rows, const int *__restrict__ rp, const int *__restrict__ ci,
                                                        unsigned long long *__restrict__ keys)
{
    int lane = threadIdx.x & 31;
    int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= rows) return;
    int pe = rp[i + 1];
    for (int p = rp[i] + lane; p < pe; p += 32) keys[p] = ((unsigned long long)(unsigned)ci[p] << 32) | (unsigned)i;
}

__global__ void __launch_bounds__(256) k_transpose_rows(int t_rows, long long nnz, const unsigned long long *__restrict__ keys,
                                                        int *__restrict__ rp)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > t_rows) return;
    unsigned long long target = (unsigned long long)(unsigned)r << 32;
    long long lo = 0, hi = nnz;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
    rp[r] = (int)lo;
}

__global__ void __launch_bounds__(256) k_transpose_cols(long long nnz, const unsigned long long *__restrict__ keys, int *__restrict__ ci)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) ci[p] = (int)(keys[p] & 0xffffffffull);
}

}  // namespace

extern "C" {

int ias_csr_transpose(const IasCsrMatrixDev *A, IasCsrMatrixDev *At)
{
    IAS_TRY(ensure_init());
    if (!A || !At) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    memset(At, 0, sizeof *At);
    At->choice = true; At->row = A->col; At->col = A->row; At->nnz = A->nnz;
    IAS_TRY(dalloc(&At->row_ind_dev, (size_t)At->row + 1));
    IAS_TRY(dalloc(&At->col_ind_dev, (size_t)At->nnz));
    IAS_TRY(dalloc(&At->values_dev, (size_t)At->nnz));
    long long nnz = A->nnz;
    DBuf<unsigned long long> keys, sorted;
    IAS_TRY(keys.alloc((size_t)nnz));
    IAS_TRY(sorted.alloc((size_t)nnz));
    if (nnz) {
        IAS_LAUNCH(k_transpose_keys, grid_for((long long)A->row * 32, 256), 256, 0, A->row, A->row_ind_dev, A->col_ind_dev, keys.p);
        int bits = 32;
        while (bits < 63 && (1LL << (bits - 32)) < (long long)A->col) ++bits;
        size_t tb = 0;
        IAS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, sorted.p, A->values_dev, At->values_dev, nnz, 0, bits, c.stream));
        DBuf<char> tmp;
        IAS_TRY(tmp.alloc(tb));
        IAS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, sorted.p, A->values_dev, At->values_dev, nnz, 0, bits, c.stream));
        c.launches += 4;
        IAS_LAUNCH(k_transpose_cols, grid_for(nnz, 256), 256, 0, nnz, sorted.p, At->col_ind_dev);
    }
    IAS_LAUNCH(k_transpose_rows, grid_for((long long)At->row + 1, 256), 256, 0, At->row, nnz, sorted.p, At->row_ind_dev);
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

int ias_mtx_write_csr64(const char *path, const IasCsr64Dev *C, int row_base)
{
    IAS_TRY(ensure_init());
    if (!path || !C) return fail(IAS_E_ARG, "NULL");
    std::vector<long long> rp((size_t)C->row + 1);
    std::vector<int> ci((size_t)C->nnz);
    std::vector<double> v((size_t)C->nnz);
    IAS_TRY(ias_download_csr64(C, rp.data(), ci.data(), v.data()));
    FILE *f = fopen(path, "w");
    if (!f) return fail(IAS_E_IO, "cannot open %s for writing", path);
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %lld\n", row_base + C->row, C->col, C->nnz);
    for (int i = 0; i < C->row; ++i)
        for (long long p = rp[i]; p < rp[i + 1]; ++p) fprintf(f, "%d %d %.17g\n", row_base + i + 1, ci[p] + 1, v[p]);
    if (fclose(f) != 0) return fail(IAS_E_IO, "write to %s failed", path);
    return IAS_OK;
}

}  // extern "C"
