// features.cu -- density representation and structure features on device.
//
// Replaces the serial host loops of the reference front end:
//   density representation   CPU/main.cpp:516-577 (GPU/main.cu:279-413)
//   GetInfo1                 CPU/detail/csr/common_csr.h:257-287
//   GetInfo2 / GetInfo3      CPU/detail/dia/common_dia.h:222-233, CPU/detail/ell/common_ell.h:222-229
// The density histogram is privatised per CTA in shared memory (128x128 u32 bins = 64 KB) and
// flushed with 64-bit atomics; index arithmetic is 64-bit (the reference's `i*128` overflows int32
// for rows > 2^24).
#include <math.h>

#include <cub/cub.cuh>

#include "common.cuh"

using namespace ias;

namespace {

__device__ __forceinline__ void density_span(long long idx, long long dim, int &s, int &e)
{
    if (dim > 128)      { s = (int)(idx * 128 / dim); e = s; }
    else if (dim < 128) { s = (int)(idx * 128 / dim); e = s + (int)(128 / dim); }
    else                { s = (int)idx; e = (int)idx; }
}

__global__ void __launch_bounds__(512) k_density(int rows, int cols, const int *__restrict__ rp, const int *__restrict__ ci,
                                                 unsigned long long *__restrict__ img)
{
    extern __shared__ unsigned s_img[];          // 16384 bins
    for (int t = threadIdx.x; t < 16384; t += blockDim.x) s_img[t] = 0;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < rows; i += nwarps) {
        int rs, re;
        density_span(i, rows, rs, re);
        int pe = rp[i + 1];
        for (int p = rp[i] + lane; p < pe; p += 32) {
            int cs, ce;
            density_span(ci[p], cols, cs, ce);
            for (int r = rs; r <= re && r < 128; ++r)
                for (int c = cs; c <= ce && c < 128; ++c) atomicAdd(&s_img[r * 128 + c], 1u);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 16384; t += blockDim.x)
        if (s_img[t]) atomicAdd(&img[t], (unsigned long long)s_img[t]);
}

struct RowStats { int mx, mn; double ss; };

// max / min row length and sum of squared deviations from the mean (mean = nnz/rows is known up front)
__global__ void __launch_bounds__(256) k_row_stats(int rows, const int *__restrict__ rp, double mean,
                                                   int *__restrict__ g_max, int *__restrict__ g_min, double *__restrict__ partial)
{
    typedef cub::BlockReduce<double, 256> RD;
    typedef cub::BlockReduce<int, 256> RI;
    __shared__ typename RD::TempStorage td;
    __shared__ typename RI::TempStorage ti;
    int mx = -1, mn = 0x7fffffff;
    double ss = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        int len = rp[i + 1] - rp[i];
        mx = max(mx, len); mn = min(mn, len);
        double d = len - mean;
        ss += d * d;
    }
    ss = RD(td).Sum(ss);
    mx = RI(ti).Reduce(mx, cub::Max());
    __syncthreads();
    mn = RI(ti).Reduce(mn, cub::Min());
    if (threadIdx.x == 0) { partial[blockIdx.x] = ss; atomicMax(g_max, mx); atomicMin(g_min, mn); }
}

}  // namespace

extern "C" {

int ias_density_image(const IasCsrMatrixDev *A, long long *img_host)
{
    IAS_TRY(ensure_init());
    if (!A || !img_host) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    DBuf<unsigned long long> img;
    IAS_TRY(img.alloc(16384));
    IAS_CUDA(cudaMemsetAsync(img.p, 0, 16384 * sizeof(unsigned long long), c.stream));
    if (A->row > 0 && A->nnz > 0) {
        size_t sm = 16384 * sizeof(unsigned);
        IAS_CUDA(cudaFuncSetAttribute(k_density, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        long long warps = A->row;
        unsigned grid = (unsigned)std::min<long long>((warps * 32 + 511) / 512, (long long)c.sm_count * 2);
        IAS_LAUNCH(k_density, grid, 512, sm, A->row, A->col, A->row_ind_dev, A->col_ind_dev, img.p);
    }
    IAS_CUDA(cudaMemcpyAsync(img_host, img.p, 16384 * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

int ias_getinfo1(const IasCsrMatrixDev *A, double *f)
{
    IAS_TRY(ensure_init());
    if (!A || !f) return fail(IAS_E_ARG, "NULL");
    if (A->row <= 0) return fail(IAS_E_ARG, "GetInfo1 needs at least one row");
    Ctx &c = ctx();
    double mean = (double)A->nnz / A->row;
    const int GRID = 1024;
    DBuf<double> partial;
    DBuf<int> mm;
    IAS_TRY(partial.alloc(GRID));
    IAS_TRY(mm.alloc(2));
    int init[2] = {-1, 0x7fffffff};
    IAS_CUDA(cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, c.stream));
    int grid = (int)std::min<long long>(GRID, grid_for(A->row, 256));
    IAS_LAUNCH(k_row_stats, grid, 256, 0, A->row, A->row_ind_dev, mean, mm.p, mm.p + 1, partial.p);
    double h_partial[GRID];
    int h_mm[2];
    IAS_CUDA(cudaMemcpyAsync(h_partial, partial.p, sizeof(double) * grid, cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaMemcpyAsync(h_mm, mm.p, sizeof h_mm, cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    double ss = 0.0;
    for (int b = 0; b < grid; ++b) ss += h_partial[b];          // fixed order: deterministic
    double var = ss / (A->row - 1);
    f[0] = A->row; f[1] = A->col; f[2] = A->nnz;
    f[3] = (double)A->nnz / ((double)A->row * (double)A->col);   // the reference evaluates row*col in int (overflows)
    f[4] = h_mm[0]; f[5] = h_mm[1]; f[6] = mean; f[7] = var; f[8] = sqrt(var) / mean;
    return IAS_OK;
}

int ias_getinfo2(int rows, int cols, int nd, double *f)
{
    if (!f) return fail(IAS_E_ARG, "NULL");
    f[0] = nd;
    f[1] = (double)nd / (double)((long long)rows + cols - 1);
    f[2] = ((double)nd * (double)rows) / ((double)rows * (double)cols);
    return IAS_OK;
}

int ias_getinfo3(int rows, long long nnz, int width, double *f)
{
    if (!f) return fail(IAS_E_ARG, "NULL");
    f[0] = (double)nnz / ((double)rows * (double)width);
    return IAS_OK;
}

// CPU/main.cpp:655-679: GetInfo1(A), GetInfo1(B), GetInfo2(A_dia), GetInfo2(B_dia), GetInfo3(A_ell), GetInfo3(B_ell).
// GetInfo2/3 only need the diagonal count and the ELL width, so no format is built here.
int ias_features26(const IasCsrMatrixDev *A, const IasCsrMatrixDev *B, double *f)
{
    if (!A || !B || !f) return fail(IAS_E_ARG, "NULL");
    IAS_TRY(ias_getinfo1(A, f));
    IAS_TRY(ias_getinfo1(B, f + 9));
    const IasCsrMatrixDev *M[2] = {A, B};
    for (int t = 0; t < 2; ++t) {
        int nd = 0, w = 0;
        IAS_TRY(ias_count_diagonals(M[t], &nd));
        IAS_TRY(ias_getinfo2(M[t]->row, M[t]->col, nd, f + 18 + 3 * t));
        IAS_TRY(ias_max_row_nnz(M[t], &w));
        IAS_TRY(ias_getinfo3(M[t]->row, M[t]->nnz, w, f + 24 + t));
    }
    return IAS_OK;
}

}  // extern "C"
