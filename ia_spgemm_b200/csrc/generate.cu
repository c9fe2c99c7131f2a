// generate.cu -- synthetic operands of the BASELINE.json configs, generated on device.
//
// Bit-identical to ia_spgemm_b200/workloads.py (same counter-based splitmix64 hash, same draw
// order), so the CPU oracle (fed by the NumPy generators) and the GPU engine (fed by these) see
// the same arrays; tests/test_generators_gpu.py checks that.  Canonical CSR: sorted, duplicate-free
// columns, int32 indices, fp64 values in [0.5, 1.5) except Poisson (4 / -1).
#include <cub/cub.cuh>

#include "common.cuh"

using namespace ias;

namespace {

__host__ __device__ __forceinline__ uint64_t mix64_hd(uint64_t x)
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
inline uint64_t seed_salt(long long seed) { return mix64_hd((uint64_t)seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull); }
__device__ __forceinline__ uint64_t key2(uint64_t salt, uint64_t i, uint64_t j) { return ((i << 32) | j) ^ salt; }
__device__ __forceinline__ double unit53(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// ---- Poisson 2-D 5-point on an nx x ny grid (row-major node order r = y*nx + x): closed-form row pointers
__device__ __forceinline__ long long poisson_rp(long long r, long long nx, long long ny)
{
    long long top = r < nx ? r : nx;                                  // rows before r on the first grid line
    long long bottom = r > nx * (ny - 1) ? r - nx * (ny - 1) : 0;     // ... on the last grid line
    long long left = (r + nx - 1) / nx;                               // ... with x == 0
    long long right = r / nx;                                         // ... with x == nx-1
    return 5 * r - top - bottom - left - right;
}

__global__ void __launch_bounds__(256) k_poisson(int nx, int ny, int *__restrict__ rp, int *__restrict__ ci, double *__restrict__ v)
{
    long long n = (long long)nx * ny;
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    long long p = poisson_rp(r, nx, ny);
    rp[r] = (int)p;
    if (r == n) return;
    long long x = r % nx, y = r / nx;
    if (y > 0)      { ci[p] = (int)(r - nx); v[p] = -1.0; ++p; }
    if (x > 0)      { ci[p] = (int)(r - 1); v[p] = -1.0; ++p; }
    ci[p] = (int)r; v[p] = 4.0; ++p;
    if (x < nx - 1) { ci[p] = (int)(r + 1); v[p] = -1.0; ++p; }
    if (y < ny - 1) { ci[p] = (int)(r + nx); v[p] = -1.0; ++p; }
}

// ---- uniform: per_row distinct sorted columns per row; a row with a collision is re-drawn whole
constexpr int UNIFORM_MAX = 64;
__global__ void __launch_bounds__(128) k_uniform(int n, int per_row, uint64_t salt_cols, uint64_t salt_vals,
                                                 int *__restrict__ rp, int *__restrict__ ci, double *__restrict__ v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    rp[i] = (int)(i * per_row);
    if (i == n) return;
    long long c[UNIFORM_MAX];
    for (int draw = 0;; ++draw) {
        for (int t = 0; t < per_row; ++t) {
            long long x = (long long)(mix64(key2(salt_cols, (uint64_t)i, (uint64_t)t + (uint64_t)draw * per_row)) % (uint64_t)n);
            int s = t - 1;                                  // insertion sort
            while (s >= 0 && c[s] > x) { c[s + 1] = c[s]; --s; }
            c[s + 1] = x;
        }
        bool dup = false;
        for (int t = 1; t < per_row; ++t) dup |= (c[t] == c[t - 1]);
        if (!dup) break;
    }
    for (int t = 0; t < per_row; ++t) {
        long long p = i * per_row + t;
        ci[p] = (int)c[t];
        v[p] = 0.5 + unit53(mix64(key2(salt_vals, (uint64_t)i, (uint64_t)c[t])));
    }
}

// ---- R-MAT
__global__ void __launch_bounds__(256) k_rmat_edges(long long m, int scale, uint64_t salt, double a, double ab, double abc,
                                                    uint64_t *__restrict__ keys)
{
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    uint64_t i = 0, j = 0;
    for (int lvl = 0; lvl < scale; ++lvl) {
        double u = unit53(mix64(key2(salt, (uint64_t)e, (uint64_t)lvl)));
        uint64_t ibit = u >= ab;
        uint64_t jbit = ((u >= a) && (u < ab)) || (u >= abc);
        i = (i << 1) | ibit;
        j = (j << 1) | jbit;
    }
    keys[e] = (i << 32) | j;
}

__global__ void __launch_bounds__(256) k_rmat_rows(int n, long long nnz, const uint64_t *__restrict__ keys, int *__restrict__ rp)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    uint64_t target = (uint64_t)r << 32;
    long long lo = 0, hi = nnz;                      // first key >= target
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
    rp[r] = (int)lo;
}

__global__ void __launch_bounds__(256) k_rmat_fill(long long nnz, const uint64_t *__restrict__ keys, uint64_t salt_vals,
                                                   int *__restrict__ ci, double *__restrict__ v)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    uint64_t k = keys[p];
    uint64_t i = k >> 32, j = k & 0xffffffffull;
    ci[p] = (int)j;
    v[p] = 0.5 + unit53(mix64(key2(salt_vals, i, j)));
}

}  // namespace

extern "C" {

int ias_gen_poisson2d(int nx, int ny, IasCsrMatrixDev *out)
{
    IAS_TRY(ensure_init());
    if (!out || nx < 1 || ny < 1) return fail(IAS_E_ARG, "bad argument");
    long long n = (long long)nx * ny, nnz = 5 * n - 2LL * nx - 2LL * ny;
    if (nnz >= 0x7fffffffLL || n >= 0x7fffffffLL) return fail(IAS_E_OVERFLOW, "Poisson grid %dx%d overflows int32 indices", nx, ny);
    memset(out, 0, sizeof *out);
    out->choice = true; out->row = (int)n; out->col = (int)n; out->nnz = (int)nnz;
    IAS_TRY(dalloc(&out->row_ind_dev, (size_t)n + 1));
    IAS_TRY(dalloc(&out->col_ind_dev, (size_t)nnz));
    IAS_TRY(dalloc(&out->values_dev, (size_t)nnz));
    IAS_LAUNCH(k_poisson, grid_for(n + 1, 256), 256, 0, nx, ny, out->row_ind_dev, out->col_ind_dev, out->values_dev);
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

int ias_gen_uniform(int n, int per_row, int seed, IasCsrMatrixDev *out)
{
    IAS_TRY(ensure_init());
    if (!out || n < 1 || per_row < 1 || per_row > UNIFORM_MAX || per_row > n) return fail(IAS_E_ARG, "bad argument (per_row <= %d)", UNIFORM_MAX);
    long long nnz = (long long)n * per_row;
    if (nnz >= 0x7fffffffLL) return fail(IAS_E_OVERFLOW, "uniform %d x %d overflows int32 indices", n, per_row);
    memset(out, 0, sizeof *out);
    out->choice = true; out->row = n; out->col = n; out->nnz = (int)nnz;
    IAS_TRY(dalloc(&out->row_ind_dev, (size_t)n + 1));
    IAS_TRY(dalloc(&out->col_ind_dev, (size_t)nnz));
    IAS_TRY(dalloc(&out->values_dev, (size_t)nnz));
    IAS_LAUNCH(k_uniform, grid_for((long long)n + 1, 128), 128, 0, n, per_row, seed_salt((long long)seed + 7919), seed_salt(seed),
               out->row_ind_dev, out->col_ind_dev, out->values_dev);
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

int ias_gen_rmat(int scale, int edge_factor, int seed, double a, double b, double c3, IasCsrMatrixDev *out)
{
    IAS_TRY(ensure_init());
    if (!out || scale < 1 || scale > 30 || edge_factor < 1) return fail(IAS_E_ARG, "bad argument");
    Ctx &c = ctx();
    int n = 1 << scale;
    long long m = (long long)edge_factor * n;
    double ab = a + b, abc = a + b + c3;
    memset(out, 0, sizeof *out);
    DBuf<uint64_t> keys, sorted, uniq;
    DBuf<long long> d_count;
    IAS_TRY(keys.alloc((size_t)m));
    IAS_TRY(sorted.alloc((size_t)m));
    IAS_LAUNCH(k_rmat_edges, grid_for(m, 256), 256, 0, m, scale, seed_salt((long long)seed + 104729), a, ab, abc, keys.p);
    {
        size_t tb = 0;
        IAS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, sorted.p, m, 0, 32 + scale, c.stream));
        DBuf<char> tmp;
        IAS_TRY(tmp.alloc(tb));
        IAS_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, sorted.p, m, 0, 32 + scale, c.stream));
    }
    IAS_TRY(d_count.alloc(1));
    uniq.p = keys.release();                      // reuse the unsorted buffer for the unique keys
    {
        size_t tb = 0;
        IAS_CUDA(cub::DeviceSelect::Unique(nullptr, tb, sorted.p, uniq.p, d_count.p, m, c.stream));
        DBuf<char> tmp;
        IAS_TRY(tmp.alloc(tb));
        IAS_CUDA(cub::DeviceSelect::Unique(tmp.p, tb, sorted.p, uniq.p, d_count.p, m, c.stream));
    }
    c.launches += 8;
    long long nnz = 0;
    IAS_CUDA(cudaMemcpyAsync(&nnz, d_count.p, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    if (nnz >= 0x7fffffffLL) return fail(IAS_E_OVERFLOW, "R-MAT scale %d overflows int32 indices", scale);
    out->choice = true; out->row = n; out->col = n; out->nnz = (int)nnz;
    IAS_TRY(dalloc(&out->row_ind_dev, (size_t)n + 1));
    IAS_TRY(dalloc(&out->col_ind_dev, (size_t)nnz));
    IAS_TRY(dalloc(&out->values_dev, (size_t)nnz));
    IAS_LAUNCH(k_rmat_rows, grid_for((long long)n + 1, 256), 256, 0, n, nnz, uniq.p, out->row_ind_dev);
    IAS_LAUNCH(k_rmat_fill, grid_for(nnz, 256), 256, 0, nnz, uniq.p, seed_salt(seed), out->col_ind_dev, out->values_dev);
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

}  // extern "C"
