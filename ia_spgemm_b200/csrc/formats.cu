// formats.cu -- ELL and COO: converters and multiplies (Algorithms 4 and 5).
//
// ELL replaces CSRtoELL (CPU/detail/ell/common_ell.h:30-77), ELL_MUL_ELL (ell:80-189) and the
// never-called ELL_MUL_ELL_DEV chain (GPU/detail/ell_dev/common_ell_dev.h:170-382: expand every
// product into a width-`max_upper` array, O(w^2) in-row dedup, serial <<<1,1>>> scans).  The
// multiply is a one-pass kernel of its own when a row's products fit a warp's register sort (k_ell_mul_ell below)
// and otherwise the same Gustavson pipeline as CSR run on fixed-width rows (EllView: no row-pointer
// gathers, row j of B starts at j*w); either way the result is a row-major ELL of width max nnz(C_i) with
// column-sorted rows and 0 / 0.0 padding (the reference pads with 0, ell:54-56).
//
// COO replaces CSRtoCOO (CPU/detail/coo/common_coo.h:29-66), COO_MUL_COO (coo:72-161) and
// COO_MUL_COO_DEV (GPU/detail/coo_dev/common_coo_dev.h:279-602, CUSP's sliced ESC): the reference
// COO carries a CSR-like row_offset, so the multiply is the CSR pipeline on (row_offset, col, val)
// followed by a row-index expansion of the result.
#include <algorithm>

#define IAS_TU tu_formats
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

// ---------------------------------------------------------------- ELL x ELL in one pass (fixed-width operands)
// Replaces ELL_MUL_ELL_DEV (GPU/detail/ell_dev/common_ell_dev.h:310-382: upper bound, expand every product into a
// width-`max_upper` array, O(w^2) in-row dedup, serial scans, compaction).  What the fixed width buys:
//   * no row pointers anywhere: product (a, b) of row i is B[A.ci[i][a]][b], found by index arithmetic -- no scan, no
//     search, and every lane reads IPL consecutive entries of one B row (128-bit loads when the width allows);
//   * the output row starts at i * width: no prefix sum over rows, hence no separate symbolic pass -- one kernel
//     expands, sorts, compresses and writes, padding included (no memset of the result);
//   * the products arrive as P2(wa) sorted runs of P2(wb) slots (B rows are column sorted when B is canonical), so the
//     register sorting network starts at the first level that merges two runs: 26 compare-exchange stages instead of
//     36 for 16 x 16.
// A warp owns a row; element e = lane * IPL + r of its 32 * IPL slots is product (a, b) = (e / R, e % R), R = slots per
// run.  Key = column << IDX_BITS | e (ties keep the arrival order, the order CSR_MUL_CSR / ELL_MUL_ELL accumulate in);
// values wait in shared memory.  The network is the all-ascending form of the bitonic sorter: a merge level first
// compares e with its mirror image e ^ (kk - 1), then halves the distance.  After the sort: segmented sum over equal
// columns, the last product of a column writes the entry (same epilogue as k_esc_warp).
// C is written with the upper-bound width w_ub = min(wa * wb, cols); the host re-strides it when the largest row
// turns out narrower (ell:117-128 defines the width as max nnz(C_i)).
template <class KeyT, int IPL>
__device__ __forceinline__ void warp_sort_runs(KeyT (&key)[IPL], int lane, int run /* slots per sorted run, power of two */)
{
    constexpr int N = 32 * IPL;
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
        if (kk <= run) continue;                                   // warp-uniform: these levels are already in order
        // mirror step: e <-> e ^ (kk - 1)
        if (kk <= IPL) {
#pragma unroll
            for (int r = 0; r < IPL; ++r) {
                const int p = r ^ (kk - 1);
                if (r < p) { KeyT lo = key[r] < key[p] ? key[r] : key[p], hi = key[r] < key[p] ? key[p] : key[r]; key[r] = lo; key[p] = hi; }
            }
        } else {
            const int lm = kk / IPL - 1;                           // lanes e and e' differ in these bits; registers are mirrored
            const bool low = (lane & (kk / (2 * IPL))) == 0;
            KeyT other[IPL];
#pragma unroll
            for (int r = 0; r < IPL; ++r) other[r] = __shfl_xor_sync(0xffffffffu, key[IPL - 1 - r], lm);
#pragma unroll
            for (int r = 0; r < IPL; ++r) key[r] = low ? (key[r] < other[r] ? key[r] : other[r]) : (key[r] < other[r] ? other[r] : key[r]);
        }
#pragma unroll
        for (int jj = kk >> 2; jj > 0; jj >>= 1) {
            if (jj >= IPL) {
                const int lj = jj / IPL;
                const bool low = (lane & lj) == 0;
#pragma unroll
                for (int r = 0; r < IPL; ++r) {
                    const KeyT other = __shfl_xor_sync(0xffffffffu, key[r], lj);
                    key[r] = low ? (key[r] < other ? key[r] : other) : (key[r] < other ? other : key[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < IPL; ++r)
                    if ((r & jj) == 0) { KeyT lo = key[r] < key[r | jj] ? key[r] : key[r | jj], hi = key[r] < key[r | jj] ? key[r | jj] : key[r]; key[r] = lo; key[r | jj] = hi; }
            }
        }
    }
}

template <class KeyT, int IPL, bool FULL /* every row of A and B holds exactly `width` entries */, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ell_mul_ell(int nrows, EllView A, EllView B, int log2_run, int sorted_runs, int vec_ok,
                                                       int w_ub, int bulk, int *__restrict__ c_nr, int *__restrict__ c_ci, double *__restrict__ c_v,
                                                       unsigned long long *__restrict__ total_nnz, int *__restrict__ max_nnz)
{
    constexpr int N = 32 * IPL;
    constexpr int IDX_BITS = IPL == 16 ? 9 : IPL == 8 ? 8 : IPL == 4 ? 7 : 6;
    constexpr int WARPS = BLOCK / 32;
    // per warp: the products' values, later the finished row (values and columns) on its way out
    __shared__ __align__(16) double s_vals[WARPS][N];
    __shared__ __align__(16) int s_cols[WARPS][N];
    __shared__ unsigned long long s_total;
    __shared__ int s_max;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_total = 0; s_max = 0; }
    __syncthreads();
    const KeyT PAD = ~(KeyT)0;
    const KeyT IDX_MASK = ((KeyT)1 << IDX_BITS) - 1;
    double *svals = s_vals[w];
    int *scols = s_cols[w];
    const int e0 = lane * IPL;                       // first slot of this lane
    const int a = e0 >> log2_run;                    // its run = A entry (a run spans run / IPL >= 1 lanes)
    const int b0 = e0 & ((1 << log2_run) - 1);       // first B position of this lane inside the run
    long long my_total = 0;
    int my_max = 0;
    bool pending = false;                            // a bulk copy of the previous row may still be reading svals / scols
    const int nwarps = gridDim.x * WARPS;
    for (int i = blockIdx.x * WARPS + w; i < nrows; i += nwarps) {
        const int na = FULL ? A.w : __ldg(A.nr + i);
        int j = -1, lenb = 0;
        double av = 0.0;
        if (a < na) {
            j = __ldg(A.ci + (long long)i * A.w + a);
            av = __ldg(A.v + (long long)i * A.w + a);
            lenb = FULL ? B.w : __ldg(B.nr + j);
        }
        // this lane's IPL consecutive entries of B row j, into registers first: the loads are in flight while the
        // copy engine finishes reading the previous row out of shared memory
        int cc[IPL];
        double vv[IPL];
        const long long qb = (long long)j * B.w + b0;
        if (FULL && vec_ok && j >= 0 && b0 + IPL <= B.w) {
#pragma unroll
            for (int r = 0; r < IPL; r += 4) {       // 128-bit loads (IPL is a multiple of 4 here, the row start 16-byte aligned)
                const int4 c4 = __ldg(reinterpret_cast<const int4 *>(B.ci + qb + r));
                const double2 v01 = __ldg(reinterpret_cast<const double2 *>(B.v + qb + r));
                const double2 v23 = __ldg(reinterpret_cast<const double2 *>(B.v + qb + r + 2));
                cc[r] = c4.x; vv[r] = v01.x;
                if (r + 1 < IPL) { cc[r + 1] = c4.y; vv[r + 1] = v01.y; }
                if (r + 2 < IPL) { cc[r + 2] = c4.z; vv[r + 2] = v23.x; }
                if (r + 3 < IPL) { cc[r + 3] = c4.w; vv[r + 3] = v23.y; }
            }
        } else {
#pragma unroll
            for (int r = 0; r < IPL; ++r) {
                const bool valid = j >= 0 && b0 + r < lenb;
                cc[r] = valid ? __ldg(B.ci + qb + r) : -1;
                vv[r] = valid ? __ldg(B.v + qb + r) : 0.0;
            }
        }
        if (pending) {                               // warp-uniform
            if (lane == 0) bulk_wait_group_read0();
            __syncwarp();
            pending = false;
        }
        KeyT key[IPL];
#pragma unroll
        for (int r = 0; r < IPL; ++r) {
            key[r] = cc[r] >= 0 ? (((KeyT)(unsigned)cc[r] << IDX_BITS) | (KeyT)(e0 + r)) : PAD;
            svals[r * 32 + lane] = av * vv[r];         // slot e = lane * IPL + r lives at r * 32 + lane: conflict-free stores
                                                       // (lane-major [e] made every store a 16-way bank conflict: ncu 2.3 G conflicts)
        }
        __syncwarp();
        warp_sort_runs<KeyT, IPL>(key, lane, sorted_runs ? (1 << log2_run) : 1);
        // tails: last product of every column
        KeyT next_first = __shfl_down_sync(0xffffffffu, key[0], 1);
        if (lane == 31) next_first = PAD;
        int tails = 0;
        bool is_tail[IPL];
#pragma unroll
        for (int r = 0; r < IPL; ++r) {
            const KeyT nxt = r + 1 < IPL ? key[r + 1] : next_first;
            is_tail[r] = key[r] != PAD && (nxt == PAD || (nxt >> IDX_BITS) != (key[r] >> IDX_BITS));
            tails += is_tail[r];
        }
        int tincl = tails;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, tincl, o); if (lane >= o) tincl += u; }
        const int count = __shfl_sync(0xffffffffu, tincl, 31);
        // segmented sums (heads = first product of a column)
        const KeyT prev_last = __shfl_up_sync(0xffffffffu, key[IPL - 1], 1);
        double v[IPL];
        bool head[IPL];
        bool any_head = false;
        double tail_sum = 0.0;
#pragma unroll
        for (int r = 0; r < IPL; ++r) {
            const bool valid = key[r] != PAD;
            const KeyT prv = r > 0 ? key[r - 1] : prev_last;
            head[r] = valid && ((r == 0 && lane == 0) || (prv >> IDX_BITS) != (key[r] >> IDX_BITS));
            const unsigned e = (unsigned)(key[r] & IDX_MASK);
            v[r] = valid ? svals[(e % IPL) * 32 + e / IPL] : 0.0;
            tail_sum = head[r] ? v[r] : tail_sum + v[r];
            any_head |= head[r];
        }
        double carry = tail_sum;
        bool flag = any_head;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double uc = __shfl_up_sync(0xffffffffu, carry, o);
            const bool uf = __shfl_up_sync(0xffffffffu, flag, o);
            if (lane >= o && !flag) { carry += uc; flag = uf; }
        }
        double carry_in = __shfl_up_sync(0xffffffffu, carry, 1);
        if (lane == 0) carry_in = 0.0;
        __syncwarp();                                              // every product value is in a register: svals becomes the output row
        const long long gs = (long long)i * w_ub;
        int opos = tincl - tails;
        double run_sum = carry_in;
#pragma unroll
        for (int r = 0; r < IPL; ++r) {
            run_sum = head[r] ? v[r] : run_sum + v[r];
            if (is_tail[r]) { scols[opos] = (int)(key[r] >> IDX_BITS); svals[opos] = run_sum; ++opos; }
        }
        for (int p = count + lane; p < w_ub; p += 32) { scols[p] = 0; svals[p] = 0.0; }      // padding = 0 / 0.0 (ell:54-56)
        __syncwarp();
        if (bulk) {
            // the finished row (w_ub entries, padding included) leaves through the copy engine: two cp.async.bulk per
            // row instead of 2 * w_ub / 32 strided stores per lane, and it overlaps the next row's loads
            if (lane == 0) {
                fence_proxy_async_shared();
                bulk_store_shared_to_global(c_ci + gs, scols, (unsigned)w_ub * 4u);
                bulk_store_shared_to_global(c_v + gs, svals, (unsigned)w_ub * 8u);
                bulk_commit_group();
            }
            pending = true;
        } else {
            for (int p = lane; p < w_ub; p += 32) { c_ci[gs + p] = scols[p]; c_v[gs + p] = svals[p]; }
            __syncwarp();
        }
        if (lane == 0) { c_nr[i] = count; my_total += count; my_max = max(my_max, count); }
    }
    if (pending && lane == 0) bulk_wait_group_read0();             // shared memory must outlive the copy's reads
    if (lane == 0) { if (my_total) atomicAdd(&s_total, (unsigned long long)my_total); atomicMax(&s_max, my_max); }
    __syncthreads();
    if (threadIdx.x == 0) { if (s_total) atomicAdd(total_nnz, s_total); if (s_max) atomicMax(max_nnz, s_max); }
}

// C was written with stride w_in; the reference's width is max nnz(C_i) = w_out <= w_in
__global__ void __launch_bounds__(256) k_ell_restride(long long cells_out, int w_in, int w_out, const int *__restrict__ ci_in,
                                                      const double *__restrict__ v_in, int *__restrict__ ci_out, double *__restrict__ v_out)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells_out) return;
    long long i = t / w_out;
    int k = (int)(t - i * w_out);
    ci_out[t] = ci_in[i * w_in + k];
    v_out[t] = v_in[i * w_in + k];
}

__global__ void __launch_bounds__(256) k_min_count(int n, const int *__restrict__ counts, int *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = i < n ? counts[i] : 0x7fffffff;
    v = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v < *out) atomicMin(out, v);
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

static int p2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// one-pass ELL x ELL (k_ell_mul_ell) when the products of a row fit a warp's register sort: P2(wa) * P2(wb) <= 512
// slots, at most 32 runs, and the upper-bound-width result fits comfortably in memory.  *done = 0: not applicable.
static int ell_mul_onepass(const IasEllDev *A, const IasEllDev *B, IasEll64Dev *C, double *elapsed_ms, int *done)
{
    *done = 0;
    Ctx &c = ctx();
    if (c.tune.ell_onepass == 0) return IAS_OK;
    const int wa = A->max_nnz_per_row, wb = B->max_nnz_per_row;
    if (A->row <= 0 || wa <= 0 || wb <= 0) return IAS_OK;
    int runs = p2_ceil(wa), run = p2_ceil(wb);
    if (runs > 32 || (long long)runs * run > 512) return IAS_OK;
    while (runs * run < 64) run <<= 1;                       // at least two slots per lane
    const int n_slots = runs * run, ipl = n_slots / 32;
    if (run < ipl) return IAS_OK;                            // a lane's slots must lie inside one run
    int log2_run = 0;
    while ((1 << log2_run) < run) ++log2_run;
    const long long w_ub = std::min<long long>((long long)wa * wb, (long long)B->col);
    const size_t f = free_device_bytes();
    const double need = (double)A->row * (double)w_ub * 12.0;
    if (need > 0.45 * (double)f) return IAS_OK;              // the two-pass pipeline sizes C exactly
    int col_bits = 1;
    while (col_bits < 31 && (1LL << col_bits) < (long long)B->col) ++col_bits;
    const int idx_bits = ipl == 16 ? 9 : ipl == 8 ? 8 : ipl == 4 ? 7 : 6;
    const bool k32 = col_bits + idx_bits <= 32;              // the all-ones key is the padding marker: keep it unreachable
    if (k32 && col_bits + idx_bits == 32 && (long long)B->col == (1LL << col_bits)) return IAS_OK;

    cudaStream_t s = c.stream;
    IAS_CUDA(cudaEventRecord(c.ev[0], s));
    EllView av{A->nnz_row_dev, A->col_ind_dev, A->values_dev, wa};
    EllView bv{B->nnz_row_dev, B->col_ind_dev, B->values_dev, wb};
    // scalars: [0] total nnz (u64), [1] max nnz, min row length of A, of B, B not canonical
    DBuf<unsigned long long> d_sc;
    IAS_TRY(d_sc.alloc(8));
    IAS_CUDA(cudaMemsetAsync(d_sc.p, 0, 8 * sizeof(unsigned long long), s));
    int *d_int = reinterpret_cast<int *>(d_sc.p + 1);        // [0] max, [1] min A, [2] min B
    const int big = 0x7fffffff;
    IAS_CUDA(cudaMemcpyAsync(d_int + 1, &big, sizeof(int), cudaMemcpyHostToDevice, s));
    IAS_CUDA(cudaMemcpyAsync(d_int + 2, &big, sizeof(int), cudaMemcpyHostToDevice, s));
    IAS_LAUNCH(k_min_count, grid_for(A->row, 256), 256, 0, A->row, A->nnz_row_dev, d_int + 1);
    IAS_LAUNCH(k_min_count, grid_for(B->row, 256), 256, 0, B->row, B->nnz_row_dev, d_int + 2);
    IAS_LAUNCH((k_rows_canonical<EllView>), grid_for(B->row, 256), 256, 0, B->row, bv, d_sc.p + 4);
    unsigned long long h_sc[8];
    IAS_CUDA(cudaMemcpyAsync(h_sc, d_sc.p, sizeof h_sc, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    const int *h_int = reinterpret_cast<const int *>(h_sc + 1);
    const bool full = h_int[1] == wa && h_int[2] == wb;
    const int sorted_runs = h_sc[4] == 0 ? 1 : 0;
    const int vec_ok = (wb % 4 == 0 && ipl % 4 == 0) ? 1 : 0;
    const int bulk = (c.tune.bulk_store != 0 && w_ub % 4 == 0 && w_ub <= n_slots) ? 1 : 0;     // 16-byte aligned rows of C

    const size_t cells = (size_t)A->row * (size_t)w_ub;
    DBuf<int> nr, ci;
    DBuf<double> cv;
    IAS_TRY(nr.alloc((size_t)A->row));
    IAS_TRY(ci.alloc(cells));
    IAS_TRY(cv.alloc(cells));
    constexpr int EB = 128;
    const unsigned grid = (unsigned)std::min<long long>(grid_for(A->row, EB / 32), (long long)c.sm_count * 16);
#define IAS_ELL1(KT, IPL, FULLROWS)                                                                                          \
    IAS_LAUNCH((k_ell_mul_ell<KT, IPL, FULLROWS, EB>), grid, EB, 0, A->row, av, bv, log2_run, sorted_runs, vec_ok, (int)w_ub, \
               bulk, nr.p, ci.p, cv.p, d_sc.p, d_int)
#define IAS_ELL2(KT, FULLROWS)                                                                     \
    do {                                                                                           \
        if (ipl == 2) IAS_ELL1(KT, 2, FULLROWS); else if (ipl == 4) IAS_ELL1(KT, 4, FULLROWS);     \
        else if (ipl == 8) IAS_ELL1(KT, 8, FULLROWS); else IAS_ELL1(KT, 16, FULLROWS);             \
    } while (0)
    if (k32) { if (full) IAS_ELL2(unsigned, true); else IAS_ELL2(unsigned, false); }
    else     { if (full) IAS_ELL2(unsigned long long, true); else IAS_ELL2(unsigned long long, false); }
#undef IAS_ELL2
#undef IAS_ELL1
    IAS_CUDA(cudaMemcpyAsync(h_sc, d_sc.p, sizeof h_sc, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    const long long nnz = (long long)h_sc[0];
    const int w = h_int[0];
    memset(C, 0, sizeof *C);
    C->choice = true; C->row = A->row; C->col = B->col; C->nnz = nnz; C->max_nnz_per_row = w;
    if (w < w_ub) {                                          // the reference's width is max nnz(C_i): re-stride
        const size_t out_cells = (size_t)A->row * (size_t)w;
        DBuf<int> ci2;
        DBuf<double> cv2;
        IAS_TRY(ci2.alloc(out_cells));
        IAS_TRY(cv2.alloc(out_cells));
        if (out_cells) IAS_LAUNCH(k_ell_restride, grid_for((long long)out_cells, 256), 256, 0, (long long)out_cells, (int)w_ub, w, ci.p, cv.p, ci2.p, cv2.p);
        IAS_CUDA(cudaEventRecord(c.ev[4], s));
        IAS_CUDA(cudaStreamSynchronize(s));
        C->col_ind_dev = ci2.release(); C->values_dev = cv2.release();
    } else {
        IAS_CUDA(cudaEventRecord(c.ev[4], s));
        IAS_CUDA(cudaStreamSynchronize(s));
        C->col_ind_dev = ci.release(); C->values_dev = cv.release();
    }
    C->nnz_row_dev = nr.release();
    if (elapsed_ms) *elapsed_ms = ev_ms(0, 4);
    *done = 1;
    return IAS_OK;
}

static int ell_mul(const IasEllDev *A, const IasEllDev *B, IasEll64Dev *C, double *elapsed_ms)
{
    IAS_TRY(ensure_init());
    if (!A || !B || !C) return fail(IAS_E_ARG, "NULL");
    if (!A->choice || !B->choice) return fail(IAS_E_GATE, "ELL operand was rejected by the size gate (choice == false)");
    if (A->col > B->row) return fail(IAS_E_ARG, "shape mismatch: A is %dx%d, B is %dx%d", A->row, A->col, B->row, B->col);
    Ctx &c = ctx();
    int done = 0;
    IAS_TRY(ell_mul_onepass(A, B, C, elapsed_ms, &done));
    if (done) return IAS_OK;
    // general case: the Gustavson pipeline on fixed-width rows (EllView: no row-pointer gathers, row j of B starts at j*w)
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

int ias_download_ell64(const IasEll64Dev *d, int *nnz_row, int *col_ind, double *values)
{
    if (!d) return fail(IAS_E_ARG, "NULL");
    IasEllDev v;
    v.choice = d->choice; v.row = d->row; v.col = d->col; v.nnz = (int)std::min<long long>(d->nnz, 0x7fffffffLL);
    v.max_nnz_per_row = d->max_nnz_per_row; v.nnz_row_dev = d->nnz_row_dev; v.col_ind_dev = d->col_ind_dev; v.values_dev = d->values_dev;
    return ias_download_ell(&v, nnz_row, col_ind, values);
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
