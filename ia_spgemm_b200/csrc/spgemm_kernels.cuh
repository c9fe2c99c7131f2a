// spgemm_kernels.cuh -- device side of the Gustavson pipeline (shared by the CSR, ELL and COO entry points).
//
// Replaces: CSR_MUL_CSR (CPU/detail/csr/common_csr.h:85-193), the thrust ESC chain of
// CSR_MUL_CSR_DEV (GPU/detail/csr_dev/common_csr_dev.h:134-254), cusp::multiply (GPU/main.cu:482),
// cusparseXcsrgemmNnz + cusparseDcsrgemm (GPU/detail/cusparse/common_cusparse.h:78-91) and the
// Gustavson loops of ELL_MUL_ELL / COO_MUL_COO (ell:80-189, coo:72-161).
//
// Pipeline (one row range [r0, r1) of C at a time):
//   analyze   k_row_ub_thread  ub[i] = sum of B row lengths over A(i,:) (= the row's share of GetFlop), symbolic
//             k_row_ub_long    bin, bin histogram, total products, canonical check of A's rows; tiny rows are
//                              counted on the spot with the k-way merge (kept if B turns out canonical)
//   symbolic  k_sym_tiny       ub <= 32       thread per row: cursor merge (canonical B) or private list
//             k_sym_hash       ub <= 24576    warp / CTA per row, shared-memory hash of column keys
//             k_sym_gwin       larger         canonical B: persistent CTA per row, bitmap of a column super-window in
//                                             shared memory, only the part of each sorted B row inside it is visited
//             k_sym_global     larger         otherwise: {bitmap, rank} cells in global memory (L2)
//   scan      cub ExclusiveSum over nnz(C_i) -> 64-bit row pointers
//   numeric   k_num_tiny       ub <= 32       thread per row k-way merge, rows staged in smem, coalesced copy-out
//             k_esc_warp       ub <= 512      warp per row: expand - bitonic sort in registers - compress
//             k_num_hash_cta   nnz <= 12288   CTA per row: shared-memory hash SPA, block radix sort of
//                                             (column, slot), striped coalesced write
//             k_num_gwin       nnz > 4096     canonical B and a column space of one super-window (<= 262 144 columns):
//                                             {bitmap, rank} cells + rank window in shared memory, no global workspace
//             k_num_global     larger         bitmap + rank cells in an L2-resident slot: mark columns (in shared memory
//                                             when B is canonical), prefix-popcount, emit sorted columns, accumulate
//                                             windows of ranks in a dense shared-memory tile
//   products are enumerated by warp_products (one warp) / cta_products (balanced over a CTA, 4 per lane and trip) /
//   gwin_build + gwin_run (the same balance over B-row segments restricted to a column window)
// No intermediate product is ever materialised in global memory (the reference's ESC chain writes
// 24 B per product, csr_dev:170,190-194).
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace ias {
// Every translation unit that includes the pipeline gets its OWN copy of the kernels (IAS_TU is defined by the .cu file):
// identical template instantiations in two units would otherwise be merged by the linker, and the phase-clock build
// (make prof) would read the counters of the copy that lost.
inline namespace IAS_TU {

enum { BIN_EMPTY = 0, BIN_T = 1, BIN_W = 2, BIN_B1 = 3, BIN_B2 = 4, BIN_G = 5, NBINS = 6 };

// symbolic bins by upper bound (hash tables hold <= 50 % load)
constexpr int T_MAX = 32;
constexpr int SYM_W_UB = 512, SYM_W_TSIZE = 1024;
constexpr int SYM_B1_UB = 4096, SYM_B1_TSIZE = 8192;
constexpr int SYM_B2_UB = 24576, SYM_B2_TSIZE = 49152;
// numeric bins by exact nnz(C_i)
constexpr int NUM_W_UB = 512;
constexpr int NUM_B1_NNZ = 4096, NUM_B1_TSIZE = 8192;
constexpr int NUM_B2_NNZ = 12288, NUM_B2_TSIZE = 16384;

__host__ __device__ __forceinline__ int sym_bin_of(long long ub)
{
    if (ub == 0) return BIN_EMPTY;
    if (ub <= T_MAX) return BIN_T;
    if (ub <= SYM_W_UB) return BIN_W;
    if (ub <= SYM_B1_UB) return BIN_B1;
    if (ub <= SYM_B2_UB) return BIN_B2;
    return BIN_G;
}
// b2_max: largest nnz(C_i) the large CTA hash takes (NUM_B2_NNZ, or NUM_B1_NNZ when the windowed kernel is cheaper)
__host__ __device__ __forceinline__ int num_bin_of(int ub, int nnz, int b2_max = NUM_B2_NNZ)
{
    if (nnz == 0) return BIN_EMPTY;
    if (ub <= T_MAX) return BIN_T;
    if (ub <= NUM_W_UB) return BIN_W;                 // expand-sort-compress in registers holds every product
    if (nnz <= NUM_B1_NNZ) return BIN_B1;
    if (nnz <= b2_max) return BIN_B2;
    return BIN_G;
}

// ---------------------------------------------------------------- operand views
struct CsrView {                       // CsrMatrixDev, GPU/detail/format.h:59-69
    const int *rp; const int *ci; const double *v;
    typedef int off_t;
    __device__ __forceinline__ off_t begin(int i) const { return __ldg(rp + i); }
    __device__ __forceinline__ off_t end(int i) const { return __ldg(rp + i + 1); }
    __device__ __forceinline__ int len(int i) const { return __ldg(rp + i + 1) - __ldg(rp + i); }
    const void *rp_base() const { return rp; }
};
struct CsrRowsView {                   // an arbitrary list of rows of a CSR matrix: row li of the view = [rb[li], re[li]) in ci / v
    const int *rb; const int *re; const int *ci; const double *v;
    typedef int off_t;
    __device__ __forceinline__ off_t begin(int i) const { return __ldg(rb + i); }
    __device__ __forceinline__ off_t end(int i) const { return __ldg(re + i); }
    __device__ __forceinline__ int len(int i) const { return __ldg(re + i) - __ldg(rb + i); }
    const void *rp_base() const { return rb; }
};
struct Csr64View {                     // CSR with 64-bit row offsets (IasCooDev::row_offset_dev)
    const long long *rp; const int *ci; const double *v;
    typedef long long off_t;
    __device__ __forceinline__ off_t begin(int i) const { return __ldg(rp + i); }
    __device__ __forceinline__ off_t end(int i) const { return __ldg(rp + i + 1); }
    __device__ __forceinline__ int len(int i) const { return (int)(__ldg(rp + i + 1) - __ldg(rp + i)); }
    const void *rp_base() const { return rp; }
};
struct EllView {                       // EllMatrixDev, GPU/detail/format.h:108-119 (row-major, fixed width)
    const int *nr; const int *ci; const double *v; int w;
    typedef long long off_t;
    __device__ __forceinline__ off_t begin(int i) const { return (long long)i * w; }
    __device__ __forceinline__ off_t end(int i) const { return (long long)i * w + __ldg(nr + i); }
    __device__ __forceinline__ int len(int i) const { return __ldg(nr + i); }
    const void *rp_base() const { return nr; }
};

// where row li of the current range goes in the output arrays
struct OutMap {
    const long long *rp;     // CSR: 64-bit row pointers of the range (rp[li] - rp0 = start)
    long long rp0;
    const int *nnz_row;      // ELL: per-row counts, start = li * stride
    long long stride;
    __device__ __forceinline__ long long start(int li) const { return rp ? rp[li] - rp0 : (long long)li * stride; }
    __device__ __forceinline__ int count(int li) const { return rp ? (int)(rp[li + 1] - rp[li]) : nnz_row[li]; }
};

// ---------------------------------------------------------------- cursors of the tiny-row k-way merge (see "tiny rows" below)
template <class BV, int MERGE, bool WITH_VALUES>
struct TinyCursors {
    typename BV::off_t q[MERGE], qe[MERGE];
    int hc[MERGE];
    double hv[WITH_VALUES ? MERGE : 1];
    template <class AV>
    __device__ __forceinline__ void init(const AV &A, const BV &B, typename AV::off_t pa, int na)
    {
#pragma unroll
        for (int a = 0; a < MERGE; ++a) {
            hc[a] = 0x7fffffff; q[a] = 0; qe[a] = 0;
            if (WITH_VALUES) hv[a] = 0.0;
            if (a < na) {
                int j = __ldg(A.ci + pa + a);
                q[a] = B.begin(j); qe[a] = B.end(j);
                if (q[a] < qe[a]) {
                    hc[a] = __ldg(B.ci + q[a]);
                    if (WITH_VALUES) hv[a] = __ldg(B.v + q[a]);
                }
            }
        }
    }
    __device__ __forceinline__ int head() const
    {
        int m = hc[0];
#pragma unroll
        for (int a = 1; a < MERGE; ++a) m = min(m, hc[a]);
        return m;
    }
    __device__ __forceinline__ void advance(const BV &B, int a)
    {
        ++q[a];
        bool more = q[a] < qe[a];
        hc[a] = more ? __ldg(B.ci + q[a]) : 0x7fffffff;
        if (WITH_VALUES && more) hv[a] = __ldg(B.v + q[a]);
    }
};

// number of distinct columns of one tiny row (symbolic k-way merge)
template <class AV, class BV, int MERGE>
__device__ __forceinline__ int tiny_merge_count(const AV &A, const BV &B, typename AV::off_t pa, int na)
{
    TinyCursors<BV, MERGE, false> cur;
    cur.init(A, B, pa, na);
    int cnt = 0;
    while (cnt < T_MAX) {
        int m = cur.head();
        if (m == 0x7fffffff) break;
#pragma unroll
        for (int a = 0; a < MERGE; ++a)
            if (cur.hc[a] == m) cur.advance(B, a);
        ++cnt;
    }
    return cnt;
}

// ---------------------------------------------------------------- analyze
// block-level histogram of bins + sum of products.  Shared atomics are 32-bit and warp-aggregated:
// a 64-bit shared atomicAdd is a CAS loop, and with every row of a CTA in the same bin (Poisson,
// uniform) 256 threads would spin on one address.
__device__ __forceinline__ void block_hist(int bin, long long ub, unsigned *s_cnt /*NBINS*/, unsigned long long *s_sum,
                                           unsigned long long *__restrict__ g_hist /*NBINS + 1*/)
{
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int b = 0; b < NBINS; ++b) {
        unsigned m = __ballot_sync(0xffffffffu, bin == b);
        if (lane == 0 && m) atomicAdd(&s_cnt[b], (unsigned)__popc(m));
    }
    long long wsum = warp_sum(ub);
    if (lane == 0 && wsum) atomicAdd(s_sum, (unsigned long long)wsum);       // one CAS per warp
    __syncthreads();
    if (threadIdx.x < NBINS && s_cnt[threadIdx.x]) atomicAdd(&g_hist[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    if (threadIdx.x == NBINS && *s_sum) atomicAdd(&g_hist[NBINS], *s_sum);
}

// one thread per row; rows with more than LONG_A entries are deferred to k_row_ub_long (a thread would
// crawl through a 100 000-entry R-MAT hub row while its 255 neighbours idle)
constexpr int LONG_A = 64;

template <class AV, class BV>
__global__ void __launch_bounds__(256) k_row_ub_thread(int nrows, int r0, AV A, BV B, int *__restrict__ ub_out,
                                                       unsigned char *__restrict__ bin_out,
                                                       unsigned long long *__restrict__ g_hist /*16 slots*/,
                                                       int *__restrict__ long_list, int *__restrict__ long_count,
                                                       int *__restrict__ nnz_row /* optimistic tiny-row counts */)
{
    __shared__ unsigned s_cnt[NBINS];
    __shared__ unsigned long long s_sum;
    __shared__ int s_max[4];
    if (threadIdx.x < NBINS) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == NBINS) s_sum = 0;
    if (threadIdx.x < 4) s_max[threadIdx.x] = 0;
    __syncthreads();
    int li = blockIdx.x * blockDim.x + threadIdx.x;
    long long ub = 0;
    int bin = -1, tiny_na = 0, tiny_ub = 0, tiny_cnt = 0;
    if (li < nrows) {
        int i = r0 + li;
        typename AV::off_t pa = A.begin(i), pe = A.end(i);
        if (pe - pa > LONG_A) {
            long_list[atomicAdd(long_count, 1)] = li;
        } else {
            int prev = -1, unsorted = 0;
            for (typename AV::off_t p = pa; p < pe; ++p) {
                int j = __ldg(A.ci + p);
                unsorted |= (j <= prev);
                prev = j;
                ub += B.len(j);
            }
            if (unsorted) g_hist[NBINS + 1] = 1;      // A (== B for A^2) is not canonical
            bin = sym_bin_of(ub);
            tiny_na = (bin == BIN_T) ? (int)(pe - pa) : 0;
            tiny_ub = (bin == BIN_T) ? (int)ub : (bin == BIN_W ? -(int)ub : 0);   // negative: a warp-bin row
            ub_out[li] = ub > 0x7fffffffLL ? 0x7fffffff : (int)ub;
            bin_out[li] = (unsigned char)bin;
            // Tiny rows with few A entries: count nnz(C_i) right here with the k-way merge, ASSUMING B is canonical
            // (A's and B's lines are in L1 now).  The host keeps these counts only if the whole of B turns out
            // canonical and no tiny row has more than 8 entries; otherwise k_sym_tiny recomputes the bin.
            if (bin == BIN_T && tiny_na <= 8) {
                int cnt = tiny_na <= 5 ? tiny_merge_count<AV, BV, 5>(A, B, pa, tiny_na) : tiny_merge_count<AV, BV, 8>(A, B, pa, tiny_na);
                nnz_row[li] = cnt;
                tiny_cnt = cnt;
            }
        }
    }
    tiny_cnt = __reduce_max_sync(0xffffffffu, tiny_cnt);
    if ((threadIdx.x & 31) == 0 && tiny_cnt > 0) atomicMax(&s_max[3], tiny_cnt);
    int warp_ub = __reduce_max_sync(0xffffffffu, tiny_ub < 0 ? -tiny_ub : 0);
    tiny_na = __reduce_max_sync(0xffffffffu, tiny_na);
    tiny_ub = __reduce_max_sync(0xffffffffu, tiny_ub > 0 ? tiny_ub : 0);
    if ((threadIdx.x & 31) == 0) {
        if (tiny_na > 0) { atomicMax(&s_max[0], tiny_na); atomicMax(&s_max[1], tiny_ub); }
        if (warp_ub > 0) atomicMax(&s_max[2], warp_ub);
    }
    block_hist(bin, ub, s_cnt, &s_sum, g_hist);
    // one global atomic per CTA and statistic, and only when it would raise the maximum (plain read is a hint)
    if (threadIdx.x < 3 && s_max[threadIdx.x] > 0 && (unsigned long long)s_max[threadIdx.x] > g_hist[NBINS + 2 + threadIdx.x])
        atomicMax(&g_hist[NBINS + 2 + threadIdx.x], (unsigned long long)s_max[threadIdx.x]);
    if (threadIdx.x == 3 && s_max[3] > 0 && (unsigned long long)s_max[3] > g_hist[12]) atomicMax(&g_hist[12], (unsigned long long)s_max[3]);
}

// the deferred long rows: one warp per row, persistent grid
template <class AV, class BV>
__global__ void __launch_bounds__(256) k_row_ub_long(const int *__restrict__ long_list, const int *__restrict__ long_count, int r0,
                                                     AV A, BV B, int *__restrict__ ub_out, unsigned char *__restrict__ bin_out,
                                                     unsigned long long *__restrict__ g_hist)
{
    __shared__ unsigned s_cnt[NBINS];
    __shared__ unsigned long long s_sum;
    if (threadIdx.x < NBINS) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == NBINS) s_sum = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n = *long_count;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < n; idx += nwarps) {
        int li = long_list[idx];
        int i = r0 + li;
        typename AV::off_t pa = A.begin(i), pe = A.end(i);
        long long ub = 0;
        int unsorted = 0;
        for (typename AV::off_t p = pa + lane; p < pe; p += 32) {
            int j = __ldg(A.ci + p);
            if (p > pa) unsorted |= (j <= __ldg(A.ci + p - 1));
            ub += B.len(j);
        }
        if (unsorted) g_hist[NBINS + 1] = 1;
        ub = warp_sum(ub);
        if (lane == 0) {
            int bin = sym_bin_of(ub);                      // never tiny: more than LONG_A entries
            ub_out[li] = ub > 0x7fffffffLL ? 0x7fffffff : (int)ub;
            bin_out[li] = (unsigned char)bin;
            atomicAdd(&s_cnt[bin], 1u);
            if (ub) atomicAdd(&s_sum, (unsigned long long)ub);
            if (bin == BIN_W && (unsigned long long)ub > g_hist[NBINS + 4]) atomicMax(&g_hist[NBINS + 4], (unsigned long long)ub);
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS && s_cnt[threadIdx.x]) atomicAdd(&g_hist[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    if (threadIdx.x == NBINS && s_sum) atomicAdd(&g_hist[NBINS], s_sum);
}

// numeric bins from exact counts; also the largest nnz(C_i) among tiny rows (sizes k_num_tiny's smem)
static __global__ void __launch_bounds__(256) k_classify_num(int nrows, const int *__restrict__ ub, const int *__restrict__ nnz_row,
                                                      unsigned char *__restrict__ bin_out,
                                                      unsigned long long *__restrict__ g_hist /*NBINS+2: [NBINS]=max tiny nnz, [NBINS+1]=max nnz*/,
                                                      int b2_max)
{
    __shared__ unsigned s_cnt[NBINS];
    __shared__ int s_max[2];
    if (threadIdx.x < NBINS) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x >= NBINS && threadIdx.x < NBINS + 2) s_max[threadIdx.x - NBINS] = 0;
    __syncthreads();
    int li = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    int bin = -1, n = 0;
    if (li < nrows) {
        n = nnz_row[li];
        bin = num_bin_of(ub[li], n, b2_max);
        bin_out[li] = (unsigned char)bin;
    }
#pragma unroll
    for (int b = 0; b < NBINS; ++b) {
        unsigned m = __ballot_sync(0xffffffffu, bin == b);
        if (lane == 0 && m) atomicAdd(&s_cnt[b], (unsigned)__popc(m));
    }
    int mt = __reduce_max_sync(0xffffffffu, bin == BIN_T ? n : 0);
    int ma = __reduce_max_sync(0xffffffffu, n);
    if (lane == 0) { atomicMax(&s_max[0], mt); atomicMax(&s_max[1], ma); }
    __syncthreads();
    if (threadIdx.x < NBINS && s_cnt[threadIdx.x]) atomicAdd(&g_hist[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    if (threadIdx.x >= NBINS && threadIdx.x < NBINS + 2) atomicMax(&g_hist[threadIdx.x], (unsigned long long)s_max[threadIdx.x - NBINS]);
}

// (key, value) = (work of the row, row) for ordering a bin's rows by decreasing work
static __global__ void __launch_bounds__(256) k_work_keys(int n, const int *__restrict__ list, const int *__restrict__ ub,
                                                   unsigned *__restrict__ keys, int *__restrict__ vals)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int li = list ? list[t] : t;
    keys[t] = (unsigned)ub[li];
    vals[t] = li;
}

static __global__ void k_iota(int n, int *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// ---------------------------------------------------------------- tiny rows: one thread per row
// Two per-thread algorithms:
//   merge   (B canonical and the A row has <= MERGE entries): the row of C is the k-way merge of the
//           sorted B rows it touches.  One cursor per A entry lives in registers (fully unrolled); every
//           step takes the smallest head column, adds up all heads equal to it (in A order, the order
//           CSR_MUL_CSR accumulates in, csr:150-170) and advances them, prefetching the next head's
//           column and value.  Output comes out column sorted: no search, no sort.
//   list    (anything else): unsorted private list with linear search, insertion-sorted at the end.
// MERGE is a template parameter (4 / 6 / 8) picked by the host from the longest A row in the bin.
// one merged row into (mc, mv); returns the number of distinct columns (at most `limit`)
template <class AV, class BV, int MERGE>
__device__ __forceinline__ int tiny_merge_numeric(const AV &A, const BV &B, typename AV::off_t pa, int na, int *mc, double *mv, int limit)
{
    TinyCursors<BV, MERGE, true> cur;
    double av[MERGE];
    cur.init(A, B, pa, na);
#pragma unroll
    for (int a = 0; a < MERGE; ++a) av[a] = a < na ? __ldg(A.v + pa + a) : 0.0;
    int cnt = 0;
    while (cnt < limit) {
        int m = cur.head();
        if (m == 0x7fffffff) break;
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int a = 0; a < MERGE; ++a)
            if (cur.hc[a] == m) {
                double x = av[a] * cur.hv[a];
                acc = first ? x : acc + x;
                first = false;
                cur.advance(B, a);
            }
        mc[cnt] = m; mv[cnt] = acc; ++cnt;
    }
    return cnt;
}

template <class AV, class BV, int BLOCK, int MERGE>
__global__ void __launch_bounds__(BLOCK) k_sym_tiny(const int *__restrict__ rows, int nrows, int r0, AV A, BV B,
                                                    int *__restrict__ nnz_row, int b_canonical,
                                                    unsigned long long *__restrict__ max_out /* largest nnz(C_i) seen */)
{
    __shared__ int list[T_MAX * BLOCK];          // [slot][thread]: conflict free (list path only)
    int idx = blockIdx.x * BLOCK + threadIdx.x;
    // lanes of this warp that own a row: taken before any divergence, so the reduction below names exactly them
    const unsigned mask = __ballot_sync(0xffffffffu, idx < nrows);
    if (idx >= nrows) return;
    int li = rows ? rows[idx] : idx;
    int i = r0 + li;
    typename AV::off_t pa = A.begin(i), pe = A.end(i);
    int na = (int)(pe - pa);
    int cnt = 0;
    if (b_canonical && na <= MERGE) {
        TinyCursors<BV, MERGE, false> cur;
        cur.init(A, B, pa, na);
        while (true) {
            int m = cur.head();
            if (m == 0x7fffffff) break;
#pragma unroll
            for (int a = 0; a < MERGE; ++a)
                if (cur.hc[a] == m) cur.advance(B, a);
            ++cnt;
        }
    } else {
        int *mine = list + threadIdx.x;
        for (typename AV::off_t p = pa; p < pe; ++p) {
            int j = __ldg(A.ci + p);
            typename BV::off_t qe = B.end(j);
            for (typename BV::off_t q = B.begin(j); q < qe; ++q) {
                int k = __ldg(B.ci + q);
                int s = 0;
                for (; s < cnt; ++s)
                    if (mine[s * BLOCK] == k) break;
                if (s == cnt) { mine[cnt * BLOCK] = k; ++cnt; }
            }
        }
    }
    nnz_row[li] = cnt;
    __syncwarp(mask);
    int wmax = __reduce_max_sync(mask, cnt);
    if ((threadIdx.x & 31) == 0 && (unsigned long long)wmax > *max_out) atomicMax(max_out, (unsigned long long)wmax);
}

template <class AV, class BV, int BLOCK, int MERGE>
__global__ void __launch_bounds__(BLOCK) k_num_tiny(const int *__restrict__ rows, int nrows, int r0, AV A, BV B, OutMap out,
                                                    int *__restrict__ c_ci, double *__restrict__ c_v, int cap, int b_canonical, int bulk)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *vals = reinterpret_cast<double *>(smem_raw);      // every thread owns exactly nnz(C_i) slots (+ 2: alignment shift)
    int *cols = reinterpret_cast<int *>(vals + (size_t)BLOCK * cap + 2);
    typedef cub::BlockScan<int, BLOCK> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;

    int idx = blockIdx.x * BLOCK + threadIdx.x;
    int li = -1, n = 0;
    long long gstart = 0;
    if (idx < nrows) {
        li = rows ? rows[idx] : idx;
        n = out.count(li);
        gstart = out.start(li);
    }
    // identity row list: the CTA's rows are consecutive, so its output is one contiguous span starting at `base`.
    // For the bulk copy-out the staged span is shifted so that a 16-byte boundary in shared memory is one in global memory.
    const bool ident = rows == nullptr && out.rp != nullptr;
    long long base = 0;
    int sv = 0, sc = 0;
    if (ident) {
        base = out.start(blockIdx.x * BLOCK);
        if (bulk) { sv = (int)(base & 1); sc = (int)(base & 3); }
    }
    int off, total;
    Scan(scan_tmp).ExclusiveSum(n, off, total);
    if (n > 0) {
        int i = r0 + li;
        int *mc = cols + sc + off;
        double *mv = vals + sv + off;
        typename AV::off_t pa = A.begin(i), pe = A.end(i);
        int na = (int)(pe - pa);
        if (b_canonical && na <= MERGE) {
            tiny_merge_numeric<AV, BV, MERGE>(A, B, pa, na, mc, mv, n);
        } else {
            int cnt = 0;
            for (typename AV::off_t p = pa; p < pe; ++p) {
                int j = __ldg(A.ci + p);
                double av = __ldg(A.v + p);
                typename BV::off_t qe = B.end(j);
                for (typename BV::off_t q = B.begin(j); q < qe; ++q) {
                    int k = __ldg(B.ci + q);
                    double x = av * __ldg(B.v + q);
                    int s = 0;
                    for (; s < cnt; ++s)
                        if (mc[s] == k) break;
                    if (s < cnt) mv[s] += x;
                    else if (cnt < n) { mc[cnt] = k; mv[cnt] = x; ++cnt; }
                }
            }
            for (int a = 1; a < cnt; ++a) {                       // insertion sort by column
                int kk = mc[a];
                double vv = mv[a];
                int b = a - 1;
                while (b >= 0 && mc[b] > kk) { mc[b + 1] = mc[b]; mv[b + 1] = mv[b]; --b; }
                mc[b + 1] = kk; mv[b + 1] = vv;
            }
        }
    }
    __syncthreads();
    if (ident && bulk && total >= 64) {
        // one elected thread hands the 16-byte aligned middle of both arrays to the copy engine (cp.async.bulk, shared ->
        // global); the few head / tail elements go out with ordinary stores
        const int hv = sv, nv = (total - hv) & ~1;                 // values: 8-byte elements
        const int hc = (4 - sc) & 3, nc = (total - hc) & ~3;       // columns: 4-byte elements
        if (threadIdx.x == 0) {
            fence_proxy_async_shared();
            bulk_store_shared_to_global(c_v + base + hv, vals + sv + hv, (unsigned)nv * 8u);
            bulk_store_shared_to_global(c_ci + base + hc, cols + sc + hc, (unsigned)nc * 4u);
            bulk_commit_group();
        }
        const int e = threadIdx.x;                                 // at most 1 + 1 values and 3 + 3 columns are left over
        if (e < hv) c_v[base + e] = vals[sv + e];
        if (e >= 8 && e < 8 + (total - hv - nv)) c_v[base + hv + nv + (e - 8)] = vals[sv + hv + nv + (e - 8)];
        if (e >= 16 && e < 16 + hc) c_ci[base + (e - 16)] = cols[sc + (e - 16)];
        if (e >= 24 && e < 24 + (total - hc - nc)) c_ci[base + hc + nc + (e - 24)] = cols[sc + hc + nc + (e - 24)];
        if (threadIdx.x == 0) bulk_wait_group_read0();             // shared memory must outlive the copy's reads
    } else if (ident) {
        for (int e = threadIdx.x; e < total; e += BLOCK) { c_ci[base + e] = cols[sc + e]; c_v[base + e] = vals[sv + e]; }
    } else {
        int lane = threadIdx.x & 31;
        for (int r = 0; r < 32; ++r) {
            int rn = __shfl_sync(0xffffffffu, n, r);
            int ro = __shfl_sync(0xffffffffu, off, r);
            long long rg = __shfl_sync(0xffffffffu, gstart, r);
            if (lane < rn) { c_ci[rg + lane] = cols[ro + lane]; c_v[rg + lane] = vals[ro + lane]; }
        }
    }
}

// B rows strictly increasing?  (needed when B is not the operand the analyze kernel walks)
template <class BV>
__global__ void __launch_bounds__(256) k_rows_canonical(int nrows, BV B, unsigned long long *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    int prev = -1, bad = 0;
    typename BV::off_t qe = B.end(i);
    for (typename BV::off_t q = B.begin(i); q < qe; ++q) { int k = __ldg(B.ci + q); bad |= (k <= prev); prev = k; }
    if (bad) *flag = 1;
}

// ---------------------------------------------------------------- hash rows: warp (TPR=32) or CTA (TPR=BLOCK) per row
template <int TPR>
__device__ __forceinline__ void group_sync()
{
    if (TPR == 32) __syncwarp(); else __syncthreads();
}

// ---- product iterator shared by the warp / CTA / global kernels
// A warp walks the A row in chunks of 32 entries: lane l fetches entry l (column j, value, start and
// length of B row j) in one coalesced + one gathered access, so the a_ci -> b_rp -> b_ci dependency
// chain is paid once per 32 entries instead of once per entry.  B rows of LONG_ROW or more entries are
// strided by the whole warp; the short ones are flattened: their products are numbered 0..T-1 with a
// warp scan and lane t finds its (entry, offset) with a 5-step shuffle search, so lanes stay busy on
// power-law operands whose B rows are mostly a handful of entries.  f(q, a_value, pidx) is called once
// per product with q = index into B.ci / B.v and pidx = dense arrival index of the product within this
// warp's share of the row.  Returns the number of products visited.  All 32 lanes must call this together.
constexpr int LONG_ROW = 24;

template <bool NEED_VALUES, class AV, class BV, class F>
__device__ __forceinline__ int warp_products(const AV &A, const BV &B, typename AV::off_t pa, typename AV::off_t pe,
                                             int w, int nw, int lane, F &&f)
{
    typedef typename AV::off_t aoff;
    typedef typename BV::off_t boff;
    int pcount = 0;
    for (aoff base = pa + (aoff)w * 32; base < pe; base += (aoff)nw * 32) {
        aoff p = base + lane;
        int len = 0;
        boff qb = 0;
        double av = 0.0;
        if (p < pe) {
            int j = __ldg(A.ci + p);
            qb = B.begin(j);
            len = (int)(B.end(j) - qb);
            if (NEED_VALUES) av = __ldg(A.v + p);
        }
        unsigned longmask = __ballot_sync(0xffffffffu, len >= LONG_ROW);
        while (longmask) {
            int src = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            boff sqb = __shfl_sync(0xffffffffu, qb, src);
            int slen = __shfl_sync(0xffffffffu, len, src);
            double sav = NEED_VALUES ? __shfl_sync(0xffffffffu, av, src) : 0.0;
            for (int t = lane; t < slen; t += 32) f(sqb + t, sav, pcount + t);
            pcount += slen;
        }
        int slen = len >= LONG_ROW ? 0 : len;
        int incl = slen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        int total = __shfl_sync(0xffffffffu, incl, 31);
        boff rel = qb - (boff)(incl - slen);          // q = rel + t for the products of this lane's entry
        for (int t0 = 0; t0 < total; t0 += 32) {
            int t = t0 + lane;
            int src = 0;                                // smallest lane whose inclusive count exceeds t
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                if (v <= t) src += step;
            }
            src = min(src, 31);
            boff srel = __shfl_sync(0xffffffffu, rel, src);
            double sav = NEED_VALUES ? __shfl_sync(0xffffffffu, av, src) : 0.0;
            if (t < total) f(srel + t, sav, pcount + t);
        }
        pcount += total;
    }
    return pcount;
}

// ---- the same for a whole CTA: balanced over all threads.
// ncu on R-MAT showed CTA-per-row kernels issuing 3 % of the time, the rest spent at barriers behind
// the one warp that drew a hub row.  Here the A row is taken in tiles of BLOCK entries: thread t loads
// entry t, a block scan numbers the tile's products 0..T-1, and the products are dealt out evenly to the
// warps (see the loop below).  Every warp gets the same number of products.
template <class BV, int BLOCK>
struct CtaTile {
    typedef cub::BlockScan<int, BLOCK> Scan;
    typename Scan::TempStorage scan;
    int incl[BLOCK];
    typename BV::off_t rel[BLOCK];
    double av[BLOCK];
};

struct NoRestrict {
    template <class O>
    __device__ __forceinline__ void operator()(O &, int &) const {}
};

// columns of a sorted B row restricted to [c_lo, c_hi): two lower bounds
template <class BV>
struct ColumnWindow {
    const BV &B;
    int c_lo, c_hi;
    __device__ __forceinline__ typename BV::off_t lower(typename BV::off_t b, int len, int c) const
    {
        int lo = 0, hi = len;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(B.ci + b + mid) < c) lo = mid + 1; else hi = mid; }
        return b + lo;
    }
    __device__ __forceinline__ void operator()(typename BV::off_t &qb, int &len) const
    {
        if (len == 0) return;
        // both bounds in one loop: two independent loads in flight per step instead of two dependent chains one after
        // the other (a thread-per-entry search is pure L2 latency; ncu/phase clocks: "acc build" of rows with more than
        // BLOCK entries in A was 13 % of the second-generation global-row kernel)
        int lo0 = 0, hi0 = c_lo > 0 ? len : 0;                    // first index with column >= c_lo
        int lo1 = c_hi == 0x7fffffff ? len : 0, hi1 = len;        // first index with column >= c_hi
        while (lo0 < hi0 || lo1 < hi1) {
            const int m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1;
            const int k0 = lo0 < hi0 ? __ldg(B.ci + qb + m0) : 0;
            const int k1 = lo1 < hi1 ? __ldg(B.ci + qb + m1) : 0;
            if (lo0 < hi0) { if (k0 < c_lo) lo0 = m0 + 1; else hi0 = m0; }
            if (lo1 < hi1) { if (k1 < c_hi) lo1 = m1 + 1; else hi1 = m1; }
        }
        qb += lo0; len = lo1 - lo0;
    }
};

constexpr int PB = 4;      // products per lane and trip in cta_products

template <bool NEED_VALUES, int BLOCK, class AV, class BV, class F, class R = NoRestrict>
__device__ __forceinline__ void cta_products(const AV &A, const BV &B, typename AV::off_t pa, typename AV::off_t pe,
                                             CtaTile<BV, BLOCK> &tile, F &&f, const R &restrict_range = R())
{
    typedef typename AV::off_t aoff;
    typedef typename BV::off_t boff;
    const int tid = threadIdx.x;
    for (aoff base = pa; base < pe; base += BLOCK) {
        aoff p = base + tid;
        int len = 0;
        boff qb = 0;
        double av = 0.0;
        if (p < pe) {
            int j = __ldg(A.ci + p);
            qb = B.begin(j);
            len = (int)(B.end(j) - qb);
            restrict_range(qb, len);
            if (NEED_VALUES) av = __ldg(A.v + p);
        }
        int incl;
        typename CtaTile<BV, BLOCK>::Scan(tile.scan).InclusiveSum(len, incl);
        tile.incl[tid] = incl;
        tile.rel[tid] = qb - (boff)(incl - len);
        if (NEED_VALUES) tile.av[tid] = av;
        __syncthreads();
        // every warp takes an equal, contiguous share of the tile's products; its lanes take consecutive
        // products (coalesced B reads).  The owning entry is found once per warp by binary search and then
        // followed forward: on hub rows the scan loop exits at once, on short rows it advances a few entries.
        const int total = tile.incl[BLOCK - 1];
        constexpr int NWARPS = BLOCK / 32;
        const int wid = tid >> 5, lane = tid & 31;
        const int share = ((total + NWARPS - 1) / NWARPS + 31) & ~31;
        const int t_begin = wid * share, t_end = min(total, t_begin + share);
        if (t_begin < t_end) {
            int lo = 0, hi = BLOCK - 1;                    // smallest entry whose inclusive count exceeds t_begin
            while (lo < hi) { int mid = (lo + hi) >> 1; if (tile.incl[mid] <= t_begin) lo = mid + 1; else hi = mid; }
            int e = lo;
            // PB products per lane and trip: f receives them together so that it can issue all its loads
            // before the first dependent use (the kernels are bound by L2 latency, not by issue slots)
            for (int t0 = t_begin; t0 < t_end; t0 += 32 * PB) {
                typename BV::off_t q[PB];
                double pav[PB];
                unsigned valid = 0;
                int me = e;
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    int t = t0 + 32 * u + lane;
                    q[u] = 0; pav[u] = 0.0;
                    if (t < t_end) {
                        while (tile.incl[me] <= t) ++me;
                        q[u] = tile.rel[me] + t;
                        if (NEED_VALUES) pav[u] = tile.av[me];
                        valid |= 1u << u;
                    }
                }
                f(q, pav, valid);
                e = __shfl_sync(0xffffffffu, me, 31);      // lane 31 holds the furthest entry reached in this trip
            }
        }
        __syncthreads();                                   // the tile arrays are rewritten by the next pass
    }
}

template <int TSIZE>
__device__ __forceinline__ unsigned hash_insert_key(int *keys, int k, int &fresh)
{
    unsigned s = __umulhi(hash_col(k), (unsigned)TSIZE);
    while (true) {
        int cur = keys[s];
        if (cur == k) break;
        if (cur == -1) {
            int old = atomicCAS(&keys[s], -1, k);
            if (old == -1) { ++fresh; break; }
            if (old == k) break;
        }
        s = (s + 1 == TSIZE) ? 0 : s + 1;
    }
    return s;
}

template <class AV, class BV, int TPR, int BLOCK, int TSIZE>
__global__ void __launch_bounds__(BLOCK) k_sym_hash(const int *__restrict__ rows, int nrows, int r0, AV A, BV B,
                                                    int *__restrict__ nnz_row)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ int s_cnt;
    int g = threadIdx.x / TPR, t = threadIdx.x % TPR, lane = t & 31;
    int idx = blockIdx.x * (BLOCK / TPR) + g;
    if (idx >= nrows) return;                       // whole group leaves together
    int *keys = reinterpret_cast<int *>(smem_raw) + (size_t)g * TSIZE;
    for (int s = t; s < TSIZE; s += TPR) keys[s] = -1;
    if (TPR > 32 && threadIdx.x == 0) s_cnt = 0;
    group_sync<TPR>();
    int li = rows ? rows[idx] : idx;
    int i = r0 + li;
    int cnt = 0;
    if (TPR == 32) {
        warp_products<false>(A, B, A.begin(i), A.end(i), 0, 1, lane, [&](typename BV::off_t q, double, int) {
            hash_insert_key<TSIZE>(keys, __ldg(B.ci + q), cnt);
        });
    } else {
        __shared__ CtaTile<BV, TPR == 32 ? 32 : BLOCK> tile;
        cta_products<false, TPR == 32 ? 32 : BLOCK>(A, B, A.begin(i), A.end(i), tile,
                                                    [&](const typename BV::off_t (&q)[PB], const double (&)[PB], unsigned valid) {
            int k[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u) k[u] = (valid >> u) & 1u ? __ldg(B.ci + q[u]) : -1;
#pragma unroll
            for (int u = 0; u < PB; ++u)
                if (k[u] >= 0) hash_insert_key<TSIZE>(keys, k[u], cnt);
        });
    }
    cnt = warp_sum(cnt);
    if (TPR == 32) {
        if (lane == 0) nnz_row[li] = cnt;
    } else {
        if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) nnz_row[li] = s_cnt;
    }
}

// ---- CTA per row: hash SPA in shared memory, then a block radix sort of (column, slot) pairs.
// The table is never compacted: empty slots carry key 0xffffffff and sort behind every column, the
// sorted keys come back striped so the CTA writes the row with fully coalesced stores, and each value
// is fetched from its slot.  The sort's scratch aliases the key table (keys are in registers by then).
template <int BLOCK, int TSIZE>
struct NumHashSmem {
    static constexpr int IPT = TSIZE / BLOCK;
    typedef cub::BlockRadixSort<unsigned, BLOCK, IPT, unsigned> Sort;
    static constexpr size_t KEY_BYTES = sizeof(typename Sort::TempStorage) > (size_t)TSIZE * 4 ? sizeof(typename Sort::TempStorage) : (size_t)TSIZE * 4;
    static constexpr size_t BYTES = (size_t)TSIZE * 8 + ((KEY_BYTES + 15) / 16) * 16;
};

template <class AV, class BV, int BLOCK, int TSIZE>
__global__ void __launch_bounds__(BLOCK) k_num_hash_cta(const int *__restrict__ rows, int nrows, int r0, AV A, BV B, OutMap out,
                                                        int *__restrict__ c_ci, double *__restrict__ c_v, int col_bits)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef NumHashSmem<BLOCK, TSIZE> SM;
    typedef typename SM::Sort Sort;
    constexpr int IPT = SM::IPT;
    double *vals = reinterpret_cast<double *>(smem_raw);
    int *keys = reinterpret_cast<int *>(smem_raw + (size_t)TSIZE * 8);
    typename Sort::TempStorage &sort_tmp = *reinterpret_cast<typename Sort::TempStorage *>(smem_raw + (size_t)TSIZE * 8);
    int idx = blockIdx.x;
    if (idx >= nrows) return;
    int t = threadIdx.x;
    for (int s = t; s < TSIZE; s += BLOCK) { keys[s] = -1; vals[s] = 0.0; }
    __syncthreads();
    int li = rows ? rows[idx] : idx;
    int i = r0 + li;
    int fresh = 0;
    __shared__ CtaTile<BV, BLOCK> tile;
    cta_products<true, BLOCK>(A, B, A.begin(i), A.end(i), tile,
                              [&](const typename BV::off_t (&q)[PB], const double (&av)[PB], unsigned valid) {
        int k[PB];
        double x[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            k[u] = -1; x[u] = 0.0;
            if ((valid >> u) & 1u) { k[u] = __ldg(B.ci + q[u]); x[u] = av[u] * __ldg(B.v + q[u]); }
        }
#pragma unroll
        for (int u = 0; u < PB; ++u)
            if (k[u] >= 0) { unsigned s = hash_insert_key<TSIZE>(keys, k[u], fresh); atomicAdd(&vals[s], x[u]); }
    });
    __syncthreads();
    unsigned k[IPT], slot[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) { slot[r] = t * IPT + r; k[r] = (unsigned)keys[t * IPT + r]; }
    __syncthreads();                                   // every key is in registers: the table may be overwritten
    Sort(sort_tmp).SortBlockedToStriped(k, slot, 0, col_bits + 1);
    int n = out.count(li);
    long long gs = out.start(li);
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        int pos = r * BLOCK + t;
        if (pos < n) { c_ci[gs + pos] = (int)k[r]; c_v[gs + pos] = vals[slot[r]]; }
    }
}

// ---- warp per row, at most 32*IPL products: expand - sort - compress entirely in registers.
// Every product gets the key (column << IDX_BITS | arrival index); a bitonic network over the warp's
// 32*IPL register-resident keys sorts by column and, inside a column, by arrival order (the order
// CSR_MUL_CSR accumulates in); a segmented scan adds up equal columns and the last product of each
// column writes the entry.  Values wait in shared memory (8 B per product) and are fetched by index.
template <class KeyT>
__device__ __forceinline__ void cex(KeyT &a, KeyT &b, bool up)
{
    KeyT lo = a < b ? a : b, hi = a < b ? b : a;
    a = up ? lo : hi; b = up ? hi : lo;
}

template <class KeyT, int IPL>
__device__ __forceinline__ void warp_bitonic_sort(KeyT (&key)[IPL], int lane)
{
    // element index e = lane*IPL + r (blocked); ascending overall
    constexpr int N = 32 * IPL;
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            if (jj >= IPL) {
                int lj = jj / IPL;                                   // partner lane distance
                bool lower = (lane & lj) == 0;
#pragma unroll
                for (int r = 0; r < IPL; ++r) {
                    int e = lane * IPL + r;
                    bool up = (e & kk) == 0;
                    KeyT other = __shfl_xor_sync(0xffffffffu, key[r], lj);
                    bool keep_min = (lower == up);
                    key[r] = keep_min ? (key[r] < other ? key[r] : other) : (key[r] < other ? other : key[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < IPL; ++r) {
                    if ((r & jj) == 0) {
                        int e = lane * IPL + r;
                        cex(key[r], key[r | jj], (e & kk) == 0);
                    }
                }
            }
        }
    }
}

template <class AV, class BV, class KeyT, int IPL, int BLOCK, bool NUMERIC>
__global__ void __launch_bounds__(BLOCK) k_esc_warp(const int *__restrict__ rows, int nrows, int r0, AV A, BV B, OutMap out,
                                                    int *__restrict__ nnz_row, int *__restrict__ c_ci, double *__restrict__ c_v)
{
    constexpr int N = 32 * IPL;
    constexpr int IDX_BITS = NUMERIC ? (IPL == 16 ? 9 : IPL == 8 ? 8 : IPL == 4 ? 7 : 6) : 0;
    constexpr int WARPS = BLOCK / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int idx = blockIdx.x * WARPS + w;
    if (idx >= nrows) return;
    KeyT *skeys = reinterpret_cast<KeyT *>(smem_raw) + (size_t)w * N;
    double *svals = reinterpret_cast<double *>(smem_raw + (size_t)WARPS * N * sizeof(KeyT)) + (size_t)w * N;
    int li = rows ? rows[idx] : idx;
    int i = r0 + li;
    const KeyT PAD = ~(KeyT)0;
    for (int s = lane; s < N; s += 32) skeys[s] = PAD;
    __syncwarp();
    warp_products<NUMERIC>(A, B, A.begin(i), A.end(i), 0, 1, lane, [&](typename BV::off_t q, double av, int pos) {
        if (pos < N) {                                  // always true: the host sizes IPL from the largest ub of the bin
            skeys[pos] = ((KeyT)(unsigned)__ldg(B.ci + q) << IDX_BITS) | (KeyT)(NUMERIC ? pos : 0);
            if (NUMERIC) svals[pos] = av * __ldg(B.v + q);
        }
    });
    __syncwarp();
    KeyT key[IPL];
#pragma unroll
    for (int r = 0; r < IPL; ++r) key[r] = skeys[lane * IPL + r];
    warp_bitonic_sort<KeyT, IPL>(key, lane);
    // neighbours: column of the element after my last one (for tail detection)
    KeyT next_first = __shfl_down_sync(0xffffffffu, key[0], 1);
    if (lane == 31) next_first = PAD;
    int tails = 0;
    bool is_tail[IPL];
#pragma unroll
    for (int r = 0; r < IPL; ++r) {
        KeyT nxt = r + 1 < IPL ? key[r + 1] : next_first;
        is_tail[r] = key[r] != PAD && (nxt == PAD || (nxt >> IDX_BITS) != (key[r] >> IDX_BITS));
        tails += is_tail[r];
    }
    int tincl = tails;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, tincl, o); if (lane >= o) tincl += u; }
    if (!NUMERIC) {
        if (lane == 31) nnz_row[li] = tincl;
        return;
    }
    // segmented sums: heads are elements whose predecessor has another column
    KeyT prev_last = __shfl_up_sync(0xffffffffu, key[IPL - 1], 1);
    double v[IPL];
    bool head[IPL];
    bool any_head = false;
    double tail_sum = 0.0;                             // sum since the last head in this lane (all items if none)
#pragma unroll
    for (int r = 0; r < IPL; ++r) {
        bool valid = key[r] != PAD;
        KeyT prv = r > 0 ? key[r - 1] : prev_last;
        head[r] = valid && ((r == 0 && lane == 0) || (prv >> IDX_BITS) != (key[r] >> IDX_BITS));
        v[r] = valid ? svals[(unsigned)(key[r] & (((KeyT)1 << IDX_BITS) - 1))] : 0.0;
        tail_sum = head[r] ? v[r] : tail_sum + v[r];
        any_head |= head[r];
    }
    // carry from the lanes to the left, stopping at the nearest lane that contains a head
    double carry = tail_sum;
    bool flag = any_head;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double uc = __shfl_up_sync(0xffffffffu, carry, o);
        bool uf = __shfl_up_sync(0xffffffffu, flag, o);
        if (lane >= o && !flag) { carry += uc; flag = uf; }
    }
    double carry_in = __shfl_up_sync(0xffffffffu, carry, 1);
    if (lane == 0) carry_in = 0.0;
    long long gs = out.start(li);
    int opos = tincl - tails;
    double run = carry_in;
#pragma unroll
    for (int r = 0; r < IPL; ++r) {
        run = head[r] ? v[r] : run + v[r];
        if (is_tail[r]) { c_ci[gs + opos] = (int)(key[r] >> IDX_BITS); c_v[gs + opos] = run; ++opos; }
    }
}

// ---------------------------------------------------------------- global rows, windowed in shared memory
// The L2 variant above pays one L2 transaction per product and pass (ncu: 31 % issue, DRAM 3 %, L2-latency
// bound at ~100 G products/s).  A microbenchmark on B200 (tools/ubench/smem_atomics.cu) gives 7.4 lanes/clk/SM
// for random ATOMS.OR / LDS and 1.9 lanes/clk/SM for the fp64 atomicAdd CAS loop in shared memory, i.e.
// 2.2 T and 0.55 T products/s for the chip, so every per-product random access of these kernels goes to
// shared memory.  Needs canonical B (strictly increasing columns per row): the column space is cut into
// super-windows of 32*swords columns, and only the part of each B row inside the current window is visited.
//   symbolic: bitmap of the super-window in shared memory (up to 1.5 M columns), popcount, clear.
//   numeric:  {bitmap word, rank of its first bit} cells of the super-window in shared memory; mark, warp-scan
//             the populations, then rank windows of at most `win` entries: sorted columns staged in scol[],
//             products added into acc[] (fp64, indexed by rank), both written out once, coalesced.
// A restricted visit needs two lower bounds per A entry: rows with few A entries (the common case, ~70 at
// R-MAT scale 22) let a whole warp search one B row with 32 probes per step (3 dependent loads instead of 11
// for a 2 000-entry row); longer A rows search one entry per thread.

// optional phase clocks (make prof): thread 0 of every CTA adds up clock64 deltas per phase
#ifdef IAS_GWIN_PROFILE
static __device__ unsigned long long g_gwin_prof[64];
#define GP_START() long long gp_t = clock64()
#define GP_ADD(slot)                                                                           \
    do {                                                                                       \
        if (threadIdx.x == 0) {                                                                \
            long long gp_n = clock64();                                                        \
            atomicAdd(&g_gwin_prof[slot], (unsigned long long)(gp_n - gp_t));                  \
            gp_t = gp_n;                                                                       \
        }                                                                                      \
    } while (0)
#define GP_CNT(slot, v)                                                                        \
    do {                                                                                       \
        if (threadIdx.x == 0) atomicAdd(&g_gwin_prof[slot], (unsigned long long)(v));          \
    } while (0)
#else
#define GP_START() do {} while (0)
#define GP_ADD(slot) do {} while (0)
#define GP_CNT(slot, v) do {} while (0)
#endif

__device__ __forceinline__ bool use_tbl_probe(bool single_tile, int nsw, int n_a, int tstride, int tbl_cap)
{
    return single_tile && nsw > 1 && (long long)n_a * tstride <= tbl_cap;
}

// first index in ci[0, len) whose column is >= c; all 32 lanes call with the same arguments
__device__ __forceinline__ int warp_lower_bound(const int *__restrict__ ci, int len, int c, int lane)
{
    int lo = 0, hi = len;
    while (lo < hi) {
        const int step = (hi - lo) / 33 + 1;
        const int pos = lo + (lane + 1) * step - 1;
        const bool less = pos < hi && __ldg(ci + pos) < c;
        const int cnt = __popc(__ballot_sync(0xffffffffu, less));     // sorted: the first cnt probes are smaller
        const int nlo = lo + cnt * step;
        hi = min(hi, nlo + step - 1);
        lo = nlo;
    }
    return lo;
}

constexpr int GWIN_COOP_MAX = 128;    // A-row tiles of at most this many entries use the warp-wide search

// restrict every thread's B row (qb, len) of a tile with n_in entries to columns [c_lo, c_hi)
template <int BLOCK, class BV>
__device__ __forceinline__ void gwin_restrict(const BV &B, int n_in, typename BV::off_t &qb, int &len,
                                              CtaTile<BV, BLOCK> &tile, int c_lo, int c_hi)
{
    typedef typename BV::off_t boff;
    constexpr int NWARPS = BLOCK / 32;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (c_lo <= 0 && c_hi == 0x7fffffff) return;
    if (n_in <= GWIN_COOP_MAX) {
        tile.rel[tid] = qb;
        tile.incl[tid] = len;
        __syncthreads();
        for (int e = wid; e < n_in; e += NWARPS) {
            const boff b = tile.rel[e];
            const int l = tile.incl[e];
            const int s = c_lo > 0 ? warp_lower_bound(B.ci + b, l, c_lo, lane) : 0;
            const int t = c_hi != 0x7fffffff ? s + warp_lower_bound(B.ci + b + s, l - s, c_hi, lane) : l;
            __syncwarp();
            if (lane == 0) { tile.rel[e] = b + s; tile.incl[e] = t - s; }
        }
        __syncthreads();
        qb = tile.rel[tid];
        len = tile.incl[tid];
        __syncthreads();
    } else {
        ColumnWindow<BV>{B, c_lo, c_hi}(qb, len);
    }
}

// scan the tile's segment lengths; leaves the tile in shared memory (ends with a barrier), returns its product count
template <bool NEED_VALUES, int BLOCK, class BV>
__device__ __forceinline__ int gwin_scan(CtaTile<BV, BLOCK> &tile, typename BV::off_t qb, int len, double av)
{
    typedef typename BV::off_t boff;
    const int tid = threadIdx.x;
    int incl;
    typename CtaTile<BV, BLOCK>::Scan(tile.scan).InclusiveSum(len, incl);
    tile.incl[tid] = incl;
    tile.rel[tid] = qb - (boff)(incl - len);
    if (NEED_VALUES) tile.av[tid] = av;
    __syncthreads();
    return tile.incl[BLOCK - 1];
}

// entry `base + tid` of the A row: B row start, length, A value
template <bool NEED_VALUES, class AV, class BV>
__device__ __forceinline__ void gwin_load(const AV &A, const BV &B, typename AV::off_t base, typename AV::off_t pe,
                                          typename BV::off_t &qb, int &len, double &av)
{
    typename AV::off_t p = base + threadIdx.x;
    len = 0; qb = 0; av = 0.0;
    if (p < pe) {
        int j = __ldg(A.ci + p);
        qb = B.begin(j);
        len = (int)(B.end(j) - qb);
        if (NEED_VALUES) av = __ldg(A.v + p);
    }
}

// one tile of BLOCK A entries starting at `base`, B rows restricted to columns [c_lo, c_hi)
template <bool NEED_VALUES, int BLOCK, class AV, class BV>
__device__ __forceinline__ int gwin_build(const AV &A, const BV &B, typename AV::off_t base, typename AV::off_t pe,
                                          CtaTile<BV, BLOCK> &tile, int c_lo, int c_hi)
{
    typename BV::off_t qb;
    int len;
    double av;
    gwin_load<NEED_VALUES>(A, B, base, pe, qb, len, av);
    gwin_restrict<BLOCK>(B, (int)min((typename AV::off_t)BLOCK, pe - base), qb, len, tile, c_lo, c_hi);
    return gwin_scan<NEED_VALUES, BLOCK>(tile, qb, len, av);
}

// the products of a built tile, dealt out evenly to the warps (same scheme as cta_products); no barrier.
// A trip of 32*PB products that lies inside one B-row segment (the common case on hub-heavy rows) takes its
// offsets from one broadcast read instead of searching the tile per product.
template <bool NEED_VALUES, int BLOCK, class BV, class F>
__device__ __forceinline__ void gwin_run(const CtaTile<BV, BLOCK> &tile, int total, F &&f)
{
    constexpr int NWARPS = BLOCK / 32;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int share = ((total + NWARPS - 1) / NWARPS + 31) & ~31;
    const int t_begin = wid * share, t_end = min(total, t_begin + share);
    if (t_begin >= t_end) return;
    int lo = 0, hi = BLOCK - 1;                    // smallest entry whose inclusive count exceeds t_begin
    while (lo < hi) { int mid = (lo + hi) >> 1; if (tile.incl[mid] <= t_begin) lo = mid + 1; else hi = mid; }
    int e = lo;
    for (int t0 = t_begin; t0 < t_end; t0 += 32 * PB) {
        typename BV::off_t q[PB];
        double pav[PB];
        unsigned valid = 0;
        while (tile.incl[e] <= t0) ++e;            // warp-uniform: the entry that owns product t0
        const int trip_end = min(t_end, t0 + 32 * PB);
        if (tile.incl[e] >= trip_end) {
            const typename BV::off_t r = tile.rel[e] + t0 + lane;
            const double a = NEED_VALUES ? tile.av[e] : 0.0;
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                q[u] = r + 32 * u;
                pav[u] = a;
                if (t0 + 32 * u + lane < trip_end) valid |= 1u << u;
            }
            f(q, pav, valid);
        } else {
            int me = e;
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                int t = t0 + 32 * u + lane;
                q[u] = 0; pav[u] = 0.0;
                if (t < t_end) {
                    while (tile.incl[me] <= t) ++me;
                    q[u] = tile.rel[me] + t;
                    if (NEED_VALUES) pav[u] = tile.av[me];
                    valid |= 1u << u;
                }
            }
            f(q, pav, valid);
            e = __shfl_sync(0xffffffffu, me, 31);
        }
    }
}

// ---------------------------------------------------------------- global rows: bitmap + rank in L2
// Workspace of one slot (all 32-bit words, zero between rows except prefix/blkpref):
//   wp[words]       {bitmap word, output rank of its first bit}   words = ceil(ncols/32) rounded up to 32
//   summary[sumw]   one bit per block of 32 bitmap words (= 1024 columns)
//   blkpref[blocks] output rank of each block           blocks = words/32, sumw = ceil(blocks/32)
//   wsum[blocks]    which of a block's 32 words are non-zero: the per-row scans read 4 bytes per block and
//                   only the populated cells instead of the whole 8*cols/32-byte cell array
struct GLayout {
    int words, blocks, sumw;
    size_t slot_words;          // words + words + sumw + blocks
    __host__ __device__ static GLayout make(int ncols)
    {
        GLayout g;
        long long w = ((long long)ncols + 31) / 32;
        w = (w + 31) / 32 * 32;
        g.words = (int)w; g.blocks = g.words / 32; g.sumw = (g.blocks + 31) / 32;
        g.slot_words = ((size_t)g.words * 2 + g.sumw + 2 * (size_t)g.blocks + 31) / 32 * 32;   // keeps every slot 128-byte aligned
        return g;
    }
};

// bitmap word and rank prefix of the same 32 columns share one 8-byte cell {bits, rank}: the accumulate
// pass needs both and gets them with a single L2 access (the global path is bound by L2 transactions)
// a word that turns non-zero registers itself in its block's word mask, a block that turns non-zero in the summary
__device__ __forceinline__ void g_first_touch(unsigned *summary, unsigned *wsum, int w)
{
    unsigned oldm = atomicOr(wsum + (w >> 5), 1u << (w & 31));
    if (oldm == 0) atomicOr(summary + (w >> 10), 1u << ((w >> 5) & 31));
}

// PB products at once: all column loads, then all cell loads, then the (rare) atomics
template <class BV>
__device__ __forceinline__ void g_mark_batch(const BV &B, uint2 *wp, unsigned *summary, unsigned *wsum,
                                             const typename BV::off_t (&q)[PB], unsigned valid, int &cnt)
{
    int k[PB];
    unsigned cur[PB];
#pragma unroll
    for (int u = 0; u < PB; ++u) k[u] = (valid >> u) & 1u ? __ldg(B.ci + q[u]) : -1;
#pragma unroll
    for (int u = 0; u < PB; ++u) cur[u] = k[u] >= 0 ? __ldcg(&wp[k[u] >> 5].x) : 0xffffffffu;
#pragma unroll
    for (int u = 0; u < PB; ++u) {
        if (k[u] < 0) continue;
        int w = k[u] >> 5;
        unsigned bit = 1u << (k[u] & 31);
        if (!(cur[u] & bit)) {
            unsigned old = atomicOr(&wp[w].x, bit);
            if (!(old & bit)) {
                ++cnt;
                if (old == 0) g_first_touch(summary, wsum, w);
            }
        }
    }
}

// zero exactly the cells that were touched (found through wsum), one thread per 1024-column block: the
// per-row passes over the workspace are latency bound, so every thread issues its block's accesses at once
template <int BLOCK>
__device__ __forceinline__ void g_clear(uint2 *wp, unsigned *summary, unsigned *wsum, const GLayout &L)
{
    for (int b = threadIdx.x; b < L.blocks; b += BLOCK) {
        unsigned m = __ldcg(wsum + b);
        if (!m) continue;
        wsum[b] = 0;
        uint2 *cell = wp + (size_t)b * 32;
#pragma unroll
        for (int x = 0; x < 32; ++x)
            if ((m >> x) & 1u) cell[x] = make_uint2(0, 0);
    }
    for (int sw = threadIdx.x; sw < L.sumw; sw += BLOCK) summary[sw] = 0;
}

template <class AV, class BV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sym_global(const int *__restrict__ rows, int nrows, int r0, AV A, BV B,
                                                      int *__restrict__ nnz_row, unsigned *__restrict__ work, GLayout L,
                                                      int *__restrict__ cursor)
{
    __shared__ int s_row, s_cnt;
    __shared__ CtaTile<BV, BLOCK> tile;
    uint2 *wp = reinterpret_cast<uint2 *>(work + (size_t)blockIdx.x * L.slot_words);
    unsigned *summary = reinterpret_cast<unsigned *>(wp + L.words);
    unsigned *wsum = summary + L.sumw + L.blocks;
    int lane = threadIdx.x & 31;
    while (true) {
        if (threadIdx.x == 0) { s_row = atomicAdd(cursor, 1); s_cnt = 0; }
        __syncthreads();
        int idx = s_row;
        if (idx >= nrows) break;
        int li = rows ? rows[idx] : idx;
        int i = r0 + li;
        int cnt = 0;
        cta_products<false, BLOCK>(A, B, A.begin(i), A.end(i), tile,
                                   [&](const typename BV::off_t (&q)[PB], const double (&)[PB], unsigned valid) {
                                       g_mark_batch(B, wp, summary, wsum, q, valid, cnt);
                                   });
        cnt = warp_sum(cnt);
        if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) nnz_row[li] = s_cnt;
        g_clear<BLOCK>(wp, summary, wsum, L);
        __syncthreads();
    }
}

// Numeric pass of a global row.  The values are NOT accumulated in global memory: with hundreds of
// rows in flight the value segments (up to 16 MB each at R-MAT scale 22) thrash the L2 and every RED
// becomes an HBM round trip.  Instead the row is cut into windows of WIN consecutive ranks; a window's
// values live in a dense shared-memory tile indexed by rank, products are added with shared-memory
// atomics, and the tile is written out once, coalesced.  With canonical B only the part of each B row
// inside the window's column range is visited (two binary searches per A entry and window); most
// global rows need one or two windows.
template <class AV, class BV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_num_global(const int *__restrict__ rows, int nrows, int r0, AV A, BV B, OutMap out,
                                                      int *__restrict__ c_ci, double *__restrict__ c_v,
                                                      unsigned *__restrict__ work, GLayout L, int *__restrict__ cursor,
                                                      int win, int b_canonical, int smem_mark, int ncols)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *acc = reinterpret_cast<double *>(smem_raw);           // win entries
    typedef cub::BlockScan<unsigned, BLOCK> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ int s_row, s_lo, s_hi;
    __shared__ unsigned s_carry;
    __shared__ CtaTile<BV, BLOCK> tile;
    uint2 *wp = reinterpret_cast<uint2 *>(work + (size_t)blockIdx.x * L.slot_words);
    unsigned *summary = reinterpret_cast<unsigned *>(wp + L.words);
    unsigned *wsum = summary + L.sumw + L.blocks;
    while (true) {
        if (threadIdx.x == 0) { s_row = atomicAdd(cursor, 1); s_carry = 0; }
        __syncthreads();
        int idx = s_row;
        if (idx >= nrows) break;
        int li = rows ? rows[idx] : idx;
        int i = r0 + li;
        long long gs = out.start(li);
        const int n = out.count(li);
        typename AV::off_t pa = A.begin(i), pe = A.end(i);
        GP_START();
        GP_CNT(32, 1);
        // 1. mark the columns of the row.  Canonical B: in shared memory, one super-window of 2*win*32 columns at a
        //    time (the bitmap borrows the accumulate tile), each flushed to the row's cells with coalesced stores --
        //    no L2 atomic per product (that pass was 38 % of this kernel at R-MAT scale 22).
        if (smem_mark & 1) {
            unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);
            const int swords = win * 2;                                  // 32-bit words in the tile's bytes
            const long long span = (long long)swords * 32;
            for (int w = threadIdx.x; w < swords; w += BLOCK) bits[w] = 0;
            __syncthreads();
            for (long long sw_lo = 0; sw_lo < ncols; sw_lo += span) {
                const int c_lo = (int)sw_lo;
                const int c_hi = sw_lo + span >= ncols ? 0x7fffffff : (int)(sw_lo + span);
                bool any = false;
                for (typename AV::off_t base = pa; base < pe; base += BLOCK) {
                    const int total = gwin_build<false, BLOCK>(A, B, base, pe, tile, c_lo, c_hi);
                    GP_ADD(33);
                    if (total) {
                        any = true;
                        gwin_run<false, BLOCK>(tile, total, [&](const typename BV::off_t (&q)[PB], const double (&)[PB], unsigned valid) {
                            int k[PB];
                            unsigned cur[PB];
#pragma unroll
                            for (int u = 0; u < PB; ++u) k[u] = (valid >> u) & 1u ? __ldg(B.ci + q[u]) - c_lo : -1;
#pragma unroll
                            for (int u = 0; u < PB; ++u) cur[u] = k[u] >= 0 ? bits[k[u] >> 5] : 0xffffffffu;
#pragma unroll
                            for (int u = 0; u < PB; ++u) {
                                const unsigned bit = 1u << (k[u] & 31);
                                if (!(cur[u] & bit)) atomicOr(&bits[k[u] >> 5], bit);
                            }
                        });
                    }
                    __syncthreads();
                    GP_ADD(34);
                }
                if (!any) continue;
                // flush: a warp per block of 32 words; the block's word mask comes from one ballot
                const int w0g = c_lo >> 5;                                // first global word of the super-window
                const int nwords = min(swords, L.words - w0g);
                for (int wb = (threadIdx.x >> 5) * 32; wb < nwords; wb += BLOCK) {
                    const int w = wb + (threadIdx.x & 31);
                    const unsigned word = bits[w];                        // nwords is a multiple of 32 (L.words and swords are)
                    const unsigned m = __ballot_sync(0xffffffffu, word != 0u);
                    if (word) { wp[w0g + w].x = word; bits[w] = 0; }
                    if (m && (threadIdx.x & 31) == 0) wsum[(w0g + wb) >> 5] = m;
                }
                __syncthreads();
                GP_ADD(35);
            }
        } else {
            int dummy = 0;
            cta_products<false, BLOCK>(A, B, pa, pe, tile,
                                       [&](const typename BV::off_t (&q)[PB], const double (&)[PB], unsigned valid) {
                                           g_mark_batch(B, wp, summary, wsum, q, valid, dummy);
                                       });
        }
        __syncthreads();
        // 2. ranks.  One thread per 1024-column block: fetch the block's populated words in one burst (32
        //    independent predicated loads), block-scan the populations, then write each word's rank and emit
        //    the sorted column list.  (A warp-per-block, entry-parallel emission with coalesced stores was tried:
        //    its per-block L2 round trips are serial per warp and it lost 20 % at R-MAT scale 20/22.)
        for (int bb = 0; bb < L.blocks; bb += BLOCK) {
            const int b = bb + threadIdx.x;
            unsigned m = b < L.blocks ? __ldcg(wsum + b) : 0u;
            const uint2 *cell = wp + (size_t)b * 32;
            unsigned word[32];
            unsigned c = 0;
#pragma unroll
            for (int x = 0; x < 32; ++x) {
                word[x] = (m >> x) & 1u ? __ldcg(&cell[x].x) : 0u;
                c += __popc(word[x]);
            }
            unsigned excl, tile_total;
            Scan(scan_tmp).ExclusiveSum(c, excl, tile_total);
            __syncthreads();                               // scan_tmp is reused by the next trip
            unsigned rank = s_carry + excl;                 // s_carry: entries of the blocks of earlier trips
            if (m) {
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    unsigned wd = word[x];
                    if (!wd) continue;
                    wp[(size_t)b * 32 + x].y = rank;
                    while (wd) {
                        int bit = __ffs(wd) - 1;
                        wd &= wd - 1;
                        c_ci[gs + rank] = (b * 32 + x) * 32 + bit;
                        ++rank;
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_carry += tile_total;
            __syncthreads();
        }
        __threadfence();          // window bounds below are read back from c_ci by another thread
        __syncthreads();
        GP_ADD(36);
        // 3. one window of `win` ranks at a time: accumulate in the shared-memory tile, then write it out
        for (int wbase = 0; wbase < n; wbase += win) {
            const int wn = min(win, n - wbase);
            if (threadIdx.x == 0) {
                s_lo = wbase == 0 ? 0 : __ldcg(c_ci + gs + wbase);
                s_hi = wbase + win < n ? __ldcg(c_ci + gs + wbase + win) : 0x7fffffff;
            }
            for (int t = threadIdx.x; t < wn; t += BLOCK) acc[t] = 0.0;
            __syncthreads();
            GP_ADD(37);
            GP_CNT(42, 1);
            const int c_lo = s_lo, c_hi = s_hi;
            auto add = [&](const typename BV::off_t (&q)[PB], const double (&av)[PB], unsigned valid) {
                int k[PB];
                double x[PB];
                uint2 cell[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    k[u] = -1; x[u] = 0.0;
                    if ((valid >> u) & 1u) { k[u] = __ldg(B.ci + q[u]); x[u] = av[u] * __ldg(B.v + q[u]); }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    if (k[u] < c_lo || k[u] >= c_hi) k[u] = -1;
                    // (smem_mark & 4: through L1 -- the cells of a dense window's column range are re-read many times and
                    // nothing else writes them during the pass; the __threadfence above dropped every stale line)
                    cell[u] = k[u] < 0 ? make_uint2(0, 0) : (smem_mark & 4) ? wp[k[u] >> 5] : __ldcg(wp + (k[u] >> 5));
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    if (k[u] < 0) continue;
                    unsigned below = cell[u].x & ((1u << (k[u] & 31)) - 1u);
                    atomicAdd(&acc[(int)(cell[u].y + __popc(below)) - wbase], x[u]);
                }
            };
            if (smem_mark & 2) {
                // canonical B: warp-wide lower bounds for short A rows, whole-trip fast path (see gwin_build / gwin_run)
                const bool cut = n > win;
                for (typename AV::off_t base = pa; base < pe; base += BLOCK) {
                    const int total = gwin_build<true, BLOCK>(A, B, base, pe, tile, cut ? c_lo : 0, cut ? c_hi : 0x7fffffff);
                    GP_ADD(38);
                    if (total) gwin_run<true, BLOCK>(tile, total, add);
                    __syncthreads();
                    GP_ADD(39);
                }
            } else if (b_canonical && n > win) cta_products<true, BLOCK>(A, B, pa, pe, tile, add, ColumnWindow<BV>{B, c_lo, c_hi});
            else cta_products<true, BLOCK>(A, B, pa, pe, tile, add);
            __syncthreads();
            for (int t = threadIdx.x; t < wn; t += BLOCK) c_v[gs + wbase + t] = acc[t];
            __syncthreads();
            GP_ADD(40);
        }
        // 4. leave the slot clean for the next row
        g_clear<BLOCK>(wp, summary, wsum, L);
        __syncthreads();
        GP_ADD(41);
    }
}

// ---------------------------------------------------------------- global rows, second generation (canonical B)
// Phase clocks of k_num_global at R-MAT scale 22 (profiles/r02_summary.md): the products themselves (mark + accumulate
// runs) were a third of the kernel; the rest went to a separate rank + emit pass over the row's cells in L2 (23 %), to
// the two lower bounds per A entry and window (acc build 13 %, mark build 5 %) and to batch tails.  This version
//   * ranks and emits straight from the shared-memory bitmap of the mark pass: a warp takes 32 words, scans their
//     populations with shuffles, writes the {bitmap, rank} cells (only needed by the accumulate look-ups now) and the
//     sorted column list -- no pass over L2 at all, no __threadfence;
//   * finds every window boundary of every B row once per row, all in parallel (split tables in shared memory:
//     tbl[e][k] = first entry of B row e with column >= boundary k), so that a window's product list is built from
//     two shared-memory reads per A entry instead of two dependent binary searches in L2;
//   * keeps the A row's entries (B row start, length, A value) in registers for rows of at most BLOCK entries.
// Rows with more than BLOCK entries in A, or whose tables do not fit, take the per-window searches of the first
// generation.  Dynamic shared memory: win doubles (accumulate tile, also the mark bitmap) + tbl_cap ints.
// four independent binary searches per thread in lock step (a single search is a chain of dependent L2 loads; the
// split tables need thousands of them per row): lo[u] = first index in [0, len[u]) of ci + b[u] with column >= c[u]
template <class BV>
__device__ __forceinline__ void lower_bound4(const BV &B, const typename BV::off_t (&b)[4], const int (&len)[4], const int (&c)[4], int (&lo)[4])
{
    int hi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { lo[u] = 0; hi[u] = len[u]; }
    while ((lo[0] < hi[0]) | (lo[1] < hi[1]) | (lo[2] < hi[2]) | (lo[3] < hi[3])) {
        int mid[4], k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { mid[u] = (lo[u] + hi[u]) >> 1; k[u] = lo[u] < hi[u] ? __ldg(B.ci + b[u] + mid[u]) : 0; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (lo[u] < hi[u]) { if (k[u] < c[u]) lo[u] = mid[u] + 1; else hi[u] = mid[u]; }
    }
}

constexpr int G2_MAX_BND = 512;          // window boundaries kept in shared memory (rows beyond read them from c_ci)

template <class AV, class BV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_num_global2(const int *__restrict__ rows, int nrows, int r0, AV A, BV B, OutMap out,
                                                       int *__restrict__ c_ci, double *__restrict__ c_v,
                                                       unsigned *__restrict__ work, GLayout L, int *__restrict__ cursor,
                                                       int win, int tbl_cap, int ncols, int *__restrict__ gscr, int gscr_cap,
                                                       const int *__restrict__ n_split_ptr, int split_cap, int split_words,
                                                       int *__restrict__ part_cnt)
{
    typedef typename AV::off_t aoff;
    typedef typename BV::off_t boff;
    constexpr int NWARPS = BLOCK / 32;
    constexpr bool QB32 = sizeof(boff) == 4;          // B row starts fit an int (CSR); otherwise the row index is kept
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *acc = reinterpret_cast<double *>(smem_raw);                       // win entries
    unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);                  // mark bitmap: 2 * win words in the same bytes
    int *tbl = reinterpret_cast<int *>(smem_raw + (size_t)win * 8);           // tbl_cap split points
    typedef cub::BlockScan<unsigned, BLOCK> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ int s_row, s_next;
    __shared__ unsigned s_carry, s_swtot, s_base;
    __shared__ unsigned s_blk[BLOCK];                                         // populations of a super-window's chunks of 128 words
    __shared__ int s_bnd[G2_MAX_BND + 1];
    __shared__ CtaTile<BV, BLOCK> tile;
    uint2 *wp = reinterpret_cast<uint2 *>(work + (size_t)blockIdx.x * L.slot_words);
    unsigned *summary = reinterpret_cast<unsigned *>(wp + L.words);
    unsigned *wsum = summary + L.sumw + L.blocks;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int swords = min(win * 2, 2 * BLOCK * 32);                           // a multiple of 128 (the host rounds win to 64)
    const long long span = (long long)swords * 32;
    const int nsw = (int)((ncols + span - 1) / span);
    // Split rows.  The first n_split rows of the (work-ordered) list are cut into P parts by column range, one work item
    // each: a part is the product of the A row with B restricted to the part's columns (the restriction is applied where
    // the B rows' starts and lengths are loaded), marked and ranked in ONE bitmap pass of split_words words.  A part's
    // place in the C row is the sum of the counts of the parts before it: every part publishes its count (part_cnt,
    // zero = not yet) as soon as its bitmap is counted and waits for its predecessors' before it emits.  Items are drawn
    // in order from the cursor, so the parts a CTA waits for were drawn earlier by CTAs that wait only for still earlier
    // ones.  Without this a hub row is one CTA's work from start to end: at R-MAT scale 25 almost half of the kernel's
    // time was tails behind single rows (profiles/r02_summary.md, section 6).
    const int P = (QB32 && rows && n_split_ptr && split_words > 0) ? (L.words + split_words - 1) / split_words : 0;
    const int n_split = P > 1 ? min(min(__ldg(n_split_ptr), split_cap), nrows) : 0;
    const int n_items = n_split * P + (nrows - n_split);
    for (int w = tid; w < swords; w += BLOCK) bits[w] = 0;
    while (true) {
        if (tid == 0) { s_row = atomicAdd(cursor, 1); s_carry = 0; s_base = 0; }
        __syncthreads();
        const int idx = s_row;
        if (idx >= n_items) break;
        int part = -1, ridx = idx - n_split * P + n_split;
        if (idx < n_split * P) { ridx = idx / max(P, 1); part = idx - ridx * P; }
        const bool split = part >= 0;
        const int row_lo = split ? part * split_words * 32 : 0;                                   // the item's column range
        const int row_hi = split && part + 1 < P ? row_lo + split_words * 32 : 0x7fffffff;
        const int li = rows ? rows[ridx] : ridx;
        const int i = r0 + li;
        long long gs = out.start(li);                  // (a part: moved to its place in the row once the counts before it are known)
        int n = out.count(li);
        const aoff pa = A.begin(i), pe = A.end(i);
        const int n_a = (int)min((aoff)0x7fffffff, pe - pa);
        const bool single_tile = pe - pa <= (aoff)BLOCK;
        GP_START();
        GP_CNT(32, 1);
        boff my_qb = 0;
        int my_len = 0;
        double my_av = 0.0;
        if (single_tile) {
            gwin_load<true>(A, B, pa, pe, my_qb, my_len, my_av);
            if (split) ColumnWindow<BV>{B, row_lo, row_hi}(my_qb, my_len);
        }
        // Rows with more than BLOCK entries in A (the hubs): the entries' B rows (start, length) and their split points
        // go to a per-CTA scratch in global memory (L2), boundary-major so that a window's tile reads them coalesced.
        // Every boundary of every B row is searched once (the upper bound of a window is the lower bound of the next),
        // two searches in flight per thread; a window's tile then costs four coalesced loads instead of two dependent
        // binary searches behind an A.ci -> B.rp gather.
        int *gx = gscr + (size_t)blockIdx.x * gscr_cap;               // n_a starts (or B row indices), then n_a lengths, then the table
        int *glen = gx + n_a;
        int *gt = glen + n_a;
        const bool long_cached = !single_tile && (long long)n_a * 4 <= gscr_cap;
        if (long_cached) {
            for (int e = tid; e < n_a; e += BLOCK) {
                const int j = __ldg(A.ci + pa + e);
                boff b = B.begin(j);
                int l = B.len(j);
                if (split) ColumnWindow<BV>{B, row_lo, row_hi}(b, l);            // (split implies QB32)
                gx[e] = QB32 ? (int)b : j;
                glen[e] = l;
            }
            __syncthreads();
        }
        // one tile of a long row straight from the cached (start, length) pairs (a part's are already restricted)
        auto cached_tile = [&](aoff base, bool need_values) {
            const int e = (int)(base - pa) + tid;
            boff qb = 0;
            int len = 0;
            double av = 0.0;
            if (e < n_a) {
                qb = (boff)__ldcg(gx + e);
                len = __ldcg(glen + e);
                if (need_values) av = __ldg(A.v + pa + e);
            }
            return need_values ? gwin_scan<true, BLOCK>(tile, qb, len, av) : gwin_scan<false, BLOCK>(tile, qb, len, 0.0);
        };
        auto long_table = [&](int K, auto boundary) {                 // gt[k * n_a + e], k = 0 .. K; boundary(k) = column of boundary k
            for (int e = tid; e < n_a; e += BLOCK) { gt[e] = 0; gt[(size_t)K * n_a + e] = __ldcg(glen + e); }
            const int items = n_a * (K - 1);
            for (int it0 = tid; it0 < items; it0 += 4 * BLOCK) {
                boff b[4];
                int len[4], cc[4], lo[4];
                size_t slot[4];
                bool on[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int it = it0 + u * BLOCK;
                    b[u] = 0; len[u] = 0; cc[u] = 0; slot[u] = 0; on[u] = it < items;
                    if (on[u]) {
                        const int k = it / n_a + 1, e = it - (k - 1) * n_a;
                        const int x = __ldcg(gx + e);
                        b[u] = QB32 ? (boff)x : B.begin(x);
                        len[u] = __ldcg(glen + e); cc[u] = boundary(k); slot[u] = (size_t)k * n_a + e;
                    }
                }
                lower_bound4(B, b, len, cc, lo);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (on[u]) gt[slot[u]] = lo[u];
            }
            __syncthreads();
        };
        // one tile of a long row restricted to window k of the current table: returns the tile's product count
        auto long_tile = [&](aoff base, int k, bool need_values) {
            const int e = (int)(base - pa) + tid;
            boff qb = 0;
            int len = 0;
            double av = 0.0;
            if (e < n_a) {
                const int o0 = __ldcg(gt + (size_t)k * n_a + e), o1 = __ldcg(gt + (size_t)(k + 1) * n_a + e);
                const int x = __ldcg(gx + e);
                qb = (QB32 ? (boff)x : B.begin(x)) + o0;
                len = o1 - o0;
                if (need_values) av = __ldg(A.v + pa + e);
            }
            return need_values ? gwin_scan<true, BLOCK>(tile, qb, len, av) : gwin_scan<false, BLOCK>(tile, qb, len, 0.0);
        };
        const bool use_gmtbl = !split && long_cached && nsw > 1 && (long long)n_a * (nsw + 3) <= gscr_cap;
        if (use_gmtbl) long_table(nsw, [&](int k) { return (int)(k * span); });
        // split table of the mark pass: where every B row crosses the super-window boundaries
        const int mstride = (nsw + 1) | 1;
        const bool use_mtbl = !split && single_tile && nsw > 1 && (long long)n_a * mstride <= tbl_cap;
        if (use_mtbl) {
            tile.rel[tid] = my_qb;
            tile.incl[tid] = my_len;
            __syncthreads();
            const int items = n_a * (nsw - 1);
            for (int it0 = tid; it0 < items; it0 += 4 * BLOCK) {
                boff b[4];
                int len[4], cc[4], lo[4], slot[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int it = it0 + u * BLOCK;
                    b[u] = 0; len[u] = 0; cc[u] = 0; slot[u] = -1;
                    if (it < items) {
                        const int e = it / (nsw - 1), k = it - e * (nsw - 1) + 1;
                        b[u] = tile.rel[e]; len[u] = tile.incl[e]; cc[u] = (int)(k * span); slot[u] = e * mstride + k;
                    }
                }
                lower_bound4(B, b, len, cc, lo);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (slot[u] >= 0) tbl[slot[u]] = lo[u];
            }
            if (tid < n_a) { tbl[tid * mstride] = 0; tbl[tid * mstride + nsw] = my_len; }
            __syncthreads();
        }
        GP_ADD(33);
        // 1. mark, rank and emit, one super-window of the column space at a time
        auto mark = [&](const boff (&q)[PB], const double (&)[PB], unsigned valid, int c_lo) {
            int k[PB];
            unsigned cur[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u) k[u] = (valid >> u) & 1u ? __ldg(B.ci + q[u]) - c_lo : -1;
#pragma unroll
            for (int u = 0; u < PB; ++u) cur[u] = k[u] >= 0 ? bits[k[u] >> 5] : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const unsigned bit = 1u << (k[u] & 31);
                if (!(cur[u] & bit)) atomicOr(&bits[k[u] >> 5], bit);
            }
        };
        for (int k = split ? part : 0; k < (split ? part + 1 : nsw); ++k) {
            const long long sw_lo = split ? (long long)row_lo : k * span;
            const int c_lo = (int)sw_lo;
            const int c_hi = split ? row_hi : k + 1 == nsw ? 0x7fffffff : (int)(sw_lo + span);
            bool any = false;
            if (single_tile) {
                boff qb = my_qb;
                int len = my_len;
                if (use_mtbl) {
                    const int o0 = tid < n_a ? tbl[tid * mstride + k] : 0, o1 = tid < n_a ? tbl[tid * mstride + k + 1] : 0;
                    qb = my_qb + o0; len = o1 - o0;
                } else if (!split && nsw > 1) gwin_restrict<BLOCK>(B, n_a, qb, len, tile, c_lo, c_hi);
                const int total = gwin_scan<false, BLOCK>(tile, qb, len, 0.0);
                if (total) {
                    any = true;
                    gwin_run<false, BLOCK>(tile, total, [&](const boff (&q)[PB], const double (&pv)[PB], unsigned valid) { mark(q, pv, valid, c_lo); });
                }
                __syncthreads();
            } else {
                for (aoff base = pa; base < pe; base += BLOCK) {
                    const int total = use_gmtbl ? long_tile(base, k, false)
                                      : split && long_cached ? cached_tile(base, false)
                                                             : gwin_build<false, BLOCK>(A, B, base, pe, tile, c_lo, c_hi);
                    if (total) {
                        any = true;
                        gwin_run<false, BLOCK>(tile, total, [&](const boff (&q)[PB], const double (&pv)[PB], unsigned valid) { mark(q, pv, valid, c_lo); });
                    }
                    __syncthreads();
                }
            }
            GP_ADD(34);
            if (!any) {
                if (split) {                                          // an empty part: its count is zero, nothing to emit or accumulate
                    n = 0;
                    if (tid == 0) atomicExch(part_cnt + idx, 1);
                }
                continue;
            }
            const int w0g = c_lo >> 5;                                // first global word of the super-window
            const int nwords = min(split ? split_words : swords, L.words - w0g);   // a multiple of 32 (L.words, swords and split_words are)
            // The bitmap is scanned in chunks of 128 words, four consecutive words per lane (one 128-bit shared-memory
            // load): one warp scan ranks 4096 columns.  (Ranking 32 words per warp scan cost more than the marking itself
            // on sparse rows: the dependent shuffle chain is paid per scan, not per entry.)  Words between nwords and
            // the next multiple of 128 exist in the bitmap and are zero.
            const int nchunk = (nwords + 127) >> 7;
            const uint4 *bits4 = reinterpret_cast<const uint4 *>(bits);
            for (int c = wid; c < nchunk; c += NWARPS) {
                const uint4 w4 = bits4[c * 32 + lane];
                const int cnt = warp_sum((int)(__popc(w4.x) + __popc(w4.y) + __popc(w4.z) + __popc(w4.w)));
                if (lane == 0) s_blk[c] = (unsigned)cnt;
            }
            __syncthreads();
            {
                const unsigned v0 = tid < nchunk ? s_blk[tid] : 0u;       // nchunk <= 2 * BLOCK * 32 / 128 = BLOCK / 2
                unsigned excl, tot;
                Scan(scan_tmp).ExclusiveSum(v0, excl, tot);
                if (tid < nchunk) s_blk[tid] = excl;
                if (tid == 0) {
                    s_swtot = tot; s_next = 0;
                    if (split) atomicExch(part_cnt + idx, (int)tot + 1);      // published before the emit: the parts behind wait for it
                }
            }
            __syncthreads();
            if (split) {
                // this part's place in the row: the counts of the parts before it
                unsigned before = 0;
                for (int t = tid; t < part; t += BLOCK) {
                    const volatile int *slot = part_cnt + (idx - part + t);
                    int v = *slot;
                    while (v == 0) { __nanosleep(200); v = *slot; }
                    before += (unsigned)(v - 1);
                }
                if (before) atomicAdd(&s_base, before);
                __syncthreads();
                gs += s_base;
                n = (int)s_swtot;
                GP_ADD(35);
            }
            // cells + sorted columns of the super-window; the bitmap is left clean.  Chunks are drawn from a counter:
            // a chunk of hub columns (4096 entries) takes twenty times longer than a sparse one, and with a fixed
            // assignment the barrier behind this loop was the largest single stall of the kernel (ncu: 15 % of samples).
            const unsigned base_rank = s_carry;
            while (true) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_next, 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= nchunk) break;
                const uint4 w4 = bits4[c * 32 + lane];
                const unsigned wd[4] = {w4.x, w4.y, w4.z, w4.w};
                const unsigned nz = (wd[0] ? 1u : 0u) | (wd[1] ? 2u : 0u) | (wd[2] ? 4u : 0u) | (wd[3] ? 8u : 0u);
                if (!__any_sync(0xffffffffu, nz != 0u)) continue;
                const int cnt = __popc(wd[0]) + __popc(wd[1]) + __popc(wd[2]) + __popc(wd[3]);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                // which of a block's 32 words are non-zero (g_clear reads this): a block = 8 lanes x 4 words
                unsigned part = nz << (4 * (lane & 7));
                part |= __shfl_xor_sync(0xffffffffu, part, 1);
                part |= __shfl_xor_sync(0xffffffffu, part, 2);
                part |= __shfl_xor_sync(0xffffffffu, part, 4);
                if ((lane & 7) == 0 && part) wsum[(w0g >> 5) + c * 4 + (lane >> 3)] = part;
                if (nz) {
                    unsigned rank = base_rank + s_blk[c] + (unsigned)(incl - cnt);
                    const int w = c * 128 + lane * 4;
                    *reinterpret_cast<uint4 *>(bits + w) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        unsigned word = wd[x];
                        if (!word) continue;
                        wp[w0g + w + x] = make_uint2(word, rank);
                        const int col0 = c_lo + (w + x) * 32;
                        while (word) {
                            const int bit = __ffs(word) - 1;
                            word &= word - 1;
                            c_ci[gs + rank] = col0 + bit;
                            ++rank;
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) s_carry = base_rank + s_swtot;
            GP_ADD(36);
        }
        // 2. window boundaries (first column of every window of `win` ranks) and the split table of the accumulate pass
        const int W = (n + win - 1) / win;
        const bool bnd_in_smem = W - 1 <= G2_MAX_BND;
        if (W > 1 && bnd_in_smem) {
            __threadfence_block();
            __syncthreads();                                          // the emitted columns are read back by other threads
            for (int w = tid + 1; w < W; w += BLOCK) s_bnd[w] = __ldcg(c_ci + gs + (long long)w * win);
        }
        const int astride = (W + 1) | 1;
        const bool use_atbl = single_tile && W > 1 && bnd_in_smem && (long long)n_a * astride <= tbl_cap;
        if (use_atbl) {
            tile.rel[tid] = my_qb;
            tile.incl[tid] = my_len;
            __syncthreads();
            const int items = n_a * (W - 1);
            for (int it0 = tid; it0 < items; it0 += 4 * BLOCK) {
                boff b[4];
                int len[4], cc[4], lo[4], slot[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int it = it0 + u * BLOCK;
                    b[u] = 0; len[u] = 0; cc[u] = 0; slot[u] = -1;
                    if (it < items) {
                        const int e = it / (W - 1), k = it - e * (W - 1) + 1;
                        b[u] = tile.rel[e]; len[u] = tile.incl[e]; cc[u] = s_bnd[k]; slot[u] = e * astride + k;
                    }
                }
                lower_bound4(B, b, len, cc, lo);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (slot[u] >= 0) tbl[slot[u]] = lo[u];
            }
            if (tid < n_a) { tbl[tid * astride] = 0; tbl[tid * astride + W] = my_len; }
        }
        __syncthreads();
        const bool use_gatbl = long_cached && W > 1 && bnd_in_smem && (long long)n_a * (W + 3) <= gscr_cap;
        const bool cut = W > 1 || split;                              // the B rows must be restricted to the window's columns
        if (use_gatbl) long_table(W, [&](int k) { return s_bnd[k]; });
        GP_ADD(38);
        // 3. one window of `win` ranks at a time: accumulate in the shared-memory tile, then write it out
        for (int w = 0; w < W; ++w) {
            const int wbase = w * win;
            const int wn = min(win, n - wbase);
            int c_lo = row_lo, c_hi = row_hi;
            if (W > 1) {
                if (bnd_in_smem) { c_lo = w == 0 ? row_lo : s_bnd[w]; c_hi = w + 1 < W ? s_bnd[w + 1] : row_hi; }
                else { c_lo = w == 0 ? row_lo : __ldcg(c_ci + gs + wbase); c_hi = w + 1 < W ? __ldcg(c_ci + gs + wbase + win) : row_hi; }
            }
            // (the tile is all zero here: the emit pass left the bitmap clean -- the same bytes -- and every write-out
            //  below zeroes what it has read)
            GP_CNT(42, 1);
            auto add = [&](const boff (&q)[PB], const double (&av)[PB], unsigned valid) {
                int k[PB];
                double x[PB];
                uint2 cell[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    k[u] = -1; x[u] = 0.0;
                    if ((valid >> u) & 1u) { k[u] = __ldg(B.ci + q[u]); x[u] = av[u] * __ldg(B.v + q[u]); }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    if (k[u] < c_lo || k[u] >= c_hi) k[u] = -1;
                    cell[u] = k[u] >= 0 ? __ldcg(wp + (k[u] >> 5)) : make_uint2(0, 0);
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    if (k[u] < 0) continue;
                    const unsigned below = cell[u].x & ((1u << (k[u] & 31)) - 1u);
                    atomicAdd(&acc[(int)(cell[u].y + __popc(below)) - wbase], x[u]);
                }
            };
            if (single_tile) {
                boff qb = my_qb;
                int len = my_len;
                if (use_atbl) {
                    const int o0 = tid < n_a ? tbl[tid * astride + w] : 0, o1 = tid < n_a ? tbl[tid * astride + w + 1] : 0;
                    qb = my_qb + o0; len = o1 - o0;
                } else if (W > 1) gwin_restrict<BLOCK>(B, n_a, qb, len, tile, c_lo, c_hi);
                const int total = gwin_scan<true, BLOCK>(tile, qb, len, my_av);
                GP_ADD(38);
                if (total) gwin_run<true, BLOCK>(tile, total, add);
            } else {
                for (aoff base = pa; base < pe; base += BLOCK) {
                    const int total = use_gatbl ? long_tile(base, w, true)
                                      : split && long_cached && W == 1 ? cached_tile(base, true)
                                      : gwin_build<true, BLOCK>(A, B, base, pe, tile, cut ? c_lo : 0, cut ? c_hi : 0x7fffffff);
                    GP_ADD(38);
                    if (total) gwin_run<true, BLOCK>(tile, total, add);
                    __syncthreads();
                }
            }
            __syncthreads();
            GP_ADD(39);
            for (int t = tid; t < wn; t += BLOCK) { c_v[gs + wbase + t] = acc[t]; acc[t] = 0.0; }
            __syncthreads();
            GP_ADD(40);
        }
        // 4. leave the slot clean for the next row (the tile, hence the bitmap, already is); a part touched its own blocks only
        if (split) {
            const int b0 = (row_lo >> 5) >> 5, b1 = min(L.blocks, b0 + (split_words >> 5));
            for (int b = b0 + tid; b < b1; b += BLOCK) {
                const unsigned m = __ldcg(wsum + b);
                if (!m) continue;
                wsum[b] = 0;
                uint2 *cell = wp + (size_t)b * 32;
#pragma unroll
                for (int x = 0; x < 32; ++x)
                    if ((m >> x) & 1u) cell[x] = make_uint2(0, 0);
            }
        } else g_clear<BLOCK>(wp, summary, wsum, L);
        __syncthreads();
        GP_ADD(41);
    }
}

// the first rows of a bin's work-ordered list that are worth splitting: rows with at least `fixed_threshold` products, or
// (fixed_threshold == 0) with more than a quarter of one CTA's even share of the launch and at least `floor_products`
static __global__ void __launch_bounds__(1024) k_count_split_rows(const unsigned *__restrict__ keys_desc, int n, long long fixed_threshold,
                                                                  long long floor_products, int ctas, int *__restrict__ out)
{
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    unsigned long long threshold = (unsigned long long)fixed_threshold;
    if (fixed_threshold <= 0) {
        unsigned long long mine = 0;
        for (int t = threadIdx.x; t < n; t += 1024) mine += keys_desc[t];
        for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_sum, mine);
        __syncthreads();
        threshold = max((unsigned long long)floor_products, s_sum / (unsigned long long)(4 * max(ctas, 1)));
    }
    // keys are sorted by decreasing work: the rows at or above the threshold are a prefix
    int lo = 0, hi = n;
    if (threadIdx.x == 0) {
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((unsigned long long)keys_desc[mid] >= threshold) lo = mid + 1; else hi = mid; }
        *out = lo;
    }
}

// ---------------------------------------------------------------- windowed kernels (helpers are defined ahead of the L2 kernels, which share them)
template <class AV, class BV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sym_gwin(const int *__restrict__ rows, int nrows, int r0, AV A, BV B,
                                                    int *__restrict__ nnz_row, int *__restrict__ cursor, int ncols, int swords)
{
    typedef typename AV::off_t aoff;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *bits = reinterpret_cast<unsigned *>(smem_raw);      // swords words, all zero between rows
    __shared__ int s_row, s_cnt;
    __shared__ CtaTile<BV, BLOCK> tile;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int w = tid; w < swords; w += BLOCK) bits[w] = 0;
    const long long span = (long long)swords * 32;
    while (true) {
        if (tid == 0) { s_row = atomicAdd(cursor, 1); s_cnt = 0; }
        __syncthreads();
        const int idx = s_row;
        if (idx >= nrows) break;
        const int li = rows ? rows[idx] : idx;
        const int i = r0 + li;
        const aoff pa = A.begin(i), pe = A.end(i);
        int cnt = 0;
        GP_START();
        GP_CNT(24, 1);
        for (long long sw_lo = 0; sw_lo < ncols; sw_lo += span) {
            const int c_lo = (int)sw_lo;
            const int c_hi = sw_lo + span >= ncols ? 0x7fffffff : (int)(sw_lo + span);
            bool any = false;
            for (aoff base = pa; base < pe; base += BLOCK) {
                const int total = gwin_build<false, BLOCK>(A, B, base, pe, tile, c_lo, c_hi);
                GP_ADD(25);
                if (total) {
                    any = true;
                    gwin_run<false, BLOCK>(tile, total, [&](const typename BV::off_t (&q)[PB], const double (&)[PB], unsigned valid) {
                        int k[PB];
                        unsigned cur[PB];
#pragma unroll
                        for (int u = 0; u < PB; ++u) k[u] = (valid >> u) & 1u ? __ldg(B.ci + q[u]) - c_lo : -1;
#pragma unroll
                        for (int u = 0; u < PB; ++u) cur[u] = k[u] >= 0 ? bits[k[u] >> 5] : 0xffffffffu;
#pragma unroll
                        for (int u = 0; u < PB; ++u) {
                            const unsigned bit = 1u << (k[u] & 31);
                            if (!(cur[u] & bit)) atomicOr(&bits[k[u] >> 5], bit);
                        }
                    });
                }
                __syncthreads();
                GP_ADD(26);
            }
            if (any) {
                const int nw = (int)min((long long)swords, ((long long)ncols - sw_lo + 31) / 32);
                for (int w = tid; w < nw; w += BLOCK) {
                    const unsigned b = bits[w];
                    if (b) { cnt += __popc(b); bits[w] = 0; }
                }
                __syncthreads();
                GP_ADD(27);
            }
        }
        cnt = warp_sum(cnt);
        if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (tid == 0) nnz_row[li] = s_cnt;
    }
}

template <class AV, class BV, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_num_gwin(const int *__restrict__ rows, int nrows, int r0, AV A, BV B, OutMap out,
                                                    int *__restrict__ c_ci, double *__restrict__ c_v, int *__restrict__ cursor,
                                                    int ncols, int swords, int win, int tbl_cap)
{
    typedef typename AV::off_t aoff;
    typedef typename BV::off_t boff;
    constexpr int NWARPS = BLOCK / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2 *cells = reinterpret_cast<uint2 *>(smem_raw);                               // swords x {bits, rank}
    double *acc = reinterpret_cast<double *>(smem_raw + (size_t)swords * 8);         // win partial sums, zero between windows
    int *scol = reinterpret_cast<int *>(smem_raw + (size_t)swords * 8 + (size_t)win * 8);   // win sorted columns
    int *tbl = scol + win;                                                            // tbl_cap split points (see below)
    __shared__ int s_row;
    __shared__ int s_wtot[32], s_wbase[33];
    __shared__ CtaTile<BV, BLOCK> tile;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    for (int w = tid; w < swords; w += BLOCK) cells[w] = make_uint2(0u, 0u);
    for (int t = tid; t < win; t += BLOCK) acc[t] = 0.0;
    const long long span = (long long)swords * 32;
    const int nsw = (int)((ncols + span - 1) / span);
    const int tstride = (nsw + 1) | 1;                 // odd stride: thread e reads tbl[e * tstride + k] without bank conflicts
    while (true) {
        if (tid == 0) s_row = atomicAdd(cursor, 1);
        __syncthreads();
        const int idx = s_row;
        if (idx >= nrows) break;
        const int li = rows ? rows[idx] : idx;
        const int i = r0 + li;
        const long long gs = out.start(li);
        const aoff pa = A.begin(i), pe = A.end(i);
        const int n_a = (int)min((aoff)0x7fffffff, pe - pa);
        const bool single_tile = pe - pa <= (aoff)BLOCK;
        int emitted = 0;
        GP_START();
        GP_CNT(8, 1);
        if (!single_tile) GP_CNT(20, 1);
        if (use_tbl_probe(single_tile, nsw, n_a, tstride, tbl_cap)) GP_CNT(21, 1);
        // A rows of at most BLOCK entries (nearly all of them): thread e owns entry e for the whole row, and the
        // positions where its B row crosses the super-window boundaries are found once, all in parallel (one
        // latency chain per row instead of one per window): tbl[e][k] = first index of B row e with column >= k*span.
        boff my_qb = 0;
        int my_len = 0;
        const bool use_tbl = single_tile && nsw > 1 && (long long)n_a * tstride <= tbl_cap;
        if (single_tile) { double unused; gwin_load<false>(A, B, pa, pe, my_qb, my_len, unused); }
        // (the A value is re-read where a tile is built: one L1/L2 hit instead of two registers held across the row)
        auto my_av = [&]() { return tid < n_a ? __ldg(A.v + pa + tid) : 0.0; };
        if (use_tbl) {
            tile.rel[tid] = my_qb;
            tile.incl[tid] = my_len;
            __syncthreads();
            const int items = n_a * (nsw - 1);
            for (int it = tid; it < items; it += BLOCK) {
                const int e = it / (nsw - 1), k = it - e * (nsw - 1) + 1;
                const boff b = tile.rel[e];
                const int l = tile.incl[e];
                const int c = (int)(k * span);
                int lo = 0, hi = l;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(B.ci + b + mid) < c) lo = mid + 1; else hi = mid; }
                tbl[e * tstride + k] = lo;
            }
            if (tid < n_a) { tbl[tid * tstride] = 0; tbl[tid * tstride + nsw] = my_len; }
            __syncthreads();
            GP_ADD(13);
        }
        for (int k = 0; k < nsw; ++k) {
            const long long sw_lo = k * span;
            const int c_lo = (int)sw_lo;
            const int c_hi = k + 1 == nsw ? 0x7fffffff : (int)(sw_lo + span);
            const int nw = (int)min((long long)swords, ((long long)ncols - sw_lo + 31) / 32);
            // this thread's B-row segment inside the super-window (single_tile only)
            auto segment = [&](boff &qb, int &len) {
                qb = my_qb; len = my_len;
                if (use_tbl) {
                    const int o0 = tid < n_a ? tbl[tid * tstride + k] : 0, o1 = tid < n_a ? tbl[tid * tstride + k + 1] : 0;
                    qb = my_qb + o0;
                    len = o1 - o0;
                }
            };
            // 1. mark the columns of this super-window (the tile of a short A row is kept for step 3)
            bool any = false;
            int tile_total = 0;
            if (single_tile) {
                boff qb;
                int len;
                segment(qb, len);
                if (!use_tbl) gwin_restrict<BLOCK>(B, n_a, qb, len, tile, c_lo, c_hi);
                tile_total = gwin_scan<true, BLOCK>(tile, qb, len, my_av());
            }
            auto mark = [&](const boff (&q)[PB], const double (&)[PB], unsigned valid) {
                int kk[PB];
                unsigned cur[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) kk[u] = (valid >> u) & 1u ? __ldg(B.ci + q[u]) - c_lo : -1;
#pragma unroll
                for (int u = 0; u < PB; ++u) cur[u] = kk[u] >= 0 ? cells[kk[u] >> 5].x : 0xffffffffu;
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const unsigned bit = 1u << (kk[u] & 31);
                    if (!(cur[u] & bit)) atomicOr(&cells[kk[u] >> 5].x, bit);
                }
            };
            if (single_tile) {
                GP_ADD(0);
                if (tile_total) { any = true; gwin_run<false, BLOCK>(tile, tile_total, mark); }
                __syncthreads();
                GP_ADD(1);
            } else {
                for (aoff base = pa; base < pe; base += BLOCK) {
                    const int total = gwin_build<false, BLOCK>(A, B, base, pe, tile, c_lo, c_hi);
                    GP_ADD(17);
                    GP_CNT(18, 1);
                    if (total) { any = true; gwin_run<false, BLOCK>(tile, total, mark); }
                    __syncthreads();
                    GP_ADD(19);
                }
            }
            if (!any) { GP_CNT(12, 1); continue; }
            GP_CNT(9, 1);
            // 2. rank of every word's first bit: each warp scans a contiguous chunk (four cells per lane and trip,
            //    128-bit shared-memory accesses), then the chunk bases are added.  Cells between nw and the next
            //    multiple of 4 are empty and may be rewritten.
            const int chunk = ((nw + NWARPS - 1) / NWARPS + 127) & ~127;
            const int nw4 = (nw + 3) & ~3;
            const int w_begin = min(nw4, wid * chunk), w_end = min(nw4, w_begin + chunk);
            {
                int carry = 0;
                for (int w0 = w_begin; w0 < w_end; w0 += 128) {
                    const int w = w0 + lane * 4;
                    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
                    if (w < w_end) { a = *reinterpret_cast<const uint4 *>(cells + w); b = *reinterpret_cast<const uint4 *>(cells + w + 2); }
                    const int c0 = __popc(a.x), c1 = __popc(a.z), c2 = __popc(b.x), c3 = __popc(b.z);
                    const int c = c0 + c1 + c2 + c3;
                    int incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                    if (w < w_end) {
                        const unsigned y0 = (unsigned)(carry + incl - c);
                        a.y = y0; a.w = y0 + c0; b.y = y0 + c0 + c1; b.w = y0 + c0 + c1 + c2;
                        *reinterpret_cast<uint4 *>(cells + w) = a;
                        *reinterpret_cast<uint4 *>(cells + w + 2) = b;
                    }
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) s_wtot[wid] = carry;
            }
            __syncthreads();
            if (wid == 0) {
                const int v = lane < NWARPS ? s_wtot[lane] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
                s_wbase[lane] = incl - v;
                if (lane == 31) s_wbase[32] = incl;
            }
            __syncthreads();
            {
                const unsigned wb = (unsigned)s_wbase[wid];
                if (wb)
                    for (int w = w_begin + lane * 2; w < w_end; w += 64) {
                        uint4 a = *reinterpret_cast<const uint4 *>(cells + w);
                        a.y += wb; a.w += wb;
                        *reinterpret_cast<uint4 *>(cells + w) = a;
                    }
            }
            const int sw_total = s_wbase[32];
            __syncthreads();
            GP_ADD(2);
            // 3. rank windows: words [ws, we) hold at most `win` entries
            int ws = 0;
            while (ws < nw) {
                const int base_rank = (int)cells[ws].y;
                if (sw_total - base_rank == 0) break;
                int we = nw;
                if (sw_total - base_rank > win) {
                    int lo = ws + 1, hi = nw - 1;          // a word holds at most 32 <= win entries: ws + 1 is always valid
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if ((int)cells[mid].y - base_rank <= win) lo = mid; else hi = mid - 1;
                    }
                    we = lo;
                }
                const int wn = (we == nw ? sw_total : (int)cells[we].y) - base_rank;
                // sorted columns of the window, rank-parallel: thread t produces entries [t*G, (t+1)*G) -- it finds the
                // word that holds its first rank by binary search over the ranks, then walks the bits.  (One thread per
                // word leaves the warp that draws the hub columns, 32 bits in every word, far behind the others.)
                {
                    const int G = (wn + BLOCK - 1) / BLOCK;
                    int r = tid * G;
                    const int r_end = min(wn, r + G);
                    if (r < r_end) {
                        const unsigned target = (unsigned)(base_rank + r);
                        int lo = ws, hi = we - 1;              // last word whose first rank is <= target
                        while (lo < hi) {
                            const int mid = (lo + hi + 1) >> 1;
                            if (cells[mid].y <= target) lo = mid; else hi = mid - 1;
                        }
                        int w = lo;
                        const uint2 c = cells[w];
                        unsigned bits = c.x;
                        for (unsigned skip = target - c.y; skip > 0; --skip) bits &= bits - 1;
                        while (r < r_end) {
                            if (!bits) { ++w; bits = cells[w].x; continue; }
                            const int q = __ffs(bits) - 1;
                            bits &= bits - 1;
                            scol[r++] = c_lo + w * 32 + q;
                        }
                    }
                }
                GP_ADD(3);
                GP_CNT(10, 1);
                // products of the window
                const bool whole = ws == 0 && we == nw;
                const int a_lo = ws == 0 ? c_lo : c_lo + ws * 32;
                const int a_hi = we == nw ? c_hi : c_lo + we * 32;
                auto add = [&](const boff (&q)[PB], const double (&av)[PB], unsigned valid) {
                    int kk[PB];
                    double x[PB];
                    uint2 cell[PB];
#pragma unroll
                    for (int u = 0; u < PB; ++u) {
                        kk[u] = -1; x[u] = 0.0;
                        if ((valid >> u) & 1u) { kk[u] = __ldg(B.ci + q[u]) - c_lo; x[u] = av[u] * __ldg(B.v + q[u]); }
                    }
#pragma unroll
                    for (int u = 0; u < PB; ++u) cell[u] = kk[u] >= 0 ? cells[kk[u] >> 5] : make_uint2(0u, 0u);
#pragma unroll
                    for (int u = 0; u < PB; ++u) {
                        if (kk[u] < 0) continue;
                        const int slot = (int)cell[u].y - base_rank + __popc(cell[u].x & ((1u << (kk[u] & 31)) - 1u));
                        if ((unsigned)slot < (unsigned)wn) atomicAdd(&acc[slot], x[u]);
                    }
                };
                if (single_tile) {
                    if (whole) {
                        GP_CNT(11, 1);
                        gwin_run<true, BLOCK>(tile, tile_total, add);          // the tile of the mark pass is still in place
                    } else {
                        // narrow this thread's segment to the rank window (short searches when the split table gave
                        // the super-window's part of the B row, which then also bounds the window on the outside)
                        boff qb;
                        int len;
                        segment(qb, len);
                        if (use_tbl) gwin_restrict<BLOCK>(B, n_a, qb, len, tile, ws == 0 ? 0 : a_lo, we == nw ? 0x7fffffff : a_hi);
                        else gwin_restrict<BLOCK>(B, n_a, qb, len, tile, a_lo, a_hi);
                        const int total = gwin_scan<true, BLOCK>(tile, qb, len, my_av());
                        GP_ADD(4);
                        if (total) gwin_run<true, BLOCK>(tile, total, add);
                    }
                } else {
                    for (aoff base = pa; base < pe; base += BLOCK) {
                        const int total = gwin_build<true, BLOCK>(A, B, base, pe, tile, a_lo, a_hi);
                        GP_ADD(14);
                        GP_CNT(15, 1);
                        if (total) gwin_run<true, BLOCK>(tile, total, add);
                        __syncthreads();
                        GP_ADD(16);
                    }
                }
                __syncthreads();
                GP_ADD(5);
                for (int t = tid; t < wn; t += BLOCK) {
                    c_ci[gs + emitted + t] = scol[t];
                    c_v[gs + emitted + t] = acc[t];
                    acc[t] = 0.0;
                }
                emitted += wn;
                ws = we;
                __syncthreads();
                GP_ADD(6);
            }
            // 4. leave the cells clean for the next super-window / row
            for (int w = tid * 2; w < nw; w += BLOCK * 2) *reinterpret_cast<uint4 *>(cells + w) = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
            GP_ADD(7);
        }
    }
}

// ---------------------------------------------------------------- consumers / small utilities
// order-independent structure hash + checksum of a batch of C: sum over entries of mix64(row<<32 | col).
// A warp takes 32 * PER consecutive entries and reads them lane-strided (coalesced); the rows they belong to lie
// between the rows of the chunk's first and last entry (two searches per warp), so the per-entry search runs over
// that span only -- zero steps for the long rows of a power-law result, five or six for a stencil's.
static __global__ void __launch_bounds__(256) k_consume(int nrows, int row_base, const long long *__restrict__ rp, long long rp0,
                                                 const int *__restrict__ ci, const double *__restrict__ v,
                                                 unsigned long long *__restrict__ hash_out, double *__restrict__ sum_out)
{
    constexpr int PER = 16;
    const long long n = rp[nrows] - rp0;
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long c0 = warp_id * (32 * PER);
    unsigned long long h = 0;
    double s = 0.0;
    if (c0 < n) {
        const long long c1 = min(c0 + 32 * PER, n);
        auto row_of = [&](long long e, int lo, int hi) {        // last row in [lo, hi] with rp[row] - rp0 <= e
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (rp[mid] - rp0 <= e) lo = mid; else hi = mid - 1; }
            return lo;
        };
        const int r_lo = row_of(c0, 0, nrows - 1), r_hi = row_of(c1 - 1, r_lo, nrows - 1);
#pragma unroll 4
        for (long long e = c0 + lane; e < c1; e += 32) {
            const int row = r_lo == r_hi ? r_lo : row_of(e, r_lo, r_hi);
            h += mix64(((unsigned long long)(unsigned)(row_base + row) << 32) | (unsigned)__ldg(ci + e));
            s += __ldg(v + e);
        }
    }
    typedef cub::BlockReduce<unsigned long long, 256> RH;
    typedef cub::BlockReduce<double, 256> RS;
    __shared__ typename RH::TempStorage th;
    __shared__ typename RS::TempStorage tsm;
    h = RH(th).Sum(h);
    s = RS(tsm).Sum(s);
    if (threadIdx.x == 0) { if (h) atomicAdd(hash_out, h); atomicAdd(sum_out, s); }
}

static __global__ void k_copy_counts(int n, const int *__restrict__ src, int *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// greedy batch boundaries: largest row ranges whose nnz fits `cap` entries
static __global__ void k_batch_bounds(int nrows, const long long *__restrict__ rp, long long cap, int max_batches,
                               int *__restrict__ bounds /*max_batches+1*/, int *__restrict__ nb)
{
    if (threadIdx.x || blockIdx.x) return;
    int b = 0, start = 0;
    bounds[0] = 0;
    while (start < nrows && b < max_batches) {
        long long limit = rp[start] + cap;
        int lo = start + 1, hi = nrows;          // at least one row per batch
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (rp[mid] <= limit) lo = mid; else hi = mid - 1; }
        start = lo;
        bounds[++b] = start;
    }
    *nb = (start >= nrows) ? b : -1;
}

}  // inline namespace IAS_TU
}  // namespace ias
