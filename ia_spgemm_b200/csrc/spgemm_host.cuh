// spgemm_host.cuh -- host orchestration of the Gustavson pipeline for one row range of C.
// Templated on the operand views so that the CSR, ELL and COO entry points share it.
#pragma once
#include <stdlib.h>

#include <algorithm>

#include <thrust/iterator/transform_iterator.h>

#include "spgemm_kernels.cuh"

namespace ias {
// Every translation unit that includes the pipeline gets its OWN copy of the kernels (IAS_TU is defined by the .cu file):
// identical template instantiations in two units would otherwise be merged by the linker, and the phase-clock build
// (make prof) would read the counters of the copy that lost.
inline namespace IAS_TU {

constexpr int TINY_BLOCK = 128;
constexpr int G_BLOCK = 512;

// row lists grouped by bin: list == nullptr means "identity" (every row of the range is in `only_bin`)
struct BinLists {
    DBuf<int> list;
    DBuf<unsigned char> keys_sorted;
    DBuf<int> iota;
    DBuf<char> tmp;
    long long count[8] = {0};
    long long offset[8] = {0};
    int only_bin = -1;
    const int *rows_of(int bin) const { return only_bin >= 0 ? nullptr : list.p + offset[bin]; }
};

inline int build_bin_lists(int nrows, const unsigned char *bin_dev, const long long *hist, BinLists &bl)
{
    bl.only_bin = -1;
    long long run = 0;
    for (int b = 0; b < NBINS; ++b) {
        bl.count[b] = hist[b];
        bl.offset[b] = run;
        run += hist[b];
        if (hist[b] == nrows) bl.only_bin = b;
    }
    if (bl.only_bin >= 0 || nrows == 0) return IAS_OK;
    IAS_TRY(bl.list.alloc(nrows));
    IAS_TRY(bl.keys_sorted.alloc(nrows));
    IAS_TRY(bl.iota.alloc(nrows));
    IAS_LAUNCH(k_iota, grid_for(nrows, 256), 256, 0, nrows, bl.iota.p);
    size_t tb = 0;
    IAS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, bin_dev, bl.keys_sorted.p, bl.iota.p, bl.list.p, nrows, 0, 3, ctx().stream));
    IAS_TRY(bl.tmp.alloc(tb));
    IAS_CUDA(cub::DeviceRadixSort::SortPairs(bl.tmp.p, tb, bin_dev, bl.keys_sorted.p, bl.iota.p, bl.list.p, nrows, 0, 3, ctx().stream));
    ctx().launches += 3;       // histogram + onesweep passes of the 3-bit stable sort
    return IAS_OK;
}

// Rows of a bin ordered by decreasing work (longest processing time first): the persistent kernels hand rows out through
// an atomic cursor, so the tail of a launch is bounded by the smallest rows instead of whatever hub row came last.
struct WorkOrder {
    DBuf<unsigned> keys, keys_sorted;
    DBuf<int> vals, list;
    DBuf<char> tmp;
};
inline int order_by_work(const int *list_in, int n, const int *ub, WorkOrder &wo, const int **out)
{
    *out = list_in;
    if (n < 2 || ctx().tune.g_lpt == 0) return IAS_OK;
    IAS_TRY(wo.keys.alloc(n));
    IAS_TRY(wo.keys_sorted.alloc(n));
    IAS_TRY(wo.vals.alloc(n));
    IAS_TRY(wo.list.alloc(n));
    IAS_LAUNCH(k_work_keys, grid_for(n, 256), 256, 0, n, list_in, ub, wo.keys.p, wo.vals.p);
    size_t tb = 0;
    IAS_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, wo.keys.p, wo.keys_sorted.p, wo.vals.p, wo.list.p, n, 0, 32, ctx().stream));
    IAS_TRY(wo.tmp.alloc(tb));
    IAS_CUDA(cub::DeviceRadixSort::SortPairsDescending(wo.tmp.p, tb, wo.keys.p, wo.keys_sorted.p, wo.vals.p, wo.list.p, n, 0, 32, ctx().stream));
    ctx().launches += 2;
    *out = wo.list.p;
    return IAS_OK;
}

template <class K>
inline int opt_in_smem(K kernel, size_t bytes)
{
    if (bytes > 32 * 1024) IAS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return IAS_OK;
}

// state of one row range [r0, r1) after analysis + symbolic
struct RangeWork {
    int r0 = 0, nrows = 0;
    DBuf<int> ub;                  // nrows
    DBuf<int> nnz_row;             // nrows + 1 (last = 0, so that the scan yields the total)
    DBuf<unsigned char> bin;       // nrows
    DBuf<unsigned long long> hist; // 16 counters
    DBuf<unsigned> gwork;          // global-row workspace slots
    DBuf<int> gscr;                // k_num_global2: split-table scratch of the rows with more than 1024 A entries
    size_t gscr_ints = 0;
    DBuf<int> split;               // k_num_global2: [0] = rows cut into parts, then one published count per part
    size_t split_ints = 0;
    DBuf<int> cursor;
    int gslots = 0;
    long long products = 0;
    long long sym_hist[8] = {0};
    long long num_hist[8] = {0};
    int b_canonical = 0;           // every B row strictly increasing in column (enables the merge kernels)
    int max_warp_ub = 0;           // largest upper bound among warp-bin rows (sizes the register sort)
    int max_tiny_na = 0;           // longest A row / largest upper bound among tiny rows
    int max_tiny_ub = 0;
    bool tiny_only = false;        // every non-empty row is tiny: numeric bins == symbolic bins
    DBuf<int> tiny_list;           // kept row list of the tiny bin (tiny_only with empty rows)
    bool tiny_identity = false;
    bool sym_timed[8] = {false};   // which bin kernels were launched (their ev_bin pairs are pending)
    bool num_timed[8] = {false};
    double ms_bin_sym[8] = {0};
    double ms_bin_num[8] = {0};
};

// events around one bin kernel: slot = bin (symbolic) or 8 + bin (numeric)
#define IAS_BIN_BEGIN(slot) IAS_CUDA(cudaEventRecord(ctx().ev_bin[2 * (slot)], ctx().stream))
#define IAS_BIN_END(slot) IAS_CUDA(cudaEventRecord(ctx().ev_bin[2 * (slot) + 1], ctx().stream))

// after a stream synchronize: fold the pending per-bin event pairs into rw.ms_bin_*
inline void collect_bin_times(RangeWork &rw)
{
    for (int b = 0; b < 8; ++b) {
        float ms = 0.f;
        if (rw.sym_timed[b] && cudaEventElapsedTime(&ms, ctx().ev_bin[2 * b], ctx().ev_bin[2 * b + 1]) == cudaSuccess) rw.ms_bin_sym[b] += ms;
        if (rw.num_timed[b] && cudaEventElapsedTime(&ms, ctx().ev_bin[2 * (8 + b)], ctx().ev_bin[2 * (8 + b) + 1]) == cudaSuccess) rw.ms_bin_num[b] += ms;
        rw.sym_timed[b] = rw.num_timed[b] = false;
    }
    cudaGetLastError();
}

inline int read_hist(RangeWork &rw, int n, long long *out)
{
    Ctx &c = ctx();
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars, rw.hist.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, c.stream));
    HostTrace ht("cudaStreamSynchronize after the analyze kernels", 0);
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    ht.done();
    for (int i = 0; i < n; ++i) out[i] = c.h_scalars[i];
    return IAS_OK;
}

// Global rows run as persistent CTAs, one workspace slot each.  Two 512-thread CTAs per SM while all slots
// together stay around the L2 size; one 1024-thread CTA per SM when the slots are large (R-MAT scale >= 22:
// 1 MB per slot), so that the bitmap / rank cells keep hitting in L2.
inline bool g_wide(int ncols)
{
    GLayout L = GLayout::make(ncols);
    return (double)L.slot_words * 4.0 * 2.0 * ctx().sm_count > 160e6;
}

inline int ensure_gwork(RangeWork &rw, int ncols, long long nrows_g, int per_sm = 0 /* 0: by slot size */)
{
    if (nrows_g <= 0) return IAS_OK;
    GLayout L = GLayout::make(ncols);
    if (per_sm == 0) per_sm = g_wide(ncols) ? 1 : 2;
    int slots = (int)std::min<long long>(nrows_g, (long long)per_sm * ctx().sm_count);
    if (rw.gslots >= slots && rw.gwork.p) return IAS_OK;
    IAS_TRY(rw.gwork.alloc((size_t)slots * L.slot_words));
    IAS_CUDA(cudaMemsetAsync(rw.gwork.p, 0, (size_t)slots * L.slot_words * sizeof(unsigned), ctx().stream));
    if (!rw.cursor.p) IAS_TRY(rw.cursor.alloc(1));
    rw.gslots = slots;
    return IAS_OK;
}

// Windowed shared-memory kernels for global rows (k_sym_gwin / k_num_gwin): one persistent 1024-thread CTA per SM
// (more when the bitmap is small), dynamic shared memory = everything the static tile leaves.
constexpr int GW_BLOCK = 1024;

inline bool use_gwin(const RangeWork &rw) { return rw.b_canonical && ctx().tune.global_rows_smem != 0; }

template <class K>
inline int gwin_dyn_max(K kernel, size_t *out)
{
    cudaFuncAttributes fa;
    IAS_CUDA(cudaFuncGetAttributes(&fa, kernel));
    size_t lim = ctx().smem_optin;
    if (fa.sharedSizeBytes + 4096 > lim) return fail(IAS_E_CUDA, "windowed global-row kernel: static shared memory %zu leaves no room", fa.sharedSizeBytes);
    size_t dyn = lim - fa.sharedSizeBytes - 256;
    if (ctx().tune.gwin_smem_kb > 0) dyn = std::min(dyn, (size_t)ctx().tune.gwin_smem_kb * 1024);
    *out = dyn;
    return IAS_OK;
}

template <class K>
inline int gwin_grid(K kernel, size_t smem, long long nrows_g, int *grid)
{
    int occ = 1;
    IAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, GW_BLOCK, smem));
    if (occ < 1) return fail(IAS_E_CUDA, "windowed global-row kernel does not fit an SM with %zu bytes of shared memory", smem);
    *grid = (int)std::min<long long>(nrows_g, (long long)occ * ctx().sm_count);
    return IAS_OK;
}

#ifdef IAS_GWIN_PROFILE
// phase clocks of the windowed kernels since the last dump (stderr), in SM clocks summed over CTAs
static void gwin_profile_dump(const char *what)
{
    unsigned long long h[64];
    cudaStreamSynchronize(ctx().stream);
    if (cudaMemcpyFromSymbol(h, g_gwin_prof, sizeof h) != cudaSuccess) return;
    fprintf(stderr, "[gwin %s]", what);
    for (int i = 0; i < 64; ++i) if (h[i]) fprintf(stderr, " %d:%llu", i, h[i]);
    fprintf(stderr, "\n");
    memset(h, 0, sizeof h);
    cudaMemcpyToSymbol(g_gwin_prof, h, sizeof h);
}
#define GWIN_PROFILE_DUMP(what) gwin_profile_dump(what)
#else
#define GWIN_PROFILE_DUMP(what) do {} while (0)
#endif

inline int words_of(int ncols) { return (int)((((long long)ncols + 31) / 32 + 31) / 32 * 32); }

// numeric: every (row, super-window) pair costs a few passes over the window's cells whatever it holds, so the
// windowed kernel wins while rows need few super-windows (R-MAT scale <= 20 at the default 8192 words) and loses
// to the L2 bitmap kernel beyond (measured at scale 22: 16 super-windows per row, 4.6 s against 3.7 s)
inline bool gwin_numeric_pays(int ncols)
{
    long long sw = ctx().tune.gwin_swords > 0 ? std::max<long long>(32, (ctx().tune.gwin_swords + 31) & ~31LL) : 8192;
    long long nsw = (words_of(ncols) + sw - 1) / sw;
    return ctx().tune.gwin_max_sw == 0 || nsw <= ctx().tune.gwin_max_sw;
}

// analysis + symbolic over [r0, r1): fills rw.ub, rw.nnz_row (exact nnz(C_i)), products
template <class AV, class BV>
int symbolic_range(const AV &A, const BV &B, int r0, int r1, int ncols_b, double avg_a_row, RangeWork &rw, IasSpgemmStats *st,
                   bool b_is_a = false, int b_rows = 0, long long b_nnz = -1)
{
    Ctx &c = ctx();
    int nrows = r1 - r0;
    rw.r0 = r0; rw.nrows = nrows;
    IAS_TRY(rw.ub.alloc(nrows));
    IAS_TRY(rw.nnz_row.alloc((size_t)nrows + 1));
    IAS_TRY(rw.bin.alloc(nrows));
    IAS_TRY(rw.hist.alloc(16));
    IAS_CUDA(cudaMemsetAsync(rw.hist.p, 0, 16 * sizeof(unsigned long long), c.stream));
    IAS_CUDA(cudaMemsetAsync(rw.nnz_row.p, 0, sizeof(int) * ((size_t)nrows + 1), c.stream));
    if (nrows == 0) return IAS_OK;
    {
        DBuf<int> long_list, long_count;
        IAS_TRY(long_list.alloc(nrows));
        IAS_TRY(long_count.alloc(1));
        IAS_CUDA(cudaMemsetAsync(long_count.p, 0, sizeof(int), c.stream));
        IAS_LAUNCH((k_row_ub_thread<AV, BV>), grid_for(nrows, 256), 256, 0, nrows, r0, A, B, rw.ub.p, rw.bin.p, rw.hist.p, long_list.p, long_count.p, rw.nnz_row.p);
        // always launched: the kernel reads *long_count on the device and leaves at once when no row was deferred
        // (a host-side guess from the operand's average row length missed long rows inside small row blocks)
        (void)avg_a_row;
        IAS_LAUNCH((k_row_ub_long<AV, BV>), c.sm_count * 8, 256, 0, long_list.p, long_count.p, r0, A, B, rw.ub.p, rw.bin.p, rw.hist.p);
    }
    // canonical B: the analyze kernel has just checked the rows of A it walked; that covers B only when B is A
    // and the range is the whole matrix, otherwise B gets its own pass (4 B per entry of B)
    bool covered = b_is_a && r0 == 0 && r1 == b_rows;
    // Opt-in ("trust_operand_cache"): B's canonical flag is remembered per operand (pointers + shape), so repeated
    // row-block multiplies against the same B (multi-GPU ranks, callers that stream blocks themselves) pay the
    // 4 B/entry pass once.  Off by default: caller-owned memory can be freed and re-allocated at the same address
    // with the same shape (caching allocators do exactly that), and a stale flag would go unnoticed.
    bool cached = c.tune.trust_operand_cache != 0 && !covered && b_nnz >= 0 && c.canon_ci == (const void *)B.ci &&
                  c.canon_rp == (const void *)B.rp_base() && c.canon_rows == b_rows && c.canon_nnz == b_nnz;
    if (!covered && !cached && b_rows > 0) {
        IAS_CUDA(cudaMemsetAsync(rw.hist.p + NBINS + 1, 0, sizeof(unsigned long long), c.stream));
        IAS_LAUNCH((k_rows_canonical<BV>), grid_for(b_rows, 256), 256, 0, b_rows, B, rw.hist.p + NBINS + 1);
    }
    long long h[NBINS + 5];
    IAS_TRY(read_hist(rw, NBINS + 5, h));
    rw.products = h[NBINS];
    rw.b_canonical = (b_rows > 0 && h[NBINS + 1] == 0) ? 1 : 0;
    if (cached) rw.b_canonical = c.canon_flag;
    else if (c.tune.trust_operand_cache != 0 && b_nnz >= 0 && b_rows > 0) {
        c.canon_ci = (const void *)B.ci; c.canon_rp = (const void *)B.rp_base(); c.canon_rows = b_rows; c.canon_nnz = b_nnz;
        c.canon_flag = rw.b_canonical;
    }
    rw.max_tiny_na = (int)h[NBINS + 2];
    rw.max_tiny_ub = (int)h[NBINS + 3];
    rw.max_warp_ub = (int)h[NBINS + 4];
    // every non-empty row is tiny: the numeric bins equal the symbolic ones (ub <= 32 decides), no re-classification
    rw.tiny_only = h[BIN_T] > 0 && h[BIN_T] + h[BIN_EMPTY] == nrows;
    for (int b = 0; b < NBINS; ++b) rw.sym_hist[b] = h[b];
    IAS_CUDA(cudaEventRecord(c.ev[1], c.stream));
    if (st) {
        st->products = rw.products;
        for (int b = 0; b < 8; ++b) st->sym_bin_rows[b] = b < NBINS ? rw.sym_hist[b] : 0;
    }
    // hist[12] = largest nnz(C_i) among tiny rows (sizes k_num_tiny's staging); the analyze kernel has already
    // counted the tiny rows optimistically: those counts stand if B is canonical and no tiny row exceeds 8 entries
    const bool tiny_counted = rw.b_canonical && rw.max_tiny_na <= 8;
    BinLists bl;
    IAS_TRY(build_bin_lists(nrows, rw.bin.p, h, bl));
    if (rw.tiny_only) {
        rw.tiny_identity = bl.only_bin == BIN_T;
        if (!rw.tiny_identity) {                  // tiny + empty rows: keep the tiny bin's row list for the numeric pass
            IAS_TRY(rw.tiny_list.alloc((size_t)bl.count[BIN_T]));
            IAS_CUDA(cudaMemcpyAsync(rw.tiny_list.p, bl.rows_of(BIN_T), sizeof(int) * (size_t)bl.count[BIN_T], cudaMemcpyDeviceToDevice, c.stream));
        }
    }
    if (bl.count[BIN_T] && !tiny_counted) {
        IAS_BIN_BEGIN(BIN_T);
        int n = (int)bl.count[BIN_T];
        int merge = rw.max_tiny_na <= 4 ? 4 : rw.max_tiny_na <= 6 ? 6 : 8;
        const int *rl = bl.rows_of(BIN_T);
        if (merge == 4) IAS_LAUNCH((k_sym_tiny<AV, BV, TINY_BLOCK, 4>), grid_for(n, TINY_BLOCK), TINY_BLOCK, 0, rl, n, r0, A, B, rw.nnz_row.p, rw.b_canonical, rw.hist.p + 12);
        else if (merge == 6) IAS_LAUNCH((k_sym_tiny<AV, BV, TINY_BLOCK, 6>), grid_for(n, TINY_BLOCK), TINY_BLOCK, 0, rl, n, r0, A, B, rw.nnz_row.p, rw.b_canonical, rw.hist.p + 12);
        else IAS_LAUNCH((k_sym_tiny<AV, BV, TINY_BLOCK, 8>), grid_for(n, TINY_BLOCK), TINY_BLOCK, 0, rl, n, r0, A, B, rw.nnz_row.p, rw.b_canonical, rw.hist.p + 12);
        IAS_BIN_END(BIN_T);
        rw.sym_timed[BIN_T] = true;
    }
    if (bl.count[BIN_W]) {
        IAS_BIN_BEGIN(BIN_W);
        int n = (int)bl.count[BIN_W];
        // key table sized from the bin's largest ub (load <= 50 %): 256 / 512 / 1024 slots per warp
        int mub = std::max(rw.max_warp_ub, T_MAX + 1);
        if (mub <= 128) {
            auto k = k_sym_hash<AV, BV, 32, 256, 256>;
            IAS_LAUNCH(k, grid_for(n, 8), 256, (size_t)8 * 256 * sizeof(int), bl.rows_of(BIN_W), n, r0, A, B, rw.nnz_row.p);
        } else if (mub <= 256) {
            auto k = k_sym_hash<AV, BV, 32, 256, 512>;
            IAS_LAUNCH(k, grid_for(n, 8), 256, (size_t)8 * 512 * sizeof(int), bl.rows_of(BIN_W), n, r0, A, B, rw.nnz_row.p);
        } else {
            auto k = k_sym_hash<AV, BV, 32, 256, SYM_W_TSIZE>;
            size_t sm = (size_t)8 * SYM_W_TSIZE * sizeof(int);
            IAS_TRY(opt_in_smem(k, sm));
            IAS_LAUNCH(k, grid_for(n, 8), 256, sm, bl.rows_of(BIN_W), n, r0, A, B, rw.nnz_row.p);
        }
        IAS_BIN_END(BIN_W);
        rw.sym_timed[BIN_W] = true;
    }
    if (bl.count[BIN_B1]) {
        IAS_BIN_BEGIN(BIN_B1);
        int n = (int)bl.count[BIN_B1];
        auto k = k_sym_hash<AV, BV, 256, 256, SYM_B1_TSIZE>;
        size_t sm = (size_t)SYM_B1_TSIZE * sizeof(int);
        IAS_TRY(opt_in_smem(k, sm));
        IAS_LAUNCH(k, n, 256, sm, bl.rows_of(BIN_B1), n, r0, A, B, rw.nnz_row.p);
        IAS_BIN_END(BIN_B1);
        rw.sym_timed[BIN_B1] = true;
    }
    if (bl.count[BIN_B2]) {
        IAS_BIN_BEGIN(BIN_B2);
        int n = (int)bl.count[BIN_B2];
        auto k = k_sym_hash<AV, BV, 1024, 1024, SYM_B2_TSIZE>;
        size_t sm = (size_t)SYM_B2_TSIZE * sizeof(int);
        IAS_TRY(opt_in_smem(k, sm));
        IAS_LAUNCH(k, n, 1024, sm, bl.rows_of(BIN_B2), n, r0, A, B, rw.nnz_row.p);
        IAS_BIN_END(BIN_B2);
        rw.sym_timed[BIN_B2] = true;
    }
    if (bl.count[BIN_G] && use_gwin(rw)) {
        IAS_BIN_BEGIN(BIN_G);
        int n = (int)bl.count[BIN_G];
        auto k = k_sym_gwin<AV, BV, GW_BLOCK>;
        size_t dyn = 0;
        IAS_TRY(gwin_dyn_max(k, &dyn));
        int swords = std::min(words_of(ncols_b), (int)(dyn / 4) & ~31);
        if (swords < 32) return fail(IAS_E_ARG, "gwin_smem_kb leaves %zu bytes of shared memory: too little for a bitmap window", dyn);
        if (c.tune.gwin_sym_swords > 0) swords = std::min<long long>(swords, std::max<long long>(32, (c.tune.gwin_sym_swords + 31) & ~31LL));
        size_t sm = (size_t)swords * 4;
        IAS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));   // static + dynamic may exceed 48 KB
        int grid = 1;
        IAS_TRY(gwin_grid(k, sm, n, &grid));
        if (!rw.cursor.p) IAS_TRY(rw.cursor.alloc(1));
        IAS_CUDA(cudaMemsetAsync(rw.cursor.p, 0, sizeof(int), c.stream));
        WorkOrder wo;
        const int *glist = nullptr;
        IAS_TRY(order_by_work(bl.rows_of(BIN_G), n, rw.ub.p, wo, &glist));
        IAS_LAUNCH(k, grid, GW_BLOCK, sm, glist, n, r0, A, B, rw.nnz_row.p, rw.cursor.p, ncols_b, swords);
        IAS_BIN_END(BIN_G);
        GWIN_PROFILE_DUMP("sym");
        rw.sym_timed[BIN_G] = true;
    } else if (bl.count[BIN_G]) {
        IAS_BIN_BEGIN(BIN_G);
        int n = (int)bl.count[BIN_G];
        IAS_TRY(ensure_gwork(rw, ncols_b, n));
        IAS_CUDA(cudaMemsetAsync(rw.cursor.p, 0, sizeof(int), c.stream));
        if (g_wide(ncols_b))
            IAS_LAUNCH((k_sym_global<AV, BV, 1024>), rw.gslots, 1024, 0, bl.rows_of(BIN_G), n, r0, A, B, rw.nnz_row.p, rw.gwork.p,
                       GLayout::make(ncols_b), rw.cursor.p);
        else
            IAS_LAUNCH((k_sym_global<AV, BV, G_BLOCK>), rw.gslots, G_BLOCK, 0, bl.rows_of(BIN_G), n, r0, A, B, rw.nnz_row.p, rw.gwork.p,
                       GLayout::make(ncols_b), rw.cursor.p);
        IAS_BIN_END(BIN_G);
        rw.sym_timed[BIN_G] = true;
    }
    if (st) {
        st->products = rw.products;
        for (int b = 0; b < 8; ++b) st->sym_bin_rows[b] = b < NBINS ? rw.sym_hist[b] : 0;
    }
    return IAS_OK;
}

// numeric bins of local rows [b0, b1) of the range
template <class AV, class BV>
int numeric_rows(const AV &A, const BV &B, RangeWork &rw, int b0, int b1, int ncols_b, const OutMap &out_in,
                 int *c_ci, double *c_v, IasSpgemmStats *st, int tiny_max_nnz = 0)
{
    Ctx &c = ctx();
    int n = b1 - b0;
    if (n <= 0) return IAS_OK;
    long long h[NBINS + 2] = {0};
    BinLists bl;
    if (rw.tiny_only && b0 == 0 && b1 == rw.nrows && tiny_max_nnz > 0) {
        // numeric bins == symbolic bins; the largest tiny nnz(C_i) came back with the scan total
        h[BIN_T] = rw.sym_hist[BIN_T]; h[BIN_EMPTY] = rw.sym_hist[BIN_EMPTY]; h[NBINS] = tiny_max_nnz;
        for (int b = 0; b < NBINS; ++b) { rw.num_hist[b] += h[b]; bl.count[b] = h[b]; }
        bl.only_bin = rw.tiny_identity ? BIN_T : -1;
        if (!rw.tiny_identity) { bl.list.p = rw.tiny_list.release(); bl.offset[BIN_T] = 0; }
    } else {
        IAS_CUDA(cudaMemsetAsync(rw.hist.p, 0, 16 * sizeof(unsigned long long), c.stream));
        // rows the windowed kernel handles in one rank window go there instead of the large CTA hash (hash insert +
        // block radix sort cost more per product than mark + rank when the whole column space is one super-window)
        const bool g2 = rw.b_canonical && c.tune.global_rows_smem != 0 && c.tune.g_v2 != 0 && c.tune.g_block != 512 && !gwin_numeric_pays(ncols_b);
        const int b2_max = (use_gwin(rw) && gwin_numeric_pays(ncols_b) && c.tune.gwin_takes_b2) || (g2 && c.tune.g2_takes_b2) ? NUM_B1_NNZ : NUM_B2_NNZ;
        IAS_LAUNCH(k_classify_num, grid_for(n, 256), 256, 0, n, rw.ub.p + b0, rw.nnz_row.p + b0, rw.bin.p + b0, rw.hist.p, b2_max);
        IAS_TRY(read_hist(rw, NBINS + 2, h));
        for (int b = 0; b < NBINS; ++b) rw.num_hist[b] += h[b];
        IAS_TRY(build_bin_lists(n, rw.bin.p + b0, h, bl));
    }
    int r0 = rw.r0 + b0;                         // lists hold indices local to [b0, b1)
    OutMap out = out_in;
    if (out.rp) out.rp += b0;
    if (out.nnz_row) out.nnz_row += b0;
    if (bl.count[BIN_T]) {
        IAS_BIN_BEGIN(8 + BIN_T);
        int m = (int)bl.count[BIN_T];
        int cap = std::max(1, (int)h[NBINS]);
        size_t sm = (size_t)TINY_BLOCK * cap * (sizeof(double) + sizeof(int)) + 48;       // + the alignment shift of the bulk copy-out
        int merge = rw.max_tiny_na <= 4 ? 4 : rw.max_tiny_na <= 5 ? 5 : rw.max_tiny_na <= 6 ? 6 : 8;
        auto k = merge == 4 ? k_num_tiny<AV, BV, TINY_BLOCK, 4> : merge == 5 ? k_num_tiny<AV, BV, TINY_BLOCK, 5>
               : merge == 6 ? k_num_tiny<AV, BV, TINY_BLOCK, 6> : k_num_tiny<AV, BV, TINY_BLOCK, 8>;
        IAS_TRY(opt_in_smem(k, sm));
        IAS_LAUNCH(k, grid_for(m, TINY_BLOCK), TINY_BLOCK, sm, bl.rows_of(BIN_T), m, r0, A, B, out, c_ci, c_v, cap, rw.b_canonical, (int)(c.tune.bulk_store != 0));
        IAS_BIN_END(8 + BIN_T);
        rw.num_timed[BIN_T] = true;
    }
    int col_bits = 1;
    while (col_bits < 31 && (1LL << col_bits) < (long long)ncols_b) ++col_bits;
    if (bl.count[BIN_W]) {
        IAS_BIN_BEGIN(8 + BIN_W);
        int m = (int)bl.count[BIN_W];
        // items per lane from the largest product count of the bin; 32-bit keys when column and index fit in 31 bits
        int max_ub = std::min(std::max(rw.max_warp_ub, T_MAX + 1), NUM_W_UB);
        int ipl = max_ub <= 64 ? 2 : max_ub <= 128 ? 4 : max_ub <= 256 ? 8 : 16;
        int idx_bits = ipl == 16 ? 9 : ipl == 8 ? 8 : ipl == 4 ? 7 : 6;
        bool k32 = col_bits + idx_bits <= 31;
        constexpr int EB = 128;                       // 4 rows per CTA
        size_t sm = (size_t)(EB / 32) * 32 * ipl * ((k32 ? 4 : 8) + 8);
        const int *rl = bl.rows_of(BIN_W);
#define IAS_ESC(KT, IPL)                                                                                      \
        do {                                                                                                  \
            auto k = k_esc_warp<AV, BV, KT, IPL, EB, true>;                                                   \
            IAS_TRY(opt_in_smem(k, sm));                                                                      \
            IAS_LAUNCH(k, grid_for(m, EB / 32), EB, sm, rl, m, r0, A, B, out, (int *)nullptr, c_ci, c_v);     \
        } while (0)
        if (k32) { if (ipl == 2) IAS_ESC(unsigned, 2); else if (ipl == 4) IAS_ESC(unsigned, 4); else if (ipl == 8) IAS_ESC(unsigned, 8); else IAS_ESC(unsigned, 16); }
        else     { if (ipl == 2) IAS_ESC(unsigned long long, 2); else if (ipl == 4) IAS_ESC(unsigned long long, 4); else if (ipl == 8) IAS_ESC(unsigned long long, 8); else IAS_ESC(unsigned long long, 16); }
#undef IAS_ESC
        IAS_BIN_END(8 + BIN_W);
        rw.num_timed[BIN_W] = true;
    }
    if (bl.count[BIN_B1]) {
        IAS_BIN_BEGIN(8 + BIN_B1);
        int m = (int)bl.count[BIN_B1];
        auto k = k_num_hash_cta<AV, BV, 512, NUM_B1_TSIZE>;
        size_t sm = NumHashSmem<512, NUM_B1_TSIZE>::BYTES;
        IAS_TRY(opt_in_smem(k, sm));
        IAS_LAUNCH(k, m, 512, sm, bl.rows_of(BIN_B1), m, r0, A, B, out, c_ci, c_v, col_bits);
        IAS_BIN_END(8 + BIN_B1);
        rw.num_timed[BIN_B1] = true;
    }
    if (bl.count[BIN_B2]) {
        IAS_BIN_BEGIN(8 + BIN_B2);
        int m = (int)bl.count[BIN_B2];
        auto k = k_num_hash_cta<AV, BV, 1024, NUM_B2_TSIZE>;
        size_t sm = NumHashSmem<1024, NUM_B2_TSIZE>::BYTES;
        IAS_TRY(opt_in_smem(k, sm));
        IAS_LAUNCH(k, m, 1024, sm, bl.rows_of(BIN_B2), m, r0, A, B, out, c_ci, c_v, col_bits);
        IAS_BIN_END(8 + BIN_B2);
        rw.num_timed[BIN_B2] = true;
    }
    if (bl.count[BIN_G] && use_gwin(rw) && gwin_numeric_pays(ncols_b)) {
        IAS_BIN_BEGIN(8 + BIN_G);
        int m = (int)bl.count[BIN_G];
        auto k = k_num_gwin<AV, BV, GW_BLOCK>;
        size_t dyn = 0;
        IAS_TRY(gwin_dyn_max(k, &dyn));
        // {bitmap, rank} cells of a super-window (8 B per 32 columns) + rank window (8 B sum + 4 B column per entry)
        long long sw_cap = c.tune.gwin_swords > 0 ? std::max<long long>(32, (c.tune.gwin_swords + 31) & ~31LL) : 8192;
        sw_cap = std::min<long long>(sw_cap, (long long)((dyn / 2) / 8) & ~31LL);
        int swords = (int)std::min<long long>(words_of(ncols_b), sw_cap);
        // split-point table (one int per A entry and super-window boundary) only when there are several super-windows
        int tbl_cap = words_of(ncols_b) > swords ? 4096 : 0;
        if (swords < 32) return fail(IAS_E_ARG, "gwin_smem_kb leaves %zu bytes of shared memory: too little for a cell window", dyn);
        if (dyn < (size_t)swords * 8 + (size_t)tbl_cap * 4 + 64 * 12) tbl_cap = 0;          // a tight budget drops the split table first
        if (dyn < (size_t)swords * 8 + 32 * 12) return fail(IAS_E_ARG, "gwin_smem_kb leaves %zu bytes of shared memory: too little for a rank window", dyn);
        int win = (int)((dyn - (size_t)swords * 8 - (size_t)tbl_cap * 4) / 12) & ~31;
        if (c.tune.gwin_win > 0) win = (int)std::min<long long>(win, std::max<long long>(32, c.tune.gwin_win & ~31LL));
        size_t sm = (size_t)swords * 8 + (size_t)win * 12 + (size_t)tbl_cap * 4;
        IAS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int grid = 1;
        IAS_TRY(gwin_grid(k, sm, m, &grid));
        if (!rw.cursor.p) IAS_TRY(rw.cursor.alloc(1));
        IAS_CUDA(cudaMemsetAsync(rw.cursor.p, 0, sizeof(int), c.stream));
        WorkOrder wo;
        const int *glist = nullptr;
        IAS_TRY(order_by_work(bl.rows_of(BIN_G), m, rw.ub.p + b0, wo, &glist));
        IAS_LAUNCH(k, grid, GW_BLOCK, sm, glist, m, r0, A, B, out, c_ci, c_v, rw.cursor.p, ncols_b, swords, win, tbl_cap);
        IAS_BIN_END(8 + BIN_G);
        GWIN_PROFILE_DUMP("num");
        rw.num_timed[BIN_G] = true;
    } else if (bl.count[BIN_G]) {
        IAS_BIN_BEGIN(8 + BIN_G);
        int m = (int)bl.count[BIN_G];
        const bool two = c.tune.g_block == 512;
        const bool v2 = rw.b_canonical && c.tune.global_rows_smem != 0 && c.tune.g_v2 != 0 && !two;
        IAS_TRY(ensure_gwork(rw, ncols_b, m, two ? 2 : 1));
        IAS_CUDA(cudaMemsetAsync(rw.cursor.p, 0, sizeof(int), c.stream));
        WorkOrder wo;
        const int *glist = nullptr;
        IAS_TRY(order_by_work(bl.rows_of(BIN_G), m, rw.ub.p + b0, wo, &glist));
        if (v2) {
            // second generation (canonical B): rank + emit from the shared-memory bitmap, split tables for the windows.
            // 128 KB tile + 32 KB of split points = the 160 KB the first generation gives to its tile alone, so the
            // L1 that is left for the B-row stream is the same.
            auto k = k_num_global2<AV, BV, 1024>;
            int win = (int)std::min<long long>(16384, std::max<long long>(64, c.tune.g_win & ~63LL));     // bitmap of 2 * win words, scanned 128 at a time
            int tbl_cap = (int)std::min<long long>(16384, std::max<long long>(0, c.tune.g_tbl));
            size_t sm = (size_t)win * sizeof(double) + (size_t)tbl_cap * sizeof(int);
            IAS_TRY(opt_in_smem(k, sm));
            // per-CTA scratch for the split tables of rows with more than 1024 entries in A (L2 / HBM, never initialised)
            const int grid = (int)std::min<long long>(rw.gslots, (long long)c.sm_count);
            const int scr_cap = (int)std::min<long long>(1 << 24, std::max<long long>(16, c.tune.g_scr));
            if (rw.gscr_ints < (size_t)grid * scr_cap) {
                IAS_TRY(rw.gscr.alloc((size_t)grid * scr_cap));
                rw.gscr_ints = (size_t)grid * scr_cap;
            }
            // rows worth splitting over several CTAs (a prefix of the work-ordered list; counted on the device)
            const GLayout lay = GLayout::make(ncols_b);
            const int split_cap = 1024;
            int split_words = 0;
            if (c.tune.g_split != 0 && glist == wo.list.p && wo.list.p && ncols_b <= 0x7f000000) {      // (column ranges are ints)
                const long long parts = std::min<long long>(2048, std::max<long long>(2, c.tune.g_split_parts));
                long long sw = ((lay.words + parts - 1) / parts + 127) / 128 * 128;
                sw = std::min<long long>(sw, std::min<long long>(2LL * win, 2LL * 1024 * 32));
                const long long P = (lay.words + sw - 1) / sw;
                if (P > 1) {
                    split_words = (int)sw;
                    if (rw.split_ints < (size_t)split_cap * P + 1) {
                        IAS_TRY(rw.split.alloc((size_t)split_cap * P + 1));
                        rw.split_ints = (size_t)split_cap * P + 1;
                    }
                    IAS_CUDA(cudaMemsetAsync(rw.split.p, 0, sizeof(int) * ((size_t)split_cap * P + 1), c.stream));
                    IAS_LAUNCH(k_count_split_rows, 1, 1024, 0, wo.keys_sorted.p, m, c.tune.g_split_ub, (long long)1 << 18, grid, rw.split.p);
                }
            }
            IAS_LAUNCH(k, grid, 1024, sm, glist, m, r0, A, B, out, c_ci, c_v, rw.gwork.p,
                       lay, rw.cursor.p, win, tbl_cap, ncols_b, rw.gscr.p, scr_cap,
                       split_words ? rw.split.p : (int *)nullptr, split_cap, split_words, split_words ? rw.split.p + 1 : (int *)nullptr);
        } else {
            // 160 KB tile of fp64 partial sums per SM (192 KB would leave 28 KB of L1 for the B-row stream: ncu/clock64
            // showed the mark pass 1.6x slower); with two 512-thread CTAs per SM each gets half
            int win = (int)std::min<long long>(two ? 10240 : 20480, std::max<long long>(16, c.tune.g_win & ~15LL));
            size_t sm = (size_t)win * sizeof(double);
            int smem_mark = rw.b_canonical && c.tune.global_rows_smem != 0 && (win % 16) == 0 ? 1 : 0;      // bit 0: mark pass in smem
            if (rw.b_canonical && c.tune.g_coop) smem_mark |= 2;                                            // bit 1: accumulate via gwin_build / gwin_run
            if (c.tune.g_ldca) smem_mark |= 4;                                                              // bit 2: cell lookups through L1
            if (two) {
                auto k = k_num_global<AV, BV, 512>;
                IAS_TRY(opt_in_smem(k, sm));
                IAS_LAUNCH(k, std::min<long long>(rw.gslots, 2LL * c.sm_count), 512, sm, glist, m, r0, A, B, out, c_ci, c_v, rw.gwork.p,
                           GLayout::make(ncols_b), rw.cursor.p, win, rw.b_canonical, smem_mark, ncols_b);
            } else {
                auto k = k_num_global<AV, BV, 1024>;
                IAS_TRY(opt_in_smem(k, sm));
                IAS_LAUNCH(k, std::min<long long>(rw.gslots, (long long)c.sm_count), 1024, sm, glist, m, r0, A, B, out, c_ci, c_v, rw.gwork.p,
                           GLayout::make(ncols_b), rw.cursor.p, win, rw.b_canonical, smem_mark, ncols_b);
            }
        }
        IAS_BIN_END(8 + BIN_G);
        GWIN_PROFILE_DUMP("num_global");
        rw.num_timed[BIN_G] = true;
    }
    (void)st;
    return IAS_OK;
}

// 64-bit exclusive scan of nnz_row[0..nrows] (nrows+1 items; the last input is 0) -> rp[0..nrows]
struct CastI64 {
    __host__ __device__ long long operator()(int x) const { return (long long)x; }
};
inline int scan_row_ptr(const int *nnz_row, int nrows, long long *rp)
{
    auto in = thrust::make_transform_iterator(nnz_row, CastI64());
    size_t tb = 0;
    IAS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, rp, nrows + 1, ctx().stream));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, rp, nrows + 1, ctx().stream));
    ctx().launches += 2;
    return IAS_OK;
}

inline double ev_ms(int a, int b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx().ev[a], ctx().ev[b]);
    return (double)ms;
}

// rows [r0,r1) of C = A*B as a materialised CSR with 64-bit row pointers and column-sorted rows.
// Timed region (ev[0]..ev[4]) mirrors CUSPARSE_MUL_CUSPARSE (cusparse:74-93): symbolic pass, allocation
// of C, numeric pass, sort; operands already resident.
template <class AV, class BV>
int spgemm_materialise(const AV &av, const BV &bv, double avg_a_row, int ncols_b, int r0, int r1, IasCsr64Dev *C,
                       IasSpgemmStats *st, bool b_is_a = false, int b_rows = 0, long long b_nnz = -1)
{
    Ctx &c = ctx();
    long long l0 = c.launches;
    IasSpgemmStats local;
    memset(&local, 0, sizeof local);
    int nrows = r1 - r0;
    C->row = nrows; C->col = ncols_b; C->nnz = 0;
    C->row_ptr_dev = nullptr; C->col_ind_dev = nullptr; C->values_dev = nullptr;

    IAS_CUDA(cudaEventRecord(c.ev[0], c.stream));
    IAS_CUDA(cudaEventRecord(c.ev[1], c.stream));
    RangeWork rw;
    IAS_TRY(symbolic_range(av, bv, r0, r1, ncols_b, avg_a_row, rw, &local, b_is_a, b_rows, b_nnz));
    IAS_CUDA(cudaEventRecord(c.ev[2], c.stream));

    DBuf<long long> rp;
    IAS_TRY(rp.alloc((size_t)nrows + 1));
    IAS_TRY(scan_row_ptr(rw.nnz_row.p, nrows, rp.p));
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 32, rp.p + nrows, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaMemcpyAsync(c.h_scalars + 33, rw.hist.p + 12, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    long long nnz = c.h_scalars[32];
    int tiny_max_nnz = (int)c.h_scalars[33];
    DBuf<int> ci;
    DBuf<double> cv;
    IAS_TRY(ci.alloc((size_t)nnz));
    IAS_TRY(cv.alloc((size_t)nnz));
    IAS_CUDA(cudaEventRecord(c.ev[3], c.stream));

    OutMap out{rp.p, 0, nullptr, 0};
    IAS_TRY(numeric_rows(av, bv, rw, 0, nrows, ncols_b, out, ci.p, cv.p, &local, tiny_max_nnz));
    IAS_CUDA(cudaEventRecord(c.ev[4], c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    collect_bin_times(rw);
    for (int b = 0; b < 8; ++b) { local.ms_bin_sym[b] = rw.ms_bin_sym[b]; local.ms_bin_num[b] = rw.ms_bin_num[b]; }

    C->nnz = nnz;
    C->row_ptr_dev = rp.release(); C->col_ind_dev = ci.release(); C->values_dev = cv.release();
    local.nnz = nnz;
    local.batches = 1;
    local.ms_analyze = ev_ms(0, 1); local.ms_symbolic = ev_ms(1, 2); local.ms_scan = ev_ms(2, 3);
    local.ms_numeric = ev_ms(3, 4); local.ms_total = ev_ms(0, 4);
    for (int b = 0; b < 8; ++b) local.num_bin_rows[b] = b < NBINS ? rw.num_hist[b] : 0;
    local.kernel_launches = (int)(c.launches - l0);
    if (st) *st = local;
    return IAS_OK;
}

}  // inline namespace IAS_TU
}  // namespace ias
