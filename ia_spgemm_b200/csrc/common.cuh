// common.cuh -- engine context, error handling and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/iaspgemm.h"

namespace ias {

// knobs of the kernel selection (ias_set_option / IAS_OPT_<NAME> in the environment at ias_init)
struct Tuning {
    long long global_rows_smem = 1;   // 1: windowed shared-memory kernels for global rows (canonical B), 0: L2 bitmap kernels
    long long gwin_swords = 8192;     // numeric: bitmap words per super-window (x32 columns)
    long long gwin_win = 0;           // numeric: entries per rank window (0: whatever shared memory is left)
    long long gwin_sym_swords = 0;    // symbolic: bitmap words per super-window (0: as many as fit)
    long long gwin_smem_kb = 0;       // cap on the dynamic shared memory of the windowed kernels (0: device limit)
    long long g_win = 20480;          // L2 bitmap kernel: entries per accumulate window (multiple of 16; capped at 20480, 16384 for the second generation)
    long long g_coop = 1;             // L2 bitmap kernel: accumulate pass enumerates products with gwin_build / gwin_run
    long long gwin_takes_b2 = 1;      // rows of the large CTA hash bin go to the windowed kernel when it is selected
    long long gwin_max_sw = 1;        // numeric: use the windowed kernel up to this many super-windows per row (0: always);
                                      // beyond, the per-window scans of the cells cost more than the L2 lookups they replace
    long long g_ldca = 0;             // L2 bitmap kernel: accumulate-pass cell lookups go through L1
    long long g_v2 = 1;               // L2 bitmap kernel, canonical B: second generation (rank + emit from shared memory, split tables)
    long long g_tbl = 8192;           // ... its split-table capacity (ints of shared memory)
    long long g_scr = 16 << 20;       // ... per-CTA global scratch (ints) for the split tables of rows with > 1024 A entries (0: off)
    long long g_split = 1;            // ... rows with very many products are cut into column-range parts, one CTA each (0: one CTA per row)
    long long g_split_ub = 0;         // ... products from which a row is cut (0: a quarter of one CTA's even share of the launch, at least 256 Ki)
    long long g_split_parts = 128;    // ... parts per cut row (the column space in equal ranges, multiples of 4096 columns)
    long long g2_takes_b2 = 0;        // rows of the large CTA hash bin (nnz 4097..12288) go to the second-generation global-row kernel
    long long g_lpt = 1;              // global rows are handed out in order of decreasing work
    long long g_block = 1024;         // L2 bitmap kernel: threads per CTA (1024: one row per SM, 512: two)
    long long bulk_store = 1;         // shared -> global bulk copies (cp.async.bulk) for staged output tiles (0: per-thread stores)
    long long e2e_pipeline = 1;       // ias_spgemm_auto_host: chunked upload / DIA multiply / download for banded A^2 (speculative, verified)
    long long dia_vec = 0;            // DIA x DIA: two adjacent rows per thread with 128-bit accesses -- measured SLOWER than the scalar
                                      // kernel on B200 (0.617 vs 0.542 ms on Poisson 4096^2), kept for the record
    long long ell_onepass = 1;        // ELL x ELL: one-pass register-sort kernel when a row's products fit (0: always the pipeline)
    long long block_cache = 1;        // freed device blocks are kept per size class and handed out again without a driver call (0: every
                                      // allocation goes to the stream-ordered pool, whose re-mapping blocks the caller for milliseconds)
    long long trust_operand_cache = 0; // 1: remember B's canonical flag per operand (pointers + shape) across calls; the caller
                                      // promises not to rewrite or re-allocate an operand without ias_forget_operand
};

struct Ctx {
    Tuning tune;
    bool ready = false;
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 227 * 1024;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;      // stream every engine kernel is launched on
    cudaStream_t s_up = nullptr, s_down = nullptr;     // copy streams of the pipelined host path (non-blocking)
    cudaEvent_t ev_pipe[96] = {};       // upload / compute / download events of its chunks (no timing)
    cudaMemPool_t pool = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_bin[32] = {};        // [2*bin], [2*bin+1]: symbolic bins 0..7, numeric bins 8..15
    long long launches = 0;             // engine kernels launched since ias_init
    char err[512] = {0};
    // pinned host arena for ias_csr_mul_csr_host results (grow only)
    void *h_arena = nullptr;
    size_t h_arena_bytes = 0;
    // last operand whose rows were checked for strictly increasing columns (see ias_forget_operand)
    const void *canon_ci = nullptr, *canon_rp = nullptr;
    long long canon_rows = -1, canon_nnz = -1;
    int canon_flag = 0;
    // small pinned scratch for scalar read-backs
    long long *h_scalars = nullptr;     // 64 entries
};

Ctx &ctx();
int device_block(void **p, size_t bytes);       // stream-ordered device memory on ctx().stream: block cache first, then the pool
void device_block_free(void *p);
void block_cache_flush();                       // cached blocks go back to the pool (the caller synchronised the stream)
size_t free_device_bytes();                     // what an allocation could get: free + the pool's idle reserve + cached blocks
int host_arena(size_t bytes, void **p);        // pinned host memory owned by the engine (grow only; ias_release_host frees it)
int ensure_pipe_streams();
// DIA helpers shared with the auto path (dia.cu)
void dia_rows_major(int rows, int nd, const double *in_diag_major, double *out_row_major, cudaStream_t s);
int auto_dia_pipelined(const IasCsrMatrix *A, double gate, IasAutoResult *out, IasCsrMatrixDev *dA, int *have_dA, int *done);
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);
int fail(int code, const char *fmt, ...);
int ensure_init();

#define IAS_CUDA(x)                                                         \
    do {                                                                    \
        cudaError_t e__ = (x);                                              \
        if (e__ != cudaSuccess) return ias::fail_cuda(e__, #x, __FILE__, __LINE__); \
    } while (0)

#define IAS_TRY(x)                 \
    do {                           \
        int rc__ = (x);            \
        if (rc__ != IAS_OK) return rc__; \
    } while (0)

// launch bookkeeping: every engine kernel goes through this so gpu_launches is a count, not a guess
#define IAS_LAUNCH(kernel, grid, block, smem, ...)                                   \
    do {                                                                             \
        kernel<<<(grid), (block), (smem), ias::ctx().stream>>>(__VA_ARGS__);         \
        ias::ctx().launches++;                                                       \
        cudaError_t le__ = cudaGetLastError();                                       \
        if (le__ != cudaSuccess) return ias::fail_cuda(le__, #kernel, __FILE__, __LINE__); \
    } while (0)

// IAS_HOST_TRACE=1: report (stderr) every allocation, free or stream wait that blocks the calling thread for more than 2 ms
struct HostTrace {
    const char *what; size_t bytes; std::chrono::steady_clock::time_point t0; bool on;
    static bool enabled() { static int e = -1; if (e < 0) { const char *v = getenv("IAS_HOST_TRACE"); e = (v && *v && *v != '0') ? 1 : 0; } return e == 1; }
    HostTrace(const char *w, size_t b) : what(w), bytes(b), on(enabled()) { if (on) t0 = std::chrono::steady_clock::now(); }
    void done()
    {
        if (!on) return;
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 2.0) fprintf(stderr, "[ias host trace] %s (%zu bytes) blocked the caller for %.1f ms\n", what, bytes, ms);
    }
};

// stream-ordered allocation from the engine pool (no cudaMalloc+cudaMemset per call like DevMalloc)
template <class T>
inline int dalloc(T **p, size_t n)
{
    *p = nullptr;
    if (n == 0) n = 1;
    return device_block((void **)p, n * sizeof(T));
}
template <class T>
inline void dfree(T *p)
{
    if (p) device_block_free((void *)p);
}

// RAII holder for temporaries inside one API call
template <class T>
struct DBuf {
    T *p = nullptr;
    ~DBuf() { dfree(p); }
    int alloc(size_t n) { dfree(p); return dalloc(&p, n); }
    T *release() { T *q = p; p = nullptr; return q; }
    operator T *() const { return p; }
};

inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// ---------------------------------------------------------------- bulk asynchronous copies (TMA, 1-D: cp.async.bulk -> SASS UBLKCP)
// shared -> global: the issuing thread's earlier generic-proxy writes to the source (and, after a barrier, everyone's) are
// made visible to the async proxy with fence.proxy.async; source, destination and size must be multiples of 16 bytes.
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_shared_to_global(void *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// global -> shared with an mbarrier that counts the bytes (SASS: UBLKCP + SYNCS)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_global_to_shared(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ unsigned hash_col(int k) { return (unsigned)k * 0x9E3779B1u; }

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ias
