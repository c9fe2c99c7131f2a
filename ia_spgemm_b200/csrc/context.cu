// context.cu -- engine context, transfers, checksums (C ABI part 1).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"

namespace ias {

static Ctx g_ctx;
Ctx &ctx() { return g_ctx; }

int fail_cuda(cudaError_t e, const char *what, const char *file, int line)
{
    snprintf(g_ctx.err, sizeof g_ctx.err, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();
    return IAS_E_CUDA;
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ctx.err, sizeof g_ctx.err, fmt, ap);
    va_end(ap);
    return code;
}

int ensure_init()
{
    if (g_ctx.ready) return IAS_OK;
    return ias_init(0);
}

int host_arena(size_t bytes, void **p)
{
    Ctx &c = g_ctx;
    if (c.h_arena_bytes < bytes) {
        if (c.h_arena) cudaFreeHost(c.h_arena);
        c.h_arena = nullptr; c.h_arena_bytes = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMallocHost(&c.h_arena, want);
        if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMallocHost(&c.h_arena, want); }      // without the growth slack
        if (e != cudaSuccess) {
            cudaGetLastError();
            c.h_arena = nullptr;
            return fail(IAS_E_NOMEM, "pinned host allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        c.h_arena_bytes = want;
    }
    *p = c.h_arena;
    return IAS_OK;
}

// ---------------------------------------------------------------- device block cache
// The stream-ordered pool alone is not enough: with differently sized requests in flight it re-maps physical memory
// inside cudaMallocFromPoolAsync, and that blocks the calling thread for 2..700 ms at a time (IAS_HOST_TRACE=1 shows
// it; profiles/r02_alloc_stalls.md).  A multiply repeats the same request sizes call after call, so freed blocks are
// kept here per size class (4 mantissa bits: at most 6.25 % larger than asked) and handed out again without a driver
// call.  Ordering is the stream's: every engine allocation and free belongs to ctx().stream (a block freed while its
// last kernel is still queued is only ever given to work queued behind it on the same stream; ias_set_stream
// synchronises when the stream changes).  Cached bytes are capped at half the device; when the pool cannot serve a
// request the cache is emptied into it and the request retried.
namespace {
struct SizeClass {
    std::vector<void *> idle;
    unsigned long long last_use = 0;
};
struct BlockCache {
    std::mutex mu;
    std::unordered_map<size_t, SizeClass> classes;             // size class -> idle blocks of that size
    std::unordered_map<void *, size_t> live;                   // block -> its size class (handed out or idle)
    unsigned long long tick = 0;
    size_t idle_bytes = 0;
    size_t cap_bytes = 0;
} g_blocks;

size_t size_class(size_t bytes)
{
    if (bytes < 512) return 512;
    int top = 63 - __builtin_clzll((unsigned long long)bytes);
    size_t step = (size_t)1 << (top - 4);
    return (bytes + step - 1) & ~(step - 1);
}

// idle blocks go back to the pool, least recently used size class first, until at most keep_bytes stay cached
void evict_locked(size_t keep_bytes)
{
    BlockCache &b = g_blocks;
    if (b.idle_bytes <= keep_bytes) return;
    std::vector<std::pair<unsigned long long, size_t>> by_age;
    for (std::unordered_map<size_t, SizeClass>::iterator it = b.classes.begin(); it != b.classes.end(); ++it)
        if (!it->second.idle.empty()) by_age.push_back(std::make_pair(it->second.last_use, it->first));
    std::sort(by_age.begin(), by_age.end());
    for (size_t k = 0; k < by_age.size() && b.idle_bytes > keep_bytes; ++k) {
        SizeClass &sc = b.classes[by_age[k].second];
        while (!sc.idle.empty() && b.idle_bytes > keep_bytes) {
            void *q = sc.idle.back();
            sc.idle.pop_back();
            b.live.erase(q);
            b.idle_bytes -= by_age[k].second;
            cudaFreeAsync(q, g_ctx.stream);
        }
    }
}
}  // namespace

int device_block(void **p, size_t bytes)
{
    Ctx &c = g_ctx;
    BlockCache &b = g_blocks;
    const bool cached = c.tune.block_cache != 0;
    const size_t cls = cached ? size_class(bytes) : bytes;
    if (cached) {
        std::lock_guard<std::mutex> g(b.mu);
        std::unordered_map<size_t, SizeClass>::iterator it = b.classes.find(cls);
        if (it != b.classes.end() && !it->second.idle.empty()) {
            *p = it->second.idle.back();
            it->second.idle.pop_back();
            it->second.last_use = ++b.tick;
            b.idle_bytes -= cls;
            return IAS_OK;
        }
    }
    HostTrace ht("cudaMallocFromPoolAsync", cls);
    cudaError_t e = cudaMallocFromPoolAsync(p, cls, c.pool, c.stream);
    if (e != cudaSuccess && b.idle_bytes) {                    // the memory may be sitting in the cache in other size classes
        cudaGetLastError();
        cudaStreamSynchronize(c.stream);
        block_cache_flush();
        e = cudaMallocFromPoolAsync(p, cls, c.pool, c.stream);
    }
    ht.done();
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(IAS_E_NOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    if (cached) {
        std::lock_guard<std::mutex> g(b.mu);
        b.live[*p] = cls;
    }
    return IAS_OK;
}

void device_block_free(void *p)
{
    if (!p) return;
    Ctx &c = g_ctx;
    BlockCache &b = g_blocks;
    {
        std::lock_guard<std::mutex> g(b.mu);
        std::unordered_map<void *, size_t>::iterator it = b.live.find(p);
        if (it != b.live.end()) {
            if (c.tune.block_cache != 0) {
                if (!b.cap_bytes) {
                    size_t f = 0, t = 0;
                    b.cap_bytes = cudaMemGetInfo(&f, &t) == cudaSuccess ? t / 2 : (size_t)64 << 30;
                }
                SizeClass &sc = b.classes[it->second];
                sc.idle.push_back(p);
                sc.last_use = ++b.tick;
                b.idle_bytes += it->second;
                if (b.idle_bytes > b.cap_bytes) evict_locked(b.cap_bytes);
                return;
            }
            b.live.erase(it);
        }
    }
    HostTrace ht("cudaFreeAsync", 0);
    cudaFreeAsync(p, c.stream);
    ht.done();
}

void block_cache_flush()
{
    std::lock_guard<std::mutex> g(g_blocks.mu);
    evict_locked(0);
}

size_t free_device_bytes()
{
    size_t f = 0, t = 0;
    if (cudaMemGetInfo(&f, &t) != cudaSuccess) { cudaGetLastError(); return 0; }
    unsigned long long reserved = 0, used = 0;
    if (g_ctx.pool && cudaMemPoolGetAttribute(g_ctx.pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(g_ctx.pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        f += (size_t)(reserved - used);
    else
        cudaGetLastError();
    return f + g_blocks.idle_bytes;
}

int ensure_pipe_streams()
{
    Ctx &c = g_ctx;
    if (!c.s_up) IAS_CUDA(cudaStreamCreateWithFlags(&c.s_up, cudaStreamNonBlocking));
    if (!c.s_down) IAS_CUDA(cudaStreamCreateWithFlags(&c.s_down, cudaStreamNonBlocking));
    for (int i = 0; i < 96; ++i)
        if (!c.ev_pipe[i]) IAS_CUDA(cudaEventCreateWithFlags(&c.ev_pipe[i], cudaEventDisableTiming));
    return IAS_OK;
}

}  // namespace ias

using namespace ias;

extern "C" {

const char *ias_version(void) { return "ia-spgemm-b200 0.1 (sm_100a)"; }
const char *ias_last_error(void) { return ctx().err; }
long long ias_kernel_launches(void) { return ctx().launches; }

static long long *option_slot(const char *name)
{
    Tuning &t = ctx().tune;
    if (!name) return nullptr;
    if (!strcmp(name, "global_rows_smem")) return &t.global_rows_smem;
    if (!strcmp(name, "gwin_swords")) return &t.gwin_swords;
    if (!strcmp(name, "gwin_win")) return &t.gwin_win;
    if (!strcmp(name, "gwin_sym_swords")) return &t.gwin_sym_swords;
    if (!strcmp(name, "gwin_smem_kb")) return &t.gwin_smem_kb;
    if (!strcmp(name, "gwin_max_sw")) return &t.gwin_max_sw;
    if (!strcmp(name, "g_win")) return &t.g_win;
    if (!strcmp(name, "g_coop")) return &t.g_coop;
    if (!strcmp(name, "gwin_takes_b2")) return &t.gwin_takes_b2;
    if (!strcmp(name, "trust_operand_cache")) return &t.trust_operand_cache;
    if (!strcmp(name, "g_ldca")) return &t.g_ldca;
    if (!strcmp(name, "g_block")) return &t.g_block;
    if (!strcmp(name, "ell_onepass")) return &t.ell_onepass;
    if (!strcmp(name, "bulk_store")) return &t.bulk_store;
    if (!strcmp(name, "dia_vec")) return &t.dia_vec;
    if (!strcmp(name, "e2e_pipeline")) return &t.e2e_pipeline;
    if (!strcmp(name, "g_v2")) return &t.g_v2;
    if (!strcmp(name, "g_tbl")) return &t.g_tbl;
    if (!strcmp(name, "g_lpt")) return &t.g_lpt;
    if (!strcmp(name, "g_scr")) return &t.g_scr;
    if (!strcmp(name, "g2_takes_b2")) return &t.g2_takes_b2;
    if (!strcmp(name, "block_cache")) return &t.block_cache;
    if (!strcmp(name, "g_split")) return &t.g_split;
    if (!strcmp(name, "g_split_ub")) return &t.g_split_ub;
    if (!strcmp(name, "g_split_parts")) return &t.g_split_parts;
    return nullptr;
}

int ias_init(int device)
{
    Ctx &c = ctx();
    if (c.ready && c.device == device) return IAS_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(IAS_E_CUDA, "no CUDA device available (%s): the engine has no CPU fallback", cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(IAS_E_ARG, "device %d out of range (have %d)", device, n);
    IAS_CUDA(cudaSetDevice(device));
    if (c.ready && c.device != device) {
        // re-bound to another device: streams, events and cached blocks belong to the device they were created on
        {
            int now = device;
            cudaSetDevice(c.device);
            cudaStreamSynchronize(c.stream);
            block_cache_flush();
            cudaSetDevice(now);
        }
        if (c.own_stream) cudaStreamDestroy(c.own_stream);
        c.own_stream = nullptr;
        for (int i = 0; i < 8; ++i) { if (c.ev[i]) cudaEventDestroy(c.ev[i]); c.ev[i] = nullptr; }
        for (int i = 0; i < 32; ++i) { if (c.ev_bin[i]) cudaEventDestroy(c.ev_bin[i]); c.ev_bin[i] = nullptr; }
        for (int i = 0; i < 96; ++i) { if (c.ev_pipe[i]) cudaEventDestroy(c.ev_pipe[i]); c.ev_pipe[i] = nullptr; }
        if (c.s_up) cudaStreamDestroy(c.s_up);
        if (c.s_down) cudaStreamDestroy(c.s_down);
        c.s_up = c.s_down = nullptr;
        c.canon_ci = c.canon_rp = nullptr; c.canon_rows = c.canon_nnz = -1;
        c.ready = false;
        cudaGetLastError();
    }
    c.device = device;
    cudaDeviceProp p;
    IAS_CUDA(cudaGetDeviceProperties(&p, device));
    c.sm_count = p.multiProcessorCount;
    c.smem_optin = p.sharedMemPerBlockOptin;
    if (!c.own_stream) IAS_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    IAS_CUDA(cudaDeviceGetDefaultMemPool(&c.pool, device));
    unsigned long long keep = ~0ull;     // keep freed blocks in the pool: allocation cost is paid once
    IAS_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    for (int i = 0; i < 8; ++i)
        if (!c.ev[i]) IAS_CUDA(cudaEventCreate(&c.ev[i]));
    for (int i = 0; i < 32; ++i)
        if (!c.ev_bin[i]) IAS_CUDA(cudaEventCreate(&c.ev_bin[i]));
    if (!c.h_scalars) IAS_CUDA(cudaMallocHost((void **)&c.h_scalars, 64 * sizeof(long long)));
    static const char *const names[] = {"global_rows_smem", "gwin_swords", "gwin_win", "gwin_sym_swords", "gwin_smem_kb", "gwin_max_sw", "g_win", "g_coop", "gwin_takes_b2",
                                        "trust_operand_cache", "g_ldca", "g_block", "ell_onepass", "g_v2", "g_tbl", "g_lpt", "bulk_store", "dia_vec", "g_scr", "g2_takes_b2", "e2e_pipeline", "block_cache", "g_split", "g_split_ub", "g_split_parts"};
    for (const char *n : names) {                  // IAS_OPT_GWIN_WIN=4096 etc.
        char env[64] = "IAS_OPT_";
        size_t k = strlen(env);
        for (const char *q = n; *q && k + 1 < sizeof env; ++q) env[k++] = (char)(*q >= 'a' && *q <= 'z' ? *q - 32 : *q);
        env[k] = 0;
        const char *v = getenv(env);
        if (v && *v) *option_slot(n) = atoll(v);
    }
    c.ready = true;
    return IAS_OK;
}

int ias_set_option(const char *name, long long value)
{
    long long *slot = option_slot(name);
    if (!slot) return fail(IAS_E_ARG, "unknown option '%s'", name ? name : "(null)");
    if (value < 0) return fail(IAS_E_ARG, "option '%s' must not be negative", name);
    *slot = value;
    return IAS_OK;
}

int ias_get_option(const char *name, long long *value)
{
    long long *slot = option_slot(name);
    if (!slot || !value) return fail(IAS_E_ARG, "unknown option '%s'", name ? name : "(null)");
    *value = *slot;
    return IAS_OK;
}

int ias_set_stream(void *s)
{
    IAS_TRY(ensure_init());
    if (ctx().stream != (cudaStream_t)s) IAS_CUDA(cudaStreamSynchronize(ctx().stream));     // cached blocks carry the old stream's ordering
    ctx().stream = (cudaStream_t)s;          // NULL is the legacy default stream (what torch's default stream handle is)
    return IAS_OK;
}

int ias_use_own_stream(void)
{
    IAS_TRY(ensure_init());
    if (ctx().stream != ctx().own_stream) IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    ctx().stream = ctx().own_stream;
    return IAS_OK;
}

int ias_trim_pool(void)
{
    IAS_TRY(ensure_init());
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    block_cache_flush();
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    IAS_CUDA(cudaMemPoolTrimTo(ctx().pool, 0));          // cached blocks of earlier problem sizes go back to the driver
    return IAS_OK;
}

int ias_sync(void)
{
    IAS_TRY(ensure_init());
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

int ias_device_info(int *sm_count, size_t *smem_optin, size_t *free_bytes, size_t *total_bytes)
{
    IAS_TRY(ensure_init());
    if (sm_count) *sm_count = ctx().sm_count;
    if (smem_optin) *smem_optin = ctx().smem_optin;
    size_t f = 0, t = 0;
    IAS_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return IAS_OK;
}

// ---------------------------------------------------------------- transfers
int ias_upload_csr(const IasCsrMatrix *h, IasCsrMatrixDev *d)
{
    IAS_TRY(ensure_init());
    if (!h || !d || h->row < 0 || h->nnz < 0) return fail(IAS_E_ARG, "ias_upload_csr: bad argument");
    d->choice = true; d->row = h->row; d->col = h->col; d->nnz = h->nnz;
    d->row_ind_dev = nullptr; d->col_ind_dev = nullptr; d->values_dev = nullptr;
    DBuf<int> rp, ci;                                  // released automatically if a later step fails
    DBuf<double> v;
    IAS_TRY(rp.alloc((size_t)h->row + 1));
    IAS_TRY(ci.alloc((size_t)h->nnz));
    IAS_TRY(v.alloc((size_t)h->nnz));
    cudaStream_t s = ctx().stream;
    IAS_CUDA(cudaMemcpyAsync(rp.p, h->row_ind, sizeof(int) * ((size_t)h->row + 1), cudaMemcpyHostToDevice, s));
    if (h->nnz) {
        IAS_CUDA(cudaMemcpyAsync(ci.p, h->col_ind, sizeof(int) * (size_t)h->nnz, cudaMemcpyHostToDevice, s));
        IAS_CUDA(cudaMemcpyAsync(v.p, h->values, sizeof(double) * (size_t)h->nnz, cudaMemcpyHostToDevice, s));
    }
    IAS_CUDA(cudaStreamSynchronize(s));
    d->row_ind_dev = rp.release(); d->col_ind_dev = ci.release(); d->values_dev = v.release();
    return IAS_OK;
}

int ias_forget_operand(const IasCsrMatrixDev *m)
{
    Ctx &c = ctx();
    if (!m || c.canon_ci == (const void *)m->col_ind_dev) { c.canon_ci = c.canon_rp = nullptr; c.canon_rows = c.canon_nnz = -1; }
    return IAS_OK;
}

int ias_free_csr_dev(IasCsrMatrixDev *m)
{
    if (!m) return IAS_OK;
    ias_forget_operand(m);
    dfree(m->row_ind_dev); dfree(m->col_ind_dev); dfree(m->values_dev);
    m->row_ind_dev = nullptr; m->col_ind_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

int ias_free_csr64_dev(IasCsr64Dev *m)
{
    if (!m) return IAS_OK;
    dfree(m->row_ptr_dev); dfree(m->col_ind_dev); dfree(m->values_dev);
    m->row_ptr_dev = nullptr; m->col_ind_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

int ias_download_csr64(const IasCsr64Dev *d, long long *rp, int *ci, double *v)
{
    IAS_TRY(ensure_init());
    if (!d) return fail(IAS_E_ARG, "ias_download_csr64: NULL");
    cudaStream_t s = ctx().stream;
    if (rp) IAS_CUDA(cudaMemcpyAsync(rp, d->row_ptr_dev, sizeof(long long) * ((size_t)d->row + 1), cudaMemcpyDeviceToHost, s));
    if (ci && d->nnz) IAS_CUDA(cudaMemcpyAsync(ci, d->col_ind_dev, sizeof(int) * (size_t)d->nnz, cudaMemcpyDeviceToHost, s));
    if (v && d->nnz) IAS_CUDA(cudaMemcpyAsync(v, d->values_dev, sizeof(double) * (size_t)d->nnz, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    return IAS_OK;
}

int ias_download_csr(const IasCsrMatrixDev *d, int *rp, int *ci, double *v)
{
    IAS_TRY(ensure_init());
    if (!d) return fail(IAS_E_ARG, "ias_download_csr: NULL");
    cudaStream_t s = ctx().stream;
    if (rp) IAS_CUDA(cudaMemcpyAsync(rp, d->row_ind_dev, sizeof(int) * ((size_t)d->row + 1), cudaMemcpyDeviceToHost, s));
    if (ci && d->nnz) IAS_CUDA(cudaMemcpyAsync(ci, d->col_ind_dev, sizeof(int) * (size_t)d->nnz, cudaMemcpyDeviceToHost, s));
    if (v && d->nnz) IAS_CUDA(cudaMemcpyAsync(v, d->values_dev, sizeof(double) * (size_t)d->nnz, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    return IAS_OK;
}

int ias_copy(void *dst, const void *src, size_t bytes, int kind)
{
    IAS_TRY(ensure_init());
    if ((!dst || !src) && bytes) return fail(IAS_E_ARG, "ias_copy: NULL");
    if (kind < 0 || kind > 2) return fail(IAS_E_ARG, "ias_copy: kind must be 0 (h2d), 1 (d2h) or 2 (d2d)");
    static const cudaMemcpyKind kinds[3] = {cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice};
    if (bytes) IAS_CUDA(cudaMemcpyAsync(dst, src, bytes, kinds[kind], ctx().stream));
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

int ias_device_alloc(void **ptr, size_t bytes)
{
    IAS_TRY(ensure_init());
    if (!ptr) return fail(IAS_E_ARG, "ias_device_alloc: NULL");
    return device_block(ptr, bytes ? bytes : 1);
}

int ias_device_free(void *ptr)
{
    IAS_TRY(ensure_init());
    device_block_free(ptr);
    return IAS_OK;
}

// ---------------------------------------------------------------- verified_sum (csr_dev:258-273)
// Deterministic: cub's reduction tree is fixed for a given n.
int ias_checksum(const double *v, long long n, double *sum)
{
    IAS_TRY(ensure_init());
    if (!sum || n < 0) return fail(IAS_E_ARG, "ias_checksum: bad argument");
    *sum = 0.0;
    if (n == 0) return IAS_OK;
    DBuf<double> out;
    IAS_TRY(out.alloc(1));
    size_t tb = 0;
    IAS_CUDA(cub::DeviceReduce::Sum(nullptr, tb, v, out.p, n, ctx().stream));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, v, out.p, n, ctx().stream));
    ctx().launches += 2;
    IAS_CUDA(cudaMemcpyAsync(sum, out.p, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    return IAS_OK;
}

__global__ void k_is_canonical(int rows, const int *__restrict__ rp, const int *__restrict__ ci, int *bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int b = 0;
    for (int p = rp[i] + 1; p < rp[i + 1]; ++p) b |= (ci[p] <= ci[p - 1]);
    if (b) *bad = 1;
}

int ias_csr_is_canonical(const IasCsrMatrixDev *m, int *canonical)
{
    IAS_TRY(ensure_init());
    if (!m || !canonical) return fail(IAS_E_ARG, "ias_csr_is_canonical: NULL");
    DBuf<int> bad;
    IAS_TRY(bad.alloc(1));
    IAS_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx().stream));
    if (m->row > 0) IAS_LAUNCH(k_is_canonical, grid_for(m->row, 256), 256, 0, m->row, m->row_ind_dev, m->col_ind_dev, bad.p);
    int h = 0;
    IAS_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    IAS_CUDA(cudaStreamSynchronize(ctx().stream));
    *canonical = !h;
    return IAS_OK;
}

double ias_sizeof_csr(int rows, long long nnz) { return 4.0 * ((double)rows + 1 + (double)nnz + 3) + 8.0 * (double)nnz; }
double ias_sizeof_dia(int rows, int cols, int nd) { return 4.0 * ((double)rows + cols - 1 + nd + 3) + 8.0 * ((double)rows * nd); }
double ias_sizeof_ell(int rows, int w) { return 4.0 * ((double)rows + (double)rows * w + 4) + 8.0 * ((double)rows * w); }
double ias_sizeof_coo(int rows, long long nnz) { return 4.0 * ((double)rows + 1 + 2.0 * (double)nnz + 3) + 8.0 * (double)nnz; }

}  // extern "C"
