// Microbenchmark: shared-memory random-access primitives on B200 (sm_100a).
// Decides how the global-row SpGEMM kernels may touch their shared-memory accumulators:
// native ATOMS.OR / ATOMS.CAS, the fp64 atomicAdd CAS loop, or plain LDS / STS.
// Prints warp-level operations per clock and SM for one resident CTA of 1024 threads per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned rng(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(1024) k(int iters, int span, unsigned long long *clk_out, double *sink)
{
    extern __shared__ __align__(16) unsigned char raw[];
    unsigned *w32 = reinterpret_cast<unsigned *>(raw);
    double *f64 = reinterpret_cast<double *>(raw);
    for (int i = threadIdx.x; i < span * 2; i += blockDim.x) w32[i] = 0;
    __syncthreads();
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    double accum = 0.0;
    unsigned acc_u = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = rng(s) % (unsigned)span;
        if (MODE == 0) {            // ATOMS.OR (no return)
#pragma unroll
            for (int u = 0; u < 4; ++u) atomicOr(&w32[a[u]], 1u << (a[u] & 31));
        } else if (MODE == 1) {     // check first, then ATOMS.OR (bits fill up quickly: mostly LDS)
#pragma unroll
            for (int u = 0; u < 4; ++u) { unsigned bit = 1u << (s >> (27 - u) & 31); if (!(w32[a[u]] & bit)) atomicOr(&w32[a[u]], bit); }
        } else if (MODE == 2) {     // ATOMS.CAS 32
#pragma unroll
            for (int u = 0; u < 4; ++u) acc_u += atomicCAS(&w32[a[u]], 0u, a[u] + 1);
        } else if (MODE == 3) {     // fp64 atomicAdd (CAS loop)
#pragma unroll
            for (int u = 0; u < 4; ++u) atomicAdd(&f64[a[u]], 1.0);
        } else if (MODE == 4) {     // plain LDS.64 + DADD + STS.64 (racy, throughput only)
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = f64[a[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u) f64[a[u]] = v[u] + 1.0;
        } else if (MODE == 5) {     // LDS.32 random
#pragma unroll
            for (int u = 0; u < 4; ++u) acc_u += w32[a[u]];
        } else if (MODE == 6) {     // STS.32 random
#pragma unroll
            for (int u = 0; u < 4; ++u) w32[a[u]] = a[u];
        } else if (MODE == 7) {     // LDS.64 random
#pragma unroll
            for (int u = 0; u < 4; ++u) accum += f64[a[u]];
        } else if (MODE == 8) {     // address generation only
#pragma unroll
            for (int u = 0; u < 4; ++u) acc_u += a[u];
        } else if (MODE == 9) {     // ATOMS.ADD u32
#pragma unroll
            for (int u = 0; u < 4; ++u) atomicAdd(&w32[a[u]], 1u);
        } else if (MODE == 10) {    // fp32 atomicAdd
            float *f32 = reinterpret_cast<float *>(raw);
#pragma unroll
            for (int u = 0; u < 4; ++u) atomicAdd(&f32[a[u]], 1.0f);
        } else if (MODE == 11) {    // u64 atomicAdd
            unsigned long long *u64 = reinterpret_cast<unsigned long long *>(raw);
#pragma unroll
            for (int u = 0; u < 4; ++u) atomicAdd(&u64[a[u]], 1ull);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc_u == 0xdeadbeefu || accum == 1.2345) sink[0] = accum + acc_u + f64[threadIdx.x];
}

template <int MODE>
static void run(const char *name, int span)
{
    int dev_sms = 148;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); dev_sms = p.multiProcessorCount;
    unsigned long long *clk; double *sink;
    cudaMalloc(&clk, dev_sms * 8); cudaMalloc(&sink, 8);
    size_t sm = (size_t)span * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const int iters = 2000;
    k<MODE><<<dev_sms, 1024, sm>>>(10, span, clk, sink);
    k<MODE><<<dev_sms, 1024, sm>>>(iters, span, clk, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, clk, dev_sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < dev_sms; ++i) avg += (double)h[i]; avg /= dev_sms;
    double warp_ops = 32.0 * iters * 4;      // per SM
    printf("%-34s span %6d  clk/warp-op %7.2f   lanes/clk/SM %6.2f   (%s)\n", name, span, avg / warp_ops, 32.0 * warp_ops / avg, cudaGetErrorString(e));
    cudaFree(clk); cudaFree(sink);
}

int main()
{
    for (int span : {1024, 12288, 24576}) {
        run<8>("address generation only", span);
        run<5>("LDS.32 random", span);
        run<7>("LDS.64 random", span);
        run<6>("STS.32 random", span);
        run<4>("LDS.64+DADD+STS.64 (non-atomic)", span);
        run<0>("ATOMS.OR", span);
        run<1>("LDS check, then ATOMS.OR", span);
        run<9>("ATOMS.ADD u32", span);
        run<2>("ATOMS.CAS 32", span);
        run<10>("atomicAdd fp32", span);
        run<11>("atomicAdd u64", span);
        run<3>("atomicAdd fp64 (CAS loop)", span);
    }
    return 0;
}
