// green_probe.cu -- can the latency-bound chain of the blocked factorisation run on a small green-context SM
// partition while bulk GEMM-like grids saturate the rest?  Creates an 8-SM and a (rest)-SM green context
// (CUDA >= 12.4 driver API), streams in each, launches RUNTIME-API kernels into them, records %smid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o green_probe green_probe.cu -lcuda
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("driver error %s at line %d\n", s_, __LINE__); return 1; } } while (0)

__global__ void busy_kernel(long long cycles, double* sink, unsigned* smids) {
    extern __shared__ double sm[];
    if (threadIdx.x == 0 && smids) {
        unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
        atomicOr(&smids[id / 32], 1u << (id % 32));
    }
    const long long t0 = clock64();
    double acc = threadIdx.x;
    while (clock64() - t0 < cycles) acc = fma(acc, 1.0000001, 1e-9);
    if (acc == 12345.678) sink[0] = acc + sm[0];
}

static int popcount_mask(const unsigned* m, int words) { int c = 0; for (int i = 0; i < words; ++i) c += __builtin_popcount(m[i]); return c; }

int main() {
    CK(cudaSetDevice(0)); CK(cudaFree(0));
    CU(cuInit(0));
    CUdevice dev; CU(cuDeviceGet(&dev, 0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const double ghz = p.clockRate * 1e-6; const long long us = (long long)(ghz * 1e3);
    CUdevResource all; CU(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs %u\n", all.sm.smCount);
    CUdevResource small_[1], rest; unsigned nb = 1;
    CU(cuDevSmResourceSplitByCount(small_, &nb, &all, &rest, 0, 8));
    printf("split: groups %u, small %u SMs, remaining %u SMs\n", nb, small_[0].sm.smCount, rest.sm.smCount);
    CUdevResourceDesc d_small, d_rest;
    CU(cuDevResourceGenerateDesc(&d_small, &small_[0], 1));
    CU(cuDevResourceGenerateDesc(&d_rest, &rest, 1));
    CUgreenCtx g_small, g_rest;
    CU(cuGreenCtxCreate(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CU(cuGreenCtxCreate(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream s_small, s_rest;
    CU(cuGreenCtxStreamCreate(&s_small, g_small, CU_STREAM_NON_BLOCKING, 0));
    CU(cuGreenCtxStreamCreate(&s_rest, g_rest, CU_STREAM_NON_BLOCKING, 0));
    cudaStream_t cs = (cudaStream_t)s_small, bs = (cudaStream_t)s_rest;
    double* sink; CK(cudaMalloc(&sink, 64));
    unsigned* smids; CK(cudaMalloc(&smids, 2 * 8 * sizeof(unsigned))); CK(cudaMemset(smids, 0, 2 * 8 * sizeof(unsigned)));
    CK(cudaFuncSetAttribute(busy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1, ex; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&ex, cudaEventDisableTiming));
    // which SMs does each partition use
    busy_kernel<<<2000, 256, 110 * 1024, bs>>>(5 * us, sink, smids);
    busy_kernel<<<200, 256, 110 * 1024, cs>>>(5 * us, sink, smids + 8);
    CK(cudaDeviceSynchronize());
    unsigned h[16]; CK(cudaMemcpy(h, smids, sizeof(h), cudaMemcpyDeviceToHost));
    unsigned overlap = 0; for (int i = 0; i < 8; ++i) overlap |= h[i] & h[8 + i];
    printf("bulk partition ran on %d SMs, small partition on %d SMs, overlap mask %s\n", popcount_mask(h, 8), popcount_mask(h + 8, 8), overlap ? "NON-EMPTY" : "empty");
    const int chain_len = 20;
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int ctas = cfg ? 16 : 1, smem = cfg ? 70 : 182, dur = cfg ? 8 : 30;
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, cs));
        for (int i = 0; i < chain_len; ++i) busy_kernel<<<ctas, 512, smem * 1024, cs>>>(dur * us, sink, nullptr);
        CK(cudaEventRecord(e1, cs)); CK(cudaEventSynchronize(e1));
        float alone; CK(cudaEventElapsedTime(&alone, e0, e1));
        const int nct = (int)(6000.0 / 150 * rest.sm.smCount * 2);
        busy_kernel<<<nct, 256, 110 * 1024, bs>>>(150 * us, sink, nullptr);
        busy_kernel<<<1, 32, 0, cs>>>(200 * us, sink, nullptr);
        CK(cudaEventRecord(e0, cs));
        for (int i = 0; i < chain_len; ++i) busy_kernel<<<ctas, 512, smem * 1024, cs>>>(dur * us, sink, nullptr);
        CK(cudaEventRecord(e1, cs)); CK(cudaEventSynchronize(e1));
        float under; CK(cudaEventElapsedTime(&under, e0, e1));
        // cross-partition event dependency: bulk stream waits on the chain
        CK(cudaEventRecord(ex, cs)); CK(cudaStreamWaitEvent(bs, ex, 0));
        busy_kernel<<<8, 256, 0, bs>>>(5 * us, sink, nullptr);
        CK(cudaDeviceSynchronize());
        printf("chain %d CTA x %d KB x %d us: alone %.3f ms, beside a saturating bulk grid in the other partition %.3f ms (%.1f us extra per kernel)\n",
               ctas, smem, dur, alone, under, (under - alone) * 1e3 / chain_len);
    }
    // bulk throughput on the reduced partition vs the whole device (primary context stream)
    cudaStream_t ps; CK(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    for (int which = 0; which < 2; ++which) {
        cudaStream_t st = which ? bs : ps;
        const int nct = 148 * 2 * 20;
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, st));
        busy_kernel<<<nct, 256, 110 * 1024, st>>>(150 * us, sink, nullptr);
        CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%s: %d CTAs x 150 us, 2 per SM: %.3f ms\n", which ? "rest partition" : "whole device ", nct, ms);
    }
    printf("ok\n");
    return 0;
}
