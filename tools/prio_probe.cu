// prio_probe.cu -- does a high-priority stream's small dependent kernels (one CTA, large shared memory, like the
// 128x128 Cholesky leaf; or 16 CTAs of 70 KB like the small GEMMs) get SM slots promptly while a low-priority
// bulk grid (GEMM-like CTAs, 1 or 2 resident per SM) saturates the machine?  Decides whether the latency-bound
// chain of the blocked factorisation can be hidden behind bulk products by stream priorities alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prio_probe prio_probe.cu && ./prio_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__global__ void busy_kernel(long long cycles, double* sink) {
    extern __shared__ double sm[];
    const long long t0 = clock64();
    double acc = threadIdx.x;
    while (clock64() - t0 < cycles) acc = fma(acc, 1.0000001, 1e-9);
    if (acc == 12345.678) sink[0] = acc + sm[0];
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    printf("device %s sms %d clock %.2f GHz priority range least %d greatest %d\n", p.name, sms, ghz, lo, hi);
    cudaStream_t bulk, crit, crit_same;
    CK(cudaStreamCreateWithPriority(&bulk, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&crit, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&crit_same, cudaStreamNonBlocking, lo));
    double* sink; CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(busy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const long long us = (long long)(ghz * 1e3);
    struct Cfg { const char* name; int bulk_smem_kb; int bulk_threads; int bulk_us; int chain_ctas; int chain_smem_kb; int chain_us; };
    const Cfg cfgs[] = {
        {"bulk 2 CTA/SM (110 KB, 150 us) | chain 1 CTA x 182 KB x 30 us", 110, 256, 150, 1, 182, 30},
        {"bulk 1 CTA/SM (123 KB, 300 us) | chain 1 CTA x 182 KB x 30 us", 123, 256, 300, 1, 182, 30},
        {"bulk 2 CTA/SM (110 KB, 150 us) | chain 16 CTA x 70 KB x 8 us", 110, 256, 150, 16, 70, 8},
        {"bulk 1 CTA/SM (123 KB, 300 us) | chain 16 CTA x 70 KB x 8 us", 123, 256, 300, 16, 70, 8},
        {"bulk 2 CTA/SM (110 KB, 150 us) | chain 1 CTA x 100 KB x 30 us", 110, 256, 150, 1, 100, 30},
        {"bulk 2 CTA/SM (110 KB, 40 us)  | chain 1 CTA x 182 KB x 30 us", 110, 256, 40, 1, 182, 30},
    };
    for (const Cfg& c : cfgs) {
        for (int prio = 0; prio < 2; ++prio) {
            cudaStream_t cs = prio ? crit : crit_same;
            const int chain_len = 20;
            // chain alone
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, cs));
            for (int i = 0; i < chain_len; ++i) busy_kernel<<<c.chain_ctas, 512, c.chain_smem_kb * 1024, cs>>>(c.chain_us * us, sink);
            CK(cudaEventRecord(e1, cs));
            CK(cudaEventSynchronize(e1));
            float alone; CK(cudaEventElapsedTime(&alone, e0, e1));
            // bulk: enough CTAs for ~6 ms, then the chain
            const int per_sm = (c.bulk_smem_kb > 113) ? 1 : 2;
            const int nct = (int)(6000.0 / c.bulk_us * sms * per_sm);
            busy_kernel<<<nct, c.bulk_threads, c.bulk_smem_kb * 1024, bulk>>>(c.bulk_us * us, sink);
            busy_kernel<<<1, 32, 0, cs>>>(200 * us, sink);   // let the bulk grid fill the machine first
            CK(cudaEventRecord(e0, cs));
            for (int i = 0; i < chain_len; ++i) busy_kernel<<<c.chain_ctas, 512, c.chain_smem_kb * 1024, cs>>>(c.chain_us * us, sink);
            CK(cudaEventRecord(e1, cs));
            CK(cudaEventSynchronize(e1));
            float under; CK(cudaEventElapsedTime(&under, e0, e1));
            CK(cudaDeviceSynchronize());
            printf("%s | %s priority: chain of %d alone %.3f ms, under bulk load %.3f ms (%.1f us extra per kernel)\n", c.name,
                   prio ? "HIGH" : "same", chain_len, alone, under, (under - alone) * 1e3 / chain_len);
        }
    }
    return 0;
}
