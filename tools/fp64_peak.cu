// FP64 peak calibration for B200 (sm_100a): DFMA issue rate and DMMA (mma.sync f64) rate for
// each legal shape.  Register-only loops: no memory traffic, so the numbers are pipe ceilings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_m8n8k4(double* out, int iters, double a, double b) {
    double c[NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_m16n8k4(double* out, int iters, double a, double b) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_m16n8k8(double* out, int iters, double a, double b) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_m16n8k16(double* out, int iters, double a, double b) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                           "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // a launch that fails (e.g. 1024 threads of a kernel that needs > 64 registers each) must not be
    // printed as a measurement: every launch and the final sync are checked, failure returns a negative time
    launch(); launch();
    if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return -1.f; }
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch();
        if (cudaGetLastError() != cudaSuccess) return -1.f;
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

static void report(const char* what, int warps, float ms, double flop) {
    if (ms <= 0.f) printf("%s warps/cta %2d: LAUNCH FAILED (not a measurement)\n", what, warps);
    else printf("%s warps/cta %2d: %.2f TFLOP/s\n", what, warps, flop / ms / 1e9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d\n", p.name, sms);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32; int blocks = sms * 2;
        double nthr = (double)blocks * threads;
        float ms = time_it([&] { k_dfma<16><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        report("dfma ilp16 (2 cta/sm)", warps, ms, 2.0 * 16 * iters * nthr);
        double nwarp = (double)blocks * warps;
        ms = time_it([&] { k_m8n8k4<16><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        report("dmma m8n8k4  ", warps, ms, 2.0 * 8 * 8 * 4 * 16 * iters * nwarp);
        ms = time_it([&] { k_m16n8k4<8><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        report("dmma m16n8k4 ", warps, ms, 2.0 * 16 * 8 * 4 * 8 * iters * nwarp);
        ms = time_it([&] { k_m16n8k8<8><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        report("dmma m16n8k8 ", warps, ms, 2.0 * 16 * 8 * 8 * 8 * iters * nwarp);
        ms = time_it([&] { k_m16n8k16<8><<<blocks, threads>>>(out, iters, 1.000001, 1e-9); });
        report("dmma m16n8k16", warps, ms, 2.0 * 16 * 8 * 16 * 8 * iters * nwarp);
    }
    cudaFree(out);
    return 0;
}
