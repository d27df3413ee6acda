// Dependent-chain latencies of the FP64 building blocks of the 8x8 pivot chain (cycles per op).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_bench tools/lat_bench.cu
#include <cstdio>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
__device__ __forceinline__ void pin(double& v) { asm volatile("" : "+d"(v)::"memory"); }
__device__ __forceinline__ double rcp_approx(double d) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
__device__ __forceinline__ double rsqrt_approx(double d) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
__device__ __forceinline__ double rcp_h(double d) { double w = rcp_approx(d); double e = fma(-d, w, 1.0); return fma(w, fma(e, e, e), w); }
__device__ __forceinline__ double rsqrt_h(double d) {
    double y = rsqrt_approx(d); double e = fma(-d * y, y, 1.0); double t = fma(0.375, e, 0.5); return fma(y, e * t, y);
}
#define CHAIN(NAME, EXPR)                                               \
    { double v = x; pin(v); long long t0 = clk();                       \
      _Pragma("unroll") for (int i = 0; i < 64; ++i) { v = EXPR; }      \
      pin(v); long long t1 = clk(); if (threadIdx.x == 0) printf("%-28s %6.1f cycles/op  (v=%g)\n", NAME, (t1 - t0) / 64.0, v); acc += v; }
__global__ void k(double x, double* out) {
    double acc = 0;
    CHAIN("dfma", fma(v, 1.0000001, 0.5))
    CHAIN("dmul", v * 1.0000001)
    CHAIN("dadd", v + 0.25)
    CHAIN("rsqrt(v)+1", rsqrt(v) + 1.0)
    CHAIN("1/v+1.5", 1.0 / v + 1.5)
    CHAIN("sqrt(v)+2", sqrt(v) + 2.0)
    CHAIN("rcp.approx+1 (MUFU.RCP64H)", rcp_approx(v) + 1.0)
    CHAIN("rsqrt.approx+1 (MUFU.RSQ64H)", rsqrt_approx(v) + 1.0)
    CHAIN("rcp_h(v)+1 (seed+cubic)", rcp_h(v) + 1.0)
    CHAIN("rsqrt_h(v)+1 (seed+cubic)", rsqrt_h(v) + 1.0)
    CHAIN("float rsqrtf roundtrip+1", (double)rsqrtf((float)v) + 1.0)
    { // accuracy of the cubic-corrected seeds
      double worst_r = 0, worst_s = 0;
      for (int i = 0; i < 2000; ++i) { double d = 0.001 + 0.37 * i + 1e-3 * threadIdx.x; double a = rcp_h(d) * d - 1.0; double b = rsqrt_h(d); b = b * b * d - 1.0;
        worst_r = fmax(worst_r, fabs(a)); worst_s = fmax(worst_s, fabs(b)); }
      if (threadIdx.x == 0) printf("max |rcp_h(d) d - 1| = %.3g   max |rsqrt_h(d)^2 d - 1| = %.3g\n", worst_r, worst_s);
    }
    out[threadIdx.x] = acc;
}
int main() { double* o; cudaMalloc(&o, 256); k<<<1, 32>>>(1.3, o); cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError())); return 0; }
