// Phase timing of the in-shared-memory 128x128 Cholesky + inverse (block_chol.cuh) with clock64().
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I. -o tools/leaf_prof tools/leaf_prof.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../portfoliooptgp_b200/csrc/block_chol.cuh"
using namespace gpb;

__global__ void __launch_bounds__(512) prof_kernel(const double* A, int n, long long* stamps, double* out) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm; double* T = sm + 128 * SLD; double* dinv = T + 64 * TLD; int* fail = (int*)(dinv + DINV_DOUBLES);
    const int tid = threadIdx.x, nt = blockDim.x;
    long long t0 = clock64();
    for (int idx = tid; idx < n * 128; idx += nt) { int i = idx >> 7, j = idx & 127; if (j < n) S[i * SLD + j] = (j <= i) ? A[i * n + j] : 0.0; }
    __syncthreads();
    long long t1 = clock64();
    block_potrf_lower(S, n, fail, dinv);
    long long t2 = clock64();
    block_trtri_lower_inplace(S, n, T, dinv);
    long long t3 = clock64();
    for (int idx = tid; idx < n * 128; idx += nt) { int i = idx >> 7, j = idx & 127; if (j < n) out[i * n + j] = S[i * SLD + j]; }
    __syncthreads();
    long long t4 = clock64();
    if (tid == 0) { stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = t3 - t2; stamps[3] = t4 - t3; }
}

// isolated micro-latencies
__global__ void lat_kernel(long long* out, double x) {
    double v = x; long long t0 = clock64();
    for (int i = 0; i < 64; ++i) v = rsqrt(v + 1.0);
    long long t1 = clock64();
    for (int i = 0; i < 64; ++i) v = 1.0 / (v + 1.5);
    long long t2 = clock64();
    double c0 = v, c1 = v;
    for (int i = 0; i < 64; ++i) dmma_8x8x4(c0, c1, v, x);
    long long t3 = clock64();
    for (int i = 0; i < 64; ++i) v = sqrt(v + 2.0);
    long long t4 = clock64();
    for (int i = 0; i < 64; ++i) v = __shfl_sync(0xffffffffu, v, (i * 7) & 31) + 1.0;
    long long t5 = clock64();
    for (int i = 0; i < 64; ++i) v = fma(v, 1.0000001, 0.5);
    long long t6 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = (t3 - t2) / 64; out[3] = (t4 - t3) / 64; out[4] = (t5 - t4) / 64; out[5] = (t6 - t5) / 64; out[6] = (long long)(v + c0 + c1); }
}

int main() {
    const int n = 128;
    std::vector<double> G(n * n), A(n * n);
    for (int i = 0; i < n * n; ++i) G[i] = sin(0.37 * i) ;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += G[i * n + k] * G[j * n + k]; A[i * n + j] = s / n + (i == j ? 0.5 : 0.0); }
    double *dA, *dO; long long* dS;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dO, n * n * 8); cudaMalloc(&dS, 64 * 8);
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    size_t smem = (128 * SLD + 64 * TLD + DINV_DOUBLES + 16) * 8;
    cudaFuncSetAttribute(prof_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int threads : {256, 512}) {
        for (int rep = 0; rep < 2; ++rep) prof_kernel<<<1, threads, smem>>>(dA, n, dS, dO);
        long long st[4]; cudaMemcpy(st, dS, 32, cudaMemcpyDeviceToHost);
        printf("threads %d cycles: load %lld potrf %lld trtri %lld store %lld (err %s)\n", threads, st[0], st[1], st[2], st[3], cudaGetErrorString(cudaGetLastError()));
    }
    lat_kernel<<<1, 32>>>(dS, 1.3);
    long long l[7]; cudaMemcpy(l, dS, 56, cudaMemcpyDeviceToHost);
    printf("latency cycles: rsqrt+add %lld  div+add %lld  dmma-chain %lld  sqrt+add %lld  shfl+add %lld  dfma %lld\n", l[0], l[1], l[2], l[3], l[4], l[5]);
    return 0;
}
