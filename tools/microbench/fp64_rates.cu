// Micro-benchmark (development tool): float64 issue rate, dependent latency and DMMA (mma.sync m8n8k4 f64) rate of one SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu && ./fp64_rates
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_tput(double* out, int iters, long long* cyc) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void dfma_lat(double* out, int iters, long long* cyc) {
    double a = threadIdx.x * 1e-3;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) a = fma(a, m, c);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void dmma_tput(double* out, int iters, long long* cyc) {
    double c0[4][2], a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 4; ++i) c0[i][0] = c0[i][1] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i][0]), "+d"(c0[i][1]) : "d"(a), "d"(b));
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i][0] + c0[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void dmma_lat(double* out, int iters, long long* cyc) {
    double c0 = 0, c1 = 0, a = threadIdx.x * 1e-3, b = 1.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void ffma_tput(float* out, int iters, long long* cyc) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float m = 1.0000001f, c = 1e-9f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMallocManaged(&cyc, 1024);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32}) {
        dfma_tput<<<1, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        printf("DFMA  %2d warps: %.2f lanes/clk/SM\n", warps, (double)warps * 32 * 8 * iters / cyc[0]);
        dmma_tput<<<1, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        printf("DMMA  %2d warps: %.2f FMA/clk/SM  (%.2f cycles per m8n8k4 per SM)\n", warps, (double)warps * 4 * 256 * iters / cyc[0], (double)cyc[0] / (warps * 4.0 * iters));
        ffma_tput<<<1, warps * 32>>>((float*)out, iters, cyc);
        cudaDeviceSynchronize();
        printf("FFMA  %2d warps: %.2f lanes/clk/SM\n", warps, (double)warps * 32 * 8 * iters / cyc[0]);
    }
    dfma_lat<<<1, 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    printf("DFMA dependent latency: %.1f cycles\n", (double)cyc[0] / iters);
    dmma_lat<<<1, 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    printf("DMMA dependent latency: %.1f cycles\n", (double)cyc[0] / iters);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
