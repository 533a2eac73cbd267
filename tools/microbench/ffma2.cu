// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/ffma2.bin tools/microbench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s)
{
    float a[8];
    u64 p[8];
    for (int j = 0; j < 8; ++j) { a[j] = s + j + threadIdx.x; p[j] = pk(a[j], a[j] + 1.f); }
    const float m = 1.0000001f, c = 1e-7f;
    const u64 m2 = pk(m, m), c2 = pk(c, c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = fmaf(a[j], m, c);
            else if (MODE == 1) p[j] = fma2(p[j], m2, c2);
            else { a[j] = fmaf(a[j], m, c); p[j] = fma2(p[j], m2, c2); }
        }
    }
    float r = 0;
    for (int j = 0; j < 8; ++j) r += a[j] + (float)(p[j] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> double run(float *d, int blocks, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, iters, 1.f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)blocks * 256 * iters * 8 * (MODE == 2 ? 2 : 1);
    return inst / (ms * 1e-3);
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int blocks = pr.multiProcessorCount * 8, iters = 20000;
    float *d; cudaMalloc(&d, sizeof(float) * blocks * 256);
    printf("FFMA   : %.2f T lane-instr/s\n", run<0>(d, blocks, iters) / 1e12);
    printf("FFMA2  : %.2f T lane-instr/s (each instruction = 2 FMAs per lane)\n", run<1>(d, blocks, iters) / 1e12);
    printf("mixed  : %.2f T lane-instr/s (FFMA + FFMA2 interleaved)\n", run<2>(d, blocks, iters) / 1e12);
    return 0;
}
