// L1 data-pipe cost model on sm_100a: how many cycles does one warp-wide load take in the l1tex data stage as a
// function of load width and of how the 32 lanes' addresses fall on 128-byte lines?  (ncu had the render kernel at
// 81-87 % of l1tex__data_pipe_lsu_wavefronts: the pipe, not instruction issue, bounds it.)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/l1pipe.bin tools/microbench/l1pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

// PATTERN: 0 = every lane its own 128-B line; 1 = 4 consecutive lanes share a line (32 B each); 2 = all lanes one line (broadcast-ish,
// 4 B apart);  3 = every lane its own line, but lane reads 4 x 32 B of that line in consecutive instructions
template <int WIDTH, int PATTERN>
__global__ void __launch_bounds__(256) k(const float4 *__restrict__ buf, float *out, int iters, int lines)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    unsigned line = (warp * 37 + lane * 5) % lines;
    for (int i = 0; i < iters; ++i) {
        line = (line * 13 + 7) % lines;
        unsigned l = PATTERN == 1 ? (line & ~3u) + 0 : (PATTERN == 2 ? (line - lane * 5 + 1000 * lines) % lines : line);
        if (PATTERN == 1) l = ((warp * 37 + (lane >> 2) * 5 + i * 13) % lines);
        const float4 *p = buf + (size_t)l * 8 + (PATTERN == 1 ? (lane & 3) * 2 : 0);
        if (WIDTH == 32) {
            float4 a, b;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
            acc += a.x + b.w;
            if (PATTERN == 3) {
#pragma unroll
                for (int q = 1; q < 4; ++q) {
                    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 2 * q));
                    acc += a.x + b.w;
                }
            }
        }
        else if (WIDTH == 16) {
            float4 a;
            asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
            acc += a.x + a.w;
        }
        else {
            float a;
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(a) : "l"((const float *)p + (PATTERN == 2 ? lane : 0)));
            acc += a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WIDTH, int PATTERN> void run(const char *name, const float4 *buf, float *out, int sms, double ghz)
{
    const int blocks = sms * 4, iters = 4000, lines = 256; // 256 lines = 32 KB: L1-resident
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WIDTH, PATTERN><<<blocks, 256>>>(buf, out, iters, lines);
    cudaEventRecord(e0);
    k<WIDTH, PATTERN><<<blocks, 256>>>(buf, out, iters, lines);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_loads = (double)blocks * 8 * iters * (PATTERN == 3 ? 4 : 1);
    const double cyc_per_sm = ms * 1e-3 * ghz * 1e9;
    printf("%-58s %7.2f cycles per warp-load per SM   %7.1f B/clk/SM\n", name, cyc_per_sm / (warp_loads / sms), warp_loads / sms * 32 * WIDTH / cyc_per_sm);
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float4 *buf; float *out;
    cudaMalloc(&buf, 256 * 128); cudaMemset(buf, 0, 256 * 128);
    cudaMalloc(&out, sizeof(float) * pr.multiProcessorCount * 4 * 256);
    printf("%s, %d SMs, %.3f GHz\n", pr.name, pr.multiProcessorCount, ghz);
    run<4, 2>("LDG.32  all lanes in one line (coalesced 128 B)", buf, out, pr.multiProcessorCount, ghz);
    run<4, 0>("LDG.32  every lane its own line", buf, out, pr.multiProcessorCount, ghz);
    run<16, 0>("LDG.128 every lane its own line", buf, out, pr.multiProcessorCount, ghz);
    run<32, 0>("LDG.256 every lane its own line", buf, out, pr.multiProcessorCount, ghz);
    run<32, 1>("LDG.256 4 lanes share a line (quad reads one 128-B node)", buf, out, pr.multiProcessorCount, ghz);
    run<32, 3>("LDG.256 x4 of the lane's own line (a 128-B node per lane)", buf, out, pr.multiProcessorCount, ghz);
    return 0;
}
