// Microbenchmark: issue throughput of packed fp32 (FADD2/FFMA2, PTX add/fma.f32x2) vs scalar FADD/FFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 f32x2_bench.cu -o f32x2_bench && ./f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed)
{
    // 8 independent chains per thread
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    u64 p[8];
    for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
    const u64 c = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { // scalar FADD x16
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = a[i] + seed;
            } else if (MODE == 1) { // FADD2 x8 (same flops as mode 0)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = add2(p[i], c);
            } else if (MODE == 2) { // scalar FFMA x16
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], seed, 1.0f);
            } else { // FFMA2 x8
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], c, c);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* d)
{
    const int iters = 2000, grid = 148, block = 512;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(d, 10, 1.5f);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, iters, 1.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop_lane_ops = (double)grid * block * iters * 8 * 16; // fp32 lane operations
    const double winst = (double)grid * (block / 32) * iters * 8 * (MODE % 2 == 0 ? 16 : 8);
    printf("%-8s %.3f ms  %.1f G lane-ops/s  %.2f warp-inst/clk/SM (at 1.965 GHz)\n", name, ms, flop_lane_ops / ms / 1e6,
           winst / (ms * 1e-3) / 1.965e9 / 148);
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 512 * 4);
    run<0>("FADD", d); run<1>("FADD2", d); run<2>("FFMA", d); run<3>("FFMA2", d);
    return 0;
}
