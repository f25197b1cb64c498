// Microbenchmark: throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, and of FSEL / FMNMX mixed in.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c)
{ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE> __global__ void k(float *out, int iters, float seed)
{
    float a[16]; unsigned long long p[16];
    for (int i = 0; i < 16; i++) { a[i] = seed + i + threadIdx.x; p[i] = ((unsigned long long) __float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    const float b = 1.0000001f, c = 1e-9f;
    const unsigned long long b2 = ((unsigned long long) __float_as_uint(b) << 32) | __float_as_uint(b), c2 = ((unsigned long long) __float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fma1(a[i], b, c);
        } else if (MODE == 1) {     // 8 FFMA2 (same flops as mode 0)
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], b2, c2);
        } else if (MODE == 2) {     // 16 FFMA2 (twice the flops)
#pragma unroll
            for (int i = 0; i < 16; i++) p[i] = fma2(p[i], b2, c2);
        } else if (MODE == 3) {     // 8 FFMA2 + 8 FSEL/FMNMX (alu pipe) interleaved
#pragma unroll
            for (int i = 0; i < 8; i++) { p[i] = fma2(p[i], b2, c2); a[i] = fminf(a[i], a[i + 8] + 0.0f); a[i + 8] = (a[i] > 3.0f) ? a[i + 8] : b; }
        } else if (MODE == 4) {     // 16 scalar FFMA + 8 alu
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = fma1(a[i], b, c); a[i + 8] = fma1(a[i + 8], b, c); }
        }
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i] + __uint_as_float((unsigned) p[i]) + __uint_as_float((unsigned) (p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double instPerIter, double flopPerIter)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0f);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 8 * 8, winst = warps * iters * instPerIter;
    printf("%-28s %8.3f ms  %.2f warp-inst/clk/SM (at 1.965 GHz)  %.1f TFLOP/s\n", name, ms, winst / (ms * 1e-3) / 1.965e9 / 148, warps * 32 * iters * flopPerIter / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main()
{
    run<0>("16 FFMA", 16, 32); run<1>("8 FFMA2", 8, 32); run<2>("16 FFMA2", 16, 64); run<3>("8 FFMA2 + 8 FMNMX + 8 FSEL", 24, 32); run<4>("16 FFMA (2 chains)", 16, 32);
    return 0;
}
