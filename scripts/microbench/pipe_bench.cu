// Microbenchmark: issue throughput (warp instructions / clk / SM) of the instruction kinds the tile force kernel uses on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 16
template <int MODE> __global__ void k(float *out, int iters, float seed)
{
    float a[CHAINS]; unsigned int u[CHAINS];
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + i + threadIdx.x * 0.001f; u[i] = threadIdx.x * 17 + i; }
    const float b = 1.0000001f, c = 1e-9f;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (MODE == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
            if (MODE == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 3) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 4) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(a[i]) : "f"(c), "f"(b));   // FSETP + FSEL
            if (MODE == 5) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b), "r"(it & 1));
            if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(0x5555u));
            if (MODE == 7) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
            if (MODE == 8) a[i] = __shfl_sync(0xffffffffu, a[i], (lane + 1) & 31);
            if (MODE == 9) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 10) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(3));
            if (MODE == 11) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, 0; selp.u32 %0, %0, %1, p;}" : "+r"(u[i]) : "r"(7u));
            if (MODE == 13) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); if ((i & 3) == 3) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b), "r"(it & 1)); }
            if (MODE == 14) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); if ((i & 1) == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(0x5555u)); }
            if (MODE == 15) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); if ((i & 3) == 3) a[i] = __shfl_sync(0xffffffffu, a[i], (lane + 1) & 31); }
            if (MODE == 16) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); if ((i & 1) == 1) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b), "r"(it & 1)); }
            if (MODE == 12) { double d = (double) a[i]; asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(1e-9)); a[i] = (float) d; }
        }
    }
    float s = 0; for (int i = 0; i < CHAINS; i++) s += a[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double instPerOp)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0f);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 8 * 8, winst = warps * iters * CHAINS * instPerOp;
    printf("%-22s %8.3f ms  %.2f warp-inst/clk/SM\n", name, ms, winst / (ms * 1e-3) / 1.965e9 / 148);
    cudaFree(out);
}
int main()
{
    run<0>("FFMA", 1); run<1>("FADD", 1); run<2>("FMUL", 1); run<3>("FMNMX", 1); run<4>("FSETP+FSEL", 2); run<5>("FSEL (+setp.ne)", 1);
    run<6>("LOP3", 1); run<7>("SHL", 1); run<8>("SHFL", 1); run<9>("MUFU.RSQ", 1); run<10>("IADD", 1); run<11>("ISETP+SEL", 2); run<12>("F2F+DADD+F2F", 3);
    run<13>("4 FFMA : 1 FSEL", 1.25); run<14>("2 FFMA : 1 LOP3", 1.5); run<15>("4 FFMA : 1 SHFL", 1.25); run<16>("2 FFMA : 1 FSEL", 1.5);
    return 0;
}
