// Microbenchmark (round 2): which instructions share a pipe on sm_100a?  Pairs of instruction kinds are interleaved 1:1 on
// independent register chains; the time per pair tells whether they overlap (different pipes) or add (same pipe / issue bound).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_bench mix_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
typedef unsigned long long u64;

enum Op { FFMA, FFMA2, FADD2, FMUL2, FMNMX, FSEL, FSETP_SEL, FSET, LOP3, IADD3, IMAD, SHF, MUFU, SHFL, LDS128B, LDS128G, LDS64B, LDS32G, PRMT, FMNMX3, NONE };

template <int OP> __device__ __forceinline__ void op(float &a, u64 &p, unsigned &u, float b, float c, u64 b2, u64 c2, int it, const float4 *sm, int lane)
{
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
    if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p) : "l"(b2), "l"(c2));
    if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(c2));
    if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(b2));
    if (OP == FMNMX) { asm volatile("min.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a) : "f"(c)); }
    if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
    if (OP == FSEL) asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(a) : "f"(b), "r"(u & 1));   // ISETP + FSEL (u varies per chain)
    if (OP == FSETP_SEL) asm volatile("{.reg .pred q; setp.gt.f32 q, %0, %1; selp.f32 %0, %0, %2, q;}" : "+f"(a) : "f"(c), "f"(b));
    if (OP == FSET) asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(a) : "f"(c));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(it), "r"(0x5555u));
    if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(u) : "r"(it));
    if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u) : "r"(it), "r"(3));
    if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(u) : "r"(u), "r"(it));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u) : "r"(it));
    if (OP == MUFU) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a));
    if (OP == SHFL) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
    if (OP == LDS128B) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned) __cvta_generic_to_shared(sm + (u & 63)))); a += v.x; }
    if (OP == LDS128G) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned) __cvta_generic_to_shared(sm + ((u + lane) & 63)))); a += v.x; }
    if (OP == LDS64B) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned) __cvta_generic_to_shared(reinterpret_cast<const float2 *>(sm) + (u & 63)))); a += v.x; }
    if (OP == LDS32G) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned) __cvta_generic_to_shared(reinterpret_cast<const float *>(sm) + ((u * 7 + lane) & 255)))); a += v; }
}

template <int A, int NA, int B, int NB> __global__ void k(float *out, int iters, float seed)
{
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float a[CH], a2[CH]; u64 p[CH], p2[CH]; unsigned u[CH], u2[CH];
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < CH; i++) { a[i] = seed + i + threadIdx.x * 0.001f; a2[i] = a[i] * 1.5f; u[i] = threadIdx.x * 17 + i; u2[i] = u[i] * 3;
        p[i] = ((u64) __float_as_uint(a[i]) << 32) | __float_as_uint(a2[i]); p2[i] = p[i] + 12345; }
    const float b = 1.0000001f, c = 1e-9f;
    const u64 b2 = ((u64) __float_as_uint(b) << 32) | __float_as_uint(b), c2 = ((u64) __float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
#pragma unroll
            for (int r = 0; r < NA; r++) op<A>(a[i], p[i], u[i], b, c, b2, c2, it, sm, lane);
#pragma unroll
            for (int r = 0; r < NB; r++) op<B>(a2[i], p2[i], u2[i], b, c, b2, c2, it, sm, lane);
        }
    }
    float s = 0;
    for (int i = 0; i < CH; i++) s += a[i] + a2[i] + u[i] + u2[i] + __uint_as_float((unsigned) p[i]) + __uint_as_float((unsigned) (p2[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int A, int NA, int B, int NB> void run(const char *name)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    k<A, NA, B, NB><<<148 * 8, 256>>>(out, 50, 1.0f);
    cudaEventRecord(e0); k<A, NA, B, NB><<<148 * 8, 256>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // SMSP cycles per (NA x A + NB x B) group: 16 warps per SMSP
    const double groups = 16.0 * iters * CH;
    printf("%-34s %8.3f ms   %.2f SMSP-cycles per group (%d + %d instr)\n", name, ms, ms * 1e-3 * 1.965e9 / groups, NA, NB);
    cudaFree(out);
}
#define R1(A) run<A, 1, NONE, 0>(#A)
#define R2(A, NA, B, NB) run<A, NA, B, NB>(#A " x" #NA " + " #B " x" #NB)
int main()
{
    R1(FFMA); R1(FFMA2); R1(FADD2); R1(FMUL2); R1(FMNMX); R1(FMNMX3); R1(FSEL); R1(FSETP_SEL); R1(FSET); R1(LOP3); R1(IADD3); R1(IMAD); R1(SHF); R1(PRMT); R1(MUFU); R1(SHFL);
    R1(LDS128B); R1(LDS128G); R1(LDS64B); R1(LDS32G);
    R2(FFMA2, 1, LOP3, 1); R2(FFMA2, 1, FSEL, 1); R2(FFMA2, 1, FMNMX, 1); R2(FFMA2, 1, FFMA, 1); R2(FFMA2, 1, IADD3, 1); R2(FFMA2, 1, IMAD, 1); R2(FFMA2, 1, FSET, 1);
    R2(FFMA, 1, FMNMX, 1); R2(FFMA, 1, LOP3, 1); R2(FFMA, 2, LOP3, 1); R2(FMNMX, 1, LOP3, 1); R2(FMNMX, 2, LOP3, 1); R2(FFMA, 1, IMAD, 1);
    R2(FFMA2, 2, MUFU, 1); R2(FFMA2, 4, MUFU, 1); R2(FFMA2, 2, SHFL, 1); R2(FFMA2, 2, LDS128B, 1); R2(FFMA2, 2, LDS128G, 1); R2(FFMA2, 1, LDS64B, 1);
    R2(FFMA2, 2, FSETP_SEL, 1); R2(FFMA2, 3, FSEL, 1); R2(MUFU, 1, LOP3, 2); R2(MUFU, 1, SHFL, 1);
    return 0;
}
