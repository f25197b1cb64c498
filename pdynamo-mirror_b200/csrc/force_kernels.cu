// force_kernels.cu -- the MM/MM energy + gradient kernels (sm_100a).
//
// Replaces PairwiseInteractionABFS_MMMMEnergy (analytic branch, pM/csource/PairwiseInteraction.c:292-429, loop :374-427,
// macros pM/cinclude/PairwiseInteraction.h:72-119) as driven by NBModelABFS_MMMMEnergy (pM/csource/NBModelABFS.c:228-301)
// and MMMMImageEnergy (:1161-1313), including the image gradient rotation (:1295-1298) and the sums needed by
// SymmetryParameterGradients_ImageDerivatives (pM/csource/SymmetryParameterGradients.c:158-238).
//
// k_cluster_forces: one warp per work item (= one i-cluster of 8 sorted atoms x up to `chunkTiles` tiles of 32 j slots of one
// image).  Lane l OWNS j slot l of the current tile (its record, its gradient accumulator, the column byte of its descriptor:
// bit i = the pair (cluster atom i, this j) is on the list) and walks the 8 cluster atoms two at a time: the i atoms are staged
// once per item in shared memory as PAIRS {x0, x1, y0, y1}, {z0, z1, q0, q1}, read back with broadcast 128-bit loads, and the two
// pairs (i0, j), (i1, j) are evaluated together with packed fp32 arithmetic (FFMA2 / FMUL2 / FADD2 of sm_100: one issue slot for
// two operations, scalar operands broadcast for free).  Measured on B200 (scripts/microbench/mix_bench.cu): a packed
// instruction occupies the FMA pipe for two cycles and the half-rate ALU instructions (FSEL, FSETP, FMNMX, LOP3) issue in its shadow,
// whereas a scalar FFMA followed by an ALU instruction costs the sum of both -- the region selects are therefore free next to packed math.
// The j gradient stays in the lane's registers (no shuffles in the pair loop: SHFL costs 4 cycles per scheduler); the i gradients are
// 24 fp32 accumulators per lane, reduced across the warp by a reduce-scatter every kFlushTiles tiles and added to the sorted
// gradient in fp64.  All atom data come from per-call RECORDS in sorted order (two float4 per atom: grid-relative fp32 coordinates
// split as K + xl with K a multiple of 8 A, charge, LJ type), pair math is fp32 in cluster-local coordinates (exact K differences
// plus one rounding), accumulation across tiles / into global memory in fp64.
// Gradients are accumulated in sorted order and scattered to atom order by k_unsort_gradients.
#include "nbb200_internal.h"
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>

namespace nbb200 {

struct ForceArgs {
    const WorkItem *items; int nitems; unsigned int *workCursor;
    const unsigned int *tileDesc; int chunkTiles; int exp;
    const float4 *recA, *recB; int n;
    const float2 *ljAB; int ntypes; const unsigned char *typeFree;
    const ImageOpDev *ops;
    AbfsF32 F; float qScale;
    double *gradSorted; double *accum;
    double origin[3];
    // spline form: per interval l of the shared abscissae three float4 = the cubics {c0, c1, c2, c3} in u = r^2 - xf_l of the electrostatic,
    // LJ-A and LJ-B splines, stored as three arrays of splN entries (tab[k * splN + l]: lanes with different l spread over all
    // shared-memory banks); xf_l = fl((l dR) * (l dR)) in fp32 is recomputed per pair; entry splN - 1 is all zero (pairs not evaluated)
    const float4 *splTab; int splN; float splInvDR, splDR;
};

// ------------------------------------------------------------------------------------------------------
// pair math, fp32, branch free.  The three regions of the reference macros (damped core r < rDamp, plain shifted
// rDamp <= r <= rOn, switched rOn < r <= rOff) are folded into ONE instruction stream with per-lane selected coefficients,
// because the lanes of a warp always mix regions (a divergent branch would execute every path anyway):
//   Coulomb  E = qij (s G + sh)          G = 1, sh = qShift1 (plain) | G = t^3 C(t), sh = 0 (switched; t = rOff - r, C cubic:
//                                         the reference polynomial a/r - b r - c r^3 - d r^5 + qShift2 has a triple zero at rOff
//                                         and cancels ~200:1 in fp32 when evaluated as written)
//            2 dE/d(r^2) = -qij s^3 Q    Q = 1 (plain) | u^2 (k1 - k2 u), u = rOff^2 - r^2 (the switch function, factored)
//   LJ       E = A pa - B pb             pa = ka (s6 - xa)^2 - wa, pb = kb (s3 - xb)^2 - wb
//                                         plain: ka = 1, xa = 0, wa = aShift12, kb = 1, xb = 0, wb = bShift6
//                                         switched: ka = aK12, xa = aF6, wa = 0, kb = bK6, xb = bF3, wb = 0
//            2 dE/d(r^2) = -6 s^2 (2 A ka (s6 - xa) s6 - B kb (s3 - xb) s3)
// The damped core (r < dampingCutoff = 0.5 A) never occurs in a physical system; lanes that hit it are patched by a
// rarely taken slow path that reproduces the reference's damped formulas (including its LJ-B sign quirk).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_fast(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct PairOut { float e1, e2, g; };      // elect, LJ, g = -2 dE/d(r^2)  (force on i = g * (xi - xj), gradient = -that)

// S = A aShift12 - B bShift6 comes with the LJ table entry (A, B, S, 0): the two plain-region energy shifts cost one select
__device__ __forceinline__ PairOut abfs_pair(const AbfsF32 &F, float r2, float qij, float A, float B, float S)
{
    float s = rsqrt_fast(r2);
#ifdef NBB_NEWTON
    {   // optional Newton step.  Measured on B200 (scripts/accuracy_report.py, 7 systems): MUFU.RSQ alone (2^-22 relative) leaves the
        // total energy at 2e-8 .. 2e-7 and the gradients at 6e-7 .. 9e-7 relative RMS -- the same as with the refinement, because the
        // error floor is the fp32 rounding of the cluster-local coordinates.  The step costs 4 of ~80 issue slots per pair: off.
        const float rr = r2 * s;
        s = fmaf(0.5f * s, fmaf(-rr, s, 1.0f), s);
    }
#endif
    const float r = r2 * s, s2 = s * s, s3 = s * s2, s6 = s3 * s3;
    const bool plain = r2 <= F.r2On;
    // Coulomb
    const float t = F.rOff - r;
    const float C = fmaf(fmaf(fmaf(F.n6, t, F.n5), t, F.n4), t, F.n3);
    const float t3C = (t * t) * (t * C);
    const float G = plain ? 1.0f : t3C, sh = plain ? F.qShift1 : 0.0f;
    const float u = F.r2Off - r2;
    const float Qs = (u * u) * fmaf(-F.k2, u, F.k1);
    const float Q = plain ? 1.0f : Qs;
    PairOut o;
    o.e1 = qij * fmaf(s, G, sh);
    const float gq = (qij * s3) * Q;
    // Lennard-Jones
    const float ka = plain ? 1.0f : F.aK12, xa = plain ? 0.0f : F.aF6;
    const float kb = plain ? 1.0f : F.bK6,  xb = plain ? 0.0f : F.bF3;
    const float Sp = plain ? S : 0.0f;
    const float la = s6 - xa, lb = s3 - xb;
    const float X = A * (ka * la), Y = B * (kb * lb);
    o.e2 = fmaf(X, la, fmaf(-Y, lb, -Sp));
    const float m = fmaf(2.0f, X * s6, -(Y * s3));
    o.g = fmaf(6.0f * s2, m, gq);
    return o;
}

// ------------------------------------------------------------------------------------------------------
// packed fp32 (sm_100 FFMA2 / FMUL2 / FADD2): two pairs per instruction.  bc() broadcasts a scalar (a register, uniform-register or
// constant operand modifier in SASS, no instruction); negation is an operand modifier as well.
// ------------------------------------------------------------------------------------------------------
typedef float2 f2;
__device__ __forceinline__ f2 bc(float v) { return make_float2(v, v); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ f2 sel2(bool p0, bool p1, f2 a, f2 b) { return make_float2(p0 ? a.x : b.x, p1 ? a.y : b.y); }

// the two pairs (i0, j), (i1, j) at once: the same formulas as abfs_pair.  eq, el: running energy sums; g = -2 dE/d(r^2) of both pairs.
// lj0, lj1: shared-memory byte addresses of the pairs' LJ entries.  Every entry is TWO float4: {A, B, A aShift12 - B bShift6, 0} for
// the plain region and {A aK12, B bK6, 0, 0} for the switched one -- the region then selects an address (one predicated add) instead of
// three coefficients per pair; the shifts inside the brackets (aF6, bF3) and the plain-region Coulomb shift are blended with
// pi = 1 / 0 (plain / switched), which turns a subtraction into an FFMA2 of the same cost.  What is left for the ALU pipe per pair: the
// region test, pi, the two selects of the Coulomb switch and the list-mask select.
__device__ __forceinline__ float4 lds128(unsigned int addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// kLJ = false: both cluster atoms of the pair are of a type without any Lennard-Jones interaction (TIP3P hydrogens ...): Coulomb only
template <bool kLJ>
__device__ __forceinline__ void abfs_pair2(const AbfsF32 &F, f2 r2, f2 qij, unsigned int lj0, unsigned int lj1, f2 &eq, f2 &el, f2 &g)
{
    const f2 s = make_float2(rsqrt_fast(r2.x), rsqrt_fast(r2.y));
    const bool p0 = r2.x <= F.r2On, p1 = r2.y <= F.r2On;
    float4 ab0, ab1;
    if (kLJ) { ab0 = lds128(p0 ? lj0 : lj0 + 16u); ab1 = lds128(p1 ? lj1 : lj1 + 16u); }
    const f2 r = mul2(r2, s), s2 = mul2(s, s), s3 = mul2(s, s2);
    const f2 one = bc(1.0f);
    const f2 sg = make_float2(p0 ? 0.0f : 1.0f, p1 ? 0.0f : 1.0f);          // sigma = 1 - pi
    // Coulomb
    const f2 t = sub2(bc(F.rOff), r);
    const f2 C = fma2(fma2(fma2(t, bc(F.n6), bc(F.n5)), t, bc(F.n4)), t, bc(F.n3));
    const f2 t3C = mul2(mul2(t, t), mul2(t, C));
    const f2 G = sel2(p0, p1, one, t3C);
    const f2 u = sub2(bc(F.r2Off), r2);
    const f2 Qs = mul2(mul2(u, u), fma2(u, bc(-F.k2), bc(F.k1)));
    const f2 Q = sel2(p0, p1, one, Qs);
    const f2 sh = fma2(sg, bc(-F.qShift1), bc(F.qShift1));               // plain: qShift1, switched: 0
    eq = fma2(qij, fma2(s, G, sh), eq);
    const f2 gq = mul2(mul2(qij, s3), Q);
    if (!kLJ) { g = gq; return; }
    // Lennard-Jones (the table entries of the two pairs sit in unrelated registers: their products are scalar)
    const f2 s6 = mul2(s3, s3);
    const f2 la = fma2(sg, bc(-F.aF6), s6), lb = fma2(sg, bc(-F.bF3), s3);
    const f2 X = make_float2(ab0.x * la.x, ab1.x * la.y), Y = make_float2(ab0.y * lb.x, ab1.y * lb.y);
    el = fma2(X, la, el); el = fma2(neg2(Y), lb, el); el.x -= ab0.z; el.y -= ab1.z;
    const f2 m = fma2(bc(2.0f), mul2(X, s6), neg2(mul2(Y, s3)));
    g = fma2(mul2(bc(6.0f), s2), m, gq);
}

// per-warp staging of the work item's i-cluster, two atoms per entry; the landing zone of the asynchronous record copies (one slot per
// lane: the two records of the lane's j atom of the NEXT tile, in flight during the current tile without holding registers); the
// fp64 sums of the item (touched once per kFlushTiles tiles)
struct __align__(16) IStage {
    float4 xy[kCluster / 2];        // {x0, x1, y0, y1} cluster-local
    float4 zq[kCluster / 2];        // {z0, z1, q0, q1} charges times the electrostatic conversion factor
    int2   row[kCluster / 2];       // byte offsets of the atoms' LJ-table rows
    int2   pad[kCluster / 2];
    float4 recA[kTile], recB[kTile];
    float4 nextA[kCluster], nextB[kCluster];   // records of the NEXT item's cluster atoms (asynchronous copies issued one item ahead)
    double sum[3][kTile];           // per lane: Coulomb energy, LJ energy, component (lane & 3) of the summed i gradients
};

// Slow path for tiles that contain a pair inside the damped core, r^2 < r2Damp (reference: PairwiseInteraction.h:72-119,
// third branches, with s = s2 = 0).  The lane's j gradient and the energies get the CORRECTION (damped minus what abfs_pair2
// produced) through c = {gxj, gyj, gzj, eq, el}; the i side goes straight to the sorted gradient (and, for images, into the G
// sums of the image) in fp64 -- this path is practically never taken.
__device__ __noinline__ void damped_tile_fix(const AbfsF32 &F, unsigned int mask, const IStage *ist, const unsigned char *ljBase, int ljoff, float xj, float yj, float zj,
                                            float qj, double *gradI, double sc, double *accImage, float *c)
{
    float fxj = 0.f, fyj = 0.f, fzj = 0.f, eq = 0.f, el = 0.f;
    const float *xy = reinterpret_cast<const float *>(ist->xy), *zq = reinterpret_cast<const float *>(ist->zq);
    const int *row = reinterpret_cast<const int *>(ist->row);
    for (int i = 0; i < kCluster; i++) {
        const int p = i >> 1, h = i & 1;
        const float dx = xy[4 * p + h] - xj, dy = xy[4 * p + 2 + h] - yj, dz = zq[4 * p + h] - zj;
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        if (((mask >> i) & 1u) && (r2 < F.r2Damp)) {
            const float4 ab = *reinterpret_cast<const float4 *>(ljBase + row[i] + ljoff);      // the plain-region half of the entry
            const float qij = zq[4 * p + 2 + h] * qj;
            const PairOut o = abfs_pair(F, F.r2Damp, qij, ab.x, ab.y, ab.z);      // what the main path evaluated: the pair at the boundary of the core
            const float e1 = qij * fmaf(-F.qAlpha, r2, F.qF0);
            const float e2 = ab.x * fmaf(-F.aAlpha, r2, F.aF0) - ab.y * fmaf(-F.bAlpha, r2, F.bF0);
            const float g = -2.0f * (-qij * F.qAlpha - ab.x * F.aAlpha + ab.y * F.bAlpha) - o.g;
            eq += e1 - o.e1; el += e2 - o.e2;
            const float gx = g * dx, gy = g * dy, gz = g * dz;
            fxj += gx; fyj += gy; fzj += gz;
            if (gradI != nullptr) {
                atomicAdd(gradI + 3 * i, -sc * (double) gx); atomicAdd(gradI + 3 * i + 1, -sc * (double) gy); atomicAdd(gradI + 3 * i + 2, -sc * (double) gz);
            }
            if (accImage != nullptr) {          // G = sum of the image atoms' gradients = minus the i-side sum
                atomicAdd(accImage + 2, sc * (double) gx); atomicAdd(accImage + 3, sc * (double) gy); atomicAdd(accImage + 4, sc * (double) gz);
            }
        }
    }
    c[0] = fxj; c[1] = fyj; c[2] = fzj; c[3] = eq; c[4] = el;
}

// ------------------------------------------------------------------------------------------------------
// spline form (PairwiseInteractionABFS_MMMMEnergy, second branch: pM/csource/PairwiseInteraction.c:431-531).  The reference finds the
// interval of r^2 by bisection over x_i = (i dR)^2 (CubicSpline_EvaluateLUDST) and evaluates the three splines there
// (CubicSpline_FastEvaluateFG); here the interval is l = floor(r / dR) and every interval holds its cubic in u = r^2 - x_l
// (the same polynomial; at a knot either neighbour gives the same value and first two derivatives).  One instruction stream for
// all r: the tables cover the damped core, the plain and the switched region.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqrt_fast(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ PairOut spline_pair(const float4 *__restrict__ tab, int nrows, float invDR, float dR, bool live, float r2, float qij, float A, float B)
{
    int l = min((int) (sqrt_fast(r2) * invDR), nrows - 2);
    l = live ? l : nrows - 1;
    const float4 *row = tab + l;
    const float4 te = row[0], ta = row[nrows], tb = row[2 * nrows];
    const float rl = __fmul_rn((float) l, dR);
    const float u = r2 - __fmul_rn(rl, rl);              // the expansion point of the stored cubics (upload_spline_tables), not contracted
    // Lennard-Jones: coefficients are linear in (A, B)
    const float c0 = fmaf(A, ta.x, B * tb.x), c1 = fmaf(A, ta.y, B * tb.y), c2 = fmaf(A, ta.z, B * tb.z), c3 = fmaf(A, ta.w, B * tb.w);
    PairOut o;
    o.e1 = qij * fmaf(u, fmaf(u, fmaf(u, te.w, te.z), te.y), te.x);
    o.e2 = fmaf(u, fmaf(u, fmaf(u, c3, c2), c1), c0);
    const float u3 = 3.0f * u;
    const float ge = fmaf(u, fmaf(u3, te.w, te.z + te.z), te.y);
    const float gl = fmaf(u, fmaf(u3, c3, c2 + c2), c1);
    o.g = -2.0f * fmaf(qij, ge, gl);                     // dG = 2 (qij dFe + A dFa + B dFb); g = -dG
    return o;
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ------------------------------------------------------------------------------------------------------
// bulk asynchronous copy (TMA engine, UBLKCP in SASS) of a work item's descriptor stream into shared memory: the tiles of an item
// are contiguous in the pool (<= chunkTiles x 128 B), so ONE lane issues ONE cp.async.bulk per item, one item ahead, and the
// completion is signalled through an mbarrier (complete_tx); the lanes then read their descriptor words with LDS.  With plain
// LDG prefetches ptxas put the descriptor and the record loads on one scoreboard and the record gather of every tile waited for
// the descriptor load issued just before it (26 % of all stall samples, profiles/README.md).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_addr(const void *p) { return (unsigned int) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned int mbar, unsigned int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned int mbar, unsigned int bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(unsigned int dst, const void *src, unsigned int bytes, unsigned int mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned int dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned int mbar, unsigned int parity)
{
    asm volatile("{\n.reg .pred p;\nNBB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra NBB_DONE;\nbra NBB_WAIT;\nNBB_DONE:\n}" :: "r"(mbar), "r"(parity) : "memory");
}

constexpr int kLJEntryBytes = 2 * (int) sizeof(float4);     // plain-region and switched-region coefficients of a type pair
constexpr int kFuseMaxAtoms = 262144;
constexpr int kFlushTiles = 8;                       // the fp32 i-gradient / energy accumulators are flushed to fp64 every 8 tiles

// sum over the warp of 24 values per lane (8 cluster atoms x 3 components, v[3 a + c]), scattered: afterwards v[0..2] of every lane
// hold the complete x, y, z sums of atom (lane >> 2) & 7 (the four lanes of a quad hold the same numbers)
__device__ __forceinline__ void reduce_scatter_24(float (&v)[24], int lane)
{
    const bool up1 = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const float keep = up1 ? v[k + 12] : v[k], send = up1 ? v[k] : v[k + 12];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const bool up2 = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const float keep = up2 ? v[k + 6] : v[k], send = up2 ? v[k] : v[k + 6];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const bool up3 = (lane & 4) != 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float keep = up3 ? v[k + 3] : v[k], send = up3 ? v[k] : v[k + 3];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { v[k] += __shfl_xor_sync(0xffffffffu, v[k], 2); v[k] += __shfl_xor_sync(0xffffffffu, v[k], 1); }
}

// kForm: 0 = analytic ABFS formulas, 1 = spline tables staged in shared memory, 2 = spline tables read from global memory (L1)
template <bool kRot, int kThreads, int kMinBlocks, int kForm>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_cluster_forces(const __grid_constant__ ForceArgs A)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    // per warp: two descriptor buffers of chunkTiles tiles (128 B each), the staged i-cluster, two mbarriers; then the tables of the CTA
    const size_t warpBytes = 2 * (size_t) A.chunkTiles * kTile * sizeof(unsigned int) + sizeof(IStage) + 32;
    unsigned char *warpBase = smemRaw + warpBytes * (threadIdx.x >> 5);
    const unsigned int *sDesc = reinterpret_cast<const unsigned int *>(warpBase);
    IStage *ist = reinterpret_cast<IStage *>(warpBase + 2 * (size_t) A.chunkTiles * kTile * sizeof(unsigned int));
    const unsigned int mbar0 = smem_addr(reinterpret_cast<unsigned char *>(ist) + sizeof(IStage));
    // LJ table with one extra, all-zero row and column (index ntypes): empty j slots and the padding rows of the last cluster carry that
    // type and a zero charge, so that their pairs contribute exactly nothing wherever their coordinates lie.  An entry is two float4:
    // {A, B, A aShift12 - B bShift6, 0} for the plain region, {A aK12, B bK6, 0, 0} for the switched region (see abfs_pair2)
    const int nt1 = A.ntypes + 1;
    float4 *sLJ = reinterpret_cast<float4 *>(smemRaw + warpBytes * (kThreads / 32));  // [nt1 * nt1][2]
    for (int i = threadIdx.x; i < nt1 * nt1; i += blockDim.x) {
        const int ti = i / nt1, tj = i - ti * nt1;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f), e2 = e;
        if (ti < A.ntypes && tj < A.ntypes) {
            const float2 ab = A.ljAB[ti * A.ntypes + tj];
            e = make_float4(ab.x, ab.y, ab.x * A.F.aShift12 - ab.y * A.F.bShift6, 0.f);
            e2 = make_float4(ab.x * A.F.aK12, ab.y * A.F.bK6, 0.f, 0.f);
        }
        sLJ[2 * i] = e; sLJ[2 * i + 1] = e2;
    }
    const float4 *splTab = A.splTab;
    if (kForm == 1) {
        float4 *sTab = sLJ + 2 * nt1 * nt1;
        for (int i = threadIdx.x; i < 3 * A.splN; i += blockDim.x) sTab[i] = A.splTab[i];
        splTab = sTab;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const AbfsF32 &F = A.F;
    const unsigned char *ljBase = reinterpret_cast<const unsigned char *>(sLJ);
    const unsigned int ljBaseS = smem_addr(sLJ);
    const int comp = lane & 3, atomOfLane = (lane >> 2) & (kCluster - 1);     // after a reduce-scatter: this lane adds component `comp` of that atom

    if ((threadIdx.x & 31) == 0) { mbar_init(mbar0, 1); mbar_init(mbar0 + 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // work items are claimed one ahead: the cursor atomic, the item record and the descriptor stream of the NEXT item are in flight during this one
    unsigned int itNext = 0;
    if (lane == 0) itNext = atomicAdd(A.workCursor, 1u);
    itNext = __shfl_sync(0xffffffffu, itNext, 0);
    WorkItem wiNext = A.items[min(itNext, (unsigned int) (A.nitems - 1))];
    unsigned int seq = 0;                                   // items processed by this warp: buffer = seq & 1, mbarrier parity = (seq >> 1) & 1
    if (lane == 0 && itNext < (unsigned int) A.nitems) {
        const unsigned int bytes = (unsigned int) wiNext.tileCount * kTile * sizeof(unsigned int);
        mbar_expect_tx(mbar0, bytes);
        bulk_copy_g2s(smem_addr(sDesc), A.tileDesc + (size_t) wiNext.tileStart * kTile, bytes, mbar0);
    }
    // the records of an item's cluster atoms are copied one item ahead as well (padding rows of the last cluster: the null record, index n)
    auto fetch_i = [&](const WorkItem &w) {
        if (lane < kCluster) {
            const int si = min(w.block * kCluster + lane, A.n);
            cp_async16(smem_addr(&ist->nextA[lane]), A.recA + si); cp_async16(smem_addr(&ist->nextB[lane]), A.recB + si);
        }
    };
    if (itNext < (unsigned int) A.nitems) fetch_i(wiNext);
    for (;; seq++) {
        if (itNext >= (unsigned int) A.nitems) break;
        const WorkItem wi = wiNext;
        if (lane == 0) itNext = atomicAdd(A.workCursor, 1u);
        itNext = __shfl_sync(0xffffffffu, itNext, 0);
        wiNext = A.items[min(itNext, (unsigned int) (A.nitems - 1))];
        const unsigned int *myDesc = sDesc + (size_t) (seq & 1u) * A.chunkTiles * kTile + lane;
        const ImageOpDev *op = A.ops + wi.image;
        const bool isImage = wi.image > 0;
        const bool pureT = kRot ? (op->pureTranslation != 0) : true;
        const double sc = op->scale;

        // cluster-local frame: K of the first cluster atom (a multiple of 8 A, exact in fp32)
        const int s0 = wi.block * kCluster;
        cp_async_wait_all();
        __syncwarp();                                       // the previous item's last tile has been consumed; this item's cluster records have landed
        const float4 kref = ist->nextB[0];
        bool ljFree = false;
        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
        if (lane < kCluster) { ra = ist->nextA[lane]; rb = ist->nextB[lane]; }
        __syncwarp();                                       // ... and have been read: the slots are free for the next item's
        if (lane < kCluster) {
            // the small part tl of a pure translation (|tl| <= 4 A) is taken off the i atoms instead of being added to every j atom
            float tlx = 0.f, tly = 0.f, tlz = 0.f;
            if (isImage && pureT) { tlx = op->tl[0]; tly = op->tl[1]; tlz = op->tl[2]; }
            const float xi = (rb.x - kref.x) + (ra.x - tlx), yi = (rb.y - kref.y) + (ra.y - tly), zi = (rb.z - kref.z) + (ra.z - tlz);
            float *xy = reinterpret_cast<float *>(ist->xy), *zq = reinterpret_cast<float *>(ist->zq);
            const int p = lane >> 1, h = lane & 1;
            xy[4 * p + h] = xi; xy[4 * p + 2 + h] = yi; zq[4 * p + h] = zi; zq[4 * p + 2 + h] = ra.w * A.qScale;
            reinterpret_cast<int *>(ist->row)[lane] = __float_as_int(rb.w) * nt1 * kLJEntryBytes;
            ljFree = A.typeFree[__float_as_int(rb.w)] != 0;
        }
        // pairs of cluster atoms that are both of a type without Lennard-Jones interaction (the builder moves such atoms to the front of
        // every cluster): Coulomb only
        const unsigned int freeBits = __ballot_sync(0xffffffffu, ljFree);
        // pattern of Coulomb-only atom PAIRS: 0..4 = the first nf pairs and no other, 5 = anything else (decided step by step)
        const unsigned int pairFree = (freeBits & (freeBits >> 1)) & 0x55u;
        const int variant = (A.exp & 8) ? 5 : pairFree == 0u ? 0 : pairFree == 0x01u ? 1 : pairFree == 0x05u ? 2 : pairFree == 0x15u ? 3 : pairFree == 0x55u ? 4 : 5;
        __syncwarp();
        // j offsets: X'_local = (K_j + ok) + xl_j, ok exact (kt: the part of a pure translation that is a multiple of 8 A)
        float okx = -kref.x, oky = -kref.y, okz = -kref.z;
        if (isImage && pureT) { okx += op->kt[0]; oky += op->kt[1]; okz += op->kt[2]; }
        double W[9];
        if (kRot) {
#pragma unroll
            for (int k = 0; k < 9; k++) W[k] = 0.0;
        }
        double *gradI = (A.gradSorted != nullptr) ? A.gradSorted + 3 * (size_t) s0 : nullptr;
        const bool flushLane = comp < 3 && s0 + atomOfLane < A.n;

        // the item's descriptors have landed in shared memory (bulk copy issued one item ago) ...
        mbar_wait(mbar0 + 8 * (seq & 1u), (seq >> 1) & 1u);
        if (itNext < (unsigned int) A.nitems) fetch_i(wiNext);
        if (lane == 0 && itNext < (unsigned int) A.nitems) {
            // ... and the next item's stream goes to the other buffer: its last reader (the item before this one) finished before the
            // __syncwarp at the top of this item
            const unsigned int nb = (seq + 1u) & 1u, bytes = (unsigned int) wiNext.tileCount * kTile * sizeof(unsigned int);
            mbar_expect_tx(mbar0 + 8 * nb, bytes);
            bulk_copy_g2s(smem_addr(sDesc + (size_t) nb * A.chunkTiles * kTile), A.tileDesc + (size_t) wiNext.tileStart * kTile, bytes, mbar0 + 8 * nb);
        }
        // the j atom of the lane for the current tile, in cluster-local coordinates; empty slots point at the null record (index n): far
        // away, zero charge, null LJ type -- their pairs contribute exactly nothing
        float xj, yj, zj, qj;
        unsigned int ljS, sj, mask;
        double xj64 = 0.0, yj64 = 0.0, zj64 = 0.0;          // primary coordinates of j (rotations only)
        unsigned int dN = 0u;
        const unsigned int myRecA = smem_addr(&ist->recA[lane]), myRecB = smem_addr(&ist->recB[lane]);
        auto fetch_j = [&](int t) {                          // records of tile t (beyond the item: the null record): asynchronous copies, no registers in flight
            dN = (t < wi.tileCount) ? myDesc[t * kTile] : kEmptySlot;
            const unsigned int s = min(dN & kEmptySlot, (unsigned int) A.n);
            cp_async16(myRecA, A.recA + s); cp_async16(myRecB, A.recB + s);
        };
        auto unpack_j = [&]() {
            cp_async_wait_all();
            const float4 ra = lds128(myRecA), rb = lds128(myRecB);
            sj = min(dN & kEmptySlot, (unsigned int) A.n);
            mask = dN >> 24;                                 // empty slots carry a zero byte
            if (kRot && isImage && !pureT) {
                const double X = (double) rb.x + (double) ra.x, Y = (double) rb.y + (double) ra.y, Z = (double) rb.z + (double) ra.z;
                xj64 = X + A.origin[0]; yj64 = Y + A.origin[1]; zj64 = Z + A.origin[2];
                const double px = op->R[0] * X + op->R[1] * Y + op->R[2] * Z + op->cr[0];
                const double py = op->R[3] * X + op->R[4] * Y + op->R[5] * Z + op->cr[1];
                const double pz = op->R[6] * X + op->R[7] * Y + op->R[8] * Z + op->cr[2];
                xj = (float) (px - (double) kref.x); yj = (float) (py - (double) kref.y); zj = (float) (pz - (double) kref.z);
            } else { xj = (rb.x + okx) + ra.x; yj = (rb.y + oky) + ra.y; zj = (rb.z + okz) + ra.z; }
            qj = ra.w;
            ljS = ljBaseS + (unsigned int) (__float_as_int(rb.w) * kLJEntryBytes);
        };
        fetch_j(0);
        unpack_j();
        fetch_j(1);
        f2 fi[kCluster / 2][3];                             // gradient on the cluster atoms, pair p = atoms (2p, 2p+1)
#pragma unroll
        for (int p = 0; p < kCluster / 2; p++) { fi[p][0] = bc(0.f); fi[p][1] = bc(0.f); fi[p][2] = bc(0.f); }
        f2 eq = bc(0.f), el = bc(0.f);
        ist->sum[0][lane] = 0.0; ist->sum[1][lane] = 0.0; ist->sum[2][lane] = 0.0;      // per lane: energies of the item, component `comp` of the summed i gradients
        for (int t = 0; t < wi.tileCount; t++) {
            f2 fj[3] = {bc(0.f), bc(0.f), bc(0.f)};
            float r2min = F.r2Off;
            // one double step: cluster atoms 2p, 2p + 1 against the lane's j atom.  LJ: 1 = with the Lennard-Jones part, 0 = Coulomb only, -1 = decided
            // at run time from the cluster's types (a warp-uniform branch)
            auto step = [&](auto P, auto LJ) {
                constexpr int p = decltype(P)::value;
                constexpr int lj = decltype(LJ)::value;
                const float4 XY = ist->xy[p], ZQ = ist->zq[p];
                const f2 dx = sub2(make_float2(XY.x, XY.y), bc(xj)), dy = sub2(make_float2(XY.z, XY.w), bc(yj)), dz = sub2(make_float2(ZQ.x, ZQ.y), bc(zj));
                f2 r2 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
                const f2 qij = mul2(make_float2(ZQ.z, ZQ.w), bc(qj));
                const bool on0 = (mask >> (2 * p)) & 1u, on1 = (mask >> (2 * p + 1)) & 1u;
                f2 g;
                if (kForm == 0) {
                    // pairs that are not on the list or beyond the outer cutoff are evaluated AT the outer cutoff, where energy and force vanish;
                    // pairs inside the damped core are evaluated at its boundary and corrected by the slow path below
                    r2.x = fminf(on0 ? r2.x : F.r2Off, F.r2Off); r2.y = fminf(on1 ? r2.y : F.r2Off, F.r2Off);
                    r2min = fminf(r2min, fminf(r2.x, r2.y));
                    r2.x = fmaxf(r2.x, F.r2Damp); r2.y = fmaxf(r2.y, F.r2Damp);
                    if (lj == 0 || (lj < 0 && ((freeBits >> (2 * p)) & 3u) == 3u)) abfs_pair2<false>(F, r2, qij, 0u, 0u, eq, el, g);
                    else {
                        const int2 rows = ist->row[p];
                        abfs_pair2<true>(F, r2, qij, ljS + (unsigned int) rows.x, ljS + (unsigned int) rows.y, eq, el, g);
                    }
                } else {
                    // pairs off the list or beyond the cutoff read the all-zero table row (the skip of PairwiseInteraction.c:489)
                    const int2 rows = ist->row[p];
                    const float4 ab0 = lds128(ljS + (unsigned int) rows.x), ab1 = lds128(ljS + (unsigned int) rows.y);
                    const PairOut o0 = spline_pair(splTab, A.splN, A.splInvDR, A.splDR, on0 && !(r2.x > F.r2Off), r2.x, qij.x, ab0.x, ab0.y);
                    const PairOut o1 = spline_pair(splTab, A.splN, A.splInvDR, A.splDR, on1 && !(r2.y > F.r2Off), r2.y, qij.y, ab1.x, ab1.y);
                    eq = add2(eq, make_float2(o0.e1, o1.e1)); el = add2(el, make_float2(o0.e2, o1.e2));
                    g = make_float2(o0.g, o1.g);
                }
                const f2 ng = neg2(g);
                fi[p][0] = fma2(ng, dx, fi[p][0]); fi[p][1] = fma2(ng, dy, fi[p][1]); fi[p][2] = fma2(ng, dz, fi[p][2]);      // gradient = -(force on i) = -g d
                fj[0] = fma2(g, dx, fj[0]); fj[1] = fma2(g, dy, fj[1]); fj[2] = fma2(g, dz, fj[2]);
            };
            // the four double steps of the tile as ONE basic block per pattern of Coulomb-only atom pairs (the builder moves the atoms without
            // Lennard-Jones interaction to the front of the cluster: the first nf pairs): without a branch between the steps ptxas overlaps
            // the reciprocal square root and the dependent chains of neighbouring steps
            auto steps = [&](auto NF) {
                constexpr int nf = decltype(NF)::value;
                using std::integral_constant;
                step(integral_constant<int, 0>{}, integral_constant<int, (nf < 0 ? -1 : (0 < nf ? 0 : 1))>{});
                step(integral_constant<int, 1>{}, integral_constant<int, (nf < 0 ? -1 : (1 < nf ? 0 : 1))>{});
                step(integral_constant<int, 2>{}, integral_constant<int, (nf < 0 ? -1 : (2 < nf ? 0 : 1))>{});
                step(integral_constant<int, 3>{}, integral_constant<int, (nf < 0 ? -1 : (3 < nf ? 0 : 1))>{});
            };
            if (kForm == 0) {
                switch (variant) {
                    case 0: steps(std::integral_constant<int, 0>{}); break;
                    case 1: steps(std::integral_constant<int, 1>{}); break;
                    case 2: steps(std::integral_constant<int, 2>{}); break;
                    case 3: steps(std::integral_constant<int, 3>{}); break;
                    case 4: steps(std::integral_constant<int, 4>{}); break;
                    default: steps(std::integral_constant<int, -1>{}); break;
                }
            } else steps(std::integral_constant<int, -2>{});
            float fxj = fj[0].x + fj[0].y, fyj = fj[1].x + fj[1].y, fzj = fj[2].x + fj[2].y;
            if (kForm == 0 && __any_sync(0xffffffffu, r2min < F.r2Damp)) {   // damped core: practically never; patch the tile with the reference formulas
                float c[5];
                damped_tile_fix(F, mask, ist, ljBase, (int) (ljS - ljBaseS), xj, yj, zj, qj, gradI, sc, isImage ? A.accum + 16 * wi.image : nullptr, c);
                fxj += c[0]; fyj += c[1]; fzj += c[2];
                eq.x += c[3]; el.x += c[4];
            }
            if (sj < (unsigned int) A.n) {
                double gx = sc * (double) fxj, gy = sc * (double) fyj, gz = sc * (double) fzj;       // gradient on the (image) atom
                if (kRot && isImage && !pureT) {
                    W[0] += gx * xj64; W[1] += gx * yj64; W[2] += gx * zj64;
                    W[3] += gy * xj64; W[4] += gy * yj64; W[5] += gy * zj64;
                    W[6] += gz * xj64; W[7] += gz * yj64; W[8] += gz * zj64;
                    const double rx = op->R[0] * gx + op->R[3] * gy + op->R[6] * gz;           // R^T g'
                    const double ry = op->R[1] * gx + op->R[4] * gy + op->R[7] * gz;
                    const double rz = op->R[2] * gx + op->R[5] * gy + op->R[8] * gz;
                    gx = rx; gy = ry; gz = rz;
                }
                if (A.gradSorted != nullptr) {
                    double *gp = A.gradSorted + 3 * (size_t) sj;
                    atomicAdd(gp, gx); atomicAdd(gp + 1, gy); atomicAdd(gp + 2, gz);
                }
            }
            // the next tile's j atom from the records that have been in flight during this tile, then the records of the tile after it
            unpack_j();
            fetch_j(t + 2);
            if (((t + 1) % kFlushTiles) == 0 || t + 1 == wi.tileCount) {
                ist->sum[0][lane] += (double) (eq.x + eq.y); ist->sum[1][lane] += (double) (el.x + el.y);
                eq = bc(0.f); el = bc(0.f);
                float v[3 * kCluster];
#pragma unroll
                for (int p = 0; p < kCluster / 2; p++)
#pragma unroll
                    for (int c = 0; c < 3; c++) { v[6 * p + c] = fi[p][c].x; v[6 * p + 3 + c] = fi[p][c].y; fi[p][c] = bc(0.f); }
                reduce_scatter_24(v, lane);
                const double mine = sc * (double) (comp == 0 ? v[0] : (comp == 1 ? v[1] : v[2]));
                if (flushLane) {
                    ist->sum[2][lane] += mine;
                    if (gradI != nullptr) atomicAdd(gradI + 3 * atomOfLane + comp, mine);
                }
            }
        }
        double *acc = A.accum + 16 * wi.image;
        const double eQ = warp_sum(ist->sum[0][lane]) * sc, eL = warp_sum(ist->sum[1][lane]) * sc;
        if (lane == 0) { atomicAdd(&acc[0], eQ); atomicAdd(&acc[1], eL); }
        if (isImage) {
            // sum over the image atoms of their gradient = minus the sum of the i-side gradients of this item (Newton's third law);
            // lanes with the same `comp` hold the partial sums of that component
            double gsum = ist->sum[2][lane];
            for (int off = 4; off < 32; off <<= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, off);
            if (lane < 3) atomicAdd(&acc[2 + lane], -gsum);
            if (kRot && !pureT) {
#pragma unroll
                for (int k = 0; k < 9; k++) { const double w = warp_sum(W[k]); if (lane == 0) atomicAdd(&acc[5 + k], w); }
            }
        }
    }
}

// pub: results of the call that are complete when this kernel starts (the accumulators of the force kernels, the displacement maximum of an
// optimistic update decision) are written straight into page-locked host memory by the first CTA -- the call then needs no copy operation of its
// own in the stream (a small device-to-host copy between two kernels costs several microseconds of engine switching)
struct PublishArgs { const double *src[2]; double *dst[2]; int count[2]; };

// ------------------------------------------------------------------------------------------------------
// per energy call: atom records in sorted order from the current coordinates, and the way back for the gradients
// ------------------------------------------------------------------------------------------------------
__global__ void k_pack_records(const double *__restrict__ x, const int *__restrict__ sAtom, int n, int ntypes, const float *__restrict__ q32, const int *__restrict__ ljtype,
                               double ox, double oy, double oz, float4 *__restrict__ recA, float4 *__restrict__ recB, double *__restrict__ zero, int zeroCount,
                               const double *__restrict__ xprune, unsigned long long *__restrict__ pruneDisp, int slot, const PublishArgs pub)
{
    if (blockIdx.x == 0) {       // nbb200_md_run: the displacement maximum of this step's first half and the kinetic energy of the previous step
#pragma unroll
        for (int k = 0; k < 2; k++)
            for (int i = threadIdx.x; i < pub.count[k]; i += blockDim.x) pub.dst[k][i] = pub.src[k][i];
    }
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    // fused mode (nbb200_md_run): the accumulators + work cursor of the force kernels are cleared here instead of by a memset of their own
    if (zero != nullptr) for (int i = s; i < zeroCount; i += gridDim.x * blockDim.x) zero[i] = 0.0;
    if (s == n) {       // the null record: what empty tile slots and the padding rows of the last cluster read (far away, no charge, null LJ type)
        recA[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        recB[n] = make_float4(-1.0e15f, -1.0e15f, -1.0e15f, __int_as_float(ntypes));
    }
    double moved = 0.0;
    const int a = (s < n) ? sAtom[s] : -1;                   // -1: a position of a cell this rank does not see (restricted sort, several ranks)
    if (a >= 0) {
        const double xa = x[3 * a], ya = x[3 * a + 1], za = x[3 * a + 2];
        const double X = xa - ox, Y = ya - oy, Z = za - oz;
        const double KX = 8.0 * rint(X * 0.125), KY = 8.0 * rint(Y * 0.125), KZ = 8.0 * rint(Z * 0.125);
        recA[s] = make_float4((float) (X - KX), (float) (Y - KY), (float) (Z - KZ), q32[a]);
        recB[s] = make_float4((float) KX, (float) KY, (float) KZ, __int_as_float(ljtype[a]));
        if (xprune != nullptr) {
            const double dx = xa - xprune[3 * a], dy = ya - xprune[3 * a + 1], dz = za - xprune[3 * a + 2];
            moved = dx * dx + dy * dy + dz * dz;
        }
    }
    if (pruneDisp != nullptr) {
        // rolling prune: how far has any atom moved since the inner lists were made?  Two slots by call parity: this call's maximum goes to
        // `slot` (k_prune of this call reads it after this kernel), the other one is cleared for the next call
        if (xprune != nullptr) {
            for (int off = 16; off > 0; off >>= 1) moved = fmax(moved, __shfl_xor_sync(0xffffffffu, moved, off));
            if ((threadIdx.x & 31) == 0 && moved > 0.0) atomicMax(&pruneDisp[slot], (unsigned long long) __double_as_longlong(moved));
        }
        if (s == 0) pruneDisp[slot ^ 1] = 0ULL;
    }
}

// ------------------------------------------------------------------------------------------------------
// rolling prune.  The lists hold every pair within listCutoff (13.5 A) at the last rebuild; the interaction vanishes beyond outerCutoff
// (12 A), so at any moment about a quarter of the j entries of a cluster hold no pair that contributes.  k_prune copies every work
// item's descriptor stream into the inner pool without the entries whose 8 pairs are all beyond outerCutoff + buffer (at the same tile
// offsets: the inner stream of an item is never longer than the outer one -- no allocation, no atomics) and clears the mask bits of the
// pairs beyond that radius.  The inner pool stays valid until an atom has moved by more than buffer / 2; the decision is taken on the
// device (`force`: the host knows that the lists or the lattice changed), so that a call costs one almost empty launch when nothing is due.
// A pair that is dropped here is beyond outerCutoff whenever the inner pool is used, i.e. it contributes exactly nothing in the
// reference as well (PairwiseInteraction.c:389): the sums are unchanged.
// ------------------------------------------------------------------------------------------------------
struct PruneArgs {
    const WorkItem *items; int nitems; const unsigned int *tileDesc;
    WorkItem *itemsIn; unsigned int *tileDescIn;
    const float4 *recA, *recB; int n;
    const ImageOpDev *ops;
    float rc2;                                   // (outerCutoff + buffer)^2 with a rounding margin
    double thr2;                                 // (buffer / 2)^2
    int force, slot;
    unsigned long long *pruneDisp;
    const double *x; double *xprune;
};

constexpr int kPruneWarps = 8;

template <bool kRot>
__global__ void __launch_bounds__(kPruneWarps * 32) k_prune(const __grid_constant__ PruneArgs P)
{
    if (!P.force && !(__longlong_as_double((long long) P.pruneDisp[P.slot]) > P.thr2)) return;
    __shared__ float4 sI[kPruneWarps][kCluster];          // per warp: the i-cluster as pairs {x0, x1, y0, y1} [4], {z0, z1, -, -} [4]
    __shared__ unsigned int sQ[kPruneWarps][2 * kTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long gw = (long) blockIdx.x * kPruneWarps + warp, nw = (long) gridDim.x * kPruneWarps;
    // the coordinates this prune refers to
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < 3L * P.n; i += (long) gridDim.x * blockDim.x) P.xprune[i] = P.x[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.pruneDisp[2], 1ULL);
    float4 *myI = sI[warp];
    unsigned int *queue = sQ[warp];
    const unsigned int ltMask = (1u << lane) - 1u;
    for (long it = gw; it < P.nitems; it += nw) {
        const WorkItem wi = P.items[it];
        const ImageOpDev *op = P.ops + wi.image;
        const bool isImage = wi.image > 0;
        const bool pureT = kRot ? (op->pureTranslation != 0) : true;
        const int s0 = wi.block * kCluster;
        const float4 kref = P.recB[s0];
        __syncwarp();
        if (lane < kCluster) {
            const int si = min(s0 + lane, P.n);
            const float4 ra = P.recA[si], rb = P.recB[si];
            float tlx = 0.f, tly = 0.f, tlz = 0.f;
            if (isImage && pureT) { tlx = op->tl[0]; tly = op->tl[1]; tlz = op->tl[2]; }
            float *xy = reinterpret_cast<float *>(myI), *zq = reinterpret_cast<float *>(myI + kCluster / 2);
            const int p = lane >> 1, h = lane & 1;
            xy[4 * p + h] = (rb.x - kref.x) + (ra.x - tlx); xy[4 * p + 2 + h] = (rb.y - kref.y) + (ra.y - tly); zq[4 * p + h] = (rb.z - kref.z) + (ra.z - tlz);
        }
        __syncwarp();
        float okx = -kref.x, oky = -kref.y, okz = -kref.z;
        if (isImage && pureT) { okx += op->kt[0]; oky += op->kt[1]; okz += op->kt[2]; }
        int cnt = 0, outTiles = 0;
        unsigned int *out = P.tileDescIn + (size_t) wi.tileStart * kTile + lane;
        const unsigned int *in = P.tileDesc + (size_t) wi.tileStart * kTile + lane;
        // two-deep software pipeline: descriptor of tile t+2 and records of tile t+1 in flight during tile t
        unsigned int d0 = in[0], d1 = (wi.tileCount > 1) ? in[kTile] : kEmptySlot;
        float4 ra0 = P.recA[min(d0 & kEmptySlot, (unsigned int) P.n)], rb0 = P.recB[min(d0 & kEmptySlot, (unsigned int) P.n)];
        __syncwarp();
        for (int t = 0; t < wi.tileCount; t++) {
            const unsigned int d2 = (t + 2 < wi.tileCount) ? in[(size_t) (t + 2) * kTile] : kEmptySlot;
            const unsigned int s1 = min(d1 & kEmptySlot, (unsigned int) P.n);
            const float4 ra1 = P.recA[s1], rb1 = P.recB[s1];
            const unsigned int d = d0;
            const float4 ra = ra0, rb = rb0;
            d0 = d1; d1 = d2; ra0 = ra1; rb0 = rb1;
            float xj, yj, zj;
            if (kRot && isImage && !pureT) {
                const double X = (double) rb.x + (double) ra.x, Y = (double) rb.y + (double) ra.y, Z = (double) rb.z + (double) ra.z;
                xj = (float) (op->R[0] * X + op->R[1] * Y + op->R[2] * Z + op->cr[0] - (double) kref.x);
                yj = (float) (op->R[3] * X + op->R[4] * Y + op->R[5] * Z + op->cr[1] - (double) kref.y);
                zj = (float) (op->R[6] * X + op->R[7] * Y + op->R[8] * Z + op->cr[2] - (double) kref.z);
            } else { xj = (rb.x + okx) + ra.x; yj = (rb.y + oky) + ra.y; zj = (rb.z + okz) + ra.z; }
            // sign bits of r^2 - rc^2, shifted in pair by pair: bit 7 - i of `sgn` <-> cluster atom i
            unsigned int sgn = 0u;
#pragma unroll
            for (int p = 0; p < kCluster / 2; p++) {
                const float4 XY = myI[p], ZQ = myI[kCluster / 2 + p];
                const f2 dx = sub2(make_float2(XY.x, XY.y), bc(xj)), dy = sub2(make_float2(XY.z, XY.w), bc(yj)), dz = sub2(make_float2(ZQ.x, ZQ.y), bc(zj));
                const f2 e = sub2(fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))), bc(P.rc2));
                sgn = __funnelshift_l(__float_as_uint(e.x), sgn, 1);
                sgn = __funnelshift_l(__float_as_uint(e.y), sgn, 1);
            }
            const unsigned int keepMask = (d >> 24) & (__brev(sgn) >> 24);
            const unsigned int bal = __ballot_sync(0xffffffffu, keepMask != 0u);
            if (keepMask != 0u) queue[cnt + __popc(bal & ltMask)] = (d & kEmptySlot) | (keepMask << 24);
            cnt += __popc(bal);
            __syncwarp();
            if (cnt >= kTile) {
                out[(size_t) outTiles * kTile] = queue[lane];
                const unsigned int rest = queue[kTile + lane];
                __syncwarp();
                queue[lane] = rest;
                cnt -= kTile; outTiles += 1;
                __syncwarp();
            }
        }
        // the remainder; an item that lost every entry keeps one empty tile (the force kernel's bulk copies are never empty)
        if (cnt > 0 || outTiles == 0) { out[(size_t) outTiles * kTile] = (lane < cnt) ? queue[lane] : kEmptySlot; outTiles += 1; }
        if (lane == 0) { WorkItem w = wi; w.tileCount = outTiles; P.itemsIn[it] = w; }
    }
}

// assign != 0: the NB term SETS the caller's gradient (every atom has exactly one sorted position) instead of accumulating into it
// kClear (fused mode of nbb200_md_run): the sorted accumulator is left zeroed for the next call (no memset of its own).  Two instantiations:
// the plain one keeps gs const / read-only (the combined one measured 9 x slower on the 1.1 M-atom box: 143 vs 16 us)
template <bool kClear>
__global__ void k_unsort_gradients(typename std::conditional<kClear, double, const double>::type *__restrict__ gs, const int *__restrict__ sAtom, int s0, int n,
                                   double *__restrict__ grad, int assign, const double *__restrict__ cond, double condThr2, const PublishArgs pub)
{
    if (blockIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 2; k++)
            for (int i = threadIdx.x; i < pub.count[k]; i += blockDim.x) pub.dst[k][i] = pub.src[k][i];
    }
    const int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    // optimistic update decision: the lists turned out to be stale (an atom moved beyond the buffer) -- this evaluation is discarded
    const bool discard = cond != nullptr && *cond > condThr2;
    const int a = sAtom[s];
    if (a < 0) return;                                       // restricted sort (several ranks): not a position this rank sees
    const double gx = gs[3 * s], gy = gs[3 * s + 1], gz = gs[3 * s + 2];
    if (!discard) {
        if (assign) { grad[3 * a] = gx; grad[3 * a + 1] = gy; grad[3 * a + 2] = gz; }
        else { grad[3 * a] += gx; grad[3 * a + 1] += gy; grad[3 * a + 2] += gz; }
    }
    if constexpr (kClear) { gs[3 * s] = 0.0; gs[3 * s + 1] = 0.0; gs[3 * s + 2] = 0.0; }
}

// nbb200_md_run: unsort pass and second half of the velocity-Verlet / Langevin step in one launch -- the gradient of sorted position s goes
// to atom sAtom[s] (set, not added: the NB term comes first, the bonded terms have been added in sorted order), a = -100 g / m, v += dt/2 a,
// kinetic energy 0.5 * 0.01 * sum m v^2; the sorted accumulator is left zeroed; the first CTA stores the accumulators of the energy call and
// the bonded energies of the step into page-locked host memory
__global__ void k_unsort_second_half(double *__restrict__ gs, const int *__restrict__ sAtom, int n, double *__restrict__ grad, const PublishArgs pub,
                                     double *__restrict__ v, double *__restrict__ a, const double *__restrict__ mass, double dt, double *__restrict__ ke,
                                     double *__restrict__ zeroOther, const double *pubSrc, double *pubDst, int pubCount)
{
    if (blockIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 2; k++)
            for (int i = threadIdx.x; i < pub.count[k]; i += blockDim.x) pub.dst[k][i] = pub.src[k][i];
        if (threadIdx.x < pubCount) pubDst[threadIdx.x] = pubSrc[threadIdx.x];
        if (zeroOther != nullptr && threadIdx.x == 0) *zeroOther = 0.0;
    }
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double local = 0.0;
    if (s < n) {
        const int at = sAtom[s];
        if (at >= 0) {
            const double mi = mass[at];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double g = gs[3 * (size_t) s + c];
                gs[3 * (size_t) s + c] = 0.0;
                grad[3 * (size_t) at + c] = g;
                const double ai = -100.0 * g / mi, vi = v[3 * (size_t) at + c] + 0.5 * dt * ai;
                a[3 * (size_t) at + c] = ai; v[3 * (size_t) at + c] = vi;
                local += mi * vi * vi;
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(ke, 0.5 * 0.01 * local);
}

// ------------------------------------------------------------------------------------------------------
// 1-4 interactions: explicit pair list, their own LJ table and electrostatic scale, never imaged
// (NBModelABFS_MMMMEnergy third call, pM/csource/NBModelABFS.c:275-294).  A few thousand pairs: plain fp64.
// ------------------------------------------------------------------------------------------------------
struct F64Factors { double v[21]; };

__global__ void k_pairs14(const int2 *__restrict__ pairs, int npairs, const double *__restrict__ x, const double *__restrict__ q, const int *__restrict__ ljtype,
                          const double2 *__restrict__ ljAB, int ntypes, F64Factors FF, double eScale, const int *__restrict__ invPerm, int ownLo, int ownHi,
                          double *gradSorted, double *acc, const double *__restrict__ spl, int splN)
{
    const double *F = FF.v;
    double eq = 0.0, el = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
        const int i = pairs[p].x, j = pairs[p].y;
        const int si = invPerm[i], sj = invPerm[j];
        if (si < ownLo || si >= ownHi) continue;             // several ranks: a pair belongs to the rank that owns its first atom
        const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > F[2]) continue;
        const double qij = eScale * q[i] * q[j];
        const double2 ab = ljAB[ljtype[i] * ntypes + ljtype[j]];
        if (spl != nullptr) {
            // spline form in fp64 exactly as the reference evaluates it: bisection (CubicSpline_EvaluateLUDST, pC/csource/CubicSpline.c:138-159)
            // and CubicSpline_FastEvaluateFG (pC/cinclude/CubicSpline.h:30-39); spl = x[n] then (y, h)[n] of the three splines
            int l = 0, u = splN - 1;
            while (u - l > 1) { const int m = (u + l) >> 1; if (spl[m] > r2) u = m; else l = m; }
            const double d = spl[u] - spl[l], sv = (r2 - spl[l]) / d, tv = (spl[u] - r2) / d;
            double f[3], g[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const double *y = spl + (size_t) (1 + 2 * k) * splN, *h = y + splN;
                const double hl = h[l] * d / 6.0, hu = h[u] * d / 6.0, yl = y[l], yu = y[u];
                f[k] = tv * yl + sv * yu + d * (tv * (tv * tv - 1.0) * hl + sv * (sv * sv - 1.0) * hu);
                g[k] = (yu - yl) / d + (-(3.0 * tv * tv - 1.0) * hl + (3.0 * sv * sv - 1.0) * hu);
            }
            eq += qij * f[0]; el += ab.x * f[1] + ab.y * f[2];
            if (gradSorted != nullptr) {
                const double dG = 2.0 * qij * g[0] + 2.0 * ab.x * g[1] + 2.0 * ab.y * g[2];
                const double gx = dG * dx, gy = dG * dy, gz = dG * dz;
                atomicAdd(&gradSorted[3 * si], gx); atomicAdd(&gradSorted[3 * si + 1], gy); atomicAdd(&gradSorted[3 * si + 2], gz);
                atomicAdd(&gradSorted[3 * sj], -gx); atomicAdd(&gradSorted[3 * sj + 1], -gy); atomicAdd(&gradSorted[3 * sj + 2], -gz);
            }
            continue;
        }
        double s = 0.0, s2 = 0.0, dF = 0.0, e1, e2;
        if (!(r2 < F[0])) { s2 = 1.0 / r2; s = sqrt(s2); }
        if (r2 > F[1]) {
            e1 = qij * (s * (F[3] - r2 * (F[4] + r2 * (F[5] + F[6] * r2))) + F[8]);
            dF += -qij * 0.5 * s * (F[3] + r2 * (F[4] + r2 * (3.0 * F[5] + 5.0 * F[6] * r2))) / r2;
        } else if (r2 > F[0]) { e1 = qij * (s + F[7]); dF += -qij * 0.5 * s / r2; }
        else { e1 = qij * (F[9] - F[10] * r2); dF += -qij * F[10]; }
        const double s6 = s2 * s2 * s2;
        if (r2 > F[1]) {
            const double l1 = s6 - F[11], l2 = (s / r2) - F[16];
            e2 = ab.x * F[12] * l1 * l1 - ab.y * F[17] * l2 * l2;
            dF += -3.0 * s6 * (2.0 * ab.x * F[12] * l1 / r2 - ab.y * F[17] * l2 / s);
        } else if (r2 > F[0]) {
            e2 = ab.x * (s6 * s6 - F[13]) - ab.y * (s6 - F[18]);
            dF += -3.0 * s6 * (2.0 * ab.x * s6 - ab.y) / r2;
        } else {
            e2 = ab.x * (F[14] - F[15] * r2) - ab.y * (F[19] - F[20] * r2);
            dF += -ab.x * F[15] + ab.y * F[20];
        }
        eq += e1; el += e2;
        if (gradSorted != nullptr) {
            const double gx = 2.0 * dF * dx, gy = 2.0 * dF * dy, gz = 2.0 * dF * dz;
            atomicAdd(&gradSorted[3 * si], gx); atomicAdd(&gradSorted[3 * si + 1], gy); atomicAdd(&gradSorted[3 * si + 2], gz);
            atomicAdd(&gradSorted[3 * sj], -gx); atomicAdd(&gradSorted[3 * sj + 1], -gy); atomicAdd(&gradSorted[3 * sj + 2], -gz);
        }
    }
    eq = warp_sum(eq); el = warp_sum(el);
    if ((threadIdx.x & 31) == 0 && (eq != 0.0 || el != 0.0)) { atomicAdd(&acc[0], eq); atomicAdd(&acc[1], el); }
}

static int g_numSMs = 0;

typedef void (*ForceKernel)(ForceArgs);
struct ForceVariant { const char *name; ForceKernel fn; int threads; };
// launch shapes (threads per CTA x resident CTAs per SM -> register budget); NBB200_FORCE_SHAPE selects one for experiments
static const ForceVariant kForceVariants[] = {
    {"128x4", k_cluster_forces<false, 128, 4, 0>, 128},      // default: 128 registers (no spills), 16 warps per SM
    {"128x5", k_cluster_forces<false, 128, 5, 0>, 128}, {"256x2", k_cluster_forces<false, 256, 2, 0>, 256},
    {"128x3", k_cluster_forces<false, 128, 3, 0>, 128}, {"64x8", k_cluster_forces<false, 64, 8, 0>, 64},
};
static const ForceVariant kForceRot = {"rot", k_cluster_forces<true, 256, 2, 0>, 256};
// spline form: [rotations][tables in shared memory (1) / global memory (0)]
static const ForceVariant kForceSpline[2][2] = {
    {{"spline-global", k_cluster_forces<false, 256, 2, 2>, 256}, {"spline", k_cluster_forces<false, 256, 2, 1>, 256}},
    {{"spline-rot-global", k_cluster_forces<true, 256, 2, 2>, 256}, {"spline-rot", k_cluster_forces<true, 256, 2, 1>, 256}},
};
constexpr size_t kSplineSmemLimit = 40 * 1024;        // per CTA, next to 67 KB of descriptor buffers: two CTAs per SM stay resident

void init_force_kernel_attributes()
{
    for (const ForceVariant &v : kForceVariants) cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(kForceRot.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int r = 0; r < 2; r++) for (int m = 0; m < 2; m++) cudaFuncSetAttribute(kForceSpline[r][m].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    g_numSMs = prop.multiProcessorCount;
}

// spline tables of the state -> device: the fp64 tables as the reference holds them (1-4 kernel) and, for the tile kernel, the cubic
// of every interval re-expanded about the fp32 point xf_l = fl(fl(l dR)^2) the kernel recomputes (u = r^2 - xf is then exact in fp32),
// coefficients rounded to fp32
bool upload_spline_tables(State &s)
{
    const int n = s.spl.points();
    if (n < 2) { set_error("spline tables are empty"); return false; }
    std::vector<double> h64((size_t) 7 * n);
    std::copy(s.spl.x.begin(), s.spl.x.end(), h64.begin());
    for (int k = 0; k < 3; k++) {
        std::copy(s.spl.y[k].begin(), s.spl.y[k].end(), h64.begin() + (size_t) (1 + 2 * k) * n);
        std::copy(s.spl.h[k].begin(), s.spl.h[k].end(), h64.begin() + (size_t) (2 + 2 * k) * n);
    }
    std::vector<float4> poly((size_t) 3 * n, make_float4(0.f, 0.f, 0.f, 0.f));
    const float dRf = (float) (s.outer / (double) (n - 1));
    for (int l = 0; l + 1 < n; l++) {
        volatile float rl = (float) l * dRf;            // the kernel's fp32 expansion point fl(fl(l dR)^2), rounded product by product
        volatile float xfv = rl * rl;
        const float xf = xfv;
        const double dlt = (double) xf - s.spl.x[l];
        float c[3][4];
        for (int k = 0; k < 3; k++) {
            double p[4];
            spline_interval_polynomial(s.spl.x, s.spl.y[k], s.spl.h[k], l, p);
            c[k][0] = (float) (p[0] + dlt * (p[1] + dlt * (p[2] + dlt * p[3])));
            c[k][1] = (float) (p[1] + dlt * (2.0 * p[2] + 3.0 * dlt * p[3]));
            c[k][2] = (float) (p[2] + 3.0 * dlt * p[3]);
            c[k][3] = (float) p[3];
        }
        for (int k = 0; k < 3; k++) poly[(size_t) k * n + l] = make_float4(c[k][0], c[k][1], c[k][2], c[k][3]);
    }
    if (!s.splF64.ensure(h64.size()) || !s.splPoly.ensure(poly.size())) return false;
    NBB_CUDA(cudaMemcpy(s.splF64.p, h64.data(), sizeof(double) * h64.size(), cudaMemcpyHostToDevice));
    NBB_CUDA(cudaMemcpy(s.splPoly.p, poly.data(), sizeof(float4) * poly.size(), cudaMemcpyHostToDevice));
    NBB_CUDA(cudaDeviceSynchronize());       // pageable copies return once staged: the kernels of s.stream must not overtake the DMA
    return true;
}

bool unsort_gradients(State &s, long s0, long s1, double *d_grad, bool assign, bool clear)
{
    if (s1 <= s0 || d_grad == nullptr || s.gs == nullptr) return true;
    const int threads = 256;
    const unsigned int blocks = (unsigned int) ((s1 - s0 + threads - 1) / threads);
    PublishArgs pub;
    for (int k = 0; k < 2; k++) { pub.src[k] = s.pubSrc[k]; pub.dst[k] = s.pubDst[k]; pub.count[k] = (s.pubSrc[k] != nullptr && s.pubDst[k] != nullptr) ? s.pubCount[k] : 0; }
    if (pub.count[0] > 0 || pub.count[1] > 0) s.pubDone = true;
    if (clear && assign && s.secondHalf.v != nullptr && s0 == 0 && s1 == s.n && s.condDisp == nullptr) {
        const State::SecondHalf &h = s.secondHalf;
        k_unsort_second_half<<<blocks, threads, 0, s.stream>>>(s.gs, s.sAtom.p, s.n, d_grad, pub, h.v, h.a, h.mass, h.dt, h.ke, h.zeroOther, h.pubSrc, h.pubDst, h.pubCount);
        s.secondHalfDone = true;
        s.launches += 1;
        return cuda_ok(cudaGetLastError(), "k_unsort_second_half");
    }
    if (clear) k_unsort_gradients<true><<<blocks, threads, 0, s.stream>>>(s.gs, s.sAtom.p, (int) s0, (int) s1, d_grad, assign ? 1 : 0, s.condDisp, s.condThr2, pub);
    else k_unsort_gradients<false><<<blocks, threads, 0, s.stream>>>(s.gs, s.sAtom.p, (int) s0, (int) s1, d_grad, assign ? 1 : 0, s.condDisp, s.condThr2, pub);
    s.launches += 1;
    return cuda_ok(cudaGetLastError(), "k_unsort_gradients");
}

// d_grad: atom-order gradient to accumulate into (nullable).  sortedOnly: leave the gradient in sorted order in s.gs (the caller
// exchanges halo ranges and unsorts its own slab, section 8e); the two are exclusive.
bool launch_forces(State &s, double *d_grad, bool sortedOnly)
{
    const bool wantGrad = d_grad != nullptr || sortedOnly;
    if (wantGrad) {
        if (s.gsExternal != nullptr) s.gs = s.gsExternal;
        else { if (!s.gradSorted.ensure(3 * (size_t) s.n)) return false; s.gs = s.gradSorted.p; }
        if (!s.gsZeroed) NBB_CUDA(cudaMemsetAsync(s.gs, 0, sizeof(double) * 3 * (size_t) s.n, s.stream));
        s.gsZeroed = false;
    }
    const int nitems = (int) s.hostCounters.itemCount;
    const size_t accumCount = (size_t) 16 * (s.nsets + 1);
    if (!s.accum.ensure(accumCount + 1)) return false;                    // + one slot that holds the work cursor: a single memset
    // fused mode: the memsets of a call are folded into neighbouring kernels (accumulators + work cursor by k_pack_records, the sorted gradient by
    // the unsort pass of the previous call).  Always inside nbb200_md_run; for ordinary calls on systems up to kFuseMaxAtoms atoms, where a call is
    // bound by the latency of its stream operations (on the 1.1 M-atom box the clearing unsort pass costs 0.13 ms more than a memset)
    static const bool fuseSmall = std::getenv("NBB200_NO_FUSE") == nullptr;
    const bool fused = s.mdFused || (fuseSmall && s.nranks == 1 && s.n <= kFuseMaxAtoms);
    const bool fusedZero = fused && nitems > 0;           // k_pack_records clears the accumulators and the cursor in fused mode
    if (!fusedZero) NBB_CUDA(cudaMemsetAsync(s.accum.p, 0, sizeof(double) * (accumCount + 1), s.stream));
    unsigned int *workCursor = reinterpret_cast<unsigned int *>(s.accum.p + accumCount);
    const double eScale = (1.0 / s.dielectric) * kE2AngstromToKJMol;
    // inside nbb200_md_run the 1-4 kernel (a few thousand pairs, fp64: ~5 us, latency bound) runs on a side stream NEXT TO the tile kernel: it
    // needs the cleared accumulators (behind k_pack_records) and nothing else of this call.  NBB200_NO_SIDE14=1: behind the tile kernel
    bool pairs14Done = false, pairs14Side = false;
    auto launch_pairs14 = [&](cudaStream_t st) {
        F64Factors FF;
        std::memcpy(FF.v, s.factors, sizeof(FF.v));
        const int threads = 128, nblk = std::max(1, std::min(1184, (s.n14 + threads - 1) / threads));
        if (s.timing) cudaEventRecord(s.ev[4], st);
        k_pairs14<<<nblk, threads, 0, st>>>(s.pairs14.p, s.n14, s.xcur, s.q64.p, s.ljtype.p, s.ljAB14.p, s.ntypes14, FF,
                                            (s.useAnalytic ? eScale : 1.0 / s.dielectric) * s.scale14, s.invPerm.p, s.ownLo, s.ownHi, wantGrad ? s.gs : nullptr,
                                            s.accum.p + 16 * s.nsets, s.useAnalytic ? nullptr : s.splF64.p, s.useAnalytic ? 0 : s.spl.points());
        if (s.timing) cudaEventRecord(s.ev[5], st);
        s.launches += 1;
        pairs14Done = true;
    };
    if (nitems > 0) {
        if (!s.recA.ensure((size_t) s.n + 1) || !s.recB.ensure((size_t) s.n + 1)) return false;
        if (g_numSMs == 0) init_force_kernel_attributes();
        // rolling prune: only when it can remove something (outer + buffer < list)
        static const double bufferOverride = []() { const char *e = std::getenv("NBB200_PRUNE_BUFFER"); return e ? std::atof(e) : -1.0; }();
        const double pruneBuffer = bufferOverride >= 0.0 ? bufferOverride : s.pruneBuffer;
        // the first energy call on freshly built lists walks the pool as built (a prune costs about as much as it saves in one call); the
        // inner pool is made at the next call on the same lists and refreshed on the device's own decision afterwards.
        // NBB200_PRUNE_EAGER=1: prune right after every rebuild
        static const bool eager = []() { const char *e = std::getenv("NBB200_PRUNE_EAGER"); return e != nullptr && std::atoi(e) != 0; }();
        bool prune = pruneBuffer > 0.0 && s.outer + pruneBuffer < s.list - 1.0e-9;
        if (prune && !eager && s.outerCallGeneration != s.numberOfUpdates) { prune = false; s.outerCallGeneration = s.numberOfUpdates; }
        const int slot = (int) (s.pruneCall & 1);
        if (prune) {
            if (!s.tileDescIn.ensure(s.tileCap * kTile) || !s.itemsIn.ensure(s.itemCap) || !s.xprune.ensure(3 * (size_t) s.n)) return false;
            if (s.pruneDisp.p == nullptr) {
                if (!s.pruneDisp.ensure(4)) return false;
                NBB_CUDA(cudaMemsetAsync(s.pruneDisp.p, 0, sizeof(unsigned long long) * 4, s.stream));
                s.pruneListGeneration = -1;
            }
        }
        const bool pruneForce = prune && (s.pruneListGeneration != s.numberOfUpdates || !s.pruneLatticeValid || std::memcmp(s.pruneLattice.v, s.lattice.M.v, sizeof(double) * 9) != 0);
        const int pthreads = 256, pblocks = (s.n + 1 + pthreads - 1) / pthreads;
        PublishArgs prePub;
        for (int k = 0; k < 2; k++) { prePub.src[k] = s.prePubSrc[k]; prePub.dst[k] = s.prePubDst[k]; prePub.count[k] = (s.prePubSrc[k] != nullptr && s.prePubDst[k] != nullptr) ? 1 : 0; }
        k_pack_records<<<pblocks, pthreads, 0, s.stream>>>(s.xcur, s.sAtom.p, s.n, s.ntypes, s.q32.p, s.ljtype.p, s.grid.lo[0], s.grid.lo[1], s.grid.lo[2], s.recA.p, s.recB.p,
                                                            fusedZero ? s.accum.p : nullptr, (int) (accumCount + 1),
                                                            (prune && !pruneForce) ? s.xprune.p : nullptr, s.pruneDisp.p, slot, prePub);
        if (prePub.count[0] > 0 || prePub.count[1] > 0 || s.prePubEvent != nullptr) s.prePubDone = true;
        if (s.prePubEvent != nullptr) cudaEventRecord(s.prePubEvent, s.stream);
        static const bool side14Off = std::getenv("NBB200_NO_SIDE14") != nullptr;
        if (s.n14 > 0 && !side14Off && !s.timing && s.mdFused) {     // inside nbb200_md_run only: for a single call the two event operations cost more than the overlap saves (measured: DHFR rebuild call 0.35 -> 0.39 ms)
            if (s.sideStream == nullptr) {
                NBB_CUDA(cudaStreamCreateWithFlags(&s.sideStream, cudaStreamNonBlocking));
                NBB_CUDA(cudaEventCreateWithFlags(&s.evPack, cudaEventDisableTiming));
                NBB_CUDA(cudaEventCreateWithFlags(&s.evSide14, cudaEventDisableTiming));
            }
            NBB_CUDA(cudaEventRecord(s.evPack, s.stream));
            NBB_CUDA(cudaStreamWaitEvent(s.sideStream, s.evPack, 0));
            launch_pairs14(s.sideStream);
            NBB_CUDA(cudaEventRecord(s.evSide14, s.sideStream));
            pairs14Side = true;
        }
        bool rot = false;                                   // any image with a genuine rotation?
        for (const RealSpaceOp &b : s.plan.baseOps) rot = rot || !b.pureTranslation;
        if (prune) {
            PruneArgs P;
            P.items = s.items.p; P.nitems = nitems; P.tileDesc = s.tileDesc.p; P.itemsIn = s.itemsIn.p; P.tileDescIn = s.tileDescIn.p;
            P.recA = s.recA.p; P.recB = s.recB.p; P.n = s.n; P.ops = s.imageOps.p;
            const double rc = s.outer + pruneBuffer;
            P.rc2 = (float) (rc * rc * (1.0 + 2.0e-5) + 1.0e-3);
            P.thr2 = 0.25 * pruneBuffer * pruneBuffer;
            P.force = pruneForce ? 1 : 0; P.slot = slot; P.pruneDisp = s.pruneDisp.p; P.x = s.xcur; P.xprune = s.xprune.p;
            const int blocks = std::max(1, std::min(g_numSMs * 8, (nitems + 7) / 8));
            if (s.timing) cudaEventRecord(s.ev[10], s.stream);
            if (rot) k_prune<true><<<blocks, 256, 0, s.stream>>>(P); else k_prune<false><<<blocks, 256, 0, s.stream>>>(P);
            if (s.timing) cudaEventRecord(s.ev[11], s.stream);
            s.launches += 1;
            s.pruneListGeneration = s.numberOfUpdates; s.pruneLattice = s.lattice.M; s.pruneLatticeValid = true;
            s.pruneCall += 1;
        }
        ForceArgs A;
        A.items = prune ? s.itemsIn.p : s.items.p; A.nitems = nitems; A.workCursor = workCursor;
        A.tileDesc = prune ? s.tileDescIn.p : s.tileDesc.p; A.chunkTiles = s.chunkTiles; { static const int e = []() { const char *v = std::getenv("NBB200_EXP"); return v ? std::atoi(v) : 0; }(); A.exp = e; } A.recA = s.recA.p; A.recB = s.recB.p; A.n = s.n;
        A.ljAB = s.ljAB.p; A.ntypes = s.ntypes; A.typeFree = s.typeFree.p;
        A.ops = s.imageOps.p;
        for (int d = 0; d < 3; d++) A.origin[d] = s.grid.lo[d];
        const double *f = s.factors;
        AbfsF32 &F = A.F;
        F.r2Damp = (float) std::min(f[0], f[1]); F.r2On = (float) f[1]; F.r2Off = (float) f[2];      // the switched region takes precedence (PairwiseInteraction.h:72-119)
        F.a = (float) f[3]; F.b = (float) f[4]; F.c = (float) f[5]; F.d = (float) f[6]; F.c3 = (float) (3.0 * f[5]); F.d5 = (float) (5.0 * f[6]);
        F.qShift1 = (float) f[7]; F.qShift2 = (float) f[8]; F.qF0 = (float) f[9]; F.qAlpha = (float) f[10];
        F.aF6 = (float) f[11]; F.aK12 = (float) f[12]; F.aShift12 = (float) f[13]; F.aF0 = (float) f[14]; F.aAlpha = (float) f[15];
        F.bF3 = (float) f[16]; F.bK6 = (float) f[17]; F.bShift6 = (float) f[18]; F.bF0 = (float) f[19]; F.bAlpha = (float) f[20];
        {   // factored switching forms (see abfs_pair)
            const double ro = s.outer, c = f[5], d = f[6], gam = (f[2] - f[1]) * (f[2] - f[1]) * (f[2] - f[1]);
            F.rOff = (float) ro;
            F.n3 = (float) (4.0 * c * ro + 20.0 * d * ro * ro * ro);
            F.n4 = (float) (-c - 15.0 * d * ro * ro);
            F.n5 = (float) (6.0 * d * ro);
            F.n6 = (float) (-d);
            F.k1 = (float) (3.0 * (f[2] - f[1]) / gam);
            F.k2 = (float) (2.0 / gam);
        }
        A.qScale = (float) eScale;
        A.splTab = nullptr; A.splN = 0; A.splInvDR = 0.f; A.splDR = 0.f;
        const bool spline = !s.useAnalytic;
        size_t splBytes = 0;
        if (spline) {
            if (!s.splValid) { set_error("spline form selected but the spline tables are not built"); return false; }
            A.splTab = s.splPoly.p; A.splN = s.spl.points();
            A.splInvDR = (float) ((double) (s.spl.points() - 1) / s.outer);
            A.splDR = (float) (s.outer / (double) (s.spl.points() - 1));
            A.qScale = (float) (1.0 / s.dielectric);          // the electrostatic spline carries the unit (PairwiseInteraction.c:474)
            splBytes = sizeof(float4) * 3 * (size_t) s.spl.points();
        }
        A.gradSorted = wantGrad ? s.gs : nullptr; A.accum = s.accum.p;
        static const bool forcedShape = std::getenv("NBB200_FORCE_SHAPE") != nullptr;
        static const ForceVariant *chosen = []() {
            const char *e = std::getenv("NBB200_FORCE_SHAPE");
            for (const ForceVariant &v : kForceVariants) if (e != nullptr && std::strcmp(e, v.name) == 0) return &v;
            return &kForceVariants[0];
        }();
        const bool splSmem = spline && splBytes <= kSplineSmemLimit;
        const size_t warpBytes = 2 * (size_t) s.chunkTiles * kTile * sizeof(unsigned int) + sizeof(IStage) + 32;
        const size_t tableBytes = kLJEntryBytes * (size_t) (s.ntypes + 1) * (s.ntypes + 1) + (splSmem ? splBytes : 0);
        auto resident_warps = [&](const ForceVariant &f, size_t &smemOut, int &perSMOut) {
            smemOut = warpBytes * (f.threads / 32) + tableBytes;
            if (smemOut > 200 * 1024) return 0;
            perSMOut = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMOut, f.fn, f.threads, smemOut);
            return perSMOut * (f.threads / 32);
        };
        // launch shape: the default unless the tables of the CTA (LJ entries of all type pairs: 41 KB for 35 types) cost it resident warps --
        // then the wide CTA, whose eight warps share one copy (DHFR: 12 -> 16 warps per SM)
        const ForceVariant *pick = spline ? &kForceSpline[rot ? 1 : 0][splSmem ? 1 : 0] : (rot ? &kForceRot : chosen);
        size_t smem = 0; int perSM = 0;
        int warps = resident_warps(*pick, smem, perSM);
        if (!spline && !rot && !forcedShape) {
            size_t smem2 = 0; int perSM2 = 0;
            const int warps2 = resident_warps(kForceVariants[2], smem2, perSM2);      // 256x2
            if (warps2 > warps) { pick = &kForceVariants[2]; smem = smem2; perSM = perSM2; warps = warps2; }
        }
        if (warps == 0) { set_error("too many LJ types for the shared-memory table"); return false; }
        const ForceVariant &v = *pick;
        const int warpsPerBlock = v.threads / 32;
        const int grid = std::max(1, std::min(g_numSMs * perSM, (nitems + warpsPerBlock - 1) / warpsPerBlock));
        if (s.timing) cudaEventRecord(s.ev[2], s.stream);
        v.fn<<<grid, v.threads, smem, s.stream>>>(A);
        if (s.timing) cudaEventRecord(s.ev[3], s.stream);
        s.launches += 2;
    }
    if (s.n14 > 0 && !pairs14Done) launch_pairs14(s.stream);
    if (pairs14Side) NBB_CUDA(cudaStreamWaitEvent(s.stream, s.evSide14, 0));      // whoever reads the accumulators next is behind the 1-4 kernel
    // device-array calls honour nbb200_set_gradient_overwrite too (the host-array call handles it with its own staging buffer: d_grad = s.grad.p)
    if (s.preUnsortEvent != nullptr && d_grad != nullptr) NBB_CUDA(cudaStreamWaitEvent(s.stream, s.preUnsortEvent, 0));
    const bool clearGs = fused && s.gsExternal == nullptr && s.nranks == 1 && d_grad != nullptr;
    if (d_grad != nullptr && !unsort_gradients(s, 0, s.n, d_grad, s.gradOverwrite && d_grad != s.grad.p && s.nranks == 1, clearGs)) return false;
    if (clearGs) s.gsZeroed = true;                          // the whole accumulator (3 n) has just been cleared by the unsort pass
    return cuda_ok(cudaGetLastError(), "force kernels");
}

}  // namespace nbb200
