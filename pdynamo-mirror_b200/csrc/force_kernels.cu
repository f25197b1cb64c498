// force_kernels.cu -- the MM/MM energy + gradient kernels (sm_100a).
//
// Replaces PairwiseInteractionABFS_MMMMEnergy (analytic branch, pM/csource/PairwiseInteraction.c:292-429, loop :374-427,
// macros pM/cinclude/PairwiseInteraction.h:72-119) as driven by NBModelABFS_MMMMEnergy (pM/csource/NBModelABFS.c:228-301)
// and MMMMImageEnergy (:1161-1313), including the image gradient rotation (:1295-1298) and the sums needed by
// SymmetryParameterGradients_ImageDerivatives (pM/csource/SymmetryParameterGradients.c:158-238).
//
// k_tile_forces: one warp per work item (= one i-block x up to 8 tiles of one image).  Lane l owns i atom l of the
// block and, per tile, j slot l.  The 32x32 tile is walked in 32 steps; at step k lane l evaluates (i = l, j = (l+k)%32)
// and then hands its j data AND its j-force accumulator to lane l-1 (warp shuffles), so both the i and the j force are
// plain register accumulations (no shared-memory atomics, Newton's third law used once per pair).  Pair math is fp32 in
// block-local coordinates (fp64 coordinates minus the i-block centre, rounded once), accumulation per tile in fp32,
// across tiles / into global memory in fp64.
#include "nbb200_internal.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace nbb200 {

struct ForceArgs {
    const WorkItem *items; int nitems; unsigned int *workCursor;
    const int *tileJ; const unsigned int *tileMask;
    const int *sAtom; int n;
    const double *x;
    const double *blockBox;
    const float *q32; const int *ljtype; const float2 *ljAB; int ntypes;
    const ImageOpDev *ops;
    AbfsF32 F; float qScale;
    double *grad; double *accum;
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

__device__ __forceinline__ PairOut abfs_pair(const AbfsF32 &F, float r2, float qij, float A, float B)
{
    float s = rsqrt_fast(r2);
#ifndef NBB_NO_NEWTON
    {   // one Newton step: MUFU.RSQ alone (~1e-7 relative) would dominate the energy error budget
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
    const float ka = plain ? 1.0f : F.aK12, xa = plain ? 0.0f : F.aF6, wa = plain ? F.aShift12 : 0.0f;
    const float kb = plain ? 1.0f : F.bK6,  xb = plain ? 0.0f : F.bF3, wb = plain ? F.bShift6 : 0.0f;
    const float la = s6 - xa, lb = s3 - xb;
    const float kla = ka * la, klb = kb * lb;
    o.e2 = fmaf(A, fmaf(kla, la, -wa), -(B * fmaf(klb, lb, -wb)));
    const float m = fmaf(2.0f * (A * kla), s6, -((B * klb) * s3));
    o.g = fmaf(6.0f * s2, m, gq);
    return o;
}

// Slow path for tiles that contain a pair inside the damped core, r^2 < r2Damp (reference: PairwiseInteraction.h:72-119,
// third branches, with s = s2 = 0): returns the CORRECTION (damped minus what abfs_pair produced) for this lane's
// accumulators c = {fxi, fyi, fzi, fxj, fyj, fzj, eq, el}; the j part is rotated home like in the main loop.
__device__ __noinline__ void damped_tile_fix(const AbfsF32 &F, unsigned int mask, const float4 *myPosq, const unsigned char *ljRow, const int *myLj,
                                            float xi, float yi, float zi, float qi, int src, float *c)
{
    float fxi = 0.f, fyi = 0.f, fzi = 0.f, fxj = 0.f, fyj = 0.f, fzj = 0.f, eq = 0.f, el = 0.f;
    for (int k = 0; k < kTile; k++) {
        const float4 p = myPosq[k];
        const float2 ab = *reinterpret_cast<const float2 *>(ljRow + myLj[k]);
        const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        if (((mask >> k) & 1u) && (r2 < F.r2Damp)) {
            const float qij = qi * p.w;
            const PairOut o = abfs_pair(F, r2, qij, ab.x, ab.y);
            const float e1 = qij * fmaf(-F.qAlpha, r2, F.qF0);
            const float e2 = ab.x * fmaf(-F.aAlpha, r2, F.aF0) - ab.y * fmaf(-F.bAlpha, r2, F.bF0);
            const float g = -2.0f * (-qij * F.qAlpha - ab.x * F.aAlpha + ab.y * F.bAlpha) - o.g;
            eq += e1 - o.e1; el += e2 - o.e2;
            const float gx = g * dx, gy = g * dy, gz = g * dz;
            fxi -= gx; fyi -= gy; fzi -= gz;
            fxj += gx; fyj += gy; fzj += gz;
        }
        fxj = __shfl_sync(0xffffffffu, fxj, src); fyj = __shfl_sync(0xffffffffu, fyj, src); fzj = __shfl_sync(0xffffffffu, fzj, src);
    }
    c[0] = fxi; c[1] = fyi; c[2] = fzi; c[3] = fxj; c[4] = fyj; c[5] = fzj; c[6] = eq; c[7] = el;
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

constexpr int kForceThreads = 256;
constexpr int kForceWarps = kForceThreads / 32;

// per-warp staging of one j tile: entries duplicated (64 slots) so that slot (lane + k) needs no wrap-around arithmetic
struct __align__(16) JStage {
    float4 posq[2 * kTile];      // x, y, z (block-local), charge
    int    ljoff[2 * kTile];     // byte offset of the LJ-table row of the j type
};

// per-warp shared scratch: staged j tile + fp64 accumulators of the work item (kept out of registers: 64 regs -> 4 CTAs/SM)
struct __align__(16) WarpScratch {
    JStage j;
    double acc[5][kTile];        // i-gradient x, y, z and the two energies of the item, one column per lane
};

template <bool kRot, int kMinBlocks>
__global__ void __launch_bounds__(kForceThreads, kMinBlocks) k_tile_forces(const __grid_constant__ ForceArgs A)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    WarpScratch *ws = reinterpret_cast<WarpScratch *>(smemRaw) + (threadIdx.x >> 5);
    JStage *stage = &ws->j;
    float2 *sLJ = reinterpret_cast<float2 *>(smemRaw + sizeof(WarpScratch) * kForceWarps);      // [ntypes*ntypes] (A, B)
    for (int i = threadIdx.x; i < A.ntypes * A.ntypes; i += blockDim.x) sLJ[i] = A.ljAB[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const AbfsF32 &F = A.F;
    const int src = (lane + 1) & 31;
    const unsigned char *ljBase = reinterpret_cast<const unsigned char *>(sLJ);
    const float4 *myPosq = stage->posq + lane;
    const int *myLj = stage->ljoff + lane;

    for (;;) {
        unsigned int it = 0;
        if (lane == 0) it = atomicAdd(A.workCursor, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= (unsigned int) A.nitems) break;
        const WorkItem wi = A.items[it];
        const ImageOpDev *op = A.ops + wi.image;
        const bool isImage = wi.image > 0;
        const bool pureT = kRot ? (op->pureTranslation != 0) : true;
        const double *centre = A.blockBox + 9 * wi.block + 6;

        // i atom of this lane
        const int si = wi.block * kTile + lane;
        const int ai = (si < A.n) ? A.sAtom[si] : -1;
        float xi = 0.f, yi = 0.f, zi = 0.f, qi = 0.f;
        const unsigned char *ljRow = ljBase;
        if (ai >= 0) {
            xi = (float) (A.x[3 * ai] - centre[0]); yi = (float) (A.x[3 * ai + 1] - centre[1]); zi = (float) (A.x[3 * ai + 2] - centre[2]);
            qi = A.q32[ai] * A.qScale;
            ljRow = ljBase + (size_t) A.ljtype[ai] * A.ntypes * sizeof(float2);
        }
#pragma unroll
        for (int c = 0; c < 5; c++) ws->acc[c][lane] = 0.0;
        double W[9];
        if (kRot) {
#pragma unroll
            for (int k = 0; k < 9; k++) W[k] = 0.0;
        }

        size_t T = (size_t) wi.tileStart * kTile + lane;
        int ajNext = A.tileJ[T];
        unsigned int maskNext = A.tileMask[T];
        for (int t = 0; t < wi.tileCount; t++) {
            const int aj = ajNext;
            const unsigned int mask = maskNext;
            if (t + 1 < wi.tileCount) { T += kTile; ajNext = A.tileJ[T]; maskNext = A.tileMask[T]; }   // descriptor of the next tile: in flight during this one
            double xj64 = 0.0, yj64 = 0.0, zj64 = 0.0;
            float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
            int lj = 0;
            if (aj >= 0) {
                xj64 = A.x[3 * aj]; yj64 = A.x[3 * aj + 1]; zj64 = A.x[3 * aj + 2];
                double px = xj64, py = yj64, pz = zj64;
                if (isImage) {
                    if (pureT) { px += op->tv[0]; py += op->tv[1]; pz += op->tv[2]; }
                    else {
                        px = op->R[0] * xj64 + op->R[1] * yj64 + op->R[2] * zj64 + op->tv[0];
                        py = op->R[3] * xj64 + op->R[4] * yj64 + op->R[5] * zj64 + op->tv[1];
                        pz = op->R[6] * xj64 + op->R[7] * yj64 + op->R[8] * zj64 + op->tv[2];
                    }
                }
                pj = make_float4((float) (px - centre[0]), (float) (py - centre[1]), (float) (pz - centre[2]), A.q32[aj]);
                lj = A.ljtype[aj] * (int) sizeof(float2);
            }
            __syncwarp();                                   // previous tile fully consumed
            stage->posq[lane] = pj; stage->posq[lane + kTile] = pj;
            stage->ljoff[lane] = lj; stage->ljoff[lane + kTile] = lj;
            __syncwarp();

            // per tile everything is fp32: 32 terms per accumulator (the energies' fp32 rounding, ~2e-5 kJ/mol per lane and tile,
            // averages to < 1e-7 of the total over the ~1e6 lane-tiles of a system); the flush to fp64 happens once per tile
            float fxi = 0.f, fyi = 0.f, fzi = 0.f, fxj = 0.f, fyj = 0.f, fzj = 0.f, eq = 0.f, el = 0.f;
            float r2min = F.r2Off;
            unsigned int mrev = __brev(mask);               // step k tests the sign bit, then shifts
#pragma unroll 8
            for (int k = 0; k < kTile; k++) {
                const float4 p = myPosq[k];                 // j slot (lane + k) % 32
                const float2 ab = *reinterpret_cast<const float2 *>(ljRow + myLj[k]);
                const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
                const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                const bool on = ((int) mrev < 0) && (r2 <= F.r2Off);
                mrev <<= 1;
                // masked pairs are evaluated AT the outer cutoff, where energy and force vanish (to ~1e-16 kJ/mol): one select
                // on the input instead of three on the outputs, and r2m doubles as the damped-core detector
                const float r2m = on ? r2 : F.r2Off;
                r2min = fminf(r2min, r2m);
                const PairOut o = abfs_pair(F, r2m, qi * p.w, ab.x, ab.y);
                eq += o.e1; el += o.e2;
                fxi = fmaf(-o.g, dx, fxi); fyi = fmaf(-o.g, dy, fyi); fzi = fmaf(-o.g, dz, fzi);      // gradient = -(force on i) = -g d
                fxj = fmaf(o.g, dx, fxj); fyj = fmaf(o.g, dy, fyj); fzj = fmaf(o.g, dz, fzj);
                // hand the j-gradient accumulator to the lane that evaluates this j slot next
                fxj = __shfl_sync(0xffffffffu, fxj, src); fyj = __shfl_sync(0xffffffffu, fyj, src); fzj = __shfl_sync(0xffffffffu, fzj, src);
            }
            if (__any_sync(0xffffffffu, r2min < F.r2Damp)) {   // damped core: practically never; patch the tile with the reference formulas
                float c[8];
                damped_tile_fix(F, mask, myPosq, ljRow, myLj, xi, yi, zi, qi, src, c);
                fxi += c[0]; fyi += c[1]; fzi += c[2]; fxj += c[3]; fyj += c[4]; fzj += c[5];
                eq += c[6]; el += c[7];
            }
            // after 32 hand-overs the accumulator of j slot `lane` is back in this lane
            ws->acc[0][lane] += (double) fxi; ws->acc[1][lane] += (double) fyi; ws->acc[2][lane] += (double) fzi;
            ws->acc[3][lane] += (double) eq;  ws->acc[4][lane] += (double) el;
            if (aj >= 0) {
                const double sc = op->scale;
                double gx = sc * (double) fxj, gy = sc * (double) fyj, gz = sc * (double) fzj;       // gradient on the (image) atom
                if (isImage) {
                    if (kRot && !pureT) {
                        W[0] += gx * xj64; W[1] += gx * yj64; W[2] += gx * zj64;
                        W[3] += gy * xj64; W[4] += gy * yj64; W[5] += gy * zj64;
                        W[6] += gz * xj64; W[7] += gz * yj64; W[8] += gz * zj64;
                        const double rx = op->R[0] * gx + op->R[3] * gy + op->R[6] * gz;           // R^T g'
                        const double ry = op->R[1] * gx + op->R[4] * gy + op->R[7] * gz;
                        const double rz = op->R[2] * gx + op->R[5] * gy + op->R[8] * gz;
                        gx = rx; gy = ry; gz = rz;
                    }
                }
                if (A.grad != nullptr) {
                    atomicAdd(&A.grad[3 * aj], gx); atomicAdd(&A.grad[3 * aj + 1], gy); atomicAdd(&A.grad[3 * aj + 2], gz);
                }
            }
        }
        const double sc = op->scale;
        const double fix = ws->acc[0][lane], fiy = ws->acc[1][lane], fiz = ws->acc[2][lane];
        if (ai >= 0 && A.grad != nullptr) {
            atomicAdd(&A.grad[3 * ai], sc * fix); atomicAdd(&A.grad[3 * ai + 1], sc * fiy); atomicAdd(&A.grad[3 * ai + 2], sc * fiz);
        }
        double *acc = A.accum + 16 * wi.image;
        const double eQ = warp_sum(ws->acc[3][lane]) * sc, eL = warp_sum(ws->acc[4][lane]) * sc;
        if (lane == 0) { atomicAdd(&acc[0], eQ); atomicAdd(&acc[1], eL); }
        if (isImage) {
            // sum over the image atoms of their gradient = minus the sum of the i-side gradients of this item (Newton's third law)
            const double G0 = -warp_sum(fix) * sc, G1 = -warp_sum(fiy) * sc, G2 = -warp_sum(fiz) * sc;
            if (lane == 0) { atomicAdd(&acc[2], G0); atomicAdd(&acc[3], G1); atomicAdd(&acc[4], G2); }
            if (kRot && !pureT) {
#pragma unroll
                for (int k = 0; k < 9; k++) { const double w = warp_sum(W[k]); if (lane == 0) atomicAdd(&acc[5 + k], w); }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// k_tile_forces_x2: the same tile walk with Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2).
// Lane (half, lmod) owns TWO i atoms of the block, lmod (element a) and lmod + 16 (element b), and every packed instruction
// works on the pairs (a, j) and (b, j) of ONE j atom, whose data are scalar (broadcast) operands: half the shared-memory
// traffic per pair of the one-atom-per-lane walk, which is what bounds a packed kernel (the LSU data pipe).  Half warp h walks
// the 16 j slots 16h .. 16h+15 in 16 steps (lane lmod sees slot 16h + (lmod + k) % 16); the scalar j accumulator travels
// inside the half warp and is back home (slot = lane) after 16 steps.  Region selection is arithmetic (p = 1 plain,
// 0 switched; pm = 1 - p) so that it stays in the packed domain: a packed instruction takes one issue slot for two pairs.
// ------------------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ float lo_of(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi_of(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct __align__(16) JStage2 {
    float4 posq[2][2 * 16];  // per half warp: its 16 j slots (x, y, z block-local, charge), duplicated for wrap-free indexing
    int    ljoff[2][2 * 16]; // byte offset (inside a table row) of the LJ entry of the j type
};

struct __align__(16) WarpScratch2 {
    JStage2 j;
    double acc[8][kTile];    // i gradient of element a (x, y, z), of element b (x, y, z), the two energies; one column per lane
    double off[4];           // per item: image translation minus block centre (x, y, z), image scale
};

// slow path of the x2 kernel for tiles with a pair inside the damped core (same contract as damped_tile_fix):
// c = {fa.xyz, fb.xyz, fj.xyz, eq, el} corrections with the main loop's signs (fa, fb: i gradients; fj: MINUS the j gradient)
__device__ __noinline__ void damped_tile_fix_x2(const AbfsF32 &F, unsigned int mm, const float4 *myPosq, const int *myLj, const unsigned char *ljRowA,
                                               const unsigned char *ljRowB, const float *xi, const float *yi, const float *zi, const float *qi, int src, float *c)
{
    float fi[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float fj[3] = {0.f, 0.f, 0.f}, eq = 0.f, el = 0.f;
    for (int k = 0; k < 16; k++) {
        const float4 pj = myPosq[k];
        const int lo = myLj[k];
        for (int h = 0; h < 2; h++) {
            const float4 ab = *reinterpret_cast<const float4 *>((h ? ljRowB : ljRowA) + lo);      // (A, -B, ., .)
            const float dx = xi[h] - pj.x, dy = yi[h] - pj.y, dz = zi[h] - pj.z;
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const bool bit = h ? ((mm & 0x8000u) != 0u) : ((int) mm < 0);
            if (bit && (r2 < F.r2Damp)) {
                const float qij = qi[h] * pj.w, Aij = ab.x, Bij = -ab.y;
                const PairOut o = abfs_pair(F, r2, qij, Aij, Bij);
                const float e1 = qij * fmaf(-F.qAlpha, r2, F.qF0);
                const float e2 = Aij * fmaf(-F.aAlpha, r2, F.aF0) - Bij * fmaf(-F.bAlpha, r2, F.bF0);
                const float g = -2.0f * (-qij * F.qAlpha - Aij * F.aAlpha + Bij * F.bAlpha) - o.g;
                eq += e1 - o.e1; el += e2 - o.e2;
                const float gx = g * dx, gy = g * dy, gz = g * dz;
                fi[h][0] -= gx; fi[h][1] -= gy; fi[h][2] -= gz;
                fj[0] -= gx; fj[1] -= gy; fj[2] -= gz;
            }
        }
        mm <<= 1;
        for (int d = 0; d < 3; d++) fj[d] = __shfl_sync(0xffffffffu, fj[d], src);
    }
    c[0] = fi[0][0]; c[1] = fi[0][1]; c[2] = fi[0][2]; c[3] = fi[1][0]; c[4] = fi[1][1]; c[5] = fi[1][2];
    c[6] = fj[0]; c[7] = fj[1]; c[8] = fj[2]; c[9] = eq; c[10] = el;
}

template <bool kRot, int kThreads, int kMinBlocks, int kUnroll>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_tile_forces_x2(const __grid_constant__ ForceArgs A)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int kWarps = kThreads / 32;
    WarpScratch2 *ws = reinterpret_cast<WarpScratch2 *>(smemRaw) + (threadIdx.x >> 5);
    JStage2 *stage = &ws->j;
    float4 *sLJ = reinterpret_cast<float4 *>(smemRaw + sizeof(WarpScratch2) * kWarps);      // [ntypes*ntypes] (A, -B, -(A aShift12 - B bShift6), 0)
    const AbfsF32 &F = A.F;
    for (int i = threadIdx.x; i < A.ntypes * A.ntypes; i += blockDim.x) {
        const float2 ab = A.ljAB[i];
        sLJ[i] = make_float4(ab.x, -ab.y, -(ab.x * F.aShift12 - ab.y * F.bShift6), 0.f);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, lmod = lane & 15, half = lane >> 4;
    // rotation source inside the half warp; bounced through shared memory so that the compiler keeps it in a register
    // instead of rematerialising it from %tid in every step
    volatile int *srcSlot = reinterpret_cast<volatile int *>(&ws->acc[0][0]) + lane;
    *srcSlot = (lane & 16) | ((lane + 1) & 15);
    const int src = *srcSlot;
    __syncwarp();
    const unsigned char *ljBase = reinterpret_cast<const unsigned char *>(sLJ);
    const float4 *myPosq = stage->posq[half] + lmod;
    const int *myLj = stage->ljoff[half] + lmod;

    // work items are claimed one ahead: the cursor atomic and the item record of the NEXT item are in flight during this one
    unsigned int itNext = 0;
    if (lane == 0) itNext = atomicAdd(A.workCursor, 1u);
    itNext = __shfl_sync(0xffffffffu, itNext, 0);
    WorkItem wiNext = A.items[min(itNext, (unsigned int) (A.nitems - 1))];
    for (;;) {
        if (itNext >= (unsigned int) A.nitems) break;
        const WorkItem wi = wiNext;
        if (lane == 0) itNext = atomicAdd(A.workCursor, 1u);
        itNext = __shfl_sync(0xffffffffu, itNext, 0);
        wiNext = A.items[min(itNext, (unsigned int) (A.nitems - 1))];
        const ImageOpDev *op = A.ops + wi.image;
        const bool isImage = wi.image > 0;
        const bool pureT = kRot ? (op->pureTranslation != 0) : true;
        const double *centre = A.blockBox + 9 * wi.block + 6;
        // fp64 offset from absolute (primary) coordinates to block-local ones, translation of the image folded in; shared per warp
        if (lane < 3) ws->off[lane] = ((isImage && pureT) ? op->tv[lane] : 0.0) - centre[lane];
        if (lane == 3) ws->off[3] = op->scale;
        __syncwarp();

        // the two i atoms of this lane: block slots lmod and lmod + 16 (both half warps hold the same two atoms)
        float xi[2] = {0.f, 0.f}, yi[2] = {0.f, 0.f}, zi[2] = {0.f, 0.f}, qi[2] = {0.f, 0.f};
        const unsigned char *ljRowA = ljBase, *ljRowB = ljBase;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int si = wi.block * kTile + lmod + 16 * h;
            const int ai = (si < A.n) ? A.sAtom[si] : -1;
            if (ai >= 0) {
                xi[h] = (float) (A.x[3 * ai] - centre[0]); yi[h] = (float) (A.x[3 * ai + 1] - centre[1]); zi[h] = (float) (A.x[3 * ai + 2] - centre[2]);
                qi[h] = A.q32[ai] * A.qScale;
                const unsigned char *row = ljBase + (size_t) A.ljtype[ai] * A.ntypes * sizeof(float4);
                if (h == 0) ljRowA = row; else ljRowB = row;
            }
        }
        const f32x2 xi2 = pk(xi[0], xi[1]), yi2 = pk(yi[0], yi[1]), zi2 = pk(zi[0], zi[1]), nqi2 = pk(-qi[0], -qi[1]);
#pragma unroll
        for (int c = 0; c < 8; c++) ws->acc[c][lane] = 0.0;
        double W[9];
        if (kRot) {
#pragma unroll
            for (int k = 0; k < 9; k++) W[k] = 0.0;
        }

        // two-deep software pipeline over the tiles of the item: the descriptor (j index, mask) of tile t+2 and the gathered
        // atom data of tile t+1 are in flight while tile t is evaluated
        size_t T = (size_t) wi.tileStart * kTile + lane;
        int ajCur = A.tileJ[T];
        unsigned int maskCur = A.tileMask[T];
        int ajB = -1;
        unsigned int maskB = 0u;
        if (wi.tileCount > 1) { ajB = A.tileJ[T + kTile]; maskB = A.tileMask[T + kTile]; }
        double gx64 = 0.0, gy64 = 0.0, gz64 = 0.0;
        float gq = 0.f;
        int gt = 0;
        if (ajCur >= 0) { gx64 = A.x[3 * ajCur]; gy64 = A.x[3 * ajCur + 1]; gz64 = A.x[3 * ajCur + 2]; gq = A.q32[ajCur]; gt = A.ljtype[ajCur]; }
        for (int t = 0; t < wi.tileCount; t++) {
            const int aj = ajCur;
            const unsigned int rotmask = maskCur;
            const double xj64 = gx64, yj64 = gy64, zj64 = gz64;
            float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
            int lj = 0;
            if (aj >= 0) {
                double px = xj64, py = yj64, pz = zj64;
                if (kRot && isImage && !pureT) {
                    px = op->R[0] * xj64 + op->R[1] * yj64 + op->R[2] * zj64 + op->tv[0];
                    py = op->R[3] * xj64 + op->R[4] * yj64 + op->R[5] * zj64 + op->tv[1];
                    pz = op->R[6] * xj64 + op->R[7] * yj64 + op->R[8] * zj64 + op->tv[2];
                }
                pj = make_float4((float) (px + ws->off[0]), (float) (py + ws->off[1]), (float) (pz + ws->off[2]), gq);
                lj = gt * (int) sizeof(float4);
            }
            // next tile: gather its atoms now, and fetch the descriptor of the tile after it
            ajCur = ajB; maskCur = maskB;
            gx64 = 0.0; gy64 = 0.0; gz64 = 0.0; gq = 0.f; gt = 0;
            if (ajCur >= 0) { gx64 = A.x[3 * ajCur]; gy64 = A.x[3 * ajCur + 1]; gz64 = A.x[3 * ajCur + 2]; gq = A.q32[ajCur]; gt = A.ljtype[ajCur]; }
            ajB = -1; maskB = 0u;
            if (t + 2 < wi.tileCount) { ajB = A.tileJ[T + 2 * kTile]; maskB = A.tileMask[T + 2 * kTile]; }
            T += kTile;

            __syncwarp();                                   // previous tile fully consumed
            stage->posq[half][lmod] = pj; stage->posq[half][lmod + 16] = pj;        // j slot = lane
            stage->ljoff[half][lmod] = lj; stage->ljoff[half][lmod + 16] = lj;
            __syncwarp();
            // masks: canonical row of an i atom (bit s <-> slot s) lives in the lane of that atom; fetch the rows of this lane's two
            // atoms, keep the 16 slots of this half warp, rotate by lmod and bit-reverse both into one word: step k tests
            // bit 31 (element a) and bit 15 (element b), then shifts left
            unsigned int mm;
            {
                const unsigned int row = __funnelshift_l(rotmask, rotmask, lane);
                const unsigned int ga = (__shfl_sync(0xffffffffu, row, lmod) >> (16 * half)) & 0xffffu;
                const unsigned int gb = (__shfl_sync(0xffffffffu, row, lmod + 16) >> (16 * half)) & 0xffffu;
                const unsigned int ma = ((ga >> lmod) | (ga << (16 - lmod))) & 0xffffu, mb = ((gb >> lmod) | (gb << (16 - lmod))) & 0xffffu;
                mm = (__brev(ma) & 0xffff0000u) | (__brev(mb) >> 16);
            }
            const unsigned int mm0 = mm;

            f32x2 fxi = 0ULL, fyi = 0ULL, fzi = 0ULL, eq = 0ULL, el = 0ULL;   // i gradients of (a, b); eq holds the NEGATED Coulomb energy
            float fxj = 0.f, fyj = 0.f, fzj = 0.f;                            // MINUS the j gradient
            float r2min = F.r2Off;
#pragma unroll kUnroll
            for (int k = 0; k < 16; k++) {
                const float4 p = myPosq[k];
                const int lo = myLj[k];
                const unsigned char *ea = ljRowA + lo, *eb = ljRowB + lo;
                const f32x2 Aij = pk(*reinterpret_cast<const float *>(ea), *reinterpret_cast<const float *>(eb));
                const f32x2 mBij = pk(*reinterpret_cast<const float *>(ea + 4), *reinterpret_cast<const float *>(eb + 4));
                const f32x2 mSij = pk(*reinterpret_cast<const float *>(ea + 8), *reinterpret_cast<const float *>(eb + 8));
                const f32x2 dx = sub2(xi2, bc(p.x)), dy = sub2(yi2, bc(p.y)), dz = sub2(zi2, bc(p.z));
                const f32x2 r2raw = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
                const bool ba = (int) mm < 0, bb = (mm & 0x8000u) != 0u;
                mm <<= 1;
                // masked pairs sit AT the outer cutoff: zero energy and force
                const float r2a = ba ? fminf(lo_of(r2raw), F.r2Off) : F.r2Off, r2b = bb ? fminf(hi_of(r2raw), F.r2Off) : F.r2Off;
                r2min = fminf(r2min, fminf(r2a, r2b));
                const f32x2 r2 = pk(r2a, r2b);
                f32x2 s = pk(rsqrt_fast(r2a), rsqrt_fast(r2b));
                {   // Newton: s <- s + (-0.5 s) (r2 s^2 - 1)
                    const f32x2 e = fma2(mul2(r2, s), s, bc(-1.0f));
                    s = fma2(mul2(s, bc(-0.5f)), e, s);
                }
                const f32x2 s2 = mul2(s, s), s3 = mul2(s, s2), s6 = mul2(s3, s3);
                const f32x2 p1 = pk(r2a <= F.r2On ? 1.0f : 0.0f, r2b <= F.r2On ? 1.0f : 0.0f);   // 1: plain region, 0: switched
                const f32x2 pm = sub2(bc(1.0f), p1);
                const f32x2 nqij = mul2(nqi2, bc(p.w));                                            // -qi qj
                // (per-element selects of the region constants were tried instead of the arithmetic blends: the compiler turns
                // them into MOV pairs, 52 instead of 43 issue slots per pair, and the kernel is 2 % slower)
                // Coulomb energy (negated): nqij (s G + p qShift1), G = p + pm t^3 C(t); with tn = r - rOff = -t the signs of the
                // odd powers live in the coefficients: t^3 C(t) = tn^2 (tn (-n3 + n4 tn - n5 tn^2 + n6 tn^3))
                const f32x2 tn = fma2(r2, s, bc(-F.rOff));
                const f32x2 Cn = fma2(fma2(fma2(bc(F.n6), tn, bc(-F.n5)), tn, bc(F.n4)), tn, bc(-F.n3));
                const f32x2 t3C = mul2(mul2(tn, tn), mul2(tn, Cn));
                const f32x2 G = fma2(pm, t3C, p1);
                eq = fma2(nqij, fma2(s, G, mul2(p1, bc(F.qShift1))), eq);
                // Coulomb force factor (negated): nqij s^3 Q, Q = p + pm u^2 (k1 - k2 u)
                const f32x2 u = sub2(bc(F.r2Off), r2);
                const f32x2 Qs = mul2(mul2(u, u), fma2(bc(-F.k2), u, bc(F.k1)));
                const f32x2 Q = fma2(pm, Qs, p1);
                const f32x2 mgq = mul2(mul2(nqij, s3), Q);
                // Lennard-Jones: A ka (s6 - xa)^2 - B kb (s3 - xb)^2 - p (A aShift12 - B bShift6), X = A ka la, Y = -B kb lb
                const f32x2 la = fma2(pm, bc(-F.aF6), s6), lb = fma2(pm, bc(-F.bF3), s3);
                const f32x2 Xa = mul2(Aij, mul2(fma2(pm, bc(F.aK12 - 1.0f), bc(1.0f)), la)), Yb = mul2(mBij, mul2(fma2(pm, bc(F.bK6 - 1.0f), bc(1.0f)), lb));
                el = fma2(Xa, la, el); el = fma2(Yb, lb, el); el = fma2(p1, mSij, el);
                const f32x2 mmLJ = fma2(bc(2.0f), mul2(Xa, s6), mul2(Yb, s3));
                const f32x2 mg = fma2(mul2(s2, bc(-6.0f)), mmLJ, mgq);                               // -g
                fxi = fma2(mg, dx, fxi); fyi = fma2(mg, dy, fyi); fzi = fma2(mg, dz, fzi);
                fxj = fmaf(lo_of(mg), lo_of(dx), fxj); fyj = fmaf(lo_of(mg), lo_of(dy), fyj); fzj = fmaf(lo_of(mg), lo_of(dz), fzj);
                fxj = fmaf(hi_of(mg), hi_of(dx), fxj); fyj = fmaf(hi_of(mg), hi_of(dy), fyj); fzj = fmaf(hi_of(mg), hi_of(dz), fzj);
                // hand the j accumulator to the lane that evaluates this j slot next
                fxj = __shfl_sync(0xffffffffu, fxj, src); fyj = __shfl_sync(0xffffffffu, fyj, src); fzj = __shfl_sync(0xffffffffu, fzj, src);
            }
            // after 16 hand-overs inside the half warp the accumulator of j slot `lane` is back in this lane
            float ca[3] = {lo_of(fxi), lo_of(fyi), lo_of(fzi)}, cb[3] = {hi_of(fxi), hi_of(fyi), hi_of(fzi)};
            float ceq = -(lo_of(eq) + hi_of(eq)), cel = lo_of(el) + hi_of(el);
            if (__any_sync(0xffffffffu, r2min < F.r2Damp)) {   // damped core: practically never; patch the tile with the reference formulas
                float c[11];
                damped_tile_fix_x2(F, mm0, myPosq, myLj, ljRowA, ljRowB, xi, yi, zi, qi, src, c);
                ca[0] += c[0]; ca[1] += c[1]; ca[2] += c[2]; cb[0] += c[3]; cb[1] += c[4]; cb[2] += c[5];
                fxj += c[6]; fyj += c[7]; fzj += c[8]; ceq += c[9]; cel += c[10];
            }
            ws->acc[0][lane] += (double) ca[0]; ws->acc[1][lane] += (double) ca[1]; ws->acc[2][lane] += (double) ca[2];
            ws->acc[3][lane] += (double) cb[0]; ws->acc[4][lane] += (double) cb[1]; ws->acc[5][lane] += (double) cb[2];
            ws->acc[6][lane] += (double) ceq;   ws->acc[7][lane] += (double) cel;
            if (aj >= 0) {
                const double sc = ws->off[3];
                double gx = -sc * (double) fxj, gy = -sc * (double) fyj, gz = -sc * (double) fzj;    // gradient on the (image) atom
                if (isImage) {
                    if (kRot && !pureT) {
                        W[0] += gx * xj64; W[1] += gx * yj64; W[2] += gx * zj64;
                        W[3] += gy * xj64; W[4] += gy * yj64; W[5] += gy * zj64;
                        W[6] += gz * xj64; W[7] += gz * yj64; W[8] += gz * zj64;
                        const double rx = op->R[0] * gx + op->R[3] * gy + op->R[6] * gz;           // R^T g'
                        const double ry = op->R[1] * gx + op->R[4] * gy + op->R[7] * gz;
                        const double rz = op->R[2] * gx + op->R[5] * gy + op->R[8] * gz;
                        gx = rx; gy = ry; gz = rz;
                    }
                }
                if (A.grad != nullptr) {
                    atomicAdd(&A.grad[3 * aj], gx); atomicAdd(&A.grad[3 * aj + 1], gy); atomicAdd(&A.grad[3 * aj + 2], gz);
                }
            }
        }
        // the two half warps hold partial i gradients (their 16 j slots each) of the same two atoms: exchange, and let lane l
        // finish block atom l (half 0: element a = lmod, half 1: element b = lmod + 16)
        const double sc = ws->off[3];
        double fix, fiy, fiz;
        {
            const double ax = ws->acc[0][lane], ay = ws->acc[1][lane], az = ws->acc[2][lane];
            const double bx = ws->acc[3][lane], by = ws->acc[4][lane], bz = ws->acc[5][lane];
            const double ox = __shfl_xor_sync(0xffffffffu, half ? ax : bx, 16), oy = __shfl_xor_sync(0xffffffffu, half ? ay : by, 16);
            const double oz = __shfl_xor_sync(0xffffffffu, half ? az : bz, 16);
            fix = (half ? bx : ax) + ox; fiy = (half ? by : ay) + oy; fiz = (half ? bz : az) + oz;
        }
        {
            const int si = wi.block * kTile + lane;
            const int ai = (si < A.n) ? A.sAtom[si] : -1;
            if (ai >= 0 && A.grad != nullptr) {
                atomicAdd(&A.grad[3 * ai], sc * fix); atomicAdd(&A.grad[3 * ai + 1], sc * fiy); atomicAdd(&A.grad[3 * ai + 2], sc * fiz);
            }
        }
        double *acc = A.accum + 16 * wi.image;
        const double eQ = warp_sum(ws->acc[6][lane]) * sc, eL = warp_sum(ws->acc[7][lane]) * sc;
        if (lane == 0) { atomicAdd(&acc[0], eQ); atomicAdd(&acc[1], eL); }
        if (isImage) {
            // sum over the image atoms of their gradient = minus the sum of the i-side gradients of this item (Newton's third law)
            const double G0 = -warp_sum(fix) * sc, G1 = -warp_sum(fiy) * sc, G2 = -warp_sum(fiz) * sc;
            if (lane == 0) { atomicAdd(&acc[2], G0); atomicAdd(&acc[3], G1); atomicAdd(&acc[4], G2); }
            if (kRot && !pureT) {
#pragma unroll
                for (int k = 0; k < 9; k++) { const double w = warp_sum(W[k]); if (lane == 0) atomicAdd(&acc[5 + k], w); }
            }
        }
        __syncwarp();                                       // acc columns are re-zeroed by the next item
    }
}

// ------------------------------------------------------------------------------------------------------
// 1-4 interactions: explicit pair list, their own LJ table and electrostatic scale, never imaged
// (NBModelABFS_MMMMEnergy third call, pM/csource/NBModelABFS.c:275-294).  A few thousand pairs: plain fp64.
// ------------------------------------------------------------------------------------------------------
struct F64Factors { double v[21]; };

__global__ void k_pairs14(const int2 *__restrict__ pairs, int npairs, const double *__restrict__ x, const double *__restrict__ q, const int *__restrict__ ljtype,
                          const double2 *__restrict__ ljAB, int ntypes, F64Factors FF, double eScale, double *grad, double *acc)
{
    const double *F = FF.v;
    double eq = 0.0, el = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
        const int i = pairs[p].x, j = pairs[p].y;
        const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > F[2]) continue;
        const double qij = eScale * q[i] * q[j];
        const double2 ab = ljAB[ljtype[i] * ntypes + ljtype[j]];
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
        if (grad != nullptr) {
            const double gx = 2.0 * dF * dx, gy = 2.0 * dF * dy, gz = 2.0 * dF * dz;
            atomicAdd(&grad[3 * i], gx); atomicAdd(&grad[3 * i + 1], gy); atomicAdd(&grad[3 * i + 2], gz);
            atomicAdd(&grad[3 * j], -gx); atomicAdd(&grad[3 * j + 1], -gy); atomicAdd(&grad[3 * j + 2], -gz);
        }
    }
    eq = warp_sum(eq); el = warp_sum(el);
    if ((threadIdx.x & 31) == 0 && (eq != 0.0 || el != 0.0)) { atomicAdd(&acc[0], eq); atomicAdd(&acc[1], el); }
}

static int g_forceBlocksPerSM = 0, g_numSMs = 0;

typedef void (*X2Kernel)(ForceArgs);
struct X2Variant { const char *name; X2Kernel fn; int threads; };
static const X2Variant kX2Variants[] = {
    {"128x4u2", k_tile_forces_x2<false, 128, 4, 2>, 128}, {"128x4u4", k_tile_forces_x2<false, 128, 4, 4>, 128},
    {"128x5u2", k_tile_forces_x2<false, 128, 5, 2>, 128}, {"128x5u4", k_tile_forces_x2<false, 128, 5, 4>, 128},
    {"128x6u2", k_tile_forces_x2<false, 128, 6, 2>, 128}, {"128x6u1", k_tile_forces_x2<false, 128, 6, 1>, 128},
    {"128x5u8", k_tile_forces_x2<false, 128, 5, 8>, 128}, {"128x5u16", k_tile_forces_x2<false, 128, 5, 16>, 128},
    {"128x4u8", k_tile_forces_x2<false, 128, 4, 8>, 128}, {"128x4u16", k_tile_forces_x2<false, 128, 4, 16>, 128},
    {"256x2u4", k_tile_forces_x2<false, 256, 2, 4>, 256}, {"128x3u4", k_tile_forces_x2<false, 128, 3, 4>, 128},
};
static const X2Variant kX2Rot = {"rot", k_tile_forces_x2<true, 128, 3, 2>, 128};

void init_force_kernel_attributes()
{
    cudaFuncSetAttribute(k_tile_forces<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_tile_forces<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_tile_forces<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (const X2Variant &v : kX2Variants) cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(kX2Rot.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    g_numSMs = prop.multiProcessorCount;
}

bool launch_forces(State &s, double *d_grad)
{
    const int nitems = (int) s.hostCounters.itemCount;
    const size_t accumCount = (size_t) 16 * (s.nsets + 1);
    if (!s.accum.ensure(accumCount + 1)) return false;                    // + one slot that holds the work cursor: a single memset
    NBB_CUDA(cudaMemsetAsync(s.accum.p, 0, sizeof(double) * (accumCount + 1), s.stream));
    unsigned int *workCursor = reinterpret_cast<unsigned int *>(s.accum.p + accumCount);
    const double eScale = (1.0 / s.dielectric) * kE2AngstromToKJMol;
    if (nitems > 0) {
        ForceArgs A;
        A.items = s.items.p; A.nitems = nitems; A.workCursor = workCursor;
        A.tileJ = s.tileJ.p; A.tileMask = s.tileMask.p; A.sAtom = s.sAtom.p; A.n = s.n;
        A.x = s.xcur; A.blockBox = s.blockBox.p;
        A.q32 = s.q32.p; A.ljtype = s.ljtype.p; A.ljAB = s.ljAB.p; A.ntypes = s.ntypes;
        A.ops = s.imageOps.p;
        const double *f = s.factors;
        AbfsF32 &F = A.F;
        F.r2Damp = (float) f[0]; F.r2On = (float) f[1]; F.r2Off = (float) f[2];
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
        A.grad = d_grad; A.accum = s.accum.p;
        const size_t smem = sizeof(WarpScratch) * kForceWarps + sizeof(float2) * (size_t) s.ntypes * s.ntypes;
        if (smem > 160 * 1024) { set_error("too many LJ types for the shared-memory table"); return false; }
        if (g_numSMs == 0) init_force_kernel_attributes();
        bool rot = false;                                   // any image with a genuine rotation?
        for (const RealSpaceOp &b : s.plan.baseOps) rot = rot || !b.pureTranslation;
        int perSM = 0;
        static const int scalarBlocks = []() { const char *e = std::getenv("NBB200_SCALAR_BLOCKS"); return (e && std::atoi(e) == 4) ? 4 : 3; }();
        void (*skern)(ForceArgs) = rot ? k_tile_forces<true, 2> : (scalarBlocks == 4 ? k_tile_forces<false, 4> : k_tile_forces<false, 3>);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, skern, kForceThreads, smem);
        if (perSM < 1) perSM = 1;
        g_forceBlocksPerSM = perSM;
        const int warpsPerBlock = kForceThreads / 32;
        const int grid = std::max(1, std::min(g_numSMs * perSM, (nitems + warpsPerBlock - 1) / warpsPerBlock));
        if (s.timing) cudaEventRecord(s.ev[2], s.stream);
        // default: the scalar-instruction kernel (24 warps/SM, issue bound at ~84 %); NBB200_FORCE_KERNEL=x2 selects the packed
        // f32x2 variant (fewer issue slots, but 128 registers -> 16 warps/SM and latency bound: 5 % slower on B200, see profiles/)
        static const bool useX2 = []() { const char *e = std::getenv("NBB200_FORCE_KERNEL"); return e != nullptr && std::strcmp(e, "x2") == 0; }();
        if (useX2) {
            static const X2Variant *chosen = []() {
                const char *e = std::getenv("NBB200_X2_VARIANT");
                for (const X2Variant &v : kX2Variants) if (e != nullptr && std::strcmp(e, v.name) == 0) return &v;
                return &kX2Variants[0];
            }();
            const X2Variant &v = rot ? kX2Rot : *chosen;
            const int warps2 = v.threads / 32;
            const size_t smem2 = sizeof(WarpScratch2) * warps2 + sizeof(float4) * (size_t) s.ntypes * s.ntypes;
            if (smem2 > 160 * 1024) { set_error("too many LJ types for the shared-memory table"); return false; }
            int perSM2 = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM2, v.fn, v.threads, smem2);
            if (perSM2 < 1) perSM2 = 1;
            const int grid2 = std::max(1, std::min(g_numSMs * perSM2, (nitems + warps2 - 1) / warps2));
            v.fn<<<grid2, v.threads, smem2, s.stream>>>(A);
        } else skern<<<grid, kForceThreads, smem, s.stream>>>(A);
        if (s.timing) cudaEventRecord(s.ev[3], s.stream);
        s.launches += 1;
    }
    if (s.n14 > 0 && s.rank == 0) {
        F64Factors FF;
        std::memcpy(FF.v, s.factors, sizeof(FF.v));
        const int threads = 128, nblk = std::max(1, std::min(1184, (s.n14 + threads - 1) / threads));
        if (s.timing) cudaEventRecord(s.ev[4], s.stream);
        k_pairs14<<<nblk, threads, 0, s.stream>>>(s.pairs14.p, s.n14, s.xcur, s.q64.p, s.ljtype.p, s.ljAB14.p, s.ntypes14, FF,
                                                    eScale * s.scale14, d_grad, s.accum.p + 16 * s.nsets);
        if (s.timing) cudaEventRecord(s.ev[5], s.stream);
        s.launches += 1;
    }
    return cuda_ok(cudaGetLastError(), "force kernels");
}

}  // namespace nbb200
