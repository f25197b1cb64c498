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
    {   // one Newton step: MUFU.RSQ alone (~1e-7 relative) would dominate the energy error budget
        const float rr = r2 * s;
        s = fmaf(0.5f * s, fmaf(-rr, s, 1.0f), s);
    }
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

template <bool kRot>
__global__ void __launch_bounds__(kForceThreads, kRot ? 2 : 3) k_tile_forces(ForceArgs A)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    JStage *stage = reinterpret_cast<JStage *>(smemRaw) + (threadIdx.x >> 5);
    float2 *sLJ = reinterpret_cast<float2 *>(smemRaw + sizeof(JStage) * kForceWarps);      // [ntypes*ntypes] (A, B)
    for (int i = threadIdx.x; i < A.ntypes * A.ntypes; i += blockDim.x) sLJ[i] = A.ljAB[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const AbfsF32 F = A.F;
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
        const double cx = A.blockBox[9 * wi.block + 6], cy = A.blockBox[9 * wi.block + 7], cz = A.blockBox[9 * wi.block + 8];
        const double sc = op->scale;

        // i atom of this lane
        const int si = wi.block * kTile + lane;
        const int ai = (si < A.n) ? A.sAtom[si] : -1;
        float xi = 0.f, yi = 0.f, zi = 0.f, qi = 0.f;
        const unsigned char *ljRow = ljBase;
        if (ai >= 0) {
            xi = (float) (A.x[3 * ai] - cx); yi = (float) (A.x[3 * ai + 1] - cy); zi = (float) (A.x[3 * ai + 2] - cz);
            qi = A.q32[ai] * A.qScale;
            ljRow = ljBase + (size_t) A.ljtype[ai] * A.ntypes * sizeof(float2);
        }
        double fix = 0.0, fiy = 0.0, fiz = 0.0, eQ = 0.0, eL = 0.0;
        double G0 = 0.0, G1 = 0.0, G2 = 0.0, W[9];
        if (kRot) {
#pragma unroll
            for (int k = 0; k < 9; k++) W[k] = 0.0;
        }

        for (int t = 0; t < wi.tileCount; t++) {
            const size_t T = ((size_t) wi.tileStart + t) * kTile + lane;
            const int aj = A.tileJ[T];
            const unsigned int mask = A.tileMask[T];
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
                pj = make_float4((float) (px - cx), (float) (py - cy), (float) (pz - cz), A.q32[aj]);
                lj = A.ljtype[aj] * (int) sizeof(float2);
            }
            __syncwarp();                                   // previous tile fully consumed
            stage->posq[lane] = pj; stage->posq[lane + kTile] = pj;
            stage->ljoff[lane] = lj; stage->ljoff[lane + kTile] = lj;
            __syncwarp();

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
                if ((k & 3) == 3) { eQ += (double) eq; eL += (double) el; eq = 0.f; el = 0.f; }   // fp32 partial sums stay short
                const float gx = o.g * dx, gy = o.g * dy, gz = o.g * dz;      // force on i; the energy gradient is the negative
                fxi -= gx; fyi -= gy; fzi -= gz;
                fxj += gx; fyj += gy; fzj += gz;
                // hand the j-gradient accumulator to the lane that evaluates this j slot next
                fxj = __shfl_sync(0xffffffffu, fxj, src); fyj = __shfl_sync(0xffffffffu, fyj, src); fzj = __shfl_sync(0xffffffffu, fzj, src);
            }
            if (__any_sync(0xffffffffu, r2min < F.r2Damp)) {   // damped core: practically never; patch the tile with the reference formulas
                float c[8];
                damped_tile_fix(F, mask, myPosq, ljRow, myLj, xi, yi, zi, qi, src, c);
                fxi += c[0]; fyi += c[1]; fzi += c[2]; fxj += c[3]; fyj += c[4]; fzj += c[5];
                eQ += (double) c[6]; eL += (double) c[7];
            }
            // after 32 hand-overs the accumulator of j slot `lane` is back in this lane
            fix += (double) fxi; fiy += (double) fyi; fiz += (double) fzi;
            if (aj >= 0) {
                double gx = sc * (double) fxj, gy = sc * (double) fyj, gz = sc * (double) fzj;       // gradient on the (image) atom
                if (isImage) {
                    G0 += gx; G1 += gy; G2 += gz;
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
        if (ai >= 0 && A.grad != nullptr) {
            atomicAdd(&A.grad[3 * ai], sc * fix); atomicAdd(&A.grad[3 * ai + 1], sc * fiy); atomicAdd(&A.grad[3 * ai + 2], sc * fiz);
        }
        double *acc = A.accum + 16 * wi.image;
        eQ = warp_sum(eQ) * sc; eL = warp_sum(eL) * sc;
        if (lane == 0) { atomicAdd(&acc[0], eQ); atomicAdd(&acc[1], eL); }
        if (isImage) {
            G0 = warp_sum(G0); G1 = warp_sum(G1); G2 = warp_sum(G2);
            if (lane == 0) { atomicAdd(&acc[2], G0); atomicAdd(&acc[3], G1); atomicAdd(&acc[4], G2); }
            if (kRot && !pureT) {
#pragma unroll
                for (int k = 0; k < 9; k++) { const double w = warp_sum(W[k]); if (lane == 0) atomicAdd(&acc[5 + k], w); }
            }
        }
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

void init_force_kernel_attributes()
{
    cudaFuncSetAttribute(k_tile_forces<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_tile_forces<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
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
    if (!s.accum.ensure(accumCount)) return false;
    NBB_CUDA(cudaMemsetAsync(s.accum.p, 0, sizeof(double) * accumCount, s.stream));
    NBB_CUDA(cudaMemsetAsync(&s.counters->workCursor, 0, sizeof(unsigned int), s.stream));
    const double eScale = (1.0 / s.dielectric) * kE2AngstromToKJMol;
    if (nitems > 0) {
        ForceArgs A;
        A.items = s.items.p; A.nitems = nitems; A.workCursor = &s.counters->workCursor;
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
        const size_t smem = sizeof(JStage) * kForceWarps + sizeof(float2) * (size_t) s.ntypes * s.ntypes;
        if (smem > 160 * 1024) { set_error("too many LJ types for the shared-memory table"); return false; }
        if (g_numSMs == 0) init_force_kernel_attributes();
        bool rot = false;                                   // any image with a genuine rotation?
        for (const RealSpaceOp &b : s.plan.baseOps) rot = rot || !b.pureTranslation;
        int perSM = 0;
        if (rot) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_tile_forces<true>, kForceThreads, smem);
        else     cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_tile_forces<false>, kForceThreads, smem);
        if (perSM < 1) perSM = 1;
        g_forceBlocksPerSM = perSM;
        const int warpsPerBlock = kForceThreads / 32;
        const int grid = std::max(1, std::min(g_numSMs * perSM, (nitems + warpsPerBlock - 1) / warpsPerBlock));
        if (s.timing) cudaEventRecord(s.ev[2], s.stream);
        if (rot) k_tile_forces<true><<<grid, kForceThreads, smem, s.stream>>>(A);
        else     k_tile_forces<false><<<grid, kForceThreads, smem, s.stream>>>(A);
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
