// force_kernels.cu -- the MM/MM energy + gradient kernels (sm_100a).
//
// Replaces PairwiseInteractionABFS_MMMMEnergy (analytic branch, pM/csource/PairwiseInteraction.c:292-429, loop :374-427,
// macros pM/cinclude/PairwiseInteraction.h:72-119) as driven by NBModelABFS_MMMMEnergy (pM/csource/NBModelABFS.c:228-301)
// and MMMMImageEnergy (:1161-1313), including the image gradient rotation (:1295-1298) and the sums needed by
// SymmetryParameterGradients_ImageDerivatives (pM/csource/SymmetryParameterGradients.c:158-238).
//
// k_cluster_forces: one warp per work item (= one i-cluster of 8 sorted atoms x up to `chunkTiles` tiles of 32 j slots of one
// image).  Lane (g, m) owns cluster atom m and, per tile, j slot `lane`; group g (8 lanes) walks the 8 slots 8g .. 8g+7 in 8
// steps: at step k lane (g, m) evaluates (i = m, j = 8g + (m + k) % 8) and then hands its j data position AND its j-force
// accumulator to the next lane of the group (warp shuffles), so both the i and the j force are plain register accumulations
// (no shared-memory atomics, Newton's third law used once per pair).  All atom data come from per-call RECORDS in sorted
// order (two float4 per atom: grid-relative fp32 coordinates split as K + xl with K a multiple of 8 A, charge, LJ type), so
// the j gather of a tile is two coalesced 128-bit loads per lane; pair math is fp32 in cluster-local coordinates
// (exact K differences plus one rounding), accumulation per 32 steps in fp32, across tiles / into global memory in fp64.
// Gradients are accumulated in sorted order and scattered to atom order by k_unsort_gradients.
#include "nbb200_internal.h"
#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>

namespace nbb200 {

struct ForceArgs {
    const WorkItem *items; int nitems; unsigned int *workCursor;
    const unsigned int *tileDesc;
    const float4 *recA, *recB; int n;
    const float2 *ljAB; int ntypes;
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

// Slow path for tiles that contain a pair inside the damped core, r^2 < r2Damp (reference: PairwiseInteraction.h:72-119,
// third branches, with s = s2 = 0): returns the CORRECTION (damped minus what abfs_pair produced) for this lane's
// accumulators c = {fxi, fyi, fzi, fxj, fyj, fzj, eq, el}; the j part is rotated home like in the main loop.
__device__ __noinline__ void damped_tile_fix(const AbfsF32 &F, unsigned int mask, const float4 *myPosq, const unsigned char *ljRow, const int *myLj,
                                            float xi, float yi, float zi, float qi, int src, float *c)
{
    float fxi = 0.f, fyi = 0.f, fzi = 0.f, fxj = 0.f, fyj = 0.f, fzj = 0.f, eq = 0.f, el = 0.f;
    for (int k = 0; k < kCluster; k++) {
        const float4 p = myPosq[k];
        const float4 ab = *reinterpret_cast<const float4 *>(ljRow + myLj[k]);
        const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
        const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        if (((mask >> k) & 1u) && (r2 < F.r2Damp)) {
            const float qij = qi * p.w;
            const PairOut o = abfs_pair(F, r2, qij, ab.x, ab.y, ab.z);
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

constexpr int kGroups = kTile / kCluster;            // slot groups of a tile = lanes per cluster atom
constexpr int kFlushTiles = 4;                       // fp32 accumulators are flushed to fp64 every 4 tiles = 32 terms

// per-warp staging of one j tile: per group of 8 slots the entries are duplicated (16) so that slot (m + k) needs no wrap-around
struct __align__(16) JStage {
    float4 posq[kGroups][2 * kCluster];      // x, y, z (cluster-local), charge
    int    ljoff[kGroups][2 * kCluster];     // byte offset of the LJ-table row of the j type
};

// per-warp shared scratch: staged j tile + fp64 accumulators of the work item (kept out of registers)
struct __align__(16) WarpScratch {
    JStage j;
    double acc[5][kTile];        // i-gradient x, y, z and the two energies of the item, one column per lane
};

// kForm: 0 = analytic ABFS formulas, 1 = spline tables staged in shared memory, 2 = spline tables read from global memory (L1)
template <bool kRot, int kThreads, int kMinBlocks, int kForm>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_cluster_forces(const __grid_constant__ ForceArgs A)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    WarpScratch *ws = reinterpret_cast<WarpScratch *>(smemRaw) + (threadIdx.x >> 5);
    JStage *stage = &ws->j;
    float4 *sLJ = reinterpret_cast<float4 *>(smemRaw + sizeof(WarpScratch) * (kThreads / 32));  // [ntypes*ntypes] (A, B, A aShift12 - B bShift6, 0)
    for (int i = threadIdx.x; i < A.ntypes * A.ntypes; i += blockDim.x) {
        const float2 ab = A.ljAB[i];
        sLJ[i] = make_float4(ab.x, ab.y, ab.x * A.F.aShift12 - ab.y * A.F.bShift6, 0.f);
    }
    const float4 *splTab = A.splTab;
    if (kForm == 1) {
        float4 *sTab = sLJ + A.ntypes * A.ntypes;
        for (int i = threadIdx.x; i < 3 * A.splN; i += blockDim.x) sTab[i] = A.splTab[i];
        splTab = sTab;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, m = lane & (kCluster - 1), g = lane >> 3;
    const AbfsF32 &F = A.F;
    const int src = (lane & 24) | ((lane + 1) & 7);
    const unsigned char *ljBase = reinterpret_cast<const unsigned char *>(sLJ);
    const float4 *myPosq = stage->posq[g] + m;
    const int *myLj = stage->ljoff[g] + m;

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
        const double sc = op->scale;

        // cluster-local frame: K of the first cluster atom (a multiple of 8 A, exact in fp32)
        const int s0 = wi.block * kCluster;
        const float4 kref = A.recB[s0];
        // i atom of this lane (the four groups hold the same 8 atoms)
        const int si = s0 + m;
        const bool ivalid = si < A.n;
        float xi = 0.f, yi = 0.f, zi = 0.f, qi = 0.f;
        const unsigned char *ljRow = ljBase;
        if (ivalid) {
            const float4 ra = A.recA[si], rb = A.recB[si];
            xi = (rb.x - kref.x) + ra.x; yi = (rb.y - kref.y) + ra.y; zi = (rb.z - kref.z) + ra.z;
            qi = ra.w * A.qScale;
            ljRow = ljBase + (size_t) __float_as_int(rb.w) * A.ntypes * sizeof(float4);
        }
        // j offsets: X'_local = (K_j + ok) + (xl_j + ol), ok exact
        float okx = -kref.x, oky = -kref.y, okz = -kref.z, olx = 0.f, oly = 0.f, olz = 0.f;
        if (isImage && pureT) { okx += op->kt[0]; oky += op->kt[1]; okz += op->kt[2]; olx = op->tl[0]; oly = op->tl[1]; olz = op->tl[2]; }
#pragma unroll
        for (int c = 0; c < 5; c++) ws->acc[c][lane] = 0.0;
        double W[9];
        if (kRot) {
#pragma unroll
            for (int k = 0; k < 9; k++) W[k] = 0.0;
        }

        // software pipeline over the tiles: the descriptor of tile t+2 and the records of tile t+1 are in flight during tile t
        size_t T = (size_t) wi.tileStart * kTile + lane;
        unsigned int dCur = A.tileDesc[T];
        unsigned int dNext = (wi.tileCount > 1) ? A.tileDesc[T + kTile] : kEmptySlot;
        float4 ja = make_float4(0.f, 0.f, 0.f, 0.f), jb = ja;
        if ((dCur & kEmptySlot) != kEmptySlot) { ja = A.recA[dCur & kEmptySlot]; jb = A.recB[dCur & kEmptySlot]; }
        float fxi = 0.f, fyi = 0.f, fzi = 0.f, eq = 0.f, el = 0.f;
        for (int t = 0; t < wi.tileCount; t++) {
            const unsigned int d = dCur;
            const float4 a = ja, b = jb;
            dCur = dNext;
            if ((dCur & kEmptySlot) != kEmptySlot) { ja = A.recA[dCur & kEmptySlot]; jb = A.recB[dCur & kEmptySlot]; }
            dNext = (t + 2 < wi.tileCount) ? A.tileDesc[T + 2 * kTile] : kEmptySlot;
            T += kTile;
            const unsigned int sj = d & kEmptySlot, mask = d >> 24;
            const bool has = sj != kEmptySlot;
            float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
            int lj = 0;
            double xj64 = 0.0, yj64 = 0.0, zj64 = 0.0;          // primary coordinates of j (rotations only)
            if (has) {
                if (kRot && isImage && !pureT) {
                    const double X = (double) b.x + (double) a.x, Y = (double) b.y + (double) a.y, Z = (double) b.z + (double) a.z;
                    xj64 = X + A.origin[0]; yj64 = Y + A.origin[1]; zj64 = Z + A.origin[2];
                    const double px = op->R[0] * X + op->R[1] * Y + op->R[2] * Z + op->cr[0];
                    const double py = op->R[3] * X + op->R[4] * Y + op->R[5] * Z + op->cr[1];
                    const double pz = op->R[6] * X + op->R[7] * Y + op->R[8] * Z + op->cr[2];
                    pj = make_float4((float) (px - (double) kref.x), (float) (py - (double) kref.y), (float) (pz - (double) kref.z), a.w);
                } else pj = make_float4((b.x + okx) + (a.x + olx), (b.y + oky) + (a.y + oly), (b.z + okz) + (a.z + olz), a.w);
                lj = __float_as_int(b.w) * (int) sizeof(float4);
            }
            __syncwarp();                                   // previous tile fully consumed
            stage->posq[g][m] = pj; stage->posq[g][m + kCluster] = pj;
            stage->ljoff[g][m] = lj; stage->ljoff[g][m + kCluster] = lj;
            __syncwarp();

            float fxj = 0.f, fyj = 0.f, fzj = 0.f;
            float r2min = F.r2Off;
#pragma unroll
            for (int k = 0; k < kCluster; k++) {
                const float4 p = myPosq[k];                 // j slot 8 g + (m + k) % 8
                const float4 ab = *reinterpret_cast<const float4 *>(ljRow + myLj[k]);
                const float dx = xi - p.x, dy = yi - p.y, dz = zi - p.z;
                const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                // masked pairs are evaluated AT the outer cutoff, where energy and force vanish (to ~1e-16 kJ/mol): pairs that are not
                // on the list get r2 = NaN (all bits set by the sign-extended mask bit), and fminf(NaN, r2Off) = r2Off takes care of
                // them together with the pairs beyond the cutoff; r2m doubles as the damped-core detector
                PairOut o;
                if (kForm == 0) {
                    const int off = ~(((int) (mask << (31 - k))) >> 31);               // 0 if bit k is set, else 0xffffffff
                    const float r2m = fminf(__int_as_float(__float_as_int(r2) | off), F.r2Off);
                    r2min = fminf(r2min, r2m);
                    o = abfs_pair(F, r2m, qi * p.w, ab.x, ab.y, ab.z);
                } else {
                    // pairs off the list or beyond the cutoff read the all-zero table row (the skip of PairwiseInteraction.c:489)
                    o = spline_pair(splTab, A.splN, A.splInvDR, A.splDR, ((mask >> k) & 1u) && !(r2 > F.r2Off), r2, qi * p.w, ab.x, ab.y);
                }
                eq += o.e1; el += o.e2;
                fxi = fmaf(-o.g, dx, fxi); fyi = fmaf(-o.g, dy, fyi); fzi = fmaf(-o.g, dz, fzi);      // gradient = -(force on i) = -g d
                fxj = fmaf(o.g, dx, fxj); fyj = fmaf(o.g, dy, fyj); fzj = fmaf(o.g, dz, fzj);
                // hand the j-gradient accumulator to the lane that evaluates this j slot next
                fxj = __shfl_sync(0xffffffffu, fxj, src); fyj = __shfl_sync(0xffffffffu, fyj, src); fzj = __shfl_sync(0xffffffffu, fzj, src);
            }
            if (kForm == 0 && __any_sync(0xffffffffu, r2min < F.r2Damp)) {   // damped core: practically never; patch the tile with the reference formulas
                float c[8];
                damped_tile_fix(F, mask, myPosq, ljRow, myLj, xi, yi, zi, qi, src, c);
                fxi += c[0]; fyi += c[1]; fzi += c[2]; fxj += c[3]; fyj += c[4]; fzj += c[5];
                eq += c[6]; el += c[7];
            }
            // after 8 hand-overs inside the group the accumulator of j slot `lane` is back in this lane
            if (has) {
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
            if (((t + 1) % kFlushTiles) == 0 || t + 1 == wi.tileCount) {
                ws->acc[0][lane] += (double) fxi; ws->acc[1][lane] += (double) fyi; ws->acc[2][lane] += (double) fzi;
                ws->acc[3][lane] += (double) eq;  ws->acc[4][lane] += (double) el;
                fxi = 0.f; fyi = 0.f; fzi = 0.f; eq = 0.f; el = 0.f;
            }
        }
        // every lane holds the partial i gradient of its group's slots
        const double fix = ws->acc[0][lane], fiy = ws->acc[1][lane], fiz = ws->acc[2][lane];
        if (ivalid && A.gradSorted != nullptr) {
            double *gp = A.gradSorted + 3 * (size_t) si;
            atomicAdd(gp, sc * fix); atomicAdd(gp + 1, sc * fiy); atomicAdd(gp + 2, sc * fiz);
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
// per energy call: atom records in sorted order from the current coordinates, and the way back for the gradients
// ------------------------------------------------------------------------------------------------------
__global__ void k_pack_records(const double *__restrict__ x, const int *__restrict__ sAtom, int n, const float *__restrict__ q32, const int *__restrict__ ljtype,
                               double ox, double oy, double oz, float4 *__restrict__ recA, float4 *__restrict__ recB, double *__restrict__ zero = nullptr, int zeroCount = 0)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    // fused mode (nbb200_md_run): the accumulators + work cursor of the force kernels are cleared here instead of by a memset of their own
    if (zero != nullptr) for (int i = s; i < zeroCount; i += gridDim.x * blockDim.x) zero[i] = 0.0;
    if (s >= n) return;
    const int a = sAtom[s];
    const double X = x[3 * a] - ox, Y = x[3 * a + 1] - oy, Z = x[3 * a + 2] - oz;
    const double KX = 8.0 * rint(X * 0.125), KY = 8.0 * rint(Y * 0.125), KZ = 8.0 * rint(Z * 0.125);
    recA[s] = make_float4((float) (X - KX), (float) (Y - KY), (float) (Z - KZ), q32[a]);
    recB[s] = make_float4((float) KX, (float) KY, (float) KZ, __int_as_float(ljtype[a]));
}

// assign != 0: the NB term SETS the caller's gradient (every atom has exactly one sorted position) instead of accumulating into it
// kClear (fused mode of nbb200_md_run): the sorted accumulator is left zeroed for the next call (no memset of its own).  Two instantiations:
// the plain one keeps gs const / read-only (the combined one measured 9 x slower on the 1.1 M-atom box: 143 vs 16 us)
template <bool kClear>
__global__ void k_unsort_gradients(typename std::conditional<kClear, double, const double>::type *__restrict__ gs, const int *__restrict__ sAtom, int s0, int n,
                                   double *__restrict__ grad, int assign)
{
    const int s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int a = sAtom[s];
    const double gx = gs[3 * s], gy = gs[3 * s + 1], gz = gs[3 * s + 2];
    if (assign) { grad[3 * a] = gx; grad[3 * a + 1] = gy; grad[3 * a + 2] = gz; }
    else { grad[3 * a] += gx; grad[3 * a + 1] += gy; grad[3 * a + 2] += gz; }
    if constexpr (kClear) { gs[3 * s] = 0.0; gs[3 * s + 1] = 0.0; gs[3 * s + 2] = 0.0; }
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
    {"128x5", k_cluster_forces<false, 128, 5, 0>, 128},      // default: 96 registers, 20 warps per SM (measured best on B200)
    {"256x3", k_cluster_forces<false, 256, 3, 0>, 256}, {"256x2", k_cluster_forces<false, 256, 2, 0>, 256},
    {"128x4", k_cluster_forces<false, 128, 4, 0>, 128}, {"128x6", k_cluster_forces<false, 128, 6, 0>, 128},
};
static const ForceVariant kForceRot = {"rot", k_cluster_forces<true, 256, 2, 0>, 256};
// spline form: [rotations][tables in shared memory (1) / global memory (0)]
static const ForceVariant kForceSpline[2][2] = {
    {{"spline-global", k_cluster_forces<false, 256, 2, 2>, 256}, {"spline", k_cluster_forces<false, 256, 2, 1>, 256}},
    {{"spline-rot-global", k_cluster_forces<true, 256, 2, 2>, 256}, {"spline-rot", k_cluster_forces<true, 256, 2, 1>, 256}},
};
constexpr size_t kSplineSmemLimit = 96 * 1024;        // per CTA; two CTAs per SM stay resident

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
    if (clear) k_unsort_gradients<true><<<blocks, threads, 0, s.stream>>>(s.gs, s.sAtom.p, (int) s0, (int) s1, d_grad, assign ? 1 : 0);
    else k_unsort_gradients<false><<<blocks, threads, 0, s.stream>>>(s.gs, s.sAtom.p, (int) s0, (int) s1, d_grad, assign ? 1 : 0);
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
    const bool fusedZero = s.mdFused && nitems > 0;           // k_pack_records clears the accumulators and the cursor in fused mode
    if (!fusedZero) NBB_CUDA(cudaMemsetAsync(s.accum.p, 0, sizeof(double) * (accumCount + 1), s.stream));
    unsigned int *workCursor = reinterpret_cast<unsigned int *>(s.accum.p + accumCount);
    const double eScale = (1.0 / s.dielectric) * kE2AngstromToKJMol;
    if (nitems > 0) {
        if (!s.recA.ensure((size_t) s.n) || !s.recB.ensure((size_t) s.n)) return false;
        const int pthreads = 256, pblocks = (s.n + pthreads - 1) / pthreads;
        k_pack_records<<<pblocks, pthreads, 0, s.stream>>>(s.xcur, s.sAtom.p, s.n, s.q32.p, s.ljtype.p, s.grid.lo[0], s.grid.lo[1], s.grid.lo[2], s.recA.p, s.recB.p,
                                                            fusedZero ? s.accum.p : nullptr, (int) (accumCount + 1));
        ForceArgs A;
        A.items = s.items.p; A.nitems = nitems; A.workCursor = workCursor;
        A.tileDesc = s.tileDesc.p; A.recA = s.recA.p; A.recB = s.recB.p; A.n = s.n;
        A.ljAB = s.ljAB.p; A.ntypes = s.ntypes;
        A.ops = s.imageOps.p;
        for (int d = 0; d < 3; d++) A.origin[d] = s.grid.lo[d];
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
        if (g_numSMs == 0) init_force_kernel_attributes();
        bool rot = false;                                   // any image with a genuine rotation?
        for (const RealSpaceOp &b : s.plan.baseOps) rot = rot || !b.pureTranslation;
        static const ForceVariant *chosen = []() {
            const char *e = std::getenv("NBB200_FORCE_SHAPE");
            for (const ForceVariant &v : kForceVariants) if (e != nullptr && std::strcmp(e, v.name) == 0) return &v;
            return &kForceVariants[0];
        }();
        const bool splSmem = spline && splBytes <= kSplineSmemLimit;
        const ForceVariant &v = spline ? kForceSpline[rot ? 1 : 0][splSmem ? 1 : 0] : (rot ? kForceRot : *chosen);
        const int warpsPerBlock = v.threads / 32;
        const size_t smem = sizeof(WarpScratch) * warpsPerBlock + sizeof(float4) * (size_t) s.ntypes * s.ntypes + (splSmem ? splBytes : 0);
        if (smem > 200 * 1024) { set_error("too many LJ types for the shared-memory table"); return false; }
        int perSM = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, v.fn, v.threads, smem);
        if (perSM < 1) perSM = 1;
        const int grid = std::max(1, std::min(g_numSMs * perSM, (nitems + warpsPerBlock - 1) / warpsPerBlock));
        if (s.timing) cudaEventRecord(s.ev[2], s.stream);
        v.fn<<<grid, v.threads, smem, s.stream>>>(A);
        if (s.timing) cudaEventRecord(s.ev[3], s.stream);
        s.launches += 2;
    }
    if (s.n14 > 0) {
        F64Factors FF;
        std::memcpy(FF.v, s.factors, sizeof(FF.v));
        const int threads = 128, nblk = std::max(1, std::min(1184, (s.n14 + threads - 1) / threads));
        if (s.timing) cudaEventRecord(s.ev[4], s.stream);
        k_pairs14<<<nblk, threads, 0, s.stream>>>(s.pairs14.p, s.n14, s.xcur, s.q64.p, s.ljtype.p, s.ljAB14.p, s.ntypes14, FF,
                                                    (s.useAnalytic ? eScale : 1.0 / s.dielectric) * s.scale14, s.invPerm.p, s.ownLo, s.ownHi, wantGrad ? s.gs : nullptr,
                                                    s.accum.p + 16 * s.nsets, s.useAnalytic ? nullptr : s.splF64.p, s.useAnalytic ? 0 : s.spl.points());
        if (s.timing) cudaEventRecord(s.ev[5], s.stream);
        s.launches += 1;
    }
    // device-array calls honour nbb200_set_gradient_overwrite too (the host-array call handles it with its own staging buffer: d_grad = s.grad.p)
    const bool clearGs = s.mdFused && s.gsExternal == nullptr && s.nranks == 1 && d_grad != nullptr;
    if (d_grad != nullptr && !unsort_gradients(s, 0, s.n, d_grad, s.gradOverwrite && d_grad != s.grad.p && s.nranks == 1, clearGs)) return false;
    if (clearGs) s.gsZeroed = true;                          // the whole accumulator (3 n) has just been cleared by the unsort pass
    return cuda_ok(cudaGetLastError(), "force kernels");
}

}  // namespace nbb200
