// list_build.cu -- device side of the pair-list rebuild (sm_100a).
//
// Replaces, for the pure-MM case, what NBModelABFS_Update drives on the CPU in the reference
// (pM = pMolecule-1.9.0/extensions, pC = pCore-1.9.0/extensions):
//   GenerateLists -> PairListGenerator_SelfPairListFromCoordinates3        pM/csource/NBModelABFS.c:1050-1108, pC/csource/PairListGenerator.c:530-553,946-1099
//   GenerateImageLists -> PairListGenerator_CrossPairListFromDoubleCoordinates3   pM/csource/NBModelABFS.c:753-1045, pC/csource/PairListGenerator.c:414-446,703-791
//   Coordinates3_MakeGridAndOccupancy / RegularGridOccupancy_Fill (counting sort) pC/csource/Coordinates3.c:1164, pC/csource/RegularGridOccupancy.c:77-147
//   CheckForUpdate                                                         pM/csource/NBModelABFS.c:691-746
//
// Output format (B200-first, not the reference's linked lists): atoms are counting-sorted into grid cells,
// consecutive runs of 32 sorted atoms form i-blocks, and for every i-block the builder emits TILES: 32 j atoms
// (individually selected: each has at least one list pair with the block) plus a 32x32 bit mask.  Mask bit (i, j)
// is set iff the reference would put the pair on its list:
//     (dx*dx + dy*dy) + dz*dz <= listCutoff^2   in fp64 WITHOUT fused multiply-add        (PairListGenerator.c:70-81,131-139)
//     and, for the primary list, i != j counted once and (i, j) not excluded            (PairListGenerator.c:100-113,1063-1085)
// with dx = x_i - x'_j and x'_j the image coordinate produced by the reference's own operation sequence
// (rotate, translate, then the +d / -d displacement walk of GenerateImageLists, :953-958,1003-1009).
// The explicit (i, j) lists the reference holds are recovered from the masks by expand_pairs() below.
//
// This file is compiled with -fmad=false; the distance predicate additionally uses __dmul_rn/__dadd_rn.
#include "nbb200_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace nbb200 {

// ------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ref_dist2(double dx, double dy, double dz)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// x' = R x + tv exactly as Coordinates3_Rotate + Coordinates3_Translate do it (pC/csource/Coordinates3.c:1473-1500,1772-1800)
__device__ __forceinline__ void ref_transform(const double *op, double x0, double y0, double z0, double &x1, double &y1, double &z1)
{
    x1 = __dadd_rn(__dadd_rn(__dmul_rn(op[0], x0), __dmul_rn(op[1], y0)), __dmul_rn(op[2], z0));
    y1 = __dadd_rn(__dadd_rn(__dmul_rn(op[3], x0), __dmul_rn(op[4], y0)), __dmul_rn(op[5], z0));
    z1 = __dadd_rn(__dadd_rn(__dmul_rn(op[6], x0), __dmul_rn(op[7], y0)), __dmul_rn(op[8], z0));
    x1 = __dadd_rn(x1, op[9]); y1 = __dadd_rn(y1, op[10]); z1 = __dadd_rn(z1, op[11]);
}

// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): two distance tests per instruction in the prefilter
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// Boustrophedon ("snake") cell order: columns (cx, cy) are walked with y reversed on odd cx and z reversed on odd columns, so
// that consecutive cells in the sorted order are always spatial neighbours.  32-atom i-blocks are runs of the sorted order;
// without the snake a block that straddles the end of a z column would span the whole box (huge bounding box, thousands of
// useless candidates in the tile builder, half-empty tiles in the force kernel).
__device__ __forceinline__ int snake_column(const BuildGrid &g, int cx, int cy) { return cx * g.dim[1] + ((cx & 1) ? g.dim[1] - 1 - cy : cy); }
__device__ __forceinline__ bool snake_reversed(int column) { return (column & 1) != 0; }
__device__ __forceinline__ int snake_cell(const BuildGrid &g, int cx, int cy, int cz)
{
    const int col = snake_column(g, cx, cy);
    return col * g.dim[2] + (snake_reversed(col) ? g.dim[2] - 1 - cz : cz);
}

__device__ __forceinline__ int cell_coord(double v, double lo, double invh, int dim)
{
    int c = (int) floor((v - lo) * invh);
    return min(dim - 1, max(0, c));
}

// ------------------------------------------------------------------------------------------------------
// bounding boxes of the primary coordinates (op 0) and of the coordinates transformed by each base operation
// (Coordinates3_EnclosingOrthorhombicBox, pC/csource/Coordinates3.c:498-...): partial min/max per CTA, then one CTA.
// ------------------------------------------------------------------------------------------------------
__global__ void k_bbox(const double *__restrict__ x, int n, const double *__restrict__ ops, int nops, double *__restrict__ partial, unsigned int *ticket,
                       double *__restrict__ out, double *__restrict__ hostOut)
{
    // partial[(blockIdx.x * (nops) + o) * 6 + {min xyz, max xyz}]
    extern __shared__ double sh[];
    for (int o = 0; o < nops; o++) {
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            double p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
            if (o > 0) ref_transform(ops + 12 * (o - 1), p[0], p[1], p[2], p[0], p[1], p[2]);
            for (int d = 0; d < 3; d++) { mn[d] = fmin(mn[d], p[d]); mx[d] = fmax(mx[d], p[d]); }
        }
        for (int d = 0; d < 3; d++) {
            for (int off = 16; off > 0; off >>= 1) {
                mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], off));
                mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], off));
            }
        }
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if ((threadIdx.x & 31) == 0) for (int d = 0; d < 3; d++) { sh[w * 6 + d] = mn[d]; sh[w * 6 + 3 + d] = mx[d]; }
        __syncthreads();
        if (threadIdx.x < 6) {
            double v = sh[threadIdx.x];
            for (int k = 1; k < nw; k++) v = (threadIdx.x < 3) ? fmin(v, sh[k * 6 + threadIdx.x]) : fmax(v, sh[k * 6 + threadIdx.x]);
            partial[((size_t) blockIdx.x * nops + o) * 6 + threadIdx.x] = v;
        }
        __syncthreads();
    }
    // the last CTA to finish reduces the partial boxes and stores the result into page-locked host memory as well: one launch, no copy operation
    if (!last_block_done(ticket)) return;
    const int lane = threadIdx.x & 31, nblk = (int) gridDim.x;
    for (int idx = threadIdx.x >> 5; idx < nops * 6; idx += blockDim.x >> 5) {        // idx = o * 6 + c, one warp each
        const int c = idx % 6;
        double v = (c < 3) ? 1e300 : -1e300;
        for (int k = lane; k < nblk; k += 32) { const double u = __ldcg(partial + (size_t) k * nops * 6 + idx); v = (c < 3) ? fmin(v, u) : fmax(v, u); }
        for (int off = 16; off > 0; off >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, v, off); v = (c < 3) ? fmin(v, u) : fmax(v, u); }
        if (lane == 0) { out[idx] = v; hostOut[idx] = v; }
    }
}

bool device_bbox(State &s, int nops, double *hostMin, double *hostExt)
{
    const int threads = 256;
    int nblk = std::min(512, (s.n + threads * 4 - 1) / (threads * 4));
    if (nblk < 1) nblk = 1;
    if (!s.bboxDev.ensure((size_t) (nblk + 1) * nops * 6)) return false;
    if ((size_t) nops * 6 > kSmallDoubles - 16) { set_error("too many transformations for the result buffer"); return false; }
    if (s.ticket.p == nullptr) {
        if (!s.ticket.ensure(4)) return false;
        NBB_CUDA(cudaMemsetAsync(s.ticket.p, 0, sizeof(unsigned int) * 4, s.stream));
    }
    double *partial = s.bboxDev.p + (size_t) nops * 6;
    k_bbox<<<nblk, threads, (threads / 32) * 6 * sizeof(double), s.stream>>>(s.xcur, s.n, s.baseOpsDev.p, nops, partial, s.ticket.p, s.bboxDev.p, s.hsmall);
    s.launches += 1;
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    for (int o = 0; o < nops; o++)
        for (int d = 0; d < 3; d++) { hostMin[3 * o + d] = s.hsmall[6 * o + d]; hostExt[3 * o + d] = s.hsmall[6 * o + 3 + d] - s.hsmall[6 * o + d]; }
    return true;
}

// ------------------------------------------------------------------------------------------------------
// CheckForUpdate (pM/csource/NBModelABFS.c:691-746): max |x - x_ref|^2 and "any atom beyond buffac".
// The reference's early exit only matters when an update happens (then the maximum is not used further).
// ------------------------------------------------------------------------------------------------------
__global__ void k_displacement(const double *__restrict__ x, const double *__restrict__ xref, int n, const unsigned char *__restrict__ fixed, double buffacsq,
                               unsigned long long *__restrict__ out, unsigned long long *__restrict__ zeroOther = nullptr)
{
    if (zeroOther != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zeroOther = 0ULL;     // two-slot use (nbb200_md_run): prepares the next step's slot
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (fixed != nullptr && fixed[i]) continue;              // NBModelABFS.c:723-739: fixed atoms do not trigger updates
        const double dx = x[3 * i] - xref[3 * i], dy = x[3 * i + 1] - xref[3 * i + 1], dz = x[3 * i + 2] - xref[3 * i + 2];
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        m = fmax(m, r2);
    }
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long) __double_as_longlong(m));   // r2 >= 0: bit pattern is monotone
}

bool displacement_check(State &s, double buffacsq, double *maxr2, int *exceeded)
{
    if (!s.bboxDev.ensure(64)) return false;
    unsigned long long *out = reinterpret_cast<unsigned long long *>(s.bboxDev.p);
    NBB_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), s.stream));
    const int threads = 256;
    const int nblk = std::max(1, std::min(1184, (s.n + threads - 1) / threads));
    if (s.timing) cudaEventRecord(s.ev[6], s.stream);
    k_displacement<<<nblk, threads, 0, s.stream>>>(s.xcur, s.xref.p, s.n, s.nfixed > 0 ? s.fixedFlag.p : nullptr, buffacsq, out);
    if (s.timing) cudaEventRecord(s.ev[7], s.stream);
    s.launches += 1;
    NBB_CUDA(cudaMemcpyAsync(s.hsmall, out, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    *maxr2 = s.hsmall[0];
    *exceeded = (*maxr2 > buffacsq) ? 1 : 0;
    return true;
}

bool displacement_enqueue(State &s, const double *d_x, double *d_out, double *d_zeroOther)
{
    if (d_zeroOther == nullptr) NBB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), s.stream));       // else: d_out was zeroed by the previous two-slot call
    const int threads = 256;
    const int nblk = std::max(1, std::min(1184, (s.n + threads - 1) / threads));
    k_displacement<<<nblk, threads, 0, s.stream>>>(d_x, s.xref.p, s.n, s.nfixed > 0 ? s.fixedFlag.p : nullptr, 0.0, reinterpret_cast<unsigned long long *>(d_out),
                                                  reinterpret_cast<unsigned long long *>(d_zeroOther));
    s.launches += 1;
    return cuda_ok(cudaGetLastError(), "k_displacement");
}

// ------------------------------------------------------------------------------------------------------
// useCentering: NBModelABFSState_InitializeCoordinates3 (pM/csource/NBModelABFSState.c:278-311).  On a list update every isolate
// (molecule) is translated by whole lattice vectors so that its centre lies in the primary cell
// (SymmetryParameters_CenterCoordinates3ByIsolate, pM/csource/SymmetryParameters.c:82-129; Coordinates3_Center,
// pC/csource/Coordinates3.c:357-440; _FindCenteringTranslation :271-290), with the reference's operation order: the lists are
// built from these coordinates bit for bit.  Between updates the translations of the last update are re-applied.
// ------------------------------------------------------------------------------------------------------
struct CentreArgs { double M[9], invM[9]; };

__global__ void k_centre_isolates(const double *__restrict__ xin, int n, const int *__restrict__ isoPtr, const int *__restrict__ isoIdx, int nisolates, CentreArgs C,
                                  double *__restrict__ xc, double *__restrict__ isoT)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nisolates) return;
    const int lo = isoPtr[k], hi = isoPtr[k + 1];
    double c[3] = {0.0, 0.0, 0.0};
    for (int i = lo; i < hi; i++) {                          // ascending atom index, one running sum per component
        const int a = isoIdx[i];
        c[0] = __dadd_rn(c[0], xin[3 * a]); c[1] = __dadd_rn(c[1], xin[3 * a + 1]); c[2] = __dadd_rn(c[2], xin[3 * a + 2]);
    }
    const double scale = __ddiv_rn(1.0, (double) (hi - lo));
    for (int d = 0; d < 3; d++) c[d] = __dmul_rn(c[d], scale);
    double t[3], nn[3];
    for (int d = 0; d < 3; d++) {                            // Matrix33_ApplyToVector3(inverseM, centre), then -floor
        const double f = __dadd_rn(__dadd_rn(__dmul_rn(c[0], C.invM[3 * d]), __dmul_rn(c[1], C.invM[3 * d + 1])), __dmul_rn(c[2], C.invM[3 * d + 2]));
        nn[d] = (double) (-(int) floor(f));
    }
    for (int d = 0; d < 3; d++)                               // SymmetryParameters_Displacement
        t[d] = __dadd_rn(__dadd_rn(__dmul_rn(nn[0], C.M[3 * d]), __dmul_rn(nn[1], C.M[3 * d + 1])), __dmul_rn(nn[2], C.M[3 * d + 2]));
    for (int i = lo; i < hi; i++) {
        const int a = isoIdx[i];
        for (int d = 0; d < 3; d++) {
            const double v = __dadd_rn(xin[3 * a + d], t[d]);
            xc[3 * a + d] = v;
            isoT[3 * a + d] = __dadd_rn(v, __dmul_rn(-1.0, xin[3 * a + d]));      // isolateTranslations3 = centred - input
        }
    }
}

__global__ void k_apply_translations(const double *__restrict__ xin, const double *__restrict__ isoT, long m, double *__restrict__ xc)
{
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) xc[i] = __dadd_rn(xin[i], isoT[i]);
}

bool centre_coordinates(State &s, const double *d_xin, bool doUpdate)
{
    const long m = 3 * (long) s.n;
    if (!s.xc.ensure((size_t) m) || !s.isoT.ensure((size_t) m)) return false;
    if (doUpdate) {
        // atoms of removed isolates (fixed atoms) keep their coordinates and a zero translation
        NBB_CUDA(cudaMemcpyAsync(s.xc.p, d_xin, sizeof(double) * m, cudaMemcpyDeviceToDevice, s.stream));
        NBB_CUDA(cudaMemsetAsync(s.isoT.p, 0, sizeof(double) * m, s.stream));
        CentreArgs C;
        std::memcpy(C.M, s.lattice.M.v, sizeof(C.M)); std::memcpy(C.invM, s.lattice.invM.v, sizeof(C.invM));
        k_centre_isolates<<<(s.nisolates + 127) / 128, 128, 0, s.stream>>>(d_xin, s.n, s.isoPtr.p, s.isoIdx.p, s.nisolates, C, s.xc.p, s.isoT.p);
    } else k_apply_translations<<<(unsigned int) ((m + 255) / 256), 256, 0, s.stream>>>(d_xin, s.isoT.p, m, s.xc.p);
    s.launches += 1;
    return cuda_ok(cudaGetLastError(), "centring");
}

// ------------------------------------------------------------------------------------------------------
// extended atoms: set 0 = the n primary atoms (entry e = atom index), sets 1.. = image atoms that fall inside
// the search box (primary bounding box dilated by the cutoff).  Image coordinates follow the reference walk.
// ------------------------------------------------------------------------------------------------------
struct ExtendArgs {
    const double *x; int n;
    const double *baseOps;             // 12 doubles per transformation
    const double *visitDisp; const int *visitInfo; int nvisits;     // visitInfo: (t, image) per visit
    double boxLo[3], boxHi[3];         // search box with a small safety margin
    BuildGrid grid;
    double *eX; int *eAtom; int *eSet; int *eKey; unsigned long long *eSort;
    unsigned int *cellCount; unsigned int extCap; DeviceCounters *counters;
};

__device__ __forceinline__ void classify(const BuildGrid &g, int set, int atom, double px, double py, double pz, int &key, unsigned long long &sortKey)
{
    const int cx = cell_coord(px, g.lo[0], g.invh, g.dim[0]), cy = cell_coord(py, g.lo[1], g.invh, g.dim[1]), cz = cell_coord(pz, g.lo[2], g.invh, g.dim[2]);
    key = set * g.ncell + snake_cell(g, cx, cy, cz);
    const bool rev = snake_reversed(snake_column(g, cx, cy));
    // position inside the cell: 8 z slabs, then 4 y rows, then 4 x columns -> runs of sorted atoms are compact boxes
    const double fx = (px - g.lo[0]) * g.invh - cx, fy = (py - g.lo[1]) * g.invh - cy, fz = (pz - g.lo[2]) * g.invh - cz;
    const int sx = min(3, max(0, (int) (fx * 4.0))), sy = min(3, max(0, (int) (fy * 4.0))), sz = min(7, max(0, (int) (fz * 8.0)));
    sortKey = ((unsigned long long) ((rev ? 7 - sz : sz) * 16 + sy * 4 + sx) << 32) | (unsigned int) atom;
}

__global__ void k_extend(ExtendArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.n) return;
    const double x0 = A.x[3 * j], y0 = A.x[3 * j + 1], z0 = A.x[3 * j + 2];
    int key; unsigned long long sk;
    classify(A.grid, 0, j, x0, y0, z0, key, sk);
    A.eX[3 * j] = x0; A.eX[3 * j + 1] = y0; A.eX[3 * j + 2] = z0;
    A.eAtom[j] = j; A.eSet[j] = 0; A.eKey[j] = key; A.eSort[j] = sk;
    atomicAdd(&A.cellCount[key], 1u);
    int tcur = -1;
    double ux = 0, uy = 0, uz = 0;
    for (int v = 0; v < A.nvisits; v++) {
        const int t = A.visitInfo[2 * v], image = A.visitInfo[2 * v + 1];
        if (t != tcur) { ref_transform(A.baseOps + 12 * t, x0, y0, z0, ux, uy, uz); tcur = t; }
        const double dx = A.visitDisp[3 * v], dy = A.visitDisp[3 * v + 1], dz = A.visitDisp[3 * v + 2];
        const double wx = __dadd_rn(ux, dx), wy = __dadd_rn(uy, dy), wz = __dadd_rn(uz, dz);
        if (image >= 0 && wx >= A.boxLo[0] && wx <= A.boxHi[0] && wy >= A.boxLo[1] && wy <= A.boxHi[1] && wz >= A.boxLo[2] && wz <= A.boxHi[2]) {
            const unsigned int slot = atomicAdd(&A.counters->extCount, 1u);
            if (slot < A.extCap) {
                const size_t e = (size_t) A.n + slot;
                classify(A.grid, 1 + image, j, wx, wy, wz, key, sk);
                A.eX[3 * e] = wx; A.eX[3 * e + 1] = wy; A.eX[3 * e + 2] = wz;
                A.eAtom[e] = j; A.eSet[e] = 1 + image; A.eKey[e] = key; A.eSort[e] = sk;
                atomicAdd(&A.cellCount[key], 1u);
            } else atomicOr(&A.counters->overflow, 1u);
        }
        ux = __dadd_rn(wx, __dmul_rn(dx, -1.0)); uy = __dadd_rn(wy, __dmul_rn(dy, -1.0)); uz = __dadd_rn(wz, __dmul_rn(dz, -1.0));
    }
}

// stand-alone cross list front end: set 1 = a second coordinate array, taken as given (points outside the
// search box cannot pair with anything and are dropped)
__global__ void k_extend_cross(const double *__restrict__ x2, int n2, int n1, BuildGrid grid, double3 boxLo, double3 boxHi, double *eX, int *eAtom, int *eSet,
                               int *eKey, unsigned long long *eSort, unsigned int *cellCount, DeviceCounters *counters)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    const double px = x2[3 * j], py = x2[3 * j + 1], pz = x2[3 * j + 2];
    if (!(px >= boxLo.x && px <= boxHi.x && py >= boxLo.y && py <= boxHi.y && pz >= boxLo.z && pz <= boxHi.z)) return;
    const size_t e = (size_t) n1 + atomicAdd(&counters->extCount, 1u);
    int key; unsigned long long sk;
    classify(grid, 1, j, px, py, pz, key, sk);
    eX[3 * e] = px; eX[3 * e + 1] = py; eX[3 * e + 2] = pz;
    eAtom[e] = j; eSet[e] = 1; eKey[e] = key; eSort[e] = sk;
    atomicAdd(&cellCount[key], 1u);
}

// ------------------------------------------------------------------------------------------------------
// exclusive scan of the cell counts (counting sort, cf. RegularGridOccupancy_Fill pC/csource/RegularGridOccupancy.c:94-126)
// ------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024, kScanPer = 4, kScanChunk = kScanThreads * kScanPer;

__device__ unsigned int block_exclusive_scan(unsigned int v, unsigned int *total)
{
    __shared__ unsigned int wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int inc = v;
    for (int off = 1; off < 32; off <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += u; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned int s = (lane < (int) (blockDim.x >> 5)) ? wsum[lane] : 0u, si = s;
        for (int off = 1; off < 32; off <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, si, off); if (lane >= off) si += u; }
        wsum[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    const unsigned int r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void k_scan_chunks(const unsigned int *__restrict__ in, unsigned int *__restrict__ out, unsigned int *__restrict__ sums, int n)
{
    __shared__ unsigned int total;
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanPer;
    unsigned int v[kScanPer], t = 0;
    for (int k = 0; k < kScanPer; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; t += v[k]; }
    unsigned int ex = block_exclusive_scan(t, &total);
    for (int k = 0; k < kScanPer; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void k_scan_sums(unsigned int *sums, int nchunks, unsigned int *grand)
{
    __shared__ unsigned int total;
    unsigned int carry = 0;
    for (int base = 0; base < nchunks; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const unsigned int v = (i < nchunks) ? sums[i] : 0u;
        const unsigned int ex = block_exclusive_scan(v, &total);
        if (i < nchunks) sums[i] = ex + carry;
        __syncthreads();
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry;
}

__global__ void k_scan_add(unsigned int *out, const unsigned int *__restrict__ sums, int n, const unsigned int *grand)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += sums[i / kScanChunk];
    if (i == n) out[n] = *grand;
}

// ---- several ranks (spatial decomposition of the sort): the global cell histogram / prefix sum is computed by everybody (the sorted
// positions are global), but a rank scatters, sorts, groups and packs only the cells its slab can see:
//   need[c]         (image entries)  cell c lies within 3 cells (>= listCutoff: the cell edge is listCutoff / 2) of a cell that holds an owned atom
//   need[ncell + c] (primary atoms)  the same, or one of the cell's atoms has an image copy in a wanted cell (the force kernel reads the
//                                    PRIMARY record of an image atom and sends its gradient to the primary atom's owner)
__device__ __forceinline__ void snake_decode(const BuildGrid &g, int idx, int &cx, int &cy, int &cz)
{
    const int col = idx / g.dim[2], zz = idx - col * g.dim[2];
    cz = snake_reversed(col) ? g.dim[2] - 1 - zz : zz;
    cx = col / g.dim[1];
    const int yy = col - cx * g.dim[1];
    cy = (cx & 1) ? g.dim[1] - 1 - yy : yy;
}

__global__ void k_mark_needed(const unsigned int *__restrict__ cellStart, BuildGrid g, unsigned int ownLo, unsigned int ownHi, unsigned char *__restrict__ need)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncell) return;
    int cx, cy, cz;
    snake_decode(g, c, cx, cy, cz);
    bool wanted = false;
    for (int ax = max(0, cx - 3); ax <= min(g.dim[0] - 1, cx + 3) && !wanted; ax++)
        for (int ay = max(0, cy - 3); ay <= min(g.dim[1] - 1, cy + 3) && !wanted; ay++) {
            // the z run of a column is contiguous in the snake order: one range test
            const int k0 = snake_cell(g, ax, ay, max(0, cz - 3)), k1 = snake_cell(g, ax, ay, min(g.dim[2] - 1, cz + 3));
            const unsigned int lo = cellStart[min(k0, k1)], hi = cellStart[max(k0, k1) + 1];
            wanted = hi > lo && lo < ownHi && hi > ownLo;
        }
    need[c] = wanted ? 1 : 0;
    need[g.ncell + c] = wanted ? 1 : 0;
}

__global__ void k_mark_image_sources(const int *__restrict__ eKey, const int *__restrict__ eSet, const int *__restrict__ eAtom, int n, unsigned int extCap,
                                     const DeviceCounters *__restrict__ counters, int ncell, unsigned char *need)
{
    const int e = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n + (int) min(counters->extCount, extCap)) return;
    if (need[eKey[e] - eSet[e] * ncell]) need[ncell + eKey[eAtom[e]]] = 1;          // entry of a primary atom = its atom index
}

__global__ void k_scatter(const int *__restrict__ eKey, const int *__restrict__ eSet, int n, unsigned int extCap, const DeviceCounters *__restrict__ counters,
                          const unsigned int *__restrict__ cellStart, unsigned int *cellFill, int *order, const unsigned char *__restrict__ need, int ncell)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int ne = n + (int) min(counters->extCount, extCap);      // read on the device: no host round trip between the kernels of a rebuild
    if (e >= ne) return;
    const int key = eKey[e];
    if (need != nullptr) {
        const int set = eSet[e];
        if (!need[set == 0 ? ncell + key : key - set * ncell]) return;
    }
    order[cellStart[key] + atomicAdd(&cellFill[key], 1u)] = e;
}

// deterministic order inside each cell: rank sort by (sub-cell key, atom index), one warp per cell; the sorted arrays (coordinates,
// atom index, inverse permutation of the primary atoms) are written in the same pass.  (One warp per 32 keys of the mostly empty image
// sets, walking the ones with work, was slower: 140 instead of 89 us on the 1.1 M-atom box; so were, in round 2, that walk for the image
// sets only (rebuild 1.48 -> 1.51 ms) and a persistent grid that claims groups of 32 keys from a cursor (1.52 ms; DHFR 215 -> 250 us):
// the rank sort of an occupied cell is a serial chain, and a warp that owns several occupied cells runs them one after the other.  CTAs
// of 1024 threads: 107 us.)
__global__ void k_sort_cells(const unsigned int *__restrict__ cellStart, int nkeys, const unsigned long long *__restrict__ eSort, const int *__restrict__ order,
                             const double *__restrict__ eX, const int *__restrict__ eAtom, const int *__restrict__ eSet,
                             double *__restrict__ sX, int *__restrict__ sAtom, int *__restrict__ invPerm, const unsigned char *__restrict__ need, int ncell)
{
    const int key = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (key >= nkeys) return;
    const int lo = (int) cellStart[key], m = (int) cellStart[key + 1] - lo;
    if (m <= 0) return;
    if (need != nullptr && !need[key < ncell ? ncell + key : key % ncell]) {
        // a cell this rank does not see: its primary positions are marked (no record is packed for them), image positions stay unwritten
        if (key < ncell) for (int k = lane; k < m; k += 32) sAtom[lo + k] = -1;
        return;
    }
    for (int base = 0; base < m; base += 32) {
        const bool mine = base + lane < m;
        const int e = mine ? order[lo + base + lane] : 0;
        int rank = base + lane;                              // pathological density (> 2048 per cell): keep the (valid, but arbitrary) scatter order
        if (m <= 2048) {
            const unsigned long long k = mine ? eSort[e] : 0ULL;
            rank = 0;
            for (int b2 = 0; b2 < m; b2 += 32) {
                const unsigned long long ko = (b2 + lane < m) ? eSort[order[lo + b2 + lane]] : ~0ULL;
                const int cnt = min(32, m - b2);
                for (int t = 0; t < cnt; t++) rank += (__shfl_sync(0xffffffffu, ko, t) < k) ? 1 : 0;
            }
        }
        if (mine) {
            const int p = lo + rank;
            sX[3 * p] = eX[3 * e]; sX[3 * p + 1] = eX[3 * e + 1]; sX[3 * p + 2] = eX[3 * e + 2];
            const int a = eAtom[e];
            sAtom[p] = a;
            if (eSet[e] == 0) invPerm[a] = p;
        }
    }
}

// Inside every aligned run of 8 sorted primary atoms (= one i-cluster of the force kernel) the atoms whose type has no Lennard-Jones
// interaction at all (TIP3P hydrogens ...) are moved to the front, stably: the force kernel evaluates the cluster atoms two at a time and
// skips the Lennard-Jones part of a pair of atoms when both are of such a type.  Atoms never leave their grid cell (the cell ranges of
// the sorted order are what the tile builder scans): a run that straddles a cell boundary is partitioned piece by piece.
__global__ void k_group_lj_free(int n, const int *__restrict__ ljtype, const unsigned char *__restrict__ typeFree, BuildGrid g, double *__restrict__ sX, int *__restrict__ sAtom,
                                int *__restrict__ invPerm, int firstBlock, int nblocks, double *__restrict__ blockBox)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;            // sorted position; 8 consecutive lanes = one cluster
    const int lane = threadIdx.x & 31, base = lane & ~(kCluster - 1), rel = lane - base;
    const int a = (s < n) ? sAtom[s] : -1;
    const bool valid = a >= 0;                                      // restricted sort: positions of cells this rank does not see hold -1
    double x = 0.0, y = 0.0, z = 0.0;
    if (valid) { x = sX[3 * s]; y = sX[3 * s + 1]; z = sX[3 * s + 2]; }
    const int key = valid ? snake_cell(g, cell_coord(x, g.lo[0], g.invh, g.dim[0]), cell_coord(y, g.lo[1], g.invh, g.dim[1]), cell_coord(z, g.lo[2], g.invh, g.dim[2])) : -1 - lane;
    const bool isFree = valid && typeFree[ljtype[a]] != 0;
    // lanes of the same cluster that sit in the same cell
    unsigned int piece = 0u;
#pragma unroll
    for (int k = 0; k < kCluster; k++) if (__shfl_sync(0xffffffffu, key, base + k) == key) piece |= 1u << k;
    const unsigned int balFree = (__ballot_sync(0xffffffffu, isFree) >> base) & piece;
    const unsigned int below = (1u << rel) - 1u;
    // stable partition of the piece: free atoms first, then the others
    const int first = __ffs((int) piece) - 1;
    const int dst = first + (isFree ? __popc(balFree & below) : __popc(balFree) + __popc(piece & ~balFree & below));
    __syncwarp();
    if (valid) {
        const int p = (s & ~(kCluster - 1)) + dst;
        sX[3 * p] = x; sX[3 * p + 1] = y; sX[3 * p + 2] = z;
        sAtom[p] = a; invPerm[a] = p;
    }
    // the box of the sort block (this warp's 32 atoms: the grouping leaves the set unchanged) -- k_block_boxes folded in
    const int b = s >> 5;
    if (blockBox != nullptr && b >= firstBlock && b < firstBlock + nblocks) {
        double v[6] = {s < n ? x : 1e300, s < n ? y : 1e300, s < n ? z : 1e300, s < n ? x : -1e300, s < n ? y : -1e300, s < n ? z : -1e300};
#pragma unroll
        for (int d = 0; d < 3; d++)
            for (int off = 16; off > 0; off >>= 1) { v[d] = fmin(v[d], __shfl_xor_sync(0xffffffffu, v[d], off)); v[3 + d] = fmax(v[3 + d], __shfl_xor_sync(0xffffffffu, v[3 + d], off)); }
        if (lane < 3) { blockBox[9 * b + lane] = v[lane]; blockBox[9 * b + 3 + lane] = v[3 + lane]; blockBox[9 * b + 6 + lane] = 0.5 * (v[lane] + v[3 + lane]); }
    }
}

// per i-block (32 consecutive sorted primary atoms): min, max, centre
__global__ void k_block_boxes(const double *__restrict__ sX, int n, int firstBlock, int nblocks, double *__restrict__ blockBox)
{
    const int b = firstBlock + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= firstBlock + nblocks) return;
    const int s = b * kTile + lane;
    double mn[3], mx[3];
    for (int d = 0; d < 3; d++) { const double v = (s < n) ? sX[3 * s + d] : 0.0; mn[d] = (s < n) ? v : 1e300; mx[d] = (s < n) ? v : -1e300; }
    for (int d = 0; d < 3; d++)
        for (int off = 16; off > 0; off >>= 1) { mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], off)); mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], off)); }
    if (lane < 3) { blockBox[9 * b + lane] = mn[lane]; blockBox[9 * b + 3 + lane] = mx[lane]; blockBox[9 * b + 6 + lane] = 0.5 * (mn[lane] + mx[lane]); }
}

// ------------------------------------------------------------------------------------------------------
// the tile builder: one WARP per (sort block of 32 atoms, set); nothing is shared between the warps of a CTA.
//
// Stage A walks the cell rows around the block box and rejects candidates against the box; the survivors are compacted
// into a warp queue, so that stage B -- the 32 distance tests per candidate, fp32 with an exact fp64 decision inside the
// rounding band -- always runs on full warps.  A candidate's 32-bit column mask is split into the four bytes of the block's
// four i-clusters (8 atoms each): every cluster gets its own stream of j entries (a cluster only lists the atoms that are
// within the cutoff of one of ITS 8 atoms: 65 % of the 8 x 32 slots of a tile are list pairs, against 47 % for 32 x 32).
// Tiles are written into a global pool that is handed out in chunks of `chunkTiles` tiles; a chunk is a work item of the
// force kernel.
// ------------------------------------------------------------------------------------------------------
struct TileArgs {
    int n, nblocks, nsets, firstBlock, myBlocks, selfEnabled, chunkTiles, rawJ, split, splitShift, imgParts, imgShift;
    unsigned int totalWarps, primaryWarps;
    double cutoff, cutoff2;
    BuildGrid grid;
    const double *sX; const int *sAtom; const int *invPerm;
    const unsigned int *cellStart;
    const double *blockBox;
    const ImageBoxDev *imageBoxes;     // [nsets], slot 0 unused
    const int *exclPtr; const int *exclCol;
    const unsigned char *fixed;        // nullable: per-atom flags of the fixed atoms
    const unsigned char *inactive;     // nullable: per-atom flags of the atoms that are on no MM/MM list (pure QC atoms: mmSelection, NBModelABFSState.c:351)
    unsigned int *tileDesc; unsigned int tileCap;
    WorkItem *items; unsigned int itemCap;
    unsigned long long *setPairs;
    DeviceCounters *counters;
};

constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kSubBlocks = kTile / kCluster;
constexpr int kBloomWords = 64;
constexpr int kScanUnroll = 4;                  // stage A keeps this many 32-candidate chunks in flight
constexpr int kCandRing = 256;                  // >= 32 * (kScanUnroll + 1), power of two

// per-warp stream state of the four i-clusters: the open chunk's first tile in shared memory (read once per emitted tile), everything else
// packed into two registers (one byte per cluster): `counts` = entries waiting in the queue, `state` = tiles used in the open chunk
// (bits 0-5) and which half of the 64-entry ring the queue starts in (bit 6).  The stream loop is not unrolled: code size matters
// (the first version of this kernel spent most of its time on instruction-cache misses)
constexpr unsigned int kNoChunk = 0xffffffffu;  // the pool overflowed when this chunk was claimed: tiles are counted, not written

struct __align__(16) BuildWarp {
    double sxi[kTile][3];                        // exact coordinates of the block atoms
    float4 sxy[kTile / 2];                       // block-local fp32 copies for the prefilter, atoms paired (i, i+16):
    float2 szz[kTile / 2];                       //   {x_i, x_i+16, y_i, y_i+16} and {z_i, z_i+16}
    int rowStart[kTile], rowCount[kTile];
    float4 cand[kCandRing];                      // ring of the candidates that survived the box reject: block-local fp32 x, y, z and the sorted position
    unsigned int bloom[kBloomWords];             // sorted positions (mod 2048) of the exclusion partners of the block atoms
    unsigned int sub[kSubBlocks][2 * kTile];     // per i-cluster: ring of j reference | column byte << 24
    unsigned int chunkBase[kSubBlocks];
    unsigned int activeMask, padw[3];            // block atoms that may appear on a list (kept in shared memory: the builder is register bound)
};

// The kernel is bound by latency (dependent loads of the scan, shared-memory queues): resident warps matter more than registers per
// thread.  Measured on B200 (M1, whole rebuild): plain bound 86 registers / 5 CTAs per SM 1.84 ms, 6 CTAs 1.75 ms, 7 CTAs (72 registers)
// 1.71 ms; an explicit minimum of 1 lets ptxas take 132 registers (2.33 ms).  -DNBB_BUILD_MINBLOCKS=n overrides.
// kQC: a QC region is present (A.inactive); a separate instantiation, so that the ordinary builder compiles exactly as without it
#ifndef NBB_BUILD_MINBLOCKS
#define NBB_BUILD_MINBLOCKS 7
#endif
template <bool kQC>
__global__ void __launch_bounds__(kBuildThreads, NBB_BUILD_MINBLOCKS) k_build_tiles(TileArgs A)
{
    __shared__ BuildWarp sw[kBuildWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int wg = blockIdx.x * kBuildWarps + warp;
    if (wg >= A.totalWarps) return;                                    // whole warps leave: no CTA barrier below
    // Units: one warp per (block, part) for the primary set, then one warp per (block, part, image part) that walks the image sets
    // whose list-time box overlaps the block (most (block, image) combinations of a large system do not: a warp per combination
    // spent 8 % of the kernel's instructions on warps that left at once).  Small systems deal the rows of a (block, set) to `split`
    // warps and the image sets of a block to `imgParts` warps (powers of two), each with its own j streams.
    // image units come FIRST in the grid: they are few, short on arithmetic and long on latency (one row walk per overlapping set)
    // (placed behind the primary units they were a tail of 0.1 ms on the 1.1 M-atom box)
    const unsigned int imageWarps = A.totalWarps - A.primaryWarps;
    const bool imageUnit = wg < imageWarps;
    const unsigned int wu = imageUnit ? wg : wg - imageWarps;
    const int part = (int) (wu & (unsigned int) (A.split - 1));
    const unsigned int wr = wu >> A.splitShift;
    const int ip = imageUnit ? (int) (wr & (unsigned int) (A.imgParts - 1)) : 0;
    const int b = A.firstBlock + (int) (imageUnit ? (wr >> A.imgShift) : wr);
    if (!imageUnit && !A.selfEnabled) return;
    const double reach = A.cutoff + 1.0e-6;
    double sbox[9];
#pragma unroll
    for (int d = 0; d < 6; d++) sbox[d] = A.blockBox[9 * b + d];
    int set = 0, setBase = 1;
    unsigned int pending = 0u;
    if (imageUnit) {                                         // first overlapping image set, or leave
        while (pending == 0u && setBase < A.nsets) {
            const int sc = setBase + lane;
            bool overlap = sc < A.nsets && ((sc - 1) & (A.imgParts - 1)) == ip;
            if (overlap) {
                const ImageBoxDev ib = A.imageBoxes[sc];
                for (int d = 0; d < 3; d++) overlap = overlap && (ib.lo[d] <= sbox[3 + d] + reach) && (ib.hi[d] >= sbox[d] - reach);
            }
            pending = __ballot_sync(0xffffffffu, overlap);
            setBase += kTile;
        }
        if (pending == 0u) return;
        set = setBase - kTile + __ffs(pending) - 1;
        pending &= pending - 1u;
    }
#pragma unroll
    for (int d = 6; d < 9; d++) sbox[d] = A.blockBox[9 * b + d];
    BuildWarp &W = sw[warp];
    {
        const int s = b * kTile + lane;
        float f[3];
        for (int d = 0; d < 3; d++) {
            const double v = (s < A.n) ? A.sX[3 * s + d] : 1.0e30;            // padding rows never pass the test
            W.sxi[lane][d] = v;
            f[d] = (s < A.n) ? (float) (v - sbox[6 + d]) : 1.0e15f;
        }
        float *pxy = reinterpret_cast<float *>(W.sxy), *pzz = reinterpret_cast<float *>(W.szz);
        const int m = lane & 15, h = lane >> 4;
        pxy[4 * m + h] = f[0]; pxy[4 * m + 2 + h] = f[1]; pzz[2 * m + h] = f[2];
        // exclusions can only remove pairs whose partner is excluded by a block atom: a small Bloom filter over the partners'
        // sorted positions spares almost every candidate the dependent global loads of its exclusion list
        W.bloom[lane] = 0u; W.bloom[lane + 32] = 0u;
        if (lane < kSubBlocks) W.chunkBase[lane] = 0u;
        __syncwarp();
        if (set == 0 && s < A.n) {
            const int ai = A.sAtom[s];
            for (int k = A.exclPtr[ai]; k < A.exclPtr[ai + 1]; k++) {
                const int sp = A.invPerm[A.exclCol[k]];
                atomicOr(&W.bloom[(sp >> 5) & (kBloomWords - 1)], 1u << (sp & 31));
            }
        }
    }
    __syncwarp();

    // fixed atoms: a pair stays on the lists only if one of its atoms is free (orSelection = freeSelection of the reference generators)
    unsigned int freeMask = 0xffffffffu;
    if (A.fixed != nullptr) {
        const int sb = b * kTile + lane;
        freeMask = __ballot_sync(0xffffffffu, !(sb < A.n && A.fixed[A.sAtom[sb]]));
    }
    if (kQC) {
        const int sb = b * kTile + lane;
        const unsigned int act = __ballot_sync(0xffffffffu, !(sb < A.n && A.inactive[A.sAtom[sb]]));
        if (lane == 0) W.activeMask = act;
        __syncwarp();
    }
    const BuildGrid g = A.grid;
    int c0[3], c1[3];
    double maxAbs = 0.0;
    for (int d = 0; d < 3; d++) {
        c0[d] = cell_coord(sbox[d] - reach, g.lo[d], g.invh, g.dim[d]);
        c1[d] = cell_coord(sbox[3 + d] + reach, g.lo[d], g.invh, g.dim[d]);
        maxAbs = fmax(maxAbs, 0.5 * (sbox[3 + d] - sbox[d]) + reach);
    }
    const int nrowsY = c1[1] - c0[1] + 1, nrowsTotal = (c1[0] - c0[0] + 1) * nrowsY;
    // box reject in fp32: half extents of the block box around its centre; candidates are at most ~maxAbs away when they matter, so the
    // rounding of the local coordinates (<= 6e-8 * maxAbs each) moves r2 by far less than the margin
    const float hx = (float) (0.5 * (sbox[3] - sbox[0])), hy = (float) (0.5 * (sbox[4] - sbox[1])), hz = (float) (0.5 * (sbox[5] - sbox[2]));
    const float reject2f = (float) (A.cutoff2 * (1.0 + 1.0e-5) + 1.0e-3);
    // fp32 prefilter: |r2_fp32 - r2_exact| <= eps for every candidate that survives the box reject (block-local coordinates,
    // magnitude <= maxAbs); decisions inside the band are taken by the exact fp64 predicate
    const double delta = 6.0e-7 * maxAbs;
    const float eps = (float) (2.0 * (3.5 * reach * delta + 3.0e-7 * A.cutoff2) + 2.0e-5);   // also covers the rounding of cutoff^2 to fp32
    const float c2f = (float) A.cutoff2;
    const unsigned int ltMask = (1u << lane) - 1u;

    for (;;) {                                   // the sets of this unit: the primary set, or the overlapping image sets one after the other
    int candHead = 0, candTail = 0;
    unsigned int counts = 0u, state = 0u;        // per cluster, one byte each: entries waiting in the queue; tiles used in the open chunk | ring half << 6
    unsigned long long myPairs = 0;
    // scan cursor: rows are taken in batches of 32 (row tables in shared memory), each row in units of kScanUnroll chunks
    int rowBase = -kTile, nrows = 0, r = 0, base = 0, rs = 0, rc = 0;
    bool done = false;

    // ONE loop body holds the scan (stage A), the distance tests (stage B) and the tile emission, each exactly once in the code
    for (;;) {
        // ---- advance the scan cursor to the next unit with work
        while (!done && base >= rc) {
            r += A.split;
            if (r >= nrows) {
                rowBase += kTile;
                if (rowBase >= nrowsTotal) { done = true; break; }
                nrows = min(kTile, nrowsTotal - rowBase);
                __syncwarp();                                    // everybody is done with the previous row tables
                if (lane < nrows) {
                    const int rr = rowBase + lane, cx = c0[0] + rr / nrowsY, cy = c0[1] + rr % nrowsY;
                    // z range of this row: only the part of the column of cells that can be within reach of the block box
                    const double xlo = g.lo[0] + cx * g.h, ylo = g.lo[1] + cy * g.h;
                    const double ex = fmax(0.0, fmax(sbox[0] - (xlo + g.h), xlo - sbox[3])), ey = fmax(0.0, fmax(sbox[1] - (ylo + g.h), ylo - sbox[4]));
                    const double rem = reach * reach - ex * ex - ey * ey;
                    int start = 0, end = 0;
                    if (rem >= 0.0) {
                        const double dz = sqrt(rem) + 1.0e-6;
                        const int z0 = cell_coord(sbox[2] - dz, g.lo[2], g.invh, g.dim[2]), z1 = cell_coord(sbox[5] + dz, g.lo[2], g.invh, g.dim[2]);
                        const int keyLo = set * g.ncell + min(snake_cell(g, cx, cy, z0), snake_cell(g, cx, cy, z1));   // the z run is contiguous either way
                        start = (int) A.cellStart[keyLo]; end = (int) A.cellStart[keyLo + (z1 - z0) + 1];
                        if (set == 0) start = max(start, b * kTile);     // primary list: each unordered pair once (own block: triangle in stage B)
                    }
                    W.rowStart[lane] = start;
                    W.rowCount[lane] = max(0, end - start);
                }
                __syncwarp();
                r = part;
                if (r >= nrows) { rc = 0; base = 0; continue; }
            }
            rs = W.rowStart[r]; rc = W.rowCount[r]; base = 0;
        }
        // ---- stage A: box reject of kScanUnroll chunks of 32 candidates; all loads first, the scan is latency bound
        if (!done) {
            double xj[kScanUnroll], yj[kScanUnroll], zj[kScanUnroll];
#pragma unroll
            for (int u = 0; u < kScanUnroll; u++) {
                const int c = base + u * kTile + lane;
                xj[u] = 1.0e30; yj[u] = 1.0e30; zj[u] = 1.0e30;                 // out of range: rejected by the box test
                if (c < rc) { const int s = rs + c; xj[u] = A.sX[3 * s]; yj[u] = A.sX[3 * s + 1]; zj[u] = A.sX[3 * s + 2]; }
            }
#pragma unroll
            for (int u = 0; u < kScanUnroll; u++) {
                if (base + u * kTile < rc) {
                    // conservative reject against the block box, in block-local fp32 (the threshold carries the rounding margin)
                    const float fx = (float) (xj[u] - sbox[6]), fy = (float) (yj[u] - sbox[7]), fz = (float) (zj[u] - sbox[8]);
                    const float ex = fmaxf(0.f, fabsf(fx) - hx), ey = fmaxf(0.f, fabsf(fy) - hy), ez = fmaxf(0.f, fabsf(fz) - hz);
                    const bool keep = fmaf(ex, ex, fmaf(ey, ey, ez * ez)) <= reject2f;
                    const unsigned int bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) W.cand[(candTail + __popc(bal & ltMask)) & (kCandRing - 1)] = make_float4(fx, fy, fz, __int_as_float(rs + base + u * kTile + lane));
                    candTail += __popc(bal);
                }
            }
            base += kTile * kScanUnroll;
            __syncwarp();
        }
        // ---- stage B + emission: full batches while scanning; at the end the remainder, then one pass that only flushes the streams
        bool flushed = false;
        while (candTail - candHead >= (done ? 1 : kTile) || (done && !flushed)) {
            const int count = min(kTile, candTail - candHead);
            flushed = count == 0;
            unsigned int colmask = 0u, jref = 0u;
            if (lane < count) {
                const float4 cj = W.cand[(candHead + lane) & (kCandRing - 1)];
                const int s = __float_as_int(cj.w);
                const float fx = cj.x, fy = cj.y, fz = cj.z;
                float band = 1.0e30f;                        // min over the block atoms of |r2 - cutoff^2|
                const f32x2 fx2 = pk2(fx, fx), fy2 = pk2(fy, fy), fz2 = pk2(fz, fz), c22 = pk2(c2f, c2f);
                unsigned int ca = 0u, cb = 0u;               // sign bits of r2 - cutoff^2, shifted in from the right
#pragma unroll
                for (int i = 0; i < kTile / 2; i++) {        // block atoms i and i + 16 in one packed evaluation
                    const float4 pxy = W.sxy[i];
                    const float2 pz = W.szz[i];
                    const f32x2 dx = sub2(pk2(pxy.x, pxy.y), fx2), dy = sub2(pk2(pxy.z, pxy.w), fy2), dz = sub2(pk2(pz.x, pz.y), fz2);
                    float ta, tb;
                    unpk2(sub2(fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))), c22), ta, tb);
                    ca = __funnelshift_l(__float_as_uint(ta), ca, 1);
                    cb = __funnelshift_l(__float_as_uint(tb), cb, 1);
                    band = fminf(band, fminf(fabsf(ta), fabsf(tb)));
                }
                colmask = (__brev(ca) >> 16) | (__brev(cb) & 0xffff0000u);     // r2 < cutoff^2 (equality sits inside the band)
                if (band <= eps) {                           // some distance is within the fp32 error band: the reference predicate decides
                    colmask = 0u;
                    const double xj = A.sX[3 * s], yj = A.sX[3 * s + 1], zj = A.sX[3 * s + 2];
                    for (int i = 0; i < kTile; i++) {
                        const double r2 = ref_dist2(W.sxi[i][0] - xj, W.sxi[i][1] - yj, W.sxi[i][2] - zj);
                        colmask |= (r2 <= A.cutoff2) ? (1u << i) : 0u;
                    }
                }
                if (colmask != 0u && A.fixed != nullptr && A.fixed[A.sAtom[s]]) colmask &= freeMask;   // fixed j: free i atoms only
                if (kQC && colmask != 0u) colmask = A.inactive[A.sAtom[s]] ? 0u : (colmask & W.activeMask);   // both atoms in the MM selection
                if (colmask != 0u) {
                    if (set == 0) {
                        if ((s >> 5) == b) colmask &= (1u << (s & 31)) - 1u;          // own block: i < j only, no self pair
                        if ((W.bloom[(s >> 5) & (kBloomWords - 1)] >> (s & 31)) & 1u) {
                            const int atom = A.sAtom[s];
                            for (int k = A.exclPtr[atom]; k < A.exclPtr[atom + 1]; k++) {
                                const int sp = A.invPerm[A.exclCol[k]];
                                if ((sp >> 5) == b) colmask &= ~(1u << (sp & 31));
                            }
                        }
                        jref = (unsigned int) s;                                      // primary atoms: sorted position = extended position
                    } else {
                        const int atom = A.sAtom[s];
                        jref = A.rawJ ? (unsigned int) atom : (unsigned int) A.invPerm[atom];
                    }
                }
            }
            candHead += count;
            myPairs += __popc(colmask);
            // push the candidate into the queues (64-entry rings) of the clusters it pairs with (unrolled, no barrier in between) ...
#pragma unroll
            for (int q = 0; q < kSubBlocks; q++) {
                const unsigned int byte = (colmask >> (kCluster * q)) & 0xffu;
                const unsigned int bal = __ballot_sync(0xffffffffu, byte != 0u);
                const unsigned int tail = ((state >> (8 * q + 1)) & 32u) + ((counts >> (8 * q)) & 0xffu);
                if (byte != 0u) W.sub[q][(tail + __popc(bal & ltMask)) & (2 * kTile - 1)] = jref | (byte << 24);
                counts += (unsigned int) __popc(bal) << (8 * q);
            }
            __syncwarp();
            // ... and emit full tiles (at the very end: whatever is left, and close the open chunks).  A tile = the first 32 entries of
            // the ring: lane l of the force kernel owns j slot l, the high byte of its word is the column byte exactly as queued
#pragma unroll                                   // compile-time byte positions: 1.53 -> 1.48 ms (M1 rebuild) against the rolled loop
            for (int q = 0; q < kSubBlocks; q++) {
                const int sh = 8 * q;
                const int cnt = (int) ((counts >> sh) & 0xffu);
                const unsigned int stq = (state >> sh) & 0xffu;
                int used = (int) (stq & 63u);
                if (cnt < kTile && !(flushed && (cnt | used) != 0)) continue;
                const int n = min(cnt, kTile);
                const unsigned int half = stq & 64u;                         // ring half the queue starts in (x 32 entries / 2)
                unsigned int base = W.chunkBase[q];
                if (used == 0 && n > 0) {                                    // open a chunk (= the next work item of this stream)
                    if (lane == 0) base = atomicAdd(&A.counters->tileTotal, (unsigned int) A.chunkTiles);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base > A.tileCap - (unsigned int) A.chunkTiles) {   // host: tileCap >= chunkTiles, tileCap * 32 < 2^31
                        if (lane == 0) atomicOr(&A.counters->overflow, 2u);
                        base = kNoChunk;
                    }
                    if (lane == 0) W.chunkBase[q] = base;
                }
                if (n > 0) {
                    const unsigned int word = (lane < n) ? W.sub[q][(half >> 1) + lane] : kEmptySlot;
                    if (base != kNoChunk) A.tileDesc[(base + (unsigned int) used) * kTile + lane] = word;
                    used += 1;
                }
                const bool close = used == A.chunkTiles || n < kTile;       // chunk full, or the (padded) last tile of the stream
                if (close && base != kNoChunk && lane == 0) {
                    const unsigned int pos = atomicAdd(&A.counters->itemCount, 1u);
                    if (pos < A.itemCap) {
                        WorkItem w;
                        w.block = kSubBlocks * b + q; w.image = set; w.tileStart = (int) base; w.tileCount = used;
                        A.items[pos] = w;
                    } else atomicOr(&A.counters->overflow, 4u);
                    atomicAdd(&A.counters->tilesUsed, (unsigned int) used);
                }
                const unsigned int nst = (close ? 0u : (unsigned int) used) | ((n == kTile) ? (half ^ 64u) : half);
                state = (state & ~(0xffu << sh)) | (nst << sh);
                counts -= (unsigned int) n << sh;
                __syncwarp();                                                // ring slots and chunkBase are reused by the next pushes / tiles
            }
        }
        if (done) break;
    }
    for (int o = 16; o > 0; o >>= 1) myPairs += __shfl_xor_sync(0xffffffffu, myPairs, o);
    if (lane == 0 && myPairs) atomicAdd(&A.setPairs[set], myPairs);
    // ---- next overlapping image set of this unit
    if (!imageUnit) break;
    while (pending == 0u && setBase < A.nsets) {
        const int sc = setBase + lane;
        bool overlap = sc < A.nsets && ((sc - 1) & (A.imgParts - 1)) == ip;
        if (overlap) {
            const ImageBoxDev ib = A.imageBoxes[sc];
            for (int d = 0; d < 3; d++) overlap = overlap && (ib.lo[d] <= sbox[3 + d] + reach) && (ib.hi[d] >= sbox[d] - reach);
        }
        pending = __ballot_sync(0xffffffffu, overlap);
        setBase += kTile;
    }
    if (pending == 0u) break;
    set = setBase - kTile + __ffs(pending) - 1;
    pending &= pending - 1u;
    }
}

// ------------------------------------------------------------------------------------------------------
// the cooperative variant for small systems: the FOUR warps of a CTA work on one (block, set) -- each walks every fourth cell row (stages A
// and B exactly as above) -- and push into ONE ring per i-cluster in shared memory (the slots of a step are dealt per
// cluster, in warp order: the warps publish their entry counts, a CTA barrier, every warp writes behind the warps before it); after a second
// barrier warp q alone emits the full tiles of cluster q and owns its chunk state.  Four times the warps of the one-warp-per-block builder
// (a 23 k-atom box fills a fifth of the GPU with it) without splitting the j streams: dealing the rows to independent warps gives every
// part its own, padded, streams and the force kernel pays for the padding at every call.  The ring (256 entries) holds what is waiting
// (< 32 after an emission) plus the entries of one step (<= 128).
// ------------------------------------------------------------------------------------------------------
constexpr int kCoopRing = 256;                   // positions are tracked bytewise
struct CoopQueues {
    unsigned int ring[kSubBlocks][kCoopRing];
    unsigned int count[kBuildWarps];             // per warp: entries of the current step per cluster, one byte each
};

template <bool kQC>
__global__ void __launch_bounds__(kBuildThreads, NBB_BUILD_MINBLOCKS) k_build_tiles_coop(TileArgs A)
{
    static_assert(kBuildWarps == kSubBlocks, "warp q of the CTA owns the stream of cluster q");
    __shared__ BuildWarp sw[kBuildWarps];
    __shared__ CoopQueues cq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int wg = blockIdx.x * kBuildWarps + warp;
    if (wg >= A.totalWarps) return;                                    // whole warps leave: no CTA barrier below
    // Units: one warp per (block, part) for the primary set, then one warp per (block, part, image part) that walks the image sets
    // whose list-time box overlaps the block (most (block, image) combinations of a large system do not: a warp per combination
    // spent 8 % of the kernel's instructions on warps that left at once).  Small systems deal the rows of a (block, set) to `split`
    // warps and the image sets of a block to `imgParts` warps (powers of two), each with its own j streams.
    // image units come FIRST in the grid: they are few, short on arithmetic and long on latency (one row walk per overlapping set)
    // (placed behind the primary units they were a tail of 0.1 ms on the 1.1 M-atom box)
    const unsigned int imageWarps = A.totalWarps - A.primaryWarps;
    const bool imageUnit = wg < imageWarps;
    const unsigned int wu = imageUnit ? wg : wg - imageWarps;
    const int part = (int) (wu & (unsigned int) (A.split - 1));
    const unsigned int wr = wu >> A.splitShift;
    const int ip = imageUnit ? (int) (wr & (unsigned int) (A.imgParts - 1)) : 0;
    const int b = A.firstBlock + (int) (imageUnit ? (wr >> A.imgShift) : wr);
    if (!imageUnit && !A.selfEnabled) return;
    const double reach = A.cutoff + 1.0e-6;
    double sbox[9];
#pragma unroll
    for (int d = 0; d < 6; d++) sbox[d] = A.blockBox[9 * b + d];
    int set = 0, setBase = 1;
    unsigned int pending = 0u;
    if (imageUnit) {                                         // first overlapping image set, or leave
        while (pending == 0u && setBase < A.nsets) {
            const int sc = setBase + lane;
            bool overlap = sc < A.nsets && ((sc - 1) & (A.imgParts - 1)) == ip;
            if (overlap) {
                const ImageBoxDev ib = A.imageBoxes[sc];
                for (int d = 0; d < 3; d++) overlap = overlap && (ib.lo[d] <= sbox[3 + d] + reach) && (ib.hi[d] >= sbox[d] - reach);
            }
            pending = __ballot_sync(0xffffffffu, overlap);
            setBase += kTile;
        }
        if (pending == 0u) return;
        set = setBase - kTile + __ffs(pending) - 1;
        pending &= pending - 1u;
    }
#pragma unroll
    for (int d = 6; d < 9; d++) sbox[d] = A.blockBox[9 * b + d];
    BuildWarp &W = sw[warp];
    {
        const int s = b * kTile + lane;
        float f[3];
        for (int d = 0; d < 3; d++) {
            const double v = (s < A.n) ? A.sX[3 * s + d] : 1.0e30;            // padding rows never pass the test
            W.sxi[lane][d] = v;
            f[d] = (s < A.n) ? (float) (v - sbox[6 + d]) : 1.0e15f;
        }
        float *pxy = reinterpret_cast<float *>(W.sxy), *pzz = reinterpret_cast<float *>(W.szz);
        const int m = lane & 15, h = lane >> 4;
        pxy[4 * m + h] = f[0]; pxy[4 * m + 2 + h] = f[1]; pzz[2 * m + h] = f[2];
        // exclusions can only remove pairs whose partner is excluded by a block atom: a small Bloom filter over the partners'
        // sorted positions spares almost every candidate the dependent global loads of its exclusion list
        W.bloom[lane] = 0u; W.bloom[lane + 32] = 0u;
        if (lane < kSubBlocks) W.chunkBase[lane] = 0u;
        __syncwarp();
        if (set == 0 && s < A.n) {
            const int ai = A.sAtom[s];
            for (int k = A.exclPtr[ai]; k < A.exclPtr[ai + 1]; k++) {
                const int sp = A.invPerm[A.exclCol[k]];
                atomicOr(&W.bloom[(sp >> 5) & (kBloomWords - 1)], 1u << (sp & 31));
            }
        }
    }
    __syncwarp();

    // fixed atoms: a pair stays on the lists only if one of its atoms is free (orSelection = freeSelection of the reference generators)
    unsigned int freeMask = 0xffffffffu;
    if (A.fixed != nullptr) {
        const int sb = b * kTile + lane;
        freeMask = __ballot_sync(0xffffffffu, !(sb < A.n && A.fixed[A.sAtom[sb]]));
    }
    if (kQC) {
        const int sb = b * kTile + lane;
        const unsigned int act = __ballot_sync(0xffffffffu, !(sb < A.n && A.inactive[A.sAtom[sb]]));
        if (lane == 0) W.activeMask = act;
        __syncwarp();
    }
    const BuildGrid g = A.grid;
    int c0[3], c1[3];
    double maxAbs = 0.0;
    for (int d = 0; d < 3; d++) {
        c0[d] = cell_coord(sbox[d] - reach, g.lo[d], g.invh, g.dim[d]);
        c1[d] = cell_coord(sbox[3 + d] + reach, g.lo[d], g.invh, g.dim[d]);
        maxAbs = fmax(maxAbs, 0.5 * (sbox[3 + d] - sbox[d]) + reach);
    }
    const int nrowsY = c1[1] - c0[1] + 1, nrowsTotal = (c1[0] - c0[0] + 1) * nrowsY;
    // box reject in fp32: half extents of the block box around its centre; candidates are at most ~maxAbs away when they matter, so the
    // rounding of the local coordinates (<= 6e-8 * maxAbs each) moves r2 by far less than the margin
    const float hx = (float) (0.5 * (sbox[3] - sbox[0])), hy = (float) (0.5 * (sbox[4] - sbox[1])), hz = (float) (0.5 * (sbox[5] - sbox[2]));
    const float reject2f = (float) (A.cutoff2 * (1.0 + 1.0e-5) + 1.0e-3);
    // fp32 prefilter: |r2_fp32 - r2_exact| <= eps for every candidate that survives the box reject (block-local coordinates,
    // magnitude <= maxAbs); decisions inside the band are taken by the exact fp64 predicate
    const double delta = 6.0e-7 * maxAbs;
    const float eps = (float) (2.0 * (3.5 * reach * delta + 3.0e-7 * A.cutoff2) + 2.0e-5);   // also covers the rounding of cutoff^2 to fp32
    const float c2f = (float) A.cutoff2;
    const unsigned int ltMask = (1u << lane) - 1u;

    for (;;) {                                   // the sets of this unit: the primary set, or the overlapping image sets one after the other
    int candHead = 0, candTail = 0;
    unsigned int head = 0u;                      // warp q: ring position of the first waiting entry of cluster q
    int avail = 0, used = 0;                     // warp q: entries waiting in the ring of cluster q; tiles used in its open chunk
    unsigned int chunkBase = 0u, tailPos = 0u;   // tailPos: ring positions (mod 256, one byte per cluster) behind the last entry, tracked by every warp
    unsigned long long myPairs = 0;
    // scan cursor: rows are taken in batches of 32 (row tables in shared memory), each row in units of kScanUnroll chunks
    int rowBase = -kTile, nrows = 0, r = 0, base = 0, rs = 0, rc = 0;
    bool done = false;

    // every step: scan (stage A) when fewer than 32 candidates are waiting, one batch of distance tests (stage B), push; barrier; warp q emits
    for (;;) {
        if (!done && candTail - candHead < kTile) {
            // ---- advance the scan cursor to the next unit with work
            while (!done && base >= rc) {
                r += A.split;
                if (r >= nrows) {
                    rowBase += kTile;
                    if (rowBase >= nrowsTotal) { done = true; break; }
                    nrows = min(kTile, nrowsTotal - rowBase);
                    __syncwarp();                                    // everybody is done with the previous row tables
                    if (lane < nrows) {
                        const int rr = rowBase + lane, cx = c0[0] + rr / nrowsY, cy = c0[1] + rr % nrowsY;
                        // z range of this row: only the part of the column of cells that can be within reach of the block box
                        const double xlo = g.lo[0] + cx * g.h, ylo = g.lo[1] + cy * g.h;
                        const double ex = fmax(0.0, fmax(sbox[0] - (xlo + g.h), xlo - sbox[3])), ey = fmax(0.0, fmax(sbox[1] - (ylo + g.h), ylo - sbox[4]));
                        const double rem = reach * reach - ex * ex - ey * ey;
                        int start = 0, end = 0;
                        if (rem >= 0.0) {
                            const double dz = sqrt(rem) + 1.0e-6;
                            const int z0 = cell_coord(sbox[2] - dz, g.lo[2], g.invh, g.dim[2]), z1 = cell_coord(sbox[5] + dz, g.lo[2], g.invh, g.dim[2]);
                            const int keyLo = set * g.ncell + min(snake_cell(g, cx, cy, z0), snake_cell(g, cx, cy, z1));   // the z run is contiguous either way
                            start = (int) A.cellStart[keyLo]; end = (int) A.cellStart[keyLo + (z1 - z0) + 1];
                            if (set == 0) start = max(start, b * kTile);     // primary list: each unordered pair once (own block: triangle in stage B)
                        }
                        W.rowStart[lane] = start;
                        W.rowCount[lane] = max(0, end - start);
                    }
                    __syncwarp();
                    r = part;
                    if (r >= nrows) { rc = 0; base = 0; continue; }
                }
                rs = W.rowStart[r]; rc = W.rowCount[r]; base = 0;
            }

            // ---- stage A: box reject of kScanUnroll chunks of 32 candidates; all loads first, the scan is latency bound
            if (!done) {
                double xj[kScanUnroll], yj[kScanUnroll], zj[kScanUnroll];
    #pragma unroll
                for (int u = 0; u < kScanUnroll; u++) {
                    const int c = base + u * kTile + lane;
                    xj[u] = 1.0e30; yj[u] = 1.0e30; zj[u] = 1.0e30;                 // out of range: rejected by the box test
                    if (c < rc) { const int s = rs + c; xj[u] = A.sX[3 * s]; yj[u] = A.sX[3 * s + 1]; zj[u] = A.sX[3 * s + 2]; }
                }
    #pragma unroll
                for (int u = 0; u < kScanUnroll; u++) {
                    if (base + u * kTile < rc) {
                        // conservative reject against the block box, in block-local fp32 (the threshold carries the rounding margin)
                        const float fx = (float) (xj[u] - sbox[6]), fy = (float) (yj[u] - sbox[7]), fz = (float) (zj[u] - sbox[8]);
                        const float ex = fmaxf(0.f, fabsf(fx) - hx), ey = fmaxf(0.f, fabsf(fy) - hy), ez = fmaxf(0.f, fabsf(fz) - hz);
                        const bool keep = fmaf(ex, ex, fmaf(ey, ey, ez * ez)) <= reject2f;
                        const unsigned int bal = __ballot_sync(0xffffffffu, keep);
                        if (keep) W.cand[(candTail + __popc(bal & ltMask)) & (kCandRing - 1)] = make_float4(fx, fy, fz, __int_as_float(rs + base + u * kTile + lane));
                        candTail += __popc(bal);
                    }
                }
                base += kTile * kScanUnroll;
                __syncwarp();
            }
        }
        unsigned int colmask = 0u, jref = 0u, bal[kSubBlocks];
        {
            const int count = (candTail - candHead >= kTile || done) ? min(kTile, candTail - candHead) : 0;
            if (lane < count) {
                const float4 cj = W.cand[(candHead + lane) & (kCandRing - 1)];
                const int s = __float_as_int(cj.w);
                const float fx = cj.x, fy = cj.y, fz = cj.z;
                float band = 1.0e30f;                        // min over the block atoms of |r2 - cutoff^2|
                const f32x2 fx2 = pk2(fx, fx), fy2 = pk2(fy, fy), fz2 = pk2(fz, fz), c22 = pk2(c2f, c2f);
                unsigned int ca = 0u, cb = 0u;               // sign bits of r2 - cutoff^2, shifted in from the right
#pragma unroll
                for (int i = 0; i < kTile / 2; i++) {        // block atoms i and i + 16 in one packed evaluation
                    const float4 pxy = W.sxy[i];
                    const float2 pz = W.szz[i];
                    const f32x2 dx = sub2(pk2(pxy.x, pxy.y), fx2), dy = sub2(pk2(pxy.z, pxy.w), fy2), dz = sub2(pk2(pz.x, pz.y), fz2);
                    float ta, tb;
                    unpk2(sub2(fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))), c22), ta, tb);
                    ca = __funnelshift_l(__float_as_uint(ta), ca, 1);
                    cb = __funnelshift_l(__float_as_uint(tb), cb, 1);
                    band = fminf(band, fminf(fabsf(ta), fabsf(tb)));
                }
                colmask = (__brev(ca) >> 16) | (__brev(cb) & 0xffff0000u);     // r2 < cutoff^2 (equality sits inside the band)
                if (band <= eps) {                           // some distance is within the fp32 error band: the reference predicate decides
                    colmask = 0u;
                    const double xj = A.sX[3 * s], yj = A.sX[3 * s + 1], zj = A.sX[3 * s + 2];
                    for (int i = 0; i < kTile; i++) {
                        const double r2 = ref_dist2(W.sxi[i][0] - xj, W.sxi[i][1] - yj, W.sxi[i][2] - zj);
                        colmask |= (r2 <= A.cutoff2) ? (1u << i) : 0u;
                    }
                }
                if (colmask != 0u && A.fixed != nullptr && A.fixed[A.sAtom[s]]) colmask &= freeMask;   // fixed j: free i atoms only
                if (kQC && colmask != 0u) colmask = A.inactive[A.sAtom[s]] ? 0u : (colmask & W.activeMask);   // both atoms in the MM selection
                if (colmask != 0u) {
                    if (set == 0) {
                        if ((s >> 5) == b) colmask &= (1u << (s & 31)) - 1u;          // own block: i < j only, no self pair
                        if ((W.bloom[(s >> 5) & (kBloomWords - 1)] >> (s & 31)) & 1u) {
                            const int atom = A.sAtom[s];
                            for (int k = A.exclPtr[atom]; k < A.exclPtr[atom + 1]; k++) {
                                const int sp = A.invPerm[A.exclCol[k]];
                                if ((sp >> 5) == b) colmask &= ~(1u << (sp & 31));
                            }
                        }
                        jref = (unsigned int) s;                                      // primary atoms: sorted position = extended position
                    } else {
                        const int atom = A.sAtom[s];
                        jref = A.rawJ ? (unsigned int) atom : (unsigned int) A.invPerm[atom];
                    }
                }
            }
            candHead += count;
            myPairs += __popc(colmask);
            // how many entries this warp adds to each cluster's ring (one byte per cluster): published, then everybody knows everybody's
            // share and the entries land in warp order -- the streams are the same from run to run (the fp32 partial sums of the force
            // kernel follow the entry order: a first version that reserved the slots with atomics was not reproducible beyond ~1e-8)
#pragma unroll
            for (int q = 0; q < kSubBlocks; q++) {
                const unsigned int byte = (colmask >> (kCluster * q)) & 0xffu;
                bal[q] = __ballot_sync(0xffffffffu, byte != 0u);
            }
            if (lane == 0) cq.count[warp] = (unsigned int) __popc(bal[0]) | ((unsigned int) __popc(bal[1]) << 8) | ((unsigned int) __popc(bal[2]) << 16) | ((unsigned int) __popc(bal[3]) << 24);
        }
        const bool allDone = __syncthreads_and(done && candTail == candHead) != 0;
        {
            unsigned int before = 0u, total = 0u;                // bytewise: at most 4 x 32 per cluster, no carries
#pragma unroll
            for (int w2 = 0; w2 < kBuildWarps; w2++) { const unsigned int c = cq.count[w2]; total += c; before += (w2 < warp) ? c : 0u; }
            const unsigned int mine = __vadd4(tailPos, before);  // ring positions (mod 256) of this warp's first entry per cluster
#pragma unroll
            for (int q = 0; q < kSubBlocks; q++) {
                const unsigned int byte = (colmask >> (kCluster * q)) & 0xffu;
                if (byte != 0u) cq.ring[q][(((mine >> (8 * q)) & 0xffu) + __popc(bal[q] & ltMask)) & (kCoopRing - 1)] = jref | (byte << 24);
            }
            tailPos = __vadd4(tailPos, total);
            avail += (int) ((total >> (8 * warp)) & 0xffu);
        }
        __syncthreads();
        // warp q: the full tiles of cluster q (at the very end: whatever is left, and close the open chunk)
        {
            const int q = warp;
            bool flushed = false;
            while (avail >= kTile || (allDone && !flushed && (avail | used) != 0)) {
                const int n = min(avail, kTile);
                flushed = n < kTile;
                if (used == 0 && n > 0) {                                    // open a chunk (= the next work item of this stream)
                    if (lane == 0) chunkBase = atomicAdd(&A.counters->tileTotal, (unsigned int) A.chunkTiles);
                    chunkBase = __shfl_sync(0xffffffffu, chunkBase, 0);
                    if (chunkBase > A.tileCap - (unsigned int) A.chunkTiles) {
                        if (lane == 0) atomicOr(&A.counters->overflow, 2u);
                        chunkBase = kNoChunk;
                    }
                }
                if (n > 0) {
                    const unsigned int word = (lane < n) ? cq.ring[q][(head + lane) & (kCoopRing - 1)] : kEmptySlot;
                    if (chunkBase != kNoChunk) A.tileDesc[(chunkBase + (unsigned int) used) * kTile + lane] = word;
                    used += 1;
                    head += (unsigned int) n;
                    avail -= n;
                }
                const bool close = used == A.chunkTiles || n < kTile;       // chunk full, or the (padded) last tile of the stream
                if (close && chunkBase != kNoChunk && lane == 0 && used > 0) {
                    const unsigned int pos = atomicAdd(&A.counters->itemCount, 1u);
                    if (pos < A.itemCap) {
                        WorkItem w;
                        w.block = kSubBlocks * b + q; w.image = set; w.tileStart = (int) chunkBase; w.tileCount = used;
                        A.items[pos] = w;
                    } else atomicOr(&A.counters->overflow, 4u);
                    atomicAdd(&A.counters->tilesUsed, (unsigned int) used);
                }
                if (close) used = 0;
            }
        }
        if (allDone) break;
    }
    for (int o = 16; o > 0; o >>= 1) myPairs += __shfl_xor_sync(0xffffffffu, myPairs, o);
    if (lane == 0 && myPairs) atomicAdd(&A.setPairs[set], myPairs);
    // ---- next overlapping image set of this unit
    if (!imageUnit) break;
    while (pending == 0u && setBase < A.nsets) {
        const int sc = setBase + lane;
        bool overlap = sc < A.nsets && ((sc - 1) & (A.imgParts - 1)) == ip;
        if (overlap) {
            const ImageBoxDev ib = A.imageBoxes[sc];
            for (int d = 0; d < 3; d++) overlap = overlap && (ib.lo[d] <= sbox[3 + d] + reach) && (ib.hi[d] >= sbox[d] - reach);
        }
        pending = __ballot_sync(0xffffffffu, overlap);
        setBase += kTile;
    }
    if (pending == 0u) break;
    set = setBase - kTile + __ffs(pending) - 1;
    pending &= pending - 1u;
    }
}


// ------------------------------------------------------------------------------------------------------
// explicit pair lists from the tiles: one warp per work item; pairs of set k land in [pairOffsets[k], ...)
// ------------------------------------------------------------------------------------------------------
__global__ void k_expand_pairs(const WorkItem *__restrict__ items, int nitems, const unsigned int *__restrict__ tileDesc,
                               const int *__restrict__ sAtom, int n, int rawJ, unsigned long long *cursor, int *__restrict__ pairs)
{
    const int lane = threadIdx.x & 31, m = lane & 7;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int it = w; it < nitems; it += nw) {
        const WorkItem wi = items[it];
        const int si = wi.block * kCluster + m;
        const int ai = (si < n) ? sAtom[si] : -1;
        for (int t = 0; t < wi.tileCount; t++) {
            const unsigned int d = tileDesc[((size_t) wi.tileStart + t) * kTile + lane];
            const unsigned int sj = d & kEmptySlot, col = (sj == kEmptySlot) ? 0u : (d >> 24);     // bit i <-> cluster atom i
            const int aj = (sj == kEmptySlot) ? -1 : ((rawJ && wi.image > 0) ? (int) sj : sAtom[sj]);
            const int cnt = __popc(col);
            int inc = cnt;
            for (int off = 1; off < 32; off <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += u; }
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            unsigned long long base = 0;
            if (lane == 0 && total) base = atomicAdd(&cursor[wi.image], (unsigned long long) total);
            base = __shfl_sync(0xffffffffu, base, 0);
            unsigned long long pos = base + (unsigned long long) (inc - cnt);
#pragma unroll
            for (int k = 0; k < kCluster; k++) {
                const int i = __shfl_sync(0xffffffffu, ai, k);               // lanes 0..7 hold the cluster atoms
                if ((col >> k) & 1u) { pairs[2 * pos] = i; pairs[2 * pos + 1] = aj; pos++; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// section 8e: which sorted positions do this rank's lists reference inside every rank's slab?  [min, max + 1) per slab:
// the halo ranges of the gradient / position exchange (the rank's own slab is always whole)
// ------------------------------------------------------------------------------------------------------
__global__ void k_touched_ranges(const WorkItem *__restrict__ items, int nitems, int chunkTiles, const unsigned int *__restrict__ tileDesc, int nblocks, int nranks, int *tab)
{
    extern __shared__ int shTab[];                            // [nranks][2 halves][min, max + 1]
    for (int k = threadIdx.x; k < 4 * nranks; k += blockDim.x) shTab[k] = (k & 1) ? 0 : 0x7fffffff;
    __syncthreads();
    // one THREAD per (work item, tile slot of its chunk): eight independent 128-bit loads per thread keep the memory system busy
    // (a warp per tile is latency bound here); the whole tile pool is read once
    const long nslots = (long) nitems * chunkTiles, nth = (long) gridDim.x * blockDim.x;
    for (long w = (long) blockIdx.x * blockDim.x + threadIdx.x; w < nslots; w += nth) {
        const int it = (int) (w / chunkTiles), t = (int) (w % chunkTiles);
        const WorkItem wi = items[it];
        if (t >= wi.tileCount) continue;
        const uint4 *tp = reinterpret_cast<const uint4 *>(tileDesc + ((size_t) wi.tileStart + t) * kTile);
        uint4 v[kTile / 4];
#pragma unroll
        for (int k = 0; k < kTile / 4; k++) v[k] = tp[k];
        unsigned int lo = kEmptySlot, hi = 0u;
#pragma unroll
        for (int k = 0; k < kTile / 4; k++) {
            const unsigned int e[4] = {v[k].x & kEmptySlot, v[k].y & kEmptySlot, v[k].z & kEmptySlot, v[k].w & kEmptySlot};
#pragma unroll
            for (int c = 0; c < 4; c++) { lo = min(lo, e[c]); hi = max(hi, e[c] == kEmptySlot ? 0u : e[c] + 1u); }
        }
        if (hi <= lo) continue;
        // a tile's j atoms are spatially compact: they fall into very few slabs; walk the slabs between min and max.
        // slab r holds blocks [nblocks r / R, nblocks (r + 1) / R); ranges that straddle slabs are clipped per slab
        const int bLo = (int) (lo >> 5), bHi = (int) ((hi - 1) >> 5);
        int r = (int) (((long) bLo * nranks) / nblocks);
        while (r > 0 && (int) (((long) nblocks * r) / nranks) > bLo) r--;
        while (r + 1 < nranks && (int) (((long) nblocks * (r + 1)) / nranks) <= bLo) r++;
        for (; r < nranks; r++) {
            const int s0 = (int) (((long) nblocks * r) / nranks) * kTile, s1 = (int) (((long) nblocks * (r + 1)) / nranks) * kTile;
            if (s0 > ((bHi + 1) << 5)) break;
            // two ranges per slab (its lower and upper half): with periodic images a rank reaches both ends of a neighbour's slab
            const int mid = (s0 + s1) >> 1;
            const int a0 = max((int) lo, s0), b0 = min((int) hi, mid), a1 = max((int) lo, mid), b1 = min((int) hi, s1);
            if (b0 > a0) { if (a0 < shTab[4 * r]) atomicMin(&shTab[4 * r], a0); if (b0 > shTab[4 * r + 1]) atomicMax(&shTab[4 * r + 1], b0); }
            if (b1 > a1) { if (a1 < shTab[4 * r + 2]) atomicMin(&shTab[4 * r + 2], a1); if (b1 > shTab[4 * r + 3]) atomicMax(&shTab[4 * r + 3], b1); }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * nranks; k += blockDim.x) {
        if (k & 1) { if (shTab[k] > 0) atomicMax(&tab[k], shTab[k]); }
        else if (shTab[k] != 0x7fffffff) atomicMin(&tab[k], shTab[k]);
    }
}

// 1-4 partners of owned atoms are excluded from the lists, but their gradients travel the same way
__global__ void k_touched_14(const int2 *__restrict__ pairs, int npairs, const int *__restrict__ invPerm, int ownLo, int ownHi, int nblocks, int nranks, int *tab)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int si = invPerm[pairs[p].x], sj = invPerm[pairs[p].y];
    if (si < ownLo || si >= ownHi) return;
    const int bj = sj >> 5;
    int r = (int) (((long) bj * nranks) / nblocks);
    while (r > 0 && (int) (((long) nblocks * r) / nranks) > bj) r--;
    while (r + 1 < nranks && (int) (((long) nblocks * (r + 1)) / nranks) <= bj) r++;
    const int s0 = (int) (((long) nblocks * r) / nranks) * kTile, s1 = (int) (((long) nblocks * (r + 1)) / nranks) * kTile;
    const int h = (sj >= ((s0 + s1) >> 1)) ? 1 : 0;
    atomicMin(&tab[4 * r + 2 * h], sj); atomicMax(&tab[4 * r + 2 * h + 1], sj + 1);
}

static __global__ void k_ranges_finish(const int *__restrict__ tab, int count, int n, int ownRank, long *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (slab, half)
    if (k >= count) return;
    long lo = tab[2 * k], hi = tab[2 * k + 1];
    if (hi <= lo || (k >> 1) == ownRank) { lo = 0; hi = 0; }  // the own slab is never exchanged
    out[2 * k] = lo; out[2 * k + 1] = hi < n ? hi : n;
}

// enqueue the range computation; the table [nranks][2][2] (long) lands in d_out (device) without any host synchronisation
bool touched_ranges_async(State &s, long *d_out)
{
    const int R = s.nranks;
    std::vector<int> init(4 * (size_t) R);
    for (int k = 0; k < 2 * R; k++) { init[2 * k] = 0x7fffffff; init[2 * k + 1] = 0; }
    if (!s.rangeTab.ensure(4 * (size_t) R)) return false;
    // pinned staging (hsmall is free between an update and the next energy call)
    int *stage = reinterpret_cast<int *>(s.hsmall);
    std::memcpy(stage, init.data(), sizeof(int) * 4 * R);
    NBB_CUDA(cudaMemcpyAsync(s.rangeTab.p, stage, sizeof(int) * 4 * R, cudaMemcpyHostToDevice, s.stream));
    const int nitems = (int) s.hostCounters.itemCount;
    if (nitems > 0) {
        const int threads = 256;
        const long slots = (long) nitems * s.chunkTiles;
        const long blocks = std::min<long>(148 * 8, (slots + threads - 1) / threads);
        k_touched_ranges<<<(unsigned int) blocks, threads, sizeof(int) * 4 * R, s.stream>>>(s.items.p, nitems, s.chunkTiles, s.tileDesc.p, s.nblocks, R, s.rangeTab.p);
        s.launches += 1;
    }
    if (s.n14 > 0) {
        k_touched_14<<<(s.n14 + 255) / 256, 256, 0, s.stream>>>(s.pairs14.p, s.n14, s.invPerm.p, s.ownLo, s.ownHi, s.nblocks, R, s.rangeTab.p);
        s.launches += 1;
    }
    k_ranges_finish<<<(2 * R + 63) / 64, 64, 0, s.stream>>>(s.rangeTab.p, 2 * R, s.n, s.rank, d_out);
    s.launches += 1;
    return cuda_ok(cudaGetLastError(), "touched ranges");
}

bool touched_ranges(State &s, long *out)
{
    const int R = s.nranks;
    if (!s.rangeOut.ensure(4 * (size_t) R)) return false;
    if (!touched_ranges_async(s, s.rangeOut.p)) return false;
    std::vector<long> tab(4 * (size_t) R);
    NBB_CUDA(cudaMemcpyAsync(tab.data(), s.rangeOut.p, sizeof(long) * 4 * R, cudaMemcpyDeviceToHost, s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    for (int k = 0; k < 4 * R; k++) out[k] = tab[k];
    // (the synchronous variant reports the own slab's ranges as empty, too)
    return true;
}

bool expand_pairs(State &s)
{
    if (s.pairsExpanded) return true;
    const int nsets = s.nsets;
    // pair counts per set
    std::vector<unsigned long long> cnt(nsets);
    NBB_CUDA(cudaMemcpyAsync(cnt.data(), s.setPairs.p, sizeof(unsigned long long) * nsets, cudaMemcpyDeviceToHost, s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    s.pairOffsets.assign(nsets + 1, 0);
    for (int k = 0; k < nsets; k++) s.pairOffsets[k + 1] = s.pairOffsets[k] + cnt[k];
    const unsigned long long total = s.pairOffsets[nsets];
    if (!s.pairBuf.ensure((size_t) 2 * total + 2)) return false;
    if (!s.pairCursor.ensure(nsets)) return false;
    NBB_CUDA(cudaMemcpyAsync(s.pairCursor.p, s.pairOffsets.data(), sizeof(unsigned long long) * nsets, cudaMemcpyHostToDevice, s.stream));
    const int nitems = (int) s.hostCounters.itemCount;
    if (nitems > 0) {
        if (s.timing) cudaEventRecord(s.ev[8], s.stream);
        const int threads = 256, nblk = std::max(1, std::min(148 * 8, (nitems + 7) / 8));
        k_expand_pairs<<<nblk, threads, 0, s.stream>>>(s.items.p, nitems, s.tileDesc.p, s.sAtom.p, s.n, s.rawJ ? 1 : 0, s.pairCursor.p, s.pairBuf.p);
        s.launches += 1;
        if (s.timing) cudaEventRecord(s.ev[9], s.stream);
    }
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    if (s.timing && nitems > 0) { float ms = 0; cudaEventElapsedTime(&ms, s.ev[8], s.ev[9]); s.timings[4] = ms; }
    s.pairsExpanded = true;
    return true;
}

// ------------------------------------------------------------------------------------------------------
// host orchestration of one rebuild
// ------------------------------------------------------------------------------------------------------
static bool exclusive_scan(State &s, unsigned int *counts, unsigned int *out, int n)
{
    const int nchunks = (n + kScanChunk - 1) / kScanChunk;
    if (!s.scanTmp.ensure((size_t) nchunks + 2)) return false;
    unsigned int *sums = s.scanTmp.p, *grand = s.scanTmp.p + nchunks;
    k_scan_chunks<<<nchunks, kScanThreads, 0, s.stream>>>(counts, out, sums, n);
    k_scan_sums<<<1, kScanThreads, 0, s.stream>>>(sums, nchunks, grand);
    k_scan_add<<<(n + 1 + 255) / 256, 256, 0, s.stream>>>(out, sums, n, grand);
    s.launches += 3;
    return true;
}

static void setup_grid(State &s, const double *lo, const double *hi)
{
    BuildGrid &g = s.grid;
    g.h = 0.5 * s.list;
    // keep the number of keys bounded for very large / very sparse systems
    for (;;) {
        long cells = 1;
        for (int d = 0; d < 3; d++) { g.dim[d] = std::max(1, (int) std::floor((hi[d] - lo[d]) / g.h) + 1); cells *= g.dim[d]; }
        if (cells * (long) s.nsets <= 6000000L) break;
        g.h *= 1.26;
    }
    g.invh = 1.0 / g.h;
    for (int d = 0; d < 3; d++) g.lo[d] = lo[d];
    g.ncell = g.dim[0] * g.dim[1] * g.dim[2];
}

// sort the extended atoms, build blocks, tiles and work items.  On entry the extended atoms (eX, eAtom, eSet, eKey,
// eSort, cell counts in cellFill) are on the device and counters->extCount holds the number of non-primary entries.
static bool sort_and_tile(State &s, bool selfEnabled, unsigned int extUpperBound)
{
    const int nkeys = s.nsets * s.grid.ncell;
    const size_t neMax = (size_t) s.n + extUpperBound;
    if (!s.cellStart.ensure((size_t) nkeys + 2) || !s.order.ensure(neMax) || !s.sX.ensure(3 * neMax) || !s.sAtom.ensure(neMax) || !s.invPerm.ensure((size_t) s.n)) return false;
    if (!exclusive_scan(s, s.cellFill.p, s.cellStart.p, nkeys)) return false;
    NBB_CUDA(cudaMemsetAsync(s.cellFill.p, 0, sizeof(unsigned int) * nkeys, s.stream));
    // the number of extended atoms actually appended stays on the device (no host synchronisation here); an overflow of the extended
    // capacity is noticed with the counters that come back after the tile builder
    s.nblocks = (s.n + kTile - 1) / kTile;
    const int b0 = (int) (((long) s.nblocks * s.rank) / s.nranks), b1 = (int) (((long) s.nblocks * (s.rank + 1)) / s.nranks);
    const int myBlocks = b1 - b0;
    s.ownLo = b0 * kTile; s.ownHi = std::min(s.n, b1 * kTile);
    // several ranks: only the cells this rank's slab can see are scattered, sorted and (later) packed; the standalone generators and a
    // single rank sort everything
    const unsigned char *need = nullptr;
    const int ncell = s.grid.ncell;
    if (s.restrictSort && s.nranks > 1 && selfEnabled && !s.rawJ) {
        if (!s.cellNeed.ensure((size_t) 2 * ncell)) return false;
        k_mark_needed<<<(ncell + 127) / 128, 128, 0, s.stream>>>(s.cellStart.p, s.grid, (unsigned int) s.ownLo, (unsigned int) s.ownHi, s.cellNeed.p);
        if (extUpperBound > 0)
            k_mark_image_sources<<<(extUpperBound + 255) / 256, 256, 0, s.stream>>>(s.eKey.p, s.eSet.p, s.eAtom.p, s.n, extUpperBound, s.counters, ncell, s.cellNeed.p);
        NBB_CUDA(cudaMemsetAsync(s.invPerm.p, 0xff, sizeof(int) * (size_t) s.n, s.stream));      // atoms of unseen cells: no sorted position (-1)
        s.launches += 2;
        need = s.cellNeed.p;
    }
    k_scatter<<<(unsigned int) ((neMax + 255) / 256), 256, 0, s.stream>>>(s.eKey.p, s.eSet.p, s.n, extUpperBound, s.counters, s.cellStart.p, s.cellFill.p, s.order.p, need, ncell);
    k_sort_cells<<<(nkeys + 7) / 8, 256, 0, s.stream>>>(s.cellStart.p, nkeys, s.eSortBuf.p, s.order.p, s.eX.p, s.eAtom.p, s.eSet.p, s.sX.p, s.sAtom.p, s.invPerm.p, need, ncell);
    if (!s.blockBox.ensure((size_t) 9 * s.nblocks)) return false;
    // (the builder reads the own blocks' boxes only; the grouping pass writes them when it runs)
    if (s.typeFree.p != nullptr) { k_group_lj_free<<<(s.n + 255) / 256, 256, 0, s.stream>>>(s.n, s.ljtype.p, s.typeFree.p, s.grid, s.sX.p, s.sAtom.p, s.invPerm.p, b0, myBlocks, s.blockBox.p); s.launches += 1; }
    else if (myBlocks > 0) { k_block_boxes<<<(myBlocks * 32 + 255) / 256, 256, 0, s.stream>>>(s.sX.p, s.n, b0, myBlocks, s.blockBox.p); s.launches += 1; }
    s.launches += 2;

    // tiles: a global pool handed out in chunks (= work items); sized from the pair density, retried once with the exact need
    // tiles per chunk / work item: long items amortise the per-item prologue of the force kernel, short ones keep small systems spread over all SMs
    static const int chunkOverride = []() { const char *e = std::getenv("NBB200_CHUNK"); return e ? std::atoi(e) : 0; }();
    const int chunk = chunkOverride > 0 ? chunkOverride : ((s.n >= 400000) ? 32 : (s.n >= 60000 ? 16 : 8));
    if (chunk != s.chunkTiles) { s.chunkTiles = chunk; s.tileCap = 0; }
    // small systems: deal the rows of a block to several warps until the GPU is full (each part has its own, padded, j streams)
    int split = 1;
    static const long splitEnv = []() { const char *e = std::getenv("NBB200_SPLIT_TARGET"); return (e && std::atol(e) > 0) ? std::atol(e) : 0L; }();
    // dynamics rebuilds every ten steps or more (nbb200_set_list_reuse_hint): whole blocks per warp -- fewer padded tiles, the force kernel
    // of every step gains more (DHFR 73 -> 68 us) than the rarer rebuild loses (0.26 -> 0.36 ms); single calls keep the faster rebuild
    const long splitTarget = splitEnv > 0 ? splitEnv : (s.listReuseHint ? 148L * 8 : 148L * 16);
    while (split < 8 && (long) myBlocks * split * 2 <= splitTarget) split *= 2;
    // cooperative builder (k_build_tiles_coop): four warps per block with shared j streams, when one warp per block leaves the GPU short of
    // warps (148 SMs x 28 builder warps).  NBB200_COOP=0 / 1 forces it off / on
    static const int coopEnv = []() { const char *e = std::getenv("NBB200_COOP"); return e ? std::atoi(e) : -1; }();
    static const long coopBlocks = []() { const char *e = std::getenv("NBB200_COOP_BLOCKS"); return (e && std::atol(e) > 0) ? std::atol(e) : 148L * 28; }();
    const bool coop = coopEnv >= 0 ? coopEnv != 0 : (long) myBlocks <= coopBlocks;
    if (coop) split = kBuildWarps;
    size_t cap = s.tileCap;
    if (cap == 0) {
        // expected list pairs from the mean density inside the search box; 8 x 32 tiles are about half full
        double vol = 1.0;
        for (int d = 0; d < 3; d++) vol *= std::max(1.0, s.plan.upper[d] - s.plan.lower[d] - 2.0 * s.list);
        const double density = std::min(0.2, (double) s.n / vol);
        const double pairsPerAtom = 0.5 * density * (4.0 / 3.0) * 3.14159265358979 * s.list * s.list * s.list + 8.0;
        const double tiles = pairsPerAtom * ((double) s.n * myBlocks / std::max(1, s.nblocks)) / (kCluster * kTile * 0.5);
        const double streams = (double) kSubBlocks * myBlocks * std::min(s.nsets, 6) * (coop ? 1 : split);
        cap = (size_t) (1.2 * tiles + streams * chunk) + 1024;
    }
    for (int attempt = 0; attempt < 4; attempt++) {
        if (cap * kTile >= ((size_t) 1 << 31)) { set_error("tile pool exceeds 2^31 descriptor words"); return false; }
        cap = std::max(cap, (size_t) chunk);
        s.tileCap = cap;
        s.itemCap = cap / chunk + 64;
        if (!s.tileDesc.ensure(cap * kTile) || !s.items.ensure(s.itemCap) || !s.setPairs.ensure((size_t) s.nsets)) return false;
        NBB_CUDA(cudaMemsetAsync(s.setPairs.p, 0, sizeof(unsigned long long) * s.nsets, s.stream));
        NBB_CUDA(cudaMemsetAsync(&s.counters->itemCount, 0, sizeof(unsigned int) * 4, s.stream));   // itemCount, tileTotal, tilesUsed, overflow
        if (myBlocks > 0) {
            TileArgs A;
            A.n = s.n; A.nblocks = s.nblocks; A.nsets = s.nsets; A.firstBlock = b0; A.myBlocks = myBlocks; A.selfEnabled = selfEnabled ? 1 : 0;
            A.chunkTiles = chunk; A.rawJ = s.rawJ ? 1 : 0;
            A.split = split; A.splitShift = 0;
            while ((1 << A.splitShift) < split) A.splitShift++;
            // image sets of a block: one warp walks them all on large systems, up to 32 warps share them when the GPU would stay empty
            static const int imgOverride = []() { const char *e = std::getenv("NBB200_IMG_PARTS"); return e ? std::atoi(e) : 0; }();
            int imgParts = 1;
            while (imgParts < 32 && imgParts < s.nsets - 1 && (long) myBlocks * split * imgParts < 4 * splitTarget) imgParts *= 2;
            if (imgOverride > 0) { imgParts = 1; while (imgParts < imgOverride && imgParts < 32) imgParts *= 2; }
            A.imgParts = imgParts; A.imgShift = 0;
            while ((1 << A.imgShift) < imgParts) A.imgShift++;
            A.primaryWarps = (unsigned int) ((long) myBlocks * split);
            A.totalWarps = A.primaryWarps + (s.nsets > 1 ? A.primaryWarps * (unsigned int) imgParts : 0u);
            A.cutoff = s.list; A.cutoff2 = s.list * s.list;
            A.grid = s.grid;
            A.sX = s.sX.p; A.sAtom = s.sAtom.p; A.invPerm = s.invPerm.p; A.cellStart = s.cellStart.p; A.blockBox = s.blockBox.p;
            A.imageBoxes = s.imageBoxes.p; A.exclPtr = s.exclPtr.p; A.exclCol = s.exclCol.p;
            A.fixed = s.nfixed > 0 ? s.fixedFlag.p : nullptr;
            A.inactive = s.nqc > 0 ? s.qcFlag.p : nullptr;
            A.tileDesc = s.tileDesc.p; A.tileCap = (unsigned int) cap;
            A.items = s.items.p; A.itemCap = (unsigned int) s.itemCap; A.setPairs = s.setPairs.p; A.counters = s.counters;
            const long warps = (long) A.totalWarps;
            const unsigned int ctas = (unsigned int) ((warps + kBuildWarps - 1) / kBuildWarps);
            if (coop) {
                if (A.inactive != nullptr) k_build_tiles_coop<true><<<ctas, kBuildThreads, 0, s.stream>>>(A);
                else k_build_tiles_coop<false><<<ctas, kBuildThreads, 0, s.stream>>>(A);
            } else if (A.inactive != nullptr) k_build_tiles<true><<<ctas, kBuildThreads, 0, s.stream>>>(A);
            else k_build_tiles<false><<<ctas, kBuildThreads, 0, s.stream>>>(A);
            s.launches += 1;
        }
        NBB_CUDA(cudaMemcpyAsync(&s.hostCounters, s.counters, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, s.stream));
        NBB_CUDA(cudaStreamSynchronize(s.stream));
        if (!cuda_ok(cudaGetLastError(), "k_build_tiles")) return false;
        if (s.hostCounters.overflow & 1u) { set_error("extended atom capacity exceeded"); return false; }      // build_lists retries with the exact bound
        if ((s.hostCounters.overflow & 6u) == 0u) return true;
        cap = (size_t) (1.15 * (double) s.hostCounters.tileTotal) + 1024;       // the cursor kept counting: this is the exact need
    }
    set_error("tile capacity exceeded after retries");
    return false;
}

bool build_lists(State &s)
{
    const int n = s.n;
    const int ntr = s.trans.n;
    // 1. base operations and bounding boxes -> image plan (host logic, symmetry_host.cpp)
    std::vector<double> ops((size_t) 12 * std::max(1, ntr));
    std::vector<RealSpaceOp> base(ntr);
    for (int t = 0; t < ntr; t++) {
        base[t] = orthogonalize(s.trans.rot[t], &s.trans.trans[3 * t], s.lattice);
        std::memcpy(&ops[12 * t], base[t].R.v, sizeof(double) * 9);
        std::memcpy(&ops[12 * t + 9], base[t].tv, sizeof(double) * 3);
    }
    if (!s.baseOpsDev.ensure(ops.size())) return false;
    NBB_CUDA(cudaMemcpyAsync(s.baseOpsDev.p, ops.data(), sizeof(double) * ops.size(), cudaMemcpyHostToDevice, s.stream));
    std::vector<double> bmin(3 * (1 + ntr)), bext(3 * (1 + ntr));
    if (!device_bbox(s, 1 + ntr, bmin.data(), bext.data())) return false;
    for (double v : bmin) if (!std::isfinite(v)) { set_error("non-finite coordinates"); return false; }
    for (double v : bext) if (!std::isfinite(v)) { set_error("non-finite coordinates"); return false; }
    if (ntr > 0) {
        if (!plan_images(s.trans, s.lattice, s.list, s.checkForInverses, s.expandFactor, bmin.data(), bext.data(), s.plan)) {
            set_error("image search range is absurd (more than 2^20 lattice translations): coordinates or lattice are not physical");
            return false;
        }
    } else {
        s.plan = ImagePlan();
        for (int d = 0; d < 3; d++) { s.plan.lower[d] = bmin[d] - s.list; s.plan.upper[d] = (bext[d] + bmin[d]) + s.list; }
    }
    const int nimg = (int) s.plan.images.size(), nvis = (int) s.plan.visits.size();
    s.nsets = 1 + nimg;
    s.rawJ = false;
    s.imagePairs.assign(nimg, -1); s.primaryPairs = -1; s.pairCountsValid = false; s.pairsExpanded = false;

    // 2. grid over the search box, per-image list-time boxes, visits
    const double margin = 1.0e-6;
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = s.plan.lower[d] - margin; hi[d] = s.plan.upper[d] + margin; }
    setup_grid(s, lo, hi);
    std::vector<ImageBoxDev> boxes(s.nsets);
    std::memset(boxes.data(), 0, sizeof(ImageBoxDev) * boxes.size());
    for (int k = 0; k < nimg; k++) for (int d = 0; d < 3; d++) { boxes[1 + k].lo[d] = s.plan.images[k].lo[d]; boxes[1 + k].hi[d] = s.plan.images[k].hi[d]; }
    std::vector<double> vdisp((size_t) 3 * std::max(1, nvis));
    std::vector<int> vinfo((size_t) 2 * std::max(1, nvis));
    for (int v = 0; v < nvis; v++) {
        for (int d = 0; d < 3; d++) vdisp[3 * v + d] = s.plan.visits[v].disp[d];
        vinfo[2 * v] = s.plan.visits[v].t; vinfo[2 * v + 1] = s.plan.visits[v].image;
    }
    if (!s.imageBoxes.ensure(boxes.size()) || !s.visitDisp.ensure(vdisp.size()) || !s.visitInfo.ensure(vinfo.size())) return false;
    NBB_CUDA(cudaMemcpyAsync(s.imageBoxes.p, boxes.data(), sizeof(ImageBoxDev) * boxes.size(), cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaMemcpyAsync(s.visitDisp.p, vdisp.data(), sizeof(double) * vdisp.size(), cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaMemcpyAsync(s.visitInfo.p, vinfo.data(), sizeof(int) * vinfo.size(), cudaMemcpyHostToDevice, s.stream));

    // 3. extended-atom capacity: image atoms inside the search box.  Upper bound from the box overlap volumes.
    double need = 0.0;
    for (int k = 0; k < nimg; k++) {
        double frac = 1.0;
        for (int d = 0; d < 3; d++) {
            const double il = s.plan.images[k].lo[d], iu = s.plan.images[k].hi[d];
            const double ov = std::min(iu, hi[d]) - std::max(il, lo[d]);
            const double len = std::max(iu - il, 1.0e-9);
            frac *= std::min(1.0, std::max(0.0, ov + 2.0) / len);           // +2 A slack for density fluctuations
        }
        need += frac * n;
    }
    unsigned int extCap = (unsigned int) std::min((double) nimg * n, need * 1.25 + 1024.0);
    if (n <= 65536) extCap = (unsigned int) ((size_t) nimg * n);
    for (int attempt = 0; attempt < 2; attempt++) {
        const size_t neMax = (size_t) n + extCap;
        const int nkeys = s.nsets * s.grid.ncell;
        if (!s.eX.ensure(3 * neMax) || !s.eAtom.ensure(neMax) || !s.eSet.ensure(neMax) || !s.eKey.ensure(neMax) || !s.eSortBuf.ensure(neMax) ||
            !s.cellFill.ensure((size_t) nkeys + 2)) return false;
        NBB_CUDA(cudaMemsetAsync(s.cellFill.p, 0, sizeof(unsigned int) * ((size_t) nkeys + 2), s.stream));
        NBB_CUDA(cudaMemsetAsync(s.counters, 0, sizeof(DeviceCounters), s.stream));
        ExtendArgs E;
        E.x = s.xcur; E.n = n; E.baseOps = s.baseOpsDev.p; E.visitDisp = s.visitDisp.p; E.visitInfo = s.visitInfo.p; E.nvisits = nvis;
        for (int d = 0; d < 3; d++) { E.boxLo[d] = lo[d]; E.boxHi[d] = hi[d]; }
        E.grid = s.grid;
        E.eX = s.eX.p; E.eAtom = s.eAtom.p; E.eSet = s.eSet.p; E.eKey = s.eKey.p; E.eSort = s.eSortBuf.p;
        E.cellCount = s.cellFill.p; E.extCap = extCap; E.counters = s.counters;
        k_extend<<<(n + 127) / 128, 128, 0, s.stream>>>(E);
        s.launches += 1;
        s.extCap = extCap;
        if (sort_and_tile(s, true, extCap)) return true;
        if (!(s.hostCounters.overflow & 1u)) return false;
        extCap = (unsigned int) ((size_t) nimg * n);                       // exact worst case, retry once
    }
    return false;
}

// stand-alone generators (PairListGenerator_* entry points): sets are given coordinate arrays, no symmetry
bool build_lists_standalone(State &s, const double *d_x2, int n2)
{
    const int n = s.n;
    std::vector<double> bmin(3), bext(3);
    if (!s.baseOpsDev.ensure(12)) return false;
    if (!device_bbox(s, 1, bmin.data(), bext.data())) return false;
    s.plan = ImagePlan();
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = bmin[d] - s.list - 1.0e-6; hi[d] = (bext[d] + bmin[d]) + s.list + 1.0e-6; s.plan.lower[d] = lo[d]; s.plan.upper[d] = hi[d]; }
    s.nsets = d_x2 ? 2 : 1;
    s.rawJ = d_x2 != nullptr;
    if (n2 > kMaxAtoms) { set_error("more than 16.7 M points in the second array"); return false; }
    s.imagePairs.assign(s.nsets - 1, -1); s.primaryPairs = -1; s.pairCountsValid = false; s.pairsExpanded = false;
    setup_grid(s, lo, hi);
    std::vector<ImageBoxDev> boxes(s.nsets);
    for (auto &bx : boxes) for (int d = 0; d < 3; d++) { bx.lo[d] = -1.0e300; bx.hi[d] = 1.0e300; }
    if (!s.imageBoxes.ensure(boxes.size())) return false;
    NBB_CUDA(cudaMemcpyAsync(s.imageBoxes.p, boxes.data(), sizeof(ImageBoxDev) * boxes.size(), cudaMemcpyHostToDevice, s.stream));
    const unsigned int extCap = d_x2 ? (unsigned int) n2 : 0u;
    const size_t neMax = (size_t) n + extCap;
    const int nkeys = s.nsets * s.grid.ncell;
    if (!s.eX.ensure(3 * neMax) || !s.eAtom.ensure(neMax) || !s.eSet.ensure(neMax) || !s.eKey.ensure(neMax) || !s.eSortBuf.ensure(neMax) ||
        !s.cellFill.ensure((size_t) nkeys + 2) || !s.visitDisp.ensure(3) || !s.visitInfo.ensure(2)) return false;
    NBB_CUDA(cudaMemsetAsync(s.cellFill.p, 0, sizeof(unsigned int) * ((size_t) nkeys + 2), s.stream));
    NBB_CUDA(cudaMemsetAsync(s.counters, 0, sizeof(DeviceCounters), s.stream));
    ExtendArgs E;
    E.x = s.xcur; E.n = n; E.baseOps = s.baseOpsDev.p; E.visitDisp = s.visitDisp.p; E.visitInfo = s.visitInfo.p; E.nvisits = 0;
    for (int d = 0; d < 3; d++) { E.boxLo[d] = lo[d]; E.boxHi[d] = hi[d]; }
    E.grid = s.grid;
    E.eX = s.eX.p; E.eAtom = s.eAtom.p; E.eSet = s.eSet.p; E.eKey = s.eKey.p; E.eSort = s.eSortBuf.p;
    E.cellCount = s.cellFill.p; E.extCap = extCap; E.counters = s.counters;
    k_extend<<<(n + 127) / 128, 128, 0, s.stream>>>(E);
    s.launches += 1;
    if (d_x2) {
        k_extend_cross<<<(n2 + 127) / 128, 128, 0, s.stream>>>(d_x2, n2, n, s.grid, make_double3(lo[0], lo[1], lo[2]), make_double3(hi[0], hi[1], hi[2]), s.eX.p, s.eAtom.p, s.eSet.p, s.eKey.p, s.eSortBuf.p, s.cellFill.p, s.counters);
        s.launches += 1;
    }
    s.extCap = extCap;
    return sort_and_tile(s, d_x2 == nullptr, extCap);
}

}  // namespace nbb200
