// mm_terms.cu -- the bonded MM terms and the Langevin half step on the device (what an MD caller needs next to the NB
// term so that coordinates, velocities and gradients never leave the GPU).
//
// Replaces, for device-resident coordinates / gradients,
//   HarmonicBondContainer_Energy      pMolecule-1.9.0/extensions/csource/HarmonicBondContainer.c:149-183   (bonds and Urey-Bradley terms)
//   HarmonicAngleContainer_Energy     pMolecule-1.9.0/extensions/csource/HarmonicAngleContainer.c:157-205
//   FourierDihedralContainer_Energy   pMolecule-1.9.0/extensions/csource/FourierDihedralContainer.c:156-241
//   HarmonicImproperContainer_Energy  pMolecule-1.9.0/extensions/csource/HarmonicImproperContainer.c:172-258
// as System.Energy calls them one after the other (pMolecule-1.9.0/pMolecule/System.py:272-318), and one Iteration of
// pCore-1.9.0/pCore/LangevinVelocityVerletIntegrator.py:117-150 in Cartesian variables.
//
// All terms of all containers are evaluated by ONE launch: a term is one thread (fp64 throughout; a few 10^4 terms), the term kinds are
// laid out one after the other so that warps are uniform except at the four seams; energies are reduced per kind through shared
// memory, gradients are accumulated with fp64 atomics.  Parameters are expanded per term when a container is defined.
#include "../../include/nbabfs_b200.h"
#include "nbb200_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace nbb200 {

constexpr int kKinds = 5;                 // 0 bond, 1 angle, 2 Urey-Bradley, 3 Fourier dihedral, 4 harmonic improper

struct TermRecord { int atom[4]; double p[4]; };    // p: bond / UB / angle {eq, fc}; dihedral {fc, cosphase, sinphase, period}; improper {fc, coseq, sineq}

struct MMTerms {
    int device = 0, natoms = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    std::vector<TermRecord> host[kKinds];
    bool dirty = true, slotsZeroed = false;
    int start[kKinds + 1] = {};
    DevBuf<TermRecord> rec;
    DevBuf<double> x, grad, energies;
    double *hx = nullptr, *hg = nullptr, *he = nullptr;
    long launches = 0;
};

struct KindStarts { int s[kKinds + 1]; };

__device__ __forceinline__ void add3(double *g, int a, double x, double y, double z)
{
    atomicAdd(g + 3 * (size_t) a, x); atomicAdd(g + 3 * (size_t) a + 1, y); atomicAdd(g + 3 * (size_t) a + 2, z);
}

// cos / sin of a dihedral i-j-k-l and the vectors its derivatives need (shared by the Fourier dihedral and the harmonic improper)
struct Torsion { double cosphi, sinphi, rkj, rkj2, m2, n2, mx, my, mz, nx, ny, nz, xij, yij, zij, xkj, ykj, zkj, xlk, ylk, zlk; };

__device__ __forceinline__ Torsion torsion(const double *__restrict__ x, int i, int j, int k, int l)
{
    Torsion t;
    t.xij = x[3 * i] - x[3 * j]; t.yij = x[3 * i + 1] - x[3 * j + 1]; t.zij = x[3 * i + 2] - x[3 * j + 2];
    t.xkj = x[3 * k] - x[3 * j]; t.ykj = x[3 * k + 1] - x[3 * j + 1]; t.zkj = x[3 * k + 2] - x[3 * j + 2];
    t.xlk = x[3 * l] - x[3 * k]; t.ylk = x[3 * l + 1] - x[3 * k + 1]; t.zlk = x[3 * l + 2] - x[3 * k + 2];
    t.rkj2 = t.xkj * t.xkj + t.ykj * t.ykj + t.zkj * t.zkj;
    t.rkj = sqrt(t.rkj2);
    t.mx = t.yij * t.zkj - t.zij * t.ykj; t.my = t.zij * t.xkj - t.xij * t.zkj; t.mz = t.xij * t.ykj - t.yij * t.xkj;
    t.nx = t.ylk * t.zkj - t.zlk * t.ykj; t.ny = t.zlk * t.xkj - t.xlk * t.zkj; t.nz = t.xlk * t.ykj - t.ylk * t.xkj;
    t.m2 = t.mx * t.mx + t.my * t.my + t.mz * t.mz;
    t.n2 = t.nx * t.nx + t.ny * t.ny + t.nz * t.nz;
    const double mn = sqrt(t.m2 * t.n2);
    t.cosphi = (t.mx * t.nx + t.my * t.ny + t.mz * t.nz) / mn;
    t.sinphi = t.rkj * (t.xij * t.nx + t.yij * t.ny + t.zij * t.nz) / mn;
    return t;
}

// gradient of a torsion term with dE/dphi = df (FourierDihedralContainer.c:212-235 = HarmonicImproperContainer.c:229-252)
__device__ __forceinline__ void torsion_gradient(const Torsion &t, double df, int i, int j, int k, int l, double *g)
{
    const double dtxi = df * t.rkj * t.mx / t.m2, dtyi = df * t.rkj * t.my / t.m2, dtzi = df * t.rkj * t.mz / t.m2;
    const double dtxl = -df * t.rkj * t.nx / t.n2, dtyl = -df * t.rkj * t.ny / t.n2, dtzl = -df * t.rkj * t.nz / t.n2;
    const double dotij = t.xij * t.xkj + t.yij * t.ykj + t.zij * t.zkj, dotlk = t.xlk * t.xkj + t.ylk * t.ykj + t.zlk * t.zkj;
    const double sx = (dotij * dtxi + dotlk * dtxl) / t.rkj2, sy = (dotij * dtyi + dotlk * dtyl) / t.rkj2, sz = (dotij * dtzi + dotlk * dtzl) / t.rkj2;
    add3(g, i, dtxi, dtyi, dtzi);
    add3(g, j, sx - dtxi, sy - dtyi, sz - dtzi);
    add3(g, k, -sx - dtxl, -sy - dtyl, -sz - dtzl);
    add3(g, l, dtxl, dtyl, dtzl);
}

// perm (nbb200_md_run): gradients go to row perm[atom] of g -- the NB state's sorted-order accumulator, so that the bonded terms can run next
// to the NB kernels of the step instead of behind its unsort pass
__global__ void k_mm_terms(const TermRecord *__restrict__ rec, KindStarts K, const double *__restrict__ x, double *g, double *energies, double *zeroOther,
                           const int *__restrict__ perm)
{
    if (zeroOther != nullptr && blockIdx.x == 0 && threadIdx.x < kKinds) zeroOther[threadIdx.x] = 0.0;   // two-slot use: prepares the next step's accumulators
    __shared__ double sE[kKinds];
    if (threadIdx.x < kKinds) sE[threadIdx.x] = 0.0;
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < K.s[kKinds]) {
        int kind = 0;
        while (n >= K.s[kind + 1]) kind++;
        const TermRecord r = rec[n];
        const int i = r.atom[0], j = r.atom[1], k = r.atom[2], l = r.atom[3];
        int gi = i, gj = j, gk = k, gl = l;                   // gradient rows
        if (perm != nullptr) { gi = perm[i]; gj = perm[j]; gk = perm[k]; gl = perm[l]; }
        double e = 0.0;
        if (kind == 0 || kind == 2) {                         // harmonic bond / Urey-Bradley: E = fc (r - eq)^2
            double xij = x[3 * i] - x[3 * j], yij = x[3 * i + 1] - x[3 * j + 1], zij = x[3 * i + 2] - x[3 * j + 2];
            const double rij = sqrt(xij * xij + yij * yij + zij * zij), disp = rij - r.p[0];
            double df = r.p[1] * disp;
            e = df * disp;
            if (g != nullptr) {
                df *= (2.0 / rij);
                xij *= df; yij *= df; zij *= df;
                add3(g, gi, xij, yij, zij); add3(g, gj, -xij, -yij, -zij);
            }
        } else if (kind == 1) {                               // harmonic angle i-j-k: E = fc (theta - eq)^2, |cos| clamped at 0.999999
            double xij = x[3 * i] - x[3 * j], yij = x[3 * i + 1] - x[3 * j + 1], zij = x[3 * i + 2] - x[3 * j + 2];
            double xkj = x[3 * k] - x[3 * j], ykj = x[3 * k + 1] - x[3 * j + 1], zkj = x[3 * k + 2] - x[3 * j + 2];
            const double rij = sqrt(xij * xij + yij * yij + zij * zij), rkj = sqrt(xkj * xkj + ykj * ykj + zkj * zkj);
            xij /= rij; yij /= rij; zij /= rij; xkj /= rkj; ykj /= rkj; zkj /= rkj;
            double dot = xij * xkj + yij * ykj + zij * zkj;
            dot = fmin(0.999999, fmax(-0.999999, dot));
            const double disp = acos(dot) - r.p[0];
            double df = r.p[1] * disp;
            e = df * disp;
            if (g != nullptr) {
                df *= (2.0 * (-1.0 / sqrt(1.0 - dot * dot)));
                const double dtxi = df * (xkj - dot * xij) / rij, dtyi = df * (ykj - dot * yij) / rij, dtzi = df * (zkj - dot * zij) / rij;
                const double dtxk = df * (xij - dot * xkj) / rkj, dtyk = df * (yij - dot * ykj) / rkj, dtzk = df * (zij - dot * zkj) / rkj;
                add3(g, gi, dtxi, dtyi, dtzi); add3(g, gk, dtxk, dtyk, dtzk); add3(g, gj, -(dtxi + dtxk), -(dtyi + dtyk), -(dtzi + dtzk));
            }
        } else if (kind == 3) {                               // Fourier dihedral: E = fc (1 + cos(period phi - phase)), by the angle-addition recurrence
            const Torsion t = torsion(x, i, j, k, l);
            double cosn = 1.0, sinn = 0.0;
            const int period = (int) r.p[3];
            for (int p = 1; p <= period; p++) {
                const double tmp = cosn * t.cosphi - sinn * t.sinphi;
                sinn = cosn * t.sinphi + sinn * t.cosphi;
                cosn = tmp;
            }
            e = r.p[0] * (1.0 + cosn * r.p[1] + sinn * r.p[2]);
            if (g != nullptr) torsion_gradient(t, r.p[0] * r.p[3] * (cosn * r.p[2] - sinn * r.p[1]), gi, gj, gk, gl, g);
        } else {                                              // harmonic improper: E = fc (phi - eq)^2, CHARMM's choice of inverse function
            const Torsion t = torsion(x, i, j, k, l);
            const double cosd = t.cosphi * r.p[1] + t.sinphi * r.p[2], sind = t.sinphi * r.p[1] - t.cosphi * r.p[2];
            double dphi;
            if (cosd > 0.1) dphi = asin(sind);
            else { dphi = fabs(acos(fmax(cosd, -1.0))); if (sind < 0.0) dphi *= -1.0; }
            const double df = r.p[0] * dphi;
            e = df * dphi;
            if (g != nullptr) torsion_gradient(t, 2.0 * df, gi, gj, gk, gl, g);
        }
        atomicAdd(&sE[kind], e);
    }
    __syncthreads();
    if (threadIdx.x < kKinds && sE[threadIdx.x] != 0.0) atomicAdd(&energies[threadIdx.x], sE[threadIdx.x]);
}

static bool upload_terms(MMTerms &m)
{
    std::vector<TermRecord> all;
    for (int k = 0; k < kKinds; k++) { m.start[k] = (int) all.size(); all.insert(all.end(), m.host[k].begin(), m.host[k].end()); }
    m.start[kKinds] = (int) all.size();
    if (!m.rec.ensure(all.size() + 1) || !m.energies.ensure(16)) return false;
    if (!all.empty()) NBB_CUDA(cudaMemcpy(m.rec.p, all.data(), sizeof(TermRecord) * all.size(), cudaMemcpyHostToDevice));
    // a pageable H2D copy returns once the data is STAGED; the DMA runs on the legacy stream, which a non-blocking stream does not wait for:
    // without this the kernel can read a partly uploaded record array (seen as wrong dihedral / improper energies, the tail of the array)
    NBB_CUDA(cudaDeviceSynchronize());
    m.dirty = false;
    return true;
}

// enqueue only: the kernel and the copy of the five energies into pinned memory; collect() waits for them
// fused (nbb200_md_run, alternating slots): the device accumulators have two slots of 8 doubles; the kernel of slot s zeroes slot 1 - s for the
// next step, so no memset is enqueued
static bool enqueue(MMTerms &m, const double *d_x, double *d_grad, int slot = 0, bool fused = false, bool noCopy = false, const int *perm = nullptr, cudaStream_t on = nullptr)
{
    const cudaStream_t st = on != nullptr ? on : m.stream;
    if (m.dirty && !upload_terms(m)) return false;
    const int total = m.start[kKinds];
    fused = fused && total > 0;
    double *acc = m.energies.p + 8 * (slot & 1);
    if (fused && !m.slotsZeroed) { NBB_CUDA(cudaMemsetAsync(m.energies.p, 0, sizeof(double) * 16, st)); m.slotsZeroed = true; }
    if (!fused) { NBB_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * kKinds, st)); m.slotsZeroed = false; }
    if (total > 0) {
        KindStarts K;
        std::memcpy(K.s, m.start, sizeof(K.s));
        k_mm_terms<<<(total + 127) / 128, 128, 0, st>>>(m.rec.p, K, d_x, d_grad, acc, fused ? m.energies.p + 8 * ((slot + 1) & 1) : nullptr, perm);
        m.launches += 1;
    }
    if (!noCopy) NBB_CUDA(cudaMemcpyAsync(m.he + 8 * (slot & 1), acc, sizeof(double) * kKinds, cudaMemcpyDeviceToHost, st));       // noCopy: a later kernel of the caller hands the slot over
    return cuda_ok(cudaGetLastError(), "k_mm_terms");
}

static bool collect(MMTerms &m, double *energies5)
{
    NBB_CUDA(cudaStreamSynchronize(m.stream));
    for (int k = 0; k < kKinds; k++) energies5[k] = m.he[k];
    return true;
}

static bool evaluate(MMTerms &m, const double *d_x, double *d_grad, double *energies5)
{
    return enqueue(m, d_x, d_grad) && collect(m, energies5);
}

// ---- Langevin velocity Verlet, first part of an Iteration (LangevinVelocityVerletIntegrator.py:117-131,139-149) in Cartesian variables:
//   x += facR1 v + facR2 a + sdR w1 / sqrt(m) ; v = facV1 v + facV2 a + (sdV1 w1 + sdV2 w2) / sqrt(m)
// with w1, w2 independent standard normal deviates per coordinate (the reference integrates mass-weighted variables, where the random
// terms carry no mass factor).  Deviates: Philox-4x32-10 keyed by (seed), counter (coordinate index, step), Box-Muller.
__device__ __forceinline__ void philox4x32(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int k0, unsigned int k1, unsigned int *out)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// the two deviates of coordinate i at step `step`: two uniforms in (0, 1] with 52 bits from two Philox words each -> Box-Muller pair
__device__ __forceinline__ void langevin_deviates(long i, unsigned long long step, unsigned long long seed, double &w1, double &w2)
{
    unsigned int r[4];
    philox4x32((unsigned int) i, (unsigned int) ((unsigned long long) i >> 32), (unsigned int) step, (unsigned int) (step >> 32),
               (unsigned int) seed, (unsigned int) (seed >> 32), r);
    const double u1 = ((double) (((unsigned long long) r[0] << 20) | (r[1] >> 12)) + 1.0) * (1.0 / 4503599627370496.0);
    const double u2 = ((double) (((unsigned long long) r[2] << 20) | (r[3] >> 12)) + 1.0) * (1.0 / 4503599627370496.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    w1 = rad * cs; w2 = rad * sn;
}

// ApplyLinearConstraints on the random vectors (LangevinVelocityVerletIntegrator.py:139-149 -> SystemGeometryObjectiveFunction.ApplyLinearConstraints,
// pMolecule-1.9.0/pMolecule/SystemGeometryObjectiveFunction.py:99-101) for the constraint set of a periodic system after RemoveRotationTranslation
// (:213-240): the three mass-weighted translation vectors c_d[i] = sqrt(m_i / M).  Projecting them out of a deviate vector w (mass-weighted
// variables) is w_id -= sqrt(m_i) S_d / M with S_d = sum_j sqrt(m_j) w_jd, i.e. in Cartesian variables the constant S_d / M is taken off every
// atom's random displacement.  The sums need all deviates of a step before the first one is used: the kernel of step k reads sums[k % 3]
// (made by the kernel of step k - 1, or by k_langevin_sums at the start of a run), accumulates the sums of step k + 1 into sums[(k + 1) % 3]
// from the deviates of that step (counter based: they depend on nothing but (seed, coordinate, step)) and clears sums[(k + 2) % 3].
// Every deviate is still generated once: the kernel that needs the sums of step k + 1 stores that step's deviates for the next launch.
struct LangevinConstraint {
    const double *sumsIn;            // [6]: S1_xyz, S2_xyz of this step; null: no constraints
    double *sumsOut, *zeroNext;      // [6] each
    const double2 *wIn;              // [3n] the deviates (w1, w2) of this step, stored by the previous launch (or k_langevin_sums)
    double2 *wOut;                   // [3n] those of the next step
    double invTotalMass;
};

__device__ __forceinline__ void langevin_accumulate_next(const LangevinConstraint &LC, double (&acc)[6])
{
#pragma unroll
    for (int k = 0; k < 6; k++) {
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) if (acc[k] != 0.0) atomicAdd(LC.sumsOut + k, acc[k]);
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) LC.zeroNext[threadIdx.x] = 0.0;
}

// the sums of one step on their own (start of a run, or a step that does not follow the previous call's)
__global__ void k_langevin_sums(const double *__restrict__ mass, int n, unsigned long long seed, unsigned long long step, double *__restrict__ sums, double2 *__restrict__ wOut)
{
    const int atom = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (atom < n) {
        const double sm = sqrt(mass[atom]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double w1, w2;
            langevin_deviates(3 * (long) atom + c, step, seed, w1, w2);
            wOut[3 * (long) atom + c] = make_double2(w1, w2);
            acc[c] = sm * w1; acc[3 + c] = sm * w2;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) if (acc[k] != 0.0) atomicAdd(sums + k, acc[k]);
    }
}

// one thread per atom (its three coordinates; the Philox counter is the coordinate index).  xref != null: fused with the displacement check of
// CheckForUpdate (pM/csource/NBModelABFS.c:691-746) for nbb200_md_run -- |x - xref|^2 into the running maximum
__global__ void k_langevin_first_disp(double *__restrict__ x, double *__restrict__ v, const double *__restrict__ a, const double *__restrict__ mass, int n,
                                      double facR1, double facR2, double facV1, double facV2, double sdR, double sdV1, double sdV2,
                                      unsigned long long seed, unsigned long long step, const double *__restrict__ xref, const unsigned char *__restrict__ fixed,
                                      unsigned long long *__restrict__ out, unsigned long long *__restrict__ zeroOther, const LangevinConstraint LC)
{
    if (zeroOther != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zeroOther = 0ULL;
    const int atom = blockIdx.x * blockDim.x + threadIdx.x;
    double r2 = 0.0;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (atom < n) {
        const double mi = mass[atom], rsm = rsqrt(mi), sm = sqrt(mi);
        double d[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const long i = 3 * (long) atom + c;
            double w1, w2, n1, n2;                           // n: the deviates in Cartesian variables
            if (LC.sumsIn != nullptr) {
                const double2 ws = LC.wIn[i];
                w1 = ws.x; w2 = ws.y;
                n1 = w1 * rsm - LC.sumsIn[c] * LC.invTotalMass; n2 = w2 * rsm - LC.sumsIn[3 + c] * LC.invTotalMass;
                double v1, v2;
                langevin_deviates(i, step + 1ULL, seed, v1, v2);
                LC.wOut[i] = make_double2(v1, v2);
                acc[c] = sm * v1; acc[3 + c] = sm * v2;
            } else {
                langevin_deviates(i, step, seed, w1, w2);
                n1 = w1 * rsm; n2 = w2 * rsm;
            }
            const double vi = v[i], ai = a[i];
            const double xn = x[i] + (facR1 * vi + facR2 * ai + sdR * n1);
            x[i] = xn;
            v[i] = facV1 * vi + facV2 * ai + (sdV1 * n1 + sdV2 * n2);
            if (xref != nullptr) d[c] = xn - xref[i];
        }
        if (xref != nullptr && (fixed == nullptr || !fixed[atom])) r2 = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
    }
    if (LC.sumsIn != nullptr) langevin_accumulate_next(LC, acc);
    if (xref == nullptr) return;
    for (int off = 16; off > 0; off >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, off));
    if ((threadIdx.x & 31) == 0 && r2 > 0.0) atomicMax(out, (unsigned long long) __double_as_longlong(r2));
}

// constraint bookkeeping of one first-half launch at `step`: which sums to read, where the next step's go (State::lc*)
static bool langevin_constraint(State &s, const double *d_mass, unsigned long long seed, unsigned long long step, LangevinConstraint &LC)
{
    LC.sumsIn = nullptr; LC.sumsOut = nullptr; LC.zeroNext = nullptr; LC.wIn = nullptr; LC.wOut = nullptr; LC.invTotalMass = 0.0;
    if (!s.lcOn) return true;
    const size_t m = 3 * (size_t) s.n;
    if (!s.lcSums.ensure(18) || !s.lcW.ensure(2 * m)) return false;
    if (!(s.lcValid && s.lcStep == step && s.lcSeed == seed)) {
        NBB_CUDA(cudaMemsetAsync(s.lcSums.p, 0, sizeof(double) * 18, s.stream));
        k_langevin_sums<<<(s.n + 127) / 128, 128, 0, s.stream>>>(d_mass, s.n, seed, step, s.lcSums.p + 6 * (step % 3ULL), s.lcW.p + (step & 1ULL) * m);
        s.launches += 1;
    }
    LC.wIn = s.lcW.p + (step & 1ULL) * m; LC.wOut = s.lcW.p + ((step + 1ULL) & 1ULL) * m;
    LC.sumsIn = s.lcSums.p + 6 * (step % 3ULL); LC.sumsOut = s.lcSums.p + 6 * ((step + 1ULL) % 3ULL); LC.zeroNext = s.lcSums.p + 6 * ((step + 2ULL) % 3ULL);
    LC.invTotalMass = 1.0 / s.lcTotalMass;
    s.lcValid = true; s.lcStep = step + 1ULL; s.lcSeed = seed;
    return true;
}

bool langevin_first_disp(State &s, double *d_x, double *d_v, const double *d_a, const double *d_mass, const double *f7, unsigned long long seed,
                         unsigned long long step, double *d_out, double *d_zeroOther)
{
    LangevinConstraint LC;
    if (!langevin_constraint(s, d_mass, seed, step, LC)) return false;
    k_langevin_first_disp<<<(s.n + 127) / 128, 128, 0, s.stream>>>(d_x, d_v, d_a, d_mass, s.n, f7[0], f7[1], f7[2], f7[3], f7[4], f7[5], f7[6], seed, step, s.xref.p,
                                                                   s.nfixed > 0 ? s.fixedFlag.p : nullptr, reinterpret_cast<unsigned long long *>(d_out),
                                                                   reinterpret_cast<unsigned long long *>(d_zeroOther), LC);
    s.launches += 1;
    return cuda_ok(cudaGetLastError(), "k_langevin_first_disp");
}

bool langevin_first(State &s, double *d_x, double *d_v, const double *d_a, const double *d_mass, const double *f7, unsigned long long seed, unsigned long long step)
{
    LangevinConstraint LC;
    if (!langevin_constraint(s, d_mass, seed, step, LC)) return false;
    k_langevin_first_disp<<<(s.n + 127) / 128, 128, 0, s.stream>>>(d_x, d_v, d_a, d_mass, s.n, f7[0], f7[1], f7[2], f7[3], f7[4], f7[5], f7[6], seed, step, nullptr,
                                                                   nullptr, nullptr, nullptr, LC);
    s.launches += 1;
    return cuda_ok(cudaGetLastError(), "k_langevin_first");
}

}  // namespace nbb200

using namespace nbb200;

static void mm_status(int *status, int value) { if (status != nullptr) *status = value; }

extern "C" {

NBB200MMTerms *MMTerms_B200_Allocate(int device, int natoms, int *status)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError(); set_error("no CUDA device available: libnbabfs_b200 has no CPU fallback"); mm_status(status, NBB200_STATUS_LOGIC_ERROR); return nullptr;
    }
    if (natoms <= 0) { set_error("invalid number of atoms"); mm_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    MMTerms *m = new (std::nothrow) MMTerms();
    if (m == nullptr) { mm_status(status, NBB200_STATUS_OUT_OF_MEMORY); return nullptr; }
    m->device = device; m->natoms = natoms;
    cudaSetDevice(device);
    bool ok = cuda_ok(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    m->ownStream = ok;
    ok = ok && m->x.ensure(3 * (size_t) natoms) && m->grad.ensure(3 * (size_t) natoms) && m->energies.ensure(16);
    ok = ok && cuda_ok(cudaMallocHost((void **) &m->hx, sizeof(double) * 3 * (size_t) natoms), "cudaMallocHost");
    ok = ok && cuda_ok(cudaMallocHost((void **) &m->hg, sizeof(double) * 3 * (size_t) natoms), "cudaMallocHost");
    ok = ok && cuda_ok(cudaMallocHost((void **) &m->he, sizeof(double) * 16), "cudaMallocHost");      // two slots of 8
    if (!ok) { NBB200MMTerms *h = reinterpret_cast<NBB200MMTerms *>(m); MMTerms_B200_Deallocate(&h); mm_status(status, NBB200_STATUS_OUT_OF_MEMORY); return nullptr; }
    return reinterpret_cast<NBB200MMTerms *>(m);
}

void MMTerms_B200_Deallocate(NBB200MMTerms **terms)
{
    if (terms == nullptr || *terms == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(*terms);
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    m->rec.release(); m->x.release(); m->grad.release(); m->energies.release();
    if (m->hx) cudaFreeHost(m->hx);
    if (m->hg) cudaFreeHost(m->hg);
    if (m->he) cudaFreeHost(m->he);
    if (m->ownStream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
    *terms = nullptr;
}

void MMTerms_B200_SetStream(NBB200MMTerms *terms, void *cudaStream)
{
    if (terms == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    cudaSetDevice(m->device);
    if (m->ownStream && m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    m->stream = reinterpret_cast<cudaStream_t>(cudaStream);
    m->ownStream = false;
}

// common part of the four Define entry points: terms with atoms / type / QACTIVE, parameters by type
static void define_terms(NBB200MMTerms *terms, int kind, int natomsPerTerm, int nterms, const int *atoms, const int *types, const unsigned char *active,
                         int nparameters, const double *p0, const double *p1, const double *p2, const double *p3, int *status)
{
    if (terms == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    if (nterms < 0 || nparameters < 0 || (nterms > 0 && (atoms == nullptr || types == nullptr || p0 == nullptr || p1 == nullptr))) {
        set_error("invalid argument to a B200 MM term container"); mm_status(status, NBB200_STATUS_INVALID_ARGUMENT); return;
    }
    std::vector<TermRecord> out;
    out.reserve((size_t) nterms);
    for (int n = 0; n < nterms; n++) {
        if (active != nullptr && !active[n]) continue;                   // QACTIVE (fixed-atom / QC-atom deactivation in the reference)
        const int t = types[n];
        if (t < 0 || t >= nparameters) { set_error("term type out of range"); mm_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
        TermRecord r;
        for (int a = 0; a < 4; a++) {
            r.atom[a] = (a < natomsPerTerm) ? atoms[natomsPerTerm * n + a] : 0;
            if (r.atom[a] < 0 || r.atom[a] >= m->natoms) { set_error("term atom out of range"); mm_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
        }
        r.p[0] = p0[t]; r.p[1] = p1[t]; r.p[2] = (p2 != nullptr) ? p2[t] : 0.0; r.p[3] = (p3 != nullptr) ? p3[t] : 0.0;
        out.push_back(r);
    }
    m->host[kind].swap(out);
    m->dirty = true;
}

void HarmonicBondContainer_B200_Define(NBB200MMTerms *terms, int isUreyBradley, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                       int nparameters, const double *eq, const double *fc, int *status)
{
    define_terms(terms, isUreyBradley ? 2 : 0, 2, nterms, atoms, types, active, nparameters, eq, fc, nullptr, nullptr, status);
}

void HarmonicAngleContainer_B200_Define(NBB200MMTerms *terms, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                        int nparameters, const double *eq, const double *fc, int *status)
{
    define_terms(terms, 1, 3, nterms, atoms, types, active, nparameters, eq, fc, nullptr, nullptr, status);
}

void FourierDihedralContainer_B200_Define(NBB200MMTerms *terms, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                          int nparameters, const double *fc, const int *period, const double *phase, int *status)
{
    if (nparameters > 0 && (period == nullptr || phase == nullptr)) { set_error("invalid argument to a B200 MM term container"); mm_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
    std::vector<double> c((size_t) std::max(nparameters, 0)), s(c.size()), per(c.size());
    for (int i = 0; i < nparameters; i++) { c[i] = std::cos(phase[i]); s[i] = std::sin(phase[i]); per[i] = (double) period[i]; }   // FourierDihedralContainer_FillCosSinPhases (:246-257)
    define_terms(terms, 3, 4, nterms, atoms, types, active, nparameters, fc, c.data(), s.data(), per.data(), status);
}

void HarmonicImproperContainer_B200_Define(NBB200MMTerms *terms, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                           int nparameters, const double *eq, const double *fc, int *status)
{
    if (nparameters > 0 && eq == nullptr) { set_error("invalid argument to a B200 MM term container"); mm_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
    std::vector<double> c((size_t) std::max(nparameters, 0)), s(c.size());
    for (int i = 0; i < nparameters; i++) { c[i] = std::cos(eq[i]); s[i] = std::sin(eq[i]); }                                      // HarmonicImproperContainer_FillCosSinValues (:154-165)
    define_terms(terms, 4, 4, nterms, atoms, types, active, nparameters, fc, c.data(), s.data(), nullptr, status);
}

void MMTerms_B200_EnergyDevice(NBB200MMTerms *terms, const double *d_xyz, double *energies5, double *d_grad, int *status)
{
    if (terms == nullptr || d_xyz == nullptr || energies5 == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    cudaSetDevice(m->device);
    if (!evaluate(*m, d_xyz, d_grad, energies5)) mm_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void MMTerms_B200_EnergyDeviceEnqueue(NBB200MMTerms *terms, const double *d_xyz, double *d_grad, int *status)
{
    if (terms == nullptr || d_xyz == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    cudaSetDevice(m->device);
    if (!enqueue(*m, d_xyz, d_grad)) mm_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void MMTerms_B200_EnergyDeviceCollect(NBB200MMTerms *terms, double *energies5, int *status)
{
    if (terms == nullptr || energies5 == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    cudaSetDevice(m->device);
    if (!collect(*m, energies5)) mm_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void MMTerms_B200_Energy(NBB200MMTerms *terms, const double *xyz, double *energies5, double *grad, int *status)
{
    if (terms == nullptr || xyz == nullptr || energies5 == nullptr) return;
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    cudaSetDevice(m->device);
    const size_t bytes = sizeof(double) * 3 * (size_t) m->natoms;
    std::memcpy(m->hx, xyz, bytes);
    bool ok = cuda_ok(cudaMemcpyAsync(m->x.p, m->hx, bytes, cudaMemcpyHostToDevice, m->stream), "H2D coordinates");
    if (ok && grad != nullptr) ok = cuda_ok(cudaMemsetAsync(m->grad.p, 0, bytes, m->stream), "memset");
    ok = ok && evaluate(*m, m->x.p, grad != nullptr ? m->grad.p : nullptr, energies5);
    if (ok && grad != nullptr) {
        ok = cuda_ok(cudaMemcpyAsync(m->hg, m->grad.p, bytes, cudaMemcpyDeviceToHost, m->stream), "D2H gradients") && cuda_ok(cudaStreamSynchronize(m->stream), "sync");
        if (ok) for (size_t i = 0; i < 3 * (size_t) m->natoms; i++) grad[i] += m->hg[i];          // accumulated, as every term of System.Energy
    }
    if (!ok) mm_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void MMTerms_B200_LastEnergies(NBB200MMTerms *terms, double *energies5)
{
    if (terms == nullptr || energies5 == nullptr) return;
    const MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    for (int k = 0; k < kKinds; k++) energies5[k] = m->he[k];
}

}  // extern "C"

// internal (nbb200_md_run): two result slots, so that a step's energies stay readable while the next step is already in flight
namespace nbb200 {
bool mmterms_enqueue_slot(NBB200MMTerms *terms, const double *d_x, double *d_grad, int slot, bool fused, bool noCopy, const int *perm, cudaStream_t on)
{
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    return enqueue(*m, d_x, d_grad, slot, fused, noCopy, perm, on);
}
void mmterms_slot_pointers(NBB200MMTerms *terms, int slot, const double **d_energies, double **h_energies)
{
    MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    *d_energies = m->energies.p + 8 * (slot & 1); *h_energies = m->he + 8 * (slot & 1);
}
void mmterms_reset_slots(NBB200MMTerms *terms)        // start of a run: the slot the first step uses may hold an earlier run's last energies
{
    reinterpret_cast<MMTerms *>(terms)->slotsZeroed = false;
}
void mmterms_read_slot(NBB200MMTerms *terms, int slot, double *energies5)
{
    const MMTerms *m = reinterpret_cast<MMTerms *>(terms);
    for (int k = 0; k < kKinds; k++) energies5[k] = m->he[8 * (slot & 1) + k];
}
}

extern "C" {

long MMTerms_B200_NumberOfTerms(NBB200MMTerms *terms, int kind)
{
    if (terms == nullptr || kind < 0 || kind >= kKinds) return 0;
    return (long) reinterpret_cast<MMTerms *>(terms)->host[kind].size();
}

void nbb200_langevin_first_half(NBB200State *state, double *d_x, double *d_v, const double *d_a, const double *d_mass, const double *factors7,
                                unsigned long long seed, unsigned long long step)
{
    if (state == nullptr || factors7 == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    langevin_first(s, d_x, d_v, d_a, d_mass, factors7, seed, step);
}

void nbb200_set_langevin_constraints(NBB200State *state, int removeTranslation, double totalMass)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    s.lcOn = removeTranslation != 0 && totalMass > 0.0;
    s.lcTotalMass = totalMass; s.lcValid = false;
}

}  // extern "C"
