// qcmm.cu -- the QC/MM entry points of NBModelABFS for a QC region without boundary atoms:
//
//   NBModelABFS_QCMMEnergyLJ    pMolecule-1.9.0/extensions/csource/NBModelABFS.c:306-378  (PairwiseInteractionABFS_MMMMEnergy, charges off; MMMMImageEnergy)
//   NBModelABFS_QCMMPotentials  NBModelABFS.c:454-498, QCMMImagePotentials :1402-1462, QCQCImagePotentials :1557-1610, PairwiseInteraction.c:612-663,741-787
//   NBModelABFS_QCMMGradients   NBModelABFS.c:383-449, QCMMImageGradients :1318-1396, QCQCImageGradients :1467-1552, PairwiseInteraction.c:542-610,664-736
//
// The reference walks pair lists (nbqcmmlj/el, inbqcmmlj/el, inbqcqclj/el).  Every pair within the outer cutoff is on a valid list and pairs
// beyond it are skipped (PairwiseInteraction.h:72-78, PairwiseInteraction.c:654), so the sums do not depend on the lists; a QC region is
// tens of atoms, so ONE fp64 launch over (QC atom) x (image) x (atom) does a whole entry point.  The images are those of the state's
// image plan (GenerateImageLists' loop: one member of each inverse pair with scale 1, self-inverse operations with scale 0.5), and for
// every image both orientations the reference lists are taken -- (QC atom, image of any atom) and (MM atom, image of the QC atom).  Keeping
// the reference's choice of images matters for dE/dM: for a lattice distortion that breaks the space-group symmetry an operation and
// its inverse do not give the same derivative (M S M^-1 is no longer orthogonal), and the reference's numbers carry that asymmetry.
// Supported: vacuum, P1 cells (any number of MM atoms), and cells with space-group operations when every atom is a QC atom (QC/QC image
// terms only: the case the reference's crystal tests and the golden vectors cover).  Without boundary atoms the 1-4 lists of these terms
// are empty.  Restated in numpy and pinned to the compiled reference in oracle/qcmm_oracle.py / tests/test_oracle_qcmm.py.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "nbb200_internal.h"
#include "../../include/nbabfs_b200.h"

namespace nbb200 {

struct QCFactors { double v[21]; };
constexpr double kHartreeToKJMol = 2625.5;          // UNITS_ENERGY_HARTREES_TO_KILOJOULES_PER_MOLE (pCore-1.9.0/extensions/cinclude/Units.h:52)

static __device__ __forceinline__ double qc_warp_sum(double v)
{
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// CubicSpline_Evaluate on the abscissae x = r^2 (pC/csource/CubicSpline.c:138-159 bisection, pC/cinclude/CubicSpline.h:30-39 value and
// first derivative); spl = x[n], y[n], h[n]
static __device__ __forceinline__ void qc_spline(const double *__restrict__ spl, int n, double r2, double &f, double &g)
{
    int l = 0, u = n - 1;
    while (u - l > 1) { const int m = (u + l) >> 1; if (spl[m] > r2) u = m; else l = m; }
    const double *y = spl + n, *h = spl + 2 * n;
    const double d = spl[u] - spl[l], sv = (r2 - spl[l]) / d, tv = (spl[u] - r2) / d;
    const double hl = h[l] * d / 6.0, hu = h[u] * d / 6.0, yl = y[l], yu = y[u];
    f = tv * yl + sv * yu + d * (tv * (tv * tv - 1.0) * hl + sv * (sv * sv - 1.0) * hu);
    g = (yu - yl) / d + (-(3.0 * tv * tv - 1.0) * hl + (3.0 * sv * sv - 1.0) * hu);
}

struct QCArgs {
    int nq, n, nimg, ntypes, splN;
    const int *qcIdx;                    // [nq] atom index of every QC atom
    const int *qcSlot;                   // [n] position in the QC container or -1
    const double *x;
    const int *ljtype; const double2 *ljAB;
    const int *exclPtr, *exclCol;
    const unsigned char *fixed;          // nullable
    const double *mmq;                   // [n] MM charges / dielectric, zero for QC atoms
    const double *qcq;                   // [nq] QC charges (gradients)
    const double *images;                // [nimg][13]: R (row major), tv, scale; image 0 = the identity
    const double *spl;                   // atomic-unit electrostatic spline
    QCFactors F;
    double invDielectric;
    double *grad;                        // [3 n] or null
    double *acc;                         // [4 + 12 nimg]: eqcmmlj, eimqcmmlj, eimqcqclj, -, then per image G[3], W[9]
    double *pot;                         // [nq]
    double *V;                           // [nq * nq] QC/QC image potentials, unsymmetrised
};

// one pair of the brute-force sums.  kMode: 0 = Lennard-Jones, 1 = potentials, 2 = electrostatic gradients.  Returns false when the pair does
// not contribute; otherwise coef (gradient on the first atom = coef * d, on the second -coef * d), the LJ energy or the spline value in val
template <int kMode>
static __device__ __forceinline__ bool qc_pair(const QCArgs &A, double r2, int tq, int j, double w, double chargeProduct, double &coef, double &val)
{
    const double *F = A.F.v;
    if (r2 > F[2]) return false;
    if (kMode == 0) {
        const double2 ab = A.ljAB[tq + A.ljtype[j]];
        double s = 0.0, s2 = 0.0, dF = 0.0, e2;
        if (!(r2 < F[0])) { s2 = 1.0 / r2; s = sqrt(s2); }
        const double s6 = s2 * s2 * s2;
        if (r2 > F[1]) {
            const double l1 = s6 - F[11], l2 = (s / r2) - F[16];
            e2 = ab.x * F[12] * l1 * l1 - ab.y * F[17] * l2 * l2;
            dF = -3.0 * s6 * (2.0 * ab.x * F[12] * l1 / r2 - ab.y * F[17] * l2 / s);
        } else if (r2 > F[0]) {
            e2 = ab.x * (s6 * s6 - F[13]) - ab.y * (s6 - F[18]);
            dF = -3.0 * s6 * (2.0 * ab.x * s6 - ab.y) / r2;
        } else {
            e2 = ab.x * (F[14] - F[15] * r2) - ab.y * (F[19] - F[20] * r2);
            dF = -ab.x * F[15] + ab.y * F[20];
        }
        val = w * e2;
        coef = w * 2.0 * dF;
        return true;
    }
    if (chargeProduct == 0.0 && kMode == 2) return false;
    double f, g;
    qc_spline(A.spl, A.splN, r2, f, g);
    val = w * f;
    coef = kHartreeToKJMol * w * chargeProduct * 2.0 * g;
    return true;
}

// image gradient bookkeeping: gimg = gradient on the IMAGE atom (image frame), xprim = the primary coordinates of the atom that was imaged;
// its own gradient is R^T gimg; SymmetryParameterGradients_ImageDerivatives needs G = sum gimg and W = sum gimg (x) xprim per image
static __device__ __forceinline__ void qc_image_gradient(const QCArgs &A, int img, const double *R, int atom, double gx, double gy, double gz, double xp, double yp, double zp)
{
    if (A.grad != nullptr) {
        atomicAdd(&A.grad[3 * atom], R[0] * gx + R[3] * gy + R[6] * gz);
        atomicAdd(&A.grad[3 * atom + 1], R[1] * gx + R[4] * gy + R[7] * gz);
        atomicAdd(&A.grad[3 * atom + 2], R[2] * gx + R[5] * gy + R[8] * gz);
    }
    double *a = A.acc + 4 + 12 * (size_t) img;
    atomicAdd(a, gx); atomicAdd(a + 1, gy); atomicAdd(a + 2, gz);
    atomicAdd(a + 3, gx * xp); atomicAdd(a + 4, gx * yp); atomicAdd(a + 5, gx * zp);
    atomicAdd(a + 6, gy * xp); atomicAdd(a + 7, gy * yp); atomicAdd(a + 8, gy * zp);
    atomicAdd(a + 9, gz * xp); atomicAdd(a + 10, gz * yp); atomicAdd(a + 11, gz * zp);
}

// grid = (chunks of the (image, atom) space, QC atoms)
template <int kMode>
static __global__ void k_qcmm(const __grid_constant__ QCArgs A)
{
    const int k = blockIdx.y, q = A.qcIdx[k];
    const double xq = A.x[3 * q], yq = A.x[3 * q + 1], zq = A.x[3 * q + 2];
    const int tq = A.ljtype[q] * A.ntypes;
    const bool qFixed = A.fixed != nullptr && A.fixed[q];
    const double qcharge = (kMode == 2) ? A.qcq[k] : 0.0;
    double e[3] = {0.0, 0.0, 0.0}, gq[3] = {0.0, 0.0, 0.0}, pot = 0.0;
    const long total = (long) A.nimg * A.n;
    for (long idx = (long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long) gridDim.x * blockDim.x) {
        const int img = (int) (idx / A.n), j = (int) (idx - (long) img * A.n);
        const double *R = A.images + 13 * img;
        const double scale = R[12];
        const double xj = A.x[3 * j], yj = A.x[3 * j + 1], zj = A.x[3 * j + 2];
        const int slotj = A.qcSlot[j];
        const bool jqc = slotj >= 0;
        if (qFixed && A.fixed[j]) continue;                  // pairs of two fixed atoms are on no list (freeSelection of the generators)
        if (img == 0) {
            if (jqc) continue;                               // pairs inside the QC region belong to the QC model
            bool excluded = false;
            for (int c = A.exclPtr[q]; c < A.exclPtr[q + 1]; c++) excluded = excluded || (A.exclCol[c] == j);
            if (excluded) continue;
        }
        const double cj = jqc ? (kMode == 2 ? A.qcq[slotj] * A.invDielectric : A.invDielectric) : A.mmq[j];      // charge of j (QC: its QC charge; potentials: unit)
        if (kMode != 0 && !jqc && cj == 0.0) continue;
        // ---- orientation A: the QC atom against the image of atom j
        {
            const double dx = xq - (R[0] * xj + R[1] * yj + R[2] * zj + R[9]), dy = yq - (R[3] * xj + R[4] * yj + R[5] * zj + R[10]),
                         dz = zq - (R[6] * xj + R[7] * yj + R[8] * zj + R[11]);
            double coef, val;
            if (qc_pair<kMode>(A, dx * dx + dy * dy + dz * dz, tq, j, scale, qcharge * cj, coef, val)) {
                if (kMode == 0) e[img == 0 ? 0 : (jqc ? 2 : 1)] += val;
                if (kMode == 1) { if (jqc) atomicAdd(&A.V[(size_t) k * A.nq + slotj], val * cj); else pot += cj * val; }
                else {
                    const double gx = coef * dx, gy = coef * dy, gz = coef * dz;
                    gq[0] += gx; gq[1] += gy; gq[2] += gz;
                    if (img == 0) { if (A.grad != nullptr) { atomicAdd(&A.grad[3 * j], -gx); atomicAdd(&A.grad[3 * j + 1], -gy); atomicAdd(&A.grad[3 * j + 2], -gz); } }
                    else qc_image_gradient(A, img, R, j, -gx, -gy, -gz, xj, yj, zj);
                }
            }
        }
        // ---- orientation B: the MM atom j against the image of the QC atom (the other half of the reference's QC/MM image lists)
        if (img > 0 && !jqc) {
            const double dx = xj - (R[0] * xq + R[1] * yq + R[2] * zq + R[9]), dy = yj - (R[3] * xq + R[4] * yq + R[5] * zq + R[10]),
                         dz = zj - (R[6] * xq + R[7] * yq + R[8] * zq + R[11]);
            double coef, val;
            if (qc_pair<kMode>(A, dx * dx + dy * dy + dz * dz, tq, j, scale, qcharge * cj, coef, val)) {
                if (kMode == 0) e[1] += val;
                if (kMode == 1) pot += cj * val;
                else {
                    const double gx = coef * dx, gy = coef * dy, gz = coef * dz;
                    if (A.grad != nullptr) { atomicAdd(&A.grad[3 * j], gx); atomicAdd(&A.grad[3 * j + 1], gy); atomicAdd(&A.grad[3 * j + 2], gz); }
                    qc_image_gradient(A, img, R, q, -gx, -gy, -gz, xq, yq, zq);
                }
            }
        }
    }
    if (kMode == 1) {
        const double p = qc_warp_sum(pot);
        if ((threadIdx.x & 31) == 0 && p != 0.0) atomicAdd(&A.pot[k], p);
    }
#pragma unroll
    for (int c = 0; c < (kMode == 1 ? 0 : 3); c++) {
        const double gs = qc_warp_sum(gq[c]);
        if ((threadIdx.x & 31) == 0 && gs != 0.0 && A.grad != nullptr) atomicAdd(&A.grad[3 * q + c], gs);
        if (kMode == 0) {
            const double es = qc_warp_sum(e[c]);
            if ((threadIdx.x & 31) == 0 && es != 0.0) atomicAdd(&A.acc[c], es);
        }
    }
}

struct QCImageHost { int t, a, b, c; };

// everything an entry point needs on the device; rebuilt by every call (a QC region is small, the calls are rare next to the MM/MM term)
struct QCContext {
    QCArgs A;
    std::vector<QCImageHost> images;
    std::vector<double> acc;
};

static bool qc_prepare(State &s, QCContext &C, const char *who, bool needSpline)
{
    auto fail = [&](const std::string &m) { set_error(std::string(who) + ": " + m); return false; };
    if (s.isNew || s.xcur == nullptr) return fail("called before the first Update");
    if (s.hostLJ64.empty()) return fail("no fp64 LJ table");
    if (s.trans.n > 1 && s.nqc < s.n) return fail("space-group operations are supported only when every atom is a QC atom (vacuum and P1 cells otherwise)");
    if (s.useCentering) return fail("useCentering is not supported with a QC region");
    const int n = s.n;
    std::vector<int> qc, slot((size_t) n, -1);
    for (int i = 0; i < n; i++) if (s.hostQC[i]) { slot[i] = (int) qc.size(); qc.push_back(i); }
    const int nq = (int) qc.size();
    if (!s.qcIdxDev.ensure((size_t) nq) || !s.qcSlotDev.ensure((size_t) n)) return false;
    NBB_CUDA(cudaMemcpyAsync(s.qcIdxDev.p, qc.data(), sizeof(int) * nq, cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaMemcpyAsync(s.qcSlotDev.p, slot.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    // image operations: the identity first, then the images of the state's plan (every image that can hold a pair within the LIST cutoff of
    // the atoms' bounding box -- built from the actual coordinates at the last update, so molecules that have diffused any number of cells
    // away are covered) with the CURRENT lattice, as energy_enqueue does for the MM/MM term (NBModelABFS.c:1246-1256)
    std::vector<double> img = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
    C.images.assign(1, QCImageHost{-1, 0, 0, 0});
    for (const CandidateImage &im : s.plan.images) {
        const double xt[3] = {s.trans.trans[3 * im.t] + (double) im.a, s.trans.trans[3 * im.t + 1] + (double) im.b, s.trans.trans[3 * im.t + 2] + (double) im.c};
        const RealSpaceOp op = orthogonalize(s.trans.rot[im.t], xt, s.lattice);
        for (int r = 0; r < 9; r++) img.push_back(op.R.v[r]);
        for (int r = 0; r < 3; r++) img.push_back(op.tv[r]);
        img.push_back(im.scale);
        C.images.push_back(QCImageHost{im.t, im.a, im.b, im.c});
    }
    const int nimg = (int) C.images.size();
    const size_t accCount = 4 + 12 * (size_t) nimg;
    std::vector<double> mmq((size_t) n);
    // mmCharges: the charges of the active MM atoms over the dielectric (the caller's charges as given to SetUp)
    NBB_CUDA(cudaMemcpyAsync(mmq.data(), s.q64.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < n; i++) mmq[i] = s.hostQC[i] ? 0.0 : mmq[i] / s.dielectric;
    std::vector<double> spl;
    int splN = 0;
    if (needSpline) {
        std::vector<double> vx, vy, vh;
        make_abfs_electrostatic_spline_au(s.damp, s.inner, s.outer, s.splineDensity, vx, vy, vh);
        splN = (int) vx.size();
        spl.insert(spl.end(), vx.begin(), vx.end()); spl.insert(spl.end(), vy.begin(), vy.end()); spl.insert(spl.end(), vh.begin(), vh.end());
    }
    if (!s.qcImagesDev.ensure(img.size()) || !s.qcLJDev.ensure(s.hostLJ64.size()) || !s.qcMMqDev.ensure((size_t) n) || !s.qcAccDev.ensure(accCount) ||
        !s.qcGradDev.ensure(3 * (size_t) n) || !s.qcSplDev.ensure(std::max<size_t>(spl.size(), 1)) || !s.qcPotDev.ensure((size_t) nq * (nq + 2))) return false;
    NBB_CUDA(cudaMemcpyAsync(s.qcImagesDev.p, img.data(), sizeof(double) * img.size(), cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaMemcpyAsync(s.qcLJDev.p, s.hostLJ64.data(), sizeof(double2) * s.hostLJ64.size(), cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaMemcpyAsync(s.qcMMqDev.p, mmq.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
    if (!spl.empty()) NBB_CUDA(cudaMemcpyAsync(s.qcSplDev.p, spl.data(), sizeof(double) * spl.size(), cudaMemcpyHostToDevice, s.stream));
    NBB_CUDA(cudaMemsetAsync(s.qcAccDev.p, 0, sizeof(double) * accCount, s.stream));
    NBB_CUDA(cudaMemsetAsync(s.qcGradDev.p, 0, sizeof(double) * 3 * (size_t) n, s.stream));
    NBB_CUDA(cudaMemsetAsync(s.qcPotDev.p, 0, sizeof(double) * (size_t) nq * (nq + 2), s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));               // the staging vectors above are pageable host memory
    QCArgs &A = C.A;
    A.nq = nq; A.n = n; A.nimg = nimg; A.ntypes = s.ntypes; A.splN = splN;
    A.qcIdx = s.qcIdxDev.p; A.qcSlot = s.qcSlotDev.p; A.x = s.xcur; A.ljtype = s.ljtype.p; A.ljAB = s.qcLJDev.p;
    A.exclPtr = s.exclPtr.p; A.exclCol = s.exclCol.p; A.fixed = s.nfixed > 0 ? s.fixedFlag.p : nullptr;
    A.mmq = s.qcMMqDev.p; A.qcq = s.qcPotDev.p + (size_t) nq * (nq + 1); A.images = s.qcImagesDev.p; A.spl = s.qcSplDev.p;
    for (int c = 0; c < 21; c++) A.F.v[c] = s.factors[c];
    A.invDielectric = 1.0 / s.dielectric;
    A.grad = s.qcGradDev.p; A.acc = s.qcAccDev.p; A.pot = s.qcPotDev.p + (size_t) nq * nq; A.V = s.qcPotDev.p;
    C.acc.assign(accCount, 0.0);
    return true;
}

template <int kMode>
static bool qc_launch(State &s, QCContext &C)
{
    const long total = (long) C.A.nimg * C.A.n;
    const dim3 grid((unsigned int) std::max<long>(1, std::min<long>(148 * 4, (total + 255) / 256)), (unsigned int) C.A.nq);
    k_qcmm<kMode><<<grid, 256, 0, s.stream>>>(C.A);
    s.launches += 1;
    NBB_CUDA(cudaGetLastError());
    NBB_CUDA(cudaMemcpyAsync(C.acc.data(), s.qcAccDev.p, sizeof(double) * C.acc.size(), cudaMemcpyDeviceToHost, s.stream));
    NBB_CUDA(cudaStreamSynchronize(s.stream));
    return true;
}

// gradients (host array, accumulated into) and the lattice derivatives of the image terms
static bool qc_finish_gradients(State &s, QCContext &C, double *grad, double *dEdM)
{
    if (grad != nullptr) {
        std::vector<double> hg(3 * (size_t) s.n);
        NBB_CUDA(cudaMemcpyAsync(hg.data(), s.qcGradDev.p, sizeof(double) * hg.size(), cudaMemcpyDeviceToHost, s.stream));
        NBB_CUDA(cudaStreamSynchronize(s.stream));
        for (size_t i = 0; i < hg.size(); i++) grad[i] += hg[i];
    }
    if (dEdM != nullptr) {
        for (size_t k = 1; k < C.images.size(); k++) {
            const QCImageHost &im = C.images[k];
            const double xt[3] = {s.trans.trans[3 * im.t] + (double) im.a, s.trans.trans[3 * im.t + 1] + (double) im.b, s.trans.trans[3 * im.t + 2] + (double) im.c};
            const double *a = C.acc.data() + 4 + 12 * k;
            bool any = false;
            for (int c = 0; c < 12; c++) any = any || a[c] != 0.0;
            if (any) image_derivatives(dEdM, s.lattice, s.trans.rot[im.t], xt, a + 3, a);
        }
    }
    return true;
}

}  // namespace nbb200

using namespace nbb200;

static State *qc_state(NBB200State *state, int *status, const char *who)
{
    if (state == nullptr) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return nullptr; }
    State *s = reinterpret_cast<State *>(state);
    cudaSetDevice(s->device);
    (void) who;
    return s;
}

extern "C" void NBModelABFS_B200_QCMMEnergyLJ(NBB200State *state, double *energies4, double *grad, double *dEdM, int *status)
{
    State *sp = qc_state(state, status, "QCMMEnergyLJ");
    if (sp == nullptr || energies4 == nullptr) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    State &s = *sp;
    for (int c = 0; c < 4; c++) energies4[c] = 0.0;
    if (s.nqc <= 0) return;                                  // no QC atoms: the term is empty (NBModelABFS.c:308)
    if (!s.useAnalytic) { set_error("QCMMEnergyLJ: the spline form of the interaction is not supported"); if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    QCContext C;
    if (!qc_prepare(s, C, "QCMMEnergyLJ", false)) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    if (grad == nullptr && dEdM == nullptr) C.A.grad = nullptr;
    if (!qc_launch<0>(s, C) || !qc_finish_gradients(s, C, grad, dEdM)) { if (status) *status = NBB200_STATUS_LOGIC_ERROR; return; }
    energies4[0] = C.acc[0]; energies4[1] = 0.0; energies4[2] = C.acc[1]; energies4[3] = C.acc[2];     // eqcmmlj, eqcmmlj14, eimqcmmlj, eimqcqclj
}

extern "C" void NBModelABFS_B200_QCMMPotentials(NBB200State *state, double *potentials, double *qcqcPotentials, int *status)
{
    State *sp = qc_state(state, status, "QCMMPotentials");
    if (sp == nullptr) return;
    State &s = *sp;
    if (s.nqc <= 0) return;
    QCContext C;
    if (!qc_prepare(s, C, "QCMMPotentials", true)) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    C.A.grad = nullptr;
    if (!qc_launch<1>(s, C)) { if (status) *status = NBB200_STATUS_LOGIC_ERROR; return; }
    const int nq = C.A.nq;
    std::vector<double> h((size_t) nq * (nq + 1));
    if (!cuda_ok(cudaMemcpyAsync(h.data(), s.qcPotDev.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, s.stream), "D2H") || !cuda_ok(cudaStreamSynchronize(s.stream), "sync")) {
        if (status) *status = NBB200_STATUS_LOGIC_ERROR;
        return;
    }
    // "potentials is incremented not reset" (NBModelABFS.c:1400): the caller's arrays are accumulated into
    if (potentials != nullptr) for (int k = 0; k < nq; k++) potentials[k] += h[(size_t) nq * nq + k];
    if (qcqcPotentials != nullptr) {
        // SymmetricMatrix_IncrementComponent (i, j) and (j, i) address the same packed element: lower triangle, row by row
        for (int a = 0; a < nq; a++) for (int b = 0; b <= a; b++) {
            const double v = (a == b) ? h[(size_t) a * nq + a] : 0.5 * (h[(size_t) a * nq + b] + h[(size_t) b * nq + a]);
            qcqcPotentials[(size_t) a * (a + 1) / 2 + b] += v;
        }
    }
}

extern "C" void NBModelABFS_B200_QCMMGradients(NBB200State *state, const double *qcCharges, double *grad, double *dEdM, int *status)
{
    State *sp = qc_state(state, status, "QCMMGradients");
    if (sp == nullptr || qcCharges == nullptr) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    State &s = *sp;
    if (s.nqc <= 0 || grad == nullptr) return;               // NBModelABFS.c:385: nothing without gradients3
    QCContext C;
    if (!qc_prepare(s, C, "QCMMGradients", true)) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    if (!cuda_ok(cudaMemcpyAsync(const_cast<double *>(C.A.qcq), qcCharges, sizeof(double) * C.A.nq, cudaMemcpyHostToDevice, s.stream), "H2D") ||
        !cuda_ok(cudaStreamSynchronize(s.stream), "sync") || !qc_launch<2>(s, C) || !qc_finish_gradients(s, C, grad, dEdM)) {
        if (status) *status = NBB200_STATUS_LOGIC_ERROR;
    }
}
