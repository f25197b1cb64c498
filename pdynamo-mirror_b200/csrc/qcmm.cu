// qcmm.cu -- QC/MM Lennard-Jones term of NBModelABFS: what NBModelABFS_QCMMEnergyLJ
// (pMolecule-1.9.0/extensions/csource/NBModelABFS.c:306-378) adds for a QC region without boundary atoms in vacuum or in a P1 cell.
//
// The reference walks four lists (nbqcmmlj, nbqcmmlj14, inbqcmmlj, inbqcqclj) with PairwiseInteractionABFS_MMMMEnergy, charges off.
// Every pair within the outer cutoff is on a valid list and pairs beyond it are skipped (PairwiseInteraction.h:72-78), so the sums do
// not depend on the lists; a QC region is tens of atoms, so one fp64 launch over (QC atom) x (all atoms of the cell and of its
// translated copies) does the whole term.  The reference keeps one image of each inverse pair and lists both (QC, MM') and (MM, QC')
// pairs for it: every translation once for the pairs (QC, MM + s); image pairs of two QC atoms count one half per translation.
// Without boundary atoms the 1-4 list of this term is empty.  Restated and pinned in oracle/qcmm_oracle.py.
#include <algorithm>
#include <cmath>
#include "nbb200_internal.h"
#include "../../include/nbabfs_b200.h"

namespace nbb200 {

struct QCFactors { double v[21]; };

static __device__ __forceinline__ double qc_warp_sum(double v)
{
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// grid: (chunks of the (shift, atom) index space, QC atoms); acc[3] = primary, image QC/MM, image QC/QC
static __global__ void k_qcmm_lj(int nq, const int *__restrict__ qcIdx, int n, const double *__restrict__ x, const int *__restrict__ ljtype,
                                 const double2 *__restrict__ ljAB, int ntypes, const unsigned char *__restrict__ qcFlag,
                                 const int *__restrict__ exclPtr, const int *__restrict__ exclCol, QCFactors FF, int nshift,
                                 const double *__restrict__ shifts, double *grad, double *acc)
{
    const double *F = FF.v;
    const int q = qcIdx[blockIdx.y];
    const double xq = x[3 * q], yq = x[3 * q + 1], zq = x[3 * q + 2];
    const int tq = ljtype[q] * ntypes;
    double e[3] = {0.0, 0.0, 0.0}, gq[3] = {0.0, 0.0, 0.0};
    const long total = (long) nshift * n;
    for (long idx = (long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long) gridDim.x * blockDim.x) {
        const int sidx = (int) (idx / n), j = (int) (idx - (long) sidx * n);
        const double dx = xq - x[3 * j] - shifts[3 * sidx], dy = yq - x[3 * j + 1] - shifts[3 * sidx + 1], dz = zq - x[3 * j + 2] - shifts[3 * sidx + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > F[2]) continue;
        const bool jqc = qcFlag[j] != 0;
        if (sidx == 0) {
            if (jqc) continue;                               // pairs inside the QC region belong to the QC model
            bool excluded = false;
            for (int k = exclPtr[q]; k < exclPtr[q + 1]; k++) excluded = excluded || (exclCol[k] == j);
            if (excluded) continue;
        }
        const double2 ab = ljAB[tq + ljtype[j]];
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
        const double w = (sidx > 0 && jqc) ? 0.5 : 1.0;
        e[sidx == 0 ? 0 : (jqc ? 2 : 1)] += w * e2;
        const double gx = w * 2.0 * dF * dx, gy = w * 2.0 * dF * dy, gz = w * 2.0 * dF * dz;
        gq[0] += gx; gq[1] += gy; gq[2] += gz;
        atomicAdd(&grad[3 * j], -gx); atomicAdd(&grad[3 * j + 1], -gy); atomicAdd(&grad[3 * j + 2], -gz);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double es = qc_warp_sum(e[c]), gs = qc_warp_sum(gq[c]);
        if ((threadIdx.x & 31) == 0) {
            if (es != 0.0) atomicAdd(&acc[c], es);
            if (gs != 0.0) atomicAdd(&grad[3 * q + c], gs);
        }
    }
}

}  // namespace nbb200

using namespace nbb200;

extern "C" void NBModelABFS_B200_QCMMEnergyLJ(NBB200State *state, double *energies4, double *grad, int *status)
{
    if (state == nullptr || energies4 == nullptr) { if (status) *status = NBB200_STATUS_INVALID_ARGUMENT; return; }
    State &s = *reinterpret_cast<State *>(state);
    auto fail = [&](int code, const char *msg) { if (msg) set_error(msg); if (status) *status = code; };
    for (int c = 0; c < 4; c++) energies4[c] = 0.0;
    if (s.nqc <= 0) return;                                  // no QC atoms: the term is empty (NBModelABFS.c:308)
    if (s.isNew || s.xcur == nullptr) { fail(NBB200_STATUS_LOGIC_ERROR, "QCMMEnergyLJ before the first Update"); return; }
    if (s.trans.n > 1) { fail(NBB200_STATUS_INVALID_ARGUMENT, "QCMMEnergyLJ: space-group operations are not supported (vacuum and P1 cells only)"); return; }
    if (!s.useAnalytic) { fail(NBB200_STATUS_INVALID_ARGUMENT, "QCMMEnergyLJ: the spline form of the interaction is not supported"); return; }
    if (s.hostLJ64.empty()) { fail(NBB200_STATUS_LOGIC_ERROR, "QCMMEnergyLJ: no fp64 LJ table"); return; }
    cudaSetDevice(s.device);
    std::vector<int> qc;
    for (int i = 0; i < s.n; i++) if (s.hostQC[i]) qc.push_back(i);
    // translations: identity first, then every lattice vector that can bring two atoms within the outer cutoff (atoms may lie up to
    // two cells outside the primary one)
    std::vector<double> shifts = {0.0, 0.0, 0.0};
    if (s.trans.n == 1) {
        int k[3];
        for (int d = 0; d < 3; d++) {
            const double *row = s.lattice.invM.v + 3 * d;
            const double height = 1.0 / std::sqrt(row[0] * row[0] + row[1] * row[1] + row[2] * row[2]);
            k[d] = (int) std::ceil(s.outer / height) + 2;
        }
        for (int a = -k[0]; a <= k[0]; a++) for (int b = -k[1]; b <= k[1]; b++) for (int c = -k[2]; c <= k[2]; c++) {
            if (a == 0 && b == 0 && c == 0) continue;
            for (int r = 0; r < 3; r++) shifts.push_back(s.lattice.M(r, 0) * a + s.lattice.M(r, 1) * b + s.lattice.M(r, 2) * c);
        }
    }
    const int nshift = (int) (shifts.size() / 3), nq = (int) qc.size();
    int *dq = nullptr; double *dsh = nullptr, *dg = nullptr, *dacc = nullptr; double2 *dab = nullptr;
    bool ok = cuda_ok(cudaMalloc((void **) &dq, sizeof(int) * nq), "cudaMalloc") && cuda_ok(cudaMalloc((void **) &dsh, sizeof(double) * shifts.size()), "cudaMalloc") &&
              cuda_ok(cudaMalloc((void **) &dg, sizeof(double) * 3 * (size_t) s.n), "cudaMalloc") && cuda_ok(cudaMalloc((void **) &dacc, sizeof(double) * 4), "cudaMalloc") &&
              cuda_ok(cudaMalloc((void **) &dab, sizeof(double2) * s.hostLJ64.size()), "cudaMalloc");
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    std::vector<double> hg;
    if (ok) {
        cudaStreamSynchronize(s.stream);
        ok = cuda_ok(cudaMemcpy(dq, qc.data(), sizeof(int) * nq, cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(dsh, shifts.data(), sizeof(double) * shifts.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(dab, s.hostLJ64.data(), sizeof(double2) * s.hostLJ64.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemset(dg, 0, sizeof(double) * 3 * (size_t) s.n), "memset") && cuda_ok(cudaMemset(dacc, 0, sizeof(double) * 4), "memset");
    }
    if (ok) {
        QCFactors FF;
        for (int c = 0; c < 21; c++) FF.v[c] = s.factors[c];
        const long total = (long) nshift * s.n;
        const dim3 grid((unsigned int) std::min<long>(148 * 4, (total + 255) / 256), (unsigned int) nq);
        k_qcmm_lj<<<grid, 256>>>(nq, dq, s.n, s.xcur, s.ljtype.p, dab, s.ntypes, s.qcFlag.p, s.exclPtr.p, s.exclCol.p, FF, nshift, dsh, dg, dacc);
        s.launches += 1;
        ok = cuda_ok(cudaGetLastError(), "k_qcmm_lj") && cuda_ok(cudaMemcpy(acc, dacc, sizeof(double) * 4, cudaMemcpyDeviceToHost), "D2H");
        if (ok && grad != nullptr) {
            hg.resize(3 * (size_t) s.n);
            ok = cuda_ok(cudaMemcpy(hg.data(), dg, sizeof(double) * hg.size(), cudaMemcpyDeviceToHost), "D2H");
            if (ok) for (size_t i = 0; i < hg.size(); i++) grad[i] += hg[i];
        }
    }
    cudaFree(dq); cudaFree(dsh); cudaFree(dg); cudaFree(dacc); cudaFree(dab);
    if (!ok) { fail(NBB200_STATUS_LOGIC_ERROR, nullptr); return; }
    energies4[0] = acc[0]; energies4[1] = 0.0; energies4[2] = acc[1]; energies4[3] = acc[2];     // eqcmmlj, eqcmmlj14, eimqcmmlj, eimqcqclj
}
