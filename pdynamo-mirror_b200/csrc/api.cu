// api.cu -- the C-ABI of libnbabfs_b200.so (see include/nbabfs_b200.h for the contract and the reference
// interfaces each entry point replaces).  Host-side control flow mirrors NBModelABFS_Update
// (pMolecule-1.9.0/extensions/csource/NBModelABFS.c:508-623) and NBModelABFS_MMMMEnergy (:228-301).
#include "../../include/nbabfs_b200.h"
#include "nbb200_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

namespace nbb200 {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return false;
}

static void set_status(int *status, int value) { if (status != nullptr) *status = value; }

// true if p is page-locked host memory (cudaMallocHost / cudaHostRegister): DMA needs no staging copy
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// a caller's pageable host array -> device through a page-locked staging array, in a few pieces: the threaded streaming-store copy of piece
// k + 1 (csrc/host_rows.cpp) overlaps the DMA of piece k.  The stream must not be reading the staging array any more.
static bool staged_upload(State &s, double *d_dst, const double *h_src, double *h_stage, long m)
{
    const long pieces = std::max(1L, std::min(4L, m / (1L << 18)));
    for (long p = 0; p < pieces; p++) {
        const long k0 = (m * p) / pieces, k1 = (m * (p + 1)) / pieces;
        nbb200_host_copy(h_stage + k0, h_src + k0, k1 - k0);
        if (!cuda_ok(cudaMemcpyAsync(d_dst + k0, h_stage + k0, sizeof(double) * (size_t) (k1 - k0), cudaMemcpyHostToDevice, s.stream), "H2D")) return false;
    }
    return true;
}

static void destroy(State *s)
{
    if (s == nullptr) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    s->q32.release(); s->q64.release(); s->ljtype.release(); s->ljAB.release(); s->ljAB14.release(); s->typeFree.release(); s->qcIdxDev.release(); s->qcSlotDev.release(); s->qcWork.release(); s->qcImagesDev.release(); s->qcMMqDev.release(); s->qcAccDev.release(); s->qcGradDev.release(); s->qcSplDev.release(); s->qcPotDev.release(); s->qcLJDev.release();
    s->exclPtr.release(); s->exclCol.release(); s->pairs14.release(); s->fixedFlag.release(); s->qcFlag.release();
    s->isoPtr.release(); s->isoIdx.release(); s->xc.release(); s->isoT.release();
    s->x.release(); s->xref.release(); s->grad.release();
    s->imageOps.release(); s->imageBoxes.release(); s->baseOpsDev.release(); s->visitDisp.release(); s->visitInfo.release(); s->bboxDev.release(); s->ticket.release(); s->lcSums.release(); s->lcW.release();
    s->eX.release(); s->eAtom.release(); s->eSet.release(); s->eKey.release(); s->eSortBuf.release();
    s->cellStart.release(); s->cellFill.release(); s->scanTmp.release(); s->order.release(); s->order2.release();
    s->sX.release(); s->sAtom.release(); s->invPerm.release(); s->blockBox.release();
    s->tileDesc.release(); s->tileDescIn.release(); s->itemsIn.release(); s->xprune.release(); s->pruneDisp.release(); s->recA.release(); s->recB.release(); s->gradSorted.release(); s->items.release(); s->rangeTab.release(); s->rangeOut.release(); s->setPairs.release(); s->accum.release();
    s->pairBuf.release(); s->pairCursor.release(); s->splF64.release(); s->splPoly.release(); s->mdScalars.release();
    for (auto &e : s->chunkEvents) if (e != nullptr) cudaEventDestroy(e);
    if (s->sideStream != nullptr) { cudaStreamSynchronize(s->sideStream); cudaStreamDestroy(s->sideStream); cudaEventDestroy(s->evPack); cudaEventDestroy(s->evSide14); }
    for (int r = 0; r < State::kMaxPeers; r++) if (s->peerChunkOpened[r]) { cudaIpcCloseMemHandle(s->peerXc[r]); cudaIpcCloseMemHandle(s->peerGc[r]); }
    for (int r = 0; r < State::kMaxPeers; r++) if (s->peerOpened[r]) { cudaIpcCloseMemHandle(s->peerGs[r]); cudaIpcCloseMemHandle(s->peerXs[r]); cudaIpcCloseMemHandle(s->peerSig[r]); }
    s->symGs.release(); s->symXs.release(); s->symSig.release(); s->sigStage.release();
    if (s->counters) cudaFree(s->counters);
    if (s->hx) cudaFreeHost(s->hx);
    if (s->hgrad) cudaFreeHost(s->hgrad);
    if (s->hsmall) cudaFreeHost(s->hsmall);
    if (s->hacc) cudaFreeHost(s->hacc);
    if (s->hops) cudaFreeHost(s->hops);
    if (s->haveEvents) for (auto &e : s->ev) cudaEventDestroy(e);
    if (s->ownStream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// (kSmallDoubles: nbb200_internal.h)
constexpr size_t kOpsBytes = 4096 * sizeof(ImageOpDev);

static State *create(int device, int n, const double *charges, const int *ljtypes,
                     int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                     int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                     int nexcl, const int *exclPairs, int n14, const int *pairs14, int ntrans, const double *rot, const double *trans, int *status)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device available: libnbabfs_b200 has no CPU fallback"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return nullptr; }
    if (n > kMaxAtoms) { set_error("more than 16.7 M atoms per state are not supported (24-bit tile slots)"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    if (n <= 0 || charges == nullptr || ljtypes == nullptr || ntypes <= 0 || tableindex == nullptr || tableA == nullptr || tableB == nullptr ||
        device < 0 || device >= ndev || (nexcl > 0 && exclPairs == nullptr) || (n14 > 0 && pairs14 == nullptr) || (ntrans > 0 && (rot == nullptr || trans == nullptr))) {
        set_error("invalid argument to NBModelABFSState_B200_SetUp"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr;
    }
    for (int i = 0; i < n; i++) if (ljtypes[i] < 0 || ljtypes[i] >= ntypes) { set_error("ljtype out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    // the reference's LJParameterContainer holds ntypes (ntypes + 1) / 2 entries (a full square table is accepted as well): an index beyond
    // ntypes^2 cannot be a valid one
    for (int i = 0; i < ntypes * ntypes; i++) if (tableindex[i] < 0 || tableindex[i] >= ntypes * ntypes) { set_error("LJ table index out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    if (tableindex14 != nullptr) {
        if (ntypes14 <= 0 || tableA14 == nullptr || tableB14 == nullptr) { set_error("invalid 1-4 LJ table"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
        for (int i = 0; i < ntypes14 * ntypes14; i++) if (tableindex14[i] < 0 || tableindex14[i] >= ntypes14 * ntypes14) { set_error("1-4 LJ table index out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    }
    for (int k = 0; k < 2 * nexcl; k++) if (exclPairs[k] < 0 || exclPairs[k] >= n) { set_error("exclusion index out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    for (int k = 0; k < 2 * n14; k++) if (pairs14[k] < 0 || pairs14[k] >= n) { set_error("1-4 index out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return nullptr; }
    State *s = new (std::nothrow) State();
    if (s == nullptr) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return nullptr; }
    s->device = device;
    bool ok = cuda_ok(cudaSetDevice(device), "cudaSetDevice");
    ok = ok && cuda_ok(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    s->ownStream = true;
    s->n = n; s->ntypes = ntypes; s->nexcl = nexcl; s->n14 = n14;
    if (tableindex14 == nullptr) { ntypes14 = ntypes; tableindex14 = tableindex; tableA14 = tableA; tableB14 = tableB; }
    s->ntypes14 = ntypes14;
    // topology -> device
    std::vector<float> q32(n);
    for (int i = 0; i < n; i++) q32[i] = (float) charges[i];
    std::vector<float2> ab((size_t) ntypes * ntypes);
    for (int i = 0; i < ntypes * ntypes; i++) { ab[i].x = (float) tableA[tableindex[i]]; ab[i].y = (float) tableB[tableindex[i]]; }
    s->hostLJ64.resize((size_t) ntypes * ntypes);
    for (int i = 0; i < ntypes * ntypes; i++) { s->hostLJ64[i].x = tableA[tableindex[i]]; s->hostLJ64[i].y = tableB[tableindex[i]]; }
    std::vector<unsigned char> typeFree((size_t) ntypes + 1, 1);
    for (int i = 0; i < ntypes; i++) for (int j = 0; j < ntypes; j++) if (ab[(size_t) i * ntypes + j].x != 0.f || ab[(size_t) i * ntypes + j].y != 0.f || ab[(size_t) j * ntypes + i].x != 0.f || ab[(size_t) j * ntypes + i].y != 0.f) typeFree[i] = 0;
    std::vector<double2> ab14((size_t) ntypes14 * ntypes14);
    for (int i = 0; i < ntypes14 * ntypes14; i++) { ab14[i].x = tableA14[tableindex14[i]]; ab14[i].y = tableB14[tableindex14[i]]; }
    // symmetric exclusion CSR (SelfPairList_MakeConnections, pCore-1.9.0/extensions/csource/PairList.c:458-526)
    std::vector<int> ptr((size_t) n + 1, 0), col;
    for (int k = 0; k < nexcl; k++) { const int i = exclPairs[2 * k], j = exclPairs[2 * k + 1]; if (i != j) { ptr[i + 1]++; ptr[j + 1]++; } }
    for (int i = 0; i < n; i++) ptr[i + 1] += ptr[i];
    col.resize((size_t) std::max(1, ptr[n]));
    {
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (int k = 0; k < nexcl; k++) { const int i = exclPairs[2 * k], j = exclPairs[2 * k + 1]; if (i != j) { col[cur[i]++] = j; col[cur[j]++] = i; } }
    }
    s->hostExclPtr = ptr; s->hostExclCol.assign(col.begin(), col.begin() + ptr[n]);
    std::vector<int2> p14((size_t) std::max(1, n14));
    for (int k = 0; k < n14; k++) { p14[k].x = pairs14[2 * k]; p14[k].y = pairs14[2 * k + 1]; }
    s->pairs14All.assign(p14.begin(), p14.begin() + n14);
    ok = ok && s->q32.ensure(n) && s->q64.ensure(n) && s->ljtype.ensure(n) && s->ljAB.ensure(ab.size()) && s->ljAB14.ensure(ab14.size()) && s->typeFree.ensure(typeFree.size()) &&
         s->exclPtr.ensure(ptr.size()) && s->exclCol.ensure(col.size()) && s->pairs14.ensure(p14.size()) &&
         s->x.ensure(3 * (size_t) n) && s->xref.ensure(3 * (size_t) n) && s->grad.ensure(3 * (size_t) n);
    ok = ok && cuda_ok(cudaMalloc((void **) &s->counters, sizeof(DeviceCounters)), "cudaMalloc counters");
    ok = ok && cuda_ok(cudaMallocHost((void **) &s->hx, sizeof(double) * 3 * (size_t) n), "cudaMallocHost");
    ok = ok && cuda_ok(cudaMallocHost((void **) &s->hgrad, sizeof(double) * 3 * (size_t) n), "cudaMallocHost");
    ok = ok && cuda_ok(cudaMallocHost((void **) &s->hsmall, sizeof(double) * kSmallDoubles), "cudaMallocHost");
    ok = ok && cuda_ok(cudaMallocHost((void **) &s->hops, kOpsBytes), "cudaMallocHost");
    if (ok) {
        ok = cuda_ok(cudaMemcpy(s->q32.p, q32.data(), sizeof(float) * n, cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->q64.p, charges, sizeof(double) * n, cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->ljtype.p, ljtypes, sizeof(int) * n, cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->ljAB.p, ab.data(), sizeof(float2) * ab.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->ljAB14.p, ab14.data(), sizeof(double2) * ab14.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->typeFree.p, typeFree.data(), typeFree.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->exclPtr.p, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->exclCol.p, col.data(), sizeof(int) * col.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemcpy(s->pairs14.p, p14.data(), sizeof(int2) * p14.size(), cudaMemcpyHostToDevice), "H2D") &&
             cuda_ok(cudaMemset(s->counters, 0, sizeof(DeviceCounters)), "memset") &&
             cuda_ok(cudaDeviceSynchronize(), "set-up copies");   // a pageable H2D copy returns once STAGED; kernels on non-blocking streams must not overtake the DMA
    }
    if (ok) {
        for (auto &e : s->ev) ok = ok && cuda_ok(cudaEventCreate(&e), "cudaEventCreate");
        s->haveEvents = ok;
    }
    if (!ok) { destroy(s); set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return nullptr; }
    if (ntrans > 0) s->trans.set(ntrans, rot, trans);
    make_abfs_factors(s->damp, s->inner, s->outer, s->factors);
    init_force_kernel_attributes();
    return s;
}

static int update_common(State &s, const double *box6, int forceNew, int *status, int decided = -1);

// an optimistic update decision is still open (Update has been called, the energy call that reads the decision has not): entry points that
// hand out the lists themselves settle it first -- wait for the displacement maximum, rebuild when it asks for it
static void settle_optimistic(State &s)
{
    if (!s.optPending) return;
    s.optPending = false;
    double disp = 0.0;
    if (cudaMemcpyAsync(s.hsmall + (kSmallDoubles - 8), s.optDisp.p, sizeof(double), cudaMemcpyDeviceToHost, s.stream) == cudaSuccess &&
        cudaStreamSynchronize(s.stream) == cudaSuccess) disp = s.hsmall[kSmallDoubles - 8];
    else disp = 1.0e300;
    if (disp > s.optThr2) {
        s.numberOfCalls -= 1;
        const bool opt = s.optimistic;
        s.optimistic = false; s.keepLattice = true;
        int st = NBB200_STATUS_CONTINUE;
        update_common(s, nullptr, 1, &st, -1);
        s.keepLattice = false; s.optimistic = opt;
    }
}

static void fetch_pair_counts(State &s)
{
    settle_optimistic(s);
    if (s.pairCountsValid) return;
    std::vector<unsigned long long> cnt((size_t) s.nsets, 0ULL);
    if (s.setPairs.p != nullptr && s.nsets > 0) {
        cudaMemcpyAsync(cnt.data(), s.setPairs.p, sizeof(unsigned long long) * s.nsets, cudaMemcpyDeviceToHost, s.stream);
        cudaStreamSynchronize(s.stream);
    }
    s.primaryPairs = (long) cnt[0];
    s.imagePairs.assign(s.nsets - 1, 0);
    for (int k = 1; k < s.nsets; k++) s.imagePairs[k - 1] = (long) cnt[k];
    s.pairCountsValid = true;
}

// candidate images with at least one pair, in plan order = the reference's ImageList order
static std::vector<int> live_images(State &s)
{
    fetch_pair_counts(s);
    std::vector<int> live;
    for (size_t k = 0; k < s.imagePairs.size(); k++) if (s.imagePairs[k] > 0) live.push_back((int) k);
    return live;
}

static void energy_finish(State &s, double *energies, bool haveGrad, double *dEdM, const double *acc = nullptr, const Lattice *lat = nullptr);

// results of a deferred energy call: valid once the stream has been synchronised after it
static void finish_pending(State &s)
{
    if (!s.pending) return;
    s.pending = false;
    energy_finish(s, s.pendEnergies, s.pendHaveGrad, s.pendDEdM, s.pendAcc != nullptr ? s.pendAcc : s.hacc, &s.pendLattice);
}
static void flush_pending(State &s)
{
    if (!s.pending) return;
    cudaStreamSynchronize(s.stream);
    finish_pending(s);
}

static int update_common(State &s, const double *box6, int forceNew, int *status, int decided)
{
    s.numberOfCalls += 1;
    if (s.pending && (forceNew || s.isNew || decided >= 0 || s.list != s.stListCutoff || s.outer != s.stOuterCutoff)) flush_pending(s);   // no displacement check below
    if (s.trans.n > 0 && !s.keepLattice) {
        if (box6 == nullptr) { set_error("box6 is required when transformations are present"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return 0; }
        s.lattice.set_crystal(box6);
    }
    if (forceNew) s.isNew = true;
    bool doUpdate = s.isNew;
    doUpdate = doUpdate || (s.list != s.stListCutoff) || (s.outer != s.stOuterCutoff);
    if (doUpdate) { s.stListCutoff = s.list; s.stOuterCutoff = s.outer; }
    double maxDisp = 0.0;
    if (s.timing) { s.timings[0] = 0.0; s.timings[3] = 0.0; }
    bool checked = false;
    if (decided >= 0) doUpdate = doUpdate || decided != 0;       // several ranks: the displacement decision was taken collectively by the caller
    s.optPending = false;
    if (!doUpdate && decided < 0 && s.optimistic && s.nranks == 1 && s.nqc == 0 && !s.useCentering && !s.pending && !s.timing &&
        (s.trans.n == 0 || (s.haveRefLattice && std::memcmp(s.lattice.M.v, s.refLattice.M.v, sizeof(double) * 9) == 0))) {
        // optimistic decision: the check is enqueued, its result comes back with the energy call's synchronisation (see State::optimistic)
        const double buffac = 0.5 * (s.list - s.stOuterCutoff);
        // (the maximum travels to the host with the energy call: energy_enqueue)
        if (!s.optDisp.ensure(2) || !displacement_enqueue(s, s.xcur, s.optDisp.p)) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
        s.optDispCopied = false;
        s.optPending = true; s.optThr2 = buffac * buffac;
        s.isNew = false;
        return 0;
    }
    if (!doUpdate && decided < 0) {
        const double buffac = 0.5 * (s.list - s.stOuterCutoff);
        double maxr2 = 0.0; int exceeded = 0;
        if (!displacement_check(s, buffac * buffac, &maxr2, &exceeded)) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
        finish_pending(s);                                   // the check synchronised the stream: a deferred energy call before it is complete
        checked = true;
        doUpdate = exceeded != 0;
        maxDisp = std::sqrt(maxr2);
    }
    if (!doUpdate && s.trans.n > 0 && s.haveRefLattice) {
        fetch_pair_counts(s);
        // NOTE: the reference rebuilds only the image lists here; this implementation rebuilds all lists at the
        // current coordinates (a valid list either way; documented in DESIGN.md).
        doUpdate = check_for_image_update(s.trans, s.lattice, s.refLattice, s.plan.images, s.imagePairs, s.list, s.stOuterCutoff, maxDisp);
    }
    const double *xin = s.xcur;
    if (s.useCentering) {                                    // NBModelABFSState_InitializeCoordinates3: lists and energies see the centred copy
        if (s.nranks > 1) { set_error("useCentering is not available with a partitioned state"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
        if (!centre_coordinates(s, xin, doUpdate)) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return 0; }
        s.xcur = s.xc.p;
    }
    if (doUpdate) {
        if (s.timing) cudaEventRecord(s.ev[0], s.stream);
        if (!cuda_ok(cudaMemcpyAsync(s.xref.p, xin, sizeof(double) * 3 * (size_t) s.n, cudaMemcpyDeviceToDevice, s.stream), "xref copy") || !build_lists(s)) {
            set_status(status, NBB200_STATUS_OUT_OF_MEMORY);
            s.isNew = true;
            return 0;
        }
        if (s.timing) cudaEventRecord(s.ev[1], s.stream);
        s.refLattice = s.lattice; s.haveRefLattice = true;
        s.numberOfUpdates += 1;
    }
    if (s.timing) {
        cudaStreamSynchronize(s.stream);
        float ms = 0.f;
        if (doUpdate) { cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]); s.timings[0] = ms; }
        if (checked) { cudaEventElapsedTime(&ms, s.ev[6], s.ev[7]); s.timings[3] = ms; }
    }
    s.isNew = false;
    return doUpdate ? 1 : 0;
}

// enqueue one energy evaluation on the state's stream: image operations (only when the lattice or the lists changed),
// force kernels, read-back of the accumulators.  No synchronisation here.
static bool energy_enqueue(State &s, double *d_grad, bool sortedOnly = false, double *target = nullptr)
{
    // energy-time image operations: Orthogonalize(S, t + (a,b,c)) with the CURRENT lattice (NBModelABFS.c:1246-1256)
    const bool latticeSame = s.opsValid && s.opsGeneration == s.numberOfUpdates && std::memcmp(s.opsLattice.v, s.lattice.M.v, sizeof(double) * 9) == 0;
    if (!latticeSame) {
        ImageOpDev *ops = reinterpret_cast<ImageOpDev *>(s.hops);
        if ((size_t) s.nsets * sizeof(ImageOpDev) > kOpsBytes) { set_error("too many images for the operation buffer"); return false; }
        std::memset(ops, 0, sizeof(ImageOpDev) * (size_t) s.nsets);
        ops[0].R[0] = ops[0].R[4] = ops[0].R[8] = 1.0; ops[0].scale = 1.0; ops[0].pureTranslation = 1;
        for (int k = 1; k < s.nsets; k++) {
            const CandidateImage &im = s.plan.images[k - 1];
            const double xt[3] = {s.trans.trans[3 * im.t] + (double) im.a, s.trans.trans[3 * im.t + 1] + (double) im.b, s.trans.trans[3 * im.t + 2] + (double) im.c};
            const RealSpaceOp op = orthogonalize(s.trans.rot[im.t], xt, s.lattice);
            std::memcpy(ops[k].R, op.R.v, sizeof(double) * 9);
            std::memcpy(ops[k].tv, op.tv, sizeof(double) * 3);
            ops[k].scale = im.scale; ops[k].pureTranslation = op.pureTranslation ? 1 : 0;
            // the same operation in the grid-relative frame of the atom records (force_kernels.cu)
            for (int r = 0; r < 3; r++) {
                const double *o = s.grid.lo;
                ops[k].cr[r] = op.R.v[3 * r] * o[0] + op.R.v[3 * r + 1] * o[1] + op.R.v[3 * r + 2] * o[2] + op.tv[r] - o[r];
                const double kt = 8.0 * std::nearbyint(op.tv[r] * 0.125);
                ops[k].kt[r] = (float) kt; ops[k].tl[r] = (float) (op.tv[r] - kt);
            }
        }
        if (!s.imageOps.ensure((size_t) s.nsets)) return false;
        NBB_CUDA(cudaMemcpyAsync(s.imageOps.p, ops, sizeof(ImageOpDev) * (size_t) s.nsets, cudaMemcpyHostToDevice, s.stream));
        s.opsLattice = s.lattice.M; s.opsGeneration = s.numberOfUpdates; s.opsValid = true;
    }
    const size_t accumCount = (size_t) 16 * (s.nsets + 1);
    if (accumCount > kSmallDoubles - 16) { set_error("too many images for the result buffer"); return false; }      // the tail holds nbb200_md_run's kinetic-energy slots
    // the accumulators (and the displacement maximum of an optimistic update decision) reach the host through the unsort pass when the call has
    // one (PublishArgs: stores into page-locked memory), through copy operations otherwise
    static const bool noPublish = std::getenv("NBB200_NO_PUBLISH") != nullptr;
    double *hostAcc = target != nullptr ? target : s.hsmall;
    const bool dispOwed = s.optPending && !s.optDispCopied;
    s.pubDone = false;
    if (!noPublish) {
        if (!s.accum.ensure(accumCount + 1)) return false;                  // (launch_forces sizes it the same way)
        s.pubSrc[0] = s.accum.p; s.pubDst[0] = hostAcc; s.pubCount[0] = (int) accumCount;
        s.pubSrc[1] = dispOwed ? s.optDisp.p : nullptr; s.pubDst[1] = s.hsmall + (kSmallDoubles - 8); s.pubCount[1] = dispOwed ? 1 : 0;
    }
    const bool okForces = launch_forces(s, d_grad, sortedOnly);
    s.pubSrc[0] = s.pubSrc[1] = nullptr; s.pubCount[0] = s.pubCount[1] = 0;
    if (!okForces) return false;
    if (!s.pubDone) {
        NBB_CUDA(cudaMemcpyAsync(hostAcc, s.accum.p, sizeof(double) * accumCount, cudaMemcpyDeviceToHost, s.stream));
        if (dispOwed) NBB_CUDA(cudaMemcpyAsync(s.hsmall + (kSmallDoubles - 8), s.optDisp.p, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    }
    if (dispOwed) s.optDispCopied = true;
    return true;
}

// after the stream has been synchronised: energies, dE/dM, timings from the accumulators
static void energy_finish(State &s, double *energies, bool haveGrad, double *dEdM, const double *acc, const Lattice *lat)
{
    if (acc == nullptr) acc = s.hsmall;
    if (lat == nullptr) lat = &s.lattice;
    for (int k = 0; k < 6; k++) energies[k] = 0.0;
    energies[NBB200_EMMEL] = acc[0]; energies[NBB200_EMMLJ] = acc[1];
    for (int k = 1; k < s.nsets; k++) { energies[NBB200_EIMMMEL] += acc[16 * k]; energies[NBB200_EIMMMLJ] += acc[16 * k + 1]; }
    energies[NBB200_EMMEL14] = acc[16 * s.nsets]; energies[NBB200_EMMLJ14] = acc[16 * s.nsets + 1];
    if (dEdM != nullptr && haveGrad) {
        for (int k = 1; k < s.nsets; k++) {
            const CandidateImage &im = s.plan.images[k - 1];
            const double xt[3] = {s.trans.trans[3 * im.t] + (double) im.a, s.trans.trans[3 * im.t + 1] + (double) im.b, s.trans.trans[3 * im.t + 2] + (double) im.c};
            image_derivatives(dEdM, *lat, s.trans.rot[im.t], xt, acc + 16 * k + 5, acc + 16 * k + 2);
        }
    }
    if (s.timing) {
        float ms = 0.f;
        s.timings[1] = s.timings[2] = 0.0;
        if (s.hostCounters.itemCount > 0) { cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]); s.timings[1] = ms; }
        if (s.n14 > 0) { cudaEventElapsedTime(&ms, s.ev[4], s.ev[5]); s.timings[2] = ms; }
        s.timings[5] = 0.0;
        if (s.hostCounters.itemCount > 0 && s.pruneCall > 0 && cudaEventElapsedTime(&ms, s.ev[10], s.ev[11]) == cudaSuccess) s.timings[5] = ms;      // the rolling prune (an almost empty launch when nothing was due)
    }
}

}  // namespace nbb200

using namespace nbb200;

extern "C" {

int nbb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char *nbb200_last_error(void) { return g_error.c_str(); }
const char *nbb200_version(void) { return "nbabfs_b200 0.1 (sm_100a)"; }

NBB200State *NBModelABFSState_B200_SetUp(int device, int n, const double *charges, const int *ljtypes,
                                         int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                                         int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                                         int nexcl, const int *exclPairs, int n14, const int *pairs14,
                                         int ntrans, const double *rot, const double *trans, int *status)
{
    return reinterpret_cast<NBB200State *>(create(device, n, charges, ljtypes, ntypes, tableindex, tableA, tableB, ntypes14, tableindex14, tableA14, tableB14,
                                                  nexcl, exclPairs, n14, pairs14, ntrans, rot, trans, status));
}

void NBModelABFSState_B200_Deallocate(NBB200State **state)
{
    if (state == nullptr || *state == nullptr) return;
    destroy(reinterpret_cast<State *>(*state));
    *state = nullptr;
}

void NBModelABFSState_B200_SetFixedAtoms(NBB200State *state, int nfixed, const int *fixed, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (nfixed < 0 || (nfixed > 0 && fixed == nullptr)) { set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
    std::vector<unsigned char> flag((size_t) s.n, 0);
    for (int k = 0; k < nfixed; k++) {
        if (fixed[k] < 0 || fixed[k] >= s.n) { set_error("fixed atom index out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
        flag[fixed[k]] = 1;
    }
    // GenerateLists14 -> SelfPairList_FromSelfPairList(interactions14, mmSelection, freeSelection): at least one free atom
    std::vector<int2> keep;
    for (const int2 &p : s.pairs14All) if ((nfixed == 0 || !(flag[p.x] && flag[p.y])) && (s.hostQC.empty() || !(s.hostQC[p.x] || s.hostQC[p.y]))) keep.push_back(p);
    bool ok = s.fixedFlag.ensure((size_t) s.n) && s.pairs14.ensure(std::max<size_t>(1, keep.size()));
    cudaStreamSynchronize(s.stream);
    ok = ok && cuda_ok(cudaMemcpy(s.fixedFlag.p, flag.data(), (size_t) s.n, cudaMemcpyHostToDevice), "H2D fixed") &&
         (keep.empty() || cuda_ok(cudaMemcpy(s.pairs14.p, keep.data(), sizeof(int2) * keep.size(), cudaMemcpyHostToDevice), "H2D 1-4")) &&
         cuda_ok(cudaDeviceSynchronize(), "set-up copies");
    if (!ok) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return; }
    s.nfixed = nfixed; s.n14 = (int) keep.size();
    s.hostFixed = (nfixed > 0) ? flag : std::vector<unsigned char>();
    s.isNew = true;
}

/* Pure QC atoms (NBModelABFSState_SetUp with qcAtoms, NBModelABFSState.c:348-353: mmSelection = complement of the pure QC
 * selection).  They leave every MM/MM list -- primary, image (GenerateLists / GenerateImageLists and-selection, NBModelABFS.c:508-623) and
 * 1-4 (GenerateLists14) -- so that NBModelABFS_B200_MMMMEnergy returns what the reference's NBModelABFS_MMMMEnergy returns with a QC region
 * present.  The QC/MM entry points themselves are NBModelABFS_B200_QCMMEnergyLJ / _QCMMPotentials / _QCMMGradients (qcmm.cu); boundary (link)
 * atoms are not handled. */
void NBModelABFSState_B200_SetQCAtoms(NBB200State *state, int nqc, const int *qcAtoms, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (nqc < 0 || (nqc > 0 && qcAtoms == nullptr)) { set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
    if (nqc > 0 && (s.useCentering || s.nranks > 1)) { set_error("QC atoms with useCentering or several partitions are not supported"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
    std::vector<unsigned char> flag((size_t) s.n, 0);
    for (int k = 0; k < nqc; k++) {
        if (qcAtoms[k] < 0 || qcAtoms[k] >= s.n) { set_error("QC atom index out of range"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
        flag[qcAtoms[k]] = 1;
    }
    std::vector<int2> keep;
    for (const int2 &p : s.pairs14All) if ((nqc == 0 || !(flag[p.x] || flag[p.y])) && (s.hostFixed.empty() || !(s.hostFixed[p.x] && s.hostFixed[p.y]))) keep.push_back(p);
    bool ok = s.qcFlag.ensure((size_t) s.n) && s.pairs14.ensure(std::max<size_t>(1, keep.size()));
    cudaStreamSynchronize(s.stream);
    ok = ok && cuda_ok(cudaMemcpy(s.qcFlag.p, flag.data(), (size_t) s.n, cudaMemcpyHostToDevice), "H2D QC flags") &&
         (keep.empty() || cuda_ok(cudaMemcpy(s.pairs14.p, keep.data(), sizeof(int2) * keep.size(), cudaMemcpyHostToDevice), "H2D 1-4")) &&
         cuda_ok(cudaDeviceSynchronize(), "set-up copies");
    if (!ok) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return; }
    s.nqc = nqc; s.n14 = (int) keep.size();
    s.hostQC = (nqc > 0) ? flag : std::vector<unsigned char>();
    s.isNew = true;
}

void NBModelABFSState_B200_SetUpCentering(NBB200State *state, int useCentering, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    s.useCentering = false; s.nisolates = 0; s.isNew = true;
    // as the reference: needs exclusions, transformations and the option; silently off otherwise (NBModelABFSState.c:427)
    if (!useCentering || s.trans.n <= 0 || s.hostExclCol.empty()) return;
    const int n = s.n;
    // SelfPairList_ToIsolateSelectionContainer (pC/csource/PairList.c:686-770): breadth-first components in index order, members sorted;
    // SelectionContainer_RemoveIsolates(fixedAtoms): molecules with a fixed atom do not move
    std::vector<int> ptr, idx((size_t) n);
    std::vector<char> assigned((size_t) n, 0);
    int m = 0;
    for (int a = 0; a < n; a++) {
        if (assigned[a]) continue;
        const int start = m;
        idx[m++] = a; assigned[a] = 1;
        for (int i = start; i < m; i++)
            for (int c = s.hostExclPtr[idx[i]]; c < s.hostExclPtr[idx[i] + 1]; c++) { const int j = s.hostExclCol[c]; if (!assigned[j]) { idx[m++] = j; assigned[j] = 1; } }
        std::sort(idx.begin() + start, idx.begin() + m);
        bool keep = true;
        if (!s.hostFixed.empty()) for (int i = start; i < m; i++) if (s.hostFixed[idx[i]]) keep = false;
        if (keep) ptr.push_back(start); else m = start;
    }
    const int niso = (int) ptr.size();
    ptr.push_back(m);
    if (niso <= 1) return;
    bool ok = s.isoPtr.ensure(ptr.size()) && s.isoIdx.ensure((size_t) std::max(1, m));
    cudaStreamSynchronize(s.stream);
    ok = ok && cuda_ok(cudaMemcpy(s.isoPtr.p, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice), "H2D isolates") &&
         cuda_ok(cudaMemcpy(s.isoIdx.p, idx.data(), sizeof(int) * (size_t) m, cudaMemcpyHostToDevice), "H2D isolates") &&
         cuda_ok(cudaDeviceSynchronize(), "set-up copies");
    if (!ok) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return; }
    s.nisolates = niso; s.useCentering = true;
}

void NBModelABFS_B200_SetOptions(NBB200State *state, double dampingCutoff, double innerCutoff, double outerCutoff, double listCutoff,
                                 double dielectric, double electrostaticScale14, int checkForInverses, int imageExpandFactor)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    const bool cutoffsChanged = s.damp != dampingCutoff || s.inner != innerCutoff || s.outer != outerCutoff;
    s.damp = dampingCutoff; s.inner = innerCutoff; s.outer = outerCutoff; s.list = listCutoff;
    s.dielectric = dielectric; s.scale14 = electrostaticScale14;
    if ((s.checkForInverses != (checkForInverses != 0)) || s.expandFactor != imageExpandFactor) s.isNew = true;
    s.checkForInverses = checkForInverses != 0; s.expandFactor = imageExpandFactor;
    make_abfs_factors(s.damp, s.inner, s.outer, s.factors);
    if (!s.useAnalytic && (cutoffsChanged || !s.splValid)) {  // the splines depend on the cutoffs (MakeSplines, pMolecule.NBModelABFS.pyx:82)
        cudaSetDevice(s.device);
        make_abfs_splines(s.damp, s.inner, s.outer, s.splineDensity, s.spl);
        s.splValid = upload_spline_tables(s);
    }
}

void PairwiseInteractionABFS_B200_SetInteractionForm(NBB200State *state, int useAnalyticForm, int splinePointDensity, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    if (!useAnalyticForm && splinePointDensity < 1) { set_error("splinePointDensity must be positive"); set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return; }
    s.useAnalytic = useAnalyticForm != 0;
    s.splineDensity = splinePointDensity;
    s.splValid = false;
    if (!s.useAnalytic) {
        cudaSetDevice(s.device);
        make_abfs_splines(s.damp, s.inner, s.outer, s.splineDensity, s.spl);
        s.splValid = upload_spline_tables(s);
        if (!s.splValid) set_status(status, NBB200_STATUS_LOGIC_ERROR);
    }
}

int PairwiseInteractionABFS_B200_MakeSpline(int which, double dampingCutoff, double innerCutoff, double outerCutoff, int splinePointDensity,
                                            double *x, double *y, double *h)
{
    if (which < 0 || which > 3 || splinePointDensity < 1) return 0;
    const int n = abfs_spline_points(outerCutoff, splinePointDensity);
    if (x == nullptr || y == nullptr || h == nullptr) return n;
    if (which == 3) {                                        // electrostatic spline in atomic units (the QC/MM and QC/QC interactions)
        std::vector<double> vx, vy, vh;
        make_abfs_electrostatic_spline_au(dampingCutoff, innerCutoff, outerCutoff, splinePointDensity, vx, vy, vh);
        for (int i = 0; i < n; i++) { x[i] = vx[i]; y[i] = vy[i]; h[i] = vh[i]; }
        return n;
    }
    SplineTables t;
    make_abfs_splines(dampingCutoff, innerCutoff, outerCutoff, splinePointDensity, t);
    for (int i = 0; i < n; i++) { x[i] = t.x[i]; y[i] = t.y[which][i]; h[i] = t.h[which][i]; }
    return n;
}

int NBModelABFS_B200_Update(NBB200State *state, const double *xyz, const double *box6, int forceNew, int *status)
{
    if (state == nullptr || xyz == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.hostXChecked != (const void *) xyz) { s.hostXChecked = xyz; s.hostXPinned = is_pinned_host(xyz); }       // (the query is slow for pageable memory)
    bool up;
    if (s.hostXPinned) up = cuda_ok(cudaMemcpyAsync(s.x.p, xyz, sizeof(double) * 3 * (size_t) s.n, cudaMemcpyHostToDevice, s.stream), "H2D coordinates");
    else { cudaStreamSynchronize(s.stream); up = staged_upload(s, s.x.p, xyz, s.hx, 3 * (long) s.n); }
    if (!up) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
    s.xcur = s.x.p;
    return update_common(s, box6, forceNew, status);
}

int NBModelABFS_B200_UpdateDevice(NBB200State *state, const double *d_xyz, const double *box6, int forceNew, int *status)
{
    if (state == nullptr || d_xyz == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    s.xcur = d_xyz;
    return update_common(s, box6, forceNew, status);
}

void NBModelABFS_B200_MMMMEnergy(NBB200State *state, double *energies, double *grad, double *dEdM, int *status)
{
    if (state == nullptr || energies == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.xcur == nullptr) { set_error("MMMMEnergy called before Update"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    flush_pending(s);
    double *dg = nullptr;
    const bool direct = grad != nullptr && is_pinned_host(grad);       // accumulate on the device, DMA straight into the caller's array
    const bool upload = direct && !s.gradOverwrite;                      // overwrite mode: the caller's values are not needed at all
    const size_t gbytes = sizeof(double) * 3 * (size_t) s.n;
    if (grad != nullptr) {
        dg = s.grad.p;
        const bool ok0 = upload ? cuda_ok(cudaMemcpyAsync(dg, grad, gbytes, cudaMemcpyHostToDevice, s.stream), "H2D grad")
                                : cuda_ok(cudaMemsetAsync(dg, 0, gbytes, s.stream), "memset grad");
        if (!ok0) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    }
    // optimistic update decision (nbb200_set_optimistic_updates): the displacement check of the Update call comes back with this call's results
    if (s.optPending) { s.condDisp = s.optDisp.p; s.condThr2 = s.optThr2; }
    bool ok = energy_enqueue(s, dg);
    if (ok && grad != nullptr) ok = cuda_ok(cudaMemcpyAsync(direct ? grad : s.hgrad, dg, gbytes, cudaMemcpyDeviceToHost, s.stream), "D2H grad");
    ok = ok && cuda_ok(cudaStreamSynchronize(s.stream), "sync");          // the one synchronisation of the call
    s.condDisp = nullptr;
    if (ok && s.optPending) {
        s.optPending = false;
        if (s.hsmall[kSmallDoubles - 8] > s.optThr2) {
            // an update was due: the unsort pass left the staging gradient as it was prepared above (the caller's values, or zeros); rebuild at
            // these coordinates and evaluate again
            s.numberOfCalls -= 1;
            const bool opt = s.optimistic;
            s.optimistic = false;
            int st = NBB200_STATUS_CONTINUE;
            s.keepLattice = true;                       // the lattice of the Update call that is being completed
            update_common(s, nullptr, 1, &st);
            s.keepLattice = false;
            s.optimistic = opt;
            ok = st == NBB200_STATUS_CONTINUE && energy_enqueue(s, dg);
            if (ok && grad != nullptr) ok = cuda_ok(cudaMemcpyAsync(direct ? grad : s.hgrad, dg, gbytes, cudaMemcpyDeviceToHost, s.stream), "D2H grad");
            ok = ok && cuda_ok(cudaStreamSynchronize(s.stream), "sync");
        }
    }
    if (ok) {
        energy_finish(s, energies, grad != nullptr, dEdM);
        if (grad != nullptr && !direct) {
            const size_t m = 3 * (size_t) s.n;
            if (s.gradOverwrite) std::memcpy(grad, s.hgrad, sizeof(double) * m);
            else nbb200_host_add(grad, s.hgrad, (long) m);
        }
    }
    if (!ok) set_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void NBModelABFS_B200_MMMMEnergyDevice(NBB200State *state, double *energies, double *d_grad, double *dEdM, int *status)
{
    if (state == nullptr || energies == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.xcur == nullptr) { set_error("MMMMEnergy called before Update"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    flush_pending(s);
    if (s.optPending) { s.condDisp = s.optDisp.p; s.condThr2 = s.optThr2; }
    bool ok = energy_enqueue(s, d_grad) && cuda_ok(cudaStreamSynchronize(s.stream), "sync");
    s.condDisp = nullptr;
    if (ok && s.optPending) {
        s.optPending = false;
        if (s.hsmall[kSmallDoubles - 8] > s.optThr2) {
            // an update was due: nothing of the evaluation above has been handed out; rebuild at these coordinates and evaluate again
            s.numberOfCalls -= 1;
            const bool opt = s.optimistic;
            s.optimistic = false;
            int st = NBB200_STATUS_CONTINUE;
            s.keepLattice = true;                       // the lattice of the Update call that is being completed
            update_common(s, nullptr, 1, &st);
            s.keepLattice = false;
            s.optimistic = opt;
            ok = st == NBB200_STATUS_CONTINUE && energy_enqueue(s, d_grad) && cuda_ok(cudaStreamSynchronize(s.stream), "sync");
        }
    }
    if (ok) energy_finish(s, energies, d_grad != nullptr, dEdM);
    else set_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void NBModelABFS_B200_MMMMEnergyDeviceDeferred(NBB200State *state, double *energies, double *d_grad, double *dEdM, int *status)
{
    if (state == nullptr || energies == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.xcur == nullptr) { set_error("MMMMEnergy called before Update"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    flush_pending(s);
    if (s.hacc == nullptr && !cuda_ok(cudaMallocHost((void **) &s.hacc, sizeof(double) * kSmallDoubles), "cudaMallocHost")) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return; }
    if (!energy_enqueue(s, d_grad, false, s.hacc)) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    s.pending = true; s.pendEnergies = energies; s.pendDEdM = dEdM; s.pendHaveGrad = d_grad != nullptr; s.pendLattice = s.lattice; s.pendAcc = s.hacc;
}

void nbb200_flush(NBB200State *state, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (!cuda_ok(cudaStreamSynchronize(s.stream), "sync")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    finish_pending(s);
}

void nbb200_copy_to_host_async(NBB200State *state, const void *d_src, void *h_pinned_dst, size_t bytes)
{
    if (state == nullptr || d_src == nullptr || h_pinned_dst == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    cuda_ok(cudaMemcpyAsync(h_pinned_dst, d_src, bytes, cudaMemcpyDeviceToHost, s.stream), "D2H");
}

long NBModelABFSState_B200_NumberOfPairs(NBB200State *state, int image)
{
    if (state == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (image < 0) { fetch_pair_counts(s); return s.primaryPairs; }
    const std::vector<int> live = live_images(s);
    if (image >= (int) live.size()) return 0;
    return s.imagePairs[live[image]];
}

int NBModelABFSState_B200_NumberOfImages(NBB200State *state)
{
    if (state == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    return (int) live_images(s).size();
}

long NBModelABFSState_B200_NumberOfImagePairs(NBB200State *state)
{
    if (state == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    fetch_pair_counts(s);
    long t = 0;
    for (long v : s.imagePairs) t += v;
    return t;
}

long NBModelABFSState_B200_NumberOf14Pairs(NBB200State *state)
{
    return (state == nullptr) ? 0 : (long) reinterpret_cast<State *>(state)->n14;
}

void NBModelABFSState_B200_GetImageInfo(NBB200State *state, int image, int *info, double *scale)
{
    if (state == nullptr || info == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const std::vector<int> live = live_images(s);
    if (image < 0 || image >= (int) live.size()) return;
    const CandidateImage &im = s.plan.images[live[image]];
    info[0] = im.t; info[1] = im.a; info[2] = im.b; info[3] = im.c; info[4] = (int) s.imagePairs[live[image]]; info[5] = 0;
    if (scale != nullptr) *scale = im.scale;
}

long NBModelABFSState_B200_GetPairs(NBB200State *state, int image, int *pairs, int *status)
{
    if (state == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    settle_optimistic(s);
    if (!expand_pairs(s)) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return -1; }
    int set = 0;
    if (image >= 0) {
        const std::vector<int> live = live_images(s);
        if (image >= (int) live.size()) return 0;
        set = 1 + live[image];
    }
    const unsigned long long lo = s.pairOffsets[set], hi = s.pairOffsets[set + 1];
    if (pairs != nullptr && hi > lo) {
        if (!cuda_ok(cudaMemcpy(pairs, s.pairBuf.p + 2 * lo, sizeof(int) * 2 * (hi - lo), cudaMemcpyDeviceToHost), "D2H pairs")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return -1; }
    }
    return (long) (hi - lo);
}

static long standalone(int device, int n1, const double *xyz1, int n2, const double *xyz2, double cutoff, int nexcl, const int *excl, int **pairs, int *status)
{
    if (pairs == nullptr || xyz1 == nullptr || n1 <= 0 || cutoff <= 0.0) { set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return -1; }
    *pairs = nullptr;
    std::vector<double> q((size_t) n1, 0.0);
    std::vector<int> lt((size_t) n1, 0);
    const int ti[1] = {0};
    const double ta[1] = {0.0};
    State *s = create(device, n1, q.data(), lt.data(), 1, ti, ta, ta, 1, ti, ta, ta, nexcl, excl, 0, nullptr, 0, nullptr, nullptr, status);
    if (s == nullptr) return -1;
    long result = -1;
    DevBuf<double> x2;
    bool ok = true;
    s->list = cutoff;
    ok = cuda_ok(cudaMemcpy(s->x.p, xyz1, sizeof(double) * 3 * (size_t) n1, cudaMemcpyHostToDevice), "H2D");
    s->xcur = s->x.p;
    if (ok && xyz2 != nullptr) {
        ok = n2 > 0 && x2.ensure(3 * (size_t) n2) && cuda_ok(cudaMemcpy(x2.p, xyz2, sizeof(double) * 3 * (size_t) n2, cudaMemcpyHostToDevice), "H2D");
    }
    ok = ok && cuda_ok(cudaDeviceSynchronize(), "input copies") && build_lists_standalone(*s, xyz2 != nullptr ? x2.p : nullptr, n2) && expand_pairs(*s);
    if (ok) {
        const int set = (xyz2 != nullptr) ? 1 : 0;
        const unsigned long long lo = s->pairOffsets[set], hi = s->pairOffsets[set + 1];
        int *out = (int *) std::malloc(sizeof(int) * 2 * (size_t) std::max<unsigned long long>(1ULL, hi - lo));
        if (out != nullptr && (hi == lo || cuda_ok(cudaMemcpy(out, s->pairBuf.p + 2 * lo, sizeof(int) * 2 * (hi - lo), cudaMemcpyDeviceToHost), "D2H pairs"))) {
            *pairs = out; result = (long) (hi - lo);
        } else { std::free(out); set_status(status, NBB200_STATUS_OUT_OF_MEMORY); }
    } else set_status(status, NBB200_STATUS_OUT_OF_MEMORY);
    x2.release();
    destroy(s);
    return result;
}

long PairListGenerator_B200_SelfPairListFromCoordinates3(int device, int n, const double *xyz, double cutoff, int nexcl, const int *exclPairs, int **pairs, int *status)
{
    return standalone(device, n, xyz, 0, nullptr, cutoff, nexcl, exclPairs, pairs, status);
}

long PairListGenerator_B200_CrossPairListFromDoubleCoordinates3(int device, int n1, const double *xyz1, int n2, const double *xyz2, double cutoff, int **pairs, int *status)
{
    if (xyz2 == nullptr) { set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return -1; }
    return standalone(device, n1, xyz1, n2, xyz2, cutoff, 0, nullptr, pairs, status);
}

void nbb200_free(void *p) { std::free(p); }

void *nbb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (!cuda_ok(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost")) return nullptr;
    return p;
}

void nbb200_host_free(void *p) { if (p != nullptr) cudaFreeHost(p); }

void PairwiseInteractionABFS_B200_MakeFactors(double dampingCutoff, double innerCutoff, double outerCutoff, double *out21)
{
    if (out21 != nullptr) make_abfs_factors(dampingCutoff, innerCutoff, outerCutoff, out21);
}

void nbb200_set_stream(NBB200State *state, void *cudaStream)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.ownStream && s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    s.stream = reinterpret_cast<cudaStream_t>(cudaStream);
    s.ownStream = false;
}

void nbb200_enable_timing(NBB200State *state, int on) { if (state != nullptr) reinterpret_cast<State *>(state)->timing = on != 0; }

void nbb200_get_timings(NBB200State *state, double *out8)
{
    if (state == nullptr || out8 == nullptr) return;
    std::memcpy(out8, reinterpret_cast<State *>(state)->timings, sizeof(double) * 8);
}

void nbb200_get_counters(NBB200State *state, long *out8)
{
    if (state == nullptr || out8 == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    fetch_pair_counts(s);
    long pairs = s.primaryPairs;
    for (long v : s.imagePairs) pairs += v;
    out8[0] = s.hostCounters.tilesUsed; out8[1] = s.hostCounters.itemCount; out8[2] = s.nblocks; out8[3] = s.hostCounters.extCount;
    out8[4] = pairs; out8[5] = s.launches; out8[6] = s.chunkTiles; out8[7] = (long) live_images(s).size();
}

/* ---- section 8e: sorted-space slabs and halo ranges ---- */
void nbb200_get_slab(NBB200State *state, long *out4)
{
    if (state == nullptr || out4 == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    out4[0] = s.ownLo; out4[1] = s.ownHi; out4[2] = s.n; out4[3] = s.nblocks;
}

void nbb200_set_sorted_gradient_buffer(NBB200State *state, double *d_buf)
{
    if (state != nullptr) reinterpret_cast<State *>(state)->gsExternal = d_buf;
}

int nbb200_touched_ranges(NBB200State *state, long *out)
{
    if (state == nullptr || out == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    return touched_ranges(s, out) ? 1 : 0;
}

int nbb200_touched_ranges_device(NBB200State *state, long *d_out)
{
    if (state == nullptr || d_out == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    return touched_ranges_async(s, d_out) ? 1 : 0;
}

double nbb200_max_displacement(NBB200State *state, const double *d_xyz, int *status)
{
    if (state == nullptr || d_xyz == nullptr) return 0.0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.isNew) return 0.0;                                   // no reference coordinates yet
    const double *keep = s.xcur;
    s.xcur = d_xyz;
    double maxr2 = 0.0; int exceeded = 0;
    const bool ok = displacement_check(s, 0.0, &maxr2, &exceeded);
    s.xcur = keep;
    if (!ok) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0.0; }
    return maxr2;
}

int NBModelABFS_B200_UpdateDeviceDecided(NBB200State *state, const double *d_xyz, const double *box6, int doUpdate, int *status)
{
    if (state == nullptr || d_xyz == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    s.xcur = d_xyz;
    return update_common(s, box6, 0, status, doUpdate != 0 ? 1 : 0);
}

void NBModelABFS_B200_MMMMEnergySorted(NBB200State *state, double *energies, double *dEdM, int *status)
{
    if (state == nullptr || energies == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.xcur == nullptr) { set_error("MMMMEnergy called before Update"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    flush_pending(s);
    if (energy_enqueue(s, nullptr, true) && cuda_ok(cudaStreamSynchronize(s.stream), "sync")) energy_finish(s, energies, true, dEdM);
    else set_status(status, NBB200_STATUS_LOGIC_ERROR);
}

/* the same call in two halves, so that work that only depends on the kernels (the gradient push to the peers) can be enqueued before the
 * host waits for the energies */
void NBModelABFS_B200_MMMMEnergySortedEnqueue(NBB200State *state, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.xcur == nullptr) { set_error("MMMMEnergy called before Update"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    flush_pending(s);
    if (!energy_enqueue(s, nullptr, true)) set_status(status, NBB200_STATUS_LOGIC_ERROR);
}

void NBModelABFS_B200_MMMMEnergySortedFinish(NBB200State *state, double *energies, double *dEdM, int *status)
{
    if (state == nullptr || energies == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (cuda_ok(cudaStreamSynchronize(s.stream), "sync")) energy_finish(s, energies, true, dEdM);
    else set_status(status, NBB200_STATUS_LOGIC_ERROR);
}

static __global__ void k_gather_sorted_x(const double *__restrict__ x, const int *__restrict__ sAtom, long s0, long count, double *__restrict__ out)
{
    const long k = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int a = sAtom[s0 + k];
    out[3 * k] = x[3 * a]; out[3 * k + 1] = x[3 * a + 1]; out[3 * k + 2] = x[3 * a + 2];
}

// the own slab as the peers read it: positions by sorted position and, behind them, the atom index of every position (a peer that
// sorts only its own neighbourhood has no valid sAtom for this slab)
static __global__ void k_publish_slab(const double *__restrict__ x, const int *__restrict__ sAtom, long s0, long count, double *__restrict__ xs, int *__restrict__ ids)
{
    const long k = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int a = sAtom[s0 + k];
    xs[3 * (s0 + k)] = x[3 * a]; xs[3 * (s0 + k) + 1] = x[3 * a + 1]; xs[3 * (s0 + k) + 2] = x[3 * a + 2];
    ids[s0 + k] = a;
}

static __global__ void k_scatter_sorted_x(const double *__restrict__ in, const int *__restrict__ sAtom, long s0, long count, double *__restrict__ x)
{
    const long k = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int a = sAtom[s0 + k];
    x[3 * a] = in[3 * k]; x[3 * a + 1] = in[3 * k + 1]; x[3 * a + 2] = in[3 * k + 2];
}

void nbb200_gather_sorted(NBB200State *state, const double *d_x, long s0, long count, double *d_out)
{
    if (state == nullptr || count <= 0) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    k_gather_sorted_x<<<(unsigned int) ((count + 255) / 256), 256, 0, s.stream>>>(d_x, s.sAtom.p, s0, count, d_out);
    s.launches += 1;
}

void nbb200_scatter_sorted(NBB200State *state, const double *d_in, long s0, long count, double *d_x)
{
    if (state == nullptr || count <= 0) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    k_scatter_sorted_x<<<(unsigned int) ((count + 255) / 256), 256, 0, s.stream>>>(d_in, s.sAtom.p, s0, count, d_x);
    s.launches += 1;
}

void nbb200_unsort_add(NBB200State *state, long s0, long count, double *d_grad)
{
    if (state == nullptr || count <= 0) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    unsort_gradients(s, s0, s0 + count, d_grad);
}

/* ---- peer memory: the halo exchanges as plain kernels over NVLink (no NCCL call on the data path) ---- */
// the exported position buffer: 3 n doubles (by sorted position), then -- 16-byte aligned -- n ints: the atom index of every position
__host__ __device__ static inline long xs_ids_offset(long n) { return (3 * n + 1) & ~1L; }
struct PeerPtrs { double *p[State::kMaxPeers]; };
struct SlabEdges { long s[State::kMaxPeers + 1]; };

// positions, step 1: copy what this rank needs from rank r = blockIdx.y >> 1 -- the whole slab of r (list rebuild: slab edges given)
// with its atom indices, or the halo range (r, h = blockIdx.y & 1) of the device table -- into the SAME sorted positions of the own
// buffer: flat, coalesced 8-byte reads over NVLink (reading the three coordinates of an atom per thread fetched every sector three times)
static __global__ void k_peer_copy(PeerPtrs xs, SlabEdges edges, const long *__restrict__ table, int rank, int nranks, int wholeSlabs, long n, double *__restrict__ mine)
{
    const int r = blockIdx.y >> 1, h = blockIdx.y & 1;
    if (r == rank) return;
    long lo, hi;
    if (wholeSlabs) { if (h) return; lo = edges.s[r]; hi = edges.s[r + 1]; }
    else { const long *t = table + ((long) r * 2 + h) * 2; lo = t[0]; hi = t[1]; }
    const double *src = xs.p[r];
    const long stride = (long) gridDim.x * blockDim.x, t0 = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (!wholeSlabs) {                                       // halo ranges: a few thousand atoms, arbitrary alignment
        for (long k = 3 * lo + t0; k < 3 * hi; k += stride) mine[k] = src[k];
        return;
    }
    // whole slabs start at a multiple of 32 atoms (768 bytes): 16-byte loads, four in flight per thread (the NVLink round trip is what
    // limits a one-load-at-a-time loop)
    const long v0 = 3 * lo / 2, v1 = 3 * hi / 2;             // double2 units; 3 * lo is even
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(mine);
    long k = v0 + t0;
    for (; k + 3 * stride < v1; k += 4 * stride) {
        const double2 a = s2[k], b = s2[k + stride], c = s2[k + 2 * stride], d = s2[k + 3 * stride];
        d2[k] = a; d2[k + stride] = b; d2[k + 2 * stride] = c; d2[k + 3 * stride] = d;
    }
    for (; k < v1; k += stride) d2[k] = s2[k];
    if (t0 == 0 && ((3 * hi) & 1)) mine[3 * hi - 1] = src[3 * hi - 1];
    const int *sid = reinterpret_cast<const int *>(src + xs_ids_offset(n));
    int *did = reinterpret_cast<int *>(mine + xs_ids_offset(n));
    const long w0 = lo / 4, w1 = hi / 4;                     // int4 units; lo is a multiple of 4
    const int4 *s4 = reinterpret_cast<const int4 *>(sid);
    int4 *d4 = reinterpret_cast<int4 *>(did);
    for (long q = w0 + t0; q < w1; q += stride) d4[q] = s4[q];
    for (long q = 4 * w1 + t0; q < hi; q += stride) did[q] = sid[q];
}

// step 2 (local): x[atom(s)] = xs[s]; atom(s) from the owner's indices (whole slabs) or the own sort (halo ranges: always inside the cells this rank sorts)
static __global__ void k_peer_scatter(const double *__restrict__ mine, SlabEdges edges, const long *__restrict__ table, int rank, int nranks, int wholeSlabs, long n,
                                      const int *__restrict__ sAtom, double *__restrict__ x)
{
    const int r = blockIdx.y >> 1, h = blockIdx.y & 1;
    if (r == rank) return;
    long lo, hi;
    if (wholeSlabs) { if (h) return; lo = edges.s[r]; hi = edges.s[r + 1]; }
    else { const long *t = table + ((long) r * 2 + h) * 2; lo = t[0]; hi = t[1]; }
    const int *ids = wholeSlabs ? reinterpret_cast<const int *>(mine + xs_ids_offset(n)) : sAtom;
    for (long s = lo + (long) blockIdx.x * blockDim.x + threadIdx.x; s < hi; s += (long) gridDim.x * blockDim.x) {
        const int a = ids[s];
        x[3 * a] = mine[3 * s]; x[3 * a + 1] = mine[3 * s + 1]; x[3 * a + 2] = mine[3 * s + 2];
    }
}

// gradients: gs_owner[s] += gs_mine[s] over this rank's halo ranges inside rank r's slab (atomics resolved in the owner's L2)
static __global__ void k_peer_push(PeerPtrs gs, const long *__restrict__ table, int rank, int nranks, const double *__restrict__ mine)
{
    const int r = blockIdx.y >> 1, h = blockIdx.y & 1;
    if (r == rank) return;
    const long *t = table + ((long) r * 2 + h) * 2;
    const long lo = 3 * t[0], hi = 3 * t[1];
    double *dst = gs.p[r];
    for (long k = lo + (long) blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += (long) gridDim.x * blockDim.x) {
        const double v = mine[k];
        if (v != 0.0) atomicAdd(&dst[k], v);
    }
}

// signal area of one rank (doubles; written by its peers, read by itself)
constexpr int kSigFlagA = 0;                                   // [kMaxPeers] step of "accumulator zeroed, positions published, displacement written"
constexpr int kSigDisp = State::kMaxPeers;                     // [kMaxPeers] the peers' displacement maxima (1e300: rebuild requested)
constexpr int kSigFlagB = 2 * State::kMaxPeers;                // [kMaxPeers] step of "gradients pushed, scalars written"
constexpr int kSigScal = 3 * State::kMaxPeers;                 // [kMaxPeers][16] the peers' 15 scalars
constexpr int kSigFlagC = 19 * State::kMaxPeers;               // [kMaxPeers] step of "host rows of my chunk uploaded" (nbb200_chunk_*)
constexpr int kSigFlagD = 20 * State::kMaxPeers;               // [kMaxPeers] step of "gradients of my slab written into the chunk buffers"
constexpr int kSigDoubles = 21 * State::kMaxPeers;

int nbb200_peer_export(NBB200State *state, char *handles192)
{
    if (state == nullptr || handles192 == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (!s.symGs.ensure(3 * (size_t) s.n) || !s.symXs.ensure((size_t) xs_ids_offset(s.n) + ((size_t) s.n + 1) / 2 + 1) || !s.symSig.ensure(kSigDoubles) || !s.sigStage.ensure(64)) return 0;
    cudaMemset(s.symSig.p, 0, sizeof(double) * kSigDoubles);
    cudaMemset(s.sigStage.p, 0, sizeof(double) * 64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hg, hx, hs;
    if (!cuda_ok(cudaIpcGetMemHandle(&hg, s.symGs.p), "cudaIpcGetMemHandle") || !cuda_ok(cudaIpcGetMemHandle(&hx, s.symXs.p), "cudaIpcGetMemHandle") ||
        !cuda_ok(cudaIpcGetMemHandle(&hs, s.symSig.p), "cudaIpcGetMemHandle")) return 0;
    std::memcpy(handles192, &hg, 64); std::memcpy(handles192 + 64, &hx, 64); std::memcpy(handles192 + 128, &hs, 64);
    s.gsExternal = s.symGs.p;
    return 1;
}

int nbb200_peer_import(NBB200State *state, int rank, const char *handles192)
{
    if (state == nullptr || handles192 == nullptr || rank < 0 || rank >= State::kMaxPeers) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (rank == s.rank) { s.peerGs[rank] = s.symGs.p; s.peerXs[rank] = s.symXs.p; s.peerSig[rank] = s.symSig.p; return 1; }
    cudaIpcMemHandle_t hg, hx, hs;
    std::memcpy(&hg, handles192, 64); std::memcpy(&hx, handles192 + 64, 64); std::memcpy(&hs, handles192 + 128, 64);
    void *pg = nullptr, *px = nullptr, *ps = nullptr;
    if (!cuda_ok(cudaIpcOpenMemHandle(&pg, hg, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle") ||
        !cuda_ok(cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle") ||
        !cuda_ok(cudaIpcOpenMemHandle(&ps, hs, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return 0;
    s.peerGs[rank] = (double *) pg; s.peerXs[rank] = (double *) px; s.peerSig[rank] = (double *) ps; s.peerOpened[rank] = true;
    s.peersReady = true;
    return 1;
}

/* same-process variant of nbb200_peer_import (an IPC handle cannot be opened by the process that exported it): rank `rank` is the
 * state `other` on the same device.  For single-GPU tests of the peer kernels and for callers that run several partitions per process. */
int nbb200_peer_attach_local(NBB200State *state, int rank, NBB200State *other)
{
    if (state == nullptr || other == nullptr || rank < 0 || rank >= State::kMaxPeers) return 0;
    State &s = *reinterpret_cast<State *>(state);
    State &o = *reinterpret_cast<State *>(other);
    if (o.symGs.p == nullptr || o.symXs.p == nullptr || o.symSig.p == nullptr) return 0;      // nbb200_peer_export allocates them
    s.peerGs[rank] = o.symGs.p; s.peerXs[rank] = o.symXs.p; s.peerSig[rank] = o.symSig.p;
    s.peersReady = true;
    return 1;
}

// ---- signalling through peer memory: a rank writes (value, step flag) into every peer's signal area; a waiting kernel spins
// (bounded) on its own area.  All ranks issue the same sequence of calls, `step` counts them.
static __global__ void k_signal_a(PeerPtrs sig, int rank, int nranks, double step, const double *__restrict__ localDisp)
{
    const int r = threadIdx.x;
    if (r >= nranks) return;
    volatile double *dst = sig.p[r];
    dst[kSigDisp + rank] = *localDisp;
    __threadfence_system();
    dst[kSigFlagA + rank] = step;
}

static __global__ void k_signal_b(PeerPtrs sig, int rank, int nranks, double step, const double *__restrict__ scal15)
{
    const int r = threadIdx.x;
    if (r >= nranks) return;
    volatile double *dst = sig.p[r];
    for (int k = 0; k < 15; k++) dst[kSigScal + 16 * rank + k] = scal15[k];
    __threadfence_system();
    dst[kSigFlagB + rank] = step;
}

// wait for all ranks' flags, then reduce.  The spin is bounded (so that a rank that died cannot hang the GPU for ever) by `maxSpins`
// sleeps of ~200 ns: NBB200_PEER_TIMEOUT_S, default 120 s -- long enough for a peer that is checkpointing, logging, collecting garbage
// or sitting in a debugger between two calls; all ranks must reach the same call within that time.  The time-out flag is cleared by
// every wait that completes, and a wait that timed out hands out NaN instead of sums of incomplete data.
static __global__ void k_wait(double *sigOwn, int flagBase, int nranks, double step, int mode /* 0: max of disp, 1: sum of scalars */, double *out, double *timeout, long maxSpins)
{
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    const int r = threadIdx.x;
    if (r < nranks) {
        volatile double *f = sigOwn + flagBase + r;
        long spins = 0;
        while (*f < step) { __nanosleep(200); if (++spins > maxSpins) { ok = 0; break; } }
    }
    __threadfence_system();
    __syncthreads();
    volatile double *v = sigOwn;
    const double bad = __longlong_as_double(0x7ff8000000000000LL);
    if (mode == 0) {
        if (threadIdx.x == 0) { double m = 0.0; for (int k = 0; k < nranks; k++) m = fmax(m, v[kSigDisp + k]); out[0] = ok ? m : bad; }
    } else if (mode == 1 && threadIdx.x < 15) {
        double t = 0.0;
        for (int k = 0; k < nranks; k++) t += v[kSigScal + 16 * k + threadIdx.x];
        out[threadIdx.x] = ok ? t : bad;
    }
    if (threadIdx.x == 0) *timeout = ok ? 0.0 : 1.0;
}

static long peer_max_spins()
{
    static const long spins = []() {
        const char *e = std::getenv("NBB200_PEER_TIMEOUT_S");
        const double seconds = (e != nullptr && std::atof(e) > 0.0) ? std::atof(e) : 120.0;
        return (long) (seconds / 250.0e-9);                  // a sleep of 200 ns plus the poll itself
    }();
    return spins;
}

/* after nbb200_peer_begin: tell every rank that this rank's accumulator is zeroed and its positions are published, together with its
 * local displacement maximum (forceRebuild != 0: 1e300).  Stream ordered, no host wait. */
void nbb200_peer_signal_begin(NBB200State *state, long step, const double *d_x, int forceRebuild)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (forceRebuild || s.isNew || d_x == nullptr) {
        const double big = 1.0e300;
        cudaMemcpyAsync(s.sigStage.p, &big, sizeof(double), cudaMemcpyHostToDevice, s.stream);      // pageable 8 bytes: staged by the driver
    } else displacement_enqueue(s, d_x, s.sigStage.p);
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerSig[r];
    k_signal_a<<<1, 32, 0, s.stream>>>(P, s.rank, s.nranks, (double) step, s.sigStage.p);
    s.launches += 1;
}

/* wait until every rank has signalled `step`; returns the maximum over the ranks of the displacement maxima (one host
 * synchronisation: the caller decides about the rebuild).  status: logic error on a time-out. */
double nbb200_peer_wait_begin(NBB200State *state, long step, int needValue, int *status)
{
    if (state == nullptr) return 0.0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    k_wait<<<1, 32, 0, s.stream>>>(s.symSig.p, kSigFlagA, s.nranks, (double) step, 0, s.sigStage.p + 17, s.sigStage.p + 33, peer_max_spins());
    s.launches += 1;
    if (!needValue) return 1.0e300;                            // the caller knows the decision (forced rebuild): ordering only, no host wait
    double out[17] = {0};
    if (!cuda_ok(cudaMemcpyAsync(s.hsmall + (kSmallDoubles - 192), s.sigStage.p + 17, sizeof(double) * 17, cudaMemcpyDeviceToHost, s.stream), "D2H") ||
        !cuda_ok(cudaStreamSynchronize(s.stream), "sync")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0.0; }
    std::memcpy(out, s.hsmall + (kSmallDoubles - 192), sizeof(out));
    if (out[16] != 0.0) { set_error("time-out waiting for the other ranks (begin)"); set_status(status, NBB200_STATUS_LOGIC_ERROR); }
    return out[0];
}

/* after nbb200_peer_push_gradients: hand the 15 scalars (6 energies, dE/dM) to every rank and tell them that this rank's pushes are done */
void nbb200_peer_signal_end(NBB200State *state, long step, const double *scal15)
{
    if (state == nullptr || scal15 == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    double *stage = s.hsmall + (kSmallDoubles - 64);           // pinned; the accumulators of the energy call use the front of hsmall
    std::memcpy(stage, scal15, sizeof(double) * 15);
    cudaMemcpyAsync(s.sigStage.p + 1, stage, sizeof(double) * 15, cudaMemcpyHostToDevice, s.stream);
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerSig[r];
    k_signal_b<<<1, 32, 0, s.stream>>>(P, s.rank, s.nranks, (double) step, s.sigStage.p + 1);
    s.launches += 1;
}

// the 15 scalars of a call (6 energies, dE/dM) from the accumulators ON THE DEVICE: dE/dM is linear in the per-image sums (W[9], G[3]) --
// dEdM = sum_k M_k [W_k; G_k] with M_k (9 x 12) made on the host from image_derivatives (symmetry_host.cpp) applied to unit vectors
static __global__ void k_scalars(const double *__restrict__ acc, int nsets, const double *__restrict__ mat, double *__restrict__ out15)
{
    const int t = threadIdx.x;
    if (t < 9) {
        double v = 0.0;
        for (int k = 1; k < nsets; k++) {
            const double *M = mat + (size_t) (k - 1) * 108 + 12 * t, *a = acc + 16 * k;
            for (int j = 0; j < 9; j++) v += M[j] * a[5 + j];
            for (int j = 0; j < 3; j++) v += M[9 + j] * a[2 + j];
        }
        out15[6 + t] = v;
    } else if (t < 15) {
        const int e = t - 9;
        double v = 0.0;
        if (e == NBB200_EMMEL) v = acc[0];
        else if (e == NBB200_EMMLJ) v = acc[1];
        else if (e == NBB200_EIMMMEL) { for (int k = 1; k < nsets; k++) v += acc[16 * k]; }
        else if (e == NBB200_EIMMMLJ) { for (int k = 1; k < nsets; k++) v += acc[16 * k + 1]; }
        else if (e == NBB200_EMMEL14) v = acc[16 * nsets];
        else if (e == NBB200_EMMLJ14) v = acc[16 * nsets + 1];
        out15[e] = v;
    }
}

/* the same hand-over as nbb200_peer_signal_end, with the scalars taken from the accumulators of the energy call that has just been enqueued
 * (NBModelABFS_B200_MMMMEnergySortedEnqueue) by a kernel: no host wait between the energy kernels and the signal.  The per-rank energies are
 * not handed to the host; nbb200_peer_read_sums gives the sums over the ranks. */
void nbb200_peer_signal_end_device(NBB200State *state, long step, int *status)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const int nimg = s.nsets - 1;
    if (!(s.scalMatValid && s.scalMatGeneration == s.numberOfUpdates && std::memcmp(s.scalMatLattice.v, s.lattice.M.v, sizeof(double) * 9) == 0)) {
        std::vector<double> mat((size_t) 108 * std::max(1, nimg), 0.0);
        for (int k = 0; k < nimg; k++) {
            const CandidateImage &im = s.plan.images[k];
            const double xt[3] = {s.trans.trans[3 * im.t] + (double) im.a, s.trans.trans[3 * im.t + 1] + (double) im.b, s.trans.trans[3 * im.t + 2] + (double) im.c};
            for (int j = 0; j < 12; j++) {
                double W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, G[3] = {0, 0, 0}, out[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                if (j < 9) W[j] = 1.0; else G[j - 9] = 1.0;
                image_derivatives(out, s.lattice, s.trans.rot[im.t], xt, W, G);
                for (int t = 0; t < 9; t++) mat[(size_t) k * 108 + 12 * t + j] = out[t];
            }
        }
        if (!s.scalMat.ensure(mat.size())) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return; }
        // (a pageable copy: staged by the driver before the call returns, ordered on the stream)
        if (!cuda_ok(cudaMemcpyAsync(s.scalMat.p, mat.data(), sizeof(double) * mat.size(), cudaMemcpyHostToDevice, s.stream), "H2D")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
        cudaStreamSynchronize(s.stream);                       // rare (list update or lattice change): keeps `mat` alive until the DMA is done
        s.scalMatValid = true; s.scalMatGeneration = s.numberOfUpdates; s.scalMatLattice = s.lattice.M;
    }
    k_scalars<<<1, 32, 0, s.stream>>>(s.accum.p, s.nsets, s.scalMat.p, s.sigStage.p + 1);
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerSig[r];
    k_signal_b<<<1, 32, 0, s.stream>>>(P, s.rank, s.nranks, (double) step, s.sigStage.p + 1);
    s.launches += 2;
}

/* wait (on the stream) until every rank has signalled the end of `step`: all pushes into this rank's accumulator are complete and the
 * scalars are summed; no host wait -- nbb200_peer_read_sums fetches them */
void nbb200_peer_wait_end(NBB200State *state, long step)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    k_wait<<<1, 32, 0, s.stream>>>(s.symSig.p, kSigFlagB, s.nranks, (double) step, 1, s.sigStage.p + 34, s.sigStage.p + 33, peer_max_spins());
    s.launches += 1;
    cudaMemcpyAsync(s.hsmall + (kSmallDoubles - 128), s.sigStage.p + 33, sizeof(double) * 17, cudaMemcpyDeviceToHost, s.stream);   // [timeout, 15 sums]
}

void nbb200_peer_read_sums(NBB200State *state, double *sum15, int *status)
{
    if (state == nullptr || sum15 == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (!cuda_ok(cudaStreamSynchronize(s.stream), "sync")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return; }
    std::memcpy(sum15, s.hsmall + (kSmallDoubles - 128) + 1, sizeof(double) * 15);
    if (s.hsmall[kSmallDoubles - 128] != 0.0) { set_error("time-out waiting for the other ranks"); set_status(status, NBB200_STATUS_LOGIC_ERROR); }
}

/* start of a call: zero the own gradient accumulator and publish the positions of the own slab (sorted order), all on the stream */
void nbb200_peer_begin(NBB200State *state, const double *d_x, long s0, long count)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    cudaMemsetAsync(s.symGs.p, 0, sizeof(double) * 3 * (size_t) s.n, s.stream);
    s.gsZeroed = true;
    if (count > 0 && d_x != nullptr) {
        k_publish_slab<<<(unsigned int) ((count + 255) / 256), 256, 0, s.stream>>>(d_x, s.sAtom.p, s0, count, s.symXs.p, reinterpret_cast<int *>(s.symXs.p + xs_ids_offset(s.n)));
        s.launches += 1;
    }
}

/* after a collective that orders it behind every rank's nbb200_peer_begin: fetch positions from their owners into d_x (atom order) */
void nbb200_peer_pull_positions(NBB200State *state, const long *d_table, const long *slabEdges /* host, nranks + 1 */, int wholeSlabs, double *d_x)
{
    if (state == nullptr || d_x == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    PeerPtrs P; SlabEdges E;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerXs[r];
    for (int r = 0; r <= State::kMaxPeers; r++) E.s[r] = (slabEdges != nullptr && r <= s.nranks) ? slabEdges[r] : 0;
    const dim3 grid(wholeSlabs ? 148 : 32, 2 * s.nranks);
    k_peer_copy<<<grid, 256, 0, s.stream>>>(P, E, d_table, s.rank, s.nranks, wholeSlabs, (long) s.n, s.symXs.p);
    k_peer_scatter<<<grid, 256, 0, s.stream>>>(s.symXs.p, E, d_table, s.rank, s.nranks, wholeSlabs, (long) s.n, s.sAtom.p, d_x);
    s.launches += 2;
}

/* after the energy call: add this rank's halo contributions into their owners' accumulators */
void nbb200_peer_push_gradients(NBB200State *state, const long *d_table)
{
    if (state == nullptr || d_table == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerGs[r];
    k_peer_push<<<dim3(32, 2 * s.nranks), 256, 0, s.stream>>>(P, d_table, s.rank, s.nranks, s.symGs.p);
    s.launches += 1;
}

/* ---- host callers on several ranks: atom-order CHUNKS between host and device, redistribution over peer memory ----
 * A caller that keeps coordinates and gradients in host arrays (DistributedNB.call_host) should not gather / scatter rows on the host: rank r
 * moves the CONTIGUOUS rows [n r / R, n (r + 1) / R) of the host arrays (one DMA each way), and the device sorts out who needs what -- a rank
 * gathers the positions of the atoms it owns from the chunk buffers of the ranks that uploaded them, and writes the gradients of its atoms into
 * the chunk buffers of the ranks that download them, with plain loads / stores over NVLink; two more flag rounds order the phases. */
static __device__ __forceinline__ int chunk_of(long a, const SlabEdges &E, int nranks, long n)
{
    int r = (int) ((a * nranks) / n);
    while (r > 0 && a < E.s[r]) r--;
    while (r < nranks - 1 && a >= E.s[r + 1]) r++;
    return r;
}

static __global__ void k_chunk_gather(PeerPtrs xc, SlabEdges E, int nranks, long n, const int *__restrict__ sAtom, long s0, long count, double *__restrict__ x)
{
    const long k = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int a = sAtom[s0 + k];
    if (a < 0) return;
    const double *src = xc.p[chunk_of(a, E, nranks, n)] + 3 * (long) a;
    x[3 * (long) a] = src[0]; x[3 * (long) a + 1] = src[1]; x[3 * (long) a + 2] = src[2];
}

static __global__ void k_chunk_scatter(PeerPtrs gc, SlabEdges E, int nranks, long n, const int *__restrict__ sAtom, long s0, long count, const double *__restrict__ gs)
{
    const long k = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int a = sAtom[s0 + k];
    if (a < 0) return;
    double *dst = gc.p[chunk_of(a, E, nranks, n)] + 3 * (long) a;
    const double *g = gs + 3 * (s0 + k);
    dst[0] = g[0]; dst[1] = g[1]; dst[2] = g[2];
}

static __global__ void k_signal_flag(PeerPtrs sig, int rank, int nranks, double step, int flagBase)
{
    const int r = threadIdx.x;
    if (r >= nranks) return;
    volatile double *dst = sig.p[r];
    __threadfence_system();
    dst[flagBase + rank] = step;
}

static SlabEdges chunk_edges(const State &s)
{
    SlabEdges E;
    for (int r = 0; r <= State::kMaxPeers; r++) E.s[r] = (r <= s.nranks) ? ((long) s.n * r) / s.nranks : (long) s.n;
    return E;
}

int nbb200_peer_export_chunks(NBB200State *state, char *handles128)
{
    if (state == nullptr || handles128 == nullptr) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (!s.symXc.ensure(3 * (size_t) s.n) || !s.symGc.ensure(3 * (size_t) s.n)) return 0;
    cudaIpcMemHandle_t hx, hg;
    if (!cuda_ok(cudaIpcGetMemHandle(&hx, s.symXc.p), "cudaIpcGetMemHandle") || !cuda_ok(cudaIpcGetMemHandle(&hg, s.symGc.p), "cudaIpcGetMemHandle")) return 0;
    std::memcpy(handles128, &hx, 64); std::memcpy(handles128 + 64, &hg, 64);
    return 1;
}

int nbb200_peer_import_chunks(NBB200State *state, int rank, const char *handles128)
{
    if (state == nullptr || handles128 == nullptr || rank < 0 || rank >= State::kMaxPeers) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (rank == s.rank) { s.peerXc[rank] = s.symXc.p; s.peerGc[rank] = s.symGc.p; return s.symXc.p != nullptr ? 1 : 0; }
    cudaIpcMemHandle_t hx, hg;
    std::memcpy(&hx, handles128, 64); std::memcpy(&hg, handles128 + 64, 64);
    void *px = nullptr, *pg = nullptr;
    if (!cuda_ok(cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle") ||
        !cuda_ok(cudaIpcOpenMemHandle(&pg, hg, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return 0;
    s.peerXc[rank] = (double *) px; s.peerGc[rank] = (double *) pg; s.peerChunkOpened[rank] = true;
    return 1;
}

/* same-process variant (see nbb200_peer_attach_local) */
int nbb200_peer_attach_local_chunks(NBB200State *state, int rank, NBB200State *other)
{
    if (state == nullptr || other == nullptr || rank < 0 || rank >= State::kMaxPeers) return 0;
    State &s = *reinterpret_cast<State *>(state);
    State &o = *reinterpret_cast<State *>(other);
    if (o.symXc.p == nullptr || o.symGc.p == nullptr) return 0;
    s.peerXc[rank] = o.symXc.p; s.peerGc[rank] = o.symGc.p;
    return 1;
}

/* host rows [a0, a0 + count) of h_x (the base of the caller's [n][3] array) -> this rank's chunk buffer, same offsets; page-locked arrays go
 * by one DMA, others through the state's page-locked staging array (threaded copy).  Stream ordered. */
void nbb200_chunk_upload(NBB200State *state, const double *h_x, long a0, long count)
{
    if (state == nullptr || h_x == nullptr || count <= 0) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.symXc.p == nullptr || a0 < 0 || a0 + count > s.n) { set_error("nbb200_chunk_upload: chunk buffers are not exported or the range is invalid"); return; }
    const double *src = h_x + 3 * a0;
    // (the query is slow for pageable memory: asked once per array)
    if (s.hostXChecked != (const void *) h_x) { s.hostXChecked = h_x; s.hostXPinned = is_pinned_host(src); }
    if (s.hostXPinned) {
        cuda_ok(cudaMemcpyAsync(s.symXc.p + 3 * a0, src, sizeof(double) * 3 * (size_t) count, cudaMemcpyHostToDevice, s.stream), "H2D chunk");
        return;
    }
    cudaStreamSynchronize(s.stream);                           // the staging array may still feed an earlier copy
    staged_upload(s, s.symXc.p + 3 * a0, src, s.hx + 3 * a0, 3 * count);
}

/* which = 0: "my chunk is uploaded"; 1: "the gradients of my slab are written".  Stream ordered, no host wait. */
void nbb200_chunk_signal(NBB200State *state, long step, int which)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerSig[r];
    k_signal_flag<<<1, 32, 0, s.stream>>>(P, s.rank, s.nranks, (double) step, which == 0 ? kSigFlagC : kSigFlagD);
    s.launches += 1;
}

void nbb200_chunk_wait(NBB200State *state, long step, int which)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    k_wait<<<1, 32, 0, s.stream>>>(s.symSig.p, which == 0 ? kSigFlagC : kSigFlagD, s.nranks, (double) step, 2, s.sigStage.p + 50, s.sigStage.p + 51, peer_max_spins());
    s.launches += 1;
}

/* after nbb200_chunk_wait(step, 0): the positions of the atoms this rank owns (its slab of the current lists), from the chunk buffers of the
 * ranks that uploaded them, into d_x (atom order) */
void nbb200_chunk_gather_owned(NBB200State *state, double *d_x)
{
    if (state == nullptr || d_x == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const long count = (long) std::max(0, s.ownHi - s.ownLo);
    if (count == 0) return;
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerXc[r];
    k_chunk_gather<<<(unsigned int) ((count + 255) / 256), 256, 0, s.stream>>>(P, chunk_edges(s), s.nranks, (long) s.n, s.sAtom.p, (long) s.ownLo, count, d_x);
    s.launches += 1;
}

/* after nbb200_peer_wait_end: the (complete) gradients of this rank's slab go to the chunk buffers of the ranks that download those atoms */
void nbb200_chunk_scatter_gradients(NBB200State *state)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const long count = (long) std::max(0, s.ownHi - s.ownLo);
    if (count == 0 || s.gs == nullptr) return;
    PeerPtrs P;
    for (int r = 0; r < State::kMaxPeers; r++) P.p[r] = s.peerGc[r];
    k_chunk_scatter<<<(unsigned int) ((count + 255) / 256), 256, 0, s.stream>>>(P, chunk_edges(s), s.nranks, (long) s.n, s.sAtom.p, (long) s.ownLo, count, s.gs);
    s.launches += 1;
}

/* after nbb200_chunk_wait(step, 1): rows [a0, a0 + count) of this rank's gradient chunk buffer are ADDED to the same rows of h_g (the base of
 * the caller's [n][3] array).  Synchronises the stream. */
int nbb200_chunk_download(NBB200State *state, double *h_g, long a0, long count, int overwrite)
{
    if (state == nullptr || h_g == nullptr || count < 0) return 0;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    if (s.symGc.p == nullptr || a0 < 0 || a0 + count > s.n) { set_error("nbb200_chunk_download: chunk buffers are not exported or the range is invalid"); return 0; }
    double *flag = s.hsmall + (kSmallDoubles - 200);           // the time-out flag of the last nbb200_chunk_wait
    bool ok = cuda_ok(cudaMemcpyAsync(flag, s.sigStage.p + 51, sizeof(double), cudaMemcpyDeviceToHost, s.stream), "D2H");
    const long m = 3 * count;
    if (s.hostGChecked != (const void *) h_g) { s.hostGChecked = h_g; s.hostGPinned = is_pinned_host(h_g + 3 * a0); }
    if (overwrite && s.hostGPinned) {
        // the caller's rows are set, and its array is page-locked: one DMA straight into it
        if (m > 0) ok = ok && cuda_ok(cudaMemcpyAsync(h_g + 3 * a0, s.symGc.p + 3 * a0, sizeof(double) * (size_t) m, cudaMemcpyDeviceToHost, s.stream), "D2H chunk");
        ok = cuda_ok(cudaStreamSynchronize(s.stream), "sync") && ok;
        if (ok && *flag != 0.0) { set_error("time-out waiting for the other ranks (chunk exchange)"); ok = false; }
        return ok ? 1 : 0;
    }
    // a few pieces: the host adds (or copies) piece k while the DMA of piece k + 1 runs
    const long pieces = std::max(1L, std::min(4L, m / (1L << 18)));
    if (s.chunkEvents[0] == nullptr) for (auto &e : s.chunkEvents) ok = ok && cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    for (long p = 0; p < pieces && ok && m > 0; p++) {
        const long k0 = (m * p) / pieces, k1 = (m * (p + 1)) / pieces;
        ok = cuda_ok(cudaMemcpyAsync(s.hgrad + 3 * a0 + k0, s.symGc.p + 3 * a0 + k0, sizeof(double) * (size_t) (k1 - k0), cudaMemcpyDeviceToHost, s.stream), "D2H chunk") &&
             cuda_ok(cudaEventRecord(s.chunkEvents[p], s.stream), "event");
    }
    for (long p = 0; p < pieces && ok && m > 0; p++) {
        const long k0 = (m * p) / pieces, k1 = (m * (p + 1)) / pieces;
        ok = cuda_ok(cudaEventSynchronize(s.chunkEvents[p]), "event wait");
        if (ok && p == 0 && *flag != 0.0) { set_error("time-out waiting for the other ranks (chunk exchange)"); ok = false; }
        if (ok && overwrite) nbb200_host_copy(h_g + 3 * a0 + k0, s.hgrad + 3 * a0 + k0, k1 - k0);
        else if (ok) nbb200_host_add(h_g + 3 * a0 + k0, s.hgrad + 3 * a0 + k0, k1 - k0);
    }
    ok = cuda_ok(cudaStreamSynchronize(s.stream), "sync") && ok;
    if (ok && m == 0 && *flag != 0.0) { set_error("time-out waiting for the other ranks (chunk exchange)"); ok = false; }
    return ok ? 1 : 0;
}

int nbb200_chunk_download_add(NBB200State *state, double *h_g, long a0, long count) { return nbb200_chunk_download(state, h_g, a0, count, 0); }

/* ---- velocity Verlet on the device (pCore-1.9.0/pCore/VelocityVerletIntegrator.py:60-81 in Cartesian variables) ---- */
// first half: x += dt v + dt^2/2 a ; v += dt/2 a
static __global__ void k_vv_first(double *__restrict__ x, double *__restrict__ v, const double *__restrict__ a, double dt, long m)
{
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double ai = a[i], vi = v[i];
    x[i] += dt * vi + 0.5 * dt * dt * ai;
    v[i] = vi + 0.5 * dt * ai;
}

// first half fused with the displacement check of CheckForUpdate (nbb200_md_run): one thread per atom, |x - xref|^2 into the running maximum
static __global__ void k_vv_first_disp(double *__restrict__ x, double *__restrict__ v, const double *__restrict__ a, double dt, int n, const double *__restrict__ xref,
                                       const unsigned char *__restrict__ fixed, unsigned long long *__restrict__ out, unsigned long long *__restrict__ zeroOther)
{
    if (zeroOther != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zeroOther = 0ULL;
    const int atom = blockIdx.x * blockDim.x + threadIdx.x;
    double r2 = 0.0;
    if (atom < n) {
        double d[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const long i = 3 * (long) atom + c;
            const double ai = a[i], vi = v[i];
            double xi = x[i];
            xi += dt * vi + 0.5 * dt * dt * ai;
            x[i] = xi;
            v[i] = vi + 0.5 * dt * ai;
            d[c] = xi - xref[i];
        }
        if (fixed == nullptr || !fixed[atom]) r2 = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
    }
    for (int off = 16; off > 0; off >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, off));
    if ((threadIdx.x & 31) == 0 && r2 > 0.0) atomicMax(out, (unsigned long long) __double_as_longlong(r2));
}

// second half: a = -100 g / m (kJ mol^-1 A^-1 amu^-1 -> A ps^-2) ; v += dt/2 a ; kinetic energy 0.5 * 0.01 * sum m v^2 (kJ/mol)
// pub (nbb200_md_run): the bonded energies of the step (complete before this kernel starts) are stored into page-locked host memory -- no copy
// operation in the stream
static __global__ void k_vv_second(double *__restrict__ v, double *__restrict__ a, const double *__restrict__ g, const double *__restrict__ mass, double dt, long m,
                                   double *__restrict__ ke, double *__restrict__ zeroOther = nullptr,
                                   const double *pubSrc = nullptr, double *pubDst = nullptr, int pubCount = 0)
{
    if (zeroOther != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zeroOther = 0.0;      // two-slot use (nbb200_md_run): prepares the next step's slot
    if (blockIdx.x == 0 && threadIdx.x < pubCount) pubDst[threadIdx.x] = pubSrc[threadIdx.x];
    double local = 0.0;
    for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long) gridDim.x * blockDim.x) {
        const double mi = mass[i / 3], ai = -100.0 * g[i] / mi;
        const double vi = v[i] + 0.5 * dt * ai;
        a[i] = ai; v[i] = vi;
        local += mi * vi * vi;
    }
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(ke, 0.5 * 0.01 * local);
}

void nbb200_vv_first_half(NBB200State *state, double *d_x, double *d_v, const double *d_a, double dt)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const long m = 3 * (long) s.n;
    k_vv_first<<<(unsigned int) ((m + 255) / 256), 256, 0, s.stream>>>(d_x, d_v, d_a, dt, m);
    s.launches += 1;
}

void nbb200_vv_second_half(NBB200State *state, double *d_v, double *d_a, const double *d_g, const double *d_mass, double dt, double *d_ke)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const long m = 3 * (long) s.n;
    cudaMemsetAsync(d_ke, 0, sizeof(double), s.stream);
    k_vv_second<<<(unsigned int) std::min<long>(148 * 8, (m + 255) / 256), 256, 0, s.stream>>>(d_v, d_a, d_g, d_mass, dt, m, d_ke);
    s.launches += 1;
}

/* ---- FP32 roofline denominator, measured: register-only FMA chains (no memory traffic), every lane of every SM busy ---- */
static __global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float b, float c)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float) (threadIdx.x + i) * 1.0e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

double nbb200_measure_fp32_peak(int device, int *status)
{
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0.0; }
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0.0; }
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    float *out = nullptr;
    if (!cuda_ok(cudaMalloc((void **) &out, sizeof(float) * (size_t) blocks * threads), "cudaMalloc")) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return 0.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_fma_peak<<<blocks, threads>>>(out, 2000, 0.999999f, 1.0e-7f);                    // warm-up: clocks ramp
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_fma_peak<<<blocks, threads>>>(out, iters, 0.999999f, 1.0e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16.0 * (double) iters * (double) blocks * threads;
        if (ms > 0.f) best = std::max(best, flops / (ms * 1.0e-3) / 1.0e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (!cuda_ok(cudaGetLastError(), "k_fma_peak")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0.0; }
    return best;
}

/* ---- the MD loop itself, native: what pdynamo-mirror_b200/md.py:Run does per step, without the interpreter between the calls
 * and without a host wait on the critical path.
 *
 * The list-update decision of a step needs the new coordinates, so a loop that waits for it leaves the GPU idle once per step (wake-up +
 * launch latency, ~30 us of a 110-200 us step).  Here a step is enqueued OPTIMISTICALLY on the current lists -- displacement check, energy
 * kernels, bonded terms, second half -- and the host reads the check's result while the GPU is busy with that step.  In 13 of 14 steps (the
 * reference's DHFR run) no update was due and the host is already enqueueing the next step; otherwise the step is taken back (v -= f a with
 * the accelerations it produced; x is not touched by a second half), the lists are rebuilt through the ordinary path and the step is
 * evaluated again.  Results of step k (energies, bonded energies, kinetic energy) land in page-locked memory in stream order, in two
 * alternating slots, and are read after the decision of step k + 1. ---- */
static __global__ void k_axpy(double *__restrict__ y, const double *__restrict__ x, double c, long m)
{
    const long i = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) y[i] += c * x[i];
}

int nbb200_md_run(NBB200State *state, NBB200MMTerms *terms, int nsteps, int updateFrequency, double *d_x, double *d_v, double *d_a, double *d_g,
                  const double *d_mass, const double *box6, double timeStep, const double *langevinFactors7, double secondHalfDt,
                  unsigned long long seed, unsigned long long firstIteration, double *d_ke, double *potential, double *kinetic,
                  double *nbEnergies6, double *bondedEnergies5, int *status)
{
    if (state == nullptr || d_x == nullptr || d_v == nullptr || d_a == nullptr || d_g == nullptr || d_mass == nullptr || d_ke == nullptr || nsteps < 0) {
        set_status(status, NBB200_STATUS_INVALID_ARGUMENT); return 0;
    }
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    flush_pending(s);
    if (s.hacc == nullptr && !cuda_ok(cudaMallocHost((void **) &s.hacc, sizeof(double) * kSmallDoubles), "cudaMallocHost")) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return 0; }
    if (!s.mdScalars.ensure(32)) { set_status(status, NBB200_STATUS_OUT_OF_MEMORY); return 0; }
    if (terms != nullptr) { MMTerms_B200_SetStream(terms, s.stream); mmterms_reset_slots(terms); }       // stream order is what hands the results over
    // page-locked result slots (two of each): accumulators of the energy call, kinetic energy, displacement maximum
    double *haccSlot[2] = {s.hacc, s.hacc + kSmallDoubles / 2};
    double *hke = s.hacc + (kSmallDoubles - 8), *hdisp = s.hacc + (kSmallDoubles - 16);
    // device scalars in two alternating slots each: the kernel that fills slot k & 1 clears the other one for the next step (no memsets)
    double *d_disp2 = s.mdScalars.p, *d_ke2 = s.mdScalars.p + 8;
    if (!cuda_ok(cudaMemsetAsync(s.mdScalars.p, 0, sizeof(double) * 32, s.stream), "memset")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
    if ((size_t) 16 * (s.nsets + 1) > kSmallDoubles / 2 - 16) { set_error("too many images for the result slots"); set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
    cudaEvent_t evDisp;
    if (!cuda_ok(cudaEventCreateWithFlags(&evDisp, cudaEventDisableTiming), "cudaEventCreate")) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
    const long m = 3 * (long) s.n;
    const bool savedOverwrite = s.gradOverwrite;
    s.gradOverwrite = true;                                            // the NB term sets d_g, the bonded terms accumulate
    // fused mode: the five memsets of a step are folded into neighbouring kernels (accumulators + cursor by k_pack_records, sorted gradient by
    // k_unsort_gradients, the two-slot scalars by the kernel that fills the other slot).  NBB200_MD_FUSED=0 keeps the separate memsets.
    static const bool fusedMode = []() { const char *e = std::getenv("NBB200_MD_FUSED"); return e == nullptr || std::atoi(e) != 0; }();      // measured: DHFR Langevin 5.08 -> 5.36 k steps/s, ionic NVE 9.9 -> 11.1 k steps/s
    s.mdFused = fusedMode;
    static const bool noSpeculation = std::getenv("NBB200_MD_NO_SPECULATION") != nullptr;
    int updates = 0, nspec = 0, dispKnown = 0;
    double dispLast = 0.0, dispBefore = 0.0;          // displacement maxima (A) of the last two optimistic steps since the last list update
    bool ok = true;
    double eStep[2][6], dEdM[9], e5[5] = {0, 0, 0, 0, 0};
    for (int c = 0; c < 6; c++) eStep[0][c] = eStep[1][c] = 0.0;

    // everything of step k after its first half, on the lists as they are: energy (deferred into slot k & 1), bonded terms, second half, kinetic energy
    // redo: the step is evaluated a second time after it was taken back -- its result slots hold the first attempt and are cleared explicitly
    // results reach the host by stores into page-locked memory from the kernels that complete them (no copy operations in the stream: each one
    // costs a few microseconds between two small kernels); NBB200_NO_PUBLISH=1 keeps the copies
    static const bool publish = std::getenv("NBB200_NO_PUBLISH") == nullptr;
    // (who stores what: the accumulators by the unsort pass of the step; the bonded energies by the second half; the displacement maximum of
    // step k and the kinetic energy of step k - 1 by the first kernel of step k's energy call -- always a kernel that STARTS after the value is
    // complete: letting the producing kernel's last CTA do it costs a __threadfence behind its bulk stores, ~4 us per kernel on this GPU)
    // the bonded terms run on a side stream NEXT TO the NB kernels of the step: they add into the NB state's sorted-order accumulator (rows
    // invPerm[atom]) and the unsort pass waits for them -- 12 us of a latency-bound kernel leave the critical path.  NBB200_MD_NO_SIDE_STREAM=1:
    // behind the unsort pass, into d_g, as the Python loop does
    static const bool sideOff = std::getenv("NBB200_MD_NO_SIDE_STREAM") != nullptr;
    const bool sideBonded = !sideOff && terms != nullptr && fusedMode && publish && s.nranks == 1 && s.gsExternal == nullptr;
    cudaStream_t sideStream = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    if (sideBonded && !(cuda_ok(cudaStreamCreateWithFlags(&sideStream, cudaStreamNonBlocking), "cudaStreamCreate") &&
                        cuda_ok(cudaEventCreateWithFlags(&evFork, cudaEventDisableTiming), "cudaEventCreate") &&
                        cuda_ok(cudaEventCreateWithFlags(&evJoin, cudaEventDisableTiming), "cudaEventCreate"))) { set_status(status, NBB200_STATUS_LOGIC_ERROR); return 0; }
    auto enqueue_step = [&](int k, bool redo) -> bool {
        for (int c = 0; c < 9; c++) dEdM[c] = 0.0;
        const bool side = sideBonded && s.hostCounters.itemCount > 0;
        if (side) {
            if (!s.gradSorted.ensure(3 * (size_t) s.n)) return false;
            if (!s.gsZeroed) {                               // nobody has cleared the accumulator for this call yet
                if (!cuda_ok(cudaMemsetAsync(s.gradSorted.p, 0, sizeof(double) * 3 * (size_t) s.n, s.stream), "memset")) return false;
                s.gsZeroed = true;
            }
            if (!cuda_ok(cudaEventRecord(evFork, s.stream), "event") || !cuda_ok(cudaStreamWaitEvent(sideStream, evFork, 0), "wait") ||
                !mmterms_enqueue_slot(terms, d_x, s.gradSorted.p, k & 1, fusedMode && !redo, publish, s.invPerm.p, sideStream) ||
                !cuda_ok(cudaEventRecord(evJoin, sideStream), "event")) return false;
            s.preUnsortEvent = evJoin;
        }
        const double *pubSrc = nullptr; double *pubDst = nullptr;
        if (terms != nullptr && publish) mmterms_slot_pointers(terms, k & 1, &pubSrc, &pubDst);
        // with the bonded terms already in the sorted accumulator (or none at all) nothing stands between the unsort pass and the second half:
        // one kernel does both
        const bool fuseSecond = fusedMode && publish && (side || terms == nullptr) && s.nranks == 1 && s.gsExternal == nullptr && s.hostCounters.itemCount > 0;
        if (fuseSecond) {
            if (redo && !cuda_ok(cudaMemsetAsync(d_ke2 + (k & 1), 0, sizeof(double), s.stream), "memset")) return false;
            s.secondHalf.v = d_v; s.secondHalf.a = d_a; s.secondHalf.mass = d_mass; s.secondHalf.dt = secondHalfDt; s.secondHalf.ke = d_ke2 + (k & 1);
            s.secondHalf.zeroOther = d_ke2 + ((k + 1) & 1); s.secondHalf.pubSrc = pubSrc; s.secondHalf.pubDst = pubDst; s.secondHalf.pubCount = pubSrc != nullptr ? 5 : 0;
        }
        s.secondHalfDone = false;
        const bool okE = energy_enqueue(s, d_g, false, haccSlot[k & 1]);
        s.preUnsortEvent = nullptr; s.secondHalf.v = nullptr;
        if (!okE) return false;
        if (s.secondHalfDone) return true;
        if (terms != nullptr && !side && !mmterms_enqueue_slot(terms, d_x, d_g, k & 1, fusedMode && !redo, publish)) return false;
        if ((redo || !fusedMode) && !fuseSecond && !cuda_ok(cudaMemsetAsync(d_ke2 + (k & 1), 0, sizeof(double), s.stream), "memset")) return false;
        k_vv_second<<<(unsigned int) std::min<long>(148 * 8, (m + 255) / 256), 256, 0, s.stream>>>(d_v, d_a, d_g, d_mass, secondHalfDt, m, d_ke2 + (k & 1),
                                                                                                  fusedMode ? d_ke2 + ((k + 1) & 1) : nullptr,
                                                                                                  pubSrc, pubDst, pubSrc != nullptr ? 5 : 0);
        s.launches += 1;
        if (publish) return cuda_ok(cudaGetLastError(), "k_vv_second");
        return cuda_ok(cudaMemcpyAsync(hke + (k & 1), d_ke2 + (k & 1), sizeof(double), cudaMemcpyDeviceToHost, s.stream), "D2H kinetic energy");
    };
    // the numbers of a completed step: accumulators -> energies (energy_finish; must run while the lists the step used are still current),
    // bonded slot, kinetic slot
    auto harvest = [&](int k, bool nbDone) {
        double *e6 = eStep[k & 1];
        if (!nbDone) energy_finish(s, e6, true, dEdM, haccSlot[k & 1], &s.lattice);
        double pot = e6[0] + e6[1] + e6[2] + e6[3] + e6[4] + e6[5];
        if (terms != nullptr) { mmterms_read_slot(terms, k & 1, e5); pot += e5[0] + e5[1] + e5[2] + e5[3] + e5[4]; }
        if (potential != nullptr) potential[k] = pot;
        if (kinetic != nullptr) kinetic[k] = hke[k & 1];
    };

    for (int k = 0; k < nsteps && ok; k++) {
        s.xcur = d_x;
        const bool forced = updateFrequency > 0 && (k + 1) % updateFrequency == 0;
        bool latticeSame = s.trans.n == 0;
        if (!latticeSame && s.haveRefLattice && box6 != nullptr) {
            Lattice now = s.lattice;
            now.set_crystal(box6);
            latticeSame = std::memcmp(now.M.v, s.refLattice.M.v, sizeof(double) * 9) == 0;
        }
        // when the displacement maximum of the last two steps extrapolates beyond the buffer, an update is probably due at this step: an
        // optimistic step would be thrown away (~a whole step of GPU time), waiting for the decision costs a short bubble.
        // NBB200_MD_NO_PREDICTION=1: always optimistic
        static const bool noPrediction = std::getenv("NBB200_MD_NO_PREDICTION") != nullptr;
        const double bufferDist = 0.5 * (s.list - s.stOuterCutoff);
        const bool updateLikely = !noPrediction && dispKnown >= 2 && dispLast + (dispLast - dispBefore) + 0.02 * bufferDist > bufferDist;
        const bool speculate = !noSpeculation && !forced && !updateLikely && !s.isNew && !s.useCentering && s.nranks == 1 && !s.timing && latticeSame &&
                               s.list == s.stListCutoff && s.outer == s.stOuterCutoff;
        const bool fusedFirst = speculate && fusedMode;         // first half and displacement check in one kernel
        double *d_disp = d_disp2 + (nspec & 1), *d_dispOther = d_disp2 + ((nspec + 1) & 1);
        if (fusedFirst) {
            if (langevinFactors7 != nullptr) ok = langevin_first_disp(s, d_x, d_v, d_a, d_mass, langevinFactors7, seed, firstIteration + (unsigned long long) k, d_disp, d_dispOther);
            else {
                k_vv_first_disp<<<(s.n + 127) / 128, 128, 0, s.stream>>>(d_x, d_v, d_a, timeStep, s.n, s.xref.p, s.nfixed > 0 ? s.fixedFlag.p : nullptr,
                                                                        reinterpret_cast<unsigned long long *>(d_disp), reinterpret_cast<unsigned long long *>(d_dispOther));
                s.launches += 1;
            }
        } else if (langevinFactors7 != nullptr) nbb200_langevin_first_half(state, d_x, d_v, d_a, d_mass, langevinFactors7, seed, firstIteration + (unsigned long long) k);
        else nbb200_vv_first_half(state, d_x, d_v, d_a, timeStep);
        bool needSync = !speculate;
        if (speculate) {
            // optimistic: the check of CheckForUpdate (NBModelABFS.c:691-746) and the whole step go out together
            nspec += 1;
            ok = ok && (fusedFirst || displacement_enqueue(s, d_x, d_disp, nullptr));
            if (ok && publish) {
                s.prePubSrc[0] = d_disp; s.prePubDst[0] = hdisp + (k & 1);
                s.prePubSrc[1] = k > 0 ? d_ke2 + ((k - 1) & 1) : nullptr; s.prePubDst[1] = hke + ((k - 1) & 1);
                s.prePubEvent = evDisp; s.prePubDone = false;
                ok = enqueue_step(k, false);
                s.prePubSrc[0] = s.prePubSrc[1] = nullptr; s.prePubEvent = nullptr;
                if (ok && !s.prePubDone)                    // no tile kernel in this call (no pairs): copy operations, event behind them
                    ok = cuda_ok(cudaMemcpyAsync(hdisp + (k & 1), d_disp, sizeof(double), cudaMemcpyDeviceToHost, s.stream), "D2H displacement") &&
                         (k == 0 || cuda_ok(cudaMemcpyAsync(hke + ((k - 1) & 1), d_ke2 + ((k - 1) & 1), sizeof(double), cudaMemcpyDeviceToHost, s.stream), "D2H kinetic energy")) &&
                         cuda_ok(cudaEventRecord(evDisp, s.stream), "event");
                ok = ok && cuda_ok(cudaEventSynchronize(evDisp), "event wait");
            } else
                ok = ok && cuda_ok(cudaMemcpyAsync(hdisp + (k & 1), d_disp, sizeof(double), cudaMemcpyDeviceToHost, s.stream), "D2H displacement") &&
                     cuda_ok(cudaEventRecord(evDisp, s.stream), "event") && enqueue_step(k, false) && cuda_ok(cudaEventSynchronize(evDisp), "event wait");
            if (!ok) { set_status(status, NBB200_STATUS_LOGIC_ERROR); break; }
            // the stream has passed the check of step k: step k - 1 is complete
            if (k > 0) harvest(k - 1, false);
            s.numberOfCalls += 1;
            const double buffac = 0.5 * (s.list - s.stOuterCutoff);
            dispBefore = dispLast; dispLast = std::sqrt(hdisp[k & 1]); dispKnown += 1;
            if (hdisp[k & 1] > buffac * buffac) {
                // an update was due: take the step back (x is untouched by a second half) and go through the ordinary path
                k_axpy<<<(unsigned int) ((m + 255) / 256), 256, 0, s.stream>>>(d_v, d_a, -0.5 * secondHalfDt, m);
                s.launches += 1;
                s.numberOfCalls -= 1;
                needSync = true;
            }
        }
        if (needSync) {
            int st = NBB200_STATUS_CONTINUE;
            const bool owed = !speculate && k > 0;                                  // step k - 1 has not been harvested yet
            if (owed && publish && !cuda_ok(cudaMemcpyAsync(hke + ((k - 1) & 1), d_ke2 + ((k - 1) & 1), sizeof(double), cudaMemcpyDeviceToHost, s.stream), "D2H kinetic energy")) { ok = false; break; }
            if (owed) {      // its accumulators are turned into energies inside update_common: after the first wait, BEFORE the lists may change
                s.pending = true; s.pendEnergies = eStep[(k - 1) & 1]; s.pendDEdM = dEdM; s.pendHaveGrad = true; s.pendLattice = s.lattice; s.pendAcc = haccSlot[(k - 1) & 1];
            }
            const int updated = update_common(s, box6, forced ? 1 : 0, &st, speculate ? 1 : -1);     // after a take-back the decision is known: no second displacement check
            updates += updated;
            if (updated) { dispKnown = 0; dispLast = dispBefore = 0.0; }             // new reference coordinates
            else if (!speculate) { const double grow = dispLast - dispBefore; dispBefore = dispLast; dispLast += grow; }   // (not measured on the host: keep extrapolating)
            if (st != NBB200_STATUS_CONTINUE) { ok = false; set_status(status, st); break; }
            if (owed) { flush_pending(s); harvest(k - 1, true); }
            if (!enqueue_step(k, speculate)) { ok = false; set_status(status, NBB200_STATUS_LOGIC_ERROR); break; }      // speculate here: the step was taken back
        }
    }
    if (nsteps > 0 && publish) cudaMemcpyAsync(hke + ((nsteps - 1) & 1), d_ke2 + ((nsteps - 1) & 1), sizeof(double), cudaMemcpyDeviceToHost, s.stream);
    cudaStreamSynchronize(s.stream);
    s.pending = false;
    if (ok && nsteps > 0) harvest(nsteps - 1, false);
    cudaEventDestroy(evDisp);
    if (sideStream != nullptr) { cudaStreamSynchronize(sideStream); cudaStreamDestroy(sideStream); cudaEventDestroy(evFork); cudaEventDestroy(evJoin); }
    if (nsteps > 0) cudaMemcpyAsync(d_ke, d_ke2 + ((nsteps - 1) & 1), sizeof(double), cudaMemcpyDeviceToDevice, s.stream);      // the caller's kinetic-energy scalar: the last step's
    cudaStreamSynchronize(s.stream);
    s.mdFused = false; s.gsZeroed = false;
    s.gradOverwrite = savedOverwrite;
    const double *last = eStep[(nsteps > 0 ? nsteps - 1 : 0) & 1];
    if (nbEnergies6 != nullptr) std::memcpy(nbEnergies6, last, sizeof(double) * 6);
    if (bondedEnergies5 != nullptr) std::memcpy(bondedEnergies5, e5, sizeof(e5));
    return updates;
}

void NBModelABFSState_B200_GetStatistics(NBB200State *state, long *numberOfCalls, long *numberOfUpdates)
{
    if (state == nullptr) return;
    const State &s = *reinterpret_cast<State *>(state);
    if (numberOfCalls != nullptr) *numberOfCalls = s.numberOfCalls;
    if (numberOfUpdates != nullptr) *numberOfUpdates = s.numberOfUpdates;
}

void nbb200_set_optimistic_updates(NBB200State *state, int on)
{
    if (state != nullptr) reinterpret_cast<State *>(state)->optimistic = on != 0;
}

void nbb200_set_list_reuse_hint(NBB200State *state, int on)
{
    if (state != nullptr) reinterpret_cast<State *>(state)->listReuseHint = on != 0;
}

void nbb200_set_gradient_overwrite(NBB200State *state, int on)
{
    if (state != nullptr) reinterpret_cast<State *>(state)->gradOverwrite = on != 0;
}

void nbb200_set_partition(NBB200State *state, int rank, int nranks)
{
    if (state == nullptr || nranks < 1 || rank < 0 || rank >= nranks || nranks > State::kMaxPeers) return;
    State &s = *reinterpret_cast<State *>(state);
    s.rank = rank; s.nranks = nranks; s.isNew = true;
}

/* several ranks, host callers: the gradient of the own slab (sorted order, 3 (s1 - s0) doubles) and the atom index of every position of
 * the slab, copied to page-locked host memory on the stream (no wait) -- what a caller that keeps coordinates and gradients on the host
 * downloads per call instead of whole arrays; the indices say which rows of its arrays they are (and which coordinates to upload next time) */
void nbb200_own_slab_to_host(NBB200State *state, double *h_grad, int *h_atoms)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    cudaSetDevice(s.device);
    const size_t cnt = (size_t) std::max(0, s.ownHi - s.ownLo);
    if (cnt == 0) return;
    if (h_grad != nullptr && s.gs != nullptr) cudaMemcpyAsync(h_grad, s.gs + 3 * (size_t) s.ownLo, sizeof(double) * 3 * cnt, cudaMemcpyDeviceToHost, s.stream);
    if (h_atoms != nullptr) cudaMemcpyAsync(h_atoms, s.sAtom.p + s.ownLo, sizeof(int) * cnt, cudaMemcpyDeviceToHost, s.stream);
}

void nbb200_set_restricted_sort(NBB200State *state, int on)
{
    if (state == nullptr) return;
    State &s = *reinterpret_cast<State *>(state);
    s.restrictSort = on != 0; s.isNew = true;
}

}  // extern "C"
