// nbb200_internal.h -- internal declarations of libnbabfs_b200.so (not installed; the public ABI is include/nbabfs_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "symmetry_host.h"

namespace nbb200 {

constexpr int kTile = 32;            // atoms per sort block (one builder warp) and j slots per tile (one force-kernel warp)
constexpr int kCluster = 8;          // atoms per i-cluster of the force kernel: a tile is 8 i atoms x 32 j slots
constexpr int kBuildThreads = 128;   // CTA size of the tile builder (four independent warps)
constexpr unsigned int kEmptySlot = 0x00FFFFFFu;    // j field of an unused tile slot; atoms / sorted positions use 24 bits
constexpr size_t kSmallDoubles = 1 << 16;           // doubles in the page-locked result buffers (State::hsmall, hacc)
constexpr int kMaxAtoms = 0x00FFFFFF;               // hence at most 16.7 M atoms per state

void set_error(const std::string &msg);
bool cuda_ok(cudaError_t e, const char *what);
#define NBB_CUDA(call) do { if (!::nbb200::cuda_ok((call), #call)) return false; } while (0)

// growable device buffer (never shrinks); plain cudaMalloc: sizes are O(N) or O(tiles), allocated at rebuilds only
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    bool ensure(size_t count)
    {
        if (count <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = count + count / 8 + 64;
        if (!cuda_ok(cudaMalloc((void **) &p, want * sizeof(T)), "cudaMalloc")) return false;
        cap = want;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// ABFS constants in fp32 for the tile kernel (derived from the 21 fp64 factors)
struct AbfsF32 {
    float r2Damp, r2On, r2Off;
    float a, b, c, d, c3, d5, qShift1, qShift2, qF0, qAlpha;
    float aF6, aK12, aShift12, aF0, aAlpha, bF3, bK6, bShift6, bF0, bAlpha;
    float rOff, n3, n4, n5, n6, k1, k2;   // factored switching forms
};

// per image, energy-time real-space operation x' = R x + tv, scale, flags (device copy, fp64)
struct ImageOpDev {
    double R[9];
    double tv[3];
    double scale;
    int pureTranslation;
    int pad;
    // in the grid-relative frame of the atom records (X = x - origin = K + xl, K a multiple of 8 A):
    double cr[3];            // rotations: X' = R X + cr, cr = R origin + tv - origin
    float kt[3], tl[3];      // pure translations: tv = kt + tl, kt a multiple of 8 A (exact in fp32), |tl| <= 4 A
};

// per image, list-time data for the tile builder
struct ImageBoxDev {
    double lo[3], hi[3];
};

struct WorkItem { int block, image, tileStart, tileCount; };   // block = i-cluster index (8 sorted atoms); tiles are contiguous

// counters living in device memory (one allocation)
struct DeviceCounters {
    unsigned int extCount;       // extended (halo) atoms appended after the n primary ones
    unsigned int itemCount;      // work items
    unsigned int tileTotal;      // tile pool cursor: tiles reserved in chunks (the tail of a stream's last chunk stays unused)
    unsigned int tilesUsed;      // tiles actually written
    unsigned int overflow;       // bit0: extended capacity, bit1: tile capacity, bit2: item capacity
    unsigned int workCursor;     // dynamic work distribution of the force kernel
    unsigned int pad[2];
};

struct BuildGrid {
    double lo[3];
    double h, invh;
    int dim[3];
    int ncell;
};

struct State;

// ---- list_build.cu
bool build_lists_standalone(State &s, const double *d_x2, int n2);
bool build_lists(State &s);                              // everything between "coordinates on device" and "tiles + items ready"
bool expand_pairs(State &s);                             // explicit (i,j) pairs per list from the tile masks
bool device_bbox(State &s, int nops, double *hostMin, double *hostExt);
bool displacement_check(State &s, double buffacsq, double *maxr2, int *exceeded);
bool displacement_enqueue(State &s, const double *d_x, double *d_out, double *d_zeroOther = nullptr);   // max |x - xref|^2 into a device double, no host synchronisation
bool centre_coordinates(State &s, const double *d_xin, bool doUpdate);   // useCentering: fills s.xc (and s.isoT on updates)
bool touched_ranges(State &s, long *out);
bool touched_ranges_async(State &s, long *d_out);      // the same table written to a device array, no host synchronisation                // per rank slab: [lo, hi) of the sorted positions this rank's lists reference

// ---- mm_terms.cu (used by nbb200_md_run)
}  // namespace nbb200
struct NBB200MMTerms;
namespace nbb200 {
bool mmterms_enqueue_slot(NBB200MMTerms *terms, const double *d_x, double *d_grad, int slot, bool fused = false, bool noCopy = false,
                          const int *perm = nullptr, cudaStream_t on = nullptr);   // perm: gradient rows perm[atom]; on: another stream than the container's
void mmterms_slot_pointers(NBB200MMTerms *terms, int slot, const double **d_energies, double **h_energies);
void mmterms_read_slot(NBB200MMTerms *terms, int slot, double *energies5);
void mmterms_reset_slots(NBB200MMTerms *terms);
bool langevin_first(State &s, double *d_x, double *d_v, const double *d_a, const double *d_mass, const double *f7, unsigned long long seed, unsigned long long step);
bool langevin_first_disp(State &s, double *d_x, double *d_v, const double *d_a, const double *d_mass, const double *f7, unsigned long long seed,
                         unsigned long long step, double *d_out, double *d_zeroOther);

// ---- force_kernels.cu
bool upload_spline_tables(State &s);                     // s.spl -> s.splF64 / s.splPoly
bool launch_forces(State &s, double *d_grad, bool sortedOnly);
bool unsort_gradients(State &s, long s0, long s1, double *d_grad, bool assign = false, bool clear = false);   // honours s.condDisp
void init_force_kernel_attributes();

struct State {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    bool timing = false;
    long launches = 0;

    // topology / parameters
    int n = 0, ntypes = 0, ntypes14 = 0, nexcl = 0, n14 = 0;
    DevBuf<float> q32;
    DevBuf<double> q64;
    DevBuf<int> ljtype;
    DevBuf<float2> ljAB;            // [nt*nt] (A, B) expanded table, fp32
    DevBuf<unsigned char> typeFree; // [nt + 1] 1 = the type has no Lennard-Jones interaction with any type (TIP3P hydrogens ...); [nt] = the null type
    DevBuf<double2> ljAB14;         // [nt14*nt14] fp64 for the 1-4 kernel
    DevBuf<int> exclPtr, exclCol;   // symmetric CSR
    DevBuf<int2> pairs14;
    std::vector<int2> pairs14All;                // as given to SetUp; pairs14 / n14 hold the ones with at least one free atom
    DevBuf<unsigned char> fixedFlag;             // per atom, 1 = fixed (NBModelABFSState_SetUp's fixedAtoms); unused when nfixed == 0
    int nfixed = 0;
    DevBuf<unsigned char> qcFlag;                // per atom, 1 = pure QC atom: on no MM/MM list (NBModelABFSState_SetUp's qcAtoms); unused when nqc == 0
    int nqc = 0;
    std::vector<unsigned char> hostQC;
    // QC/MM entry points (qcmm.cu): device copies of what a call needs
    DevBuf<int> qcIdxDev, qcSlotDev;
    DevBuf<double> qcWork, qcImagesDev, qcMMqDev, qcAccDev, qcGradDev, qcSplDev, qcPotDev;
    DevBuf<double2> qcLJDev;
    std::vector<double2> hostLJ64;               // [nt*nt] (A, B) expanded table in fp64 (QC/MM LJ term, qcmm.cu)
    std::vector<int> hostExclPtr, hostExclCol;   // host copy of the exclusion CSR (isolates for useCentering)
    std::vector<unsigned char> hostFixed;
    // useCentering (NBModelABFSState_SetUpCentering): isolates = connected components of the exclusion graph
    bool useCentering = false;
    int nisolates = 0;
    DevBuf<int> isoPtr, isoIdx;
    DevBuf<double> xc, isoT;                     // centred coordinates, per-atom isolate translations of the last update
    Transformations trans;

    // options (NBModelABFS / PairwiseInteractionABFS)
    double damp = 0.5, inner = 8.0, outer = 12.0, list = 13.5, dielectric = 1.0, scale14 = 1.0;
    bool checkForInverses = true;
    int expandFactor = 0;
    double factors[21];
    // interaction form (PairwiseInteractionABFS.useAnalyticForm / splinePointDensity); tables rebuilt when the cutoffs change
    bool useAnalytic = true, splValid = false;
    int splineDensity = 50;
    SplineTables spl;
    DevBuf<double> splF64;          // x[n], then y[n], h[n] of the electrostatic, LJ-A and LJ-B splines (1-4 kernel, fp64)
    DevBuf<float4> splPoly;         // [3][n] per-interval cubics in fp32 (tile kernel), last row zero

    // reference-state bookkeeping (NBModelABFSState)
    bool isNew = true;
    double stListCutoff = 0.0, stOuterCutoff = 0.0;
    Lattice lattice, refLattice;
    bool haveRefLattice = false;
    long numberOfCalls = 0, numberOfUpdates = 0;

    // coordinates
    DevBuf<double> x, xref, grad;
    const double *xcur = nullptr;   // coordinates of the current call (s.x.p or a caller-owned device array)
    double *hx = nullptr, *hgrad = nullptr;      // pinned staging
    double *hsmall = nullptr;                    // pinned small results
    // deferred energy call (MMMMEnergyDeviceDeferred): the accumulators land in their own pinned buffer; energies / dE/dM are written to the
    // caller's arrays at the state's next synchronisation point (the list-update decision of the next Update, or nbb200_flush)
    double *hacc = nullptr;
    bool pending = false, pendHaveGrad = false;
    double *pendEnergies = nullptr, *pendDEdM = nullptr, *pendAcc = nullptr;
    Lattice pendLattice;
    void *hops = nullptr;                        // pinned staging of the image operations
    Mat3 opsLattice{}; long opsGeneration = -1; bool opsValid = false;
    size_t hcap = 0;

    // image plan of the current lists
    ImagePlan plan;
    std::vector<long> imagePairs;                // per candidate image (list pairs), -1 = not fetched yet
    long primaryPairs = -1;
    bool pairCountsValid = false;
    DevBuf<ImageOpDev> imageOps;                 // [1 + nimages], slot 0 = identity (primary)
    DevBuf<ImageBoxDev> imageBoxes;
    DevBuf<double> baseOpsDev;                   // per transformation 12 doubles
    DevBuf<double> visitDisp;                    // per visit 3 doubles
    DevBuf<int> visitInfo;                       // per visit: t, image
    DevBuf<double> bboxDev;                      // reduction output
    DevBuf<unsigned int> ticket;                 // CTA counters of the kernels that finish with last_block_done (left at zero by them)
    DevBuf<double> mdScalars;                    // nbb200_md_run: two-slot device scalars (displacement maximum, kinetic energy)

    // extended atoms and the sort
    BuildGrid grid;
    int nsets = 1;                               // 1 + candidate images
    size_t extCap = 0;
    DevBuf<double> eX;  DevBuf<int> eAtom; DevBuf<int> eSet; DevBuf<int> eKey; DevBuf<unsigned long long> eSortBuf;
    DevBuf<unsigned int> cellStart, cellFill, scanTmp;
    DevBuf<int> order, order2;
    DevBuf<double> sX; DevBuf<int> sAtom; DevBuf<int> invPerm;
    int nblocks = 0;
    DevBuf<double> blockBox;                     // per block: min[3], max[3], center[3]
    // tiles: 32 descriptor words each (low 24 bits: sorted position of the j atom, high 8 bits: row mask of the lane)
    int chunkTiles = 0;                          // tiles per pool chunk = max tiles per work item
    size_t tileCap = 0;                          // pool capacity in tiles
    bool rawJ = false;                           // stand-alone cross lists: the j field of sets > 0 is an index into the second array
    DevBuf<unsigned int> tileDesc;
    DevBuf<WorkItem> items; size_t itemCap = 0;
    // rolling prune (force_kernels.cu): the force kernel walks an INNER copy of the tile pool that holds only the j entries with a pair inside
    // outerCutoff + pruneBuffer; it is refreshed on the device when the lists were rebuilt, the lattice changed or an atom moved by more
    // than pruneBuffer / 2 since the last prune (decided on the device, no host round trip)
    double pruneBuffer = 0.5;                    // A; <= 0 or outer + buffer >= list: no pruning, the force kernel walks the pool as built
    DevBuf<unsigned int> tileDescIn; DevBuf<WorkItem> itemsIn;
    DevBuf<double> xprune;                       // coordinates at the last prune
    DevBuf<unsigned long long> pruneDisp;        // two slots (call parity): max |x - xprune|^2 of the current call as the bits of a double, [2]: number of prunes
    long pruneCall = 0, pruneListGeneration = -1, outerCallGeneration = -1;   // outerCallGeneration: the lists of that update have had their first energy call (on the pool as built)
    Mat3 pruneLattice{}; bool pruneLatticeValid = false;
    // per energy call: atom records in sorted order (A: xl, yl, zl, q; B: Kx, Ky, Kz, LJ type) and the sorted-order gradient
    DevBuf<float4> recA, recB;
    DevBuf<double> gradSorted;
    double *gs = nullptr, *gsExternal = nullptr;   // sorted-order gradient of the last call: own buffer or one set by the caller (section 8e)
    int ownLo = 0, ownHi = 0;                    // sorted positions of the i-blocks this rank owns (whole system for one rank)
    bool listReuseHint = false;                  // the caller evaluates many calls per list (dynamics): the builder favours the force kernel (nbb200_set_list_reuse_hint)
    bool restrictSort = false;                   // several ranks: sort / pack only the cells this rank's slab can see (nbb200_set_restricted_sort)
    DevBuf<unsigned char> cellNeed;              // [2 * ncell]: image entries wanted in this cell | primary atoms wanted in this cell
    DevBuf<int> rangeTab; DevBuf<long> rangeOut;
    // peer memory (CUDA IPC, one process per GPU of one node): every rank's sorted-order gradient accumulator and sorted positions
    static constexpr int kMaxPeers = 16;
    DevBuf<double> symGs, symXs;                 // this rank's buffers (cudaMalloc: exportable)
    DevBuf<double> symSig;                       // signal area written by the peers: step flags, displacement maxima, the 15 scalars
    double *peerGs[kMaxPeers] = {}, *peerXs[kMaxPeers] = {}, *peerSig[kMaxPeers] = {};
    DevBuf<double> sigStage;                     // device staging: [0] local displacement maximum, [1..16] scalars, [17..32] results, [33] timeout flag
    bool peerOpened[kMaxPeers] = {};
    DevBuf<double> scalMat; bool scalMatValid = false; long scalMatGeneration = -1; Mat3 scalMatLattice;      // nbb200_peer_signal_end_device: dE/dM as a linear map of the image sums
    // host callers, several ranks: atom-order chunk buffers (positions as uploaded by the rank that holds the host rows; gradients as downloaded by it)
    DevBuf<double> symXc, symGc;
    double *peerXc[kMaxPeers] = {}, *peerGc[kMaxPeers] = {};
    bool peerChunkOpened[kMaxPeers] = {};
    cudaEvent_t chunkEvents[4] = {nullptr, nullptr, nullptr, nullptr};
    const void *hostGChecked = nullptr; bool hostGPinned = false;
    const void *hostXChecked = nullptr; bool hostXPinned = false;      // nbb200_chunk_upload: is the caller's array page-locked?
    bool peersReady = false;
    // optimistic update decision (nbb200_set_optimistic_updates): Update enqueues the displacement check without waiting for it, the
    // energy call that follows evaluates on the current lists and its one synchronisation brings the decision back; when an update was
    // due after all the gradient is not handed out (the unsort pass looks at the device-side maximum), the lists are rebuilt and the
    // call is repeated
    bool optimistic = false, optPending = false, keepLattice = false, optDispCopied = false;
    double optThr2 = 0.0;
    DevBuf<double> optDisp;                      // device: max |x - xref|^2 of the pending decision
    const double *condDisp = nullptr;            // unsort pass: skip when *condDisp > condThr2
    double condThr2 = 0.0;
    bool gradOverwrite = false;                  // MMMMEnergy (host arrays) sets the caller's gradient instead of accumulating
    // results published by the unsort pass of an energy call instead of a copy operation (force_kernels.cu: PublishArgs); set by energy_enqueue
    const double *pubSrc[2] = {nullptr, nullptr}; double *pubDst[2] = {nullptr, nullptr}; int pubCount[2] = {0, 0}; bool pubDone = false;
    // Langevin dynamics: the random vectors are projected on the translation constraints (mm_terms.cu: LangevinConstraint)
    bool lcOn = false, lcValid = false; double lcTotalMass = 0.0; unsigned long long lcStep = 0, lcSeed = 0; DevBuf<double> lcSums; DevBuf<double2> lcW;
    // nbb200_md_run: two scalars stored into page-locked memory by the first kernel of the energy call (k_pack_records), and an event behind it
    const double *prePubSrc[2] = {nullptr, nullptr}; double *prePubDst[2] = {nullptr, nullptr}; cudaEvent_t prePubEvent = nullptr; bool prePubDone = false;
    // nbb200_md_run: the unsort pass also does the second half of the integrator step (force_kernels.cu: k_unsort_second_half)
    struct SecondHalf { double *v = nullptr, *a = nullptr; const double *mass = nullptr; double dt = 0.0; double *ke = nullptr, *zeroOther = nullptr;
                        const double *pubSrc = nullptr; double *pubDst = nullptr; int pubCount = 0; } secondHalf;
    bool secondHalfDone = false;
    cudaStream_t sideStream = nullptr; cudaEvent_t evPack = nullptr, evSide14 = nullptr;      // launch_forces: the 1-4 kernel next to the tile kernel
    cudaEvent_t preUnsortEvent = nullptr;        // nbb200_md_run: the unsort pass waits for it (bonded terms accumulated into the sorted gradient on a side stream)
    bool mdFused = false;                        // inside nbb200_md_run: memsets folded into neighbouring kernels (accumulators by k_pack_records, sorted gradient by k_unsort_gradients)
    bool gsZeroed = false;                       // the caller zeroed the sorted gradient for this call already (before the ranks' barrier)                        // touched sorted range per rank slab (min, max+1)
    DevBuf<unsigned long long> setPairs;         // per set list-pair counts
    DeviceCounters *counters = nullptr;
    DeviceCounters hostCounters{};
    // accumulators of the force kernels: per set 16 doubles {eq, elj, G[3], W[9], pad}; slot nsets = 1-4 terms
    DevBuf<double> accum;
    // explicit pairs
    DevBuf<int> pairBuf; DevBuf<unsigned long long> pairCursor;
    std::vector<unsigned long long> pairOffsets;
    bool pairsExpanded = false;

    // partition over ranks
    int rank = 0, nranks = 1;

    // timing
    cudaEvent_t ev[12] = {};
    double timings[8] = {};
    bool haveEvents = false;
};


#ifdef __CUDACC__
// true in every thread of the LAST CTA of the grid to arrive here, after the global writes and atomics of all CTAs before this point have
// become visible: that CTA may read what the grid reduced with atomics and hand it on (here: store it into page-locked host memory, which
// spares the stream a small copy operation).  `ticket` counts the CTAs and is left at zero for the next launch.  Every thread of the CTA
// must call it.
__device__ __forceinline__ bool last_block_done(unsigned int *ticket)
{
    __shared__ bool lastBlock;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        lastBlock = t == gridDim.x - 1;
        if (lastBlock) { *ticket = 0u; __threadfence(); }
    }
    __syncthreads();
    return lastBlock;
}
#endif

}  // namespace nbb200
