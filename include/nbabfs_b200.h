/* include/nbabfs_b200.h -- C-ABI of libnbabfs_b200.so (B200 / sm_100a).
 *
 * Drop-in boundary for ONE hot path of pDynamo 1.9.0: the NBModelABFS MM/MM non-bonded term
 * (force-switched Coulomb + Lennard-Jones energy and gradients), the cutoff pair lists it is evaluated
 * over (PairListGenerator) and the explicit periodic images (ImageList / SymmetryParameters).
 *
 * Plain pointers and sizes only; no torch / CUDA types in the signatures.  Every entry point names the
 * reference interface it replaces (paths relative to the pDynamo tree; pM = pMolecule-1.9.0/extensions,
 * pC = pCore-1.9.0/extensions).  INTEGRATION.md shows the Cython binding a pDynamo maintainer would add.
 *
 * Conventions kept from the reference boundary:
 *   - Status out-parameter, written only on failure; NBB200_STATUS_CONTINUE (=16, pC/cinclude/Status.h:37)
 *     means OK; NULL return = allocation / device failure.  CUDA errors never abort: they map to a status.
 *   - all functions are NULL tolerant;
 *   - coordinates and gradients are N x 3 row-major fp64 (Coordinates3, pC/cinclude/Coordinates3.h:26),
 *     gradients and dE/dM are ACCUMULATED into the caller's arrays (System.Energy adds bonded terms first);
 *   - synchronous at return of NBModelABFS_B200_MMMMEnergy (energies are read immediately by the caller,
 *     pM/pyrex/pMolecule.NBModelABFS.pyx:119-121).
 */
#ifndef NBABFS_B200_H
#define NBABFS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pC/cinclude/Status.h:21-70 (values used at this boundary) */
#define NBB200_STATUS_NORMAL            0
#define NBB200_STATUS_LOGIC_ERROR      10
#define NBB200_STATUS_CONTINUE         16
#define NBB200_STATUS_INVALID_ARGUMENT 22
#define NBB200_STATUS_OUT_OF_MEMORY    41

/* Opaque device-resident state: replaces NBModelABFSState (pM/cinclude/NBModelABFSState.h:33-124) together
 * with the option blocks of NBModelABFS (pM/cinclude/NBModelABFS.h) and PairwiseInteractionABFS
 * (pM/cinclude/PairwiseInteraction.h:24-37) it is evaluated with. */
typedef struct NBB200State NBB200State;

/* Energy slots, in the order of NBModelABFSState.GetEnergies (pM/pyrex/pMolecule.NBModelABFSState.pyx:41-59). */
enum { NBB200_EMMEL = 0, NBB200_EMMLJ = 1, NBB200_EMMEL14 = 2, NBB200_EMMLJ14 = 3, NBB200_EIMMMEL = 4, NBB200_EIMMMLJ = 5 };

/* ---- library ------------------------------------------------------------------------------------- */
int         nbb200_device_count(void);               /* number of visible CUDA devices (0 without a GPU/driver) */
const char *nbb200_last_error(void);                 /* text of the last failure on this thread */
const char *nbb200_version(void);

/* ---- state ----------------------------------------------------------------------------------------
 * replaces NBModelABFSState_SetUp (pM/csource/NBModelABFSState.c:316-420) for a pure-MM system:
 *   charges[n], ljtypes[n]            <- MMAtomContainer.data[i].{charge,ljtype} (pM/cinclude/MMAtomContainer.h:24-35)
 *   tableindex[nt*nt], tableA/B[...]  <- LJParameterContainer (pM/cinclude/LJParameterContainer.h), normal and 1-4
 *   exclPairs[2*nexcl], pairs14[2*n14]<- exclusions / interactions14 PairLists as index pairs
 *   rot[9*ntrans], trans[3*ntrans]    <- Transformation3Container items (fractional); ntrans = 0: no symmetry
 * device: CUDA ordinal.  Inputs are copied (the reference aliases them, NBModelABFSState.c:384-395). */
NBB200State *NBModelABFSState_B200_SetUp(int device, int n, const double *charges, const int *ljtypes,
                                         int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                                         int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                                         int nexcl, const int *exclPairs, int n14, const int *pairs14,
                                         int ntrans, const double *rot, const double *trans, int *status);
/* the fixedAtoms argument of NBModelABFSState_SetUp (freeSelection = complement, pM/csource/NBModelABFSState.c:345): the lists keep a
 * pair only if at least one of its atoms is free (orSelection of PairListGenerator_*, pC/csource/PairListGenerator.c:118-139; 1-4 list:
 * GenerateLists14, pM/csource/NBModelABFS.c:1113-1128) and CheckForUpdate ignores fixed atoms (:723-739).  Call after SetUp, before the
 * first Update; nfixed = 0 clears.  Marks the state new (lists are rebuilt). */
void NBModelABFSState_B200_SetFixedAtoms(NBB200State *state, int nfixed, const int *fixed, int *status);
/* the qcAtoms argument of NBModelABFSState_SetUp, first part (pM/csource/NBModelABFSState.c:348-353: mmSelection = complement of the pure
 * QC selection): the listed atoms leave every MM/MM list -- primary and image lists (and-selection of GenerateLists / GenerateImageLists,
 * pM/csource/NBModelABFS.c:508-623) and the 1-4 list (GenerateLists14, :1113-1128) -- so that NBModelABFS_B200_MMMMEnergy returns what
 * NBModelABFS_MMMMEnergy returns with a QC region present.  QC regions without boundary (link) atoms only; not with useCentering or
 * partitions.  The QC/MM entry points (NBModelABFS_QCMMEnergyLJ / _QCMMPotentials / _QCMMGradients, NBModelABFS.c:306-498) follow.
 * Call after SetUp, before the first Update; nqc = 0 clears.  Marks the state new. */
void NBModelABFSState_B200_SetQCAtoms(NBB200State *state, int nqc, const int *qcAtoms, int *status);
/* The QC/MM entry points for a QC region set with NBModelABFSState_B200_SetQCAtoms (no boundary atoms), after an Update.  Supported:
 * vacuum, P1 cells, and cells with space-group operations when every atom is a QC atom (QC/QC image terms only); analytic form of the
 * MM/MM interaction for the LJ term.  One fp64 launch per call over (QC atom) x (image operation) x (atom): the sums do not depend on
 * the reference's pair lists (every pair within the outer cutoff is on a valid list, pairs beyond it are skipped).
 *
 * NBModelABFS_QCMMEnergyLJ (pM/csource/NBModelABFS.c:306-378): energies4 = {eqcmmlj, eqcmmlj14 (0 without boundary atoms), eimqcmmlj,
 * eimqcqclj}; grad[3 n] (host, nullable) and dEdM[9] (nullable: the image terms' SymmetryParameterGradients_ImageDerivatives) are
 * accumulated into. */
void NBModelABFS_B200_QCMMEnergyLJ(NBB200State *state, double *energies4, double *grad, double *dEdM, int *status);
/* NBModelABFS_QCMMPotentials (NBModelABFS.c:454-498): electrostatic potentials of the MM charges (and their images) on the QC atoms in
 * atomic units, potentials[nqc], and the QC/QC image potentials qcqcPotentials[nqc (nqc + 1) / 2] (packed lower triangle, SymmetricMatrix
 * order); both are incremented, not reset, like the reference's; either may be NULL.  QC atoms are numbered in ascending atom index. */
void NBModelABFS_B200_QCMMPotentials(NBB200State *state, double *potentials, double *qcqcPotentials, int *status);
/* NBModelABFS_QCMMGradients (NBModelABFS.c:383-449): gradients of the QC/MM and QC/QC image electrostatic interactions for the QC charges
 * qcCharges[nqc], regular units; grad[3 n] (host) and dEdM[9] (nullable) are accumulated into. */
void NBModelABFS_B200_QCMMGradients(NBB200State *state, const double *qcCharges, double *grad, double *dEdM, int *status);
/* replaces NBModelABFSState_SetUpCentering (pM/csource/NBModelABFSState.c:425-450; NBModelABFS option useCentering): the isolates
 * (connected components of the exclusion graph; those with a fixed atom stay) are moved into the primary cell by whole lattice vectors at
 * every list update and carried along in between (NBModelABFSState_InitializeCoordinates3, :278-311); lists and energies are evaluated on
 * the centred coordinates.  Call after SetUp / SetFixedAtoms.  Off without exclusions, without transformations or with a single isolate,
 * as in the reference. */
void NBModelABFSState_B200_SetUpCentering(NBB200State *state, int useCentering, int *status);
/* replaces NBModelABFSState_Deallocate (pM/csource/NBModelABFSState.c:137-188) */
void NBModelABFSState_B200_Deallocate(NBB200State **state);

/* replaces the option fields set by NBModelABFS.SetOptions (pM/pyrex/pMolecule.NBModelABFS.pyx:140-179):
 * NBModelABFS.{dampingCutoff,innerCutoff,outerCutoff,listCutoff,dielectric,electrostaticScale14,
 * checkForInverses,imageExpandFactor} and the three cutoffs of PairwiseInteractionABFS. */
void NBModelABFS_B200_SetOptions(NBB200State *state, double dampingCutoff, double innerCutoff, double outerCutoff,
                                 double listCutoff, double dielectric, double electrostaticScale14,
                                 int checkForInverses, int imageExpandFactor);

/* replaces PairwiseInteractionABFS.SetOptions(useAnalyticForm, splinePointDensity) + MakeSplines for the MM/MM interaction
 * (pM/pyrex/pMolecule.PairwiseInteraction.pyx:204-239, called from NBModelABFS.CheckPairwiseInteractions, pMolecule.NBModelABFS.pyx:73-82):
 * useAnalyticForm = 0 selects the cubic-spline branch of PairwiseInteractionABFS_MMMMEnergy (pM/csource/PairwiseInteraction.c:431-531)
 * for the primary, image and 1-4 terms; the three Delta/Delta tables (PairwiseInteractionABFS_Make{Electrostatic,LennardJonesA,
 * LennardJonesB}Spline, :148-285, splinePointDensity points per Angstrom, reference default 50) are built from the state's cutoffs
 * and rebuilt by NBModelABFS_B200_SetOptions.  Default: analytic form (pM/csource/PairwiseInteraction.c:34). */
void PairwiseInteractionABFS_B200_SetInteractionForm(NBB200State *state, int useAnalyticForm, int splinePointDensity, int *status);
/* host helper: the table PairwiseInteractionABFS_Make*Spline builds (which = 0 electrostatic in kJ/mol, 1 LJ-A, 2 LJ-B, 3 electrostatic in
 * atomic units = useAtomicUnits, the spline of the QC/MM and QC/QC interactions, pM/csource/PairwiseInteraction.c:187): abscissae
 * x = r^2, ordinates y and second derivatives h as CubicSpline_MakeFromReal1DArrays leaves them (pC/csource/CubicSpline.c:309-420).
 * Returns the number of points (pM/cinclude/PairwiseInteraction.h:166-167); with x, y or h NULL only that. */
int PairwiseInteractionABFS_B200_MakeSpline(int which, double dampingCutoff, double innerCutoff, double outerCutoff, int splinePointDensity,
                                            double *x, double *y, double *h);

/* replaces NBModelABFSState_Initialize (pM/csource/NBModelABFSState.c:231-273) + NBModelABFS_Update
 * (pM/csource/NBModelABFS.c:508-623): takes this call's coordinates (host, xyz[3n]) and lattice
 * box6 = {a,b,c,alpha,beta,gamma} (ignored when ntrans = 0), decides with the reference's heuristics
 * (CheckForUpdate :691-746, CheckForImageUpdate :635-684) whether the lists must be rebuilt and rebuilds
 * them on the device.  forceNew != 0 sets state->isNew first.  Returns 1 if lists were rebuilt, else 0. */
int NBModelABFS_B200_Update(NBB200State *state, const double *xyz, const double *box6, int forceNew, int *status);
/* same, coordinates already resident on the device (d_xyz: device pointer, fp64 [3n]) */
int NBModelABFS_B200_UpdateDevice(NBB200State *state, const double *d_xyz, const double *box6, int forceNew, int *status);

/* replaces NBModelABFS_MMMMEnergy (pM/csource/NBModelABFS.c:228-301): energies[6] (slots above) are set;
 * grad[3n] (host, nullable) and dEdM[9] (nullable) are accumulated into. */
void NBModelABFS_B200_MMMMEnergy(NBB200State *state, double *energies, double *grad, double *dEdM, int *status);
/* on != 0: NBModelABFS_B200_MMMMEnergy SETS grad[3n] instead of accumulating into it (dE/dM is still accumulated).  For a caller
 * that evaluates the NB term first: System.Energy's zero fill of gradients3 (pMolecule-1.9.0/pMolecule/System.py:272-318) and the
 * upload of the array are then not needed.  Default 0 = the reference's accumulation.  The device-array calls (...MMMMEnergyDevice,
 * ...MMMMEnergyDeviceDeferred) honour it as well: d_grad is then SET by the NB term (no zero fill by the caller). */
void nbb200_set_gradient_overwrite(NBB200State *state, int on);
/* Optimistic update decision for the device-array calls (one host synchronisation per Update + MMMMEnergyDevice pair instead of two):
 * NBModelABFS_B200_UpdateDevice enqueues CheckForUpdate's displacement test (pM/csource/NBModelABFS.c:691-746) and returns 0 at once;
 * the NBModelABFS_B200_MMMMEnergyDevice that follows evaluates on the current lists and reads the decision with its own results.  When
 * an update was due, nothing of that evaluation is handed out: the lists are rebuilt at the same coordinates and the call is evaluated
 * again, so the results are those of the plain call sequence; NBModelABFSState numberOfUpdates is current after the energy call. */
void nbb200_set_optimistic_updates(NBB200State *state, int on);
/* Hint: the caller evaluates many calls per list (dynamics, minimisation).  Small systems then give the tile builder whole sort blocks per
 * warp instead of dealing their cell rows to several warps: the rebuild takes longer, every energy call on the lists is faster (fewer
 * padded tiles).  Results do not depend on it beyond the fp32 summation order. */
void nbb200_set_list_reuse_hint(NBB200State *state, int on);
/* NBModelABFSState.numberOfCalls / .numberOfUpdates (pM/cinclude/NBModelABFSState.h:38,42, printed by NBModelABFSState.StatisticsSummary) */
void NBModelABFSState_B200_GetStatistics(NBB200State *state, long *numberOfCalls, long *numberOfUpdates);
/* same, gradients accumulated into a device array d_grad[3n] (nullable) */
void NBModelABFS_B200_MMMMEnergyDevice(NBB200State *state, double *energies, double *d_grad, double *dEdM, int *status);

/* The device call WITHOUT its host synchronisation, for MD loops that keep everything on the device: the kernels are enqueued and the call
 * returns; energies[6] and dEdM[9] (caller-owned, must stay valid) are written at the state's next synchronisation point -- the
 * list-update decision of the next NBModelABFS_B200_Update* call, any other energy call, or nbb200_flush.  d_grad is complete in stream
 * order as usual.  With this, one MD step needs a single host wait (the update decision) instead of three. */
void NBModelABFS_B200_MMMMEnergyDeviceDeferred(NBB200State *state, double *energies, double *d_grad, double *dEdM, int *status);
void nbb200_flush(NBB200State *state, int *status);
/* enqueue a device-to-host copy of a few bytes (e.g. the kinetic energy of nbb200_vv_second_half) into page-locked host memory on the state's
 * stream; valid after the next synchronisation point */
void nbb200_copy_to_host_async(NBB200State *state, const void *d_src, void *h_pinned_dst, size_t bytes);

/* ---- list inspection: what PairList / ImageList hold in the reference ------------------------------
 * statistics of NBModelABFSState (pM/cinclude/NBModelABFSState.h:38-71) */
long NBModelABFSState_B200_NumberOfPairs(NBB200State *state, int image /* -1: primary list nbmmmm; k >= 0: image k */);
int  NBModelABFSState_B200_NumberOfImages(NBB200State *state);
long NBModelABFSState_B200_NumberOfImagePairs(NBB200State *state);
long NBModelABFSState_B200_NumberOf14Pairs(NBB200State *state);
/* Image record (pM/cinclude/ImageList.h:19-28): info = {t, a, b, c, npairs, 0}, scale */
void NBModelABFSState_B200_GetImageInfo(NBB200State *state, int image, int *info, double *scale);
/* explicit (i, j) pairs of one list, expanded from the tile masks on the device:
 * replaces PairList_ToIntegerPairArray (pC/csource/PairList.c:253).  pairs[2*npairs] host; returns npairs. */
long NBModelABFSState_B200_GetPairs(NBB200State *state, int image, int *pairs, int *status);

/* ---- stand-alone generators -----------------------------------------------------------------------
 * replace PairListGenerator_SelfPairListFromCoordinates3 (pC/csource/PairListGenerator.c:530-553) and
 * PairListGenerator_CrossPairListFromDoubleCoordinates3 (:414-446) for the no-radii / no-selection case:
 * all pairs with (dx*dx + dy*dy) + dz*dz <= cutoff*cutoff in fp64, minus exclusions for the self list.
 * On return *pairs is a malloc'ed int[2*npairs] (free with nbb200_free); returns npairs or -1. */
long PairListGenerator_B200_SelfPairListFromCoordinates3(int device, int n, const double *xyz, double cutoff,
                                                         int nexcl, const int *exclPairs, int **pairs, int *status);
long PairListGenerator_B200_CrossPairListFromDoubleCoordinates3(int device, int n1, const double *xyz1, int n2, const double *xyz2,
                                                                double cutoff, int **pairs, int *status);
void nbb200_free(void *p);

/* Page-locked host arrays.  Coordinates3 / gradient arrays allocated here (instead of Memory_Allocate_Array_Real,
 * pC/csource/Memory.c) are transferred by DMA without a staging copy; Update / MMMMEnergy detect them automatically
 * (cudaPointerGetAttributes) and fall back to an internal pinned staging buffer for ordinary pageable arrays. */
void *nbb200_host_alloc(size_t bytes);
void  nbb200_host_free(void *p);

/* PairwiseInteractionABFS_MakeFactors (pM/csource/PairwiseInteraction.c:89-141): host helper, out[21] */
void PairwiseInteractionABFS_B200_MakeFactors(double dampingCutoff, double innerCutoff, double outerCutoff, double *out21);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------
 * Use the caller's CUDA stream (e.g. torch's current stream) so that the caller's events bracket the work. */
void nbb200_set_stream(NBB200State *state, void *cudaStream);
/* the FP32 (non-tensor) roofline denominator measured on this device: sustained TFLOP/s of register-only FMA chains on every lane of
 * every SM (about 60 ms; SURVEY.md 8d asks for the measured figure instead of lanes x clock) */
double nbb200_measure_fp32_peak(int device, int *status);
/* Device time of the phases of the LAST Update / MMMMEnergy pair, from CUDA events on the state's stream (ms):
 * out[0] list rebuild (all kernels), out[1] tile-pair force kernel, out[2] 1-4 kernel, out[3] displacement check,
 * out[4] explicit pair expansion (last GetPairs), out[5..7] reserved.  Enabled by nbb200_enable_timing(state, 1). */
void nbb200_enable_timing(NBB200State *state, int on);
void nbb200_get_timings(NBB200State *state, double *out8);
/* counters of the current lists: out[0] tiles, out[1] work items, out[2] i-blocks, out[3] extended (halo) atoms,
 * out[4] list pairs (popcount of all masks), out[5] kernel launches since SetUp, out[6] tiles per pool chunk (= per work item), out[7] images */
void nbb200_get_counters(NBB200State *state, long *out8);
/* spatial decomposition over ranks (section 8e): this state evaluates only the i-blocks b with b % nranks == rank
 * in units of contiguous chunks; energies/gradients are then partial sums to be reduced by the caller (NCCL). */
void nbb200_set_partition(NBB200State *state, int rank, int nranks);
/* Several ranks: restrict this rank's sort (scatter, in-cell sort, grouping, block boxes, record pack) to the grid cells its slab can see
 * -- the cells within listCutoff of an owned atom plus the home cells of the atoms whose images fall there; the cell histogram and its
 * prefix sum (the GLOBAL sorted positions all ranks agree on) are still computed from all atoms.  Outside those cells the rank has no
 * sorted-position -> atom map: callers exchange positions through the peer-memory transport below (owners publish their atoms' indices
 * with their positions); the message transport (nbb200_gather_sorted / _scatter_sorted over whole slabs) needs the unrestricted sort.
 * The reference has no counterpart (its only parallelism is an OpenMP team, NBModelABFSState.c:401). */
void nbb200_set_restricted_sort(NBB200State *state, int on);
/* Host callers on several ranks: enqueue the copies of the own slab's gradient (sorted order, 3 (s1 - s0) doubles) and of the atom index of
 * every position of the slab into page-locked host memory; stream ordered, no wait.  (s0, s1 from nbb200_get_slab.) */
void nbb200_own_slab_to_host(NBB200State *state, double *h_grad, int *h_atoms);
/* host-side companions (plain CPU loops over rows of three doubles, a few threads: NBB200_HOST_THREADS, default 4):
 * out[k] = x[atoms[k]] (what to upload) and g[atoms[k]] += in[k] (where the downloaded gradients go); the atoms of a slab are distinct */
void nbb200_host_gather_rows(const double *x, const int *atoms, long count, double *out);
void nbb200_host_scatter_add_rows(double *g, const int *atoms, long count, const double *in);

/* Host callers on several ranks, without row gathers on the host (DistributedNB.call_host): rank r moves the contiguous rows
 * [n r / R, n (r + 1) / R) of the caller's host arrays, the device redistributes over peer memory.  Per call:
 *   nbb200_chunk_upload(h_x, a0, count); nbb200_chunk_signal(step, 0); nbb200_chunk_wait(step, 0); nbb200_chunk_gather_owned(d_x);
 *   ... the distributed call (nbb200_peer_begin ... nbb200_peer_wait_end) ...
 *   nbb200_chunk_scatter_gradients(); nbb200_chunk_signal(step, 1); nbb200_chunk_wait(step, 1); nbb200_chunk_download_add(h_g, a0, count);
 * The chunk buffers are exported / imported like the other peer buffers (two IPC handles, 128 bytes).  There is no counterpart in the
 * reference (one process, host arrays). */
int nbb200_peer_export_chunks(NBB200State *state, char *handles128);
int nbb200_peer_import_chunks(NBB200State *state, int rank, const char *handles128);
int nbb200_peer_attach_local_chunks(NBB200State *state, int rank, NBB200State *other);
void nbb200_chunk_upload(NBB200State *state, const double *h_x, long a0, long count);
void nbb200_chunk_signal(NBB200State *state, long step, int which);
void nbb200_chunk_wait(NBB200State *state, long step, int which);
void nbb200_chunk_gather_owned(NBB200State *state, double *d_x);
void nbb200_chunk_scatter_gradients(NBB200State *state);
int nbb200_chunk_download_add(NBB200State *state, double *h_g, long a0, long count);   /* 0: failed (time-out of a peer, bad range) */
/* overwrite != 0: the rows are SET instead of added to (System.Energy's own, freshly zeroed gradient array with the NB term first: the zero fill
 * is folded into the call, as nbb200_set_gradient_overwrite does on one GPU); a page-locked h_g then receives them by one DMA */
int nbb200_chunk_download(NBB200State *state, double *h_g, long a0, long count, int overwrite);
void nbb200_host_copy(double *dst, const double *src, long m);
void nbb200_host_add(double *dst, const double *src, long m);

/* ---- velocity Verlet on the device -----------------------------------------------------
 * One step of pCore-1.9.0/pCore/VelocityVerletIntegrator.py:60-81 (Iteration) in Cartesian variables, for callers that keep
 * coordinates, velocities, accelerations and gradients resident on the device between NB calls (device arrays, 3 n doubles; units
 * A, A/ps, A/ps^2, kJ/mol/A, amu, ps as in pMolecule/SystemGeometryObjectiveFunction.py:18-28):
 *   first half : x += dt v + dt^2/2 a ; v += dt/2 a
 *   (energy + gradient call)
 *   second half: a = -100 g / m ; v += dt/2 a ; *d_ke = kinetic energy in kJ/mol */
void nbb200_vv_first_half(NBB200State *state, double *d_x, double *d_v, const double *d_a, double dt);
void nbb200_vv_second_half(NBB200State *state, double *d_v, double *d_a, const double *d_g, const double *d_mass, double dt, double *d_ke);

/* Langevin velocity Verlet (pCore-1.9.0/pCore/LangevinVelocityVerletIntegrator.py:117-150, the integrator of the reference's own DHFR benchmark,
 * benchmarks/SystemBenchmarks.py:95-101), first part of an Iteration in Cartesian variables:
 *   x += facR1 v + facR2 a + sdR w1 / sqrt(m) ; v = facV1 v + facV2 a + (sdV1 w1 + sdV2 w2) / sqrt(m)
 * factors7 = {facR1, facR2, facV1, facV2, sdR, sdV1, sdV2} as CalculateIntegrationConstants (:54-115) gives them; w1, w2: standard normal
 * deviates from a counter-based generator keyed by (seed, step, coordinate).  The second part (a = -100 g / m; v += facV3 a; kinetic
 * energy) is nbb200_vv_second_half with dt = 2 facV3.  The projection of the deviates on the linear constraints (ApplyLinearConstraints,
 * :143,148) is applied when nbb200_set_langevin_constraints has switched it on. */
void nbb200_langevin_first_half(NBB200State *state, double *d_x, double *d_v, const double *d_a, const double *d_mass, const double *factors7,
                                unsigned long long seed, unsigned long long step);
/* ApplyLinearConstraints of the Langevin integrator for the constraint set RemoveRotationTranslation builds for a periodic system
 * (pMolecule-1.9.0/pMolecule/SystemGeometryObjectiveFunction.py:213-240: translation only; MolecularDynamics.py:27-32 switches it on by default):
 * both random vectors of a step are projected on the complement of the three mass-weighted translation vectors, i.e. the random terms carry no
 * net momentum.  totalMass = the sum of the masses passed as d_mass.  Applies to nbb200_langevin_first_half and nbb200_md_run.  (Systems with
 * fixed atoms have no such constraints in the reference, :220; rotations are only removed for non-periodic systems: not built.) */
void nbb200_set_langevin_constraints(NBB200State *state, int removeTranslation, double totalMass);

/* forward declaration (bonded terms, below) */
typedef struct NBB200MMTerms NBB200MMTerms;
/* nsteps MD steps in one call (the loop of VelocityVerletIntegrator / LangevinVelocityVerletIntegrator around Accelerations(), everything on
 * the device): per step first half (velocity Verlet with timeStep, or Langevin when langevinFactors7 != NULL: seed, iteration = firstIteration + k),
 * list-update decision (the reference's heuristic; updateFrequency > 0 forces a rebuild every that many steps), NB energy + gradients, bonded
 * terms (terms nullable), second half with secondHalfDt (= timeStep, or 2 facV3 for Langevin).  d_x, d_v, d_a, d_g, d_mass, d_ke: device arrays
 * (3n, 3n, 3n, 3n, n, 1 doubles).  potential[nsteps], kinetic[nsteps] (nullable): per-step energies in kJ/mol; nbEnergies6 / bondedEnergies5
 * (nullable): the terms of the last step.  One host wait per step.  Returns the number of list updates. */
int nbb200_md_run(NBB200State *state, NBB200MMTerms *terms, int nsteps, int updateFrequency, double *d_x, double *d_v, double *d_a, double *d_g,
                  const double *d_mass, const double *box6, double timeStep, const double *langevinFactors7, double secondHalfDt,
                  unsigned long long seed, unsigned long long firstIteration, double *d_ke, double *potential, double *kinetic,
                  double *nbEnergies6, double *bondedEnergies5, int *status);

/* ---- bonded MM terms on the device --------------------------------------------------
 * What System.Energy evaluates next to the NB model (pMolecule-1.9.0/pMolecule/System.py:272-318), for callers that keep coordinates and
 * gradients on the device.  One opaque object holds the terms of all containers; they are evaluated by a single launch in fp64.
 * Terms and parameters are given as the reference's containers hold them: per term the atom indices, a parameter type and QACTIVE
 * (active == NULL: all active); per parameter type the values.  Defining a container again replaces it; nterms = 0 removes it. */
NBB200MMTerms *MMTerms_B200_Allocate(int device, int natoms, int *status);
void MMTerms_B200_Deallocate(NBB200MMTerms **terms);
void MMTerms_B200_SetStream(NBB200MMTerms *terms, void *cudaStream);
/* HarmonicBondContainer (pM/cinclude/HarmonicBondContainer.h:17-36; E = fc (r - eq)^2): atoms[2 nterms]; isUreyBradley selects the
 * second container of this kind System holds (label "Urey-Bradley", pBabel CHARMMPSFFileReader.ToHarmonicUreyBradleyContainer) */
void HarmonicBondContainer_B200_Define(NBB200MMTerms *terms, int isUreyBradley, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                       int nparameters, const double *eq, const double *fc, int *status);
/* HarmonicAngleContainer (pM/cinclude/HarmonicAngleContainer.h:17-36; E = fc (theta - eq)^2, radians): atoms[3 nterms] */
void HarmonicAngleContainer_B200_Define(NBB200MMTerms *terms, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                        int nparameters, const double *eq, const double *fc, int *status);
/* FourierDihedralContainer (pM/cinclude/FourierDihedralContainer.h:14-47; E = fc (1 + cos(period phi - phase))): atoms[4 nterms] */
void FourierDihedralContainer_B200_Define(NBB200MMTerms *terms, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                          int nparameters, const double *fc, const int *period, const double *phase, int *status);
/* HarmonicImproperContainer (pM/cinclude/HarmonicImproperContainer.h:14-46; E = fc (phi - eq)^2): atoms[4 nterms] */
void HarmonicImproperContainer_B200_Define(NBB200MMTerms *terms, int nterms, const int *atoms, const int *types, const unsigned char *active,
                                           int nparameters, const double *eq, const double *fc, int *status);
/* replace Harmonic{Bond,Angle,Improper}Container_Energy / FourierDihedralContainer_Energy (pM/csource/HarmonicBondContainer.c:149,
 * HarmonicAngleContainer.c:157, FourierDihedralContainer.c:156, HarmonicImproperContainer.c:172): energies5 = {bond, angle, Urey-Bradley,
 * dihedral, improper} are set; grad[3 natoms] (nullable) is ACCUMULATED into.  Host arrays / device arrays. */
void MMTerms_B200_Energy(NBB200MMTerms *terms, const double *xyz, double *energies5, double *grad, int *status);
void MMTerms_B200_EnergyDevice(NBB200MMTerms *terms, const double *d_xyz, double *energies5, double *d_grad, int *status);
/* the device call in two halves, for callers that overlap it with other work on the same stream (e.g. enqueue before the NB call, collect
 * after it: one host synchronisation less per MD step): Enqueue launches the kernel, Collect waits and returns the energies */
void MMTerms_B200_EnergyDeviceEnqueue(NBB200MMTerms *terms, const double *d_xyz, double *d_grad, int *status);
void MMTerms_B200_EnergyDeviceCollect(NBB200MMTerms *terms, double *energies5, int *status);
/* the energies of the last completed Enqueue WITHOUT a synchronisation: for callers that know the stream has been synchronised since */
void MMTerms_B200_LastEnergies(NBB200MMTerms *terms, double *energies5);
long MMTerms_B200_NumberOfTerms(NBB200MMTerms *terms, int kind /* 0 bond, 1 angle, 2 Urey-Bradley, 3 dihedral, 4 improper: active terms */);

/* ---- several GPUs ------------------------------------------------------------
 * Every rank sorts all atoms the same way (cell order); rank r owns the contiguous slab of sorted positions
 * [s0, s1) = the i-blocks nbb200_set_partition gave it, i.e. a spatial slab.  Its lists reference, inside every other
 * rank's slab, a contiguous range of sorted positions (the halo).  Per call the ranks exchange exactly these ranges:
 * positions of halo atoms to the ranks that list them, gradient contributions back to the owners
 * (pdynamo-mirror_b200/parallel.py drives this with NCCL send/recv); energies and dE/dM are all-reduced.
 * There is no counterpart in the reference (its only parallelism is an OpenMP team, NBModelABFSState.c:401). */
void nbb200_get_slab(NBB200State *state, long *out4);          /* s0, s1, n, number of i-blocks */
/* after a rebuild: out[4 r + 2 h], out[4 r + 2 h + 1] = [lo, hi) sorted positions this rank's lists reference in the lower (h = 0) /
 * upper (h = 1) half of rank r's slab (0, 0: none).  Two ranges per slab: periodic images reach both ends of a neighbour's slab. */
int  nbb200_touched_ranges(NBB200State *state, long *out);
/* the same table written to a DEVICE array d_out[4 nranks] on the state's stream, without host synchronisation (the exchange of the
 * tables between the ranks then overlaps the energy kernels) */
int  nbb200_touched_ranges_device(NBB200State *state, long *d_out);
/* sorted-order gradient accumulator owned by the caller (device, 3 n doubles): zeroed and filled by ...MMMMEnergySorted */
void nbb200_set_sorted_gradient_buffer(NBB200State *state, double *d_buf);
/* max_i |x_i - xref_i|^2 of CheckForUpdate (pM/csource/NBModelABFS.c:691-746) for a collective update decision */
double nbb200_max_displacement(NBB200State *state, const double *d_xyz, int *status);
/* NBModelABFS_B200_UpdateDevice with the displacement decision taken by the caller (all ranks must rebuild together) */
int  NBModelABFS_B200_UpdateDeviceDecided(NBB200State *state, const double *d_xyz, const double *box6, int doUpdate, int *status);
/* NBModelABFS_B200_MMMMEnergyDevice that leaves the gradient in SORTED order in the sorted-gradient buffer (partial: own i atoms and
 * the j atoms of the own lists, 1-4 pairs whose first atom is owned) */
void NBModelABFS_B200_MMMMEnergySorted(NBB200State *state, double *energies, double *dEdM, int *status);
/* ... in two halves: Enqueue launches the kernels, Finish waits and hands the energies out; nbb200_peer_push_gradients may go in between */
void NBModelABFS_B200_MMMMEnergySortedEnqueue(NBB200State *state, int *status);
void NBModelABFS_B200_MMMMEnergySortedFinish(NBB200State *state, double *energies, double *dEdM, int *status);
/* out[k] = x[atom(s0 + k)], x[atom(s0 + k)] = in[k], grad[atom(s)] += sortedGradient[s] for k < count; device arrays, 3 doubles per atom */
void nbb200_gather_sorted(NBB200State *state, const double *d_x, long s0, long count, double *d_out);
void nbb200_scatter_sorted(NBB200State *state, const double *d_in, long s0, long count, double *d_x);
void nbb200_unsort_add(NBB200State *state, long s0, long count, double *d_grad);

/* Peer memory (CUDA IPC; one process per GPU of one NVLink / NVSwitch node): the halo exchanges AND the synchronisation between the ranks
 * as plain kernels that read / write the other ranks' buffers -- no library collective in a call.  Every rank exports its sorted-order
 * gradient accumulator, its sorted positions and a small signal area (3 x 64-byte handles), imports everybody else's, and then per
 * call (`step` counts the calls, the same on all ranks):
 *   nbb200_peer_begin           zero the own accumulator, publish the positions of the own slab
 *   nbb200_peer_signal_begin    write (displacement maximum, step) into every rank's signal area
 *   nbb200_peer_wait_begin      wait (bounded spin on the own signal area) for all ranks; returns the global displacement maximum
 *   nbb200_peer_pull_positions  x[atom(s)] = positions of rank r for the own halo ranges (or whole slabs: rebuild)
 *   ...UpdateDeviceDecided, nbb200_touched_ranges_device (when rebuilt), ...MMMMEnergySorted
 *   nbb200_peer_push_gradients  atomically add the own halo contributions into their owners' accumulators
 *   nbb200_peer_signal_end      write (15 scalars, step) into every rank's signal area
 *   nbb200_peer_wait_end        wait for all ranks (their pushes into the own accumulator are complete); sum of the scalars (read_sums)
 *   nbb200_unsort_add           own slab -> atom order
 * d_row: the rank's own range table, device array [nranks][2][2] (long) as written by nbb200_touched_ranges_device. */
int  nbb200_peer_export(NBB200State *state, char *handles192);
int  nbb200_peer_import(NBB200State *state, int rank, const char *handles192);
int  nbb200_peer_attach_local(NBB200State *state, int rank, NBB200State *other);   /* same process, same device: rank `rank` is `other` */
void nbb200_peer_begin(NBB200State *state, const double *d_x, long s0, long count);
void nbb200_peer_signal_begin(NBB200State *state, long step, const double *d_x, int forceRebuild);
double nbb200_peer_wait_begin(NBB200State *state, long step, int needValue, int *status);   /* needValue = 0: ordering only, no host wait */
void nbb200_peer_pull_positions(NBB200State *state, const long *d_row, const long *slabEdges, int wholeSlabs, double *d_x);
void nbb200_peer_push_gradients(NBB200State *state, const long *d_row);
void nbb200_peer_signal_end(NBB200State *state, long step, const double *scal15);
/* the same with the scalars computed from the accumulators of the enqueued energy call by a kernel (no host wait before the signal; the per-rank
 * energies stay on the device, nbb200_peer_read_sums hands out the sums) */
void nbb200_peer_signal_end_device(NBB200State *state, long step, int *status);
void nbb200_peer_wait_end(NBB200State *state, long step);                                   /* on the stream; no host wait */
void nbb200_peer_read_sums(NBB200State *state, double *sum15, int *status);                 /* synchronises; the sums of the last wait_end */

#ifdef __cplusplus
}
#endif
#endif
