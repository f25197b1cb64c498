/* compat_shim.c -- the reference's own C interface of this path on top of libnbabfs_b200.so.
 *
 * Exports, with the reference's names, signatures and struct types (pM = pMolecule-1.9.0/extensions),
 *   NBModelABFS_Allocate / _Clone / _Deallocate / _Update / _MMMMEnergy / _QCMMEnergyLJ / _QCMMPotentials / _QCMMGradients   pM/cinclude/NBModelABFS.h:39-46
 *   NBModelABFSState_Allocate / _Deallocate / _SetUp / _SetUpCentering / _Initialize / _InitializeCoordinates3 /
 *   _GridInitialize / _GridFinalize / _StatisticsAccumulate / _StatisticsInitialize                                   pM/cinclude/NBModelABFSState.h:132-161
 * i.e. everything the translation units pM/csource/NBModelABFS.c and NBModelABFSState.c export.  A pDynamo build that compiles this file
 * INSTEAD of those two (and links libnbabfs_b200.so) keeps its Cython layer (pM/pyrex/pMolecule.NBModelABFS.pyx, pMolecule.NBModelABFSState.pyx)
 * byte for byte: the containers it passes in (MMAtomContainer, LJParameterContainer, PairList, Selection, Coordinates3, SymmetryParameters,
 * Transformation3Container, QCAtomContainer) are unpacked into the flat C-ABI of include/nbabfs_b200.h, and the NBModelABFSState struct it
 * reads back (energies, list pointers for the non-NULL tests of GetEnergies, counts for the summaries, numberOfCalls / numberOfUpdates) is
 * kept current.
 *
 * It needs the reference's headers, so it is compiled only where the reference tree is present: oracle/Makefile builds
 * oracle/_ref/libshim_nbabfs.so = oracle/ref_driver.c + this file + the reference objects WITHOUT NBModelABFS.o / NBModelABFSState.o,
 * and tests/test_parity_gpu.py drives the reference driver's call sequence through it on the GPU.
 *
 * The pair lists themselves live on the device.  The PairList / ImageList objects hung into the state carry the right counts; their
 * entries are materialised (PairList_FromIntegerPairArray) only when the lists are small (NBB200_COMPAT_MATERIALIZE, default: up to
 * 4 M pairs), which is what inspection code and the parity tests need.
 */
#include <stdlib.h>
#include <string.h>

#include "NBModelABFS.h"
#include "NBModelABFSState.h"
#include "Memory.h"
#include "IndexedSelection.h"
#include "List.h"
#include "nbabfs_b200.h"

#define DEFAULT_CHECKFORINVERSES     True
#define DEFAULT_DAMPINGCUTOFF        0.5e+00
#define DEFAULT_DIELECTRIC           1.0e+00
#define DEFAULT_ELECTROSTATICSCALE14 1.0e+00
#define DEFAULT_IMAGEEXPANDFACTOR    0
#define DEFAULT_INNERCUTOFF          8.0e+00
#define DEFAULT_LISTCUTOFF           13.5e+00
#define DEFAULT_OUTERCUTOFF          12.0e+00
#define DEFAULT_QCMMCOUPLING         QCMMLinkAtomCoupling_RC
#define DEFAULT_USECENTERING         False

/* ---------------------------------------------------------------------------------------------------------------------------------
 * side table: reference state -> device state
 * -------------------------------------------------------------------------------------------------------------------------------*/
typedef struct ShimLink {
    NBModelABFSState *state;
    NBB200State      *handle;
    int               n, nqc, formAnalytic, formDensity;
    double           *x, *g;             /* contiguous staging */
    struct ShimLink  *next;
} ShimLink;

static ShimLink *links = NULL;

static ShimLink *find_link(const NBModelABFSState *st)
{
    ShimLink *l;
    for (l = links; l != NULL; l = l->next) if (l->state == st) return l;
    return NULL;
}

static int shim_device(void)
{
    const char *e = getenv("NBB200_DEVICE");
    return (e != NULL) ? atoi(e) : 0;
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * NBModelABFS (options only)
 * -------------------------------------------------------------------------------------------------------------------------------*/
NBModelABFS *NBModelABFS_Allocate(void)
{
    NBModelABFS *self = (NBModelABFS *) Memory_Allocate(sizeof(NBModelABFS));
    if (self != NULL) {
        self->checkForInverses     = DEFAULT_CHECKFORINVERSES;
        self->dampingCutoff        = DEFAULT_DAMPINGCUTOFF;
        self->dielectric           = DEFAULT_DIELECTRIC;
        self->electrostaticScale14 = DEFAULT_ELECTROSTATICSCALE14;
        self->imageExpandFactor    = DEFAULT_IMAGEEXPANDFACTOR;
        self->innerCutoff          = DEFAULT_INNERCUTOFF;
        self->listCutoff           = DEFAULT_LISTCUTOFF;
        self->outerCutoff          = DEFAULT_OUTERCUTOFF;
        self->qcmmCoupling         = DEFAULT_QCMMCOUPLING;
        self->useCentering         = DEFAULT_USECENTERING;
    }
    return self;
}

NBModelABFS *NBModelABFS_Clone(const NBModelABFS *self)
{
    NBModelABFS *other = NULL;
    if (self != NULL) {
        other = NBModelABFS_Allocate();
        if (other != NULL) *other = *self;
    }
    return other;
}

void NBModelABFS_Deallocate(NBModelABFS **self)
{
    if ((self != NULL) && (*self != NULL)) { free(*self); *self = NULL; }
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * NBModelABFSState
 * -------------------------------------------------------------------------------------------------------------------------------*/
NBModelABFSState *NBModelABFSState_Allocate(const Integer n)
{
    NBModelABFSState *self = (NBModelABFSState *) Memory_Allocate(sizeof(NBModelABFSState));
    if (self != NULL) {
        memset(self, 0, sizeof(NBModelABFSState));
        self->isNew        = True;
        self->useCentering = False;
        self->qcmmCoupling = DEFAULT_QCMMCOUPLING;
        (void) n;
    }
    return self;
}

static void drop_lists(NBModelABFSState *self)
{
    ImageList_Deallocate(&(self->inbmmmm));
    ImageList_Deallocate(&(self->inbqcmmlj)); ImageList_Deallocate(&(self->inbqcmmel));
    ImageList_Deallocate(&(self->inbqcqclj)); ImageList_Deallocate(&(self->inbqcqcel));
    PairList_Deallocate(&(self->nbmmmm));
    PairList_Deallocate(&(self->nbmmmm14));
    PairList_Deallocate(&(self->nbqcmmlj)); PairList_Deallocate(&(self->nbqcmmel));
}

void NBModelABFSState_Deallocate(NBModelABFSState **self)
{
    if ((self != NULL) && (*self != NULL)) {
        ShimLink *l, **p;
        for (p = &links; (l = *p) != NULL; p = &(l->next)) {
            if (l->state == *self) {
                *p = l->next;
                NBModelABFSState_B200_Deallocate(&(l->handle));
                free(l->x); free(l->g); free(l);
                break;
            }
        }
        drop_lists(*self);
        Selection_Deallocate(&((*self)->freeSelection));
        Selection_Deallocate(&((*self)->mmSelection));
        Selection_Deallocate(&((*self)->qcpSelection));
        free(*self);
        *self = NULL;
    }
}

/* the grid of the reference's generator has no counterpart: the device builder always bins */
void NBModelABFSState_GridFinalize(NBModelABFSState *self) { (void) self; }
void NBModelABFSState_GridInitialize(NBModelABFSState *self, const PairListGenerator *generator, Status *status) { (void) self; (void) generator; (void) status; }
/* centring is done on the device inside Update (NBModelABFSState_B200_SetUpCentering) */
void NBModelABFSState_InitializeCoordinates3(NBModelABFSState *self, const Boolean doUpdate) { (void) self; (void) doUpdate; }

void NBModelABFSState_Initialize(NBModelABFSState *self, Coordinates3 *coordinates3, SymmetryParameters *symmetryParameters, Coordinates3 *gradients3,
                                 SymmetryParameterGradients *symmetryParameterGradients)
{
    if (self != NULL) {                                       /* pM/csource/NBModelABFSState.c:231-273 */
        self->numberOfCalls += 1;
        self->eimmmel = self->eimmmlj = self->eimqcmmlj = self->eimqcqclj = 0.0e+00;
        self->emmel = self->emmel14 = self->emmlj = self->emmlj14 = self->eqcmmlj = self->eqcmmlj14 = 0.0e+00;
        self->gradients3                 = gradients3;
        self->inputCoordinates3          = coordinates3;
        self->coordinates3               = coordinates3;
        self->symmetryParameters         = symmetryParameters;
        self->symmetryParameterGradients = symmetryParameterGradients;
        Real1DArray_Set(self->qcCharges, 0.0e+00);
        Real1DArray_Set(self->qcmmPotentials, 0.0e+00);
        SymmetricMatrix_Set(self->qcqcPotentials, 0.0e+00);
    }
}

NBModelABFSState *NBModelABFSState_SetUp(MMAtomContainer *mmAtoms, QCAtomContainer *qcAtoms, Selection *fixedAtoms, PairList *exclusions, PairList *interactions14,
                                         LJParameterContainer *ljParameters, LJParameterContainer *ljParameters14, Real1DArray *qcCharges, Real1DArray *qcmmPotentials,
                                         SymmetricMatrix *qcqcPotentials, Transformation3Container *transformations, const QCMMLinkAtomCoupling qcmmCoupling)
{
    NBModelABFSState *self = NULL;
    if ((mmAtoms != NULL) && (ljParameters != NULL)) {
        const int n = mmAtoms->natoms;
        int i, t, r, c, nexcl = 0, n14 = 0, ntrans = 0, status = NBB200_STATUS_CONTINUE;
        double *q = NULL, *rot = NULL, *trans = NULL;
        int *lt = NULL, *ex = NULL, *p14 = NULL;
        LJParameterContainer *lj14 = (ljParameters14 != NULL) ? ljParameters14 : ljParameters;
        ShimLink *link = NULL;
        if ((qcAtoms != NULL) && (qcAtoms->nboundary > 0)) return NULL;      /* QC regions with boundary atoms: not on the device */
        self = NBModelABFSState_Allocate(n);
        if (self == NULL) return NULL;
        q = (double *) malloc(sizeof(double) * (size_t) (n > 0 ? n : 1)); lt = (int *) malloc(sizeof(int) * (size_t) (n > 0 ? n : 1));
        for (i = 0; i < n; i++) { q[i] = mmAtoms->data[i].QACTIVE ? mmAtoms->data[i].charge : 0.0e+00; lt[i] = mmAtoms->data[i].ljtype; }
        if (exclusions != NULL)     { nexcl = PairList_Length(exclusions);     if (nexcl > 0) ex  = PairList_ToIntegerPairArray(exclusions); }
        if (interactions14 != NULL) { n14   = PairList_Length(interactions14); if (n14 > 0)   p14 = PairList_ToIntegerPairArray(interactions14); }
        if (transformations != NULL) {
            ntrans = transformations->nitems;
            rot = (double *) malloc(sizeof(double) * 9 * (size_t) (ntrans > 0 ? ntrans : 1)); trans = (double *) malloc(sizeof(double) * 3 * (size_t) (ntrans > 0 ? ntrans : 1));
            for (t = 0; t < ntrans; t++) {
                for (r = 0; r < 3; r++) {
                    for (c = 0; c < 3; c++) rot[9 * t + 3 * r + c] = Matrix33_Item(transformations->items[t]->rotation, r, c);
                    trans[3 * t + r] = Vector3_Item(transformations->items[t]->translation, r);
                }
            }
        }
        link = (ShimLink *) calloc(1, sizeof(ShimLink));
        link->handle = NBModelABFSState_B200_SetUp(shim_device(), n, q, lt, ljParameters->ntypes, ljParameters->tableindex, ljParameters->tableA, ljParameters->tableB,
                                                   lj14->ntypes, lj14->tableindex, lj14->tableA, lj14->tableB, nexcl, ex, n14, p14, ntrans, rot, trans, &status);
        if ((link->handle != NULL) && (status == NBB200_STATUS_CONTINUE) && (fixedAtoms != NULL) && (fixedAtoms->nindices > 0)) {
            NBModelABFSState_B200_SetFixedAtoms(link->handle, fixedAtoms->nindices, fixedAtoms->indices, &status);
            self->freeSelection = Selection_Complement(fixedAtoms, n);
        }
        if ((link->handle != NULL) && (status == NBB200_STATUS_CONTINUE) && (qcAtoms != NULL)) {
            self->qcpSelection = QCAtomContainer_MakePureSelection(qcAtoms);
            if (self->qcpSelection != NULL) {
                Selection_Sort(self->qcpSelection);
                NBModelABFSState_B200_SetQCAtoms(link->handle, self->qcpSelection->nindices, self->qcpSelection->indices, &status);
                self->mmSelection = Selection_Complement(self->qcpSelection, n);
                link->nqc = self->qcpSelection->nindices;
            }
        }
        free(q); free(lt); free(rot); free(trans);
        free(ex); free(p14);
        if ((link->handle == NULL) || (status != NBB200_STATUS_CONTINUE)) {
            if (link->handle != NULL) NBModelABFSState_B200_Deallocate(&(link->handle));
            free(link);
            free(self);
            return NULL;
        }
        link->state = self; link->n = n; link->formAnalytic = 1; link->formDensity = 0;
        link->x = (double *) malloc(sizeof(double) * 3 * (size_t) (n > 0 ? n : 1)); link->g = (double *) malloc(sizeof(double) * 3 * (size_t) (n > 0 ? n : 1));
        link->next = links; links = link;
        /* aliases (pM/csource/NBModelABFSState.c:384-400) */
        self->exclusions = exclusions; self->fixedAtoms = fixedAtoms; self->interactions14 = interactions14;
        self->ljParameters = ljParameters; self->ljParameters14 = ljParameters14; self->mmAtoms = mmAtoms; self->qcAtoms = qcAtoms;
        self->qcCharges = qcCharges; self->qcmmPotentials = qcmmPotentials; self->qcqcPotentials = qcqcPotentials;
        self->transformations = transformations; self->qcmmCoupling = qcmmCoupling;
    }
    return self;
}

void NBModelABFSState_SetUpCentering(NBModelABFSState *self, const Boolean useCentering, Status *status)
{
    ShimLink *l = find_link(self);
    if (l != NULL) {
        int st = NBB200_STATUS_CONTINUE;
        NBModelABFSState_B200_SetUpCentering(l->handle, useCentering ? 1 : 0, &st);
        self->useCentering = useCentering;
        if ((st != NBB200_STATUS_CONTINUE) && (status != NULL)) *status = Status_OutOfMemory;
    }
}

void NBModelABFSState_StatisticsInitialize(NBModelABFSState *self)
{
    if (self != NULL) {
        self->numberOfCalls = 0; self->numberOfUpdates = 0;
        self->numberOfMMMMPairs = self->numberOfMMMMImageImages = self->numberOfMMMMImagePairs = 0.0e+00;
    }
}

void NBModelABFSState_StatisticsAccumulate(NBModelABFSState *self)
{
    if (self != NULL) {                                       /* pM/csource/NBModelABFSState.c: running sums over the updates */
        self->numberOfUpdates += 1;
        if (self->nbmmmm  != NULL) self->numberOfMMMMPairs += (Real) self->nbmmmm->npairs;
        if (self->inbmmmm != NULL) {
            self->numberOfMMMMImageImages += (Real) ImageList_NumberOfImages(self->inbmmmm);
            self->numberOfMMMMImagePairs  += (Real) ImageList_NumberOfPairs(self->inbmmmm);
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * the lists as the reference's inspection code sees them
 * -------------------------------------------------------------------------------------------------------------------------------*/
static int pair_compare(const void *a, const void *b)
{
    const int *p = (const int *) a, *q = (const int *) b;
    if (p[0] != q[0]) return (p[0] < q[0]) ? -1 : 1;
    if (p[1] != q[1]) return (p[1] < q[1]) ? -1 : 1;
    return 0;
}

/* a PairList in the reference's basic representation (one IndexedSelection per first index, pCore PairList.h:25-36).  Unlike
 * PairList_FromIntegerPairArray nothing is dropped: an image list may pair an atom with its own image (i == j). */
static PairList *make_pairlist(ShimLink *l, int image, long count, Boolean qself, int materialize)
{
    PairList *pl = PairList_Allocate(qself);
    if (pl == NULL) return NULL;
    pl->npairs = (Integer) count;                            /* counts only, unless the entries are materialised below */
    if (materialize && (count > 0)) {
        int st = NBB200_STATUS_CONTINUE;
        int *pairs = (int *) malloc(sizeof(int) * 2 * (size_t) count), *js = (int *) malloc(sizeof(int) * (size_t) count);
        if ((pairs != NULL) && (js != NULL) && (NBModelABFSState_B200_GetPairs(l->handle, image, pairs, &st) == count) && (st == NBB200_STATUS_CONTINUE)) {
            long p, first = 0;
            int maxi = 0, maxj = 0;
            if (qself) for (p = 0; p < count; p++) if (pairs[2 * p] < pairs[2 * p + 1]) { const int t = pairs[2 * p]; pairs[2 * p] = pairs[2 * p + 1]; pairs[2 * p + 1] = t; }
            qsort(pairs, (size_t) count, 2 * sizeof(int), pair_compare);
            for (p = 0; p <= count; p++) {
                if ((p == count) || (pairs[2 * p] != pairs[2 * first])) {
                    const long m = p - first;
                    long k;
                    IndexedSelection *sel;
                    for (k = 0; k < m; k++) js[k] = pairs[2 * (first + k) + 1];
                    sel = IndexedSelection_Allocate(pairs[2 * first], (int) m, js);
                    if (sel != NULL) List_Element_Append(pl->pairs, (void *) sel);
                    first = p;
                }
                if (p < count) { if (pairs[2 * p] > maxi) maxi = pairs[2 * p]; if (pairs[2 * p + 1] > maxj) maxj = pairs[2 * p + 1]; }
            }
            pl->upperBoundi = maxi + 1; pl->upperBoundj = maxj + 1;
        }
        free(pairs); free(js);
    }
    return pl;
}

static void refresh_lists(ShimLink *l)
{
    NBModelABFSState *self = l->state;
    const char *e = getenv("NBB200_COMPAT_MATERIALIZE");
    const long limit = (e != NULL) ? atol(e) : 4000000L;
    const long np = NBModelABFSState_B200_NumberOfPairs(l->handle, -1), n14 = NBModelABFSState_B200_NumberOf14Pairs(l->handle);
    const long nip = NBModelABFSState_B200_NumberOfImagePairs(l->handle);
    const int nimages = NBModelABFSState_B200_NumberOfImages(l->handle), materialize = (np + nip) <= limit;
    int k;
    drop_lists(self);
    self->nbmmmm = make_pairlist(l, -1, np, True, materialize);
    if ((self->interactions14 != NULL) && (n14 > 0)) { self->nbmmmm14 = PairList_Allocate(True); if (self->nbmmmm14 != NULL) self->nbmmmm14->npairs = (Integer) n14; }
    if ((self->transformations != NULL) && (nimages > 0)) {
        self->inbmmmm = ImageList_Allocate();
        for (k = 0; k < nimages; k++) {
            int info[6]; double scale = 1.0e+00;
            PairList *pl;
            NBModelABFSState_B200_GetImageInfo(l->handle, k, info, &scale);
            pl = make_pairlist(l, k, (long) info[4], False, materialize);
            ImageList_CreateImage(self->inbmmmm, info[1], info[2], info[3], scale, self->transformations->items[info[0]], &pl);
        }
    }
    if (l->nqc > 0) {                                         /* non-NULL markers for NBModelABFSState.GetEnergies (pMolecule.NBModelABFSState.pyx:41-59) */
        self->nbqcmmlj = PairList_Allocate(False); self->nbqcmmel = PairList_Allocate(False);
        if (self->transformations != NULL) { self->inbqcmmlj = ImageList_Allocate(); self->inbqcmmel = ImageList_Allocate(); self->inbqcqclj = ImageList_Allocate(); self->inbqcqcel = ImageList_Allocate(); }
    }
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * NBModelABFS_Update (pM/csource/NBModelABFS.c:508-623)
 * -------------------------------------------------------------------------------------------------------------------------------*/
Boolean NBModelABFS_Update(const NBModelABFS *self, const PairListGenerator *generator, NBModelABFSState *nbState, Status *status)
{
    Boolean updated = False;
    ShimLink *l = find_link(nbState);
    (void) generator;                                        /* cell sizes etc. of the CPU generator do not change the lists */
    if ((self != NULL) && (l != NULL) && (nbState->inputCoordinates3 != NULL)) {
        int i, st = NBB200_STATUS_CONTINUE, done;
        double box6[6];
        const Coordinates3 *x = nbState->inputCoordinates3;
        NBModelABFS_B200_SetOptions(l->handle, self->dampingCutoff, self->innerCutoff, self->outerCutoff, self->listCutoff, self->dielectric,
                                    self->electrostaticScale14, self->checkForInverses ? 1 : 0, self->imageExpandFactor);
        for (i = 0; i < l->n; i++) { l->x[3 * i] = Coordinates3_Item(x, i, 0); l->x[3 * i + 1] = Coordinates3_Item(x, i, 1); l->x[3 * i + 2] = Coordinates3_Item(x, i, 2); }
        if (nbState->symmetryParameters != NULL) {
            const SymmetryParameters *sp = nbState->symmetryParameters;
            box6[0] = sp->a; box6[1] = sp->b; box6[2] = sp->c; box6[3] = sp->alpha; box6[4] = sp->beta; box6[5] = sp->gamma;
        }
        done = NBModelABFS_B200_Update(l->handle, l->x, (nbState->symmetryParameters != NULL) ? box6 : NULL, nbState->isNew ? 1 : 0, &st);
        if (st != NBB200_STATUS_CONTINUE) { if (status != NULL) *status = Status_OutOfMemory; return False; }
        nbState->listCutoff = self->listCutoff; nbState->outerCutoff = self->outerCutoff;
        nbState->isNew = False;
        nbState->useGridSearch = True;
        if (done) {
            updated = True;
            refresh_lists(l);
            NBModelABFSState_StatisticsAccumulate(nbState);
        }
    }
    return updated;
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * energies
 * -------------------------------------------------------------------------------------------------------------------------------*/
static void add_gradients(ShimLink *l, Coordinates3 *g3)
{
    int i;
    for (i = 0; i < l->n; i++) {
        Coordinates3_Item(g3, i, 0) += l->g[3 * i]; Coordinates3_Item(g3, i, 1) += l->g[3 * i + 1]; Coordinates3_Item(g3, i, 2) += l->g[3 * i + 2];
    }
}

static void add_dEdM(NBModelABFSState *st, const double *m9)
{
    if ((st->symmetryParameterGradients != NULL) && (st->symmetryParameters != NULL)) {
        int r, c;
        for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) Matrix33_Item(st->symmetryParameterGradients->dEdM, r, c) += m9[3 * r + c];
    }
}

void NBModelABFS_MMMMEnergy(const NBModelABFS *self, const PairwiseInteractionABFS *mmmmPairwiseInteraction, NBModelABFSState *nbState)
{
    ShimLink *l = find_link(nbState);
    if ((self != NULL) && (mmmmPairwiseInteraction != NULL) && (l != NULL)) {
        int st = NBB200_STATUS_CONTINUE;
        double e[6], m9[9];
        const int analytic = mmmmPairwiseInteraction->useAnalyticForm ? 1 : 0, density = mmmmPairwiseInteraction->splinePointDensity;
        const int wantG = nbState->gradients3 != NULL;
        if ((analytic != l->formAnalytic) || (!analytic && (density != l->formDensity))) {
            PairwiseInteractionABFS_B200_SetInteractionForm(l->handle, analytic, density, &st);
            l->formAnalytic = analytic; l->formDensity = density;
        }
        memset(m9, 0, sizeof(m9));
        if (wantG) memset(l->g, 0, sizeof(double) * 3 * (size_t) l->n);
        NBModelABFS_B200_MMMMEnergy(l->handle, e, wantG ? l->g : NULL, wantG ? m9 : NULL, &st);
        if (st != NBB200_STATUS_CONTINUE) return;
        nbState->emmel   = e[NBB200_EMMEL];   nbState->emmlj   = e[NBB200_EMMLJ];
        nbState->emmel14 = e[NBB200_EMMEL14]; nbState->emmlj14 = e[NBB200_EMMLJ14];
        nbState->eimmmel = e[NBB200_EIMMMEL]; nbState->eimmmlj = e[NBB200_EIMMMLJ];
        if (wantG) { add_gradients(l, nbState->gradients3); add_dEdM(nbState, m9); }
    }
}

void NBModelABFS_QCMMEnergyLJ(const NBModelABFS *self, const PairwiseInteractionABFS *mmmmPairwiseInteraction, NBModelABFSState *nbState)
{
    ShimLink *l = find_link(nbState);
    if ((self != NULL) && (mmmmPairwiseInteraction != NULL) && (l != NULL) && (nbState->qcAtoms != NULL)) {
        int st = NBB200_STATUS_CONTINUE;
        double e4[4], m9[9];
        const int wantG = nbState->gradients3 != NULL;
        memset(m9, 0, sizeof(m9));
        if (wantG) memset(l->g, 0, sizeof(double) * 3 * (size_t) l->n);
        NBModelABFS_B200_QCMMEnergyLJ(l->handle, e4, wantG ? l->g : NULL, wantG ? m9 : NULL, &st);
        if (st != NBB200_STATUS_CONTINUE) return;
        nbState->eqcmmlj = e4[0]; nbState->eqcmmlj14 = e4[1]; nbState->eimqcmmlj = e4[2]; nbState->eimqcqclj = e4[3];
        if (wantG) { add_gradients(l, nbState->gradients3); add_dEdM(nbState, m9); }
    }
}

void NBModelABFS_QCMMPotentials(const NBModelABFS *self, const PairwiseInteractionABFS *qcmmPairwiseInteraction, const PairwiseInteractionABFS *qcqcPairwiseInteraction,
                                NBModelABFSState *nbState)
{
    ShimLink *l = find_link(nbState);
    if ((self != NULL) && (qcmmPairwiseInteraction != NULL) && (qcqcPairwiseInteraction != NULL) && (l != NULL) && (nbState->qcAtoms != NULL)) {
        int st = NBB200_STATUS_CONTINUE, k;
        const int nq = l->nqc;
        double *pot = (double *) calloc((size_t) (nq > 0 ? nq : 1), sizeof(double)), *v = (double *) calloc((size_t) (nq * (nq + 1) / 2 + 1), sizeof(double));
        NBModelABFS_B200_QCMMPotentials(l->handle, pot, (nbState->qcqcPotentials != NULL) ? v : NULL, &st);
        if (st == NBB200_STATUS_CONTINUE) {
            if (nbState->qcmmPotentials != NULL) for (k = 0; k < nq; k++) Real1DArray_Item(nbState->qcmmPotentials, k) += pot[k];
            if (nbState->qcqcPotentials != NULL) for (k = 0; k < nq * (nq + 1) / 2; k++) nbState->qcqcPotentials->data[k] += v[k];
        }
        free(pot); free(v);
    }
}

void NBModelABFS_QCMMGradients(const NBModelABFS *self, const PairwiseInteractionABFS *qcmmPairwiseInteraction, const PairwiseInteractionABFS *qcqcPairwiseInteraction,
                               NBModelABFSState *nbState)
{
    ShimLink *l = find_link(nbState);
    if ((self != NULL) && (qcmmPairwiseInteraction != NULL) && (qcqcPairwiseInteraction != NULL) && (l != NULL) && (nbState->qcAtoms != NULL) &&
        (nbState->gradients3 != NULL) && (nbState->qcCharges != NULL)) {
        int st = NBB200_STATUS_CONTINUE, k;
        const int nq = l->nqc;
        double m9[9], *qc = (double *) malloc(sizeof(double) * (size_t) (nq > 0 ? nq : 1));
        for (k = 0; k < nq; k++) qc[k] = Real1DArray_Item(nbState->qcCharges, k);
        memset(m9, 0, sizeof(m9));
        memset(l->g, 0, sizeof(double) * 3 * (size_t) l->n);
        NBModelABFS_B200_QCMMGradients(l->handle, qc, l->g, m9, &st);
        if (st == NBB200_STATUS_CONTINUE) { add_gradients(l, nbState->gradients3); add_dEdM(nbState, m9); }
        free(qc);
    }
}
