/* oracle/ref_driver.c -- TEST INFRASTRUCTURE (never linked into the product library).
 *
 * Drives the UNMODIFIED reference C implementation (headers/sources used where they lie under
 * /root/reference) through the same sequence the Cython layer uses:
 *   NBModelABFS.SetUp  : pMolecule-1.9.0/extensions/pyrex/pMolecule.NBModelABFS.pyx:181-273
 *   NBModelABFS.Energy : pMolecule-1.9.0/extensions/pyrex/pMolecule.NBModelABFS.pyx:108-122
 * Default generator options: pMolecule.NBModelABFS.pyx:59-71 ; cellSize = factor*cutoff as in
 * pCore-1.9.0/extensions/pyrex/pCore.PairListGenerator.pyx:104.
 * Identity container construction: pCore.Transformation3Container.pyx:159-168.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef USEOPENMP
#include <omp.h>
#endif

#include "NBModelABFS.h"
#include "NBModelABFSState.h"
#include "PairwiseInteraction.h"
#include "PairListGenerator.h"
#include "LJParameterContainer.h"
#include "MMAtomContainer.h"
#include "SymmetryParameters.h"
#include "SymmetryParameterGradients.h"
#include "Transformation3Container.h"
#include "ImageList.h"
#include "Selection.h"
#include "CubicSpline.h"
#include "HarmonicBondContainer.h"
#include "HarmonicAngleContainer.h"
#include "FourierDihedralContainer.h"
#include "HarmonicImproperContainer.h"
#include "ref_driver.h"

struct RefNB {
    int n, ntrans;
    MMAtomContainer            *mm;
    LJParameterContainer       *lj, *lj14;
    PairList                   *excl, *i14;
    Transformation3Container   *tc;
    SymmetryParameters         *sp;
    SymmetryParameterGradients *spg;
    NBModelABFS                *nb;
    PairListGenerator          *gen;
    PairwiseInteractionABFS    *pw;
    NBModelABFSState           *st;
    Coordinates3               *x, *g;
    Selection                  *fixed;
    int                         centering;
};

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec;
}

static LJParameterContainer *make_lj(int nt, const int *tableindex, const double *tA, const double *tB)
{
    LJParameterContainer *lj = LJParameterContainer_Allocate(nt);
    int i, m = (nt * (nt + 1)) / 2;
    for (i = 0; i < nt * nt; i++) lj->tableindex[i] = tableindex[i];
    for (i = 0; i < m; i++) { lj->tableA[i] = tA[i]; lj->tableB[i] = tB[i]; }
    return lj;
}

RefNB *refnb_create(int n, const double *charges, const int *ljtypes,
                    int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                    int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                    int nexcl, const int *exclPairs, int n14, const int *pairs14,
                    int ntrans, const double *rot, const double *trans)
{
    int i, t, r, c;
    RefNB *h = (RefNB *) calloc(1, sizeof(RefNB));
    if (h == NULL) return NULL;
    h->n = n; h->ntrans = ntrans;
    h->mm = MMAtomContainer_Allocate(n);
    for (i = 0; i < n; i++) {
        h->mm->data[i].QACTIVE  = True;
        h->mm->data[i].atomtype = ljtypes[i];
        h->mm->data[i].ljtype   = ljtypes[i];
        h->mm->data[i].charge   = charges[i];
    }
    h->lj = make_lj(ntypes, tableindex, tableA, tableB);
    if (tableindex14 != NULL) h->lj14 = make_lj(ntypes14, tableindex14, tableA14, tableB14);
    else                      h->lj14 = make_lj(ntypes, tableindex, tableA, tableB);
    if (nexcl > 0) {
        int *tmp = (int *) malloc(sizeof(int) * 2 * (size_t) nexcl);
        memcpy(tmp, exclPairs, sizeof(int) * 2 * (size_t) nexcl);
        h->excl = PairList_FromIntegerPairArray(True, nexcl, tmp);
        free(tmp);
    }
    if (n14 > 0) {
        int *tmp = (int *) malloc(sizeof(int) * 2 * (size_t) n14);
        memcpy(tmp, pairs14, sizeof(int) * 2 * (size_t) n14);
        h->i14 = PairList_FromIntegerPairArray(True, n14, tmp);
        free(tmp);
    }
    if (ntrans > 0) {
        h->tc = Transformation3Container_Allocate(ntrans);
        h->tc->QOWNER = True;
        for (t = 0; t < ntrans; t++) {
            Transformation3 *T = Transformation3_Allocate();
            for (r = 0; r < 3; r++) {
                for (c = 0; c < 3; c++) Matrix33_Item(T->rotation, r, c) = rot[9 * t + 3 * r + c];
                Vector3_Item(T->translation, r) = trans[3 * t + r];
            }
            h->tc->items[t] = T;
        }
        Transformation3Container_FindIdentity(h->tc);
        Transformation3Container_FindInverses(h->tc);
        h->sp  = SymmetryParameters_Allocate();
        h->spg = SymmetryParameterGradients_Allocate();
    }
    h->nb  = NBModelABFS_Allocate();
    h->gen = PairListGenerator_Allocate();
    h->pw  = PairwiseInteractionABFS_Allocate();
    h->x   = Coordinates3_Allocate(n);
    h->g   = Coordinates3_Allocate(n);
    refnb_set_options(h, 0.5, 8.0, 12.0, 13.5, 1.0, 1.0, 1, 0, 0.5, 0, 1, 0);
    h->st = NBModelABFSState_SetUp(h->mm, NULL, NULL, h->excl, h->i14, h->lj, h->lj14, NULL, NULL, NULL, h->tc, h->nb->qcmmCoupling);
    if (h->st == NULL) { refnb_destroy(h); return NULL; }
    return h;
}

/* fixedAtoms argument of NBModelABFSState_SetUp (pMolecule.NBModelABFS.pyx:181-273 passes system.hardConstraints.fixedAtoms):
 * the state is created anew with the selection, as SetUp does for a new configuration.  nfixed = 0 clears. */
int refnb_set_fixed(RefNB *h, int nfixed, const int *fixed)
{
    int i;
    if (h == NULL) return 0;
    NBModelABFSState_Deallocate(&(h->st));
    Selection_Deallocate(&(h->fixed));
    if (nfixed > 0) {
        h->fixed = Selection_Allocate(nfixed);
        for (i = 0; i < nfixed; i++) h->fixed->indices[i] = fixed[i];
    }
    h->st = NBModelABFSState_SetUp(h->mm, NULL, h->fixed, h->excl, h->i14, h->lj, h->lj14, NULL, NULL, NULL, h->tc, h->nb->qcmmCoupling);
    if (h->st != NULL && h->centering) { Status status = Status_Continue; NBModelABFSState_SetUpCentering(h->st, True, &status); }
    return h->st != NULL;
}

/* NBModelABFS option useCentering: NBModelABFSState_SetUpCentering right after the state is created (pMolecule.NBModelABFS.pyx:245).
 * The state is created anew, as SetUp does for a new configuration. */
int refnb_set_centering(RefNB *h, int on)
{
    if (h == NULL) return 0;
    h->centering = on != 0;
    {
        int nf = (h->fixed != NULL) ? h->fixed->nindices : 0, ok;
        int *idx = (nf > 0) ? (int *) malloc(sizeof(int) * (size_t) nf) : NULL, i;
        for (i = 0; i < nf; i++) idx[i] = h->fixed->indices[i];
        ok = refnb_set_fixed(h, nf, idx);
        free(idx);
        return ok;
    }
}

/* PairwiseInteractionABFS options useAnalyticForm / splinePointDensity (pMolecule.PairwiseInteraction.pyx:226-239) followed by
 * MakeSplines as NBModelABFS.CheckPairwiseInteractions does for the MM/MM interaction (pMolecule.NBModelABFS.pyx:78-82): the three
 * Delta/Delta splines in kJ/mol.  Call after refnb_set_options (the splines depend on the cutoffs). */
void refnb_set_interaction_form(RefNB *h, int useAnalyticForm, int splinePointDensity)
{
    if (h == NULL) return;
    h->pw->useAnalyticForm    = useAnalyticForm ? True : False;
    h->pw->splinePointDensity = splinePointDensity;
    CubicSpline_Deallocate(&(h->pw->electrostaticSpline));
    CubicSpline_Deallocate(&(h->pw->lennardJonesASpline));
    CubicSpline_Deallocate(&(h->pw->lennardJonesBSpline));
    h->pw->electrostaticSpline = PairwiseInteractionABFS_MakeElectrostaticSpline(h->pw, False, NULL);
    h->pw->lennardJonesASpline = PairwiseInteractionABFS_MakeLennardJonesASpline(h->pw, NULL);
    h->pw->lennardJonesBSpline = PairwiseInteractionABFS_MakeLennardJonesBSpline(h->pw, NULL);
}

void refnb_destroy(RefNB *h)
{
    if (h == NULL) return;
    NBModelABFSState_Deallocate(&(h->st));
    Selection_Deallocate(&(h->fixed));
    Coordinates3_Deallocate(&(h->x));
    Coordinates3_Deallocate(&(h->g));
    PairwiseInteractionABFS_Deallocate(&(h->pw));
    PairListGenerator_Deallocate(&(h->gen));
    NBModelABFS_Deallocate(&(h->nb));
    if (h->spg != NULL) SymmetryParameterGradients_Deallocate(&(h->spg));
    if (h->sp  != NULL) SymmetryParameters_Deallocate(&(h->sp));
    if (h->tc  != NULL) Transformation3Container_Deallocate(&(h->tc));
    PairList_Deallocate(&(h->i14));
    PairList_Deallocate(&(h->excl));
    LJParameterContainer_Deallocate(&(h->lj14));
    LJParameterContainer_Deallocate(&(h->lj));
    MMAtomContainer_Deallocate(&(h->mm));
    free(h);
}

void refnb_set_options(RefNB *h, double damp, double inner, double outer, double list,
                       double dielectric, double elecScale14, int checkForInverses, int imageExpandFactor,
                       double cellSizeFactor, int method, int useGridByCell, int sortIndices)
{
    h->nb->dampingCutoff        = damp;
    h->nb->innerCutoff          = inner;
    h->nb->outerCutoff          = outer;
    h->nb->listCutoff           = list;
    h->nb->dielectric           = dielectric;
    h->nb->electrostaticScale14 = elecScale14;
    h->nb->checkForInverses     = checkForInverses ? True : False;
    h->nb->imageExpandFactor    = imageExpandFactor;
    h->pw->dampingCutoff = damp; h->pw->innerCutoff = inner; h->pw->outerCutoff = outer;
    h->gen->cutoff               = list;
    h->gen->cutoffCellSizeFactor = cellSizeFactor;
    h->gen->cellSize             = cellSizeFactor * list;
    h->gen->minimumCellExtent    = 2;
    h->gen->minimumPoints        = 500;
    h->gen->sortIndices          = sortIndices ? True : False;
    h->gen->useGridByCell        = useGridByCell ? True : False;
    if (method == 1) h->gen->minimumPoints = 2000000000;              /* never use the grid */
    if (method == 2) { h->gen->minimumPoints = 0; h->gen->minimumCellExtent = 0; }
}

int refnb_energy(RefNB *h, const double *xyz, const double *box, int forceNew,
                 double *energies, double *grad, double *dEdM, double *timings)
{
    int i, updated;
    Status status = Status_Continue;
    double t0, t1, t2;
    for (i = 0; i < h->n; i++) {
        Coordinates3_Item(h->x, i, 0) = xyz[3 * i];
        Coordinates3_Item(h->x, i, 1) = xyz[3 * i + 1];
        Coordinates3_Item(h->x, i, 2) = xyz[3 * i + 2];
    }
    Coordinates3_Set(h->g, 0.0);
    if (h->ntrans > 0) {
        SymmetryParameters_SetCrystalParameters(h->sp, box[0], box[1], box[2], box[3], box[4], box[5]);
        Matrix33_Set(h->spg->dEdM, 0.0);
    }
    if (forceNew) h->st->isNew = True;
    NBModelABFSState_Initialize(h->st, h->x, h->sp, (grad != NULL) ? h->g : NULL, (grad != NULL) ? h->spg : NULL);
    t0 = now_s();
    updated = (int) NBModelABFS_Update(h->nb, h->gen, h->st, &status);
    t1 = now_s();
    if (status != Status_Continue) return -1;
    NBModelABFS_MMMMEnergy(h->nb, h->pw, h->st);
    t2 = now_s();
    if (timings != NULL) { timings[0] = t1 - t0; timings[1] = t2 - t1; }
    energies[0] = h->st->emmel;   energies[1] = h->st->emmlj;
    energies[2] = h->st->emmel14; energies[3] = h->st->emmlj14;
    energies[4] = h->st->eimmmel; energies[5] = h->st->eimmmlj;
    if (grad != NULL) {
        for (i = 0; i < h->n; i++) {
            grad[3 * i]     += Coordinates3_Item(h->g, i, 0);
            grad[3 * i + 1] += Coordinates3_Item(h->g, i, 1);
            grad[3 * i + 2] += Coordinates3_Item(h->g, i, 2);
        }
        if ((dEdM != NULL) && (h->spg != NULL)) {
            int r, c;
            for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) dEdM[3 * r + c] += Matrix33_Item(h->spg->dEdM, r, c);
        }
    }
    return updated;
}

long refnb_num_primary_pairs(RefNB *h) { return (h->st->nbmmmm == NULL) ? 0 : (long) h->st->nbmmmm->npairs; }
long refnb_num_14_pairs(RefNB *h)      { return (h->st->nbmmmm14 == NULL) ? 0 : (long) h->st->nbmmmm14->npairs; }
int  refnb_num_images(RefNB *h)        { return (h->st->inbmmmm == NULL) ? 0 : ImageList_NumberOfImages(h->st->inbmmmm); }
long refnb_num_image_pairs(RefNB *h)   { return (h->st->inbmmmm == NULL) ? 0 : (long) ImageList_NumberOfPairs(h->st->inbmmmm); }
int  refnb_uses_grid(RefNB *h)         { return (int) h->st->useGridSearch; }
int  refnb_num_threads(void)
{
#ifdef USEOPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void dump_pairs(PairList *pl, int *pairs)
{
    int r, k; long m = 0;
    if (pl == NULL) return;
    PairList_MakeRecords(pl);
    for (r = 0; r < pl->numberOfRecords; r++) {
        IndexedSelection *rec = pl->records[r];
        for (k = 0; k < rec->nindices; k++) { pairs[2 * m] = rec->index; pairs[2 * m + 1] = rec->indices[k]; m++; }
    }
}

void refnb_get_primary_pairs(RefNB *h, int *pairs) { dump_pairs(h->st->nbmmmm, pairs); }

void refnb_get_image_info(RefNB *h, int k, int *info, double *scale)
{
    ImageList *il = h->st->inbmmmm;
    Image *im; int t;
    ImageList_MakeRecords(il);
    im = il->records[k];
    info[0] = -1;
    for (t = 0; t < h->tc->nitems; t++) if (h->tc->items[t] == im->transformation3) info[0] = t;
    info[1] = im->a; info[2] = im->b; info[3] = im->c;
    info[4] = (im->pairlist == NULL) ? 0 : im->pairlist->npairs;
    info[5] = 0;
    scale[0] = im->scale;
}

void refnb_get_image_pairs(RefNB *h, int k, int *pairs)
{
    ImageList *il = h->st->inbmmmm;
    ImageList_MakeRecords(il);
    dump_pairs(il->records[k]->pairlist, pairs);
}

void refnb_make_factors(double damp, double inner, double outer, double *out)
{
    PairwiseInteractionABFS *pw = PairwiseInteractionABFS_Allocate();
    PairwiseInteractionABFSFactors f;
    pw->dampingCutoff = damp; pw->innerCutoff = inner; pw->outerCutoff = outer;
    PairwiseInteractionABFS_MakeFactors(pw, &f);
    out[0] = f.r2Damp; out[1] = f.r2On; out[2] = f.r2Off;
    out[3] = f.a; out[4] = f.b; out[5] = f.c; out[6] = f.d;
    out[7] = f.qShift1; out[8] = f.qShift2; out[9] = f.qF0; out[10] = f.qAlpha;
    out[11] = f.aF6; out[12] = f.aK12; out[13] = f.aShift12; out[14] = f.aF0; out[15] = f.aAlpha;
    out[16] = f.bF3; out[17] = f.bK6; out[18] = f.bShift6; out[19] = f.bF0; out[20] = f.bAlpha;
    PairwiseInteractionABFS_Deallocate(&pw);
}

/* the reference's spline tables: which = 0 electrostatic (kJ/mol; 3: atomic units), 1 LJ-A, 2 LJ-B; x, y, h hold npoints values
 * (PairwiseInteractionABFS_Make*Spline, pM/csource/PairwiseInteraction.c:148-285; CubicSpline_MakeFromReal1DArrays,
 * pC/csource/CubicSpline.c:309-420).  Returns the number of points (call with x = NULL to size the arrays). */
int refnb_make_spline(int which, double damp, double inner, double outer, int density, double *x, double *y, double *hh)
{
    PairwiseInteractionABFS *pw = PairwiseInteractionABFS_Allocate();
    CubicSpline *sp = NULL;
    int n = 0, i;
    pw->dampingCutoff = damp; pw->innerCutoff = inner; pw->outerCutoff = outer; pw->splinePointDensity = density;
    if (which == 0)      sp = PairwiseInteractionABFS_MakeElectrostaticSpline(pw, False, NULL);
    else if (which == 3) sp = PairwiseInteractionABFS_MakeElectrostaticSpline(pw, True, NULL);
    else if (which == 1) sp = PairwiseInteractionABFS_MakeLennardJonesASpline(pw, NULL);
    else                 sp = PairwiseInteractionABFS_MakeLennardJonesBSpline(pw, NULL);
    if (sp != NULL) {
        n = sp->length;
        if (x != NULL) for (i = 0; i < n; i++) { x[i] = Real1DArray_Item(sp->x, i); y[i] = Real1DArray_Item(sp->y, i); hh[i] = Real1DArray_Item(sp->h, i); }
        CubicSpline_Deallocate(&sp);
    }
    PairwiseInteractionABFS_Deallocate(&pw);
    return n;
}

/* CubicSpline_Evaluate (pC/csource/CubicSpline.c) of a spline built from the given (x, y): f and g = df/dx at x0 */
void refnb_spline_evaluate(int n, const double *x, const double *y, double x0, double *f, double *g)
{
    Real1DArray *ax = Real1DArray_Allocate(n, NULL), *ay = Real1DArray_Allocate(n, NULL);
    CubicSpline *sp = NULL;
    int i;
    for (i = 0; i < n; i++) { Real1DArray_Item(ax, i) = x[i]; Real1DArray_Item(ay, i) = y[i]; }
    CubicSpline_MakeFromReal1DArrays(&sp, &ax, &ay, 1, 0.0e+00, 1, 0.0e+00);
    CubicSpline_Evaluate(sp, x0, f, g, NULL);
    CubicSpline_Deallocate(&sp);
}

void refnb_lj_table(int ntypes, const double *eps, const double *sigma, int amber,
                    int *tableindex, double *tableA, double *tableB)
{
    LJParameterContainer *lj = LJParameterContainer_Allocate(ntypes);
    int i, m = (ntypes * (ntypes + 1)) / 2;
    for (i = 0; i < ntypes; i++) { lj->epsilon[i] = eps[i]; lj->sigma[i] = sigma[i]; }
    if (amber) LJParameterContainer_MakeTableAMBER(lj); else LJParameterContainer_MakeTableOPLS(lj);
    for (i = 0; i < ntypes * ntypes; i++) tableindex[i] = lj->tableindex[i];
    for (i = 0; i < m; i++) { tableA[i] = lj->tableA[i]; tableB[i] = lj->tableB[i]; }
    LJParameterContainer_Deallocate(&lj);
}

void refnb_make_M(const double *box, double *M9, double *invM9)
{
    SymmetryParameters *sp = SymmetryParameters_Allocate();
    int r, c;
    SymmetryParameters_SetCrystalParameters(sp, box[0], box[1], box[2], box[3], box[4], box[5]);
    for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) {
        M9[3 * r + c]    = Matrix33_Item(sp->M, r, c);
        invM9[3 * r + c] = Matrix33_Item(sp->inverseM, r, c);
    }
    SymmetryParameters_Deallocate(&sp);
}

/* The reference's bonded MM terms as System.Energy evaluates them (pMolecule-1.9.0/pMolecule/System.py:272-318): containers are filled
 * with one parameter record per term (type = term index) and the unmodified *_Energy routines are called
 * (pMolecule-1.9.0/extensions/csource/HarmonicBondContainer.c:149, HarmonicAngleContainer.c:157, FourierDihedralContainer.c:156,
 * HarmonicImproperContainer.c:172).  energies5 = {bond, angle, Urey-Bradley, dihedral, improper}; grad (nullable) is accumulated into. */
void refmm_energy(int n, const double *xyz,
                  int nbond, const int *bonds, const double *bondEq, const double *bondFc,
                  int nangle, const int *angles, const double *angleEq, const double *angleFc,
                  int nub, const int *ubs, const double *ubEq, const double *ubFc,
                  int ndih, const int *dihedrals, const double *dihFc, const int *dihPeriod, const double *dihPhase,
                  int nimp, const int *impropers, const double *impEq, const double *impFc,
                  double *energies5, double *grad)
{
    Coordinates3 *x = Coordinates3_Allocate(n), *g = (grad != NULL) ? Coordinates3_Allocate(n) : NULL;
    int i, pass;
    for (i = 0; i < n; i++) { Coordinates3_Item(x, i, 0) = xyz[3 * i]; Coordinates3_Item(x, i, 1) = xyz[3 * i + 1]; Coordinates3_Item(x, i, 2) = xyz[3 * i + 2]; }
    if (g != NULL) Coordinates3_Set(g, 0.0);
    for (i = 0; i < 5; i++) energies5[i] = 0.0;
    for (pass = 0; pass < 2; pass++) {
        int nt = pass ? nub : nbond; const int *a = pass ? ubs : bonds; const double *eq = pass ? ubEq : bondEq, *fc = pass ? ubFc : bondFc;
        if (nt > 0) {
            HarmonicBondContainer *c = HarmonicBondContainer_Allocate(nt, nt);
            for (i = 0; i < nt; i++) {
                c->terms[i].QACTIVE = True; c->terms[i].atom1 = a[2 * i]; c->terms[i].atom2 = a[2 * i + 1]; c->terms[i].type = i;
                c->parameters[i].eq = eq[i]; c->parameters[i].fc = fc[i];
            }
            energies5[pass ? 2 : 0] = HarmonicBondContainer_Energy(c, x, g);
            HarmonicBondContainer_Deallocate(&c);
        }
    }
    if (nangle > 0) {
        HarmonicAngleContainer *c = HarmonicAngleContainer_Allocate(nangle, nangle);
        for (i = 0; i < nangle; i++) {
            c->terms[i].QACTIVE = True; c->terms[i].atom1 = angles[3 * i]; c->terms[i].atom2 = angles[3 * i + 1]; c->terms[i].atom3 = angles[3 * i + 2]; c->terms[i].type = i;
            c->parameters[i].eq = angleEq[i]; c->parameters[i].fc = angleFc[i];
        }
        energies5[1] = HarmonicAngleContainer_Energy(c, x, g);
        HarmonicAngleContainer_Deallocate(&c);
    }
    if (ndih > 0) {
        FourierDihedralContainer *c = FourierDihedralContainer_Allocate(ndih, ndih);
        for (i = 0; i < ndih; i++) {
            c->terms[i].QACTIVE = True; c->terms[i].atom1 = dihedrals[4 * i]; c->terms[i].atom2 = dihedrals[4 * i + 1];
            c->terms[i].atom3 = dihedrals[4 * i + 2]; c->terms[i].atom4 = dihedrals[4 * i + 3]; c->terms[i].type = i;
            c->parameters[i].fc = dihFc[i]; c->parameters[i].period = dihPeriod[i]; c->parameters[i].phase = dihPhase[i];
        }
        FourierDihedralContainer_FillCosSinPhases(c);
        energies5[3] = FourierDihedralContainer_Energy(c, x, g);
        FourierDihedralContainer_Deallocate(&c);
    }
    if (nimp > 0) {
        HarmonicImproperContainer *c = HarmonicImproperContainer_Allocate(nimp, nimp);
        for (i = 0; i < nimp; i++) {
            c->terms[i].QACTIVE = True; c->terms[i].atom1 = impropers[4 * i]; c->terms[i].atom2 = impropers[4 * i + 1];
            c->terms[i].atom3 = impropers[4 * i + 2]; c->terms[i].atom4 = impropers[4 * i + 3]; c->terms[i].type = i;
            c->parameters[i].eq = impEq[i]; c->parameters[i].fc = impFc[i];
        }
        HarmonicImproperContainer_FillCosSinValues(c);
        energies5[4] = HarmonicImproperContainer_Energy(c, x, g);
        HarmonicImproperContainer_Deallocate(&c);
    }
    if (g != NULL) {
        for (i = 0; i < n; i++) { grad[3 * i] += Coordinates3_Item(g, i, 0); grad[3 * i + 1] += Coordinates3_Item(g, i, 1); grad[3 * i + 2] += Coordinates3_Item(g, i, 2); }
        Coordinates3_Deallocate(&g);
    }
    Coordinates3_Deallocate(&x);
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * QC/MM entry points of NBModelABFS (second half; NOT yet built on the B200 side -- these drivers exist so that the
 * golden vectors of that row are pinned by the compiled reference before any kernel is written).
 *
 * Call order of the Cython layer with QC atoms present (pMolecule.NBModelABFS.pyx:108-135,181-273): SetUp creates the state with
 * qcAtoms and the three QC/MM work arrays of QCMMInteractionState; Energy calls NBModelABFS_MMMMEnergy then
 * NBModelABFS_QCMMEnergyLJ (NBModelABFS.c:306-378); the QC model calls QCMMPotentials (NBModelABFS.c:449-498) before the SCF and
 * QCMMGradients (NBModelABFS.c:383-444) after it with the QC charges.  The QC region here has no boundary (link) atoms and uses
 * the MM link-atom coupling, so mmCharges are the MM charges with the QC atoms deactivated (EnergyModel.DeactivateQCAtomMMTerms,
 * pMolecule/EnergyModel.py:129-140 -> MMAtomContainer QACTIVE flags) and mmCoordinates3 aliases coordinates3
 * (NBModelABFSState.c:300).  The qcmm / qcqc pairwise interactions carry the point-charge electrostatic spline in atomic units
 * (NBModelABFS.CheckPairwiseInteractions, pMolecule.NBModelABFS.pyx:83-97).
 * ------------------------------------------------------------------------------------------------------------------------------ */
#include "QCAtomContainer.h"
#include "Real1DArray.h"
#include "SymmetricMatrix.h"

struct RefQC {
    RefNB                   *nb;
    int                      nqc;
    QCAtomContainer         *qc;
    Real1DArray             *qcCharges, *qcmmPotentials;
    SymmetricMatrix         *qcqcPotentials;
    PairwiseInteractionABFS *pwqcmm, *pwqcqc;
};

void refqc_destroy(RefQC *q)
{
    if (q == NULL) return;
    refnb_destroy(q->nb);                                   /* the state only aliases the QC arrays */
    PairwiseInteractionABFS_Deallocate(&(q->pwqcmm));
    PairwiseInteractionABFS_Deallocate(&(q->pwqcqc));
    SymmetricMatrix_Deallocate(&(q->qcqcPotentials));
    Real1DArray_Deallocate(&(q->qcmmPotentials));
    Real1DArray_Deallocate(&(q->qcCharges));
    QCAtomContainer_Deallocate(&(q->qc));
    free(q);
}

static PairwiseInteractionABFS *make_qc_interaction(const NBModelABFS *nb, int density)
{
    PairwiseInteractionABFS *pw = PairwiseInteractionABFS_Allocate();
    if (pw == NULL) return NULL;
    pw->dampingCutoff = nb->dampingCutoff; pw->innerCutoff = nb->innerCutoff; pw->outerCutoff = nb->outerCutoff;
    pw->splinePointDensity = density;
    pw->electrostaticSpline = PairwiseInteractionABFS_MakeElectrostaticSpline(pw, True, NULL);      /* atomic units */
    return pw;
}

RefQC *refqc_create(int n, const double *charges, const int *ljtypes,
                    int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                    int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                    int nexcl, const int *exclPairs, int n14, const int *pairs14,
                    int ntrans, const double *rot, const double *trans,
                    int nqc, const int *qcIndex, const int *qcAtomicNumber, int splinePointDensity)
{
    int i;
    Status status = Status_Continue;
    RefQC *q = (RefQC *) calloc(1, sizeof(RefQC));
    if (q == NULL || nqc <= 0) { free(q); return NULL; }
    q->nb = refnb_create(n, charges, ljtypes, ntypes, tableindex, tableA, tableB, ntypes14, tableindex14, tableA14, tableB14,
                         nexcl, exclPairs, n14, pairs14, ntrans, rot, trans);
    if (q->nb == NULL) { free(q); return NULL; }
    q->nqc = nqc;
    q->qc  = QCAtomContainer_Allocate(nqc);
    for (i = 0; i < nqc; i++) {
        q->qc->data[i].index        = qcIndex[i];
        q->qc->data[i].atomicNumber = qcAtomicNumber[i];
        q->qc->data[i].center       = i;
        q->nb->mm->data[qcIndex[i]].QACTIVE = False;
    }
    q->qcCharges      = Real1DArray_Allocate(nqc, &status);
    q->qcmmPotentials = Real1DArray_Allocate(nqc, &status);
    if (ntrans > 0) q->qcqcPotentials = SymmetricMatrix_Allocate(nqc);
    q->nb->nb->qcmmCoupling = QCMMLinkAtomCoupling_MM;
    /* the state is created anew with the QC atoms, as SetUp does for a new configuration */
    NBModelABFSState_Deallocate(&(q->nb->st));
    q->nb->st = NBModelABFSState_SetUp(q->nb->mm, q->qc, NULL, q->nb->excl, q->nb->i14, q->nb->lj, q->nb->lj14,
                                       q->qcCharges, q->qcmmPotentials, q->qcqcPotentials, q->nb->tc, q->nb->nb->qcmmCoupling);
    q->pwqcmm = make_qc_interaction(q->nb->nb, splinePointDensity);
    q->pwqcqc = make_qc_interaction(q->nb->nb, splinePointDensity);
    if (q->nb->st == NULL || q->pwqcmm == NULL || q->pwqcqc == NULL || status != Status_Continue) { refqc_destroy(q); return NULL; }
    return q;
}

/* One full pass: Update, MM/MM energy, QC/MM LJ energy, QC/MM potentials, and (given QC charges) the QC/MM electrostatic gradients.
 * energies[10] = emmel, emmlj, emmel14, emmlj14, eimmmel, eimmmlj, eqcmmlj, eqcmmlj14, eimqcmmlj, eimqcqclj
 * potentials[nqc] (atomic units), qcqc[nqc (nqc + 1) / 2] (packed lower triangle; image QC/QC potentials, only with symmetry),
 * gradLJ[3n] = gradients after the MM/MM + LJ calls, gradEl[3n] = what QCMMGradients adds for the given qcCharges. */
int refqc_energy(RefQC *q, const double *xyz, const double *box, const double *qcCharges,
                 double *energies, double *potentials, double *qcqc, double *gradLJ, double *gradEl, double *dEdM)
{
    RefNB *h = q->nb;
    NBModelABFSState *st = h->st;
    int i, updated;
    Status status = Status_Continue;
    for (i = 0; i < h->n; i++) {
        Coordinates3_Item(h->x, i, 0) = xyz[3 * i]; Coordinates3_Item(h->x, i, 1) = xyz[3 * i + 1]; Coordinates3_Item(h->x, i, 2) = xyz[3 * i + 2];
    }
    Coordinates3_Set(h->g, 0.0);
    if (h->ntrans > 0) {
        SymmetryParameters_SetCrystalParameters(h->sp, box[0], box[1], box[2], box[3], box[4], box[5]);
        Matrix33_Set(h->spg->dEdM, 0.0);
    }
    st->isNew = True;
    NBModelABFSState_Initialize(st, h->x, h->sp, h->g, h->spg);
    updated = (int) NBModelABFS_Update(h->nb, h->gen, st, &status);
    if (status != Status_Continue) return -1;
    NBModelABFS_MMMMEnergy(h->nb, h->pw, st);
    NBModelABFS_QCMMEnergyLJ(h->nb, h->pw, st);
    energies[0] = st->emmel;   energies[1] = st->emmlj;   energies[2] = st->emmel14; energies[3] = st->emmlj14;
    energies[4] = st->eimmmel; energies[5] = st->eimmmlj;
    energies[6] = st->eqcmmlj; energies[7] = st->eqcmmlj14; energies[8] = st->eimqcmmlj; energies[9] = st->eimqcqclj;
    for (i = 0; i < h->n; i++) {
        gradLJ[3 * i] = Coordinates3_Item(h->g, i, 0); gradLJ[3 * i + 1] = Coordinates3_Item(h->g, i, 1); gradLJ[3 * i + 2] = Coordinates3_Item(h->g, i, 2);
    }
    Real1DArray_Set(q->qcmmPotentials, 0.0);
    if (q->qcqcPotentials != NULL) SymmetricMatrix_Set_Zero(q->qcqcPotentials);
    NBModelABFS_QCMMPotentials(h->nb, q->pwqcmm, q->pwqcqc, st);
    for (i = 0; i < q->nqc; i++) potentials[i] = Real1DArray_Item(q->qcmmPotentials, i);
    if (q->qcqcPotentials != NULL && qcqc != NULL) {
        int j, k = 0;
        for (i = 0; i < q->nqc; i++) for (j = 0; j <= i; j++) qcqc[k++] = SymmetricMatrix_Get_Component(q->qcqcPotentials, i, j);
    }
    for (i = 0; i < q->nqc; i++) Real1DArray_Item(q->qcCharges, i) = qcCharges[i];
    NBModelABFS_QCMMGradients(h->nb, q->pwqcmm, q->pwqcqc, st);
    for (i = 0; i < h->n; i++) {
        gradEl[3 * i]     = Coordinates3_Item(h->g, i, 0) - gradLJ[3 * i];
        gradEl[3 * i + 1] = Coordinates3_Item(h->g, i, 1) - gradLJ[3 * i + 1];
        gradEl[3 * i + 2] = Coordinates3_Item(h->g, i, 2) - gradLJ[3 * i + 2];
    }
    if (dEdM != NULL && h->spg != NULL) {
        int r, c;
        for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) dEdM[3 * r + c] = Matrix33_Item(h->spg->dEdM, r, c);
    }
    return updated;
}

/* list sizes: nbmmmm, nbqcmmlj, nbqcmmel, nbmmmm14, nbqcmmlj14, nbqcmmel14, and images / pairs of inbmmmm, inbqcmmlj, inbqcmmel, inbqcqclj, inbqcqcel */
void refqc_counts(RefQC *q, long *out)
{
    NBModelABFSState *st = q->nb->st;
    PairList  *pl[6] = { st->nbmmmm, st->nbqcmmlj, st->nbqcmmel, st->nbmmmm14, st->nbqcmmlj14, st->nbqcmmel14 };
    ImageList *il[5] = { st->inbmmmm, st->inbqcmmlj, st->inbqcmmel, st->inbqcqclj, st->inbqcqcel };
    int k;
    for (k = 0; k < 6; k++) out[k] = (pl[k] == NULL) ? 0 : (long) pl[k]->npairs;
    for (k = 0; k < 5; k++) {
        out[6 + 2 * k]     = (il[k] == NULL) ? 0 : (long) ImageList_NumberOfImages(il[k]);
        out[6 + 2 * k + 1] = (il[k] == NULL) ? 0 : (long) ImageList_NumberOfPairs(il[k]);
    }
}

/* which = 0 nbmmmm, 1 nbqcmmlj, 2 nbqcmmel (first index of a QC/MM pair = position in the QC container) */
void refqc_get_pairs(RefQC *q, int which, int *pairs)
{
    NBModelABFSState *st = q->nb->st;
    dump_pairs(which == 0 ? st->nbmmmm : (which == 1 ? st->nbqcmmlj : st->nbqcmmel), pairs);
}
