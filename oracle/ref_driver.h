/* oracle/ref_driver.h -- TEST INFRASTRUCTURE (never linked into the product library).
 *
 * Flat C driver around the UNMODIFIED reference C sources of pDynamo 1.9.0 (compiled in place from
 * /root/reference by oracle/Makefile).  It reproduces the call order of the Cython class
 * NBModelABFS.SetUp / NBModelABFS.Energy
 *   (pMolecule-1.9.0/extensions/pyrex/pMolecule.NBModelABFS.pyx:181-273,108-122)
 * and of System.Energy (pMolecule-1.9.0/pMolecule/System.py:272-318), which cannot be imported here
 * (Python 2 only).  All arrays are plain pointers so the driver can be used from ctypes.
 */
#ifndef REF_DRIVER_H
#define REF_DRIVER_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RefNB RefNB;

/* charges[n], ljtypes[n]; LJ tables as in LJParameterContainer (tableindex[nt*nt], tableA/B[nt(nt+1)/2]);
 * the 1-4 tables may be NULL (then the normal tables are used, as System does when lj14 is absent).
 * exclPairs[2*nexcl], pairs14[2*n14]: index pairs (copied).  Transformations: ntrans fractional
 * rotations rot[9*ntrans] (row-major) and translations trans[3*ntrans]; ntrans==0 => no symmetry (vacuum). */
RefNB *refnb_create(int n, const double *charges, const int *ljtypes,
                    int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                    int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                    int nexcl, const int *exclPairs, int n14, const int *pairs14,
                    int ntrans, const double *rot, const double *trans);
void   refnb_destroy(RefNB *h);
/* fixedAtoms of NBModelABFSState_SetUp: re-creates the state with the selection (0 clears); returns 1 on success */
int    refnb_set_fixed(RefNB *h, int nfixed, const int *fixed);
/* NBModelABFS.useCentering: NBModelABFSState_SetUpCentering on a newly created state; returns 1 on success */
int    refnb_set_centering(RefNB *h, int on);

/* NBModelABFS + generator options (pMolecule.NBModelABFS.pyx:59-71,140-179). method: 0 = automatic
 * (DetermineMethod), 1 = force O(N^2) direct (minimumPoints huge), 2 = force grid. useGridByCell as in generator. */
void   refnb_set_options(RefNB *h, double damp, double inner, double outer, double list,
                         double dielectric, double elecScale14, int checkForInverses, int imageExpandFactor,
                         double cellSizeFactor, int method, int useGridByCell, int sortIndices);

/* PairwiseInteractionABFS.{useAnalyticForm, splinePointDensity} + MakeSplines (pMolecule.PairwiseInteraction.pyx:204-239) for the
 * MM/MM interaction; call after refnb_set_options.  useAnalyticForm = 0: the cubic-spline branch of
 * PairwiseInteractionABFS_MMMMEnergy (pM/csource/PairwiseInteraction.c:431-531). */
void   refnb_set_interaction_form(RefNB *h, int useAnalyticForm, int splinePointDensity);

/* One System.Energy-style call: Initialize -> Update -> MMMMEnergy.
 * xyz[3n]; box = {a,b,c,alpha,beta,gamma} (ignored when ntrans==0); grad[3n] is ACCUMULATED into (may be NULL);
 * dEdM[9] accumulated into (may be NULL); energies[6] = {emmel, emmlj, emmel14, emmlj14, eimmmel, eimmmlj}.
 * forceNew != 0 marks the state new before the call (forces a list rebuild).
 * timings[2] (nullable) = seconds spent in NBModelABFS_Update and in NBModelABFS_MMMMEnergy.
 * Returns 1 if the lists were updated, 0 if not, <0 on error. */
int    refnb_energy(RefNB *h, const double *xyz, const double *box, int forceNew,
                    double *energies, double *grad, double *dEdM, double *timings);

/* list inspection after a call */
long   refnb_num_primary_pairs(RefNB *h);
int    refnb_num_images(RefNB *h);
long   refnb_num_image_pairs(RefNB *h);
void   refnb_get_primary_pairs(RefNB *h, int *pairs /* [2*npairs] (i,j) */);
/* image k: info[6] = {t, a, b, c, npairs, 0}; scale[1] */
void   refnb_get_image_info(RefNB *h, int k, int *info, double *scale);
void   refnb_get_image_pairs(RefNB *h, int k, int *pairs /* [2*npairs] (i, j) */);
long   refnb_num_14_pairs(RefNB *h);
int    refnb_uses_grid(RefNB *h);
int    refnb_num_threads(void);

/* helpers exposed for pinning the restatement */
void   refnb_make_factors(double damp, double inner, double outer, double *out21);
int    refnb_make_spline(int which, double damp, double inner, double outer, int density, double *x, double *y, double *h);
void   refnb_spline_evaluate(int n, const double *x, const double *y, double x0, double *f, double *g);
void   refnb_lj_table(int ntypes, const double *eps, const double *sigma, int amber,
                      int *tableindex, double *tableA, double *tableB);
void   refnb_make_M(const double *box6, double *M9, double *invM9);

/* bonded MM terms through the reference's own containers and *_Energy routines; per-term parameters; energies5 = {bond, angle,
 * Urey-Bradley, dihedral, improper}; grad[3n] (nullable) accumulated into */
void   refmm_energy(int n, const double *xyz,
                    int nbond, const int *bonds, const double *bondEq, const double *bondFc,
                    int nangle, const int *angles, const double *angleEq, const double *angleFc,
                    int nub, const int *ubs, const double *ubEq, const double *ubFc,
                    int ndih, const int *dihedrals, const double *dihFc, const int *dihPeriod, const double *dihPhase,
                    int nimp, const int *impropers, const double *impEq, const double *impFc,
                    double *energies5, double *grad);

/* QC/MM entry points of NBModelABFS with a QC region without boundary atoms, MM link-atom coupling (see ref_driver.c) */
typedef struct RefQC RefQC;
RefQC *refqc_create(int n, const double *charges, const int *ljtypes,
                    int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                    int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                    int nexcl, const int *exclPairs, int n14, const int *pairs14,
                    int ntrans, const double *rot, const double *trans,
                    int nqc, const int *qcIndex, const int *qcAtomicNumber, int splinePointDensity);
void   refqc_destroy(RefQC *q);
int    refqc_energy(RefQC *q, const double *xyz, const double *box, const double *qcCharges,
                    double *energies10, double *potentials, double *qcqc, double *gradLJ, double *gradEl, double *dEdM);
void   refqc_counts(RefQC *q, long *out16);
void   refqc_get_pairs(RefQC *q, int which, int *pairs);

#ifdef __cplusplus
}
#endif
#endif
