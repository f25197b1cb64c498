/* oracle/nbabfs_oracle.h -- TEST INFRASTRUCTURE: CPU restatement of the reference algorithm.
 *
 * Plain-C, flat-array restatement of pDynamo 1.9.0's NBModelABFS MM/MM path (see nbabfs_oracle.c for the
 * reference file:line each function follows).  Pinned against the compiled reference (oracle/_ref) and the
 * committed golden fixtures by tests/test_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library; the product (pdynamo-mirror_b200/) never does.
 */
#ifndef NBABFS_ORACLE_H
#define NBABFS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcNB OrcNB;

OrcNB *orc_create(int n, const double *charges, const int *ljtypes,
                  int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                  int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                  int nexcl, const int *exclPairs, int n14, const int *pairs14,
                  int ntrans, const double *rot, const double *trans);
void   orc_destroy(OrcNB *h);
void   orc_set_fixed(OrcNB *h, int nfixed, const int *fixed);
void   orc_set_centering(OrcNB *h, int on);                      /* NBModelABFS.useCentering (call after orc_set_fixed) */   /* NBModelABFSState_SetUp's fixedAtoms; 0 clears */
void   orc_set_options(OrcNB *h, double damp, double inner, double outer, double list,
                       double dielectric, double elecScale14, int checkForInverses, int imageExpandFactor);
/* PairwiseInteractionABFS.{useAnalyticForm, splinePointDensity} + MakeSplines (pMolecule.PairwiseInteraction.pyx:204-239); after orc_set_options */
void   orc_set_interaction_form(OrcNB *h, int useAnalyticForm, int splinePointDensity);
/* same contract as refnb_energy (oracle/ref_driver.h); timings[2] = list update, energy */
int    orc_energy(OrcNB *h, const double *xyz, const double *box, int forceNew,
                  double *energies, double *grad, double *dEdM, double *timings);
long   orc_num_primary_pairs(OrcNB *h);
int    orc_num_images(OrcNB *h);
long   orc_num_image_pairs(OrcNB *h);
long   orc_num_14_pairs(OrcNB *h);
void   orc_get_primary_pairs(OrcNB *h, int *pairs);
void   orc_get_image_info(OrcNB *h, int k, int *info, double *scale);
void   orc_get_image_pairs(OrcNB *h, int k, int *pairs);
/* list-time image coordinates of image k (the coordinates the reference's cross list was built from) */
void   orc_get_image_coordinates(OrcNB *h, int k, double *xyz);

void   orc_make_factors(double damp, double inner, double outer, double *out21);
/* spline tables of PairwiseInteractionABFS_Make*Spline: which = 0 electrostatic kJ/mol, 1 LJ-A, 2 LJ-B, 3 electrostatic atomic units;
 * returns the number of points (x = NULL: only that); x, y, h: abscissae r^2, ordinates, second derivatives */
int    orc_make_spline(int which, double damp, double inner, double outer, int density, double *x, double *y, double *h);
void   orc_spline_evaluate(int n, const double *x, const double *y, const double *h, double x0, double *f, double *g);
void   orc_make_M(const double *box6, double *M9, double *invM9);
/* single pair: returns energies e[2] = {elect, lj} and dF = dE/d(r^2) for qij (already scaled), Aij, Bij */
void   orc_pair(const double *f21, double r2, double qij, double Aij, double Bij, double *e2, double *dF);

/* bonded MM terms (oracle/mmterms_oracle.c): same contract as refmm_energy (oracle/ref_driver.h) */
void   orc_mm_energy(int n, const double *xyz,
                     int nbond, const int *bonds, const double *bondEq, const double *bondFc,
                     int nangle, const int *angles, const double *angleEq, const double *angleFc,
                     int nub, const int *ubs, const double *ubEq, const double *ubFc,
                     int ndih, const int *dihedrals, const double *dihFc, const int *dihPeriod, const double *dihPhase,
                     int nimp, const int *impropers, const double *impEq, const double *impFc,
                     double *energies5, double *grad);

#ifdef __cplusplus
}
#endif
#endif
