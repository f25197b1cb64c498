#-------------------------------------------------------------------------------
# . File      : pMolecule.NBModelABFSB200.pxd
# . Binding of libnbabfs_b200.so (include/nbabfs_b200.h) for pDynamo 1.9.0: goes next to pMolecule.NBModelABFS.pxd in
# . pMolecule-1.9.0/extensions/pyrex.  tests/test_integration_cython.py cythonizes it against the reference's own .pxd files.
#-------------------------------------------------------------------------------
from pCore.cDefinitions                    cimport Boolean, CFalse, CTrue, Integer, Real
from pCore.Coordinates3                    cimport CCoordinates3, Coordinates3
from pCore.Memory                          cimport Memory_Allocate_Array_Integer, Memory_Allocate_Array_Real, Memory_Deallocate_Integer, Memory_Deallocate_Real
from pCore.PairList                        cimport CPairList, SelfPairList, PairList_ToIntegerPairArray
from pCore.Selection                       cimport Selection, CSelection
from pCore.Status                          cimport Status, Status_Continue
from pCore.Transformation3Container        cimport Transformation3Container, CTransformation3Container
from pCore.Matrix33                        cimport CMatrix33, Matrix33_GetItem
from pCore.Real2DArray                     cimport CReal2DArray
from pMolecule.LJParameterContainer        cimport LJParameterContainer, CLJParameterContainer
from pMolecule.MMAtomContainer             cimport MMAtomContainer, CMMAtomContainer
from pMolecule.NBModel                     cimport NBModel
from pMolecule.PairwiseInteraction         cimport CPairwiseInteractionABFS, PairwiseInteractionABFS
from pMolecule.QCAtomContainer             cimport QCAtomContainer, CQCAtomContainer
from pMolecule.SymmetryParameterGradients  cimport SymmetryParameterGradients, CSymmetryParameterGradients
from pMolecule.SymmetryParameters          cimport SymmetryParameters, CSymmetryParameters

cdef extern from "nbabfs_b200.h":

    ctypedef struct NBB200State:
        pass

    cdef int          nbb200_device_count ( )
    cdef const char  *nbb200_last_error   ( )
    cdef NBB200State *NBModelABFSState_B200_SetUp ( int device, int n, const double *charges, const int *ljtypes,
                                                    int ntypes,   const int *tableindex,   const double *tableA,   const double *tableB,
                                                    int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                                                    int nexcl, const int *exclPairs, int n14, const int *pairs14,
                                                    int ntrans, const double *rot, const double *trans, int *status )
    cdef void  NBModelABFSState_B200_SetFixedAtoms  ( NBB200State *state, int nfixed, const int *fixed, int *status )
    cdef void  NBModelABFSState_B200_SetQCAtoms     ( NBB200State *state, int nqc, const int *qcAtoms, int *status )
    cdef void  NBModelABFSState_B200_SetUpCentering ( NBB200State *state, int useCentering, int *status )
    cdef void  NBModelABFSState_B200_Deallocate     ( NBB200State **state )
    cdef void  NBModelABFSState_B200_GetStatistics  ( NBB200State *state, long *numberOfCalls, long *numberOfUpdates )
    cdef void  NBModelABFS_B200_SetOptions          ( NBB200State *state, double dampingCutoff, double innerCutoff, double outerCutoff, double listCutoff,
                                                      double dielectric, double electrostaticScale14, int checkForInverses, int imageExpandFactor )
    cdef void  PairwiseInteractionABFS_B200_SetInteractionForm ( NBB200State *state, int useAnalyticForm, int splinePointDensity, int *status )
    cdef int   NBModelABFS_B200_Update              ( NBB200State *state, const double *xyz, const double *box6, int forceNew, int *status )
    cdef void  NBModelABFS_B200_MMMMEnergy          ( NBB200State *state, double *energies, double *grad, double *dEdM, int *status )
    cdef void  NBModelABFS_B200_QCMMEnergyLJ        ( NBB200State *state, double *energies4, double *grad, double *dEdM, int *status )
    cdef void  NBModelABFS_B200_QCMMPotentials      ( NBB200State *state, double *potentials, double *qcqcPotentials, int *status )
    cdef void  NBModelABFS_B200_QCMMGradients       ( NBB200State *state, const double *qcCharges, double *grad, double *dEdM, int *status )
    cdef void  nbb200_set_optimistic_updates        ( NBB200State *state, int on )
    cdef long  NBModelABFSState_B200_NumberOfPairs      ( NBB200State *state, int image )
    cdef int   NBModelABFSState_B200_NumberOfImages     ( NBB200State *state )
    cdef long  NBModelABFSState_B200_NumberOfImagePairs ( NBB200State *state )
    cdef long  NBModelABFSState_B200_NumberOf14Pairs    ( NBB200State *state )

cdef class NBModelABFSB200State:

    cdef NBB200State *cObject
    cdef double       energies[6]
    cdef double       qcEnergies[4]
    cdef public object isOwner

cdef class NBModelABFSB200 ( NBModel ):

    cdef public PairwiseInteractionABFS mmmmPairwiseInteraction
    cdef public object  label
    cdef public Real    dampingCutoff, innerCutoff, outerCutoff, listCutoff, dielectric, electrostaticScale14
    cdef public object  checkForInverses, useCentering
    cdef public Integer imageExpandFactor, device
