#-------------------------------------------------------------------------------
# . File      : pMolecule.NBModelABFSB200.pyx
# . The NB model of pMolecule.NBModelABFS.pyx on libnbabfs_b200.so (B200): same options, same SetUp / Energy protocol,
# . system.DefineNBModel ( NBModelABFSB200 ( ) ) is the only change a pDynamo script needs.
#-------------------------------------------------------------------------------
"""Defines an ABFS NB model that runs on a B200 GPU."""

from pCore import CLibraryError, logFile, LogFileActive

cdef class NBModelABFSB200State:
    """The device-side state (pMolecule.NBModelABFSState.pyx)."""

    def __dealloc__ ( self ):
        if self.isOwner:
            NBModelABFSState_B200_Deallocate ( &self.cObject )
            self.isOwner = False

    def _Initialize ( self ):
        cdef int i
        self.cObject = NULL
        self.isOwner = False
        for i from 0 <= i < 6: self.energies[i]   = 0.0
        for i from 0 <= i < 4: self.qcEnergies[i] = 0.0

    def GetEnergies ( self, energies ):
        """Append the energies (labels of pMolecule.NBModelABFSState.pyx:41-59)."""
        if self.cObject != NULL:
            energies.append ( ( "MM/MM Elect."   , self.energies[0] ) )
            energies.append ( ( "MM/MM LJ"       , self.energies[1] ) )
            energies.append ( ( "MM/MM 1-4 Elect.", self.energies[2] ) )
            energies.append ( ( "MM/MM 1-4 LJ"   , self.energies[3] ) )
            if NBModelABFSState_B200_NumberOfImages ( self.cObject ) > 0:
                energies.append ( ( "MM/MM Image Elect.", self.energies[4] ) )
                energies.append ( ( "MM/MM Image LJ"    , self.energies[5] ) )

    @classmethod
    def Raw ( selfClass ):
        self = selfClass.__new__ ( selfClass )
        self._Initialize ( )
        return self

    def Summary ( self, log = logFile ):
        cdef long ncalls = 0, nupdates = 0
        if LogFileActive ( log ) and ( self.cObject != NULL ):
            NBModelABFSState_B200_GetStatistics ( self.cObject, &ncalls, &nupdates )
            summary = log.GetSummary ( )
            summary.Start ( "ABFS NB Model State Summary (B200)" )
            summary.Entry ( "MM/MM Pairs"       , "{:d}".format ( NBModelABFSState_B200_NumberOfPairs      ( self.cObject, -1 ) ) )
            summary.Entry ( "MM/MM 1-4 Pairs"   , "{:d}".format ( NBModelABFSState_B200_NumberOf14Pairs    ( self.cObject     ) ) )
            summary.Entry ( "MM/MM Images"      , "{:d}".format ( NBModelABFSState_B200_NumberOfImages     ( self.cObject     ) ) )
            summary.Entry ( "MM/MM Image Pairs" , "{:d}".format ( NBModelABFSState_B200_NumberOfImagePairs ( self.cObject     ) ) )
            summary.Entry ( "Number of Calls"   , "{:d}".format ( ncalls   ) )
            summary.Entry ( "Number of Updates" , "{:d}".format ( nupdates ) )
            summary.Stop ( )

cdef class NBModelABFSB200 ( NBModel ):
    """The ABFS NB model on a B200."""

    defaultAttributes = { "checkForInverses" : True , "dampingCutoff" : 0.5 , "device" : 0 , "dielectric" : 1.0 , "electrostaticScale14" : 1.0 ,
                          "imageExpandFactor" : 0 , "innerCutoff" : 8.0 , "label" : "ABFS NB Model (B200)" , "listCutoff" : 13.5 ,
                          "outerCutoff" : 12.0 , "useCentering" : False }

    def __init__ ( self, **keywordArguments ):
        for ( key, value ) in self.__class__.defaultAttributes.iteritems ( ): setattr ( self, key, value )
        self.mmmmPairwiseInteraction = None
        self.SetOptions ( **keywordArguments )

    def __getstate__ ( self ):
        return dict ( [ ( key, getattr ( self, key ) ) for key in self.__class__.defaultAttributes ] )

    def __setstate__ ( self, state ):
        self.__init__ ( **state )

    def Clear ( self, configuration ):
        """Clear up temporary data (pMolecule.NBModelABFS.pyx:108-113)."""
        if configuration is not None:
            for attribute in ( "nbState", ):
                if hasattr ( configuration, attribute ): delattr ( configuration, attribute )

    def SetOptions ( self, **keywordArguments ):
        """Option handling of pMolecule.NBModelABFS.pyx:140-179: unknown keys and a non-increasing cutoff sequence are errors."""
        for ( key, value ) in keywordArguments.iteritems ( ):
            if key == "mmmmPairwiseInteraction": self.mmmmPairwiseInteraction = value
            elif key in self.__class__.defaultAttributes: setattr ( self, key, value )
            else: raise ValueError ( "Invalid option: " + key + "." )
        if not ( 0.0 <= self.dampingCutoff <= self.innerCutoff < self.outerCutoff <= self.listCutoff ):
            raise ValueError ( "Invalid cutoff values." )

    def SetUp ( self, MMAtomContainer mmAtoms, QCAtomContainer qcAtoms, LJParameterContainer ljParameters, LJParameterContainer ljParameters14, Selection fixedAtoms,
                      SelfPairList interactions14, SelfPairList exclusions, symmetry, isolates, configuration, log = logFile ):
        """Set up the energy calculation (pMolecule.NBModelABFS.pyx:181-273)."""
        cdef NBModelABFSB200State       nbState
        cdef Coordinates3               coordinates3
        cdef SymmetryParameters         symmetryParameters
        cdef Transformation3Container   transformations
        cdef CTransformation3Container *ctransformations
        cdef CLJParameterContainer     *clj
        cdef CLJParameterContainer     *clj14
        cdef int      status, i, j, t, n, ntrans, nexcl, n14, updateDone
        cdef Real    *q
        cdef Real    *rot
        cdef Real    *trans
        cdef Integer *lt
        cdef Integer *excl
        cdef Integer *i14
        cdef Integer *qc
        cdef double   box6[6]
        cdef double  *cbox
        if configuration is None: return
        if not hasattr ( configuration, "nbState" ):
            if nbb200_device_count ( ) <= 0: raise CLibraryError ( "No CUDA device: NBModelABFSB200 has no CPU path." )
            status = Status_Continue
            n      = mmAtoms.cObject.natoms
            q      = Memory_Allocate_Array_Real    ( n )
            lt     = Memory_Allocate_Array_Integer ( n )
            for i from 0 <= i < n:
                q [i] = mmAtoms.cObject.data[i].charge
                lt[i] = mmAtoms.cObject.data[i].ljtype
            excl  = NULL ; nexcl = 0
            i14   = NULL ; n14   = 0
            if exclusions     is not None: excl = PairList_ToIntegerPairArray ( exclusions.cObject     ) ; nexcl = exclusions.cObject.npairs
            if interactions14 is not None: i14  = PairList_ToIntegerPairArray ( interactions14.cObject ) ; n14   = interactions14.cObject.npairs
            # . Symmetry operations: 3x3 rotations and translations, fractional (Transformation3Container.items[t]).
            ctransformations = NULL ; ntrans = 0 ; rot = NULL ; trans = NULL
            if ( symmetry is not None ) and hasattr ( symmetry, "transformations" ) and ( symmetry.transformations is not None ):
                transformations  = symmetry.transformations
                ctransformations = transformations.cObject
                ntrans           = ctransformations.nitems
                rot              = Memory_Allocate_Array_Real ( 9 * ntrans )
                trans            = Memory_Allocate_Array_Real ( 3 * ntrans )
                for t from 0 <= t < ntrans:
                    for i from 0 <= i < 3:
                        for j from 0 <= j < 3: rot[9*t+3*i+j] = Matrix33_GetItem ( ctransformations.items[t].rotation, i, j, NULL )
                        trans[3*t+i] = ctransformations.items[t].translation.data[i]
            clj   = ljParameters.cObject
            clj14 = clj
            if ljParameters14 is not None: clj14 = ljParameters14.cObject
            nbState         = NBModelABFSB200State.Raw ( )
            nbState.cObject = NBModelABFSState_B200_SetUp ( self.device, n, q, lt, clj.ntypes, clj.tableindex, clj.tableA, clj.tableB,
                                                            clj14.ntypes, clj14.tableindex, clj14.tableA, clj14.tableB,
                                                            nexcl, excl, n14, i14, ntrans, rot, trans, &status )
            Memory_Deallocate_Real ( &q ) ; Memory_Deallocate_Integer ( &lt ) ; Memory_Deallocate_Integer ( &excl ) ; Memory_Deallocate_Integer ( &i14 )
            Memory_Deallocate_Real ( &rot ) ; Memory_Deallocate_Real ( &trans )
            if ( nbState.cObject == NULL ) or ( status != Status_Continue ): raise CLibraryError ( "Unable to create the NB state: " + nbb200_last_error ( ) )
            nbState.isOwner = True
            if fixedAtoms is not None:
                NBModelABFSState_B200_SetFixedAtoms ( nbState.cObject, fixedAtoms.cObject.nindices, fixedAtoms.cObject.indices, &status )
            if ( qcAtoms is not None ) and ( qcAtoms.size > 0 ):
                boundary = set ( qcAtoms.BoundaryAtomSelection ( ) )          # . Boundary atoms stay on the MM/MM lists (NBModelABFSState.c:351).
                pure     = [ i for i in qcAtoms.QCAtomSelection ( ) if i not in boundary ]
                qc   = Memory_Allocate_Array_Integer ( len ( pure ) )
                for i from 0 <= i < len ( pure ): qc[i] = pure[i]
                NBModelABFSState_B200_SetQCAtoms ( nbState.cObject, len ( pure ), qc, &status )
                Memory_Deallocate_Integer ( &qc )
            NBModelABFSState_B200_SetUpCentering ( nbState.cObject, 1 if self.useCentering else 0, &status )   # . pMolecule.NBModelABFS.pyx:245.
            if status != Status_Continue: raise CLibraryError ( "Unable to set up the NB state: " + nbb200_last_error ( ) )
            setattr ( configuration, "nbState", nbState )
        nbState = configuration.nbState
        status  = Status_Continue
        NBModelABFS_B200_SetOptions ( nbState.cObject, self.dampingCutoff, self.innerCutoff, self.outerCutoff, self.listCutoff, self.dielectric,
                                      self.electrostaticScale14, 1 if self.checkForInverses else 0, self.imageExpandFactor )
        if self.mmmmPairwiseInteraction is not None:
            PairwiseInteractionABFS_B200_SetInteractionForm ( nbState.cObject, 1 if self.mmmmPairwiseInteraction.cObject.useAnalyticForm == CTrue else 0,
                                                              self.mmmmPairwiseInteraction.cObject.splinePointDensity, &status )
        coordinates3 = configuration.coordinates3
        cbox         = NULL
        if getattr ( configuration, "symmetryParameters", None ) is not None:
            symmetryParameters = configuration.symmetryParameters
            box6[0] = symmetryParameters.cObject.a     ; box6[1] = symmetryParameters.cObject.b    ; box6[2] = symmetryParameters.cObject.c
            box6[3] = symmetryParameters.cObject.alpha ; box6[4] = symmetryParameters.cObject.beta ; box6[5] = symmetryParameters.cObject.gamma
            cbox    = box6
        updateDone = NBModelABFS_B200_Update ( nbState.cObject, coordinates3.cObject.data, cbox, 0, &status )
        if status != Status_Continue: raise CLibraryError ( "Unable to create the NB lists: " + nbb200_last_error ( ) )
        if updateDone: nbState.Summary ( log = log )

    def Energy ( self, configuration ):
        """Energy and gradients (pMolecule.NBModelABFS.pyx:115-138)."""
        cdef NBModelABFSB200State       nbState
        cdef Coordinates3               gradients3
        cdef SymmetryParameterGradients symmetryParameterGradients
        cdef CReal2DArray *cmatrix
        cdef Real *cgradients
        cdef Real *cdEdM
        cdef int   status
        energies = []
        if ( configuration is not None ) and hasattr ( configuration, "nbState" ):
            nbState    = configuration.nbState
            status     = Status_Continue
            cgradients = NULL
            cdEdM      = NULL
            if getattr ( configuration, "gradients3", None ) is not None:
                gradients3 = configuration.gradients3
                cgradients = gradients3.cObject.data
            if getattr ( configuration, "symmetryParameterGradients", None ) is not None:
                symmetryParameterGradients = configuration.symmetryParameterGradients
                cmatrix = < CReal2DArray * > symmetryParameterGradients.cObject.dEdM     # . Matrix33 is a Real2DArray (Matrix33.h): compact 3 x 3.
                cdEdM   = &cmatrix.data[cmatrix.offset]
            NBModelABFS_B200_MMMMEnergy ( nbState.cObject, nbState.energies, cgradients, cdEdM, &status )
            if status != Status_Continue: raise CLibraryError ( "NB energy evaluation failed: " + nbb200_last_error ( ) )
            nbState.GetEnergies ( energies )
        return energies
