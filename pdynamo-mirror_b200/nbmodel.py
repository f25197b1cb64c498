"""Host-side mirror of pDynamo's NB model plugin surface for the ABFS model, backed by libnbabfs_b200.so.

Mirrors (same names, option keys, argument meaning and error behaviour):
  NBModel             pMolecule-1.9.0/extensions/pyrex/pMolecule.NBModel.pyx:23-114
  NBModelABFS         pMolecule-1.9.0/extensions/pyrex/pMolecule.NBModelABFS.pyx
  NBModelABFSState    pMolecule-1.9.0/extensions/pyrex/pMolecule.NBModelABFSState.pyx
  PairListGenerator   pCore-1.9.0/extensions/pyrex/pCore.PairListGenerator.pyx (option holder + B200 generators)
  PairwiseInteractionABFS  pMolecule-1.9.0/extensions/pyrex/pMolecule.PairwiseInteraction.pyx (option holder)
Only the MM/MM path is implemented on the device; QC/MM entry points are no-ops exactly as the reference's
are when there are no QC atoms (pMolecule-1.9.0/extensions/csource/NBModelABFS.c:308).
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import CLibraryError, d_, i_

_STATENAMES = ("nbState", "qcmmstate")          # pMolecule.NBModel.pyx:17


class PairwiseInteractionABFS:
    """Option holder for the ABFS interaction (pMolecule.PairwiseInteraction.pyx:111-251): analytic form (default) or the cubic-spline
    form (useAnalyticForm = False, splinePointDensity points per Angstrom); the spline tables themselves live in the device state
    (PairwiseInteractionABFS_B200_SetInteractionForm) and can be inspected with MakeSplines()."""

    def __init__(self, **options):
        # defaults of PairwiseInteractionABFS_Allocate (pMolecule-1.9.0/extensions/csource/PairwiseInteraction.c:28-42)
        self.dampingCutoff, self.innerCutoff, self.outerCutoff = 0.5, 8.0, 12.0
        self.useAnalyticForm, self.splinePointDensity = True, 50
        self.electrostaticModel, self.width1, self.width2 = "Delta/Delta", 0.0, 0.0
        self.SetOptions(**options)

    @classmethod
    def FromOptions(cls, **options):
        return cls(**options)

    def SetOptions(self, **kw):
        for key in ("dampingCutoff", "electrostaticModel", "innerCutoff", "outerCutoff", "splinePointDensity", "width1", "width2"):
            if key in kw:
                setattr(self, key, kw.pop(key))
        if "useAnalyticForm" in kw:
            self.useAnalyticForm = bool(kw.pop("useAnalyticForm"))
        if len(kw) > 0:
            raise ValueError("Invalid options: " + ", ".join(sorted(kw.keys())) + ".")
        self.CheckOptions()

    def CheckOptions(self):
        if self.electrostaticModel is None:
            self.electrostaticModel = "Delta/Delta"
        elif self.electrostaticModel not in ("Delta/Delta", "Delta/Gaussian", "Gaussian/Gaussian"):
            raise ValueError("Invalid pairwise interaction electrostatic model: " + self.electrostaticModel + ".")
        if self.electrostaticModel != "Delta/Delta":
            raise NotImplementedError("Gaussian electrostatic models (QC/MM couplings) are not implemented on the device")

    def MakeSplines(self, electrostatic=True, lennardJones=True, useAtomicUnits=False):
        """The tables MakeSplines builds in the reference (pMolecule.PairwiseInteraction.pyx:204-214), as a dict of (x, y, h) arrays:
        x = r^2, ordinates, second derivatives.  The device state builds the same tables itself when the spline form is selected."""
        out = {}
        # useAtomicUnits: the electrostatic spline of the QC/MM and QC/QC interactions (PairwiseInteractionABFS_MakeElectrostaticSpline, True)
        which = ([("electrostatic", 3 if useAtomicUnits else 0)] if electrostatic else []) + ([("lennardJonesA", 1), ("lennardJonesB", 2)] if lennardJones else [])
        L = _lib.lib()
        for name, w in which:
            n = L.PairwiseInteractionABFS_B200_MakeSpline(w, self.dampingCutoff, self.innerCutoff, self.outerCutoff, int(self.splinePointDensity), None, None, None)
            if n <= 0:
                raise ValueError("Invalid spline point density.")
            x, y, h = np.zeros(n), np.zeros(n), np.zeros(n)
            L.PairwiseInteractionABFS_B200_MakeSpline(w, self.dampingCutoff, self.innerCutoff, self.outerCutoff, int(self.splinePointDensity), d_(x), d_(y), d_(h))
            out[name] = (x, y, h)
        return out

    def MakeFactors(self):
        out = np.zeros(21)
        _lib.lib().PairwiseInteractionABFS_B200_MakeFactors(self.dampingCutoff, self.innerCutoff, self.outerCutoff, d_(out))
        return out

    def __getstate__(self):
        return dict(dampingCutoff=self.dampingCutoff, electrostaticModel=self.electrostaticModel, innerCutoff=self.innerCutoff, outerCutoff=self.outerCutoff,
                    splinePointDensity=self.splinePointDensity, useAnalyticForm=self.useAnalyticForm, width1=self.width1, width2=self.width2)

    def __setstate__(self, state):
        self.__init__(**state)


class PairListGenerator:
    """Option holder (pCore.PairListGenerator.pyx:36-110) plus the device generators.  The grid options are kept for
    interface compatibility; the device builder always uses its own cell grid and returns the same pair SET."""
    _KEYS = ("cutoff", "cutoffCellSizeFactor", "minimumCellExtent", "minimumCellSize", "minimumExtentFactor", "minimumPoints", "sortIndices", "useGridByCell")

    def __init__(self, **options):
        self.cutoff, self.cutoffCellSizeFactor, self.minimumCellExtent, self.minimumCellSize = 0.0, 0.5, 2, 1.0
        self.minimumExtentFactor, self.minimumPoints, self.sortIndices, self.useGridByCell = 1.5, 500, True, False
        self.device = 0
        self.SetOptions(**options)

    @classmethod
    def FromOptions(cls, **options):
        return cls(**options)

    def SetOptions(self, **kw):
        for key in self._KEYS:
            if key in kw:
                setattr(self, key, kw.pop(key))
        if len(kw) > 0:
            raise ValueError("Invalid options: " + ", ".join(sorted(kw.keys())) + ".")
        self.cellSize = self.cutoffCellSizeFactor * self.cutoff

    def __getstate__(self):
        return {k: getattr(self, k) for k in self._KEYS}

    def __setstate__(self, state):
        self.__init__(**state)

    def _take(self, n, pp, status):
        if n < 0 or status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("Unable to create pair list: " + _lib.last_error())
        out = np.ctypeslib.as_array(pp, shape=(max(n, 1), 2))[:n].copy()
        _lib.lib().nbb200_free(pp)
        return out

    def SelfPairListFromCoordinates3(self, coordinates3, exclusions=None):
        """All pairs (i, j), i != j once, within the cutoff and not excluded: PairListGenerator_SelfPairListFromCoordinates3."""
        x = np.ascontiguousarray(coordinates3, np.float64).reshape(-1, 3)
        ex = np.zeros((0, 2), np.int32) if exclusions is None else np.ascontiguousarray(exclusions, np.int32).reshape(-1, 2)
        pp, status = _lib.ip(), C.c_int(_lib.STATUS_CONTINUE)
        n = _lib.lib().PairListGenerator_B200_SelfPairListFromCoordinates3(self.device, len(x), d_(x), float(self.cutoff), len(ex),
                                                                           i_(ex) if len(ex) else None, C.byref(pp), C.byref(status))
        return self._take(n, pp, status)

    def CrossPairListFromDoubleCoordinates3(self, coordinates31, coordinates32):
        """All (i, j) with |x1_i - x2_j| within the cutoff: PairListGenerator_CrossPairListFromDoubleCoordinates3."""
        x1 = np.ascontiguousarray(coordinates31, np.float64).reshape(-1, 3)
        x2 = np.ascontiguousarray(coordinates32, np.float64).reshape(-1, 3)
        pp, status = _lib.ip(), C.c_int(_lib.STATUS_CONTINUE)
        n = _lib.lib().PairListGenerator_B200_CrossPairListFromDoubleCoordinates3(self.device, len(x1), d_(x1), len(x2), d_(x2), float(self.cutoff),
                                                                                  C.byref(pp), C.byref(status))
        return self._take(n, pp, status)


class QCMMInteractionState:
    """pMolecule.QCMMInteractionState (pM/pyrex/pMolecule.QCMMInteractionState.pyx:12-52, pM/cinclude/QCMMInteractionState.h): the arrays through
    which a QC model and the NB model talk -- QC charges in, potentials on the QC atoms (atomic units) and, with symmetry, the packed QC/QC image
    potentials out.  QC atoms in ascending atom index."""

    def __init__(self, extent, includeQCQC=False):
        self.qcCharges = np.zeros(extent)
        self.qcmmPotentials = np.zeros(extent)
        self.qcqcPotentials = np.zeros(extent * (extent + 1) // 2) if includeQCQC else None

    @classmethod
    def WithExtent(cls, extent, includeQCQC=False):
        return cls(extent, includeQCQC=includeQCQC)

    def Initialize(self):
        """QCMMInteractionState_Initialize: potentials are incremented by the NB model, so the QC model clears them first"""
        self.qcmmPotentials[:] = 0.0
        if self.qcqcPotentials is not None:
            self.qcqcPotentials[:] = 0.0


class NBModelABFSState:
    """Owner of the device-resident NB state (lists, reference coordinates, statistics)."""
    LABELS = ("MM/MM Elect.", "MM/MM LJ", "MM/MM 1-4 Elect.", "MM/MM 1-4 LJ", "MM/MM Image Elect.", "MM/MM Image LJ")

    def __init__(self):
        self.cObject = None
        self.isOwner = False
        self.n = 0
        self.energies = np.zeros(6)
        self.hasSymmetry = False
        self.numberOfCalls = 0
        self.numberOfUpdates = 0
        self.nqc = 0
        self.qcEnergies = None

    def __del__(self):
        try:
            self.Deallocate()
        except Exception:
            pass

    def Deallocate(self):
        if self.isOwner and self.cObject:
            h = C.c_void_p(self.cObject)
            _lib.lib().NBModelABFSState_B200_Deallocate(C.byref(h))
        self.cObject, self.isOwner = None, False

    def SetQCAtoms(self, indices):
        """Pure QC atoms leave every MM/MM list (NBModelABFSState_SetUp's qcAtoms -> mmSelection, NBModelABFSState.c:348-353); the QC/MM
        entry points (QCMMEnergyLJ / QCMMPotentials / QCMMGradients) then act on them."""
        idx = np.ascontiguousarray(indices, np.int32).reshape(-1)
        self.nqc = len(idx)
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().NBModelABFSState_B200_SetQCAtoms(self.cObject, len(idx), i_(idx) if len(idx) else None, C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("Unable to set the QC atoms. " + _lib.last_error())

    @staticmethod
    def _host_array(a, what):
        if a is None:
            return None
        if not (isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] and a.dtype == np.float64):
            raise ValueError(what + " must be a C-contiguous float64 array")
        return d_(a)

    def QCMMEnergyLJ(self, gradients3=None, dEdM=None):
        """NBModelABFS_QCMMEnergyLJ for the QC atoms of SetQCAtoms: returns (eqcmmlj, eqcmmlj14, eimqcmmlj, eimqcqclj); gradients3[n, 3] and
        dEdM[3, 3] (host, optional) are accumulated into."""
        e = np.zeros(4)
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().NBModelABFS_B200_QCMMEnergyLJ(self.cObject, d_(e), self._host_array(gradients3, "gradients3"), self._host_array(dEdM, "dEdM"), C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("QC/MM LJ energy failed. " + _lib.last_error())
        self.qcEnergies = e
        return e

    def QCMMPotentials(self, qcmmPotentials, qcqcPotentials=None):
        """NBModelABFS_QCMMPotentials: potentials on the QC atoms (atomic units) and the packed QC/QC image potentials, incremented in place."""
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().NBModelABFS_B200_QCMMPotentials(self.cObject, self._host_array(qcmmPotentials, "qcmmPotentials"), self._host_array(qcqcPotentials, "qcqcPotentials"), C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("QC/MM potentials failed. " + _lib.last_error())

    def QCMMGradients(self, qcCharges, gradients3, dEdM=None):
        """NBModelABFS_QCMMGradients: electrostatic QC/MM and QC/QC image gradients for the given QC charges, accumulated in place."""
        status = C.c_int(_lib.STATUS_CONTINUE)
        q = np.ascontiguousarray(qcCharges, np.float64)
        _lib.lib().NBModelABFS_B200_QCMMGradients(self.cObject, d_(q), self._host_array(gradients3, "gradients3"), self._host_array(dEdM, "dEdM"), C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("QC/MM gradients failed. " + _lib.last_error())

    def GetEnergies(self, energies):
        """Append (label, value) tuples; same non-NULL-list gating as pMolecule.NBModelABFSState.pyx:41-59."""
        e = self.energies
        if self.NumberOfPairs() > 0:
            energies.append((self.LABELS[0], float(e[0])))
            energies.append((self.LABELS[1], float(e[1])))
        if self.NumberOf14Pairs() > 0:
            energies.append((self.LABELS[2], float(e[2])))
            energies.append((self.LABELS[3], float(e[3])))
        qe = getattr(self, "qcEnergies", None)
        if qe is not None and self.nqc > 0:                     # order of pMolecule.NBModelABFSState.pyx:41-59
            energies.append(("QC/MM LJ", float(qe[0])))
        if self.NumberOfImages() > 0:
            energies.append((self.LABELS[4], float(e[4])))
            energies.append((self.LABELS[5], float(e[5])))
        if qe is not None and self.nqc > 0 and self.hasSymmetry:
            energies.append(("QC/MM Image LJ", float(qe[2])))
            energies.append(("QC/QC Image LJ", float(qe[3])))

    # list inspection -------------------------------------------------------------------------------
    def NumberOfPairs(self, image=-1):
        return _lib.lib().NBModelABFSState_B200_NumberOfPairs(self.cObject, image)

    def NumberOf14Pairs(self):
        return _lib.lib().NBModelABFSState_B200_NumberOf14Pairs(self.cObject)

    def NumberOfImages(self):
        return _lib.lib().NBModelABFSState_B200_NumberOfImages(self.cObject)

    def NumberOfImagePairs(self):
        return _lib.lib().NBModelABFSState_B200_NumberOfImagePairs(self.cObject)

    def Pairs(self, image=-1):
        n = self.NumberOfPairs(image)
        out = np.zeros((max(n, 1), 2), np.int32)
        status = C.c_int(_lib.STATUS_CONTINUE)
        m = _lib.lib().NBModelABFSState_B200_GetPairs(self.cObject, image, i_(out), C.byref(status))
        if m != n or status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("Unable to expand the pair list: " + _lib.last_error())
        return out[:n]

    def Images(self, pairs=True):
        out = []
        for k in range(self.NumberOfImages()):
            info, sc = np.zeros(6, np.int32), np.zeros(1)
            _lib.lib().NBModelABFSState_B200_GetImageInfo(self.cObject, k, i_(info), d_(sc))
            d = dict(t=int(info[0]), a=int(info[1]), b=int(info[2]), c=int(info[3]), scale=float(sc[0]), npairs=int(info[4]))
            if pairs:
                d["pairs"] = self.Pairs(k)
            out.append(d)
        return out

    def Counters(self):
        out = (C.c_long * 8)()
        _lib.lib().nbb200_get_counters(self.cObject, out)
        keys = ("tiles", "workItems", "iBlocks", "haloAtoms", "listPairs", "kernelLaunches", "chunkTiles", "images")
        return dict(zip(keys, [int(v) for v in out]))

    def Timings(self):
        out = np.zeros(8)
        _lib.lib().nbb200_get_timings(self.cObject, d_(out))
        return dict(listRebuild=out[0], tileForces=out[1], pairs14=out[2], displacementCheck=out[3], pairExpansion=out[4], prune=out[5])

    def Summary(self, log=None):
        if log is not None:
            log("ABFS NB Model State Summary: MM/MM Pairs %d, MM/MM 1-4 Pairs %d, MM/MM Image Images %d, MM/MM Image Pairs %d" %
                (self.NumberOfPairs(), self.NumberOf14Pairs(), self.NumberOfImages(), self.NumberOfImagePairs()))

    def StatisticsSummary(self, log=None):
        if log is not None:
            n = max(self.numberOfUpdates, 1)
            log("ABFS NB Model State Statistics: calls %d, updates %d, calls per update %.1f" % (self.numberOfCalls, self.numberOfUpdates, self.numberOfCalls / n))


class NBModel:
    """Abstract base (pMolecule.NBModel.pyx:23-114)."""

    def __init__(self, **options):
        self._Initialize()
        self._Allocate()
        self.SetOptions(**options)

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self._Initialize()
        self._Allocate()
        self.SetOptions(**state)

    def __copy__(self):
        return self.__class__(**self.__getstate__())

    def __deepcopy__(self, memo):
        return self.__copy__()

    def _Allocate(self):
        pass

    def _Initialize(self):
        pass

    def Clear(self, configuration):
        if configuration is not None:
            for name in _STATENAMES:
                if hasattr(configuration, name):
                    delattr(configuration, name)

    def Energy(self, configuration):
        pass

    @classmethod
    def FromOptions(cls, **options):
        return cls(**options)

    def QCMMGradients(self, configuration):
        pass

    def QCMMPotentials(self, configuration):
        pass

    def SetOptions(self, **keywordArguments):
        pass

    def SetUp(self, mmAtoms, qcAtoms, ljParameters, ljParameters14, fixedAtoms, interactions14, exclusions, symmetry, isolates, configuration, log=None):
        pass

    def Summary(self, log=None):
        pass


class NBModelABFS(NBModel):
    """Atom-based force-switching NB model on the GPU; drop-in for pMolecule.NBModelABFS.

    Extra (non-reference) options, all optional: device (CUDA ordinal, default 0), updateFrequency (force a list rebuild every
    k-th call; default 0 = the reference's displacement heuristic only), overwriteGradients (default False = the reference's
    accumulation into configuration.gradients3; True: the NB call SETS the gradients, for callers that evaluate the NB term first
    and so need neither a zero fill of the host array nor its upload) and optimisticUpdates (default False; True: SetUp enqueues
    CheckForUpdate's displacement test without waiting for it and Energy reads the decision with its results -- one host wait per
    SetUp + Energy pair, the same numbers and update counts; a call in which an update turns out to be due is evaluated twice)."""

    def _Initialize(self):
        self.generator = None
        self.label = "ABFS"
        self.mmmmPairwiseInteraction = None
        self.qcmmPairwiseInteraction = None
        self.qcqcPairwiseInteraction = None

    def _Allocate(self):
        # defaults of NBModelABFS_Allocate (pMolecule-1.9.0/extensions/csource/NBModelABFS.c:28-37,174-190)
        self.checkForInverses, self.dampingCutoff, self._dielectric, self._electrostaticScale14 = True, 0.5, 1.0, 1.0
        self.imageExpandFactor, self._innerCutoff, self._listCutoff, self._outerCutoff = 0, 8.0, 13.5, 12.0
        self.qcmmCoupling, self.useCentering = "RC Coupling", False
        self.device, self.updateFrequency, self.overwriteGradients, self.optimisticUpdates = 0, 0, False, False

    def __getstate__(self):
        state = dict(checkForInverses=self.checkForInverses, imageExpandFactor=self.imageExpandFactor, dampingCutoff=self.dampingCutoff,
                     dielectric=self._dielectric, electrostaticScale14=self._electrostaticScale14, innerCutoff=self._innerCutoff,
                     listCutoff=self._listCutoff, outerCutoff=self._outerCutoff, qcmmCoupling=self.qcmmCoupling, useCentering=self.useCentering,
                     device=self.device, updateFrequency=self.updateFrequency, overwriteGradients=self.overwriteGradients,
                     optimisticUpdates=getattr(self, "optimisticUpdates", False))
        if self.generator is not None:
            state["generator"] = self.generator
        if self.mmmmPairwiseInteraction is not None:
            state["mmmmPairwiseInteraction"] = self.mmmmPairwiseInteraction
        return state

    def CheckGenerator(self):
        if self.generator is None:
            self.generator = PairListGenerator.FromOptions(cutoff=self._listCutoff, cutoffCellSizeFactor=0.5, minimumCellExtent=2, minimumCellSize=3.0,
                                                           minimumExtentFactor=1.5, minimumPoints=500, sortIndices=False, useGridByCell=True)
        else:
            self.generator.SetOptions(cutoff=self._listCutoff)
        self.generator.device = self.device

    def CheckPairwiseInteractions(self):
        if self.mmmmPairwiseInteraction is None:
            self.mmmmPairwiseInteraction = PairwiseInteractionABFS.FromOptions(dampingCutoff=self.dampingCutoff, innerCutoff=self._innerCutoff, outerCutoff=self._outerCutoff)

    def ClearPairwiseInteractions(self):
        for label in ("mmmmPairwiseInteraction", "qcmmPairwiseInteraction", "qcqcPairwiseInteraction"):
            setattr(self, label, None)

    def SetOptions(self, **kw):
        simple = dict(dampingCutoff="dampingCutoff", dielectric="_dielectric", electrostaticScale14="_electrostaticScale14", generator="generator",
                      imageExpandFactor="imageExpandFactor", innerCutoff="_innerCutoff", listCutoff="_listCutoff", outerCutoff="_outerCutoff",
                      mmmmPairwiseInteraction="mmmmPairwiseInteraction", qcmmPairwiseInteraction="qcmmPairwiseInteraction",
                      qcqcPairwiseInteraction="qcqcPairwiseInteraction", device="device", updateFrequency="updateFrequency",
                      overwriteGradients="overwriteGradients", optimisticUpdates="optimisticUpdates")
        for key, attr in simple.items():
            if key in kw:
                setattr(self, attr, kw.pop(key))
        if "checkForInverses" in kw:
            self.checkForInverses = bool(kw.pop("checkForInverses"))
        if "useCentering" in kw:
            self.useCentering = bool(kw.pop("useCentering"))
        if "qcmmCoupling" in kw:
            coupling = kw.pop("qcmmCoupling")
            if coupling not in ("MM Coupling", "RC Coupling", "RD Coupling"):
                raise TypeError("Unrecognized QC/MM coupling option: " + str(coupling) + ".")
            self.qcmmCoupling = coupling
        if len(kw) > 0:
            raise ValueError("Invalid options: " + ", ".join(sorted(kw.keys())) + ".")
        if (self.dampingCutoff < 0.0) or (self._innerCutoff < self.dampingCutoff) or (self._outerCutoff < self._innerCutoff) or (self._listCutoff < self._outerCutoff):
            raise TypeError("Invalid cutoff values: damping - {:.3f}; inner - {:.3f}; outer - {:.3f}; list - {:.3f}.".format(
                self.dampingCutoff, self._innerCutoff, self._outerCutoff, self._listCutoff))
        self.CheckGenerator()
        self.CheckPairwiseInteractions()

    # read-only properties (pMolecule.NBModelABFS.pyx:292-302)
    dielectric = property(lambda self: self._dielectric)
    electrostaticScale14 = property(lambda self: self._electrostaticScale14)
    innerCutoff = property(lambda self: self._innerCutoff)
    listCutoff = property(lambda self: self._listCutoff)
    outerCutoff = property(lambda self: self._outerCutoff)

    def _push_options(self, nbState):
        pw = self.mmmmPairwiseInteraction
        _lib.lib().NBModelABFS_B200_SetOptions(nbState.cObject, pw.dampingCutoff, pw.innerCutoff, pw.outerCutoff, self._listCutoff,
                                               self._dielectric, self._electrostaticScale14, int(self.checkForInverses), int(self.imageExpandFactor))
        _lib.lib().nbb200_set_optimistic_updates(nbState.cObject, 1 if getattr(self, "optimisticUpdates", False) else 0)
        form = (True, 0) if pw.useAnalyticForm else (False, int(pw.splinePointDensity))
        if getattr(nbState, "_interactionForm", (True, 0)) != form:      # the library rebuilds the tables itself when the cutoffs change
            status = C.c_int(_lib.STATUS_CONTINUE)
            _lib.lib().PairwiseInteractionABFS_B200_SetInteractionForm(nbState.cObject, int(form[0]), form[1], C.byref(status))
            if status.value != _lib.STATUS_CONTINUE:
                raise CLibraryError("Unable to make the interaction splines. " + _lib.last_error())
            nbState._interactionForm = form

    def SetUp(self, mmAtoms, qcAtoms, ljParameters, ljParameters14, fixedAtoms, interactions14, exclusions, symmetry, isolates, configuration, log=None):
        """Create / reuse configuration.nbState, hand over this call's coordinates and update the lists if needed."""
        if configuration is None:
            return
        qcIndices = None
        if qcAtoms is not None and len(qcAtoms) > 0:
            # QC regions without boundary (link) atoms: the indices of the pure QC atoms (QCAtomContainer_MakePureSelection)
            if getattr(qcAtoms, "nboundary", 0):
                raise NotImplementedError("QC regions with boundary atoms are not implemented on the device")
            qcIndices = np.ascontiguousarray(getattr(qcAtoms, "indices", qcAtoms), np.int32).reshape(-1)
        L = _lib.lib()
        if not hasattr(configuration, "nbState"):
            transformations = getattr(symmetry, "transformations", None) if symmetry is not None else None
            q = np.ascontiguousarray(mmAtoms.charges, np.float64)
            lt = np.ascontiguousarray(mmAtoms.ljtypes, np.int32)
            ex = np.zeros((0, 2), np.int32) if exclusions is None else np.ascontiguousarray(exclusions.pairs, np.int32).reshape(-1, 2)
            p14 = np.zeros((0, 2), np.int32) if interactions14 is None else np.ascontiguousarray(interactions14.pairs, np.int32).reshape(-1, 2)
            lj14 = ljParameters14 if ljParameters14 is not None else ljParameters
            if transformations is None:
                ntr, rot, trn = 0, None, None
            else:
                rot = np.ascontiguousarray(transformations.rotations, np.float64).reshape(-1)
                trn = np.ascontiguousarray(transformations.translations, np.float64).reshape(-1)
                ntr = len(trn) // 3
            status = C.c_int(_lib.STATUS_CONTINUE)
            nbState = NBModelABFSState()
            nbState.cObject = L.NBModelABFSState_B200_SetUp(int(self.device), len(q), d_(q), i_(lt),
                                                            ljParameters.ntypes, i_(ljParameters.tableindex), d_(ljParameters.tableA), d_(ljParameters.tableB),
                                                            lj14.ntypes, i_(lj14.tableindex), d_(lj14.tableA), d_(lj14.tableB),
                                                            len(ex), i_(ex) if len(ex) else None, len(p14), i_(p14) if len(p14) else None,
                                                            ntr, d_(rot), d_(trn), C.byref(status))
            nbState.isOwner = True
            nbState.n = len(q)
            nbState.hasSymmetry = ntr > 0
            if (not nbState.cObject) or (status.value != _lib.STATUS_CONTINUE):
                raise CLibraryError("Unable to create NB state. " + _lib.last_error())
            if fixedAtoms is not None and len(fixedAtoms) > 0:     # NBModelABFSState_SetUp's fixedAtoms (a Selection in the reference)
                fx = np.ascontiguousarray(getattr(fixedAtoms, "indices", fixedAtoms), np.int32).reshape(-1)
                L.NBModelABFSState_B200_SetFixedAtoms(nbState.cObject, len(fx), i_(fx), C.byref(status))
                if status.value != _lib.STATUS_CONTINUE:
                    raise CLibraryError("Unable to create NB state. " + _lib.last_error())
            # NBModelABFSState_SetUpCentering ( nbState.cObject, self.cObject.useCentering, &status )  (pMolecule.NBModelABFS.pyx:245)
            L.NBModelABFSState_B200_SetUpCentering(nbState.cObject, 1 if self.useCentering else 0, C.byref(status))
            if status.value != _lib.STATUS_CONTINUE:
                raise CLibraryError("Unable to create NB state. " + _lib.last_error())
            if qcIndices is not None:
                # qcmmstate = QCMMInteractionState.WithExtent ( qcAtoms.size, includeQCQC = ( ctransformations != NULL ) )  (pMolecule.NBModelABFS.pyx:218-219)
                setattr(configuration, "qcmmstate", QCMMInteractionState(len(qcIndices), includeQCQC=ntr > 0))
                nbState.SetQCAtoms(np.sort(qcIndices))
            setattr(configuration, "nbState", nbState)
        nbState = configuration.nbState
        self._push_options(nbState)
        coordinates3 = getattr(configuration, "coordinates3", None)
        symmetryParameters = getattr(configuration, "symmetryParameters", None)
        if coordinates3 is None:
            return
        x = np.ascontiguousarray(coordinates3, np.float64).reshape(-1, 3)
        if x.shape[0] != nbState.n:
            raise CLibraryError("Unable to create NB lists. Coordinate array has the wrong extent.")
        box = None if symmetryParameters is None else np.ascontiguousarray(symmetryParameters.box6, np.float64)
        nbState.numberOfCalls += 1
        force = 1 if (self.updateFrequency > 0 and (nbState.numberOfCalls - 1) % self.updateFrequency == 0) else 0
        status = C.c_int(_lib.STATUS_CONTINUE)
        updateDone = L.NBModelABFS_B200_Update(nbState.cObject, d_(x), d_(box), force, C.byref(status)) == 1
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("Unable to create NB lists. " + _lib.last_error())
        if updateDone:
            nbState.numberOfUpdates += 1
            nbState.Summary(log=log)

    def Energy(self, configuration):
        """Energies (list of (label, value)) and, when the configuration carries them, gradients and dE/dM (accumulated)."""
        energies = []
        if hasattr(configuration, "nbState"):
            nbState = configuration.nbState
            g = getattr(configuration, "gradients3", None)
            spg = getattr(configuration, "symmetryParameterGradients", None)
            if g is not None and not (isinstance(g, np.ndarray) and g.dtype == np.float64 and g.flags["C_CONTIGUOUS"]):
                raise TypeError("gradients3 must be a C-contiguous float64 array (it is accumulated into in place)")
            dEdM = None if spg is None else spg.dEdM
            status = C.c_int(_lib.STATUS_CONTINUE)
            # System.Energy of the mirror evaluates the NB term first on a gradient array it has declared zero: setting = accumulating
            known_zero = bool(getattr(configuration, "gradientsAreZero", False)) and g is not None
            if known_zero:
                configuration.gradientsAreZero = False
            _lib.lib().nbb200_set_gradient_overwrite(nbState.cObject, 1 if (self.overwriteGradients or known_zero) else 0)
            _lib.lib().NBModelABFS_B200_MMMMEnergy(nbState.cObject, d_(nbState.energies), d_(g), d_(dEdM), C.byref(status))
            if known_zero and not self.overwriteGradients:       # the state keeps the model's own option between calls
                _lib.lib().nbb200_set_gradient_overwrite(nbState.cObject, 0)
            if status.value != _lib.STATUS_CONTINUE:
                raise CLibraryError("NB energy evaluation failed. " + _lib.last_error())
            if getattr(self, "optimisticUpdates", False):        # an update may have been decided (and done) inside the energy call: the library counts
                ncalls, nupd = C.c_long(0), C.c_long(0)
                _lib.lib().NBModelABFSState_B200_GetStatistics(nbState.cObject, C.byref(ncalls), C.byref(nupd))
                nbState.numberOfUpdates = int(nupd.value)
            if nbState.nqc > 0:                                  # NBModelABFS_QCMMEnergyLJ ( ... )  (pMolecule.NBModelABFS.pyx:120)
                nbState.QCMMEnergyLJ(g, dEdM)
            nbState.GetEnergies(energies)
        return energies

    def QCMMGradients(self, configuration):
        """Calculate the QC/MM electrostatic gradients (pMolecule.NBModelABFS.pyx:124-130): configuration.qcmmstate.qcCharges in,
        configuration.gradients3 / symmetryParameterGradients accumulated into."""
        if hasattr(configuration, "nbState") and hasattr(configuration, "qcmmstate"):
            g = getattr(configuration, "gradients3", None)
            if g is None:
                return
            spg = getattr(configuration, "symmetryParameterGradients", None)
            configuration.nbState.QCMMGradients(configuration.qcmmstate.qcCharges, g, None if spg is None else spg.dEdM)

    def QCMMPotentials(self, configuration):
        """Calculate the QC/MM electrostatic potentials (pMolecule.NBModelABFS.pyx:132-138) into configuration.qcmmstate."""
        if hasattr(configuration, "nbState") and hasattr(configuration, "qcmmstate"):
            st = configuration.qcmmstate
            configuration.nbState.QCMMPotentials(st.qcmmPotentials, st.qcqcPotentials)

    def Summary(self, log=None):
        if log is not None:
            log("ABFS NB Model Summary: Dielectric %.6f, El. 1-4 Scaling %.6f, Damping Cutoff %.6f, Inner Cutoff %.6f, List Cutoff %.6f, Outer Cutoff %.6f, Use Centering %r" %
                (self._dielectric, self._electrostaticScale14, self.dampingCutoff, self._innerCutoff, self._listCutoff, self._outerCutoff, self.useCentering))
