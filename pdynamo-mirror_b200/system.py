"""A minimal Python-3 stand-in for the parts of pDynamo's System / EnergyModel / Configuration that the NB
plugin is called through, so that the drop-in can be exercised with the reference's call order:
  System.DefineNBModel / DefineSymmetry / Energy   pMolecule-1.9.0/pMolecule/System.py:182-197,243-258,272-318
  Configuration (persistent nbState, temporary gradients3)   pMolecule-1.9.0/pMolecule/Configuration.py:25-52
  SystemWithTimings keys "NB Set Up" / "NB Evaluation"       pMolecule-1.9.0/pMolecule/SystemWithTimings.py:20-96
The real pDynamo Python layer is Python 2 and cannot be imported here; INTEGRATION.md shows the binding into it."""
import time
import numpy as np

from .nbmodel import NBModel


class MMAtomContainer:
    def __init__(self, charges, ljtypes):
        self.charges = np.ascontiguousarray(charges, np.float64)
        self.ljtypes = np.ascontiguousarray(ljtypes, np.int32)

    def __len__(self):
        return len(self.charges)


class LJParameterContainer:
    def __init__(self, tableindex, tableA, tableB):
        self.tableindex = np.ascontiguousarray(tableindex, np.int32)
        self.tableA = np.ascontiguousarray(tableA, np.float64)
        self.tableB = np.ascontiguousarray(tableB, np.float64)
        self.ntypes = int(round(len(self.tableindex) ** 0.5))


class SelfPairList:
    def __init__(self, pairs):
        self.pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)

    def __len__(self):
        return len(self.pairs)


class Transformation3Container:
    def __init__(self, rotations, translations):
        self.rotations = np.ascontiguousarray(rotations, np.float64).reshape(-1, 3, 3)
        self.translations = np.ascontiguousarray(translations, np.float64).reshape(-1, 3)

    @classmethod
    def Identity(cls):
        """pCore.Transformation3Container.pyx:159-168"""
        return cls(np.eye(3)[None], np.zeros((1, 3)))


class SymmetryParameters:
    def __init__(self, a, b, c, alpha=90.0, beta=90.0, gamma=90.0):
        self.box6 = np.array([a, b, c, alpha, beta, gamma], dtype=np.float64)

    def SetCrystalParameters(self, a, b, c, alpha, beta, gamma):
        """pMolecule.SymmetryParameters.SetCrystalParameters (pM/pyrex/pMolecule.SymmetryParameters.pyx): new cell lengths and angles"""
        self.box6 = np.array([a, b, c, alpha, beta, gamma], dtype=np.float64)


class SymmetryParameterGradients:
    def __init__(self):
        self.dEdM = np.zeros((3, 3))


class Symmetry:
    def __init__(self, transformations):
        self.transformations = transformations


class Configuration:
    _TEMPORARY = ("energyTerms", "gradients3", "symmetryParameterGradients", "gradientsAreZero")

    def __init__(self):
        self.coordinates3 = None
        self.symmetryParameters = None

    def ClearTemporaryAttributes(self):
        for name in self._TEMPORARY:
            if hasattr(self, name):
                delattr(self, name)

    def SetTemporaryAttribute(self, name, value):
        setattr(self, name, value)


class EnergyModel:
    def __init__(self):
        self.mmAtoms = self.ljParameters = self.ljParameters14 = self.exclusions = self.interactions14 = self.nbModel = None
        self.electrostaticScale14 = 1.0
        self.mmTerms = []                      # bonded term containers (pMolecule.MMModel / EnergyModel.mmTerms), evaluated before the NB model

    def ClearNBModel(self, configuration):
        if self.nbModel is not None:
            self.nbModel.Clear(configuration)
        self.nbModel = None


class System:
    """System.FromWorkload(dict) builds the MM data the NB model needs from a workloads.py system dict."""

    def __init__(self):
        self.configuration = Configuration()
        self.energyModel = None
        self.symmetry = None
        self.timings = {"NB Set Up": 0.0, "NB Evaluation": 0.0, "Energy": 0.0}
        self.fixedAtoms = None                                 # system.hardConstraints.fixedAtoms of the reference: indices of atoms that do not move
        self._gradients = None

    @classmethod
    def FromWorkload(cls, w):
        self = cls()
        em = self.energyModel = EnergyModel()
        em.mmAtoms = MMAtomContainer(w["charges"], w["ljtypes"])
        em.ljParameters = LJParameterContainer(w["tableindex"], w["tableA"], w["tableB"])
        em.ljParameters14 = LJParameterContainer(w["tableindex14"], w["tableA14"], w["tableB14"])
        em.exclusions = SelfPairList(w["exclusions"]) if len(w["exclusions"]) else None
        em.interactions14 = SelfPairList(w["pairs14"]) if len(w["pairs14"]) else None
        em.electrostaticScale14 = w.get("electrostaticScale14", 1.0)
        self.masses = w.get("masses")
        if w.get("bonded") is not None:
            from .mmterms import containers_from_bonded
            em.mmTerms = containers_from_bonded(w["bonded"])
        if w.get("fixed") is not None and len(w["fixed"]) > 0:
            self.fixedAtoms = np.ascontiguousarray(w["fixed"], np.int32)
        from ._lib import pinned_array
        self.coordinates3 = pinned_array(w["xyz"].shape)        # page-locked: DMA without a staging copy
        self.coordinates3[...] = w["xyz"]
        if w["box"] is not None:
            b = w["box"]
            self.DefineSymmetry(a=b[0], b=b[1], c=b[2], alpha=b[3], beta=b[4], gamma=b[5],
                                transformations=Transformation3Container(w["rot"], w["trans"]))
        return self

    coordinates3 = property(lambda self: self.configuration.coordinates3,
                            lambda self, v: setattr(self.configuration, "coordinates3", v))
    symmetryParameters = property(lambda self: self.configuration.symmetryParameters,
                                  lambda self, v: setattr(self.configuration, "symmetryParameters", v))

    def DefineQCRegion(self, indices):
        """Stand-in for System.DefineQCModel ( qcModel, qcSelection = ... ) (pMolecule/System.py): the atoms of the QC region (no boundary
        atoms).  The NB model then keeps them off the MM/MM lists and serves the QC/MM entry points; the QC model itself is out of scope."""
        if self.energyModel is None:
            self.energyModel = EnergyModel()
        self.energyModel.qcAtoms = None if indices is None or len(indices) == 0 else np.sort(np.asarray(indices, dtype=np.int32))
        if self.energyModel.nbModel is not None:
            self.energyModel.nbModel.Clear(self.configuration)          # the NB state is set up anew with the QC region

    def DefineNBModel(self, nbModel):
        if isinstance(nbModel, NBModel):
            if self.energyModel is None:
                self.energyModel = EnergyModel()
            else:
                self.energyModel.ClearNBModel(self.configuration)
            self.energyModel.nbModel = nbModel
            try:
                nbModel.SetOptions(electrostaticScale14=float(self.energyModel.electrostaticScale14))
            except Exception:
                pass

    def DefineSymmetry(self, a=None, alpha=90.0, b=None, beta=90.0, c=None, gamma=90.0, transformations=None):
        if transformations is None:
            transformations = Transformation3Container.Identity()
        self.symmetry = Symmetry(transformations)
        b = a if b is None else b
        c = a if c is None else c
        self.symmetryParameters = SymmetryParameters(a, b, c, alpha, beta, gamma)

    def Energy(self, log=None, doGradients=False):
        if (self.coordinates3 is None) or (self.energyModel is None):
            return None
        t0 = time.perf_counter()
        cfg = self.configuration
        cfg.ClearTemporaryAttributes()
        if doGradients:
            n = len(self.energyModel.mmAtoms)
            # System.Energy zeroes gradients3 and every term accumulates (pMolecule-1.9.0/pMolecule/System.py:272-318).  In this mirror the NB
            # model is evaluated FIRST, on an array this method has just "zeroed": 0 + g = g exactly, so the zero fill, its upload and the
            # accumulation collapse into "the NB term sets the array" (flag gradientsAreZero, consumed by NBModelABFS.Energy) -- the same
            # numbers as fill + accumulate, without touching 2 x 24 n bytes on the host.  A caller that hands its own gradient array to
            # NBModelABFS.Energy gets the reference's accumulate semantics (option overwriteGradients = False).
            if self._gradients is None or self._gradients.shape[0] != n:
                from ._lib import pinned_array
                self._gradients = pinned_array((n, 3))          # reused, page-locked gradient storage
            if self.energyModel.nbModel is not None:
                cfg.SetTemporaryAttribute("gradientsAreZero", True)
            else:
                self._gradients.fill(0.0)
            cfg.SetTemporaryAttribute("gradients3", self._gradients)
            if self.symmetry is not None:
                cfg.SetTemporaryAttribute("symmetryParameterGradients", SymmetryParameterGradients())
        em = self.energyModel
        nbTerms, mmTerms = [], []
        if em.nbModel is not None:
            t1 = time.perf_counter()
            em.nbModel.SetUp(em.mmAtoms, getattr(em, "qcAtoms", None), em.ljParameters, em.ljParameters14, self.fixedAtoms, em.interactions14, em.exclusions, self.symmetry, None, cfg, log=log)
            t2 = time.perf_counter()
            nbTerms = em.nbModel.Energy(cfg)
            t3 = time.perf_counter()
            self.timings["NB Set Up"] += t2 - t1
            self.timings["NB Evaluation"] += t3 - t2
        if len(em.mmTerms) > 0:
            # System.Energy: `for mmterm in em.mmTerms: ...Energy(coordinates3, gradients3)` (System.py:294-297); here ONE device call for all
            # containers.  Evaluated after the NB model (which may SET the gradients, see above); the terms are reported in the reference's order.
            if getattr(self, "_mmTermsDevice", None) is None or self._mmTermsDevice.containers != em.mmTerms:
                from .mmterms import MMTermsB200
                self._mmTermsDevice = MMTermsB200(len(em.mmAtoms), em.mmTerms, device=int(getattr(em.nbModel, "device", 0) or 0))
            self._mmTermsDevice.Energy(self.coordinates3, cfg.gradients3 if doGradients else None)
            mmTerms = self._mmTermsDevice.EnergyTerms()
        terms = mmTerms + nbTerms
        if doGradients and self.fixedAtoms is not None and len(self.fixedAtoms) > 0:
            cfg.gradients3[np.asarray(self.fixedAtoms, np.int64)] = 0.0      # System.Energy: gradients3.SetRowSelection(fixedAtoms, 0.0) (System.py:292,313)
        cfg.SetTemporaryAttribute("energyTerms", terms)
        self.timings["Energy"] += time.perf_counter() - t0
        return sum(v for _, v in terms)
