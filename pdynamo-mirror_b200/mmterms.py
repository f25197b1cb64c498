"""Host-side mirror of pDynamo's bonded MM term containers, evaluated on the device by libnbabfs_b200.so.

Mirrors (labels, term / parameter layout, Energy(coordinates3, gradients3) accumulating into the gradients):
  HarmonicBondContainer      pMolecule-1.9.0/extensions/pyrex/pMolecule.HarmonicBondContainer.pyx      (also the "Urey-Bradley" container)
  HarmonicAngleContainer     pMolecule-1.9.0/extensions/pyrex/pMolecule.HarmonicAngleContainer.pyx
  FourierDihedralContainer   pMolecule-1.9.0/extensions/pyrex/pMolecule.FourierDihedralContainer.pyx
  HarmonicImproperContainer  pMolecule-1.9.0/extensions/pyrex/pMolecule.HarmonicImproperContainer.pyx
as System.Energy loops over them (pMolecule-1.9.0/pMolecule/System.py:272-318: `for mmterm in em.mmTerms: energies.append(mmterm.Energy(...))`).
All containers of a system share ONE device object (MMTermsB200): one launch evaluates every term; the per-container Energy() of the
reference interface is kept for callers that want a single term.  There is no CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CLibraryError, d_, i_

KINDS = ("Harmonic Bond", "Harmonic Angle", "Urey-Bradley", "Fourier Dihedral", "Harmonic Improper")


class _Container:
    label, natoms, kind = None, 0, -1

    def __init__(self, terms, types, parameters, active=None, label=None):
        """terms[nterms, natoms] atom indices; types[nterms] parameter index of each term; parameters: dict of per-type arrays"""
        self.terms = np.ascontiguousarray(terms, np.int32).reshape(-1, self.natoms)
        self.types = np.ascontiguousarray(types, np.int32).reshape(-1)
        self.parameters = {k: np.ascontiguousarray(v) for k, v in parameters.items()}
        self.active = None if active is None else np.ascontiguousarray(active, np.uint8)
        if label is not None:
            self.label = label
        if len(self.types) != len(self.terms):
            raise ValueError("Inconsistent term and type arrays.")

    def __len__(self):
        return len(self.terms)

    @classmethod
    def FromPerTermParameters(cls, terms, **parameters):
        """one parameter record per term (type = term index)"""
        terms = np.ascontiguousarray(terms, np.int32).reshape(-1, cls.natoms)
        return cls(terms, np.arange(len(terms), dtype=np.int32), parameters)

    def _active_ptr(self):
        return None if self.active is None else self.active.ctypes.data_as(C.c_char_p)

    def Energy(self, coordinates3, gradients3=None, device=0):
        """This container alone (the reference's per-container call)."""
        own = MMTermsB200(len(coordinates3), [self], device=device)
        return own.Energy(coordinates3, gradients3)[self.kind]


class HarmonicBondContainer(_Container):
    label, natoms, kind = "Harmonic Bond", 2, 0

    def __init__(self, terms, types, parameters, active=None, label=None, is12Interaction=True):
        super().__init__(terms, types, parameters, active, label)
        self.is12Interaction = is12Interaction
        if self.label == "Urey-Bradley":
            self.kind = 2

    def _define(self, h, status):
        p = self.parameters
        _lib.lib().HarmonicBondContainer_B200_Define(h, 1 if self.kind == 2 else 0, len(self), i_(self.terms), i_(self.types), self._active_ptr(),
                                                     len(p["eq"]), d_(np.ascontiguousarray(p["eq"], np.float64)), d_(np.ascontiguousarray(p["fc"], np.float64)), C.byref(status))


class HarmonicAngleContainer(_Container):
    label, natoms, kind = "Harmonic Angle", 3, 1

    def _define(self, h, status):
        p = self.parameters
        _lib.lib().HarmonicAngleContainer_B200_Define(h, len(self), i_(self.terms), i_(self.types), self._active_ptr(),
                                                      len(p["eq"]), d_(np.ascontiguousarray(p["eq"], np.float64)), d_(np.ascontiguousarray(p["fc"], np.float64)), C.byref(status))


class FourierDihedralContainer(_Container):
    label, natoms, kind = "Fourier Dihedral", 4, 3

    def _define(self, h, status):
        p = self.parameters
        _lib.lib().FourierDihedralContainer_B200_Define(h, len(self), i_(self.terms), i_(self.types), self._active_ptr(), len(p["fc"]),
                                                        d_(np.ascontiguousarray(p["fc"], np.float64)), i_(np.ascontiguousarray(p["period"], np.int32)),
                                                        d_(np.ascontiguousarray(p["phase"], np.float64)), C.byref(status))


class HarmonicImproperContainer(_Container):
    label, natoms, kind = "Harmonic Improper", 4, 4

    def _define(self, h, status):
        p = self.parameters
        _lib.lib().HarmonicImproperContainer_B200_Define(h, len(self), i_(self.terms), i_(self.types), self._active_ptr(),
                                                         len(p["eq"]), d_(np.ascontiguousarray(p["eq"], np.float64)), d_(np.ascontiguousarray(p["fc"], np.float64)), C.byref(status))


class MMTermsB200:
    """Device object holding the terms of a system's containers (at most one container per kind)."""

    def __init__(self, natoms, containers, device=0):
        self.natoms, self.containers = int(natoms), list(containers)
        status = C.c_int(_lib.STATUS_CONTINUE)
        self.cObject = _lib.lib().MMTerms_B200_Allocate(int(device), self.natoms, C.byref(status))
        if (not self.cObject) or status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("Unable to create the MM terms. " + _lib.last_error())
        seen = set()
        for c in self.containers:
            if c.kind in seen:
                raise ValueError("Two containers of the kind " + KINDS[c.kind] + ".")
            seen.add(c.kind)
            c._define(self.cObject, status)
            if status.value != _lib.STATUS_CONTINUE:
                raise CLibraryError("Unable to define MM terms. " + _lib.last_error())
        self.energies = np.zeros(5)

    def __del__(self):
        try:
            if self.cObject:
                h = C.c_void_p(self.cObject)
                _lib.lib().MMTerms_B200_Deallocate(C.byref(h))
                self.cObject = None
        except Exception:
            pass

    def SetStream(self, cuda_stream):
        _lib.lib().MMTerms_B200_SetStream(self.cObject, C.c_void_p(cuda_stream))

    def Energy(self, coordinates3, gradients3=None):
        """energies of the five kinds (host arrays; gradients3 is accumulated into)"""
        x = np.ascontiguousarray(coordinates3, np.float64)
        if gradients3 is not None and not (isinstance(gradients3, np.ndarray) and gradients3.dtype == np.float64 and gradients3.flags["C_CONTIGUOUS"]):
            raise TypeError("gradients3 must be a C-contiguous float64 array (it is accumulated into in place)")
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().MMTerms_B200_Energy(self.cObject, d_(x), d_(self.energies), d_(gradients3), C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("MM term evaluation failed. " + _lib.last_error())
        return self.energies.copy()

    def EnergyDevice(self, x_ptr, g_ptr):
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().MMTerms_B200_EnergyDevice(self.cObject, C.c_void_p(x_ptr), d_(self.energies), C.c_void_p(g_ptr) if g_ptr else None, C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("MM term evaluation failed. " + _lib.last_error())
        return self.energies

    def EnqueueDevice(self, x_ptr, g_ptr):
        """launch only (same stream as the caller's other work); CollectDevice() waits and returns the energies"""
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().MMTerms_B200_EnergyDeviceEnqueue(self.cObject, C.c_void_p(x_ptr), C.c_void_p(g_ptr) if g_ptr else None, C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("MM term evaluation failed. " + _lib.last_error())

    def CollectDevice(self):
        status = C.c_int(_lib.STATUS_CONTINUE)
        _lib.lib().MMTerms_B200_EnergyDeviceCollect(self.cObject, d_(self.energies), C.byref(status))
        if status.value != _lib.STATUS_CONTINUE:
            raise CLibraryError("MM term evaluation failed. " + _lib.last_error())
        return self.energies

    def LastEnergies(self):
        """the energies of the last enqueued evaluation, no synchronisation (the caller knows the stream has been synchronised since)"""
        _lib.lib().MMTerms_B200_LastEnergies(self.cObject, d_(self.energies))
        return self.energies

    def EnergyTerms(self):
        """(label, value) pairs in System.Energy's order for the containers present"""
        return [(c.label, float(self.energies[c.kind])) for c in self.containers]


def containers_from_bonded(b):
    """containers from a bonded-term dict with per-term parameters (tests/golden/dhfr_bonded.npz layout), in the order
    CHARMMPSFFileReader.ToSystem adds them: bond, angle, Urey-Bradley, dihedral, improper"""
    out = []
    if len(b.get("bonds", ())):
        out.append(HarmonicBondContainer.FromPerTermParameters(b["bonds"], eq=b["bond_eq"], fc=b["bond_fc"]))
    if len(b.get("angles", ())):
        out.append(HarmonicAngleContainer.FromPerTermParameters(b["angles"], eq=b["angle_eq"], fc=b["angle_fc"]))
    if len(b.get("ureybradleys", ())):
        ub = HarmonicBondContainer(np.asarray(b["ureybradleys"]), np.arange(len(b["ureybradleys"]), dtype=np.int32), dict(eq=b["ub_eq"], fc=b["ub_fc"]),
                                   label="Urey-Bradley", is12Interaction=False)
        out.append(ub)
    if len(b.get("dihedrals", ())):
        out.append(FourierDihedralContainer.FromPerTermParameters(b["dihedrals"], fc=b["dihedral_fc"], period=b["dihedral_period"], phase=b["dihedral_phase"]))
    if len(b.get("impropers", ())):
        out.append(HarmonicImproperContainer.FromPerTermParameters(b["impropers"], eq=b["improper_eq"], fc=b["improper_fc"]))
    return out
