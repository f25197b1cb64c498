"""Velocity-Verlet dynamics with everything resident on the device.

Mirrors one Iteration of pCore-1.9.0/pCore/VelocityVerletIntegrator.py:60-81 as driven by VelocityVerletDynamics_SystemGeometry
(pMolecule-1.9.0/pMolecule/...), in Cartesian variables: x += dt v + dt^2/2 a; v += dt/2 a; E, g = NB(x); a = -100 g / m; v += dt/2 a.
Coordinates, velocities, accelerations and gradients never leave the GPU; per step the host sees the 6 energies and the kinetic
energy.  Only the NB term exists in this repository, so this is a complete force field only for bond-free systems
(workloads.ionic_fluid)."""
import ctypes as C

import numpy as np

_KB_KJMOL = 1.3806505e-23 * 6.0221415e+23 * 1.0e-3          # kJ mol^-1 K^-1 (constants of pCore/Constants.py)


class VelocityVerletDynamics:
    def __init__(self, system, timeStep=0.001, temperature=300.0, seed=491831, device=0):
        import torch
        from . import _lib
        self.torch, self.L, self._lib = torch, _lib.lib(), _lib
        self.system, self.dt = system, float(timeStep)
        cfg = system.configuration
        system.Energy(doGradients=True)                      # creates the NB state through the plugin surface
        self.state = cfg.nbState
        self.h = self.state.cObject
        self.n = self.state.n
        dev = "cuda:%d" % device
        self.L.nbb200_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        masses = np.asarray(getattr(system, "masses", None) if getattr(system, "masses", None) is not None else np.ones(self.n), np.float64)
        self.mass = torch.from_numpy(masses).to(dev)
        self.x = torch.from_numpy(np.ascontiguousarray(system.coordinates3, np.float64)).to(dev)
        self.g = torch.zeros_like(self.x)
        self.a = torch.zeros_like(self.x)
        rng = np.random.Generator(np.random.PCG64(seed))
        sigma = np.sqrt(_KB_KJMOL * temperature / (0.01 * masses))[:, None]           # A/ps: (1/2) 0.01 m v^2 = (1/2) kT per degree of freedom
        v = rng.standard_normal((self.n, 3)) * sigma
        v -= (v * masses[:, None]).sum(0) / masses.sum()                               # no net momentum
        self.v = torch.from_numpy(v).to(dev)
        self.ke_dev = torch.zeros(1, dtype=torch.float64, device=dev)
        sp = system.symmetryParameters
        self.box = None if sp is None else np.ascontiguousarray(sp.box6, np.float64)
        self.energies, self.dEdM = np.zeros(6), np.zeros(9)
        self.updates = 0
        self.potential = self._forces(True)
        self.L.nbb200_vv_second_half(self.h, self._p(self.v), self._p(self.a), self._p(self.g), self._p(self.mass), 0.0, self._p(self.ke_dev))
        self.kinetic = float(self.ke_dev.item())

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def _forces(self, force_new):
        st = C.c_int(16)
        self.g.zero_()
        self.updates += self.L.NBModelABFS_B200_UpdateDevice(self.h, self._p(self.x), self._lib.d_(self.box), 1 if force_new else 0, C.byref(st))
        self.L.NBModelABFS_B200_MMMMEnergyDevice(self.h, self._lib.d_(self.energies), self._p(self.g), self._lib.d_(self.dEdM), C.byref(st))
        if st.value != 16:
            raise RuntimeError("NB call failed: " + self._lib.last_error())
        return float(self.energies.sum())

    def Run(self, steps, updateFrequency=0, log=None):
        """steps velocity-Verlet steps; updateFrequency > 0 forces a list rebuild every that many steps (0: the reference's displacement
        heuristic only).  Returns the list of (potential, kinetic) per step."""
        out = []
        for k in range(steps):
            self.L.nbb200_vv_first_half(self.h, self._p(self.x), self._p(self.v), self._p(self.a), self.dt)
            self.potential = self._forces(updateFrequency > 0 and (k + 1) % updateFrequency == 0)
            self.L.nbb200_vv_second_half(self.h, self._p(self.v), self._p(self.a), self._p(self.g), self._p(self.mass), self.dt, self._p(self.ke_dev))
            self.kinetic = float(self.ke_dev.item())
            out.append((self.potential, self.kinetic))
            if log is not None and (k + 1) % 100 == 0:
                log("step %d: potential %.4f kinetic %.4f total %.4f" % (k + 1, self.potential, self.kinetic, self.potential + self.kinetic))
        return out
