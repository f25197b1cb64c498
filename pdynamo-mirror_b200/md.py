"""Velocity-Verlet dynamics with everything resident on the device.

Mirrors one Iteration of pCore-1.9.0/pCore/VelocityVerletIntegrator.py:60-81 as driven by VelocityVerletDynamics_SystemGeometry
(pMolecule-1.9.0/pMolecule/...), in Cartesian variables: x += dt v + dt^2/2 a; v += dt/2 a; E, g = NB(x); a = -100 g / m; v += dt/2 a.
Coordinates, velocities, accelerations and gradients never leave the GPU; per step the host sees the 6 NB energies, the 5 bonded
energies (when the system has bonded term containers, mmterms.py: evaluated on the device into the same gradient array) and the
kinetic energy.

LangevinDynamics is the integrator of the reference's own DHFR benchmark (benchmarks/SystemBenchmarks.py:95-101 ->
LangevinDynamics_SystemGeometry -> pCore-1.9.0/pCore/LangevinVelocityVerletIntegrator.py): CalculateIntegrationConstants (:54-115) on the
host, the Iteration (:117-150) as two kernels around the force evaluation."""
import ctypes as C
import math

import numpy as np

_KB_KJMOL = 1.3806505e-23 * 6.0221415e+23 * 1.0e-3          # kJ mol^-1 K^-1 (constants of pCore/Constants.py)


class VelocityVerletDynamics:
    """Velocity Verlet dynamics on the device; options as VelocityVerletDynamics_SystemGeometry / VelocityVerletIntegrator
    (pCore-1.9.0/pCore/VelocityVerletIntegrator.py:17-104): timeStep (ps), temperature = the start temperature of the Maxwell velocities,
    and the temperature handling temperatureScaleFrequency / temperatureScaleOption ("constant", "exponential", "linear") /
    temperatureStart / temperatureStop (velocities are scaled to the target temperature every temperatureScaleFrequency steps)."""

    def __init__(self, system, timeStep=0.001, temperature=300.0, seed=491831, device=0, temperatureScaleFrequency=0, temperatureScaleOption=None,
                 temperatureStart=None, temperatureStop=None, removeRotationTranslation=True):
        self.temperatureScaleFrequency = int(temperatureScaleFrequency)
        self.temperatureScaleOption = temperatureScaleOption.capitalize() if isinstance(temperatureScaleOption, str) else None
        if self.temperatureScaleOption not in ("Constant", "Exponential", "Linear") or self.temperatureScaleFrequency <= 0:
            self.temperatureScaleOption = None
        self.temperatureStart = float(temperature if temperatureStart is None else temperatureStart)
        self.temperatureStop = self.temperatureStart if (temperatureStop is None or self.temperatureScaleOption == "Constant") else float(temperatureStop)
        self.time, self.numberOfIterations = 0.0, 0
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
        self.L.nbb200_set_list_reuse_hint(self.h, 1)         # dynamics: ten or more calls per list
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
        # removeRotationTranslation (pMoleculeScripts/MolecularDynamics.py:27-32 -> SystemGeometryObjectiveFunction.RemoveRotationTranslation,
        # pMolecule-1.9.0/pMolecule/SystemGeometryObjectiveFunction.py:213-240): a periodic system loses its three translations (rotations only
        # without symmetry: not built), nothing is done when there are fixed atoms; the temperature is taken over the remaining degrees of freedom
        fixed = getattr(system, "fixedAtoms", None)
        self.removeTranslation = bool(removeRotationTranslation) and (fixed is None or len(fixed) == 0)
        self.totalMass = float(masses.sum())
        self.degreesOfFreedom = 3 * self.n - (3 if self.removeTranslation else 0)
        sp = system.symmetryParameters
        self.box = None if sp is None else np.ascontiguousarray(sp.box6, np.float64)
        self.energies, self.dEdM = np.zeros(6), np.zeros(9)
        self.updates = 0
        self.mmterms, self.bonded = None, np.zeros(5)
        em = system.energyModel
        if len(getattr(em, "mmTerms", ())) > 0:
            from .mmterms import MMTermsB200
            self.mmterms = MMTermsB200(self.n, em.mmTerms, device=device)
            self.mmterms.SetStream(torch.cuda.current_stream().cuda_stream)
        self.potential = self._forces(True)
        self.L.nbb200_vv_second_half(self.h, self._p(self.v), self._p(self.a), self._p(self.g), self._p(self.mass), 0.0, self._p(self.ke_dev))
        self.kinetic = float(self.ke_dev.item())

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def _forces(self, force_new):
        st = C.c_int(16)
        self.g.zero_()
        if self.mmterms is not None:
            self.mmterms.EnqueueDevice(self.x.data_ptr(), self.g.data_ptr())       # runs ahead of the NB kernels on the same stream; collected below
        self.updates += self.L.NBModelABFS_B200_UpdateDevice(self.h, self._p(self.x), self._lib.d_(self.box), 1 if force_new else 0, C.byref(st))
        self.L.NBModelABFS_B200_MMMMEnergyDevice(self.h, self._lib.d_(self.energies), self._p(self.g), self._lib.d_(self.dEdM), C.byref(st))
        if st.value != 16:
            raise RuntimeError("NB call failed: " + self._lib.last_error())
        if self.mmterms is not None:
            self.bonded = self.mmterms.CollectDevice()
            return float(self.energies.sum() + self.bonded.sum())
        return float(self.energies.sum())

    def _first_half(self):
        self.L.nbb200_vv_first_half(self.h, self._p(self.x), self._p(self.v), self._p(self.a), self.dt)

    def TargetTemperature(self, time, total_time):
        """VelocityVerletIntegrator.TargetTemperature / TemperatureHandlingOptions (:85-104)"""
        t0, t1 = self.temperatureStart, self.temperatureStop
        if self.temperatureScaleOption == "Exponential":
            return max(0.0, t0 * math.exp(math.log(t1 / t0) / total_time * time))
        if self.temperatureScaleOption == "Linear":
            return max(0.0, t0 + (t1 - t0) / total_time * time)
        return max(0.0, t0)

    def _second_half_dt(self):
        return self.dt

    def _langevin_factors(self):
        return None

    def RunNative(self, steps, updateFrequency=0):
        """The same loop inside the library (nbb200_md_run): no interpreter between the launches of a step."""
        pot, kin = np.zeros(steps), np.zeros(steps)
        st = C.c_int(16)
        fac = self._langevin_factors()
        first = getattr(self, "iteration", 0) + 1
        self.updates += self.L.nbb200_md_run(self.h, C.c_void_p(self.mmterms.cObject) if self.mmterms is not None else None, int(steps), int(updateFrequency),
                                             self._p(self.x), self._p(self.v), self._p(self.a), self._p(self.g), self._p(self.mass), self._lib.d_(self.box),
                                             self.dt, self._lib.d_(fac) if fac is not None else None, self._second_half_dt(), C.c_ulonglong(getattr(self, "seed", 0)),
                                             C.c_ulonglong(first), self._p(self.ke_dev), self._lib.d_(pot), self._lib.d_(kin), self._lib.d_(self.energies),
                                             self._lib.d_(self.bonded), C.byref(st))
        if st.value != 16:
            raise RuntimeError("MD run failed: " + self._lib.last_error())
        if hasattr(self, "iteration"):
            self.iteration += steps
        self.numberOfIterations += steps
        self.time += steps * self.dt
        if steps > 0:
            self.potential, self.kinetic = float(pot[-1]), float(kin[-1])
        return list(zip(pot.tolist(), kin.tolist()))

    def Run(self, steps, updateFrequency=0, log=None, native=None):
        """steps integration steps; updateFrequency > 0 forces a list rebuild every that many steps (0: the reference's displacement
        heuristic only).  Returns the list of (potential, kinetic) per step.  native (default: when there is neither logging nor
        temperature scaling): run the loop inside the library (RunNative).

        The loop is pipelined: per step the host waits ONCE, for the list-update decision (NBModelABFS_B200_UpdateDevice); the energy call is
        deferred (NBModelABFS_B200_MMMMEnergyDeviceDeferred), the bonded energies and the kinetic energy are copied to page-locked memory in
        stream order, and the numbers of step k are picked up after the decision of step k + 1 (or the final flush)."""
        scaling = self.temperatureScaleOption is not None and self.temperatureScaleFrequency < steps
        if native is None:
            native = log is None and not scaling
        if native and steps > 0:
            return self.RunNative(steps, updateFrequency)
        out = []
        st = C.c_int(16)
        if getattr(self, "_ke_host", None) is None:
            self._ke_host = self._lib.pinned_array((2,))
        ke_ptr = [C.c_void_p(self._ke_host.ctypes.data), C.c_void_p(self._ke_host.ctypes.data + 8)]
        dt2 = self._second_half_dt()

        t_begin, total_time = self.time, steps * self.dt

        def harvest(k, ke_scale=1.0):
            if self.mmterms is not None:
                self.bonded = self.mmterms.LastEnergies()
            self.potential = float(self.energies.sum() + (self.bonded.sum() if self.mmterms is not None else 0.0))
            self.kinetic = float(self._ke_host[k & 1]) * ke_scale
            out.append((self.potential, self.kinetic))
            if log is not None and (k + 1) % 100 == 0:
                log("step %d: potential %.4f kinetic %.4f total %.4f temperature %.2f" % (k + 1, self.potential, self.kinetic, self.potential + self.kinetic,
                                                                                       2.0 * self.kinetic / (self.degreesOfFreedom * _KB_KJMOL)))
        self.L.nbb200_set_gradient_overwrite(self.h, 1)        # the NB term sets g (no zero fill), the bonded terms then accumulate
        try:
            for k in range(steps):
                self._first_half()
                force_new = updateFrequency > 0 and (k + 1) % updateFrequency == 0
                self.updates += self.L.NBModelABFS_B200_UpdateDevice(self.h, self._p(self.x), self._lib.d_(self.box), 1 if force_new else 0, C.byref(st))
                # the decision synchronised the stream: step k - 1 is complete, its NB energies are in self.energies.  Enqueue first, read after.
                self.L.NBModelABFS_B200_MMMMEnergyDeviceDeferred(self.h, self._lib.d_(self.energies), self._p(self.g), self._lib.d_(self.dEdM), C.byref(st))
                if k > 0 and len(out) < k:
                    harvest(k - 1)
                if self.mmterms is not None:
                    self.mmterms.EnqueueDevice(self.x.data_ptr(), self.g.data_ptr())
                self.L.nbb200_vv_second_half(self.h, self._p(self.v), self._p(self.a), self._p(self.g), self._p(self.mass), dt2, self._p(self.ke_dev))
                self.L.nbb200_copy_to_host_async(self.h, self._p(self.ke_dev), ke_ptr[k & 1], 8)
                self.numberOfIterations += 1
                self.time += self.dt
                if st.value != 16:
                    raise RuntimeError("NB call failed: " + self._lib.last_error())
                if scaling and self.numberOfIterations % self.temperatureScaleFrequency == 0:
                    # temperature scaling needs this step's kinetic energy: one extra host wait on these (rare) steps
                    self.L.nbb200_flush(self.h, C.byref(st))
                    temp = 2.0 * float(self._ke_host[k & 1]) / (self.degreesOfFreedom * _KB_KJMOL)
                    scale = self.TargetTemperature(self.time - t_begin, total_time) / temp if temp > 0.0 else 1.0
                    self.v.mul_(math.sqrt(scale))
                    harvest(k, scale)
        finally:
            self.L.nbb200_set_gradient_overwrite(self.h, 0)
        if steps > 0 and len(out) < steps:
            self.L.nbb200_flush(self.h, C.byref(st))
            harvest(steps - 1)
        return out


def langevin_constants(timeStep, collisionFrequency, temperature, forceSeries=None):
    """The integration constants of LangevinVelocityVerletIntegrator.CalculateIntegrationConstants (pCore-1.9.0/pCore/LangevinVelocityVerletIntegrator.py:54-115):
    returns (factors7, facV3) with factors7 = (facR1, facR2, facV1, facV2, sdR, sdV1, sdV2), the random-term factors already multiplied by
    sqrt(kT) in dynamics units (amu A^2 ps^-2).  Exponential formulae for collisionFrequency * timeStep > 0.009, series expansions (valid to
    fact^5) below; forceSeries (testing) overrides the choice."""
    dt = float(timeStep)
    fact = collisionFrequency * dt
    series = (fact <= 0.009) if forceSeries is None else bool(forceSeries)
    if not series:
        c0 = math.exp(-fact)
        c1 = (1.0 - c0) / fact
        c2 = (1.0 - c1) / fact
        sdR = math.sqrt(dt ** 2 * (2.0 - (3.0 - 4.0 * c0 + c0 * c0) / fact) / fact)
        sdV = math.sqrt(1.0 - c0 * c0)
        cRV1 = dt * (1.0 - c0) ** 2 / (fact * sdR * sdV)
        cRV2 = math.sqrt(1.0 - cRV1 * cRV1)
    else:
        c0 = 1.0 - fact + fact ** 2 / 2.0 - fact ** 3 / 6.0 + fact ** 4 / 24.0 - fact ** 5 / 120.0
        c1 = 1.0 - fact / 2.0 + fact ** 2 / 6.0 - fact ** 3 / 24.0 + fact ** 4 / 120.0 - fact ** 5 / 720.0
        c2 = 0.5 - fact / 6.0 + fact ** 2 / 24.0 - fact ** 3 / 120.0 + fact ** 4 / 720.0 - fact ** 5 / 5040.0
        sdR = (2.0 - 3.0 * fact / 4.0 + 67.0 * fact ** 2 / 320.0 - 119.0 * fact ** 3 / 2560.0) * math.sqrt(fact / 6.0) * dt
        sdV = (2.0 - fact + 5.0 * fact ** 2 / 12.0 - fact ** 3 / 8.0 + 79.0 * fact ** 4 / 2880.0) * math.sqrt(fact / 2.0)
        cRV1 = (3.0 / 2.0 - 3.0 * fact / 16.0 - 51.0 * fact ** 2 / 1280.0 + 17.0 * fact ** 3 / 2048.0 + 40967.0 * fact ** 4 / 11468800.0 -
                57203.0 * fact ** 5 / 91750400.0) / math.sqrt(3.0)
        cRV2 = (0.5 + 3.0 * fact / 16.0 - 9.0 * fact ** 2 / 1280.0 - 109.0 * fact ** 3 / 10240.0 + 10077.0 * fact ** 4 / 11468800.0 +
                14887.0 * fact ** 5 / 18350080.0)
    kT = math.sqrt(100.0 * _KB_KJMOL * temperature)          # sqrt(K -> amu A^2 ps^-2), SystemGeometryObjectiveFunction.TemperatureConversionFactor
    factors = np.array([c1 * dt, c2 * dt ** 2, c0, (c1 - c2) * dt, sdR * kT, cRV1 * sdV * kT, cRV2 * sdV * kT], dtype=np.float64)
    return factors, c2 * dt


class LangevinDynamics(VelocityVerletDynamics):
    """Langevin velocity Verlet dynamics on the device (pCore-1.9.0/pCore/LangevinVelocityVerletIntegrator.py); options as
    LangevinDynamics_SystemGeometry: collisionFrequency (ps^-1), temperature (K), timeStep (ps)."""

    def __init__(self, system, timeStep=0.001, temperature=300.0, collisionFrequency=25.0, seed=491831, device=0, removeRotationTranslation=True):
        if collisionFrequency <= 0.0 or temperature < 0.0:
            raise ValueError("Invalid temperature handling options.")
        super().__init__(system, timeStep=timeStep, temperature=temperature, seed=seed, device=device, removeRotationTranslation=removeRotationTranslation)
        self.collisionFrequency, self.temperature, self.seed, self.iteration = float(collisionFrequency), float(temperature), int(seed), 0
        self.CalculateIntegrationConstants()
        # RandomForces: ApplyLinearConstraints on both random vectors (LangevinVelocityVerletIntegrator.py:139-149)
        self.L.nbb200_set_langevin_constraints(self.h, 1 if self.removeTranslation else 0, self.totalMass)

    def CalculateIntegrationConstants(self):
        """LangevinVelocityVerletIntegrator.CalculateIntegrationConstants (:54-115)"""
        self.factors, self.facV3 = langevin_constants(self.dt, self.collisionFrequency, self.temperature)

    def _first_half(self):
        self.iteration += 1
        self.L.nbb200_langevin_first_half(self.h, self._p(self.x), self._p(self.v), self._p(self.a), self._p(self.mass), self._lib.d_(self.factors),
                                          C.c_ulonglong(self.seed), C.c_ulonglong(self.iteration))

    def _second_half_dt(self):
        return 2.0 * self.facV3                # v += facV3 a  (LangevinVelocityVerletIntegrator.py:133)

    def _langevin_factors(self):
        return self.factors
