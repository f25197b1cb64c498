"""CPU tests of the host-side mirror of the reference plugin surface (option handling, pickling, call order)."""
import copy
import pickle

import numpy as np
import pytest


def test_defaults_match_reference(pkg):
    nb = pkg.NBModelABFS()
    # NBModelABFS_Allocate defaults: pMolecule-1.9.0/extensions/csource/NBModelABFS.c:28-37
    assert (nb.dampingCutoff, nb.innerCutoff, nb.outerCutoff, nb.listCutoff, nb.dielectric) == (0.5, 8.0, 12.0, 13.5, 1.0)
    assert nb.checkForInverses is True and nb.useCentering is False and nb.label == "ABFS"
    g = nb.generator       # pMolecule.NBModelABFS.pyx:59-71
    assert (g.cutoff, g.cutoffCellSizeFactor, g.minimumCellExtent, g.minimumPoints, g.sortIndices, g.useGridByCell) == (13.5, 0.5, 2, 500, False, True)
    assert g.cellSize == 6.75
    pw = nb.mmmmPairwiseInteraction
    assert (pw.dampingCutoff, pw.innerCutoff, pw.outerCutoff) == (0.5, 8.0, 12.0)


def test_set_options_validation(pkg):
    nb = pkg.NBModelABFS(innerCutoff=6.0, outerCutoff=9.0, listCutoff=10.5)
    assert nb.generator.cutoff == 10.5
    with pytest.raises(ValueError, match="Invalid options: bogus"):
        nb.SetOptions(bogus=1)
    with pytest.raises(TypeError, match="Invalid cutoff values"):
        pkg.NBModelABFS(innerCutoff=13.0)
    with pytest.raises(TypeError, match="Unrecognized QC/MM coupling"):
        pkg.NBModelABFS(qcmmCoupling="nope")
    with pytest.raises(AttributeError):
        nb.dielectric = 2.0          # read-only properties, pMolecule.NBModelABFS.pyx:292-302


def test_pickle_roundtrip_is_options_only(pkg):
    nb = pkg.NBModelABFS(dielectric=2.0, listCutoff=14.0, electrostaticScale14=0.5)
    for clone in (pickle.loads(pickle.dumps(nb)), copy.deepcopy(nb)):
        assert clone.dielectric == 2.0 and clone.listCutoff == 14.0 and clone.electrostaticScale14 == 0.5
        assert clone.generator.cutoff == 14.0


def test_system_call_order_and_state_lifetime(pkg):
    calls = []

    class Spy(pkg.NBModel):
        def SetUp(self, *a, **k):
            calls.append("SetUp")
            a[9].nbState = "state"

        def Energy(self, configuration):
            calls.append("Energy")
            return [("MM/MM Elect.", -1.0), ("MM/MM LJ", 0.25)]

    s = pkg.System.FromWorkload(pkg.workloads.WORKLOADS["w216"]())
    spy = Spy()
    s.DefineNBModel(spy)
    assert s.Energy(doGradients=True) == -0.75
    assert calls == ["SetUp", "Energy"]
    assert s.configuration.gradients3.shape == (648, 3) and hasattr(s.configuration, "symmetryParameterGradients")
    s.configuration.ClearTemporaryAttributes()
    assert hasattr(s.configuration, "nbState") and not hasattr(s.configuration, "gradients3")   # nbState is persistent
    s.DefineNBModel(Spy())                                                                       # Clear() drops the old state
    assert not hasattr(s.configuration, "nbState")


def test_workloads_are_deterministic_and_sized(pkg):
    w = pkg.workloads.WORKLOADS["jac"]()
    assert w["n"] == 23558 and w["ntypes"] == 35 and len(w["pairs14"]) == 2486 and abs(w["charges"].sum() + 11.0) < 1e-9
    w2 = pkg.workloads.WORKLOADS["jac"]()
    assert np.array_equal(w["xyz"], w2["xyz"])
    w = pkg.workloads.WORKLOADS["w216"]()
    assert w["n"] == 648 and w["box"][0] == 18.641
    a = pkg.workloads.lcg_uniform(12345, 5)
    s, ref = 12345, []
    for _ in range(5):
        s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        ref.append((s >> 11) / float(1 << 53))
    assert np.array_equal(a, np.array(ref))


def test_dhfr_workloads_nb_only_and_complete(pkg):
    """'dhfr' is the NB-only system of the NB parity tests; 'dhfr_mm' adds the bonded containers in the order CHARMMPSFFileReader.ToSystem does"""
    a, b = pkg.workloads.WORKLOADS["dhfr"](), pkg.workloads.WORKLOADS["dhfr_mm"]()
    assert "bonded" not in a and len(pkg.System.FromWorkload(a).energyModel.mmTerms) == 0
    labels = [c.label for c in pkg.System.FromWorkload(b).energyModel.mmTerms]
    assert labels == ["Harmonic Bond", "Harmonic Angle", "Urey-Bradley", "Fourier Dihedral", "Harmonic Improper"]
    assert [len(c) for c in pkg.System.FromWorkload(b).energyModel.mmTerms] == [23592, 11584, 2117, 7000, 418]
    assert abs(b["masses"].sum() - a["masses"].sum()) == 0.0 and len(a["masses"]) == a["n"]


def test_langevin_integration_constants(pkg):
    """CalculateIntegrationConstants (pCore-1.9.0/pCore/LangevinVelocityVerletIntegrator.py:54-115): the exponential formulae and the series
    expansions (valid to fact^5) are two transcriptions of the same functions -- they must agree where both are accurate; plus the limits
    (no friction: velocity Verlet, no noise) and the fluctuation-dissipation balance of the velocity update."""
    from pdynamo_mirror_b200.md import langevin_constants, _KB_KJMOL
    dt, T = 0.001, 300.0
    for gamma in (2.0, 5.0, 9.0, 20.0):                       # fact = 0.002 ... 0.02: both branches are accurate to better than 1e-9 relative
        a, va = langevin_constants(dt, gamma, T, forceSeries=False)
        b, vb = langevin_constants(dt, gamma, T, forceSeries=True)
        # the exponential form of sdR / cRV1 cancels badly at small fact (that is why the reference switches): compare at the accuracy it has
        tol = np.array([1e-10, 1e-9, 1e-12, 1e-9, 3e-6, 3e-6, 3e-5])
        assert np.all(np.abs(a - b) <= tol * np.abs(b)), (gamma, (a - b) / b)
        assert abs(va - vb) <= 1e-9 * abs(vb)
    f, v3 = langevin_constants(dt, 1.0e-9, 0.0)                # no friction, no noise
    assert np.allclose(f[:4], [dt, 0.5 * dt * dt, 1.0, 0.5 * dt], rtol=1e-8) and np.all(f[4:] == 0.0) and abs(v3 - 0.5 * dt) < 1e-12
    # stationary velocity variance: v' = c0 v + noise with variance sdV1^2 + sdV2^2 must keep <v^2> = kT (per unit mass, dynamics units)
    for gamma in (1.0, 25.0, 200.0):
        f, _ = langevin_constants(dt, gamma, T)
        kT = 100.0 * _KB_KJMOL * T
        assert abs((f[5] ** 2 + f[6] ** 2) - kT * (1.0 - f[2] ** 2)) <= 1e-9 * kT
    # the reference benchmark's values (25 ps^-1, 1 fs): exponential branch
    f, _ = langevin_constants(0.001, 25.0, 300.0)
    assert 0.975 < f[2] < 0.9754 and f[4] > 0 and f[5] > 0 and f[6] > 0


def test_host_row_helpers_of_the_multi_gpu_host_path(pkg):
    """csrc/host_rows.cpp (no GPU involved): the threaded streaming-store copy into staging memory, the threaded add, and the row gather /
    scatter-add companions -- sizes below and above the threading threshold, a destination that is not 16-byte aligned, repeated calls on
    the persistent pool."""
    import ctypes as C
    from pdynamo_mirror_b200 import _lib
    L = _lib.lib()
    rng = np.random.Generator(np.random.PCG64(7))
    for m in (5, 32768 + 3, 3 * 250001):
        a = rng.random(m)
        b = np.zeros(m + 1)[1:]                              # 8 mod 16: the head of the streaming copy
        for _ in range(3):
            L.nbb200_host_copy(C.c_void_p(b.ctypes.data), C.c_void_p(a.ctypes.data), m)
        assert np.array_equal(a, b)
        g = np.ones(m)
        for _ in range(4):
            L.nbb200_host_add(C.c_void_p(g.ctypes.data), C.c_void_p(a.ctypes.data), m)
        assert np.allclose(g, 1.0 + 4.0 * a, rtol=0, atol=1e-14)
    n = 100000
    x = rng.random((n, 3))
    ids = rng.permutation(n).astype(np.int32)[:70000]
    out = np.zeros((len(ids), 3))
    L.nbb200_host_gather_rows(C.c_void_p(x.ctypes.data), C.c_void_p(ids.ctypes.data), len(ids), C.c_void_p(out.ctypes.data))
    assert np.array_equal(out, x[ids])
    acc = np.zeros((n, 3))
    L.nbb200_host_scatter_add_rows(C.c_void_p(acc.ctypes.data), C.c_void_p(ids.ctypes.data), len(ids), C.c_void_p(out.ctypes.data))
    ref = np.zeros((n, 3)); ref[ids] = x[ids]
    assert np.array_equal(acc, ref)
