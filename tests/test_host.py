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
