"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/nbabfs_b200.h
declares, has no torch / oracle dependency, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nbabfs_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:NBModelABFS|PairListGenerator|PairwiseInteractionABFS|nbb200|MMTerms_B200|HarmonicBondContainer|HarmonicAngleContainer|FourierDihedralContainer|HarmonicImproperContainer)\w*)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    from pdynamo_mirror_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run python __graft_entry__.py (build) first"
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), "missing export: " + name
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)


def test_reference_interface_shim_exports_the_reference_names():
    """csrc/compat_shim.c replaces the translation units NBModelABFS.c / NBModelABFSState.c: the shim build (oracle/_ref/libshim_nbabfs.so, made
    where the reference headers exist) must define every function those units export (pM/cinclude/NBModelABFS.h:39-46,
    NBModelABFSState.h:132-161) and resolve the computation in libnbabfs_b200.so, not in the reference."""
    path = os.path.join(ROOT, "oracle", "_ref", "libshim_nbabfs.so")
    if not os.path.exists(path):
        pytest.skip("shim library not built (no reference tree here)")
    syms = subprocess.run(["nm", "-D", path], capture_output=True, text=True).stdout
    defined = set(re.findall(r" T (\w+)", syms))
    undefined = set(re.findall(r" U (\w+)", syms))
    for name in ("NBModelABFS_Allocate", "NBModelABFS_Clone", "NBModelABFS_Deallocate", "NBModelABFS_Update", "NBModelABFS_MMMMEnergy", "NBModelABFS_QCMMEnergyLJ",
                 "NBModelABFS_QCMMPotentials", "NBModelABFS_QCMMGradients", "NBModelABFSState_Allocate", "NBModelABFSState_Deallocate", "NBModelABFSState_SetUp",
                 "NBModelABFSState_SetUpCentering", "NBModelABFSState_Initialize", "NBModelABFSState_InitializeCoordinates3", "NBModelABFSState_GridInitialize",
                 "NBModelABFSState_GridFinalize", "NBModelABFSState_StatisticsAccumulate", "NBModelABFSState_StatisticsInitialize"):
        assert name in defined, name
    assert {"NBModelABFSState_B200_SetUp", "NBModelABFS_B200_Update", "NBModelABFS_B200_MMMMEnergy"} <= undefined
    # the reference's own pair-list generator / pair kernels are not behind these entry points any more
    text = open(os.path.join(ROOT, "pdynamo-mirror_b200", "csrc", "compat_shim.c")).read()
    assert "PairListGenerator_SelfPairList" not in text and "PairwiseInteractionABFS_MMMMEnergy" not in text


def test_library_is_self_contained(pkg):
    from pdynamo_mirror_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "ref_nbabfs" not in out


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "pdynamo-mirror_b200")):
        if os.path.basename(base) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "refnb" not in text and "liboracle" not in text and "nbabfs_oracle" not in text, f


def test_make_factors_matches_oracle(pkg, orc):
    pw = pkg.PairwiseInteractionABFS(dampingCutoff=0.5, innerCutoff=8.0, outerCutoff=12.0)
    assert np.array_equal(pw.MakeFactors(), orc.make_factors(0.5, 8.0, 12.0))
    pw = pkg.PairwiseInteractionABFS(dampingCutoff=1.0, innerCutoff=6.0, outerCutoff=9.0)
    assert np.array_equal(pw.MakeFactors(), orc.make_factors(1.0, 6.0, 9.0))


def test_spline_tables_match_oracle(pkg, orc):
    """host helper PairwiseInteractionABFS_B200_MakeSpline (what the device state uploads) against the restatement (itself bit-exact
    against the compiled reference, tests/test_oracle.py): abscissae, ordinates and second derivatives identical"""
    for cut in [(0.5, 8.0, 12.0, 50), (1.0, 6.0, 9.0, 20), (0.5, 8.0, 12.0, 333)]:
        pw = pkg.PairwiseInteractionABFS(dampingCutoff=cut[0], innerCutoff=cut[1], outerCutoff=cut[2], splinePointDensity=cut[3], useAnalyticForm=False)
        tables = pw.MakeSplines()
        for which, name in enumerate(("electrostatic", "lennardJonesA", "lennardJonesB")):
            for a, b in zip(tables[name], orc.make_spline(which, *cut)):
                assert np.array_equal(a, b), (cut, name)
        au = pw.MakeSplines(lennardJones=False, useAtomicUnits=True)            # the spline of the QC/MM and QC/QC interactions
        assert list(au) == ["electrostatic"]
        for a, b in zip(au["electrostatic"], orc.make_spline(3, *cut)):
            assert np.array_equal(a, b), (cut, "electrostatic, atomic units")
    with pytest.raises(ValueError):
        pkg.PairwiseInteractionABFS(splinePoints=3)
    with pytest.raises(ValueError):
        pkg.PairwiseInteractionABFS(electrostaticModel="Point/Point")
    state = pkg.PairwiseInteractionABFS(useAnalyticForm=False, splinePointDensity=20).__getstate__()
    assert state["useAnalyticForm"] is False and state["splinePointDensity"] == 20 and state["electrostaticModel"] == "Delta/Delta"


def test_no_cpu_fallback(pkg):
    from pdynamo_mirror_b200 import _lib
    if _lib.lib().nbb200_device_count() > 0:
        pytest.skip("a GPU is present")
    system = pkg.System.FromWorkload(pkg.workloads.WORKLOADS["w216"]())
    system.DefineNBModel(pkg.NBModelABFS())
    with pytest.raises(pkg.CLibraryError, match="Unable to create NB state"):
        system.Energy(doGradients=True)
    with pytest.raises(pkg.CLibraryError):
        pkg.PairListGenerator(cutoff=5.0).SelfPairListFromCoordinates3(np.zeros((4, 3)))
