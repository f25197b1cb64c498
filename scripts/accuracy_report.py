"""Print the measured errors of the CUDA path against the golden outputs / the oracle for a few systems (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pdynamo_mirror_b200 as p
import oracle
names = sys.argv[1:] or ["w216", "w216_triclinic", "bala", "jac", "dhfr", "crystal_GLYGLY", "crystal_ALAALA"]
for name in names:
    w = p.workloads.WORKLOADS[name]()
    s = p.System.FromWorkload(w); s.DefineNBModel(p.NBModelABFS()); s.Energy(doGradients=True)
    ref = oracle.OracleNB(w).energy(force_new=True)
    e, g = s.configuration.nbState.energies, s.configuration.gradients3
    re, rg = ref["energies"], ref["grad"]
    dm = s.configuration.symmetryParameterGradients.dEdM
    print("%-16s E rel %.2e | per-term max |dE|/sum|E| %.2e | grad rel RMS %.2e | dEdM rel %.2e" % (
        name, abs(e.sum() - re.sum()) / abs(re.sum()), np.abs(e - re).max() / np.abs(re).sum(),
        np.sqrt(((g - rg) ** 2).mean()) / np.sqrt((rg ** 2).mean()), np.linalg.norm(dm - ref["dEdM"]) / max(1e-300, np.linalg.norm(ref["dEdM"]))), flush=True)
