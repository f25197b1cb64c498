"""debug helper: drive the reference-interface shim twice on one crystal"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pdynamo_mirror_b200 as p
import refnb
name = sys.argv[1] if len(sys.argv) > 1 else "crystal_GLYGLY"
maker, opts, _ = p.workloads.GOLDEN_CASES[name]
w = maker()
r = refnb.RefNB(w, omp="shim", **opts)
out = r.energy(force_new=True)
print("first", out["energies"], out["updated"], flush=True)
u = p.workloads.lcg_uniform(7, 3 * w["n"]).reshape(-1, 3)
for amp in (0.2, 0.9):
    x = w["xyz"] + (2 * u - 1) * amp
    a = r.energy(xyz=x)
    print(amp, a["energies"], a["updated"], flush=True)
r.close()
print("done")
