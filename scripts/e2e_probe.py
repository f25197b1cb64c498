import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pdynamo_mirror_b200 as p
w = p.workloads.WORKLOADS["m1"]()
s = p.System.FromWorkload(w); s.DefineNBModel(p.NBModelABFS(updateFrequency=1))
for _ in range(3): s.Energy(doGradients=True)
s.timings = {k: 0.0 for k in s.timings}
t0 = time.perf_counter()
for _ in range(10): s.Energy(doGradients=True)
tot = (time.perf_counter() - t0) / 10
print("e2e %.3f ms: NB Set Up %.3f, NB Evaluation %.3f, other (zeroing etc.) %.3f" % (1e3 * tot, 100 * s.timings["NB Set Up"], 100 * s.timings["NB Evaluation"], 1e3 * tot - 100 * (s.timings["NB Set Up"] + s.timings["NB Evaluation"])))
g = s._gradients
t0 = time.perf_counter()
for _ in range(10): g.fill(0.0)
print("fill %.3f ms" % ((time.perf_counter() - t0) * 100))
