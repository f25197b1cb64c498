import sys, os, statistics
sys.path.insert(0, os.getcwd())
import torch, bench
w = bench.make_workload("dhfr")
m = bench.DeviceModel(w, 0)
m.step(rebuild=True); torch.cuda.synchronize()
ms = bench.timed_steps(torch, lambda: m.step(rebuild=True), 30, 5, lambda: None)
fk, lb = [], []
for _ in range(20):
    m.step(rebuild=True); t = m.state.Timings(); fk.append(t["tileForces"]); lb.append(t["listRebuild"])
ms_nr = bench.timed_steps(torch, lambda: m.step(rebuild=False), 100, 10, lambda: None)
c = m.state.Counters()
print("CHUNK=%s SPLIT=%s: rebuild call %.1f us, no-rebuild call %.1f us, force kernel %.1f us, rebuild kernels %.1f us, tiles %d items %d" % (
    os.environ.get("NBB200_CHUNK"), os.environ.get("NBB200_SPLIT_TARGET"), 1e3 * ms, 1e3 * ms_nr, 1e3 * statistics.mean(fk), 1e3 * statistics.mean(lb), c["tiles"], c["workItems"]))
