"""Kernel times of ONE rank's share of the M1 box on one GPU (development aid): this process plays rank r of R (restricted sort, no peers)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
r, R = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3/8").split("/"))
w = bench.make_workload("m1")
m = bench.DeviceModel(w, 0)
m.L.nbb200_set_partition(m.h, r, R)
m.L.nbb200_set_restricted_sort(m.h, 1)
m.step(rebuild=True); torch.cuda.synchronize()
ms = bench.timed_steps(torch, lambda: m.step(rebuild=True), 10, 3, lambda: None)
fk, lb = [], []
for _ in range(10):
    m.step(rebuild=True); t = m.state.Timings(); fk.append(t["tileForces"]); lb.append(t["listRebuild"])
c = m.state.Counters()
print("rank %d/%d COOP=%s: step %.3f ms, force kernel %.3f ms, rebuild kernels %.3f ms, tiles %d items %d" % (r, R, os.environ.get("NBB200_COOP"), ms, statistics.mean(fk), statistics.mean(lb), c["tiles"], c["workItems"]))
