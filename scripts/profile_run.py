"""Short driver for ncu captures: a few rebuild + energy steps on one workload through the C-ABI (device-resident)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "jac"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.make_workload(name)
m = bench.DeviceModel(w, 0)
for _ in range(steps):
    m.step(rebuild=True)
torch.cuda.synchronize()
print(name, m.state.Counters(), m.state.Timings())
