"""Short driver for ncu captures: a few rebuild + energy steps on one workload through the C-ABI (device-resident).

    python scripts/profile_run.py <workload> <steps> [analytic|spline|md|norebuild]

spline: the same steps with the interaction in its spline form (PairwiseInteractionABFS_B200_SetInteractionForm);
norebuild: one rebuild step, then <steps> calls on slightly displaced coordinates that keep the lists (rolling prune + inner pool);
md:     <steps> Langevin velocity-Verlet steps with bonded terms on the device (workload dhfr_mm)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "jac"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "analytic"
w = bench.make_workload(name)
if mode == "md":
    import pdynamo_mirror_b200 as p
    sysm = p.System.FromWorkload(w)
    sysm.DefineNBModel(p.NBModelABFS())
    md = p.md.LangevinDynamics(sysm)
    md.Run(steps)
    torch.cuda.synchronize()
    print(name, "md", md.potential, md.kinetic, md.updates)
else:
    m = bench.DeviceModel(w, 0)
    if os.environ.get("PROFILE_PARTITION"):                 # "r/R": this process plays rank r of R (restricted sort), no peers: kernel times of one rank
        r, R = (int(v) for v in os.environ["PROFILE_PARTITION"].split("/"))
        m.L.nbb200_set_partition(m.h, r, R)
        m.L.nbb200_set_restricted_sort(m.h, int(os.environ.get("PROFILE_RESTRICT", "1")))
    if mode == "spline":
        st = C.c_int(16)
        m.L.PairwiseInteractionABFS_B200_SetInteractionForm(m.h, 0, 50, C.byref(st))
    if mode == "norebuild":
        m.step(rebuild=True)
        g = torch.Generator(device=m.x.device).manual_seed(5)
        x0 = m.x.clone()
        for k in range(steps):
            m.x.copy_(x0 + 0.03 * (k + 1) * torch.randn(x0.shape, generator=g, device=x0.device, dtype=x0.dtype).clamp_(-2, 2))
            m.step(rebuild=False)
    else:
        for _ in range(steps):
            m.step(rebuild=True)
    torch.cuda.synchronize()
    print(name, mode, m.state.Counters(), m.state.Timings(), m.e)
