"""Where does an MD step's wall time go?  Host timers around every call of LangevinDynamics.Run (DHFR, all terms)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pdynamo_mirror_b200 as p

w = p.workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "dhfr_mm"]()
sysm = p.System.FromWorkload(w)
sysm.DefineNBModel(p.NBModelABFS())
md = p.md.LangevinDynamics(sysm)
md.Run(50)
T = {}


def tick(name, t0):
    t = time.perf_counter()
    T[name] = T.get(name, 0.0) + t - t0
    return t


steps = 500
torch.cuda.synchronize()
tstart = time.perf_counter()
for k in range(steps):
    t = time.perf_counter()
    md.iteration += 1
    md.L.nbb200_langevin_first_half(md.h, md._p(md.x), md._p(md.v), md._p(md.a), md._p(md.mass), md._lib.d_(md.factors), C.c_ulonglong(md.seed), C.c_ulonglong(md.iteration))
    t = tick("langevin launch", t)
    st = C.c_int(16)
    md.g.zero_()
    t = tick("g.zero_", t)
    if md.mmterms is not None:
        md.mmterms.EnqueueDevice(md.x.data_ptr(), md.g.data_ptr())
    t = tick("bonded enqueue", t)
    md.updates += md.L.NBModelABFS_B200_UpdateDevice(md.h, md._p(md.x), md._lib.d_(md.box), 0, C.byref(st))
    t = tick("UpdateDevice (check + sync [+ rebuild])", t)
    md.L.NBModelABFS_B200_MMMMEnergyDevice(md.h, md._lib.d_(md.energies), md._p(md.g), md._lib.d_(md.dEdM), C.byref(st))
    t = tick("MMMMEnergyDevice (launch + sync)", t)
    if md.mmterms is not None:
        md.mmterms.CollectDevice()
    t = tick("bonded collect", t)
    md.L.nbb200_vv_second_half(md.h, md._p(md.v), md._p(md.a), md._p(md.g), md._p(md.mass), 2.0 * md.facV3, md._p(md.ke_dev))
    t = tick("second half launch", t)
    ke = float(md.ke_dev.item())
    t = tick("ke.item()", t)
torch.cuda.synchronize()
total = time.perf_counter() - tstart
print("steps/s %.0f, us/step %.1f, updates %d" % (steps / total, 1e6 * total / steps, md.updates))
for k, v in T.items():
    print("  %-44s %7.1f us/step" % (k, 1e6 * v / steps))
