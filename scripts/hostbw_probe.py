import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import pdynamo_mirror_b200
from pdynamo_mirror_b200 import _lib
for T in (1, 2, 4, 8, 16):
    pass
L = _lib.lib()
m = 3 * 559872
src = np.random.rand(m); 
pin = torch.empty(m, dtype=torch.float64).pin_memory(); dst = pin.numpy()
dev = torch.empty(m, dtype=torch.float64, device="cuda")
def t(f, reps=10):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("threads env", os.environ.get("NBB200_HOST_THREADS"), "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
print("host_copy pageable->pinned %.3f ms" % t(lambda: L.nbb200_host_copy(C.c_void_p(dst.ctypes.data), C.c_void_p(src.ctypes.data), m)))
print("numpy copyto %.3f ms" % t(lambda: np.copyto(dst, src)))
print("H2D from pinned %.3f ms" % t(lambda: dev.copy_(pin, non_blocking=True)))
srct = torch.from_numpy(src)
print("H2D from pageable %.3f ms" % t(lambda: dev.copy_(srct)))
print("D2H to pinned %.3f ms" % t(lambda: pin.copy_(dev, non_blocking=True)))
g = np.zeros(m)
print("host_add %.3f ms" % t(lambda: L.nbb200_host_add(C.c_void_p(g.ctypes.data), C.c_void_p(dst.ctypes.data), m)))
print("numpy += %.3f ms" % t(lambda: np.add(g, dst, out=g)))
