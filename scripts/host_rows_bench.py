"""Timing of the host row gather / scatter-add helpers (csrc/host_rows.cpp) on the box's CPU: python scripts/host_rows_bench.py [threads]"""
import ctypes as C, os, sys, time
import numpy as np
if len(sys.argv) > 1:
    os.environ["NBB200_HOST_THREADS"] = sys.argv[1]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
importlib.import_module("pdynamo_mirror_b200")
from pdynamo_mirror_b200 import _lib
L = _lib.lib()
n = 1119744
rng = np.random.default_rng(0)
x = rng.normal(size=(n, 3)); g = np.zeros((n, 3))
for cnt in (n // 8, n // 2):
    # a slab: atoms of a contiguous region, locally shuffled (sorted order inside cells)
    ids = np.arange(cnt, dtype=np.int32).reshape(-1, 24)
    ids = ids[:, rng.permutation(24)].reshape(-1).copy()
    out = np.zeros((cnt, 3))
    for rep in range(4):
        t = time.perf_counter(); L.nbb200_host_gather_rows(C.c_void_p(x.ctypes.data), C.c_void_p(ids.ctypes.data), cnt, C.c_void_p(out.ctypes.data)); t1 = time.perf_counter()
        L.nbb200_host_scatter_add_rows(C.c_void_p(g.ctypes.data), C.c_void_p(ids.ctypes.data), cnt, C.c_void_p(out.ctypes.data)); t2 = time.perf_counter()
    print("threads", os.environ.get("NBB200_HOST_THREADS", "4"), "rows", cnt, "gather %.3f ms scatter-add %.3f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3))
    t = time.perf_counter(); np.take(x, ids, axis=0, out=out); t1 = time.perf_counter(); print("  numpy take %.3f ms" % ((t1 - t) * 1e3))
