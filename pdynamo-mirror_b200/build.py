"""Build libnbabfs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python pdynamo-mirror_b200/build.py [--force] [--verbose]

list_build.cu is compiled with -fmad=false: its fp64 distance predicate must round like the reference's
(gcc -O2, no FMA contraction).  force_kernels.cu keeps FMA contraction (fp32 pair math).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnbabfs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function"] + os.environ.get("NBB200_NVCC_FLAGS", "").split()
UNITS = [("api.cu", []), ("list_build.cu", ["-fmad=false"]), ("force_kernels.cu", []), ("mm_terms.cu", ["-fmad=false"]), ("qcmm.cu", ["-fmad=false"]), ("symmetry_host.cpp", []), ("host_rows.cpp", [])]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(HERE, "..", "include", "nbabfs_b200.h")]
    objs, relink = [], force or not os.path.exists(OUT)
    for name, extra in UNITS:
        src = os.path.join(CSRC, name)
        obj = os.path.join(HERE, "build", name.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            relink = True
    if relink:
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
