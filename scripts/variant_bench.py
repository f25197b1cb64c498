"""Time the force kernel for the kernel variants selected through environment variables (development aid)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "m1"
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["128x4u2", "128x4u4", "128x5u2", "128x5u4", "128x6u2"]
variants = [dict(NBB200_FORCE_KERNEL="scalar", NBB200_SCALAR_BLOCKS="3")] + [dict(NBB200_FORCE_KERNEL="x2", NBB200_X2_VARIANT=n) for n in names]
for v in variants:
    env = dict(os.environ); env.update(v)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu", "--no-jac", "--workload", wl],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().split("\n")[-1])
        print(v, "tile_forces %.3f ms  rebuild %.3f ms  step %.3f ms" % (d["kernels_ms"]["tile_forces"], d["kernels_ms"]["list_rebuild"], d["ms_per_step"]), flush=True)
    except Exception as e:
        print(v, "failed", e, out.stderr[-500:])
