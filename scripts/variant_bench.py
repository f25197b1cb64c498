"""Time the force kernel for the launch shapes selectable through NBB200_FORCE_SHAPE (development aid)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "m1"
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["256x3", "256x2", "128x5", "128x4", "128x6"]
for n in names:
    env = dict(os.environ, NBB200_FORCE_SHAPE=n)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu", "--no-jac", "--workload", wl],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().split("\n")[-1])
        print(n, "tile_forces %.3f ms  rebuild %.3f ms  step %.3f ms  no-rebuild %.3f ms" % (d["kernels_ms"]["tile_forces"], d["kernels_ms"]["list_rebuild"], d["ms_per_step"], d["no_rebuild"]["ms_per_call"]), flush=True)
    except Exception as e:
        print(n, "failed", e, out.stderr[-500:])
