"""Quick GPU parity probe (development aid): CUDA path vs oracle restatement on a few workloads."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pdynamo_mirror_b200 as p
import oracle

def run(name, w=None):
    w = w or p.workloads.WORKLOADS[name]()
    t = time.time()
    o = oracle.OracleNB(w)
    ro = o.energy(force_new=True)
    to = time.time() - t
    s = p.System.FromWorkload(w)
    s.DefineNBModel(p.NBModelABFS())
    t = time.time()
    s.Energy(doGradients=True)
    tg = time.time() - t
    cfg = s.configuration
    st = cfg.nbState
    e = st.energies.copy()
    g = cfg.gradients3
    dm = cfg.symmetryParameterGradients.dEdM if hasattr(cfg, "symmetryParameterGradients") else np.zeros((3, 3))
    print("==", name, "n", w["n"], "oracle %.2fs gpu first call %.2fs" % (to, tg))
    print("  counters", st.Counters())
    oc = o.counts()
    print("  pairs primary gpu/orc", st.NumberOfPairs(), oc["primary"], "images", st.NumberOfImages(), oc["images"], "image pairs", st.NumberOfImagePairs(), oc["image_pairs"])
    for k, lab in enumerate(st.LABELS):
        ref = ro["energies"][k]
        rel = abs(e[k] - ref) / max(abs(ref), 1e-30) if ref != 0 else abs(e[k])
        print("  %-20s gpu %18.8f  orc %18.8f  rel %.2e" % (lab, e[k], ref, rel))
    gr = ro["grad"]
    print("  grad rel RMS err %.3e (rms %.4f) max abs %.3e" % (np.sqrt(((g - gr) ** 2).mean()) / np.sqrt((gr ** 2).mean()), np.sqrt((gr ** 2).mean()), np.abs(g - gr).max()))
    print("  dEdM rel Frobenius err %.3e" % (np.linalg.norm(dm - ro["dEdM"]) / max(np.linalg.norm(ro["dEdM"]), 1e-30)))
    # pair sets
    t = time.time()
    same = np.array_equal(oracle.canonical_primary(st.Pairs(-1)), oracle.canonical_primary(o.primary_pairs()))
    gi, oi = st.Images(), o.images()
    meta = [(x["t"], x["a"], x["b"], x["c"], x["scale"]) for x in gi] == [(x["t"], x["a"], x["b"], x["c"], x["scale"]) for x in oi]
    simg = meta and all(np.array_equal(oracle.canonical_cross(x["pairs"]), oracle.canonical_cross(y["pairs"])) for x, y in zip(gi, oi))
    print("  primary set equal:", same, " image meta equal:", meta, " image sets equal:", simg, "(%.2fs)" % (time.time() - t))
    # second call, no rebuild, timing
    t = time.time(); s.Energy(doGradients=True); t2 = time.time() - t
    print("  second call %.4fs  timings" % t2, st.Timings())
    return same and simg

if __name__ == "__main__":
    names = sys.argv[1:] or ["w216", "w1728", "jac"]
    ok = True
    for nm in names:
        ok = run(nm) and ok
    print("ALL PAIR SETS EQUAL" if ok else "PAIR SET MISMATCH")
