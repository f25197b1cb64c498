"""Generate the committed fixtures under tests/golden/ from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  python tests/golden/make_fixtures.py inputs    # input-structure fixtures read from reference data files
  python tests/golden/make_fixtures.py golden    # golden outputs of the COMPILED reference (oracle/_ref) on committed inputs

Inputs:
  water216_cubicBox.npz  <- book/data/mol/water216_cubicBox.mol (equilibrated 216-water box of book Example 20,
                            a = 18.641: book/examples/Example20.py:6-17).  Reordered O,H,H per molecule.
  bala_c7eq.npz          <- book/data/mol/bala_c7eq.mol (blocked alanine dipeptide, coordinates + bonds)
Golden outputs (one npz per case): six energies, gradients, dE/dM, pair counts, a hash of the canonical pair sets and the
full pair sets for the small cases -- all produced by the unmodified reference C code through oracle/ref_driver.c.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def read_mol(path):
    with open(path) as f:
        lines = f.read().split("\n")
    counts = lines[3]
    na, nb = int(counts[0:3]), int(counts[3:6])
    xyz = np.array([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in lines[4:4 + na]])
    sym = [l[31:34].strip() for l in lines[4:4 + na]]
    bonds = np.array([[int(l[0:3]) - 1, int(l[3:6]) - 1] for l in lines[4 + na:4 + na + nb]], dtype=np.int32)
    return xyz, sym, bonds


def make_inputs():
    xyz, sym, bonds = read_mol(os.path.join(REF, "book/data/mol/water216_cubicBox.mol"))
    # regroup as O,H,H per molecule following the bond table
    partners = {}
    for i, j in bonds:
        o, h = (i, j) if sym[i] == "O" else (j, i)
        partners.setdefault(int(o), []).append(int(h))
    order = []
    for o in sorted(partners):
        assert len(partners[o]) == 2
        order += [o] + sorted(partners[o])
    assert sorted(order) == list(range(len(sym)))
    np.savez_compressed(os.path.join(HERE, "water216_cubicBox.npz"), xyz=xyz[order], a=18.641)
    xyz, sym, bonds = read_mol(os.path.join(REF, "book/data/mol/bala_c7eq.mol"))
    np.savez_compressed(os.path.join(HERE, "bala_c7eq.npz"), xyz=xyz, symbols=np.array(sym), bonds=bonds)
    print("inputs written")


def pair_hash(keys):
    return hashlib.sha256(np.ascontiguousarray(keys, dtype=np.int64).tobytes()).hexdigest()


def make_golden():
    import pdynamo_mirror_b200 as p
    import oracle
    import refnb
    cases = p.workloads.GOLDEN_CASES
    for name, (maker, opts, store_pairs) in cases.items():
        w = maker()
        r = refnb.RefNB(w, **opts)
        out = r.energy(force_new=True)
        prim = oracle.canonical_primary(r.primary_pairs())
        imgs = r.images()
        data = dict(energies=out["energies"], grad=out["grad"], dEdM=out["dEdM"], nprimary=len(prim),
                    image_meta=np.array([[im["t"], im["a"], im["b"], im["c"], len(im["pairs"])] for im in imgs], dtype=np.int64).reshape(-1, 5),
                    image_scale=np.array([im["scale"] for im in imgs]),
                    primary_hash=pair_hash(prim),
                    image_hashes=np.array([pair_hash(oracle.canonical_cross(im["pairs"])) for im in imgs]))
        if store_pairs:
            data["primary_keys"] = prim
            for k, im in enumerate(imgs):
                data["image_keys_%d" % k] = oracle.canonical_cross(im["pairs"])
        np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % name), **data)
        print(name, out["energies"], len(prim), len(imgs))
        r.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("inputs", "all"):
        make_inputs()
    if what in ("golden", "all"):
        make_golden()
