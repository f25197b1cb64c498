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


# ----------------------------------------------------------------------------------------------------
# DHFR / JAC benchmark system of the reference (benchmarks/data/dhfr): CHARMM22 PSF (XPLOR) + par_all22_prot.prm + xyz.
# Minimal Python-3 readers, written for this fixture only, that follow what the reference's readers extract for the NB
# path: atom types / charges (pBabel CHARMMPSFFileReader.ToMMAtomContainer :680-711), 1-2/1-3/1-4 exclusions from the bond
# list (ToExclusionPairLists :335-396), LJ parameters incl. separate 1-4 values (CHARMMParameterFileReader.ProcessNonBond
# :315-344, ToLJParameterContainers :653-678, AMBER/arithmetic table form).
# ----------------------------------------------------------------------------------------------------
def _psf_sections(path):
    with open(path) as f:
        lines = f.read().split("\n")
    i, out = 1, {}
    while i < len(lines):
        line = lines[i]
        if "!" in line:
            head, tag = line.split("!", 1)
            tag = tag.split(":")[0].strip()
            count = int(head.split()[0]) if head.split() else 0
            body = []
            i += 1
            while i < len(lines) and "!" not in lines[i]:
                if lines[i].strip():
                    body.append(lines[i])
                i += 1
            out[tag] = (count, body)
        else:
            i += 1
    return out


def read_dhfr():
    base = os.path.join(REF, "benchmarks/data/dhfr")
    sec = _psf_sections(os.path.join(base, "dhfr.psfx"))
    natom, body = sec["NATOM"]
    types, charges = [], []
    for line in body[:natom]:
        f = line.split()
        types.append(f[5].upper())
        charges.append(float(f[6]))
    nbond, body = sec["NBOND"]
    flat = [int(v) for line in body for v in line.split()]
    bonds = np.array(flat[:2 * nbond], dtype=np.int64).reshape(-1, 2) - 1
    nnb = sec.get("NNB", (0, []))[0]
    assert nnb == 0, "explicit NNB exclusions not handled"
    # exclusions as ToExclusionPairLists builds them
    i12 = set((max(int(i), int(j)), min(int(i), int(j))) for i, j in bonds)
    conn = [[] for _ in range(natom)]
    for i, j in i12:
        conn[i].append(j)
        conn[j].append(i)
    for c in conn:
        c.sort()
    i13 = set()
    for j in range(natom):
        jb = conn[j]
        for a in range(1, len(jb)):
            for b in range(a):
                i13.add((jb[a], jb[b]))
    i123 = i12 | i13
    i14p = set()
    for j, k in i12:
        for i in conn[j]:
            if i != k:
                for l in conn[k]:
                    if l != i and l != j:
                        i14p.add((max(i, l), min(i, l)))
    excl = np.array(sorted(i123 | i14p), dtype=np.int32)
    p14 = np.array(sorted(i14p - i123), dtype=np.int32)
    # non-bonded parameters
    nb, nb14, on = {}, {}, False
    with open(os.path.join(base, "par_all22_prot.prm")) as f:
        for raw in f:
            line = raw.split("!")[0].strip()
            if not line:
                continue
            key = line.split()[0].upper()
            if key.startswith("NONB"):
                on = True
                continue
            if key in ("HBOND", "NBFIX", "END", "BONDS", "ANGLES", "DIHEDRALS", "IMPROPER", "CMAP"):
                on = False
                continue
            if on:
                f4 = line.split()
                if f4[0].lower().startswith("cutnb"):
                    continue
                t = f4[0].upper()
                nb[t] = (abs(float(f4[2])) * 4.184, 2.0 * float(f4[3]))
                if len(f4) >= 7:
                    nb14[t] = (abs(float(f4[5])) * 4.184, 2.0 * float(f4[6]))
    uniq = sorted(set(types))
    index = {t: k for k, t in enumerate(uniq)}
    eps = [nb[t][0] for t in uniq]
    sig = [nb[t][1] for t in uniq]
    eps14 = [nb14.get(t, nb[t])[0] for t in uniq]
    sig14 = [nb14.get(t, nb[t])[1] for t in uniq]
    xyz = np.loadtxt(os.path.join(base, "dhfr.xyz"), skiprows=2, usecols=(1, 2, 3))
    assert xyz.shape == (natom, 3)
    return dict(xyz=xyz, charges=np.array(charges), ljtypes=np.array([index[t] for t in types], dtype=np.int32),
                eps=np.array(eps), sigma=np.array(sig), eps14=np.array(eps14), sigma14=np.array(sig14),
                exclusions=excl, pairs14=p14, a=62.23, types=np.array(uniq))


def make_dhfr():
    d = read_dhfr()
    # per-term energies published by the reference for exactly this input (benchmarks/log/systemBenchmarks_Serial_1ps.log:397-403)
    golden = np.array([-388870.0641, 26802.2738, 17884.0130, 1421.1985, -62084.6864, 4280.7766])
    np.savez_compressed(os.path.join(HERE, "dhfr_jac.npz"), published_energies=golden, published_counts=np.array([8885288, 2942151, 13, 6556, 34709, 35]), **d)
    print("dhfr: %d atoms, %d exclusions, %d 1-4 pairs, %d LJ types, total charge %.4f" %
          (len(d["charges"]), len(d["exclusions"]), len(d["pairs14"]), len(d["types"]), d["charges"].sum()))


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
        make_dhfr()
    if what in ("golden", "all"):
        make_golden()
