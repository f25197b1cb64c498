"""Generate the committed fixtures under tests/golden/ from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  python tests/golden/make_fixtures.py inputs    # input-structure fixtures read from reference data files
  python tests/golden/make_fixtures.py golden    # golden outputs of the COMPILED reference (oracle/_ref) on committed inputs
  python tests/golden/make_fixtures.py qcmm      # golden outputs of the QC/MM entry points (QC region without boundary atoms)

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


# ----------------------------------------------------------------------------------------------------
# Bonded terms of the same DHFR input: PSF angle / dihedral / improper sections and the CHARMM parameter
# sections, matched as the reference does (pBabel CHARMMParameterFileReader.ProcessBond/Angle/Dihedral/Improper :160-313 --
# canonical keys, "X A B X" dihedral and "A X X B" improper wild cards, multi-term dihedrals, Urey-Bradley terms only where
# the angle line carries them; CHARMMPSFFileReader.ToHarmonicBondContainer / ...Angle / ...UreyBradley / ToFourierDihedral /
# ToHarmonicImproperContainer :410-651).  Units: kcal -> kJ (4.184), degrees -> radians; E = fc (q - q0)^2 (no 1/2).
# ----------------------------------------------------------------------------------------------------
_PRM_SECTIONS = ("ANGL", "ATOM", "BOND", "CMAP", "DIHE", "END", "EQUI", "HBON", "IMPH", "IMPR", "NBFI", "NBON", "NONB", "PHI", "PRIN", "SPAS", "THET")


def _prm_sections(path):
    out, current, pending = {}, None, ""
    with open(path) as f:
        for raw in f:
            line = raw.strip()
            k = line.find("!")
            if k >= 0:
                line = line[:k].strip()
            if line.endswith("-"):
                pending += line[:-1].strip() + " "
                continue
            line = (pending + line).strip()
            pending = ""
            if not line or line.startswith("*"):
                continue
            head = line.split(" ", 1)[0].upper()
            sec = next((nm for nm in _PRM_SECTIONS if head.startswith(nm)), None)
            if sec is not None:
                if sec == "END":
                    break
                current = sec
                continue
            if current is not None:
                out.setdefault(current, []).append(line.split())
    return out


def read_dhfr_bonded():
    base = os.path.join(REF, "benchmarks/data/dhfr")
    sec = _psf_sections(os.path.join(base, "dhfr.psfx"))
    natom, body = sec["NATOM"]
    types = [line.split()[5].upper() for line in body[:natom]]
    masses = np.array([float(line.split()[7]) for line in body[:natom]])

    def ints(tag, width):
        count, lines = sec[tag]
        flat = [int(v) for line in lines for v in line.split()]
        return np.array(flat[:width * count], dtype=np.int64).reshape(-1, width) - 1
    bonds, angles, dihedrals, impropers = ints("NBOND", 2), ints("NTHETA", 3), ints("NPHI", 4), ints("NIMPHI", 4)
    prm = _prm_sections(os.path.join(base, "par_all22_prot.prm"))
    KC, DEG = 4.184, np.pi / 180.0
    pb, pa, pub, pd, pdw, pi_, piw = {}, {}, {}, {}, {}, {}, {}
    for d in prm["BOND"]:
        t1, t2 = d[0].upper(), d[1].upper()
        pb.setdefault((max(t1, t2), min(t1, t2)), (KC * float(d[2]), float(d[3])))
    for d in prm["ANGL"]:
        t1, t2, t3 = d[0].upper(), d[1].upper(), d[2].upper()
        key = (max(t1, t3), t2, min(t1, t3))
        pa.setdefault(key, (KC * float(d[3]), DEG * float(d[4])))
        if len(d) > 5:
            pub.setdefault(key, (KC * float(d[5]), float(d[6])))
    for d in prm["DIHE"]:
        t1, t2, t3, t4 = [v.upper() for v in d[:4]]
        key = (t1, t2, t3, t4) if t2 > t3 else ((max(t1, t4), t2, t3, min(t1, t4)) if t2 == t3 else (t4, t3, t2, t1))
        tab = pdw if "X" in key else pd
        terms = tab.get(key, [])
        nper = int(d[5])
        if all(t[0] != nper for t in terms):
            terms.append((nper, KC * float(d[4]), DEG * float(d[6])))
            terms.sort()
            tab[key] = terms
    for d in prm["IMPR"]:
        t1, t2, t3, t4 = [v.upper() for v in d[:4]]
        key = (t1, t2, t3, t4) if t1 > t4 else ((t1, max(t2, t3), min(t2, t3), t4) if t1 == t4 else (t4, t3, t2, t1))
        (piw if "X" in key else pi_).setdefault(key, (KC * float(d[4]), DEG * float(d[6])))
    out = dict(masses=masses)
    # bonds
    out["bonds"] = bonds.astype(np.int32)
    bp = [pb[(max(types[i], types[j]), min(types[i], types[j]))] for i, j in bonds]
    out["bond_fc"], out["bond_eq"] = np.array([v[0] for v in bp]), np.array([v[1] for v in bp])
    # angles + Urey-Bradley
    keys = [(max(types[i], types[k]), types[j], min(types[i], types[k])) for i, j, k in angles]
    out["angles"] = angles.astype(np.int32)
    out["angle_fc"], out["angle_eq"] = np.array([pa[k][0] for k in keys]), np.array([pa[k][1] for k in keys])
    ub = [(a[0], a[2], pub[k]) for a, k in zip(angles, keys) if k in pub]
    out["ureybradleys"] = np.array([[u[0], u[1]] for u in ub], dtype=np.int32).reshape(-1, 2)
    out["ub_fc"], out["ub_eq"] = np.array([u[2][0] for u in ub]), np.array([u[2][1] for u in ub])
    # dihedrals: one term per (dihedral, multiplicity)
    dt, dp = [], []
    for i, j, k, l in dihedrals:
        ti, tj, tk, tl = types[i], types[j], types[k], types[l]
        key = (ti, tj, tk, tl) if tj > tk else ((max(ti, tl), tj, tk, min(ti, tl)) if tj == tk else (tl, tk, tj, ti))
        terms = pd[key] if key in pd else pdw[("X", max(tj, tk), min(tj, tk), "X")]
        for nper, fc, phase in terms:
            dt.append((i, j, k, l)); dp.append((fc, nper, phase))
    out["dihedrals"] = np.array(dt, dtype=np.int32).reshape(-1, 4)
    out["dihedral_fc"], out["dihedral_period"], out["dihedral_phase"] = np.array([v[0] for v in dp]), np.array([v[1] for v in dp], dtype=np.int32), np.array([v[2] for v in dp])
    # impropers
    ip = []
    for i, j, k, l in impropers:
        ti, tj, tk, tl = types[i], types[j], types[k], types[l]
        key = (ti, tj, tk, tl) if ti > tl else ((ti, max(tj, tk), min(tj, tk), tl) if ti == tl else (tl, tk, tj, ti))
        ip.append(pi_[key] if key in pi_ else piw[(max(ti, tl), "X", "X", min(ti, tl))])
    out["impropers"] = impropers.astype(np.int32)
    out["improper_fc"], out["improper_eq"] = np.array([v[0] for v in ip]), np.array([v[1] for v in ip])
    # published by the reference for this input (benchmarks/log/systemBenchmarks_Serial_1ps.log:397-400): bond, angle, Urey-Bradley, dihedral, improper; total PE; RMS gradient
    out["published_bonded"] = np.array([12444.3942, 9454.9073, 130.3783, 3015.4953, 52.1977])
    out["published_total"] = np.array([-375469.1160, 1.4766])
    return out


def make_dhfr_bonded():
    d = read_dhfr_bonded()
    np.savez_compressed(os.path.join(HERE, "dhfr_bonded.npz"), **d)
    print("dhfr bonded: %d bonds, %d angles, %d Urey-Bradley, %d dihedral terms, %d impropers" %
          (len(d["bonds"]), len(d["angles"]), len(d["ureybradleys"]), len(d["dihedrals"]), len(d["impropers"])))
    # golden output of the compiled reference's own containers (HarmonicBondContainer_Energy ... through oracle/ref_driver.c: refmm_energy)
    import refnb
    xyz = np.load(os.path.join(HERE, "dhfr_jac.npz"))["xyz"]
    e, g = refnb.mm_energy(d, xyz)
    np.savez_compressed(os.path.join(HERE, "golden_dhfr_bonded.npz"), energies=e, grad=g)
    print("reference bonded energies", e, "published", d["published_bonded"])


# ----------------------------------------------------------------------------------------------------
# The 12 molecular crystals of pMolecule-1.9.0/tests/CrystalMMEnergies.py (:30-63): AMBER top/crd files from
# pMolecule-1.9.0/data/molecularCrystals, space-group operations and cell parameters from the test file.  Non-P1 space
# groups exercise rotations S != I, self-inverse images (scale 0.5), inverse-pair skipping and triclinic cells.
# Only what the NB path needs is read (charges, LJ coefficient tables, bonds); the published crystal energies include
# bonded terms and are therefore not usable as NB known answers -- parity on these inputs is against the compiled reference.
# ----------------------------------------------------------------------------------------------------
_CRYSTAL_OPS = {"C2": ["(x,y,z)", "(-x,y,-z)"],
                "I4": ["(x,y,z)", "(-x,-y,z)", "(-y,x,z)", "(y,-x,z)", "(x+1/2,y+1/2,z+1/2)", "(-x+1/2,-y+1/2,z+1/2)", "(-y+1/2,x+1/2,z+1/2)", "(y+1/2,-x+1/2,z+1/2)"],
                "P1": ["(x,y,z)"],
                "P21a": ["(x,y,z)", "(-x+1/2,y+1/2,-z)", "(-x,-y,-z)", "(x+1/2,-y+1/2,z)"],
                "P212121": ["(x,y,z)", "(-x+1/2,-y,z+1/2)", "(-x,y+1/2,-z+1/2)", "(x+1/2,-y+1/2,-z)"],
                "P61": ["(x,y,z)", "(-y,x-y,z+1/3)", "(-x+y,-x,z+2/3)", "(-x,-y,z+1/2)", "(y,-x+y,z+5/6)", "(x-y,x,z+1/6)"],
                "P21c": ["(x,y,z)", "(-x,y+1/2,-z+1/2)", "(-x,-y,-z)", "(x,-y+1/2,z+1/2)"],
                "R3": ["(x,y,z)", "(z,x,y)", "(y,z,x)"]}
_CRYSTALS = [("ALAALA", "I4", dict(a=17.9850, b=17.9850, c=5.1540)),
             ("ALAMET01", "P21c", dict(a=13.089, b=5.329, c=15.921, beta=108.57)),
             ("AQARUF", "P61", dict(a=14.3720, b=14.3720, c=9.8282, gamma=120.0)),
             ("BEVXEF01", "P212121", dict(a=9.6590, b=9.6720, c=10.7390)),
             ("GLYALB", "P212121", dict(a=9.6930, b=9.5240, c=7.5370)),
             ("GLYGLY", "P21a", dict(a=7.812, b=9.566, c=9.410, beta=124.60)),
             ("GUFQON", "P212121", dict(a=7.2750, b=9.0970, c=10.5070)),
             ("HXACAN19", "P21a", dict(a=12.8720, b=9.3700, c=7.0850, beta=115.6200)),
             ("IWANID", "C2", dict(a=23.091, b=5.494, c=17.510, beta=117.88)),
             ("LCDMPP10", "P1", dict(a=8.067, b=6.082, c=5.155, alpha=131.7, beta=82.4, gamma=106.6)),
             ("WIRYEB", "P61", dict(a=14.4240, b=14.4240, c=9.9960, gamma=120.0)),
             ("WABZOO", "R3", dict(a=12.5940, b=12.5940, c=12.5940, alpha=118.03, beta=118.03, gamma=118.03))]


def _symop(ostring):
    """pCore.Transformation3.pyx:127-163 (Transformation3_FromSymmetryOperationString)"""
    t = ostring.replace(" ", "").strip("()")
    rot, tr = np.zeros((3, 3)), np.zeros(3)
    for i, s in enumerate(t.split(",")):
        ns = [s[0:1]]
        for j, c in enumerate(s[1:]):
            if s[j:j + 1].isdigit() and not (c.isdigit() or c == "."):
                ns.append(".")
            ns.append(c.lower())
        if ns[-1].isdigit():
            ns.append(".")
        e = "".join(ns)
        t0 = eval(e, {}, dict(x=0.0, y=0.0, z=0.0))
        rx = eval(e, {}, dict(x=1.0, y=0.0, z=0.0)) - t0
        ry = eval(e, {}, dict(x=1.0, y=1.0, z=0.0)) - t0 - rx
        rz = eval(e, {}, dict(x=1.0, y=1.0, z=1.0)) - t0 - rx - ry
        tr[i], rot[i] = t0, (rx, ry, rz)
    return rot, tr


def _amber_top(path):
    flags, cur = {}, None
    with open(path) as f:
        for line in f:
            if line.startswith("%FLAG"):
                cur = line.split()[1]
                flags[cur] = []
            elif line.startswith("%FORMAT") or line.startswith("%VERSION"):
                fmt = line
                if cur is not None:
                    flags[cur].append(("fmt", line.strip()))
            elif cur is not None:
                flags[cur].append(("data", line.rstrip("\n")))
    out = {}
    for k, items in flags.items():
        fmt = [v for t, v in items if t == "fmt"][0]
        width = int(fmt.split("(")[1].rstrip(")").lower().replace("e", "a").replace("i", "a").split("a")[1].split(".")[0])
        vals = []
        for t, v in items:
            if t == "data":
                vals += [v[i:i + width] for i in range(0, len(v), width) if v[i:i + width].strip()]
        out[k] = vals
    return out


def make_crystals():
    base = os.path.join(REF, "pMolecule-1.9.0/data/molecularCrystals")
    data = {"names": np.array([c[0] for c in _CRYSTALS])}
    for name, group, cell in _CRYSTALS:
        top = _amber_top(os.path.join(base, name + ".top"))
        n = int(top["POINTERS"][0])
        ntypes = int(top["POINTERS"][1])
        q = np.array([float(v) for v in top["CHARGE"]]) / 18.2223
        ati = np.array([int(v) for v in top["ATOM_TYPE_INDEX"]], dtype=np.int32) - 1
        nbi = np.array([int(v) for v in top["NONBONDED_PARM_INDEX"]], dtype=np.int32).reshape(ntypes, ntypes) - 1
        acoef = np.array([float(v) for v in top["LENNARD_JONES_ACOEF"]]) * 4.184
        bcoef = np.array([float(v) for v in top["LENNARD_JONES_BCOEF"]]) * 4.184
        bonds = [int(v) for v in top.get("BONDS_INC_HYDROGEN", [])] + [int(v) for v in top.get("BONDS_WITHOUT_HYDROGEN", [])]
        bonds = np.array(bonds, dtype=np.int64).reshape(-1, 3)[:, :2] // 3
        with open(os.path.join(base, name + ".crd")) as f:
            lines = f.read().split("\n")
        vals = [float(l[i:i + 12]) for l in lines[2:] for i in range(0, len(l), 12) if l[i:i + 12].strip()]
        xyz = np.array(vals[:3 * n]).reshape(n, 3)
        ops = [_symop(o) for o in _CRYSTAL_OPS[group]]
        box = [cell["a"], cell["b"], cell["c"], cell.get("alpha", 90.0), cell.get("beta", 90.0), cell.get("gamma", 90.0)]
        assert len(q) == n and nbi.min() >= 0
        data[name + "_xyz"], data[name + "_q"], data[name + "_type"] = xyz, q, ati
        data[name + "_nbindex"], data[name + "_A"], data[name + "_B"] = nbi.astype(np.int32), acoef, bcoef
        data[name + "_bonds"] = bonds.astype(np.int32)
        data[name + "_rot"], data[name + "_trans"] = np.array([o[0] for o in ops]), np.array([o[1] for o in ops])
        data[name + "_box"] = np.array(box)
        print(name, group, n, "atoms", ntypes, "types", len(bonds), "bonds", len(ops), "operations")
    np.savez_compressed(os.path.join(HERE, "crystals.npz"), **data)


def pair_hash(keys):
    return hashlib.sha256(np.ascontiguousarray(keys, dtype=np.int64).tobytes()).hexdigest()


def make_golden(only=None):
    import pdynamo_mirror_b200 as p
    import oracle
    import refnb
    cases = p.workloads.GOLDEN_CASES
    for name, (maker, opts, store_pairs) in cases.items():
        if only and name not in only:
            continue
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



# ----------------------------------------------------------------------------------------------------
# the headline workload itself (1 119 744-atom water box, bench.py's m1) through the compiled reference: energies, dE/dM, list sizes, a
# 65 536-row sample of the gradient and eight seeded random projections of the whole gradient (what a GPU test can compare at the
# size the bench runs at; the full fp64 gradient would be 27 MB).  About a minute and ~20 GB on the OpenMP build.
# ----------------------------------------------------------------------------------------------------
def m1_projection_vectors(n, count=8, seed=20261017):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((n, 3)) for _ in range(count)]


def make_golden_m1():
    import pdynamo_mirror_b200 as p
    import refnb
    w = p.workloads.WORKLOADS["m1"]()
    r = refnb.RefNB(w, omp=True)
    out = r.energy(force_new=True)
    c = r.counts()
    g = out["grad"]
    rows = np.sort(np.random.default_rng(7).choice(w["n"], 65536, replace=False))
    proj = np.array([float((v * g).sum()) for v in m1_projection_vectors(w["n"])])
    data = dict(energies=out["energies"], dEdM=out["dEdM"], counts=np.array([c[k] for k in sorted(c)], dtype=np.int64), count_keys=np.array(sorted(c)),
                grad_rows=rows, grad_sample=g[rows], grad_rms=float(np.sqrt((g * g).mean())), grad_sum=g.sum(axis=0), grad_proj=proj)
    np.savez_compressed(os.path.join(HERE, "golden_m1.npz"), **data)
    print("m1", out["energies"], c)
    r.close()


# ----------------------------------------------------------------------------------------------------
# QC/MM entry points of NBModelABFS (second half): golden vectors of the COMPILED reference for a QC region
# without boundary atoms (oracle/ref_driver.c refqc_*, oracle/refnb.py RefQC).  The B200 side of this row is not built yet;
# the vectors pin it in advance.  QC charges: the MM charges of the QC atoms scaled by 0.9 (any fixed vector would do).
# ----------------------------------------------------------------------------------------------------
QCMM_CASES = {"w216": ("w216", [0, 1, 2], [8, 1, 1], False), "w216_vacuum": ("w216", [3, 4, 5], [8, 1, 1], True),
              "bala": ("bala", list(range(22)), [1] * 22, False),
              "crystal_GLYGLY": ("crystal_GLYGLY", list(range(17)), [1] * 17, False)}      # the whole asymmetric unit: QC/QC image lists only


def qcmm_case(name):
    import pdynamo_mirror_b200 as p
    wname, idx, zs, vacuum = QCMM_CASES[name]
    w = p.workloads.WORKLOADS[wname]()
    if vacuum:
        w = dict(w)
        w["box"], w["rot"], w["trans"] = None, np.zeros((0, 3, 3)), np.zeros((0, 3))
    return w, np.array(idx, np.int32), np.array(zs, np.int32), 0.9 * np.asarray(w["charges"], np.float64)[idx]


def make_qcmm():
    import refnb
    for name in QCMM_CASES:
        w, idx, zs, qcq = qcmm_case(name)
        r = refnb.RefQC(w, idx, zs)
        out = r.energy(qcq)
        keys = {k: np.array(sorted(map(tuple, r.pairs(which).tolist())), dtype=np.int32).reshape(-1, 2) for which, k in ((1, "nbqcmmlj"), (2, "nbqcmmel"))}
        np.savez_compressed(os.path.join(HERE, "golden_qcmm_%s.npz" % name), qc_index=idx, qc_charges=qcq, energies=out["energies"],
                            potentials=out["potentials"], qcqc_potentials=out["qcqc_potentials"], grad_lj=out["grad_lj"], grad_el=out["grad_el"],
                            dEdM=out["dEdM"], count_labels=np.array(refnb.RefQC.COUNT_LABELS), counts=np.array([out["counts"][k] for k in refnb.RefQC.COUNT_LABELS]),
                            nbqcmmlj=keys["nbqcmmlj"], nbqcmmel=keys["nbqcmmel"])
        print(name, out["energies"][6:], out["potentials"][:3], out["counts"])
        r.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("inputs", "all"):
        make_inputs()
        make_dhfr()
        make_dhfr_bonded()
        make_crystals()
    if what == "bonded":
        make_dhfr_bonded()
    if what in ("qcmm", "all"):
        make_qcmm()
    if what in ("golden", "all"):
        make_golden(only=sys.argv[2:])          # python make_fixtures.py golden [case ...]
    if what == "m1":
        make_golden_m1()                        # the 1.1 M-atom bench workload (not part of "all": a minute of CPU, ~20 GB)
