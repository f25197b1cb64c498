"""Synthetic MM systems for the NBModelABFS hot path.

Everything here is plain numpy and deterministic (64-bit LCG, no numpy RNG state) so that the same
arrays are produced in this container, on the GPU box, for the CUDA path, the oracle and the compiled
reference.  A system is a dict of flat arrays -- the same data the reference keeps in MMAtomContainer
(pMolecule-1.9.0/extensions/cinclude/MMAtomContainer.h:24-35), LJParameterContainer
(pMolecule-1.9.0/extensions/cinclude/LJParameterContainer.h) and the exclusion / 1-4 PairLists
(pMolecule-1.9.0/pMolecule/MMModel.py builds them from connectivity).

Units follow pDynamo: Angstrom, kJ/mol, elementary charges.
"""
import numpy as np

KCAL = 4.184

# TIP3P (parameters/forceFields/opls/bookSmallExamples/atomTypes.yaml:13,18, lennardJonesParameters.yaml:23,28)
TIP3P_QO, TIP3P_QH = -0.834, 0.417
TIP3P_EPS_O = 0.1521 * KCAL
TIP3P_SIG_O = 3.15061
TIP3P_ROH = 0.9572
TIP3P_HOH = 104.52


# ----------------------------------------------------------------------------------------------------
# LJ tables: restatement of LJParameterContainer_MakeTable
# (pMolecule-1.9.0/extensions/csource/LJParameterContainer.c:184-210)
# ----------------------------------------------------------------------------------------------------
def make_lj_table(eps, sigma, style="opls"):
    """Return (tableindex[nt*nt] int32, tableA[nt(nt+1)/2], tableB[...]).

    style "opls" : geometric sigma, A = 4 e s^12, B = 4 e s^6          (MakeTableOPLS)
    style "amber": arithmetic sigma (= r_min), A = e s^12, B = 2 e s^6 (MakeTableAMBER)
    """
    eps = np.asarray(eps, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    nt = len(eps)
    tindex = np.zeros(nt * nt, dtype=np.int32)
    tA = np.zeros(nt * (nt + 1) // 2)
    tB = np.zeros(nt * (nt + 1) // 2)
    n = 0
    for i in range(nt):
        for j in range(i + 1):
            eij = np.sqrt(eps[i] * eps[j])
            if style == "opls":
                sij = np.sqrt(sigma[i] * sigma[j])
            else:
                sij = 0.5 * (sigma[i] + sigma[j])
            sij6 = sij ** 6
            sij12 = sij6 * sij6
            if style == "opls":
                eij = eij * 4.0
            else:
                sij6 = sij6 * 2.0
            tA[n] = eij * sij12
            tB[n] = eij * sij6
            tindex[j + i * nt] = n
            tindex[i + j * nt] = n
            n += 1
    return tindex, tA, tB


# ----------------------------------------------------------------------------------------------------
# deterministic uniform stream: s = s*6364136223846793005 + 1442695040888963407 ; u = (s>>11)/2^53
# vectorised through the closed form s_k = a^k s_0 + c (a^{k-1} + ... + 1)  (mod 2^64)
# ----------------------------------------------------------------------------------------------------
_LCG_A = np.uint64(6364136223846793005)
_LCG_C = np.uint64(1442695040888963407)


def lcg_uniform(seed, count):
    with np.errstate(over="ignore"):
        a = np.full(count, _LCG_A, dtype=np.uint64)
        ak = np.cumprod(a)                                   # a^1 .. a^count (mod 2^64)
        geo = np.concatenate(([np.uint64(1)], ak[:-1]))      # a^0 .. a^(count-1)
        sk = ak * np.uint64(seed) + _LCG_C * np.cumsum(geo)  # s_1 .. s_count
    return (sk >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def _water_geometry(o, u):
    """o[nw,3] oxygen positions, u[nw,6] uniforms (3 jitter already applied by caller, 3 orientation)."""
    nw = o.shape[0]
    ct = 1.0 - 2.0 * u[:, 3]
    st = np.sqrt(np.maximum(0.0, 1.0 - ct * ct))
    ph = 2.0 * np.pi * u[:, 4]
    psi = 2.0 * np.pi * u[:, 5]
    b = np.stack([st * np.cos(ph), st * np.sin(ph), ct], axis=1)          # bisector
    ref = np.where(np.abs(b[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    e1 = np.cross(b, ref)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(b, e1)
    p = e1 * np.cos(psi)[:, None] + e2 * np.sin(psi)[:, None]             # in-plane normal to bisector
    half = np.deg2rad(TIP3P_HOH) / 2.0
    h1 = o + TIP3P_ROH * (np.cos(half) * b + np.sin(half) * p)
    h2 = o + TIP3P_ROH * (np.cos(half) * b - np.sin(half) * p)
    xyz = np.empty((nw, 3, 3))
    xyz[:, 0], xyz[:, 1], xyz[:, 2] = o, h1, h2
    return xyz.reshape(3 * nw, 3)


def water_lattice(ncell, a, seed=12345, jitter=0.2, keep=None):
    """ncell^3 TIP3P waters on a cubic lattice in a box of side a.  Returns (xyz[3nw,3], nw).
    keep: optional boolean mask over the ncell^3 lattice sites (uniforms are drawn for all sites so that
    the positions of the kept waters do not depend on the mask)."""
    nw = ncell ** 3
    sp = a / ncell
    u = lcg_uniform(seed, 6 * nw).reshape(nw, 6)
    g = np.arange(ncell)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    site = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    o = (site + 0.5) * sp + (2.0 * u[:, :3] - 1.0) * jitter
    if keep is not None:
        o, u = o[keep], u[keep]
    return _water_geometry(o, u), o.shape[0]


def _water_topology(nw, first=0):
    o = first + 3 * np.arange(nw, dtype=np.int32)
    excl = np.concatenate([np.stack([o + 1, o], 1), np.stack([o + 2, o], 1), np.stack([o + 2, o + 1], 1)])
    return excl.astype(np.int32)


def _finish(xyz, charges, ljtypes, eps, sigma, style, excl, pairs14, a, name, eps14=None, sigma14=None, scale14=1.0):
    tindex, tA, tB = make_lj_table(eps, sigma, style)
    sysd = dict(name=name, n=int(xyz.shape[0]),
                xyz=np.ascontiguousarray(xyz, dtype=np.float64),
                charges=np.ascontiguousarray(charges, dtype=np.float64),
                ljtypes=np.ascontiguousarray(ljtypes, dtype=np.int32),
                ntypes=int(len(eps)), tableindex=tindex, tableA=tA, tableB=tB,
                exclusions=np.ascontiguousarray(excl, dtype=np.int32).reshape(-1, 2),
                pairs14=np.ascontiguousarray(pairs14, dtype=np.int32).reshape(-1, 2),
                electrostaticScale14=float(scale14))
    if eps14 is None:
        eps14, sigma14 = eps, sigma
    t14 = make_lj_table(eps14, sigma14, style)
    sysd.update(tableindex14=t14[0], tableA14=t14[1], tableB14=t14[2])
    if a is None:
        sysd.update(box=None, rot=np.zeros((0, 3, 3)), trans=np.zeros((0, 3)))
    else:
        box = np.array([a, a, a, 90.0, 90.0, 90.0]) if np.isscalar(a) else np.asarray(a, dtype=np.float64)
        sysd.update(box=box, rot=np.eye(3)[None].copy(), trans=np.zeros((1, 3)))      # P1: identity only
    return sysd


def water_box(ncell=6, a=None, seed=12345, jitter=0.2, name=None):
    """Config 1 / 5 family: ncell^3 TIP3P waters, OPLS (geometric) LJ table, P1 cubic box.
    ncell=6 -> W216 (648 atoms, a=18.63); ncell=20 -> 24 000 atoms (a=62.1); ncell=70 -> M1 (1 029 000 atoms)."""
    if a is None:
        a = 3.105 * ncell
    xyz, nw = water_lattice(ncell, a, seed, jitter)
    charges = np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nw)
    ljtypes = np.tile([0, 1, 1], nw)
    return _finish(xyz, charges, ljtypes, [TIP3P_EPS_O, 0.0], [TIP3P_SIG_O, 0.0], "opls",
                   _water_topology(nw), np.zeros((0, 2), np.int32), a, name or "water%d" % nw, scale14=0.5)


def _chain_topology(first, n):
    i = first + np.arange(n, dtype=np.int32)
    e12 = np.stack([i[1:], i[:-1]], 1)
    e13 = np.stack([i[2:], i[:-2]], 1)
    e14 = np.stack([i[3:], i[:-3]], 1)
    return np.concatenate([e12, e13, e14]), e14


def _snake_in_sphere(center, radius, spacing, count):
    """Boustrophedon path over cubic lattice sites inside a sphere: consecutive atoms are lattice neighbours
    inside a row; rows and planes are traversed back and forth."""
    m = int(np.ceil(radius / spacing))
    pts = []
    fy = 1
    for kx in range(-m, m + 1):
        ys = range(-m, m + 1) if fy > 0 else range(m, -m - 1, -1)
        fz = 1
        for ky in ys:
            zs = range(-m, m + 1) if fz > 0 else range(m, -m - 1, -1)
            row = [(kx, ky, kz) for kz in zs if (kx * kx + ky * ky + kz * kz) * spacing * spacing <= radius * radius]
            if row:
                pts.extend(row)
                fz = -fz
        fy = -fy
    pts = np.array(pts, dtype=np.float64) * spacing + center
    if pts.shape[0] < count:
        raise ValueError("sphere too small for chain: %d < %d" % (pts.shape[0], count))
    return pts[:count]


def jac_like(natoms=23558, nchain=2489, ncell=20, a=62.23, seed=12345, ntypes=35):
    """Config 3 stand-in for DHFR/JAC (benchmarks/data/dhfr/systemData.yaml:3-5: 23 558 atoms, a = 62.23):
    a TIP3P lattice with a central sphere of waters replaced by an nchain-atom heteropolymer with
    ntypes-2 extra LJ types, 1-2/1-3/1-4 exclusions, a 1-4 list with its own LJ table and CHARMM/AMBER
    (arithmetic r_min) combination, net charge -11 like DHFR."""
    nwat = (natoms - nchain) // 3
    if 3 * nwat + nchain != natoms:
        raise ValueError("natoms - nchain must be a multiple of 3")
    sp = a / ncell
    g = (np.arange(ncell) + 0.5) * sp
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    d2 = (ix.ravel() - a / 2) ** 2 + (iy.ravel() - a / 2) ** 2 + (iz.ravel() - a / 2) ** 2
    order = np.argsort(d2, kind="stable")
    keep = np.ones(ncell ** 3, dtype=bool)
    nremove = ncell ** 3 - nwat
    keep[order[:nremove]] = False
    rcav = np.sqrt(d2[order[nremove - 1]])
    wxyz, nw = water_lattice(ncell, a, seed, 0.2, keep)
    assert nw == nwat
    spacing = 1.9
    cxyz = _snake_in_sphere(np.array([a / 2, a / 2, a / 2]), rcav - 1.6, spacing, nchain)
    u = lcg_uniform(seed + 77, 5 * nchain + 4 * ntypes).reshape(-1)
    cxyz = cxyz + (2.0 * u[:3 * nchain].reshape(nchain, 3) - 1.0) * 0.12
    ctype = 2 + np.minimum((u[3 * nchain:4 * nchain] * (ntypes - 2)).astype(np.int32), ntypes - 3)
    cq = (2.0 * u[4 * nchain:5 * nchain] - 1.0) * 0.55
    cq += (-11.0 - cq.sum()) / nchain
    ut = u[5 * nchain:].reshape(ntypes, 4)
    eps = np.empty(ntypes)
    rmin = np.empty(ntypes)
    eps[0], rmin[0] = TIP3P_EPS_O, TIP3P_SIG_O * 2.0 ** (1.0 / 6.0)      # CHARMM TIP3P
    eps[1], rmin[1] = 0.046 * KCAL, 2 * 0.2245
    eps[2:] = (0.02 + 0.18 * ut[2:, 0]) * KCAL
    rmin[2:] = 2.0 * (0.70 + 0.45 * ut[2:, 1])
    eps14 = eps.copy()
    rmin14 = rmin.copy()
    eps14[2:] *= 0.5 + 0.5 * ut[2:, 2]
    rmin14[2:] *= 0.9 + 0.1 * ut[2:, 3]
    xyz = np.concatenate([cxyz, wxyz])
    charges = np.concatenate([cq, np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nwat)])
    ljtypes = np.concatenate([ctype, np.tile([0, 1, 1], nwat)])
    cexcl, c14 = _chain_topology(0, nchain)
    excl = np.concatenate([cexcl, _water_topology(nwat, nchain)])
    return _finish(xyz, charges, ljtypes, eps, rmin, "amber", excl, c14, a, "jac%d" % natoms,
                   eps14=eps14, sigma14=rmin14, scale14=1.0)


def solvated_solute(solute_xyz, solute_q, solute_types, solute_eps, solute_rmin, bonds, ncell=9, a=27.945,
                    seed=12345, exclude_radius=2.8, name="bala_water"):
    """Config 2: a small solute (e.g. blocked alanine dipeptide) centred in an ncell^3 TIP3P box with
    overlapping waters removed, CHARMM-style (arithmetic) LJ, bonded 1-2/1-3 exclusions and a 1-4 list."""
    solute_xyz = np.asarray(solute_xyz, dtype=np.float64)
    ns = solute_xyz.shape[0]
    solute_xyz = solute_xyz - solute_xyz.mean(0) + a / 2
    sp = a / ncell
    wfull, nwfull = water_lattice(ncell, a, seed, 0.2)
    o = wfull[0::3]
    d = np.sqrt(((o[:, None, :] - solute_xyz[None, :, :]) ** 2).sum(-1)).min(1)
    keep = d > exclude_radius
    wxyz, nw = water_lattice(ncell, a, seed, 0.2, keep)
    # bonded graph distances on the solute
    adj = [set() for _ in range(ns)]
    for i, j in bonds:
        adj[i].add(j)
        adj[j].add(i)
    excl, p14 = set(), set()
    for i in range(ns):
        d1 = adj[i]
        d2 = set().union(*[adj[j] for j in d1]) - d1 - {i} if d1 else set()
        d3 = (set().union(*[adj[j] for j in d2]) if d2 else set()) - d2 - d1 - {i}
        for j in d1 | d2 | d3:
            excl.add((max(i, j), min(i, j)))
        for j in d3:
            p14.add((max(i, j), min(i, j)))
    excl = np.array(sorted(excl), dtype=np.int32).reshape(-1, 2)
    p14 = np.array(sorted(p14), dtype=np.int32).reshape(-1, 2)
    nst = len(solute_eps)
    eps = np.concatenate([[TIP3P_EPS_O, 0.046 * KCAL], np.asarray(solute_eps, dtype=np.float64)])
    rmin = np.concatenate([[TIP3P_SIG_O * 2.0 ** (1.0 / 6.0), 2 * 0.2245], np.asarray(solute_rmin, dtype=np.float64)])
    xyz = np.concatenate([solute_xyz, wxyz])
    charges = np.concatenate([np.asarray(solute_q, dtype=np.float64), np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nw)])
    ljtypes = np.concatenate([2 + np.asarray(solute_types, dtype=np.int32), np.tile([0, 1, 1], nw)])
    excl = np.concatenate([excl, _water_topology(nw, ns)])
    return _finish(xyz, charges, ljtypes, eps, rmin, "amber", excl, p14, a, name, scale14=1.0)


# ----------------------------------------------------------------------------------------------------
# systems built from the reference's own equilibrated 216-water box (fixture tests/golden/water216_cubicBox.npz,
# generated from book/data/mol/water216_cubicBox.mol by tests/golden/make_fixtures.py)
# ----------------------------------------------------------------------------------------------------
import os as _os

_GOLDEN = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "golden")


def _load_w216():
    d = np.load(_os.path.join(_GOLDEN, "water216_cubicBox.npz"))
    return np.array(d["xyz"], dtype=np.float64), float(d["a"])


def water216_real(box=None, name="w216", wrap=False):
    """Config 1: the shipped, equilibrated box of book Example 20 exactly as stored (molecules have diffused out of the
    primary cell, so more than the first shell of images is visited).  box: optional [a,b,c,alpha,beta,gamma] override."""
    xyz, a = _load_w216()
    nw = xyz.shape[0] // 3
    if wrap:
        mol = xyz.reshape(-1, 3, 3)
        xyz = (mol - np.floor(mol[:, 0:1, :] / a) * a).reshape(-1, 3)
    return _finish(xyz, np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nw), np.tile([0, 1, 1], nw), [TIP3P_EPS_O, 0.0], [TIP3P_SIG_O, 0.0], "opls",
                   _water_topology(nw), np.zeros((0, 2), np.int32), a if box is None else box, name, scale14=0.5)


def water216_example20(name="w216_mm"):
    """Book Example 20 as the reference sets it up: the 216-water box with the OPLS "bookSmallExamples" MM model, i.e. the NB terms of
    `w216` plus the flexible-water bonded terms (parameters/forceFields/opls/bookSmallExamples/harmonicBondParameters.yaml:26 HW-OW
    0.9572 A / 529.60, harmonicAngleParameters.yaml:34 HW-OW-HW 104.52 deg / 34.05, ureyBradleyParameters.yaml:20 HW..HW 1.5139 A / 38.25,
    kcal/mol units).  Known answer: the potential energy the reference prints for the stored coordinates,
    book/logs/Example20.log:70 (time 0): -8285.33515551 kJ/mol."""
    w = water216_real(name=name)
    nw = w["n"] // 3
    o = 3 * np.arange(nw, dtype=np.int32)
    bonds = np.stack([np.concatenate([o, o]), np.concatenate([o + 1, o + 2])], axis=1).astype(np.int32)
    w["bonded"] = dict(bonds=bonds, bond_eq=np.full(len(bonds), 0.9572), bond_fc=np.full(len(bonds), 529.60 * KCAL),
                       angles=np.stack([o + 1, o, o + 2], axis=1).astype(np.int32), angle_eq=np.full(nw, np.deg2rad(104.52)), angle_fc=np.full(nw, 34.05 * KCAL),
                       ureybradleys=np.stack([o + 1, o + 2], axis=1).astype(np.int32), ub_eq=np.full(nw, 1.5139), ub_fc=np.full(nw, 38.25 * KCAL))
    w["masses"] = np.tile([15.9994, 1.00794, 1.00794], nw)
    w["published_potential_energy"] = -8285.33515551
    return w


def replicated_water_xyz(nx, ny, nz, jitter=0.02, seed=4242):
    """The equilibrated box wrapped by molecule into [0,a)^3 and replicated nx x ny x nz times, plus a small jitter so
    that replicas are not exact copies.  Returns (xyz, nwaters, (ax, ay, az))."""
    xyz, a = _load_w216()
    mol = xyz.reshape(-1, 3, 3)
    shift = -np.floor(mol[:, 0:1, :] / a) * a
    mol = mol + shift
    reps = []
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                reps.append(mol + np.array([ix, iy, iz], dtype=np.float64) * a)
    out = np.concatenate(reps).reshape(-1, 3)
    if jitter > 0.0:
        u = lcg_uniform(seed, out.size).reshape(out.shape)
        out = out + (2.0 * u - 1.0) * jitter
    return out, out.shape[0] // 3, (a * nx, a * ny, a * nz)


def replicated_water(nx, ny=None, nz=None, jitter=0.02, name=None):
    """Config 5 family: nx*ny*nz replicas of the equilibrated 216-water box (12^3 -> 1 119 744 atoms, a = 223.69)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    xyz, nw, (ax, ay, az) = replicated_water_xyz(nx, ny, nz, jitter)
    return _finish(xyz, np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nw), np.tile([0, 1, 1], nw), [TIP3P_EPS_O, 0.0], [TIP3P_SIG_O, 0.0], "opls",
                   _water_topology(nw), np.zeros((0, 2), np.int32), [ax, ay, az, 90.0, 90.0, 90.0], name or "water%dx%dx%d" % (nx, ny, nz), scale14=0.5)


def jac_protein_water(natoms=23558, nchain=2489, ntypes=35, seed=12345):
    """Config 3: JAC/DHFR-sized system (benchmarks/data/dhfr/systemData.yaml:3-5: 23 558 atoms, 35 LJ types): 3x3x4 replicas
    of the equilibrated water box (orthorhombic 55.92 x 55.92 x 74.56) with a central sphere of waters replaced by an
    nchain-atom heteropolymer: ntypes-2 extra LJ types, CHARMM-style arithmetic combination, 1-2/1-3/1-4 exclusions, a 1-4
    list with its own LJ table, net charge -11 like DHFR."""
    nwat = (natoms - nchain) // 3
    if 3 * nwat + nchain != natoms:
        raise ValueError("natoms - nchain must be a multiple of 3")
    wxyz, nwfull, (ax, ay, az) = replicated_water_xyz(3, 3, 4)
    centre = np.array([ax, ay, az]) / 2.0
    o = wxyz[0::3]
    d2 = ((o - centre) ** 2).sum(1)
    order = np.argsort(d2, kind="stable")
    nremove = nwfull - nwat
    keep = np.ones(nwfull, dtype=bool)
    keep[order[:nremove]] = False
    rcav = np.sqrt(d2[order[nremove - 1]])
    wxyz = wxyz.reshape(-1, 3, 3)[keep].reshape(-1, 3)
    cxyz = _snake_in_sphere(centre, rcav - 1.5, 1.85, nchain)
    u = lcg_uniform(seed + 77, 5 * nchain + 4 * ntypes).reshape(-1)
    cxyz = cxyz + (2.0 * u[:3 * nchain].reshape(nchain, 3) - 1.0) * 0.12
    ctype = 2 + np.minimum((u[3 * nchain:4 * nchain] * (ntypes - 2)).astype(np.int32), ntypes - 3)
    cq = (2.0 * u[4 * nchain:5 * nchain] - 1.0) * 0.55
    cq += (-11.0 - cq.sum()) / nchain
    ut = u[5 * nchain:].reshape(ntypes, 4)
    eps, rmin = np.empty(ntypes), np.empty(ntypes)
    eps[0], rmin[0] = TIP3P_EPS_O, TIP3P_SIG_O * 2.0 ** (1.0 / 6.0)      # CHARMM TIP3P
    eps[1], rmin[1] = 0.046 * KCAL, 2 * 0.2245
    eps[2:] = (0.02 + 0.18 * ut[2:, 0]) * KCAL
    rmin[2:] = 2.0 * (0.70 + 0.45 * ut[2:, 1])
    eps14, rmin14 = eps.copy(), rmin.copy()
    eps14[2:] *= 0.5 + 0.5 * ut[2:, 2]
    rmin14[2:] *= 0.9 + 0.1 * ut[2:, 3]
    xyz = np.concatenate([cxyz, wxyz])
    charges = np.concatenate([cq, np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nwat)])
    ljtypes = np.concatenate([ctype, np.tile([0, 1, 1], nwat)])
    cexcl, c14 = _chain_topology(0, nchain)
    excl = np.concatenate([cexcl, _water_topology(nwat, nchain)])
    return _finish(xyz, charges, ljtypes, eps, rmin, "amber", excl, c14, [ax, ay, az, 90.0, 90.0, 90.0], "jac%d" % natoms,
                   eps14=eps14, sigma14=rmin14, scale14=1.0)


def bala_water(name="bala_water"):
    """Config 2: blocked alanine dipeptide (book/data/mol/bala_c7eq.mol via tests/golden/bala_c7eq.npz) solvated in 2x2x2
    replicas of the equilibrated water box (a = 37.282; waters whose O lies within 2.8 A of the solute removed; ~5.1k atoms),
    CHARMM-style LJ combination, bonded 1-2/1-3/1-4 exclusions and a 1-4 list.  The solute parameters are a fixed plausible
    table (parity is judged against the oracle on identical inputs)."""
    d = np.load(_os.path.join(_GOLDEN, "bala_c7eq.npz"))
    sxyz, sym, bonds = np.array(d["xyz"], dtype=np.float64), [str(x) for x in d["symbols"]], [tuple(int(v) for v in b) for b in d["bonds"]]
    table = {"C": (0.110 * KCAL, 4.00), "H": (0.022 * KCAL, 2.64), "O": (0.120 * KCAL, 3.40), "N": (0.200 * KCAL, 3.70)}
    elements = sorted(set(sym))
    stypes = np.array([elements.index(x) for x in sym], dtype=np.int32)
    q = np.array([{"C": 0.10, "H": 0.09, "O": -0.51, "N": -0.47}[x] for x in sym])
    q = q - q.mean()
    wxyz, nwfull, (ax, ay, az) = replicated_water_xyz(2, 2, 2)
    ns = len(sym)
    sxyz = sxyz - sxyz.mean(0) + np.array([ax, ay, az]) / 2.0
    o = wxyz[0::3]
    dmin = np.sqrt(((o[:, None, :] - sxyz[None, :, :]) ** 2).sum(-1)).min(1)
    wxyz = wxyz.reshape(-1, 3, 3)[dmin > 2.8].reshape(-1, 3)
    nw = wxyz.shape[0] // 3
    adj = [set() for _ in range(ns)]
    for i, j in bonds:
        adj[i].add(j)
        adj[j].add(i)
    excl, p14 = set(), set()
    for i in range(ns):
        d1 = adj[i]
        d2 = (set().union(*[adj[j] for j in d1]) if d1 else set()) - d1 - {i}
        d3 = (set().union(*[adj[j] for j in d2]) if d2 else set()) - d2 - d1 - {i}
        for j in d1 | d2 | d3:
            excl.add((max(i, j), min(i, j)))
        for j in d3:
            p14.add((max(i, j), min(i, j)))
    excl = np.array(sorted(excl), dtype=np.int32).reshape(-1, 2)
    p14 = np.array(sorted(p14), dtype=np.int32).reshape(-1, 2)
    eps = np.concatenate([[TIP3P_EPS_O, 0.046 * KCAL], [table[e][0] for e in elements]])
    rmin = np.concatenate([[TIP3P_SIG_O * 2.0 ** (1.0 / 6.0), 2 * 0.2245], [table[e][1] for e in elements]])
    xyz = np.concatenate([sxyz, wxyz])
    charges = np.concatenate([q, np.tile([TIP3P_QO, TIP3P_QH, TIP3P_QH], nw)])
    ljtypes = np.concatenate([2 + stypes, np.tile([0, 1, 1], nw)])
    excl = np.concatenate([excl, _water_topology(nw, ns)])
    return _finish(xyz, charges, ljtypes, eps, rmin, "amber", excl, p14, [ax, ay, az, 90.0, 90.0, 90.0], name, scale14=1.0)


def dhfr_jac(name="dhfr", bonded=False):
    """The reference's own JAC benchmark system (benchmarks/data/dhfr: 23 558 atoms, CHARMM22, cubic a = 62.23), from the
    fixture tests/golden/dhfr_jac.npz (tests/golden/make_fixtures.py).  Its per-term NB energies are published in
    benchmarks/log/systemBenchmarks_Serial_1ps.log:397-403 -- a known-answer vector for this path."""
    d = np.load(_os.path.join(_GOLDEN, "dhfr_jac.npz"))
    w = _finish(d["xyz"], d["charges"], d["ljtypes"], d["eps"], d["sigma"], "amber", d["exclusions"], d["pairs14"], float(d["a"]), name,
                eps14=d["eps14"], sigma14=d["sigma14"], scale14=1.0)
    w["published_energies"] = np.array(d["published_energies"])
    w["published_counts"] = np.array(d["published_counts"])
    bonded_path = _os.path.join(_GOLDEN, "dhfr_bonded.npz")
    if _os.path.exists(bonded_path):
        # bonded terms of the same input: per-term parameters, masses, the energies the reference publishes for them
        b = dict(np.load(bonded_path))
        w["masses"] = b.pop("masses")
        w["published_bonded"] = b.pop("published_bonded")
        w["published_total"] = b.pop("published_total")
        if bonded:
            w["bonded"] = b
    return w


CRYSTAL_NAMES = ("ALAALA", "ALAMET01", "AQARUF", "BEVXEF01", "GLYALB", "GLYGLY", "GUFQON", "HXACAN19", "IWANID", "LCDMPP10", "WIRYEB", "WABZOO")


def _bond_exclusions(n, bonds):
    adj = [set() for _ in range(n)]
    for i, j in bonds:
        adj[int(i)].add(int(j))
        adj[int(j)].add(int(i))
    excl, p14 = set(), set()
    for i in range(n):
        d1 = adj[i]
        d2 = (set().union(*[adj[j] for j in d1]) if d1 else set()) - d1 - {i}
        d3 = (set().union(*[adj[j] for j in d2]) if d2 else set()) - d2 - d1 - {i}
        for j in d1 | d2 | d3:
            excl.add((max(i, j), min(i, j)))
        for j in d3:
            p14.add((max(i, j), min(i, j)))
    return (np.array(sorted(excl), dtype=np.int32).reshape(-1, 2), np.array(sorted(p14), dtype=np.int32).reshape(-1, 2))


def crystal(name):
    """One of the 12 molecular crystals of pMolecule-1.9.0/tests/CrystalMMEnergies.py (fixture tests/golden/crystals.npz):
    genuine space-group operations (rotations, screw axes, inversions), tiny cells (tens to hundreds of images inside the
    13.5 A list cutoff), monoclinic / hexagonal / rhombohedral / triclinic lattices.  LJ tables are the topology's A/B
    coefficient tables; 1-4 pairs use half the LJ table and electrostaticScale14 = 0.5 (OPLS convention of the test)."""
    d = np.load(_os.path.join(_GOLDEN, "crystals.npz"))
    xyz, q, types, nbi = d[name + "_xyz"], d[name + "_q"], d[name + "_type"], d[name + "_nbindex"]
    nt = nbi.shape[0]
    # triangular table in the reference's layout (LJParameterContainer: tableindex[nt*nt] -> n(n+1)/2 entries)
    tindex = np.zeros(nt * nt, dtype=np.int32)
    tA, tB, k = np.zeros(nt * (nt + 1) // 2), np.zeros(nt * (nt + 1) // 2), 0
    for i in range(nt):
        for j in range(i + 1):
            tA[k], tB[k] = d[name + "_A"][nbi[i, j]], d[name + "_B"][nbi[i, j]]
            tindex[j + i * nt] = tindex[i + j * nt] = k
            k += 1
    excl, p14 = _bond_exclusions(len(q), d[name + "_bonds"])
    return dict(name=name, n=int(len(q)), xyz=np.ascontiguousarray(xyz, np.float64), charges=np.ascontiguousarray(q, np.float64),
                ljtypes=np.ascontiguousarray(types, np.int32), ntypes=int(nt), tableindex=tindex, tableA=tA, tableB=tB,
                tableindex14=tindex.copy(), tableA14=0.5 * tA, tableB14=0.5 * tB, exclusions=excl, pairs14=p14,
                electrostaticScale14=0.5, box=np.array(d[name + "_box"]), rot=np.array(d[name + "_rot"]), trans=np.array(d[name + "_trans"]))


def perturbed(system, amplitude, seed=999):
    """Copy of a system with every coordinate displaced uniformly in [-amplitude, amplitude] (for update-heuristic tests)."""
    s = dict(system)
    u = lcg_uniform(seed, 3 * system["n"]).reshape(-1, 3)
    s["xyz"] = system["xyz"] + (2.0 * u - 1.0) * amplitude
    return s


def ionic_fluid(nx=28, ny=28, nz=30, spacing=3.4, jitter=0.3, seed=777, name=None):
    """A bond-free Lennard-Jones + Coulomb fluid (alternating charges +-0.4 e on a jittered simple cubic lattice, two LJ types,
    masses 23 / 35.5 amu): the NB term is the WHOLE force field, so velocity-Verlet dynamics with NBModelABFS alone is physical
    (BASELINE config 4 needs bonded terms for water / protein boxes, which are out of scope).  28 x 28 x 30 = 23 520 atoms, JAC size."""
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    n = nx * ny * nz
    u = lcg_uniform(seed, 3 * n).reshape(n, 3)
    xyz = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], 1).astype(np.float64) * spacing + (2.0 * u - 1.0) * jitter + 0.5 * spacing
    kind = ((ix + iy + iz) & 1).ravel()
    out = _finish(xyz, np.where(kind == 0, 0.4, -0.4), kind, [0.4, 0.6], [2.6, 3.4], "opls", np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32),
                  [nx * spacing, ny * spacing, nz * spacing, 90.0, 90.0, 90.0], name or "ionic%dx%dx%d" % (nx, ny, nz))
    out["masses"] = np.where(kind == 0, 23.0, 35.5)
    return out


def with_fixed(system, fixed, name):
    """The same system with a set of fixed atoms (system.hardConstraints.fixedAtoms of the reference)."""
    out = dict(system)
    out["fixed"] = np.ascontiguousarray(fixed, np.int32)
    out["name"] = name
    return out


def _bala_fixed():
    w = bala_water()
    n = w["n"]
    # half of the solute and the oxygens of the solvent: pairs of two fixed atoms leave the lists, mixed pairs stay
    return with_fixed(w, np.concatenate([np.arange(0, 22, 2), np.arange(22, n, 3)]), "bala_fixed")


def _w216_fixed():
    w = water216_real()
    return with_fixed(w, np.arange(0, 324), "w216_fixed")          # the first 108 molecules do not move


WORKLOADS = {
    "w216": lambda: water216_real(),
    "bala_fixed": _bala_fixed,
    "w216_fixed": _w216_fixed,
    "w216_mm": water216_example20,                        # book Example 20: the same box with the OPLS flexible-water bonded terms
    "w216_lattice": lambda: water_box(6, name="w216_lattice"),
    "w216_triclinic": lambda: water216_real(box=[21.5, 22.0, 23.0, 85.0, 95.0, 100.0], name="w216_triclinic", wrap=True),
    "bala": lambda: bala_water(),
    "w1728_lattice": lambda: water_box(12, name="w1728_lattice"),
    "water3x3x3": lambda: replicated_water(3),
    "water4x4x4": lambda: replicated_water(4),           # 41 472 atoms: the bounded CPU sample of m1 (same generator, same density)
    "jac": lambda: jac_protein_water(),
    "dhfr": lambda: dhfr_jac(),
    "dhfr_mm": lambda: dhfr_jac("dhfr_mm", bonded=True),        # the same system with its bonded terms (the reference's complete CHARMM22 energy model)
    "jac_lattice": lambda: jac_like(),
    "water24k_lattice": lambda: water_box(20, name="water24k_lattice"),
    "m1": lambda: replicated_water(12, name="m1"),
    "ionic23k": lambda: ionic_fluid(name="ionic23k"),
    "ionic1k": lambda: ionic_fluid(10, 10, 10, name="ionic1k"),
}

for _c in CRYSTAL_NAMES:
    WORKLOADS["crystal_" + _c] = (lambda c=_c: crystal(c))

# cases with committed golden outputs of the compiled reference: name -> (maker, reference options, store full pair sets)
GOLDEN_CASES = {
    "w216": (WORKLOADS["w216"], {}, True),
    "w216_lattice": (WORKLOADS["w216_lattice"], {}, False),
    "w216_triclinic": (WORKLOADS["w216_triclinic"], {}, False),
    "w216_cut": (WORKLOADS["w216"], dict(dampingCutoff=1.0, innerCutoff=6.0, outerCutoff=9.0, listCutoff=10.5), False),
    "bala": (WORKLOADS["bala"], {}, False),
    "jac": (WORKLOADS["jac"], {}, False),
    "bala_fixed": (WORKLOADS["bala_fixed"], {}, False),
    "w216_fixed": (WORKLOADS["w216_fixed"], {}, False),
    # useCentering: the stored 216-water box has molecules outside the primary cell (42 images without, 13 with centring)
    "w216_centred": (WORKLOADS["w216"], dict(useCentering=True), False),
    "w216_triclinic_centred": (WORKLOADS["w216_triclinic"], dict(useCentering=True), False),
    "w216_fixed_centred": (WORKLOADS["w216_fixed"], dict(useCentering=True), False),
}
# spline form of the interaction (PairwiseInteractionABFS useAnalyticForm = False): default density, a coarse table with
# other cutoffs, 1-4 pairs + several LJ types, a crystal with rotations
GOLDEN_CASES["w216_spline"] = (WORKLOADS["w216"], dict(useAnalyticForm=False), False)
GOLDEN_CASES["w216_cut_spline"] = (WORKLOADS["w216"], dict(useAnalyticForm=False, splinePointDensity=20, dampingCutoff=1.0, innerCutoff=6.0, outerCutoff=9.0, listCutoff=10.5), False)
GOLDEN_CASES["bala_spline"] = (WORKLOADS["bala"], dict(useAnalyticForm=False), False)
for _c in CRYSTAL_NAMES:
    GOLDEN_CASES["crystal_" + _c] = (WORKLOADS["crystal_" + _c], {}, False)
GOLDEN_CASES["crystal_GLYGLY_spline"] = (WORKLOADS["crystal_GLYGLY"], dict(useAnalyticForm=False), False)
