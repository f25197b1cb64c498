"""TEST INFRASTRUCTURE (never imported by the product).  numpy restatement of the QC/MM entry points of NBModelABFS for a QC region
without boundary atoms, MM link-atom coupling, in vacuum or in a P1 cell (lattice translations only):

  NBModelABFS_QCMMEnergyLJ   pMolecule-1.9.0/extensions/csource/NBModelABFS.c:306-378   (PairwiseInteractionABFS_MMMMEnergy with charges off)
  NBModelABFS_QCMMPotentials NBModelABFS.c:449-498, PairwiseInteraction.c:612-663           (potentials on the QC atoms, atomic units)
  NBModelABFS_QCMMGradients  NBModelABFS.c:383-444, PairwiseInteraction.c:542-610           (regular units)

The reference walks pair lists; every pair within the outer cutoff is on a valid list and pairs beyond it are skipped
(PairwiseInteraction.h:72-78, PairwiseInteraction.c:654), so the sums do not depend on the lists: this restatement takes ALL pairs of a
QC atom with the atoms of the cell and of its translated copies.  The reference keeps one image of each inverse pair and lists both
(QC, MM') and (MM, QC') pairs for it (GenerateImageLists, NBModelABFS.c:753-1045), which is every translation once for the pairs
(QC, MM + s); image pairs of two QC atoms are counted with one half per translation.  Pinned by tests/test_oracle_qcmm.py against the
golden vectors of the compiled reference (tests/golden/golden_qcmm_*.npz).  This is the algorithm a device kernel for this row would
implement: the QC region is tens of atoms, so brute force over the extended atoms is the natural shape."""
import numpy as np

import oracle

HARTREE_KJ = 2625.5          # UNITS_ENERGY_HARTREES_TO_KILOJOULES_PER_MOLE (pCore-1.9.0/extensions/cinclude/Units.h:52)


def _shifts(w, outer):
    if w["box"] is None:
        return np.zeros((1, 3))
    M, invM = oracle.make_M(w["box"])                      # columns of M = lattice vectors
    frac = np.asarray(w["xyz"]) @ invM.T
    spread = np.ceil(frac.max(0) - frac.min(0)).astype(int)
    heights = 1.0 / np.linalg.norm(invM, axis=1)           # distance between the lattice planes of each axis
    k = np.ceil(outer / heights).astype(int) + spread
    rng = [np.arange(-k[d], k[d] + 1) for d in range(3)]
    abc = np.array(np.meshgrid(*rng, indexing="ij")).reshape(3, -1).T
    abc = np.concatenate([abc[(abc == 0).all(1)], abc[~(abc == 0).all(1)]])        # identity first
    return abc @ M.T


def qcmm(w, qc_index, qc_charges, damp=0.5, inner=8.0, outer=12.0, density=50, dielectric=1.0):
    """Returns dict(eqcmmlj, eimqcmmlj, eimqcqclj, potentials[nqc], grad_lj[n, 3], grad_el[n, 3])."""
    assert w["box"] is None or len(w["trans"]) == 1, "P1 cells only"
    x = np.asarray(w["xyz"], np.float64)
    n = len(x)
    qc = np.asarray(qc_index)
    isqc = np.zeros(n, bool)
    isqc[qc] = True
    mmq = np.where(isqc, 0.0, np.asarray(w["charges"], np.float64)) / dielectric      # mmCharges: active MM atoms only
    nt = w["ntypes"]
    ti = np.asarray(w["tableindex"]).reshape(nt, nt)
    tA, tB = np.asarray(w["tableA"]), np.asarray(w["tableB"])
    lt = np.asarray(w["ljtypes"])
    excl = set()
    for a, b in np.asarray(w["exclusions"]).reshape(-1, 2):
        excl.add((int(a), int(b))); excl.add((int(b), int(a)))
    f = oracle.make_factors(damp, inner, outer)
    sx, sy, sh = oracle.make_spline(3, damp, inner, outer, density)
    r2off = outer * outer
    e = dict(eqcmmlj=0.0, eimqcmmlj=0.0, eimqcqclj=0.0)
    pot = np.zeros(len(qc))
    g_lj, g_el = np.zeros((n, 3)), np.zeros((n, 3))
    for si, s in enumerate(_shifts(w, outer)):
        primary = si == 0
        for k, q in enumerate(qc):
            d = x[q] - x - s
            r2 = (d * d).sum(1)
            for j in np.nonzero(r2 <= r2off)[0]:
                j = int(j)
                if primary and (isqc[j] or (int(q), j) in excl):
                    continue                                # QC/QC pairs of the cell belong to the QC model; exclusions (qcmmExclusions)
                t = ti[lt[q], lt[j]]
                _, elj, dF = oracle.pair(f, float(r2[j]), 0.0, float(tA[t]), float(tB[t]))
                wgt = 0.5 if isqc[j] else 1.0
                e["eqcmmlj" if primary else ("eimqcqclj" if isqc[j] else "eimqcmmlj")] += wgt * elj
                g_lj[q] += wgt * 2.0 * dF * d[j]
                g_lj[j] -= wgt * 2.0 * dF * d[j]
                if not isqc[j] and mmq[j] != 0.0:
                    fe, dfe = oracle.spline_evaluate(sx, sy, sh, float(r2[j]))
                    pot[k] += mmq[j] * fe
                    c = HARTREE_KJ * qc_charges[k] * mmq[j] * 2.0 * dfe
                    g_el[q] += c * d[j]
                    g_el[j] -= c * d[j]
    e.update(potentials=pot, grad_lj=g_lj, grad_el=g_el)
    return e


def qcqc_images(w, qc_index, damp=0.5, inner=8.0, outer=12.0, density=50, dielectric=1.0):
    """The QC/QC image terms for a cell with space-group operations (MMMMImageEnergy on inbqcqclj, NBModelABFS.c:356-376; QCQCImagePotentials,
    NBModelABFS.c:470-476): images x' = (M S M^-1) x + M (t + (a, b, c)) (NBModelABFS.c:1161-1313) of the QC atoms, every operation except the
    identity with weight one half (the reference keeps one operation of each inverse pair with scale 1 and self-inverse ones with scale 0.5).
    Returns dict(eimqcqclj, qcqc_potentials (packed lower triangle, atomic units), grad_lj[n, 3])."""
    x = np.asarray(w["xyz"], np.float64)
    qc = np.asarray(qc_index)
    nq = len(qc)
    M, invM = oracle.make_M(w["box"])
    frac = x @ invM.T
    spread = np.ceil(frac.max(0) - frac.min(0)).astype(int) + 1          # the operations move the molecule inside the cell
    heights = 1.0 / np.linalg.norm(invM, axis=1)
    k = np.ceil(outer / heights).astype(int) + spread
    rng = [np.arange(-k[d], k[d] + 1) for d in range(3)]
    abc = np.array(np.meshgrid(*rng, indexing="ij")).reshape(3, -1).T
    nt = w["ntypes"]
    ti = np.asarray(w["tableindex"]).reshape(nt, nt)
    tA, tB = np.asarray(w["tableA"]), np.asarray(w["tableB"])
    lt = np.asarray(w["ljtypes"])
    f = oracle.make_factors(damp, inner, outer)
    sx, sy, sh = oracle.make_spline(3, damp, inner, outer, density)
    r2off = outer * outer
    e, W, g = 0.0, np.zeros((nq, nq)), np.zeros((len(x), 3))
    for S, t in zip(np.asarray(w["rot"]).reshape(-1, 3, 3), np.asarray(w["trans"]).reshape(-1, 3)):
        R = M @ S @ invM
        identity_op = np.array_equal(S, np.eye(3)) and not np.any(t)
        base = x[qc] @ R.T
        for s in abc:
            if identity_op and not np.any(s):
                continue
            xi = base + M @ (t + s)
            d = x[qc][:, None, :] - xi[None, :, :]                         # primary q (rows) against image q' (columns)
            r2 = (d * d).sum(2)
            for a, b in zip(*np.nonzero(r2 <= r2off)):
                tt = ti[lt[qc[a]], lt[qc[b]]]
                _, elj, dF = oracle.pair(f, float(r2[a, b]), 0.0, float(tA[tt]), float(tB[tt]))
                e += 0.5 * elj
                g[qc[a]] += 0.5 * 2.0 * dF * d[a, b]
                g[qc[b]] -= 0.5 * 2.0 * dF * (R.T @ d[a, b])
                W[a, b] += 0.5 * oracle.spline_evaluate(sx, sy, sh, float(r2[a, b]))[0] / dielectric
    W = 0.5 * (W + W.T)
    return dict(eimqcqclj=e, qcqc_potentials=W[np.tril_indices(nq)], grad_lj=g)
