/* oracle/mmterms_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's bonded MM terms (flat arrays, per-term parameters).
 * Pinned against the compiled reference (oracle/ref_driver.c: refmm_energy) and the published DHFR energies by tests/test_oracle.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.  Reference paths: pM = pMolecule-1.9.0/extensions. */
#include <math.h>
#include <stddef.h>
#include "nbabfs_oracle.h"

/* HarmonicBondContainer_Energy, pM/csource/HarmonicBondContainer.c:149-183 */
static double bonds_energy(int nt, const int *a, const double *eq, const double *fc, const double *x, double *g)
{
    double energy = 0.0e+00;
    int n;
    for (n = 0; n < nt; n++) {
        int i = a[2 * n], j = a[2 * n + 1];
        double xij = x[3 * i] - x[3 * j], yij = x[3 * i + 1] - x[3 * j + 1], zij = x[3 * i + 2] - x[3 * j + 2];
        double rij = sqrt(xij * xij + yij * yij + zij * zij), disp = rij - eq[n], df = fc[n] * disp;
        energy += (df * disp);
        if (g != NULL) {
            df *= (2.0e+00 / rij);
            xij *= df; yij *= df; zij *= df;
            g[3 * i] += xij; g[3 * i + 1] += yij; g[3 * i + 2] += zij;
            g[3 * j] -= xij; g[3 * j + 1] -= yij; g[3 * j + 2] -= zij;
        }
    }
    return energy;
}

/* HarmonicAngleContainer_Energy, pM/csource/HarmonicAngleContainer.c:157-205 (DOT_LIMIT 0.999999, :27) */
static double angles_energy(int nt, const int *a, const double *eq, const double *fc, const double *x, double *g)
{
    double energy = 0.0e+00;
    int n;
    for (n = 0; n < nt; n++) {
        int i = a[3 * n], j = a[3 * n + 1], k = a[3 * n + 2];
        double xij = x[3 * i] - x[3 * j], yij = x[3 * i + 1] - x[3 * j + 1], zij = x[3 * i + 2] - x[3 * j + 2];
        double xkj = x[3 * k] - x[3 * j], ykj = x[3 * k + 1] - x[3 * j + 1], zkj = x[3 * k + 2] - x[3 * j + 2];
        double rij = sqrt(xij * xij + yij * yij + zij * zij), rkj = sqrt(xkj * xkj + ykj * ykj + zkj * zkj), dot, theta, disp, df;
        xij /= rij; yij /= rij; zij /= rij;
        xkj /= rkj; ykj /= rkj; zkj /= rkj;
        dot = xij * xkj + yij * ykj + zij * zkj;
        dot = (-0.999999 > dot) ? -0.999999 : dot;
        dot = (0.999999 < dot) ? 0.999999 : dot;
        theta = acos(dot);
        disp = theta - eq[n];
        df = fc[n] * disp;
        energy += (df * disp);
        if (g != NULL) {
            double dtdx = -1.0 / sqrt(1.0 - dot * dot), dtxi, dtyi, dtzi, dtxk, dtyk, dtzk;
            df *= (2.0e+00 * dtdx);
            dtxi = df * (xkj - dot * xij) / rij; dtyi = df * (ykj - dot * yij) / rij; dtzi = df * (zkj - dot * zij) / rij;
            dtxk = df * (xij - dot * xkj) / rkj; dtyk = df * (yij - dot * ykj) / rkj; dtzk = df * (zij - dot * zkj) / rkj;
            g[3 * i] += dtxi; g[3 * i + 1] += dtyi; g[3 * i + 2] += dtzi;
            g[3 * k] += dtxk; g[3 * k + 1] += dtyk; g[3 * k + 2] += dtzk;
            g[3 * j] -= (dtxi + dtxk); g[3 * j + 1] -= (dtyi + dtyk); g[3 * j + 2] -= (dtzi + dtzk);
        }
    }
    return energy;
}

/* common to FourierDihedralContainer_Energy (pM/csource/FourierDihedralContainer.c:156-241) and HarmonicImproperContainer_Energy
 * (pM/csource/HarmonicImproperContainer.c:172-258): which = 0 Fourier {fc, period, phase}, 1 improper {fc, eq} */
static double torsions_energy(int which, int nt, const int *a, const double *fc, const int *period, const double *angle, const double *x, double *g)
{
    double energy = 0.0e+00;
    int n, p;
    for (n = 0; n < nt; n++) {
        int i = a[4 * n], j = a[4 * n + 1], k = a[4 * n + 2], l = a[4 * n + 3];
        double xij = x[3 * i] - x[3 * j], yij = x[3 * i + 1] - x[3 * j + 1], zij = x[3 * i + 2] - x[3 * j + 2];
        double xkj = x[3 * k] - x[3 * j], ykj = x[3 * k + 1] - x[3 * j + 1], zkj = x[3 * k + 2] - x[3 * j + 2];
        double xlk = x[3 * l] - x[3 * k], ylk = x[3 * l + 1] - x[3 * k + 1], zlk = x[3 * l + 2] - x[3 * k + 2];
        double rkj2 = xkj * xkj + ykj * ykj + zkj * zkj, rkj = sqrt(rkj2);
        double mx = yij * zkj - zij * ykj, my = zij * xkj - xij * zkj, mz = xij * ykj - yij * xkj;
        double nx = ylk * zkj - zlk * ykj, ny = zlk * xkj - xlk * zkj, nz = xlk * ykj - ylk * xkj;
        double m2 = mx * mx + my * my + mz * mz, n2 = nx * nx + ny * ny + nz * nz, mn = sqrt(m2 * n2);
        double cosphi = (mx * nx + my * ny + mz * nz) / mn, sinphi = rkj * (xij * nx + yij * ny + zij * nz) / mn, df;
        if (which == 0) {
            double cosnphi = 1.0e+00, sinnphi = 0.0e+00, temp, cosphase = cos(angle[n]), sinphase = sin(angle[n]);
            for (p = 1; p <= period[n]; p++) {
                temp = cosnphi * cosphi - sinnphi * sinphi;
                sinnphi = cosnphi * sinphi + sinnphi * cosphi;
                cosnphi = temp;
            }
            energy += fc[n] * (1.0e+00 + cosnphi * cosphase + sinnphi * sinphase);
            df = fc[n] * ((double) period[n]) * (cosnphi * sinphase - sinnphi * cosphase);
        } else {
            double coseq = cos(angle[n]), sineq = sin(angle[n]), dphi;
            double cosdphi = cosphi * coseq + sinphi * sineq, sindphi = sinphi * coseq - cosphi * sineq;
            if (cosdphi > 0.1e+00) dphi = asin(sindphi);                                   /* LOWCOSPHI, :171 */
            else {
                dphi = fabs(acos((cosdphi > -1.0e+00) ? cosdphi : -1.0e+00));
                if (sindphi < 0.0e+00) dphi *= -1.0e+00;
            }
            df = fc[n] * dphi;
            energy += df * dphi;
            df *= 2.0e+00;
        }
        if (g != NULL) {
            double dtxi = df * rkj * mx / m2, dtyi = df * rkj * my / m2, dtzi = df * rkj * mz / m2;
            double dtxl = -df * rkj * nx / n2, dtyl = -df * rkj * ny / n2, dtzl = -df * rkj * nz / n2;
            double dotij = xij * xkj + yij * ykj + zij * zkj, dotlk = xlk * xkj + ylk * ykj + zlk * zkj;
            double sx = (dotij * dtxi + dotlk * dtxl) / rkj2, sy = (dotij * dtyi + dotlk * dtyl) / rkj2, sz = (dotij * dtzi + dotlk * dtzl) / rkj2;
            g[3 * i] += dtxi; g[3 * i + 1] += dtyi; g[3 * i + 2] += dtzi;
            g[3 * j] += sx - dtxi; g[3 * j + 1] += sy - dtyi; g[3 * j + 2] += sz - dtzi;
            g[3 * k] += -sx - dtxl; g[3 * k + 1] += -sy - dtyl; g[3 * k + 2] += -sz - dtzl;
            g[3 * l] += dtxl; g[3 * l + 1] += dtyl; g[3 * l + 2] += dtzl;
        }
    }
    return energy;
}

void orc_mm_energy(int n, const double *xyz,
                   int nbond, const int *bonds, const double *bondEq, const double *bondFc,
                   int nangle, const int *angles, const double *angleEq, const double *angleFc,
                   int nub, const int *ubs, const double *ubEq, const double *ubFc,
                   int ndih, const int *dihedrals, const double *dihFc, const int *dihPeriod, const double *dihPhase,
                   int nimp, const int *impropers, const double *impEq, const double *impFc,
                   double *energies5, double *grad)
{
    (void) n;
    energies5[0] = bonds_energy(nbond, bonds, bondEq, bondFc, xyz, grad);
    energies5[1] = angles_energy(nangle, angles, angleEq, angleFc, xyz, grad);
    energies5[2] = bonds_energy(nub, ubs, ubEq, ubFc, xyz, grad);
    energies5[3] = torsions_energy(0, ndih, dihedrals, dihFc, dihPeriod, dihPhase, xyz, grad);
    energies5[4] = torsions_energy(1, nimp, impropers, impFc, NULL, impEq, xyz, grad);
}
