/* oracle/nbabfs_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference algorithm.
 *
 * A plain-C, flat-array restatement of pDynamo 1.9.0's NBModelABFS MM/MM hot path.  It is the parity
 * checker for the CUDA product; it is never linked into, imported by or executed from the product.
 * Pinning: tests/test_oracle.py checks every function below against the compiled, unmodified reference
 * (oracle/_ref/libref_nbabfs.so) and against fixtures generated from it (tests/golden/).
 *
 * All paths below are relative to /root/reference.  pM = pMolecule-1.9.0/extensions, pC = pCore-1.9.0/extensions.
 * All arithmetic is fp64 and must be compiled with -ffp-contract=off (the reference is built with plain -O2
 * and no -march flag, i.e. without FMA contraction: installation/InstallUtilities.py:47-49).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "nbabfs_oracle.h"

/* pC/cinclude/Units.h:18-49 */
#define ORC_AVOGADRO 6.0221415e+23
#define ORC_ECHARGE  1.60217653e-19
#define ORC_EPS0     8.854187817e-12
#define ORC_E2A_TO_KJMOL ((1.0e+7 * ORC_AVOGADRO * ORC_ECHARGE * ORC_ECHARGE) / (4.0e+00 * M_PI * ORC_EPS0))
#define ORC_ANGSTROMS_TO_BOHRS (1.0e-10 / 5.291772083e-11)   /* pC/cinclude/Units.h:21,55-57 */
#define ORC_DEG2RAD (M_PI / 180.0e+00)

typedef struct {
    int t, a, b, c;
    double scale;
    long npairs;
    int *pairs;      /* (i, j) ordered: i primary atom, j image atom */
    double *xyz;     /* list-time image coordinates */
} OrcImage;

struct OrcNB {
    int n, ntypes, ntypes14, ntrans, identity;
    double *q;
    int *ljtype;
    int *tindex, *tindex14;
    double *tA, *tB, *tA14, *tB14;
    int *exclPtr, *exclCol;           /* symmetric CSR of exclusions (pC/csource/PairList.c:458-526) */
    long n14, n14all;
    int *p14, *p14all;               /* 1-4 pairs after / before the fixed-atom filter */
    char *fixed;                      /* NULL or flags: fixed atoms (freeSelection = complement, NBModelABFSState.c:345) */
    int centering, nisolates;         /* useCentering: isolates = connected components of the exclusion graph */
    int *isoPtr, *isoIdx;
    double *xc, *isoT;                /* centred coordinates, per-atom isolate translations (NBModelABFSState.c:278-311) */
    double *rot, *trans;
    int *inverses;
    /* options */
    double damp, inner, outer, list, dielectric, scale14;
    int checkForInverses, expandFactor;
    int useAnalytic, density;         /* PairwiseInteractionABFS.useAnalyticForm / splinePointDensity */
    int nsp;                          /* spline points */
    double *spx, *spy[3], *sph[3];    /* shared abscissae x = r^2; ordinates and second derivatives of the electrostatic, LJ-A, LJ-B splines */
    /* state */
    int isNew;
    double stListCutoff, stOuterCutoff;
    double *xref;
    double refM[9];
    int haveRefM;
    long nprimary;
    int *primary;
    int nimages;
    OrcImage *images;
};

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1.0e-9 * (double) ts.tv_nsec;
}

/* ---------------------------------------------------------------------------------------------------
 * 3x3 helpers.  Row-major m[3*r+c].
 * Matrix33_Determinant pC/csource/Matrix33.c:110-124 ; Matrix33_Invert :190-219 ;
 * Matrix33_PostMultiplyBy :348-377 ; Matrix33_PreMultiplyBy :382-408 ; Matrix33_ApplyToVector3 :23-43
 * ------------------------------------------------------------------------------------------------- */
static double m33_det(const double *m)
{
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[6] * m[5]) + m[2] * (m[3] * m[7] - m[6] * m[4]);
}

static void m33_invert(double *s, const double *o)
{
    double a00 = o[0], a01 = o[1], a02 = o[2], a10 = o[3], a11 = o[4], a12 = o[5], a20 = o[6], a21 = o[7], a22 = o[8];
    double f;
    int i;
    s[0] = (a11 * a22 - a12 * a21); s[1] = (a02 * a21 - a22 * a01); s[2] = (a01 * a12 - a11 * a02);
    s[3] = (a12 * a20 - a10 * a22); s[4] = (a00 * a22 - a02 * a20); s[5] = (a02 * a10 - a00 * a12);
    s[6] = (a10 * a21 - a11 * a20); s[7] = (a01 * a20 - a00 * a21); s[8] = (a00 * a11 - a10 * a01);
    f = 1.0e+00 / m33_det(o);
    for (i = 0; i < 9; i++) s[i] *= f;
}

static void m33_postmul(double *s, const double *o)   /* s = s . o */
{
    double s00 = s[0], s01 = s[1], s02 = s[2], s10 = s[3], s11 = s[4], s12 = s[5], s20 = s[6], s21 = s[7], s22 = s[8];
    int i;
    for (i = 0; i < 3; i++) {
        double o0 = o[i], o1 = o[3 + i], o2 = o[6 + i];
        s[i]     = s00 * o0 + s01 * o1 + s02 * o2;
        s[3 + i] = s10 * o0 + s11 * o1 + s12 * o2;
        s[6 + i] = s20 * o0 + s21 * o1 + s22 * o2;
    }
}

static void m33_premul(double *s, const double *o)    /* s = o . s */
{
    double s00 = s[0], s01 = s[1], s02 = s[2], s10 = s[3], s11 = s[4], s12 = s[5], s20 = s[6], s21 = s[7], s22 = s[8];
    int i;
    for (i = 0; i < 3; i++) {
        double o0 = o[3 * i], o1 = o[3 * i + 1], o2 = o[3 * i + 2];
        s[3 * i]     = s00 * o0 + s10 * o1 + s20 * o2;
        s[3 * i + 1] = s01 * o0 + s11 * o1 + s21 * o2;
        s[3 * i + 2] = s02 * o0 + s12 * o1 + s22 * o2;
    }
}

static void m33_apply(const double *m, double *v)
{
    double x = v[0], y = v[1], z = v[2];
    v[0] = x * m[0] + y * m[1] + z * m[2];
    v[1] = x * m[3] + y * m[4] + z * m[5];
    v[2] = x * m[6] + y * m[7] + z * m[8];
}

/* Matrix33_InverseDerivative pC/csource/Matrix33.c:130-183 */
static void m33_inverse_derivative(const double *m, int i, int j, double *o)
{
    double m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    double ddet = 0.0, det, f;
    int k;
    m33_invert(o, m);
    switch (3 * i + j) {
        case 0: ddet = -m12 * m21 + m11 * m22; break;
        case 1: ddet =  m12 * m20 - m10 * m22; break;
        case 2: ddet = -m11 * m20 + m10 * m21; break;
        case 3: ddet =  m02 * m21 - m01 * m22; break;
        case 4: ddet = -m02 * m20 + m00 * m22; break;
        case 5: ddet =  m01 * m20 - m00 * m21; break;
        case 6: ddet = -m02 * m11 + m01 * m12; break;
        case 7: ddet =  m02 * m10 - m00 * m12; break;
        case 8: ddet = -m01 * m10 + m00 * m11; break;
    }
    f = -ddet;
    for (k = 0; k < 9; k++) o[k] *= f;
    switch (3 * i + j) {
        case 0: o[4] += m22; o[5] -= m12; o[7] -= m21; o[8] += m11; break;
        case 1: o[1] -= m22; o[2] += m12; o[7] += m20; o[8] -= m10; break;
        case 2: o[1] += m21; o[2] -= m11; o[4] -= m20; o[5] += m10; break;
        case 3: o[3] -= m22; o[5] += m02; o[6] += m21; o[8] -= m01; break;
        case 4: o[0] += m22; o[2] -= m02; o[6] -= m20; o[8] += m00; break;
        case 5: o[0] -= m21; o[2] += m01; o[3] += m20; o[5] -= m00; break;
        case 6: o[3] += m12; o[4] -= m02; o[6] -= m11; o[7] += m01; break;
        case 7: o[0] -= m12; o[1] += m02; o[6] += m10; o[7] -= m00; break;
        case 8: o[0] += m11; o[1] -= m01; o[3] -= m10; o[4] += m00; break;
    }
    det = m33_det(m);
    f = 1.0e+00 / det;
    for (k = 0; k < 9; k++) o[k] *= f;
}

/* SymmetryParameters_MakeM pM/csource/SymmetryParameters.c:370-409 */
void orc_make_M(const double *box, double *M, double *invM)
{
    double alpha = box[3] * ORC_DEG2RAD, beta = box[4] * ORC_DEG2RAD, gamma = box[5] * ORC_DEG2RAD;
    double cosalpha = cos(alpha), cosbeta = cos(beta), cosgamma = cos(gamma), singamma = sin(gamma);
    int i;
    for (i = 0; i < 9; i++) M[i] = 0.0e+00;
    M[0] = box[0];
    M[1] = box[1] * cosgamma;
    M[4] = box[1] * singamma;
    M[2] = box[2] * cosbeta;
    M[5] = box[2] * (cosalpha - cosbeta * cosgamma) / singamma;
    M[8] = box[2] * sqrt(1.0e+00 - cosalpha * cosalpha - cosbeta * cosbeta - cosgamma * cosgamma +
                         2.0e+00 * cosalpha * cosbeta * cosgamma) / singamma;
    m33_invert(invM, M);
}

/* PairwiseInteractionABFS_MakeFactors pM/csource/PairwiseInteraction.c:89-141.
 * out: r2Damp r2On r2Off | a b c d qShift1 qShift2 qF0 qAlpha | aF6 aK12 aShift12 aF0 aAlpha | bF3 bK6 bShift6 bF0 bAlpha */
enum { R2DAMP, R2ON, R2OFF, FA, FB, FC, FD, QSHIFT1, QSHIFT2, QF0, QALPHA, AF6, AK12, ASHIFT12, AF0, AALPHA, BF3, BK6, BSHIFT6, BF0, BALPHA };

void orc_make_factors(double damp, double inner, double outer, double *o)
{
    double f, g, gamma, s12, s6;
    o[R2DAMP] = damp * damp;
    o[R2OFF]  = outer * outer;
    o[R2ON]   = inner * inner;
    gamma     = pow(o[R2OFF] - o[R2ON], 3);
    o[FA] = o[R2OFF] * o[R2OFF] * (o[R2OFF] - 3.0e+00 * o[R2ON]) / gamma;
    o[FB] = 6.0e+00 * o[R2OFF] * o[R2ON] / gamma;
    o[FC] = -(o[R2OFF] + o[R2ON]) / gamma;
    o[FD] = 0.4e+00 / gamma;
    o[QSHIFT1] = 8.0e+00 * (o[R2OFF] * o[R2ON] * (outer - inner) -
                 0.2e+00 * (outer * o[R2OFF] * o[R2OFF] - inner * o[R2ON] * o[R2ON])) / gamma;
    o[QSHIFT2] = -(o[FA] / outer) + o[FB] * outer + o[FC] * outer * o[R2OFF] + o[FD] * outer * o[R2OFF] * o[R2OFF];
    f = 1.0e+00 / damp + o[QSHIFT1];
    g = -1.0e+00 / o[R2DAMP];
    o[QF0]    = f - 0.5e+00 * damp * g;
    o[QALPHA] = -0.5e+00 * g / damp;
    o[AF6]      = 1.0e+00 / (o[R2OFF] * o[R2OFF] * o[R2OFF]);
    o[AK12]     = pow(o[R2OFF], 3) / (pow(o[R2OFF], 3) - pow(o[R2ON], 3));
    o[ASHIFT12] = 1.0e+00 / pow(inner * outer, 6);
    s12 = 1.0e+00 / pow(damp, 12);
    f   = s12 - o[ASHIFT12];
    g   = -12.0e+00 * s12 / damp;
    o[AF0]    = f - 0.5e+00 * damp * g;
    o[AALPHA] = -0.5e+00 * g / damp;
    o[BF3]     = 1.0e+00 / (outer * o[R2OFF]);
    o[BK6]     = (outer * o[R2OFF]) / (outer * o[R2OFF] - inner * o[R2ON]);
    o[BSHIFT6] = 1.0e+00 / pow(inner * outer, 3);
    s6 = 1.0e+00 / pow(damp, 6);
    f  = -s6 + o[BSHIFT6];
    g  = 6.0e+00 * s6 / damp;
    o[BF0]    = f - 0.5e+00 * damp * g;
    o[BALPHA] = -0.5e+00 * g / damp;
}

/* One pair: macros PairwiseInteractionABFS_CheckDistances / _ElectrostaticTerm / _LennardJonesTerm,
 * pM/cinclude/PairwiseInteraction.h:72-119.  Returns 0 if the pair is skipped (r2 > r2Off). */
static inline int pair_terms(const double *F, double r2, double qij, double Aij, double Bij, double *eq, double *elj, double *dFout)
{
    double s, s2, s6, dF = 0.0e+00;
    if (r2 > F[R2OFF]) return 0;
    else if (r2 < F[R2DAMP]) { s2 = 0.0e+00; s = 0.0e+00; }
    else { s2 = 1.0e+00 / r2; s = sqrt(s2); }
    if (r2 > F[R2ON]) {
        *eq = qij * (s * (F[FA] - r2 * (F[FB] + r2 * (F[FC] + F[FD] * r2))) + F[QSHIFT2]);
        dF += -qij * 0.5e+00 * s * (F[FA] + r2 * (F[FB] + r2 * (3.0e+00 * F[FC] + 5.0e+00 * F[FD] * r2))) / r2;
    } else if (r2 > F[R2DAMP]) {
        *eq = qij * (s + F[QSHIFT1]);
        dF += -qij * 0.5e+00 * s / r2;
    } else {
        *eq = qij * (F[QF0] - F[QALPHA] * r2);
        dF += -qij * F[QALPHA];
    }
    s6 = s2 * s2 * s2;
    if (r2 > F[R2ON]) {
        double l1 = s6 - F[AF6], l2 = (s / r2) - F[BF3];
        *elj = Aij * F[AK12] * pow(l1, 2) - Bij * F[BK6] * pow(l2, 2);
        dF += -3.0e+00 * s6 * (2.0e+00 * Aij * F[AK12] * l1 / r2 - Bij * F[BK6] * l2 / s);
    } else if (r2 > F[R2DAMP]) {
        *elj = Aij * (s6 * s6 - F[ASHIFT12]) - Bij * (s6 - F[BSHIFT6]);
        dF += -3.0e+00 * s6 * (2.0e+00 * Aij * s6 - Bij) / r2;
    } else {
        *elj = Aij * (F[AF0] - F[AALPHA] * r2) - Bij * (F[BF0] - F[BALPHA] * r2);
        dF += -Aij * F[AALPHA] + Bij * F[BALPHA];
    }
    *dFout = dF;
    return 1;
}

void orc_pair(const double *f21, double r2, double qij, double Aij, double Bij, double *e2, double *dF)
{
    e2[0] = e2[1] = 0.0; *dF = 0.0;
    pair_terms(f21, r2, qij, Aij, Bij, &e2[0], &e2[1], dF);
}

/* ---------------------------------------------------------------------------------------------------
 * spline form (PairwiseInteractionABFS.useAnalyticForm = False)
 * ------------------------------------------------------------------------------------------------- */
/* CubicSpline_MakeFromReal1DArrays with both boundary conditions "first derivative = 0" (pC/csource/CubicSpline.c:309-420): second
 * derivatives h from the tridiagonal system, solved as LAPACK dgtsv does when no row interchange is needed (the system is diagonally
 * dominant: |d_i| = (dl + du)/3 against off-diagonals dl/6, du/6) -- forward elimination, back substitution. */
static void spline_second_derivatives(int n, const double *x, const double *y, double *h)
{
    double *dl = (double *) malloc(sizeof(double) * (size_t) n), *d = (double *) malloc(sizeof(double) * (size_t) n), *du = (double *) malloc(sizeof(double) * (size_t) n);
    double l, u;
    int i;
    l = x[1] - x[0];
    d[0] = l / 3.0e+00; du[0] = l / 6.0e+00; h[0] = (y[1] - y[0]) / l - 0.0e+00;
    for (i = 1; i < n - 1; i++) {
        l = x[i] - x[i - 1]; u = x[i + 1] - x[i];
        dl[i - 1] = l / 6.0e+00; d[i] = (l + u) / 3.0e+00; du[i] = u / 6.0e+00;
        h[i] = (y[i + 1] - y[i]) / u + (y[i - 1] - y[i]) / l;
    }
    u = x[n - 1] - x[n - 2];
    dl[n - 2] = u / 6.0e+00; d[n - 1] = u / 3.0e+00; h[n - 1] = 0.0e+00 - (y[n - 1] - y[n - 2]) / u;
    for (i = 0; i < n - 1; i++) {                       /* dgtsv, branch |d_i| >= |dl_i| */
        double fact = dl[i] / d[i];
        d[i + 1] = d[i + 1] - fact * du[i];
        h[i + 1] = h[i + 1] - fact * h[i];
    }
    h[n - 1] = h[n - 1] / d[n - 1];
    if (n > 1) h[n - 2] = (h[n - 2] - du[n - 2] * h[n - 1]) / d[n - 2];
    for (i = n - 3; i >= 0; i--) h[i] = (h[i] - du[i] * h[i + 1] - 0.0e+00 * h[i + 2]) / d[i];
    free(dl); free(d); free(du);
}

/* PairwiseInteractionABFS_Make{Electrostatic,LennardJonesA,LennardJonesB}Spline (pM/csource/PairwiseInteraction.c:148-285):
 * which = 0 electrostatic in kJ/mol (3: atomic units), 1 LJ-A, 2 LJ-B; abscissae x_i = (i dR)^2, last point (r2Off, 0).
 * Returns the number of points NumberOfSplinePoints gives (pM/cinclude/PairwiseInteraction.h:166-167); x == NULL: only that. */
int orc_make_spline(int which, double damp, double inner, double outer, int density, double *x, double *y, double *hh)
{
    double F[21], dR, a = (double) density * outer;
    int n = ((a >= 0) ? (int) (a + 0.5) : (int) (a - 0.5)) + 1, i;
    if (n < 2) n = 2;
    if (x == NULL) return n;
    orc_make_factors(damp, inner, outer, F);
    dR = outer / (double) (n - 1);
    for (i = 0; i < n - 1; i++) {
        double r = dR * (double) i, r2 = r * r, s, s2, s6, f;
        x[i] = r2;
        if (r2 < F[R2DAMP]) { s2 = 0.0e+00; s = 0.0e+00; } else { s2 = 1.0e+00 / r2; s = sqrt(s2); }
        s6 = s2 * s2 * s2;
        if (which == 0 || which == 3) {
            if (r2 > F[R2ON])        f = 1.0e+00 * (s * (F[FA] - r2 * (F[FB] + r2 * (F[FC] + F[FD] * r2))) + F[QSHIFT2]);
            else if (r2 > F[R2DAMP]) f = 1.0e+00 * (s + F[QSHIFT1]);
            else                     f = 1.0e+00 * (F[QF0] - F[QALPHA] * r2);
        } else if (which == 1) {
            if (r2 > F[R2ON])        { double l1 = s6 - F[AF6]; f = 1.0e+00 * F[AK12] * pow(l1, 2); }
            else if (r2 > F[R2DAMP]) f = 1.0e+00 * (s6 * s6 - F[ASHIFT12]);
            else                     f = 1.0e+00 * (F[AF0] - F[AALPHA] * r2);
        } else {
            if (r2 > F[R2ON])        { double l2 = (s / r2) - F[BF3]; f = -1.0e+00 * F[BK6] * pow(l2, 2); }
            else if (r2 > F[R2DAMP]) f = -1.0e+00 * (s6 - F[BSHIFT6]);
            else                     f = -1.0e+00 * (F[BF0] - F[BALPHA] * r2);
        }
        y[i] = f;
    }
    x[n - 1] = F[R2OFF]; y[n - 1] = 0.0e+00;
    if (which == 0) for (i = 0; i < n; i++) y[i] *= ORC_E2A_TO_KJMOL;                     /* Real1DArray_Scale = cblas_dscal */
    if (which == 3) { double sc = 1.0e+00 / ORC_ANGSTROMS_TO_BOHRS; for (i = 0; i < n; i++) y[i] *= sc; }
    spline_second_derivatives(n, x, y, hh);
    return n;
}

/* CubicSpline_EvaluateLUDST + CubicSpline_FastEvaluateFG (pC/csource/CubicSpline.c:138-159, pC/cinclude/CubicSpline.h:30-39) */
static inline void spline_ludst(int n, const double *x, double x0, int *l, int *u, double *d, double *s, double *t)
{
    int l0 = 0, u0 = n - 1, i;
    while ((u0 - l0) > 1) { i = (u0 + l0) >> 1; if (x[i] > x0) u0 = i; else l0 = i; }
    *l = l0; *u = u0;
    *d = (x[u0] - x[l0]);
    *s = (x0 - x[l0]) / (*d);
    *t = (x[u0] - x0) / (*d);
}
static inline void spline_fg(const double *y, const double *h, int l, int u, double d, double s, double t, double *f, double *g)
{
    double hl = h[l] * d / 6.0e+00, hu = h[u] * d / 6.0e+00, yl = y[l], yu = y[u];
    *f = t * yl + s * yu + d * (t * (t * t - 1.0e+00) * hl + s * (s * s - 1.0e+00) * hu);
    *g = (yu - yl) / d + (-(3.0e+00 * t * t - 1.0e+00) * hl + (3.0e+00 * s * s - 1.0e+00) * hu);
}
void orc_spline_evaluate(int n, const double *x, const double *y, const double *h, double x0, double *f, double *g)
{
    int l, u; double d, s, t;
    spline_ludst(n, x, x0, &l, &u, &d, &s, &t);
    spline_fg(y, h, l, u, d, s, t, f, g);
}

static void make_splines(OrcNB *h)
{
    int k, n = orc_make_spline(0, h->damp, h->inner, h->outer, h->density, NULL, NULL, NULL);
    free(h->spx); for (k = 0; k < 3; k++) { free(h->spy[k]); free(h->sph[k]); }
    h->nsp = n;
    h->spx = (double *) malloc(sizeof(double) * (size_t) n);
    for (k = 0; k < 3; k++) {
        h->spy[k] = (double *) malloc(sizeof(double) * (size_t) n); h->sph[k] = (double *) malloc(sizeof(double) * (size_t) n);
        orc_make_spline(k, h->damp, h->inner, h->outer, h->density, h->spx, h->spy[k], h->sph[k]);
    }
}

/* PairwiseInteractionABFS.SetOptions(useAnalyticForm, splinePointDensity) + MakeSplines; call after orc_set_options */
void orc_set_interaction_form(OrcNB *h, int useAnalyticForm, int splinePointDensity)
{
    h->useAnalytic = useAnalyticForm; h->density = splinePointDensity;
    make_splines(h);
}

/* PairwiseInteractionABFS_MMMMEnergy, spline branch: pM/csource/PairwiseInteraction.c:431-531.  Note the charge scale: the
 * splines carry the kJ/mol unit, so qi = electrostaticScale * q_i (:474). */
static void mmmm_energy_spline(const OrcNB *h, const int *tindex, const double *tA, const double *tB, int nt,
                               long npairs, const int *pairs, double electrostaticScale, double ljScale,
                               const double *crd1, const double *crd2, double *eElect, double *eLJ, double *grd1, double *grd2)
{
    double eQQ = 0.0e+00, eL = 0.0e+00, r2Off = h->outer * h->outer;
    long p;
    for (p = 0; p < npairs; p++) {
        int i = pairs[2 * p], j = pairs[2 * p + 1], tij, l, u;
        double qi = electrostaticScale * h->q[i], qij, Aij, Bij, d, s, t, f, dF, dG;
        double xij = crd1[3 * i] - crd2[3 * j], yij = crd1[3 * i + 1] - crd2[3 * j + 1], zij = crd1[3 * i + 2] - crd2[3 * j + 2];
        double r2 = (xij * xij + yij * yij + zij * zij);
        if (r2 > r2Off) continue;
        spline_ludst(h->nsp, h->spx, r2, &l, &u, &d, &s, &t);
        dG = 0.0e+00;
        qij = qi * h->q[j];
        spline_fg(h->spy[0], h->sph[0], l, u, d, s, t, &f, &dF);
        eQQ += (qij * f);
        dG  += (2.0e+00 * qij * dF);
        tij = tindex[nt * h->ljtype[i] + h->ljtype[j]];
        Aij = tA[tij] * ljScale;
        Bij = tB[tij] * ljScale;
        spline_fg(h->spy[1], h->sph[1], l, u, d, s, t, &f, &dF);
        eL += Aij * f;
        dG += (2.0e+00 * Aij * dF);
        spline_fg(h->spy[2], h->sph[2], l, u, d, s, t, &f, &dF);
        eL += Bij * f;
        dG += (2.0e+00 * Bij * dF);
        if (grd1 != NULL) {
            xij *= dG; yij *= dG; zij *= dG;
            grd1[3 * i] += xij; grd1[3 * i + 1] += yij; grd1[3 * i + 2] += zij;
            grd2[3 * j] -= xij; grd2[3 * j + 1] -= yij; grd2[3 * j + 2] -= zij;
        }
    }
    *eElect = eQQ; *eLJ = eL;
}

/* PairwiseInteractionABFS_MMMMEnergy, analytic branch: pM/csource/PairwiseInteraction.c:292-429 (loop :374-427).
 * pairs[(i,j)]: i indexes crd1/grd1, j indexes crd2/grd2. */
static void mmmm_energy(const OrcNB *h, const double *F, const int *tindex, const double *tA, const double *tB, int nt,
                        long npairs, const int *pairs, double electrostaticScale, double ljScale,
                        const double *crd1, const double *crd2, double *eElect, double *eLJ, double *grd1, double *grd2)
{
    double eScale = electrostaticScale * ORC_E2A_TO_KJMOL, eQQ = 0.0e+00, eL = 0.0e+00;
    long p;
    if (!h->useAnalytic) { mmmm_energy_spline(h, tindex, tA, tB, nt, npairs, pairs, electrostaticScale, ljScale, crd1, crd2, eElect, eLJ, grd1, grd2); return; }
    for (p = 0; p < npairs; p++) {
        int i = pairs[2 * p], j = pairs[2 * p + 1], tij;
        double qi = eScale * h->q[i], qij, Aij, Bij, eq, el, dF;
        double xij = crd1[3 * i] - crd2[3 * j], yij = crd1[3 * i + 1] - crd2[3 * j + 1], zij = crd1[3 * i + 2] - crd2[3 * j + 2];
        double r2 = (xij * xij + yij * yij + zij * zij);
        qij = qi * h->q[j];
        tij = tindex[nt * h->ljtype[i] + h->ljtype[j]];
        Aij = tA[tij] * ljScale;
        Bij = tB[tij] * ljScale;
        if (!pair_terms(F, r2, qij, Aij, Bij, &eq, &el, &dF)) continue;
        eQQ += eq;
        eL  += el;
        if (grd1 != NULL) {
            xij *= (2.0e+00 * dF); yij *= (2.0e+00 * dF); zij *= (2.0e+00 * dF);
            grd1[3 * i] += xij; grd1[3 * i + 1] += yij; grd1[3 * i + 2] += zij;
            grd2[3 * j] -= xij; grd2[3 * j + 1] -= yij; grd2[3 * j + 2] -= zij;
        }
    }
    *eElect = eQQ; *eLJ = eL;
}

/* ---------------------------------------------------------------------------------------------------
 * construction
 * ------------------------------------------------------------------------------------------------- */
static void *dupmem(const void *p, size_t bytes) { void *q = malloc(bytes ? bytes : 1); if (p != NULL && bytes) memcpy(q, p, bytes); return q; }

/* Matrix33_IsEqual / IsIdentity tolerance 1e-6: pC/csource/Matrix33.c:224-284 ; Vector3_IsNull pC/csource/Vector3.c:72-89 */
static int m33_is_equal(const double *a, const double *b) { int i; for (i = 0; i < 9; i++) if (fabs(a[i] - b[i]) > 1.0e-6) return 0; return 1; }

OrcNB *orc_create(int n, const double *charges, const int *ljtypes,
                  int ntypes, const int *tableindex, const double *tableA, const double *tableB,
                  int ntypes14, const int *tableindex14, const double *tableA14, const double *tableB14,
                  int nexcl, const int *exclPairs, int n14, const int *pairs14,
                  int ntrans, const double *rot, const double *trans)
{
    OrcNB *h = (OrcNB *) calloc(1, sizeof(OrcNB));
    int i, j, t, k;
    static const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    h->n = n; h->ntypes = ntypes; h->ntrans = ntrans;
    h->q = (double *) dupmem(charges, sizeof(double) * n);
    h->ljtype = (int *) dupmem(ljtypes, sizeof(int) * n);
    h->tindex = (int *) dupmem(tableindex, sizeof(int) * ntypes * ntypes);
    h->tA = (double *) dupmem(tableA, sizeof(double) * (ntypes * (ntypes + 1)) / 2);
    h->tB = (double *) dupmem(tableB, sizeof(double) * (ntypes * (ntypes + 1)) / 2);
    if (tableindex14 == NULL) { ntypes14 = ntypes; tableindex14 = tableindex; tableA14 = tableA; tableB14 = tableB; }
    h->ntypes14 = ntypes14;
    h->tindex14 = (int *) dupmem(tableindex14, sizeof(int) * ntypes14 * ntypes14);
    h->tA14 = (double *) dupmem(tableA14, sizeof(double) * (ntypes14 * (ntypes14 + 1)) / 2);
    h->tB14 = (double *) dupmem(tableB14, sizeof(double) * (ntypes14 * (ntypes14 + 1)) / 2);
    /* symmetric exclusion CSR: SelfPairList_MakeConnections pC/csource/PairList.c:458-526 */
    h->exclPtr = (int *) calloc((size_t) n + 1, sizeof(int));
    for (k = 0; k < nexcl; k++) { i = exclPairs[2 * k]; j = exclPairs[2 * k + 1]; if (i != j) { h->exclPtr[i + 1]++; h->exclPtr[j + 1]++; } }
    for (i = 0; i < n; i++) h->exclPtr[i + 1] += h->exclPtr[i];
    h->exclCol = (int *) malloc(sizeof(int) * (size_t) (h->exclPtr[n] ? h->exclPtr[n] : 1));
    {
        int *cur = (int *) dupmem(h->exclPtr, sizeof(int) * ((size_t) n + 1));
        for (k = 0; k < nexcl; k++) { i = exclPairs[2 * k]; j = exclPairs[2 * k + 1]; if (i != j) { h->exclCol[cur[i]++] = j; h->exclCol[cur[j]++] = i; } }
        free(cur);
    }
    /* 1-4 list: GenerateLists14 -> SelfPairList_FromSelfPairList with all atoms MM and free = the input list
     * (pM/csource/NBModelABFS.c:1113-1156, pC/csource/PairList.c:342). */
    h->n14 = h->n14all = n14;
    h->p14 = (int *) dupmem(pairs14, sizeof(int) * 2 * (size_t) n14);
    h->p14all = (int *) dupmem(pairs14, sizeof(int) * 2 * (size_t) n14);
    h->rot = (double *) dupmem(rot, sizeof(double) * 9 * (size_t) ntrans);
    h->trans = (double *) dupmem(trans, sizeof(double) * 3 * (size_t) ntrans);
    /* Transformation3Container_FindIdentity / _FindInverses pC/csource/Transformation3Container.c:58-72,120-154 */
    h->identity = -1;
    for (t = 0; t < ntrans; t++) {
        const double *tv = h->trans + 3 * t;
        if (m33_is_equal(h->rot + 9 * t, eye) && fabs(tv[0]) <= 1.0e-6 && fabs(tv[1]) <= 1.0e-6 && fabs(tv[2]) <= 1.0e-6) { h->identity = t; break; }
    }
    h->inverses = (int *) malloc(sizeof(int) * (size_t) (ntrans ? ntrans : 1));
    for (t = 0; t < ntrans; t++) h->inverses[t] = -1;
    for (i = 0; i < ntrans; i++) {
        if (h->inverses[i] < 0) {
            double m[9];
            m33_invert(m, h->rot + 9 * i);
            for (j = 0; j <= i; j++) {
                if (h->inverses[j] < 0 && m33_is_equal(m, h->rot + 9 * j)) { h->inverses[i] = j; h->inverses[j] = i; break; }
            }
        }
    }
    h->xref = (double *) calloc(3 * (size_t) n, sizeof(double));
    h->isNew = 1;
    h->stListCutoff = h->stOuterCutoff = 0.0;
    h->useAnalytic = 1; h->density = 50;                        /* pM/csource/PairwiseInteraction.c:34-35 */
    orc_set_options(h, 0.5, 8.0, 12.0, 13.5, 1.0, 1.0, 1, 0);   /* pM/csource/NBModelABFS.c:28-37 */
    return h;
}

static void free_images(OrcNB *h)
{
    int k;
    for (k = 0; k < h->nimages; k++) { free(h->images[k].pairs); free(h->images[k].xyz); }
    free(h->images); h->images = NULL; h->nimages = 0;
}

void orc_destroy(OrcNB *h)
{
    if (h == NULL) return;
    free_images(h);
    free(h->primary); free(h->xref); free(h->inverses); free(h->rot); free(h->trans); free(h->p14); free(h->p14all); free(h->fixed); free(h->isoPtr); free(h->isoIdx); free(h->xc); free(h->isoT);
    free(h->exclPtr); free(h->exclCol); free(h->tindex); free(h->tindex14); free(h->tA); free(h->tB); free(h->tA14); free(h->tB14);
    free(h->spx); { int k; for (k = 0; k < 3; k++) { free(h->spy[k]); free(h->sph[k]); } }
    free(h->q); free(h->ljtype); free(h);
}

/* fixedAtoms of NBModelABFSState_SetUp (pM/csource/NBModelABFSState.c:316-345): lists keep a pair only if one of its atoms is free
 * (orSelection of the generators; GenerateLists14 -> SelfPairList_FromSelfPairList, NBModelABFS.c:1113-1128, pC/csource/PairList.c:342),
 * and CheckForUpdate ignores fixed atoms.  nfixed = 0 clears. */
void orc_set_fixed(OrcNB *h, int nfixed, const int *fixed)
{
    long k, m = 0;
    int i;
    free(h->fixed); h->fixed = NULL;
    if (nfixed > 0) {
        h->fixed = (char *) calloc((size_t) h->n, 1);
        for (i = 0; i < nfixed; i++) if (fixed[i] >= 0 && fixed[i] < h->n) h->fixed[fixed[i]] = 1;
    }
    for (k = 0; k < h->n14all; k++) {
        const int a = h->p14all[2 * k], b = h->p14all[2 * k + 1];
        if (h->fixed != NULL && h->fixed[a] && h->fixed[b]) continue;
        h->p14[2 * m] = a; h->p14[2 * m + 1] = b; m++;
    }
    h->n14 = m;
    h->isNew = 1;
}

static int cmp_int(const void *a, const void *b) { return (*(const int *) a > *(const int *) b) - (*(const int *) a < *(const int *) b); }

/* NBModelABFSState_SetUpCentering (pM/csource/NBModelABFSState.c:425-450): isolates from the exclusions
 * (SelfPairList_ToIsolateSelectionContainer, pC/csource/PairList.c:686-770: breadth-first components, indices sorted),
 * isolates with a fixed atom removed; centring needs exclusions, transformations and more than one isolate. */
void orc_set_centering(OrcNB *h, int on)
{
    int n = h->n, s, i, c, m = 0, niso = 0, *ptr, *idx;
    char *assigned;
    free(h->isoPtr); free(h->isoIdx); free(h->xc); free(h->isoT);
    h->isoPtr = h->isoIdx = NULL; h->xc = h->isoT = NULL; h->centering = 0; h->nisolates = 0; h->isNew = 1;
    if (!on || h->ntrans <= 0 || h->exclPtr[n] == 0) return;
    ptr = (int *) malloc(sizeof(int) * ((size_t) n + 1)); idx = (int *) malloc(sizeof(int) * (size_t) n);
    assigned = (char *) calloc((size_t) n, 1);
    for (s = 0; s < n; s++) {
        int start = m, keep = 1;
        if (assigned[s]) continue;
        idx[m++] = s; assigned[s] = 1;
        for (i = start; i < m; i++)
            for (c = h->exclPtr[idx[i]]; c < h->exclPtr[idx[i] + 1]; c++) { const int j = h->exclCol[c]; if (!assigned[j]) { idx[m++] = j; assigned[j] = 1; } }
        qsort(idx + start, (size_t) (m - start), sizeof(int), cmp_int);
        if (h->fixed != NULL) for (i = start; i < m; i++) if (h->fixed[idx[i]]) keep = 0;     /* SelectionContainer_RemoveIsolates */
        if (keep) ptr[niso++] = start; else m = start;
    }
    ptr[niso] = m;
    free(assigned);
    if (niso > 1) {
        h->isoPtr = ptr; h->isoIdx = idx; h->nisolates = niso; h->centering = 1;
        h->xc = (double *) calloc(3 * (size_t) n, sizeof(double)); h->isoT = (double *) calloc(3 * (size_t) n, sizeof(double));
    } else { free(ptr); free(idx); }
}

/* NBModelABFSState_InitializeCoordinates3 (pM/csource/NBModelABFSState.c:278-311) with SymmetryParameters_CenterCoordinates3ByIsolate
 * (pM/csource/SymmetryParameters.c:82-129), Coordinates3_Center (pC/csource/Coordinates3.c:357-440), _FindCenteringTranslation (:271-290) */
static const double *centre_coordinates(OrcNB *h, const double *x, const double *M, const double *invM, int doUpdate)
{
    int n = h->n, k, i, d;
    if (!h->centering) return x;
    if (!doUpdate) { for (i = 0; i < 3 * n; i++) h->xc[i] = x[i] + h->isoT[i]; return h->xc; }
    memcpy(h->xc, x, sizeof(double) * 3 * (size_t) n);
    for (k = 0; k < h->nisolates; k++) {
        double c[3] = {0.0, 0.0, 0.0}, f[3], t[3], na, nb, nc;
        const int lo = h->isoPtr[k], hi = h->isoPtr[k + 1];
        const double scale = 1.0e+00 / (double) (hi - lo);
        for (i = lo; i < hi; i++) for (d = 0; d < 3; d++) c[d] += h->xc[3 * h->isoIdx[i] + d];
        for (d = 0; d < 3; d++) c[d] *= scale;
        for (d = 0; d < 3; d++) f[d] = c[0] * invM[3 * d] + c[1] * invM[3 * d + 1] + c[2] * invM[3 * d + 2];
        na = (double) (-(int) floor(f[0])); nb = (double) (-(int) floor(f[1])); nc = (double) (-(int) floor(f[2]));
        for (d = 0; d < 3; d++) t[d] = na * M[3 * d] + nb * M[3 * d + 1] + nc * M[3 * d + 2];
        for (i = lo; i < hi; i++) for (d = 0; d < 3; d++) h->xc[3 * h->isoIdx[i] + d] += t[d];
    }
    for (i = 0; i < 3 * n; i++) h->isoT[i] = h->xc[i] + (-1.0e+00) * x[i];
    return h->xc;
}

void orc_set_options(OrcNB *h, double damp, double inner, double outer, double list,
                     double dielectric, double elecScale14, int checkForInverses, int imageExpandFactor)
{
    h->damp = damp; h->inner = inner; h->outer = outer; h->list = list;
    h->dielectric = dielectric; h->scale14 = elecScale14;
    h->checkForInverses = checkForInverses; h->expandFactor = imageExpandFactor;
    if (!h->useAnalytic) orc_set_interaction_form(h, 0, h->density);      /* the splines depend on the cutoffs */
}

/* ---------------------------------------------------------------------------------------------------
 * pair search: a uniform-grid search of our own.  The SET it returns is the reference's:
 *   { (i, j) : (dx*dx + dy*dy) + dz*dz <= cutoff^2 }, dx = x1_i - x2_j in fp64 without contraction
 * (CheckForGrid*InteractionWithNoRadii, pC/csource/PairListGenerator.c:70-81,131-139; the reference's
 * includeAll shortcut for fully-enclosed cells, RegularGrid.c:557-563, is the same set mathematically).
 * self != 0: unordered pairs i > j, minus exclusions (MakeSelfPairListFromCoordinates3*, :946-1216).
 * self == 0: all ordered (i, j) (MakeCrossPairListFromCoordinates3CellCell, :703-791, exclusions NULL).
 * ------------------------------------------------------------------------------------------------- */
typedef struct { long n, cap; int *p; } PairBuf;
static void pb_push(PairBuf *b, int i, int j)
{
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 1 << 16; b->p = (int *) realloc(b->p, sizeof(int) * 2 * (size_t) b->cap); }
    b->p[2 * b->n] = i; b->p[2 * b->n + 1] = j; b->n++;
}

static void pair_search(const OrcNB *h, int n, const double *x1, const double *x2, double cutoff, int self, PairBuf *out)
{
    double lo[3], hi[3], c2 = cutoff * cutoff, hcell = cutoff;
    int dim[3], d, i, *cellOf, *start, *items;
    char *flag = NULL;
    long ncell;
    for (d = 0; d < 3; d++) { lo[d] = hi[d] = x2[d]; }
    for (i = 0; i < n; i++) for (d = 0; d < 3; d++) { double v = x2[3 * i + d]; if (v < lo[d]) lo[d] = v; if (v > hi[d]) hi[d] = v; }
    for (d = 0; d < 3; d++) { dim[d] = (int) floor((hi[d] - lo[d]) / hcell) + 1; }
    ncell = (long) dim[0] * dim[1] * dim[2];
    cellOf = (int *) malloc(sizeof(int) * (size_t) n);
    start = (int *) calloc((size_t) ncell + 1, sizeof(int));
    items = (int *) malloc(sizeof(int) * (size_t) n);
    for (i = 0; i < n; i++) {
        int c[3];
        for (d = 0; d < 3; d++) { c[d] = (int) floor((x2[3 * i + d] - lo[d]) / hcell); if (c[d] >= dim[d]) c[d] = dim[d] - 1; if (c[d] < 0) c[d] = 0; }
        cellOf[i] = (c[0] * dim[1] + c[1]) * dim[2] + c[2];
        start[cellOf[i] + 1]++;
    }
    for (i = 0; i < ncell; i++) start[i + 1] += start[i];
    { int *cur = (int *) dupmem(start, sizeof(int) * ((size_t) ncell + 1)); for (i = 0; i < n; i++) items[cur[cellOf[i]]++] = i; free(cur); }
    if (self) flag = (char *) calloc((size_t) n, 1);
    for (i = 0; i < n; i++) {
        double xi = x1[3 * i], yi = x1[3 * i + 1], zi = x1[3 * i + 2];
        int c[3], clo[3], chi[3], cx, cy, cz, k;
        for (d = 0; d < 3; d++) {
            double v = x1[3 * i + d];
            c[d] = (int) floor((v - lo[d]) / hcell);
            clo[d] = c[d] - 1; chi[d] = c[d] + 1;
            if (clo[d] < 0) clo[d] = 0;
            if (chi[d] > dim[d] - 1) chi[d] = dim[d] - 1;
        }
        if (self) for (k = h->exclPtr[i]; k < h->exclPtr[i + 1]; k++) flag[h->exclCol[k]] = 1;
        for (cx = clo[0]; cx <= chi[0]; cx++) for (cy = clo[1]; cy <= chi[1]; cy++) for (cz = clo[2]; cz <= chi[2]; cz++) {
            int cc = (cx * dim[1] + cy) * dim[2] + cz;
            for (k = start[cc]; k < start[cc + 1]; k++) {
                int j = items[k];
                double dx, dy, dz;
                if (self && (j >= i || flag[j])) continue;
                if (h->fixed != NULL && h->fixed[i] && h->fixed[j]) continue;   /* orSelection = freeSelection: QORI || QOR[j] (PairListGenerator.c:118-139) */
                dx = xi - x2[3 * j]; dy = yi - x2[3 * j + 1]; dz = zi - x2[3 * j + 2];
                if ((dx * dx + dy * dy + dz * dz) <= c2) pb_push(out, i, j);
            }
        }
        if (self) for (k = h->exclPtr[i]; k < h->exclPtr[i + 1]; k++) flag[h->exclCol[k]] = 0;
    }
    free(flag); free(items); free(start); free(cellOf);
}

/* Coordinates3_EnclosingOrthorhombicBox pC/csource/Coordinates3.c:498-... (origin = min, extents = max - min) */
static void enclosing_box(int n, const double *x, double *origin, double *extents)
{
    double mx[3], mn[3]; int i, d;
    for (d = 0; d < 3; d++) mx[d] = mn[d] = x[d];
    for (i = 1; i < n; i++) for (d = 0; d < 3; d++) { double v = x[3 * i + d]; if (v > mx[d]) mx[d] = v; if (v < mn[d]) mn[d] = v; }
    for (d = 0; d < 3; d++) { origin[d] = mn[d]; extents[d] = mx[d] - mn[d]; }
}

/* GetLimits pM/csource/SymmetryParameters.c:199-219 */
static int get_limits(double bl, double bu, double il, double iu, double t, int *low, int *high)
{
    int n = 0;
    while (iu >= bl) { il -= t; iu -= t; n--; }
    while (iu <  bl) { il += t; iu += t; n++; }
    if (il <= bu) {
        *low = n;
        while (il <= bu) { il += t; iu += t; n++; }
        *high = n - 1;
        return 1;
    }
    return 0;
}

/* SymmetryParameters_FindBoxSearchLimits pM/csource/SymmetryParameters.c:221-265 */
static void find_box_search_limits(const double *M, const double *lower, const double *upper, const double *ilower, const double *iupper, int *lim)
{
    int ok;
    double bl, bu, d1, d2;
    lim[0] = lim[2] = lim[4] = 0; lim[1] = lim[3] = lim[5] = -1;   /* alow ahigh blow bhigh clow chigh */
    ok = get_limits(lower[2], upper[2], ilower[2], iupper[2], M[8], &lim[4], &lim[5]);
    if (ok) {
        d1 = lim[4] * M[5]; d2 = lim[5] * M[5];
        bl = lower[1] - (d1 > d2 ? d1 : d2);
        bu = upper[1] - (d1 < d2 ? d1 : d2);
        ok = get_limits(bl, bu, ilower[1], iupper[1], M[4], &lim[2], &lim[3]);
        if (ok) {
            d1 = lim[4] * M[2]; d2 = lim[5] * M[2];
            bl = lower[0] - (d1 > d2 ? d1 : d2);
            bu = upper[0] - (d1 < d2 ? d1 : d2);
            d1 = lim[2] * M[1]; d2 = lim[3] * M[1];
            bl -= (d1 > d2 ? d1 : d2);
            bu -= (d1 < d2 ? d1 : d2);
            get_limits(bl, bu, ilower[0], iupper[0], M[0], &lim[0], &lim[1]);
        }
    }
}

#define ORC_ROUND(a) (((a) >= 0) ? (int) ((a) + 0.5) : (int) ((a) - 0.5))   /* pC/cinclude/Macros.h:34 */

/* Transformation3Container_FindInverseIntegerTranslation pC/csource/Transformation3Container.c:80-113 */
static void find_inverse_translation(const OrcNB *h, int t, int a, int b, int c, int *inv)
{
    int tinv = h->inverses[t], i;
    double v[3];
    inv[0] = inv[1] = inv[2] = -999999;
    if (tinv < 0) return;
    v[0] = h->trans[3 * t] + (double) a; v[1] = h->trans[3 * t + 1] + (double) b; v[2] = h->trans[3 * t + 2] + (double) c;
    m33_apply(h->rot + 9 * tinv, v);
    for (i = 0; i < 3; i++) { v[i] *= -1.0e+00; v[i] += -1.0e+00 * h->trans[3 * tinv + i]; }
    for (i = 0; i < 3; i++) {
        int r = ORC_ROUND(v[i]);
        inv[i] = (fabs(v[i] - (double) r) < 1.0e-4) ? r : -999999;
    }
}

/* Real-space form of a fractional transformation: Transformation3_Orthogonalize pC/csource/Transformation3.c:95-103 */
static void orthogonalize(const double *rotF, const double *transF, const double *M, const double *invM, double *R, double *tv)
{
    memcpy(R, rotF, sizeof(double) * 9);
    m33_premul(R, M);
    m33_postmul(R, invM);
    tv[0] = transF[0]; tv[1] = transF[1]; tv[2] = transF[2];
    m33_apply(M, tv);
}

/* Coordinates3_Transform = Rotate then Translate: pC/csource/Coordinates3.c:1473-1500,1760-1800 */
static void transform_coords(int n, const double *x, const double *R, const double *tv, double *out)
{
    int i;
    for (i = 0; i < n; i++) {
        double x0 = x[3 * i], y0 = x[3 * i + 1], z0 = x[3 * i + 2];
        double x1 = R[0] * x0 + R[1] * y0 + R[2] * z0;
        double y1 = R[3] * x0 + R[4] * y0 + R[5] * z0;
        double z1 = R[6] * x0 + R[7] * y0 + R[8] * z0;
        x1 += tv[0]; y1 += tv[1]; z1 += tv[2];
        out[3 * i] = x1; out[3 * i + 1] = y1; out[3 * i + 2] = z1;
    }
}

static void translate_coords(int n, double *x, const double *d)
{
    int i;
    for (i = 0; i < n; i++) { x[3 * i] += d[0]; x[3 * i + 1] += d[1]; x[3 * i + 2] += d[2]; }
}

/* GenerateImageLists, MM/MM part: pM/csource/NBModelABFS.c:753-1045 */
static void generate_image_lists(OrcNB *h, const double *x, const double *M, const double *invM)
{
    int n = h->n, t, d;
    double lower[3], upper[3], ilower[3], iupper[3], disp[3];
    double *ix = (double *) malloc(sizeof(double) * 3 * (size_t) n);
    free_images(h);
    enclosing_box(n, x, lower, upper);
    for (d = 0; d < 3; d++) { upper[d] += lower[d]; }
    for (d = 0; d < 3; d++) { lower[d] += -h->list; upper[d] += h->list; }
    for (t = 0; t < h->ntrans; t++) {
        int tinverse = h->inverses[t], lim[6], a, b, c;
        double defaultscale = 0.5e+00, R[9], tv[3];
        if (!h->checkForInverses) tinverse = -1;
        if (tinverse >= 0) {
            if (t < tinverse) continue;
            defaultscale = 1.0e+00;
        }
        orthogonalize(h->rot + 9 * t, h->trans + 3 * t, M, invM, R, tv);
        transform_coords(n, x, R, tv, ix);
        enclosing_box(n, ix, ilower, iupper);
        for (d = 0; d < 3; d++) iupper[d] += ilower[d];
        find_box_search_limits(M, lower, upper, ilower, iupper, lim);
        if (h->expandFactor > 0) { lim[0] -= h->expandFactor; lim[1] += h->expandFactor; lim[2] -= h->expandFactor; lim[3] += h->expandFactor; lim[4] -= h->expandFactor; lim[5] += h->expandFactor; }
        for (a = lim[0]; a <= lim[1]; a++) for (b = lim[2]; b <= lim[3]; b++) for (c = lim[4]; c <= lim[5]; c++) {
            double scale = defaultscale;
            if ((a == 0) && (b == 0) && (c == 0) && (h->identity == t)) continue;
            if ((tinverse >= 0) && (tinverse == t)) {
                int inv[3];
                find_inverse_translation(h, t, a, b, c, inv);
                if ((inv[0] >= lim[0]) && (inv[0] <= lim[1]) && (inv[1] >= lim[2]) && (inv[1] <= lim[3]) && (inv[2] >= lim[4]) && (inv[2] <= lim[5])) {
                    if ((a == inv[0]) && (b == inv[1]) && (c == inv[2])) scale = 0.5e+00;
                    else {
                        scale = 1.0e+00;
                        if ((a < inv[0]) || ((a == inv[0]) && (b < inv[1])) || ((a == inv[0]) && (b == inv[1]) && (c < inv[2]))) continue;
                    }
                } else scale = 1.0e+00;
            }
            /* SymmetryParameters_Displacement pM/csource/SymmetryParameters.c:178-192 */
            for (d = 0; d < 3; d++) disp[d] = ((double) a) * M[3 * d] + ((double) b) * M[3 * d + 1] + ((double) c) * M[3 * d + 2];
            translate_coords(n, ix, disp);
            for (d = 0; d < 3; d++) { ilower[d] += disp[d]; iupper[d] += disp[d]; }
            if ((ilower[0] <= upper[0]) && (ilower[1] <= upper[1]) && (ilower[2] <= upper[2]) &&
                (iupper[0] >= lower[0]) && (iupper[1] >= lower[1]) && (iupper[2] >= lower[2])) {
                PairBuf pb = {0, 0, NULL};
                pair_search(h, n, x, ix, h->list, 0, &pb);
                if (pb.n > 0) {     /* ImageList_CreateImage keeps non-empty lists only: pM/csource/ImageList.c:37-70 */
                    OrcImage *im;
                    h->images = (OrcImage *) realloc(h->images, sizeof(OrcImage) * (size_t) (h->nimages + 1));
                    im = &h->images[h->nimages++];
                    im->t = t; im->a = a; im->b = b; im->c = c; im->scale = scale; im->npairs = pb.n; im->pairs = pb.p;
                    im->xyz = (double *) dupmem(ix, sizeof(double) * 3 * (size_t) n);
                } else free(pb.p);
            }
            for (d = 0; d < 3; d++) disp[d] *= -1.0e+00;
            translate_coords(n, ix, disp);
            for (d = 0; d < 3; d++) { ilower[d] += disp[d]; iupper[d] += disp[d]; }
        }
    }
    free(ix);
}

/* CheckForUpdate pM/csource/NBModelABFS.c:691-746 (no fixed atoms) */
static int check_for_update(int n, const double *x, const double *xref, const char *fixed, double listCutoff, double outerCutoff, double *maxDisp)
{
    int i, doUpdate = 0;
    double buffac = 0.5e+00 * (listCutoff - outerCutoff), buffacsq = buffac * buffac, maxr2 = 0.0e+00;
    for (i = 0; i < n; i++) {
        double dx = x[3 * i] - xref[3 * i], dy = x[3 * i + 1] - xref[3 * i + 1], dz = x[3 * i + 2] - xref[3 * i + 2];
        double r2 = dx * dx + dy * dy + dz * dz;
        if (fixed != NULL && fixed[i]) continue;             /* NBModelABFS.c:723-739 */
        maxr2 = (maxr2 > r2) ? maxr2 : r2;
        if (r2 > buffacsq) { doUpdate = 1; break; }
    }
    *maxDisp = sqrt(maxr2);
    return doUpdate;
}

/* CheckForImageUpdate pM/csource/NBModelABFS.c:635-684 */
static int check_for_image_update(const OrcNB *h, const double *M, double maxDisp)
{
    double buffac = h->list - h->stOuterCutoff - maxDisp, dM[9];
    int k, i;
    if (h->images == NULL || !h->haveRefM) return 0;
    for (i = 0; i < 9; i++) dM[i] = M[i] + (-1.0e+00) * h->refM[i];
    for (k = 0; k < h->nimages; k++) {
        const OrcImage *im = &h->images[k];
        double v[3], di;
        v[0] = h->trans[3 * im->t] + (double) im->a; v[1] = h->trans[3 * im->t + 1] + (double) im->b; v[2] = h->trans[3 * im->t + 2] + (double) im->c;
        m33_apply(dM, v);
        di = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (di > buffac) return 1;
    }
    return 0;
}

/* SymmetryParameterGradients_ImageDerivatives pM/csource/SymmetryParameterGradients.c:158-238 */
static void image_derivatives(double *dEdM, const double *M, const double *invM, const double *rotF, const double *transF,
                              int n, const double *x, const double *g)
{
    double ms[9], si[9], di[9];
    int a, b, i;
    memcpy(ms, rotF, sizeof(ms)); m33_premul(ms, M);
    memcpy(si, rotF, sizeof(si)); m33_postmul(si, invM);
    for (a = 0; a < 3; a++) for (b = 0; b < 3; b++) {
        double sum = 0.0e+00, t = transF[b];
        m33_inverse_derivative(M, a, b, di);
        m33_premul(di, ms);
        di[3 * a] += si[3 * b]; di[3 * a + 1] += si[3 * b + 1]; di[3 * a + 2] += si[3 * b + 2];
        for (i = 0; i < n; i++) {
            double xx = x[3 * i], y = x[3 * i + 1], z = x[3 * i + 2], gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
            double dx = di[0] * xx + di[1] * y + di[2] * z;
            double dy = di[3] * xx + di[4] * y + di[5] * z;
            double dz = di[6] * xx + di[7] * y + di[8] * z;
            switch (a) { case 0: dx += t; break; case 1: dy += t; break; case 2: dz += t; break; }
            sum += dx * gx + dy * gy + dz * gz;
        }
        dEdM[3 * a + b] += sum;
    }
}

/* NBModelABFS_Update (pM/csource/NBModelABFS.c:508-623) + NBModelABFS_MMMMEnergy (:228-301) + MMMMImageEnergy (:1161-1313) */
int orc_energy(OrcNB *h, const double *xin, const double *box, int forceNew,
               double *energies, double *grad, double *dEdM, double *timings)
{
    const double *x = xin;                 /* the coordinates the lists and energies see: the input, or the centred copy */
    int n = h->n, doUpdate, k, i;
    double M[9], invM[9], F[21], maxDisp = 0.0e+00, eScale, t0, t1, t2;
    double *ix = NULL, *ig = NULL;
    int hasSym = (h->ntrans > 0);
    if (hasSym) orc_make_M(box, M, invM);
    if (forceNew) h->isNew = 1;
    t0 = now_s();
    doUpdate = h->isNew;
    doUpdate = doUpdate || (h->list != h->stListCutoff) || (h->outer != h->stOuterCutoff);
    if (doUpdate) { h->stListCutoff = h->list; h->stOuterCutoff = h->outer; }
    doUpdate = doUpdate || check_for_update(n, xin, h->xref, h->fixed, h->list, h->stOuterCutoff, &maxDisp);
    if (hasSym) x = centre_coordinates(h, xin, M, invM, doUpdate);
    if (doUpdate) {
        PairBuf pb = {0, 0, NULL};
        free(h->primary);
        pair_search(h, n, x, x, h->list, 1, &pb);
        h->primary = pb.p; h->nprimary = pb.n;
        memcpy(h->xref, xin, sizeof(double) * 3 * (size_t) n);
    }
    if (hasSym) {
        doUpdate = doUpdate || check_for_image_update(h, M, maxDisp);
        if (doUpdate) {
            generate_image_lists(h, x, M, invM);
            memcpy(h->refM, M, sizeof(M)); h->haveRefM = 1;
        }
    }
    h->isNew = 0;
    t1 = now_s();

    orc_make_factors(h->damp, h->inner, h->outer, F);
    eScale = 1.0e+00 / h->dielectric;
    for (i = 0; i < 6; i++) energies[i] = 0.0;
    /* images */
    if (hasSym && h->nimages > 0) {
        double eQQ = 0.0, eLJ = 0.0;
        ix = (double *) malloc(sizeof(double) * 3 * (size_t) n);
        if (grad != NULL) ig = (double *) malloc(sizeof(double) * 3 * (size_t) n);
        for (k = 0; k < h->nimages; k++) {
            const OrcImage *im = &h->images[k];
            double xt[3], R[9], tv[3], eq, el;
            xt[0] = h->trans[3 * im->t] + (double) im->a; xt[1] = h->trans[3 * im->t + 1] + (double) im->b; xt[2] = h->trans[3 * im->t + 2] + (double) im->c;
            orthogonalize(h->rot + 9 * im->t, xt, M, invM, R, tv);
            transform_coords(n, x, R, tv, ix);
            if (ig != NULL) memset(ig, 0, sizeof(double) * 3 * (size_t) n);
            mmmm_energy(h, F, h->tindex, h->tA, h->tB, h->ntypes, im->npairs, im->pairs, eScale * im->scale, im->scale, x, ix, &eq, &el, grad, ig);
            eQQ += eq; eLJ += el;
            if (ig != NULL) {
                double Rt[9];
                if (dEdM != NULL) image_derivatives(dEdM, M, invM, h->rot + 9 * im->t, xt, n, x, ig);
                Rt[0] = R[0]; Rt[1] = R[3]; Rt[2] = R[6]; Rt[3] = R[1]; Rt[4] = R[4]; Rt[5] = R[7]; Rt[6] = R[2]; Rt[7] = R[5]; Rt[8] = R[8];
                for (i = 0; i < n; i++) {
                    double x0 = ig[3 * i], y0 = ig[3 * i + 1], z0 = ig[3 * i + 2];
                    grad[3 * i]     += 1.0e+00 * (Rt[0] * x0 + Rt[1] * y0 + Rt[2] * z0);
                    grad[3 * i + 1] += 1.0e+00 * (Rt[3] * x0 + Rt[4] * y0 + Rt[5] * z0);
                    grad[3 * i + 2] += 1.0e+00 * (Rt[6] * x0 + Rt[7] * y0 + Rt[8] * z0);
                }
            }
        }
        energies[4] = eQQ; energies[5] = eLJ;
    }
    /* primary */
    if (h->nprimary > 0)
        mmmm_energy(h, F, h->tindex, h->tA, h->tB, h->ntypes, h->nprimary, h->primary, eScale, 1.0e+00, x, x, &energies[0], &energies[1], grad, grad);
    /* 1-4 */
    eScale *= h->scale14;
    if (h->n14 > 0)
        mmmm_energy(h, F, h->tindex14, h->tA14, h->tB14, h->ntypes14, h->n14, h->p14, eScale, 1.0e+00, x, x, &energies[2], &energies[3], grad, grad);
    free(ix); free(ig);
    t2 = now_s();
    if (timings != NULL) { timings[0] = t1 - t0; timings[1] = t2 - t1; }
    return doUpdate;
}

long orc_num_primary_pairs(OrcNB *h) { return h->nprimary; }
int  orc_num_images(OrcNB *h) { return h->nimages; }
long orc_num_image_pairs(OrcNB *h) { long s = 0; int k; for (k = 0; k < h->nimages; k++) s += h->images[k].npairs; return s; }
long orc_num_14_pairs(OrcNB *h) { return h->n14; }
void orc_get_primary_pairs(OrcNB *h, int *pairs) { memcpy(pairs, h->primary, sizeof(int) * 2 * (size_t) h->nprimary); }
void orc_get_image_info(OrcNB *h, int k, int *info, double *scale)
{
    const OrcImage *im = &h->images[k];
    info[0] = im->t; info[1] = im->a; info[2] = im->b; info[3] = im->c; info[4] = (int) im->npairs; info[5] = 0;
    scale[0] = im->scale;
}
void orc_get_image_pairs(OrcNB *h, int k, int *pairs) { memcpy(pairs, h->images[k].pairs, sizeof(int) * 2 * (size_t) h->images[k].npairs); }
void orc_get_image_coordinates(OrcNB *h, int k, double *xyz) { memcpy(xyz, h->images[k].xyz, sizeof(double) * 3 * (size_t) h->n); }
