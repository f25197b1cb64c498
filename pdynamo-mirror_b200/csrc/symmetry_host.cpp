// symmetry_host.cpp -- see symmetry_host.h.  Compile with -ffp-contract=off: the values produced here feed
// fp64 pair-list membership tests that must agree bit for bit with the reference's (built without FMA).
#include "symmetry_host.h"
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace nbb200 {

static const double kDegToRad = 3.14159265358979323846 / 180.0e+00;

double det3(const Mat3 &m)
{
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2)) -
           m(0, 1) * (m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1));
}

Mat3 inverse3(const Mat3 &o)
{
    Mat3 s;
    s(0, 0) = o(1, 1) * o(2, 2) - o(1, 2) * o(2, 1);
    s(0, 1) = o(0, 2) * o(2, 1) - o(2, 2) * o(0, 1);
    s(0, 2) = o(0, 1) * o(1, 2) - o(1, 1) * o(0, 2);
    s(1, 0) = o(1, 2) * o(2, 0) - o(1, 0) * o(2, 2);
    s(1, 1) = o(0, 0) * o(2, 2) - o(0, 2) * o(2, 0);
    s(1, 2) = o(0, 2) * o(1, 0) - o(0, 0) * o(1, 2);
    s(2, 0) = o(1, 0) * o(2, 1) - o(1, 1) * o(2, 0);
    s(2, 1) = o(0, 1) * o(2, 0) - o(0, 0) * o(2, 1);
    s(2, 2) = o(0, 0) * o(1, 1) - o(1, 0) * o(0, 1);
    const double f = 1.0e+00 / det3(o);
    for (double &e : s.v) e *= f;
    return s;
}

Mat3 mul3(const Mat3 &a, const Mat3 &b)
{
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
    return r;
}

void apply3(const Mat3 &m, double *v)
{
    const double x = v[0], y = v[1], z = v[2];
    v[0] = x * m(0, 0) + y * m(0, 1) + z * m(0, 2);
    v[1] = x * m(1, 0) + y * m(1, 1) + z * m(1, 2);
    v[2] = x * m(2, 0) + y * m(2, 1) + z * m(2, 2);
}

Mat3 inverse_derivative3(const Mat3 &m, int i, int j)
{
    Mat3 o = inverse3(m);
    const double m00 = m(0, 0), m01 = m(0, 1), m02 = m(0, 2), m10 = m(1, 0), m11 = m(1, 1), m12 = m(1, 2),
                 m20 = m(2, 0), m21 = m(2, 1), m22 = m(2, 2);
    // d(det)/dM_ij and d(adj)/dM_ij, element by element
    static const int minorIdx[9][4] = {{4, 8, 5, 7}, {5, 6, 3, 8}, {3, 7, 4, 6}, {2, 7, 1, 8}, {0, 8, 2, 6},
                                       {1, 6, 0, 7}, {1, 5, 2, 4}, {2, 3, 0, 5}, {0, 4, 1, 3}};
    const int e = 3 * i + j;
    const double ddet = m.v[minorIdx[e][0]] * m.v[minorIdx[e][1]] - m.v[minorIdx[e][2]] * m.v[minorIdx[e][3]];
    for (double &x : o.v) x *= -ddet;
    switch (e) {
        case 0: o(1, 1) += m22; o(1, 2) -= m12; o(2, 1) -= m21; o(2, 2) += m11; break;
        case 1: o(0, 1) -= m22; o(0, 2) += m12; o(2, 1) += m20; o(2, 2) -= m10; break;
        case 2: o(0, 1) += m21; o(0, 2) -= m11; o(1, 1) -= m20; o(1, 2) += m10; break;
        case 3: o(1, 0) -= m22; o(1, 2) += m02; o(2, 0) += m21; o(2, 2) -= m01; break;
        case 4: o(0, 0) += m22; o(0, 2) -= m02; o(2, 0) -= m20; o(2, 2) += m00; break;
        case 5: o(0, 0) -= m21; o(0, 2) += m01; o(1, 0) += m20; o(1, 2) -= m00; break;
        case 6: o(1, 0) += m12; o(1, 1) -= m02; o(2, 0) -= m11; o(2, 1) += m01; break;
        case 7: o(0, 0) -= m12; o(0, 1) += m02; o(2, 0) += m10; o(2, 1) -= m00; break;
        case 8: o(0, 0) += m11; o(0, 1) -= m01; o(1, 0) -= m10; o(1, 1) += m00; break;
    }
    const double f = 1.0e+00 / det3(m);
    for (double &x : o.v) x *= f;
    return o;
}

void Lattice::set_crystal(const double *box)
{
    const double alpha = box[3] * kDegToRad, beta = box[4] * kDegToRad, gamma = box[5] * kDegToRad;
    const double ca = std::cos(alpha), cb = std::cos(beta), cg = std::cos(gamma), sg = std::sin(gamma);
    for (double &e : M.v) e = 0.0;
    M(0, 0) = box[0];
    M(0, 1) = box[1] * cg;
    M(1, 1) = box[1] * sg;
    M(0, 2) = box[2] * cb;
    M(1, 2) = box[2] * (ca - cb * cg) / sg;
    M(2, 2) = box[2] * std::sqrt(1.0e+00 - ca * ca - cb * cb - cg * cg + 2.0e+00 * ca * cb * cg) / sg;
    invM = inverse3(M);
}

static bool near3(const Mat3 &a, const Mat3 &b, double tol)
{
    for (int i = 0; i < 9; i++) if (std::fabs(a.v[i] - b.v[i]) > tol) return false;
    return true;
}

void Transformations::set(int ntrans, const double *rot9, const double *trans3)
{
    n = ntrans; identity = -1;
    rot.resize(n); trans.assign(trans3, trans3 + 3 * (size_t) n); inverses.assign(n, -1);
    for (int t = 0; t < n; t++) std::memcpy(rot[t].v, rot9 + 9 * t, sizeof(double) * 9);
    Mat3 eye{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    for (int t = 0; t < n; t++) {                                   // first identity wins
        const double *tv = &trans[3 * t];
        if (near3(rot[t], eye, 1.0e-6) && std::fabs(tv[0]) <= 1.0e-6 && std::fabs(tv[1]) <= 1.0e-6 && std::fabs(tv[2]) <= 1.0e-6) { identity = t; break; }
    }
    for (int i = 0; i < n; i++) {                                   // rotational inverses only, as the reference
        if (inverses[i] >= 0) continue;
        const Mat3 inv = inverse3(rot[i]);
        for (int j = 0; j <= i; j++)
            if (inverses[j] < 0 && near3(inv, rot[j], 1.0e-6)) { inverses[i] = j; inverses[j] = i; break; }
    }
}

void Transformations::inverse_integer_translation(int t, int a, int b, int c, int *inv) const
{
    const int big = -999999;
    inv[0] = inv[1] = inv[2] = big;
    const int ti = inverses[t];
    if (ti < 0) return;
    double v[3] = {trans[3 * t] + (double) a, trans[3 * t + 1] + (double) b, trans[3 * t + 2] + (double) c};
    apply3(rot[ti], v);
    for (int i = 0; i < 3; i++) {
        v[i] *= -1.0e+00;
        v[i] += -1.0e+00 * trans[3 * ti + i];
        const int r = (v[i] >= 0) ? (int) (v[i] + 0.5) : (int) (v[i] - 0.5);
        inv[i] = (std::fabs(v[i] - (double) r) < 1.0e-4) ? r : big;
    }
}

RealSpaceOp orthogonalize(const Mat3 &rotF, const double *transF, const Lattice &lat)
{
    RealSpaceOp op;
    op.R = mul3(mul3(lat.M, rotF), lat.invM);                     // (M S) M^-1
    op.tv[0] = transF[0]; op.tv[1] = transF[1]; op.tv[2] = transF[2];
    apply3(lat.M, op.tv);
    static const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    op.pureTranslation = (std::memcmp(rotF.v, eye, sizeof(eye)) == 0);
    return op;
}

// while-loops of GetLimits: slide an interval [il, iu] by multiples of t over [bl, bu]
static bool interval_limits(double bl, double bu, double il, double iu, double t, int &low, int &high)
{
    // the reference loops unconditionally; a degenerate lattice or exploded coordinates would spin forever, so bail out
    // (plan_images then reports an error instead of hanging) when the shift count becomes absurd
    const int guard = 100000;
    int n = 0;
    if (!(t > 0.0) || !std::isfinite(bl) || !std::isfinite(bu) || !std::isfinite(il) || !std::isfinite(iu)) { low = 0; high = guard; return true; }
    while (iu >= bl && n > -guard) { il -= t; iu -= t; n--; }
    while (iu <  bl && n <  guard) { il += t; iu += t; n++; }
    if (!(il <= bu)) return false;
    low = n;
    while (il <= bu && n < guard) { il += t; iu += t; n++; }
    high = n - 1;
    return true;
}

static void box_search_limits(const Mat3 &M, const double *lower, const double *upper, const double *ilower, const double *iupper, int *lim)
{
    lim[0] = lim[2] = lim[4] = 0; lim[1] = lim[3] = lim[5] = -1;
    if (!interval_limits(lower[2], upper[2], ilower[2], iupper[2], M(2, 2), lim[4], lim[5])) return;
    double d1 = lim[4] * M(1, 2), d2 = lim[5] * M(1, 2);
    double bl = lower[1] - (d1 > d2 ? d1 : d2), bu = upper[1] - (d1 < d2 ? d1 : d2);
    if (!interval_limits(bl, bu, ilower[1], iupper[1], M(1, 1), lim[2], lim[3])) return;
    d1 = lim[4] * M(0, 2); d2 = lim[5] * M(0, 2);
    bl = lower[0] - (d1 > d2 ? d1 : d2); bu = upper[0] - (d1 < d2 ? d1 : d2);
    d1 = lim[2] * M(0, 1); d2 = lim[3] * M(0, 1);
    bl -= (d1 > d2 ? d1 : d2); bu -= (d1 < d2 ? d1 : d2);
    interval_limits(bl, bu, ilower[0], iupper[0], M(0, 0), lim[0], lim[1]);
}

bool plan_images(const Transformations &tr, const Lattice &lat, double cutoff, bool checkForInverses, int expandFactor,
                 const double *bboxMin, const double *bboxExt, ImagePlan &plan)
{
    const size_t maxVisits = 1u << 20;                // far beyond any physical system (crystals: a few hundred)
    plan.visits.clear(); plan.images.clear(); plan.activeT.clear();
    plan.baseOps.resize(tr.n);
    for (int t = 0; t < tr.n; t++) plan.baseOps[t] = orthogonalize(tr.rot[t], &tr.trans[3 * t], lat);
    for (int d = 0; d < 3; d++) {
        plan.lower[d] = bboxMin[d];
        plan.upper[d] = bboxExt[d] + plan.lower[d];               // extents + origin, as the reference forms it
        plan.lower[d] += -cutoff;
        plan.upper[d] += cutoff;
    }
    for (int t = 0; t < tr.n; t++) {
        int tinverse = checkForInverses ? tr.inverses[t] : -1;
        double defaultScale = 0.5e+00;
        if (tinverse >= 0) {
            if (t < tinverse) continue;                            // its inverse comes later and covers it
            defaultScale = 1.0e+00;
        }
        double ilower[3], iupper[3];
        for (int d = 0; d < 3; d++) { ilower[d] = bboxMin[3 * (1 + t) + d]; iupper[d] = bboxExt[3 * (1 + t) + d] + ilower[d]; }
        int lim[6];
        box_search_limits(lat.M, plan.lower, plan.upper, ilower, iupper, lim);
        if (expandFactor > 0) for (int k = 0; k < 3; k++) { lim[2 * k] -= expandFactor; lim[2 * k + 1] += expandFactor; }
        {
            double count = 1.0;
            for (int k = 0; k < 3; k++) count *= (double) std::max(0, lim[2 * k + 1] - lim[2 * k] + 1);
            if (count + (double) plan.visits.size() > (double) maxVisits) return false;
        }
        bool any = false;
        for (int a = lim[0]; a <= lim[1]; a++) for (int b = lim[2]; b <= lim[3]; b++) for (int c = lim[4]; c <= lim[5]; c++) {
            if (a == 0 && b == 0 && c == 0 && tr.identity == t) continue;
            double scale = defaultScale;
            if (tinverse >= 0 && tinverse == t) {
                int inv[3];
                tr.inverse_integer_translation(t, a, b, c, inv);
                const bool inRange = inv[0] >= lim[0] && inv[0] <= lim[1] && inv[1] >= lim[2] && inv[1] <= lim[3] && inv[2] >= lim[4] && inv[2] <= lim[5];
                if (inRange) {
                    if (a == inv[0] && b == inv[1] && c == inv[2]) scale = 0.5e+00;     // pure self-inverse image
                    else {
                        scale = 1.0e+00;
                        if (a < inv[0] || (a == inv[0] && b < inv[1]) || (a == inv[0] && b == inv[1] && c < inv[2])) continue;
                    }
                } else scale = 1.0e+00;
            }
            ImageVisit v;
            v.t = t; v.a = a; v.b = b; v.c = c; v.image = -1;
            for (int d = 0; d < 3; d++) v.disp[d] = ((double) a) * lat.M(d, 0) + ((double) b) * lat.M(d, 1) + ((double) c) * lat.M(d, 2);
            for (int d = 0; d < 3; d++) { ilower[d] += v.disp[d]; iupper[d] += v.disp[d]; }
            if (ilower[0] <= plan.upper[0] && ilower[1] <= plan.upper[1] && ilower[2] <= plan.upper[2] &&
                iupper[0] >= plan.lower[0] && iupper[1] >= plan.lower[1] && iupper[2] >= plan.lower[2]) {
                CandidateImage im;
                im.t = t; im.a = a; im.b = b; im.c = c; im.scale = scale;
                for (int d = 0; d < 3; d++) { im.lo[d] = ilower[d]; im.hi[d] = iupper[d]; }
                v.image = (int) plan.images.size();
                plan.images.push_back(im);
            }
            plan.visits.push_back(v);
            any = true;
            for (int d = 0; d < 3; d++) { const double nd = v.disp[d] * -1.0e+00; ilower[d] += nd; iupper[d] += nd; }
        }
        if (any) plan.activeT.push_back(t);
    }
    return true;
}

bool check_for_image_update(const Transformations &tr, const Lattice &now, const Lattice &ref,
                            const std::vector<CandidateImage> &images, const std::vector<long> &imagePairs,
                            double listCutoff, double outerCutoff, double maximumDisplacement)
{
    const double buffac = listCutoff - outerCutoff - maximumDisplacement;
    Mat3 dM;
    for (int i = 0; i < 9; i++) dM.v[i] = now.M.v[i] + (-1.0e+00) * ref.M.v[i];
    for (size_t k = 0; k < images.size(); k++) {
        if (imagePairs[k] <= 0) continue;                          // the reference list holds non-empty images only
        const CandidateImage &im = images[k];
        double v[3] = {tr.trans[3 * im.t] + (double) im.a, tr.trans[3 * im.t + 1] + (double) im.b, tr.trans[3 * im.t + 2] + (double) im.c};
        apply3(dM, v);
        if (std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) > buffac) return true;
    }
    return false;
}

void image_derivatives(double *dEdM, const Lattice &lat, const Mat3 &rotF, const double *tF, const double *W, const double *G)
{
    const Mat3 ms = mul3(lat.M, rotF), si = mul3(rotF, lat.invM);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
        Mat3 di = mul3(ms, inverse_derivative3(lat.M, a, b));
        for (int d = 0; d < 3; d++) di(a, d) += si(b, d);
        double sum = 0.0;
        for (int c = 0; c < 3; c++) for (int d = 0; d < 3; d++) sum += di(c, d) * W[3 * c + d];   // sum_i g'_ic (D x_i)_c
        sum += tF[b] * G[a];
        dEdM[3 * a + b] += sum;
    }
}

void make_abfs_factors(double damp, double inner, double outer, double *o)
{
    // order: r2Damp r2On r2Off | a b c d qShift1 qShift2 qF0 qAlpha | aF6 aK12 aShift12 aF0 aAlpha | bF3 bK6 bShift6 bF0 bAlpha
    const double r2Damp = damp * damp, r2Off = outer * outer, r2On = inner * inner;
    const double gamma = std::pow(r2Off - r2On, 3);
    o[0] = r2Damp; o[1] = r2On; o[2] = r2Off;
    o[3] = r2Off * r2Off * (r2Off - 3.0e+00 * r2On) / gamma;
    o[4] = 6.0e+00 * r2Off * r2On / gamma;
    o[5] = -(r2Off + r2On) / gamma;
    o[6] = 0.4e+00 / gamma;
    o[7] = 8.0e+00 * (r2Off * r2On * (outer - inner) - 0.2e+00 * (outer * r2Off * r2Off - inner * r2On * r2On)) / gamma;
    o[8] = -(o[3] / outer) + o[4] * outer + o[5] * outer * r2Off + o[6] * outer * r2Off * r2Off;
    double f = 1.0e+00 / damp + o[7], g = -1.0e+00 / r2Damp;
    o[9]  = f - 0.5e+00 * damp * g;
    o[10] = -0.5e+00 * g / damp;
    o[11] = 1.0e+00 / (r2Off * r2Off * r2Off);
    o[12] = std::pow(r2Off, 3) / (std::pow(r2Off, 3) - std::pow(r2On, 3));
    o[13] = 1.0e+00 / std::pow(inner * outer, 6);
    const double s12 = 1.0e+00 / std::pow(damp, 12);
    f = s12 - o[13]; g = -12.0e+00 * s12 / damp;
    o[14] = f - 0.5e+00 * damp * g;
    o[15] = -0.5e+00 * g / damp;
    o[16] = 1.0e+00 / (outer * r2Off);
    o[17] = (outer * r2Off) / (outer * r2Off - inner * r2On);
    o[18] = 1.0e+00 / std::pow(inner * outer, 3);
    const double s6 = 1.0e+00 / std::pow(damp, 6);
    f = -s6 + o[18]; g = 6.0e+00 * s6 / damp;
    o[19] = f - 0.5e+00 * damp * g;
    o[20] = -0.5e+00 * g / damp;
}

// ------------------------------------------------------------------------------------------------------
// spline form of the interaction (PairwiseInteractionABFS.useAnalyticForm = False)
// ------------------------------------------------------------------------------------------------------
// One ordinate of PairwiseInteractionABFS_Make{Electrostatic,LennardJonesA,LennardJonesB}Spline (pM/csource/PairwiseInteraction.c:148-285):
// the reference macros (pM/cinclude/PairwiseInteraction.h:72-165) for a unit charge product / unit A / unit B, in its operation order.
static double abfs_ordinate(const double *F, int which, double r2)
{
    const double r2Damp = F[0], r2On = F[1];
    double s = 0.0e+00, s2 = 0.0e+00;
    if (!(r2 < r2Damp)) { s2 = 1.0e+00 / r2; s = std::sqrt(s2); }
    const double s6 = s2 * s2 * s2;
    if (which == 0) {
        if (r2 > r2On) return 1.0e+00 * (s * (F[3] - r2 * (F[4] + r2 * (F[5] + F[6] * r2))) + F[8]);
        if (r2 > r2Damp) return 1.0e+00 * (s + F[7]);
        return 1.0e+00 * (F[9] - F[10] * r2);
    }
    if (which == 1) {
        if (r2 > r2On) { const double l1 = s6 - F[11]; return 1.0e+00 * F[12] * std::pow(l1, 2); }
        if (r2 > r2Damp) return 1.0e+00 * (s6 * s6 - F[13]);
        return 1.0e+00 * (F[14] - F[15] * r2);
    }
    if (r2 > r2On) { const double l2 = (s / r2) - F[16]; return -1.0e+00 * F[17] * std::pow(l2, 2); }
    if (r2 > r2Damp) return -1.0e+00 * (s6 - F[18]);
    return -1.0e+00 * (F[19] - F[20] * r2);
}

// second derivatives of the cubic spline through (x, y) with zero first derivative at both ends
// (CubicSpline_MakeFromReal1DArrays with lower/upper derivative = 1, value 0: pC/csource/CubicSpline.c:309-420).  The reference hands
// the tridiagonal system to LAPACK dgtsv; the system is diagonally dominant, so dgtsv never interchanges rows and reduces to the plain
// elimination below (same operations, same order: the tables agree with the reference's to the last bit).
void spline_second_derivatives(const std::vector<double> &x, const std::vector<double> &y, std::vector<double> &h)
{
    const int n = (int) x.size();
    std::vector<double> sub(n, 0.0), dia(n, 0.0), sup(n, 0.0);
    h.assign(n, 0.0);
    if (n < 2) return;
    for (int i = 0; i < n; i++) {
        const double dl = (i > 0) ? x[i] - x[i - 1] : 0.0, du = (i < n - 1) ? x[i + 1] - x[i] : 0.0;
        if (i == 0)          { dia[0] = du / 3.0e+00; sup[0] = du / 6.0e+00; h[0] = (y[1] - y[0]) / du - 0.0e+00; }
        else if (i == n - 1) { sub[n - 2] = dl / 6.0e+00; dia[n - 1] = dl / 3.0e+00; h[n - 1] = 0.0e+00 - (y[n - 1] - y[n - 2]) / dl; }
        else                 { sub[i - 1] = dl / 6.0e+00; dia[i] = (dl + du) / 3.0e+00; sup[i] = du / 6.0e+00; h[i] = (y[i + 1] - y[i]) / du + (y[i - 1] - y[i]) / dl; }
    }
    for (int i = 0; i + 1 < n; i++) {
        const double m = sub[i] / dia[i];
        dia[i + 1] = dia[i + 1] - m * sup[i];
        h[i + 1] = h[i + 1] - m * h[i];
    }
    h[n - 1] = h[n - 1] / dia[n - 1];
    h[n - 2] = (h[n - 2] - sup[n - 2] * h[n - 1]) / dia[n - 2];
    for (int i = n - 3; i >= 0; i--) h[i] = (h[i] - sup[i] * h[i + 1] - 0.0e+00 * h[i + 2]) / dia[i];
}

int abfs_spline_points(double outer, int density)
{
    const double a = (double) density * outer;                         // Round(), pC/cinclude/Macros.h:34
    const int n = ((a >= 0) ? (int) (a + 0.5) : (int) (a - 0.5)) + 1;
    return n > 2 ? n : 2;
}

void make_abfs_splines(double damp, double inner, double outer, int density, SplineTables &t)
{
    double F[21];
    make_abfs_factors(damp, inner, outer, F);
    const int n = abfs_spline_points(outer, density);
    const double dR = outer / (double) (n - 1);
    t.x.assign(n, 0.0);
    for (int k = 0; k < 3; k++) t.y[k].assign(n, 0.0);
    for (int i = 0; i < n - 1; i++) {
        const double r = dR * (double) i, r2 = r * r;
        t.x[i] = r2;
        for (int k = 0; k < 3; k++) t.y[k][i] = abfs_ordinate(F, k, r2);
    }
    t.x[n - 1] = F[2];                                                 // (r2Off, 0) closes every table
    for (int i = 0; i < n; i++) t.y[0][i] *= kE2AngstromToKJMol;       // the electrostatic spline carries the kJ/mol unit
    for (int k = 0; k < 3; k++) spline_second_derivatives(t.x, t.y[k], t.h[k]);
}

// PairwiseInteractionABFS_MakeElectrostaticSpline with useAtomicUnits (pM/csource/PairwiseInteraction.c:149-200, scale :187): the spline the
// QC/MM and QC/QC interactions carry (NBModelABFS.CheckPairwiseInteractions, pMolecule.NBModelABFS.pyx:83-97) -- potentials in atomic units
// over r^2 in Angstrom^2
void make_abfs_electrostatic_spline_au(double damp, double inner, double outer, int density, std::vector<double> &x, std::vector<double> &y, std::vector<double> &h)
{
    double F[21];
    make_abfs_factors(damp, inner, outer, F);
    const int n = abfs_spline_points(outer, density);
    const double dR = outer / (double) (n - 1);
    const double scale = 1.0e+00 / (1.0e-10 / 5.291772083e-11);       // 1 / UNITS_LENGTH_ANGSTROMS_TO_BOHRS (pC/cinclude/Units.h:21,55-57)
    x.assign(n, 0.0); y.assign(n, 0.0);
    for (int i = 0; i < n - 1; i++) {
        const double r = dR * (double) i, r2 = r * r;
        x[i] = r2;
        y[i] = abfs_ordinate(F, 0, r2);
    }
    x[n - 1] = F[2];
    for (int i = 0; i < n; i++) y[i] *= scale;
    spline_second_derivatives(x, y, h);
}

// per interval [x_l, x_l+1] the cubic in u = x - x_l:  f = c0 + u (c1 + u (c2 + u c3)); algebraically the reference's
// CubicSpline_FastEvaluateFG (pC/cinclude/CubicSpline.h:30-39) with s = u / d, t = 1 - s
void spline_interval_polynomial(const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &h, int l, double *c4)
{
    const double d = x[l + 1] - x[l];
    c4[0] = y[l];
    c4[1] = (y[l + 1] - y[l]) / d - d * (2.0 * h[l] + h[l + 1]) / 6.0;
    c4[2] = 0.5 * h[l];
    c4[3] = (h[l + 1] - h[l]) / (6.0 * d);
}

}  // namespace nbb200
