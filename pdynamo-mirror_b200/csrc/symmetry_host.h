// symmetry_host.h -- host-side (C++) lattice / image bookkeeping of the NBModelABFS path.
//
// B200 product code.  The device kernels need, per list rebuild, a small amount of fp64 host logic that the
// reference keeps in SymmetryParameters.c, Transformation3*.c, Matrix33.c and GenerateImageLists: which
// periodic images to visit, in which order, with which scale, and the exact fp64 operation sequence that
// produces their coordinates (pair-list membership is decided on those coordinates bit for bit).
// Reference (paths under pDynamo 1.9.0; pM = pMolecule-1.9.0/extensions, pC = pCore-1.9.0/extensions):
//   SymmetryParameters_MakeM                 pM/csource/SymmetryParameters.c:370-409
//   SymmetryParameters_FindBoxSearchLimits   pM/csource/SymmetryParameters.c:199-265
//   Transformation3_Orthogonalize            pC/csource/Transformation3.c:95-103
//   Transformation3Container_Find*           pC/csource/Transformation3Container.c:58-154
//   GenerateImageLists                       pM/csource/NBModelABFS.c:753-1045
//   CheckForImageUpdate                      pM/csource/NBModelABFS.c:635-684
//   SymmetryParameterGradients_ImageDerivatives pM/csource/SymmetryParameterGradients.c:158-238
#pragma once
#include <vector>

namespace nbb200 {

struct Mat3 {
    double v[9];                      // row-major
    double  operator()(int r, int c) const { return v[3 * r + c]; }
    double &operator()(int r, int c) { return v[3 * r + c]; }
};

double det3(const Mat3 &m);
Mat3   inverse3(const Mat3 &m);                    // cofactors / determinant, reference operation order
Mat3   mul3(const Mat3 &a, const Mat3 &b);         // a . b with the reference's left-to-right sums
void   apply3(const Mat3 &m, double *v);           // v <- m v
Mat3   inverse_derivative3(const Mat3 &m, int i, int j);

struct Lattice {
    Mat3 M, invM;
    void set_crystal(const double *box6);          // a, b, c, alpha, beta, gamma (degrees)
};

// fractional symmetry operations (Transformation3Container)
struct Transformations {
    int n = 0, identity = -1;
    std::vector<Mat3> rot;
    std::vector<double> trans;                     // 3 per item
    std::vector<int> inverses;
    void set(int ntrans, const double *rot9, const double *trans3);
    void inverse_integer_translation(int t, int a, int b, int c, int *inv) const;
};

// real-space form of one operation: x' = R x + tv
struct RealSpaceOp {
    Mat3 R;
    double tv[3];
    bool pureTranslation;                          // fractional rotation is exactly the identity matrix
};
RealSpaceOp orthogonalize(const Mat3 &rotF, const double *transF, const Lattice &lat);

// One (a,b,c) visit of GenerateImageLists' inner loop for transformation t: coordinates are displaced by
// +disp, possibly used for a cross list (image >= 0), then displaced by -disp (drift included).
struct ImageVisit {
    int t, a, b, c;
    double disp[3];
    int image;                                     // candidate image index or -1 (box prefilter failed)
};
struct CandidateImage {
    int t, a, b, c;
    double scale;
    double lo[3], hi[3];                           // bounding box of the image coordinates at list time
};
struct ImagePlan {
    std::vector<RealSpaceOp> baseOps;              // per transformation: Orthogonalize(items[t])
    std::vector<int> activeT;                      // transformations that are visited at all
    std::vector<ImageVisit> visits;                // grouped by t, loop order
    std::vector<CandidateImage> images;
    double lower[3], upper[3];                     // search box: primary bounding box dilated by the cutoff
};

// bboxMin/bboxExt: [0] primary coordinates, [1+t] coordinates transformed by baseOps[t] (min and max-min, as
// Coordinates3_EnclosingOrthorhombicBox returns them).
bool plan_images(const Transformations &tr, const Lattice &lat, double cutoff, bool checkForInverses, int expandFactor,
                 const double *bboxMin, const double *bboxExt, ImagePlan &plan);

bool check_for_image_update(const Transformations &tr, const Lattice &now, const Lattice &ref,
                            const std::vector<CandidateImage> &images, const std::vector<long> &imagePairs,
                            double listCutoff, double outerCutoff, double maximumDisplacement);

// dE/dM += sum_cd D_ab[c][d] W[c][d] + t_b G_a   with W = sum_i g'_i (x) x_i, G = sum_i g'_i
void image_derivatives(double *dEdM, const Lattice &lat, const Mat3 &rotF, const double *transFplusABC,
                       const double *W9, const double *G3);

void make_abfs_factors(double damp, double inner, double outer, double *out21);

// spline form (PairwiseInteractionABFS.useAnalyticForm = False): the reference's three tables on shared abscissae x = r^2
struct SplineTables {
    std::vector<double> x, y[3], h[3];          // [0] electrostatic (kJ/mol for unit charges), [1] LJ-A, [2] LJ-B; h = second derivatives
    int points() const { return (int) x.size(); }
};
int  abfs_spline_points(double outer, int density);
void make_abfs_splines(double damp, double inner, double outer, int density, SplineTables &t);
void spline_second_derivatives(const std::vector<double> &x, const std::vector<double> &y, std::vector<double> &h);
void make_abfs_electrostatic_spline_au(double damp, double inner, double outer, int density, std::vector<double> &x, std::vector<double> &y, std::vector<double> &h);
void spline_interval_polynomial(const std::vector<double> &x, const std::vector<double> &y, const std::vector<double> &h, int l, double *c4);
constexpr double kE2AngstromToKJMol = (1.0e+7 * 6.0221415e+23 * 1.60217653e-19 * 1.60217653e-19) / (4.0e+00 * 3.14159265358979323846 * 8.854187817e-12);

}  // namespace nbb200
