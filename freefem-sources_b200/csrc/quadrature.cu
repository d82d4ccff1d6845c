// quadrature.cu — the simplex quadrature rules FreeFEM selects for a volume integral when the script gives no
// explicit rule: CDomainOfIntegration::FIT / FIV (fflib/problem.cpp:14102-14145) ask QF_Simplex for the rule with
// the fewest points that is exact for degree  qforder-1  (femlib/QuadratureFormular.cpp:73-115; default qforder 6).
// The FreeFEM plugin does not need this (it forwards the points of the GQuadratureFormular FreeFEM picked); it is
// here so that a standalone caller of the C ABI (bench.py, tests) integrates with exactly the same rules.
//
// Rules are stored as symmetric orbits of barycentric points and expanded on request; weights sum to 1 (FreeFEM's
// convention: the element measure is applied by the caller, QuadratureFormular.cpp:725-743 multiplies by 6).
#include "common.cuh"

namespace {
struct Orbit {
    int kind;      // 0: centroid, 1: (a,b,..,b) vertex-type orbit (d+1 points), 2: tetrahedron edge-type (a,a,b,b) (6 points)
    double a, b, w;
};

// expands to reference coordinates xhat_r = lambda_r, r = 1..dim (lambda_0 = 1 - sum)
int expand(int dim, const Orbit *orb, int norb, double *pts, double *w)
{
    int n = 0;
    auto put = [&](const double *lam, double wt) {
        if (pts)
            for (int r = 0; r < dim; ++r) pts[(size_t)n * dim + r] = lam[r + 1];
        if (w) w[n] = wt;
        ++n;
    };
    for (int o = 0; o < norb; ++o) {
        const Orbit &O = orb[o];
        double lam[4];
        if (O.kind == 0) {
            for (int i = 0; i <= dim; ++i) lam[i] = 1.0 / (dim + 1);
            put(lam, O.w);
        } else if (O.kind == 1) {
            for (int p = 0; p <= dim; ++p) {
                for (int i = 0; i <= dim; ++i) lam[i] = (i == p) ? O.a : O.b;
                put(lam, O.w);
            }
        } else {
            for (int p = 0; p < 4; ++p)
                for (int q = p + 1; q < 4; ++q) {
                    for (int i = 0; i < 4; ++i) lam[i] = (i == p || i == q) ? O.a : O.b;
                    put(lam, O.w);
                }
        }
    }
    return n;
}
} // namespace

extern "C" int ffcuda_quadrature(int dim, int qforder, int *nq, double *qpts, double *qw)
{
    FF_API_BEGIN
    FF_REQUIRE(dim == 2 || dim == 3, "ffcuda_quadrature: dim must be 2 or 3");
    FF_REQUIRE(nq, "ffcuda_quadrature: null output");
    int exact = qforder - 1;
    if (exact < 0) exact = 0;
    if (dim == 2) {
        FF_REQUIRE(exact <= 5, "ffcuda_quadrature: triangle rules beyond qforder 6 are not tabulated here; pass the rule explicitly");
        if (exact <= 1) {
            const Orbit R[] = {{0, 0, 0, 1.0}};
            *nq = expand(2, R, 1, qpts, qw);
        } else if (exact == 2) { // edge midpoints
            const Orbit R[] = {{1, 0.0, 0.5, 1.0 / 3.0}};
            *nq = expand(2, R, 1, qpts, qw);
        } else { // 7-point degree-5 rule (Radon / Stroud T2:5-1)
            const double s15 = 3.87298334620741688517926539978;
            const double r = (6 - s15) / 21, s = (9 + 2 * s15) / 21, u = (6 + s15) / 21, v = (9 - 2 * s15) / 21;
            const Orbit R[] = {{0, 0, 0, 0.225}, {1, s, r, (155 - s15) / 1200}, {1, v, u, (155 + s15) / 1200}};
            *nq = expand(2, R, 3, qpts, qw);
        }
    } else {
        // beyond degree 5 FreeFEM itself warns and falls back to the 14-point rule (problem.cpp:14119-14124)
        if (exact <= 1) {
            const Orbit R[] = {{0, 0, 0, 1.0}};
            *nq = expand(3, R, 1, qpts, qw);
        } else if (exact == 2) {
            const Orbit R[] = {{1, 0.58541019662496845446137605030968, 0.138196601125010515179541316563436, 0.25}};
            *nq = expand(3, R, 1, qpts, qw);
        } else { // 14-point degree-5 rule, weights scaled to sum 1
            const Orbit R[] = {
                {1, 0.7217942490673263207930282587889082, 0.0927352503108912264023239137370306, 6 * 0.0122488405193936582572850342477212},
                {1, 0.067342242210098170607962798709629, 0.310885919263300609797345733763457, 6 * 0.0187813209530026417998642753888810},
                {2, 0.454496295874350350508119473720660, 0.045503704125649649491880526279339, 6 * 7.09100346284691107301157135337624e-3}};
            *nq = expand(3, R, 3, qpts, qw);
        }
    }
    FF_API_END(nullptr)
}
