// rect_row.cuh — one node row of a RECTANGULAR matrix, `matrix B = vb(Uh,Vh)` with two different spaces on one mesh
// (rows = dofs of the test space Vh, columns = dofs of the space of the unknown Uh).
//
// Replaces Element_Op with Ku != Kv (fflib/problem.cpp:6337-6437 in 3-D, :6063-6160 in 2-D: `same` false, the basis
// functions of both elements tabulated at every quadrature point, n = Kv.NbDoF rows, m = Ku.NbDoF columns) followed by
// HashMatrix::operator+= (femlib/HashMatrix.cpp:1295-1332).  Same ownership as every other assembly kernel of this
// library: the node row is owned by ONE thread, which walks the (element, local node) records of its node in element
// order and adds the a-th block row of every element matrix into its CSR segment — the summation order of an entry is
// the reference's (elements in order; inside an element quadrature points outermost, then the terms), no atomics.
//
// This file is plain C++ behind FF_HD so that the very same row routine is compiled into the kernel (assemble.cu,
// k_asm_rect) and, by the test-suite only, into a host harness (tests/host_rect.cpp) that runs it row by row against the
// oracle on machines without a GPU.  Nothing in the library calls it on the host.
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifdef __CUDACC__
#define FF_HD __host__ __device__ __forceinline__
#else
#define FF_HD inline
#endif

static constexpr int RECT_MAXQ = 32;   // quadrature points of the volume rule
static constexpr int RECT_MAXT = 64;   // terms of the form
static constexpr int RECT_MAXLAB = 16; // region labels of the integral

struct RectTerm {
    double coef;
    int vcomp, ucomp, vslot, uslot; // slot: 0 value, 1..3 d/dx, d/dy, d/dz
};
struct RectParams {
    int nq, nterms, nlab; // nlab < 0: every region
    int order_v, order_u, ncv, ncu, nloc_u;
    int labels[RECT_MAXLAB];
    double w[RECT_MAXQ];      // weights of the rule (sum = 1)
    double lam[RECT_MAXQ][4]; // barycentric coordinates of its points
    RectTerm t[RECT_MAXT];
};

// gradients of the barycentric coordinates and signed measure of element K (femlib/Mesh3dn.hpp:126-136, fem.hpp:321-324)
template <int DIM>
FF_HD void rect_geometry(const int32_t *K, const double *xyz, int vstride, double (&G)[DIM + 1][DIM], double &mes)
{
    double X[DIM + 1][DIM];
    for (int a = 0; a <= DIM; ++a)
        for (int d = 0; d < DIM; ++d) X[a][d] = xyz[(size_t)K[a] * vstride + d];
    if (DIM == 2) {
        const double ax = X[1][0] - X[0][0], ay = X[1][1] - X[0][1], bx = X[2][0] - X[0][0], by = X[2][1] - X[0][1];
        const double det = ax * by - ay * bx;
        G[1][0] = by / det;  G[1][1] = -bx / det;
        G[2][0] = -ay / det; G[2][1] = ax / det;
        mes = det * 0.5;
    } else {
        double e[3][3], c[3][3]; // e[r] = X[r+1] - X[0];  c[r] = e[r+1] x e[r+2]
        for (int r = 0; r < 3; ++r)
            for (int d = 0; d < 3; ++d) e[r][d] = X[r + 1][d] - X[0][d];
        for (int r = 0; r < 3; ++r) {
            const double *u = e[(r + 1) % 3], *v = e[(r + 2) % 3];
            c[r][0] = u[1] * v[2] - u[2] * v[1];
            c[r][1] = u[2] * v[0] - u[0] * v[2];
            c[r][2] = u[0] * v[1] - u[1] * v[0];
        }
        const double det = e[0][0] * c[0][0] + e[0][1] * c[0][1] + e[0][2] * c[0][2];
        for (int r = 0; r < 3; ++r)
            for (int d = 0; d < DIM; ++d) G[r + 1][d] = c[r][d] / det;
        mes = det * (1.0 / 6.0);
    }
    for (int d = 0; d < DIM; ++d) {
        double s = 0.0;
        for (int a = 1; a <= DIM; ++a) s += G[a][d];
        G[0][d] = -s;
    }
}

// value and gradient of the basis function of local node a at the point with barycentric coordinates l: out[0] value,
// out[1..DIM] derivatives.  P2 edge nodes: {01,02,03,12,13,23} on a tetrahedron, the edge opposite vertex e on a triangle
// (femlib/P012_3d.cpp:199-300, femlib/FESpace.cpp:1219-1262)
template <int DIM>
FF_HD void rect_basis(int order, int a, const double *l, const double (&G)[DIM + 1][DIM], double (&out)[4])
{
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (order == 1) {
        out[0] = l[a];
        for (int d = 0; d < DIM; ++d) out[1 + d] = G[a][d];
        return;
    }
    if (a <= DIM) {
        out[0] = l[a] * (2.0 * l[a] - 1.0);
        for (int d = 0; d < DIM; ++d) out[1 + d] = (4.0 * l[a] - 1.0) * G[a][d];
        return;
    }
    int p, r;
    if (DIM == 3) {
        const int x = a - 4; // 01 02 03 12 13 23
        p = x < 3 ? 0 : (x < 5 ? 1 : 2);
        r = x < 3 ? x + 1 : (x < 5 ? x - 1 : 3);
    } else {
        p = (a - 3 + 1) % 3;
        r = (a - 3 + 2) % 3;
    }
    out[0] = 4.0 * l[p] * l[r];
    for (int d = 0; d < DIM; ++d) out[1 + d] = 4.0 * (l[p] * G[r][d] + l[r] * G[p][d]);
}

// The row of test node i.  rec(e) = e-th incidence record of the node, (element << 4) | local node, in element order
// (0xffffffff: padding).  ncol[rb .. rb+L): the sorted column nodes of the row; row: its values, component-block layout
// row[(cv * L + p) * ncu + cu] = B(i*ncv + cv, ncol[rb+p]*ncu + cu), zeroed by the caller.
template <int DIM, class RecFn>
FF_HD void rect_row(int ninc, RecFn rec, const int32_t *conn, const int32_t *elab, const double *xyz, int vstride,
                    const int32_t *e2n_u, const RectParams &P, const int32_t *ncol, int rb, int L, double *row)
{
    const int ncu = P.ncu, ncv = P.ncv, nlu = P.nloc_u, nq = P.nq, nterms = P.nterms;
    for (int e = 0; e < ninc; ++e) {
        const uint32_t it = rec(e);
        if (it == 0xffffffffu) continue;
        const int el = (int)(it >> 4), a = (int)(it & 15u);
        if (P.nlab >= 0) {
            const int lab = elab[el];
            bool ok = false;
            for (int x = 0; x < P.nlab; ++x) ok |= (P.labels[x] == lab);
            if (!ok) continue;
        }
        double G[DIM + 1][DIM], mes;
        rect_geometry<DIM>(conn + (size_t)(DIM + 1) * el, xyz, vstride, G, mes);
        const int32_t *Nu = e2n_u + (size_t)nlu * el;
        for (int b = 0; b < nlu; ++b) {
            double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int q = 0; q < nq; ++q) {
                double va[4], ub[4];
                rect_basis<DIM>(P.order_v, a, P.lam[q], G, va);
                rect_basis<DIM>(P.order_u, b, P.lam[q], G, ub);
                const double w = mes * P.w[q];
                for (int t = 0; t < nterms; ++t) {
                    const RectTerm &T = P.t[t];
                    acc[T.vcomp][T.ucomp] += (T.coef * w) * va[T.vslot] * ub[T.uslot];
                }
            }
            const int j = Nu[b];
            int lo = 0, hi = L - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ncol[rb + mid] < j) lo = mid + 1;
                else hi = mid;
            }
            for (int cv = 0; cv < ncv; ++cv)
                for (int cu = 0; cu < ncu; ++cu)
                    if (acc[cv][cu] != 0.0) row[((size_t)cv * L + lo) * ncu + cu] += acc[cv][cu];
        }
    }
}
