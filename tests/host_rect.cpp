// host_rect.cpp — TEST INFRASTRUCTURE ONLY (compiled by tests/test_rect_row_host.py with g++, never part of the library):
// runs the row routine of the rectangular assembly kernel (freefem-sources_b200/csrc/rect_row.cuh, the same source the
// kernel k_asm_rect compiles) row by row on the host, so that its arithmetic can be compared with the oracle where there
// is no GPU.  Incidence lists and the node-level pattern come from the test (numpy).
#include "../freefem-sources_b200/csrc/rect_row.cuh"
#include <cstring>

struct RecCsr {
    const uint32_t *inc;
    int base;
    uint32_t operator()(int e) const { return inc[base + e]; }
};

extern "C" int host_rect_sizeof_params() { return (int)sizeof(RectParams); }

extern "C" int host_rect_assemble(int dim, int nrows, const int32_t *incptr, const uint32_t *inc, const int32_t *conn,
                                  const int32_t *elab, const double *xyz, int vstride, const int32_t *e2n_u, int order_v,
                                  int ncv, int order_u, int ncu, int nloc_u, int nq, const double *lam /* nq x 4 */,
                                  const double *w, int nterms, const double *coef, const int32_t *tdesc /* nterms x 4: vcomp ucomp vslot uslot */,
                                  int nlab, const int32_t *labels, const int32_t *nrowptr, const int32_t *ncol, double *vals)
{
    if (nq > RECT_MAXQ || nterms > RECT_MAXT || nlab > RECT_MAXLAB) return 1;
    RectParams P;
    memset(&P, 0, sizeof(P));
    P.nq = nq; P.nterms = nterms; P.nlab = labels ? nlab : -1;
    P.order_v = order_v; P.order_u = order_u; P.ncv = ncv; P.ncu = ncu; P.nloc_u = nloc_u;
    for (int i = 0; i < nlab && labels; ++i) P.labels[i] = labels[i];
    for (int q = 0; q < nq; ++q) {
        P.w[q] = w[q];
        for (int a = 0; a < 4; ++a) P.lam[q][a] = lam[4 * q + a];
    }
    for (int t = 0; t < nterms; ++t) P.t[t] = RectTerm{coef[t], tdesc[4 * t], tdesc[4 * t + 1], tdesc[4 * t + 2], tdesc[4 * t + 3]};
    for (int i = 0; i < nrows; ++i) {
        const int rb = nrowptr[i], L = nrowptr[i + 1] - rb;
        double *row = vals + (size_t)ncv * ncu * rb;
        RecCsr rec{inc, incptr[i]};
        if (dim == 3) rect_row<3>(incptr[i + 1] - incptr[i], rec, conn, elab, xyz, vstride, e2n_u, P, ncol, rb, L, row);
        else rect_row<2>(incptr[i + 1] - incptr[i], rec, conn, elab, xyz, vstride, e2n_u, P, ncol, rb, L, row);
    }
    return 0;
}
