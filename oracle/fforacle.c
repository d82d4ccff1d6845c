/* fforacle.c — CPU ORACLE (test infrastructure only; see fforacle.h for the contract and the
 * reference file:line each routine restates).  Plain C99, single thread, like the reference path. */
#include "fforacle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ quadrature ---------- */
/* Published rules (Stroud 1971 p.314; Keast / Grundmann-Moeller 14-point degree 5), as tabulated in
 * femlib/QuadratureFormular.cpp:127-188 (triangle) and :690-743 (tetrahedron; weights sum to 1). */
static int q_set(int dim, int n, const double *P, const double *W, double *pts, double *w)
{
    for (int i = 0; i < n; ++i) {
        w[i] = W[i];
        for (int d = 0; d < dim; ++d) pts[i * dim + d] = P[i * dim + d];
    }
    return n;
}

int ffo_quadrature(int dim, const char *name, double *pts, double *w)
{
    if (dim == 2) {
        if (!strcmp(name, "qf1pT")) {
            const double P[] = {1. / 3., 1. / 3.}, W[] = {1.};
            return q_set(2, 1, P, W, pts, w);
        }
        if (!strcmp(name, "qf1pTlump")) {
            const double P[] = {0., 0., 1., 0., 0., 1.}, W[] = {1. / 3., 1. / 3., 1. / 3.};
            return q_set(2, 3, P, W, pts, w);
        }
        if (!strcmp(name, "qf2pT")) {
            const double P[] = {0.5, 0.5, 0.0, 0.5, 0.5, 0.0}, W[] = {1. / 3., 1. / 3., 1. / 3.};
            return q_set(2, 3, P, W, pts, w);
        }
        if (!strcmp(name, "qf5pT")) {
            const double sqrt15 = 3.87298334620741688517926539978;
            const double t = 1.E0 / 3.E0, A = 0.225E0;
            const double r = (6 - sqrt15) / 21, s = (9 + 2 * sqrt15) / 21, B = (155 - sqrt15) / 1200;
            const double u = (6 + sqrt15) / 21, v = (9 - 2 * sqrt15) / 21, C = (155 + sqrt15) / 1200;
            const double P[] = {t, t, r, r, r, s, s, r, u, u, u, v, v, u};
            const double W[] = {A, B, B, B, C, C, C};
            return q_set(2, 7, P, W, pts, w);
        }
        return -1;
    }
    if (dim == 3) {
        if (!strcmp(name, "qfV1")) {
            const double P[] = {0.25, 0.25, 0.25}, W[] = {1.};
            return q_set(3, 1, P, W, pts, w);
        }
        if (!strcmp(name, "qfV1lump")) {
            const double P[] = {0, 0, 0, 1., 0, 0, 0, 1., 0, 0, 0, 1.}, W[] = {0.25, 0.25, 0.25, 0.25};
            return q_set(3, 4, P, W, pts, w);
        }
        if (!strcmp(name, "qfV2")) {
            const double a = 0.58541019662496845446137605030968, b = 0.138196601125010515179541316563436;
            const double P[] = {a, b, b, b, a, b, b, b, a, b, b, b}, W[] = {0.25, 0.25, 0.25, 0.25};
            return q_set(3, 4, P, W, pts, w);
        }
        if (!strcmp(name, "qfV5")) {
            const double a1 = 0.7217942490673263207930282587889082, b1 = 0.0927352503108912264023239137370306;
            const double a2 = 0.067342242210098170607962798709629, b2 = 0.310885919263300609797345733763457;
            const double a3 = 0.454496295874350350508119473720660, b3 = 0.045503704125649649491880526279339;
            const double w1 = 0.0122488405193936582572850342477212 * 6.;
            const double w2 = 0.0187813209530026417998642753888810 * 6.;
            const double w3 = 7.09100346284691107301157135337624E-3 * 6.;
            const double P[] = {a1, b1, b1, b1, a1, b1, b1, b1, a1, b1, b1, b1,
                                a2, b2, b2, b2, a2, b2, b2, b2, a2, b2, b2, b2,
                                a3, a3, b3, a3, b3, a3, b3, a3, a3, b3, b3, a3, b3, a3, b3, a3, b3, b3};
            const double W[] = {w1, w1, w1, w1, w2, w2, w2, w2, w3, w3, w3, w3, w3, w3};
            return q_set(3, 14, P, W, pts, w);
        }
        return -1;
    }
    return -1;
}

/* ------------------------------------------------------------------ meshes -------------- */
static const int nvfaceTet[4][3] = {{3, 2, 1}, {0, 2, 3}, {3, 1, 0}, {0, 1, 2}};
static const int nvedgeTet[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static const int nvedgeTri[3][2] = {{1, 2}, {2, 0}, {0, 1}};

/* the kind=6 split of BuildCube (the hex_decoupe row whose 6 boundary diagonals all have code 0):
 * all six tets share the diagonal 0-7 of the cell; local corner id = a + 2b + 4c. */
static const int cubeTets[6][4] = {{4, 0, 6, 7}, {0, 4, 5, 7}, {1, 0, 5, 7}, {0, 1, 3, 7}, {2, 0, 3, 7}, {0, 2, 6, 7}};

void ffo_cube_sizes(int nx, int ny, int nz, int *nv, int *nt, int *nbe)
{
    *nv = (nx + 1) * (ny + 1) * (nz + 1);
    *nt = 6 * nx * ny * nz;
    *nbe = 4 * (nx * ny + nx * nz + ny * nz);
}

void ffo_cube(int nx, int ny, int nz, double *xyz, int32_t *conn, int32_t *elab,
              int32_t *bconn, int32_t *blab, int32_t *belem, int32_t *bface)
{
    const int nj = nx + 1, nk = nj * (ny + 1);
    const double xd = 1. / nx, yd = 1. / ny, zd = 1. / nz;
    const int nff[6] = {3, 1, 0, 2, 4, 5}; /* plane bit -> label index; labels are 1..6 */
    int nv = (nx + 1) * (ny + 1) * (nz + 1);
    int *vlab = (int *)malloc(sizeof(int) * (size_t)nv);
    int p = 0;
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i, ++p) {
                xyz[3 * p + 0] = 0 + xd * i;
                xyz[3 * p + 1] = 0 + yd * j;
                xyz[3 * p + 2] = 0 + zd * k;
                vlab[p] = 1 * (i == 0) + 2 * (i == nx) + 4 * (j == 0) + 8 * (j == ny) + 16 * (k == 0) + 32 * (k == nz);
            }
    int t = 0, kf = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                int n[8];
                for (int c = 0; c < 2; ++c)
                    for (int b = 0; b < 2; ++b)
                        for (int a = 0; a < 2; ++a) n[a + 2 * b + 4 * c] = (i + a) + nj * (j + b) + nk * (k + c);
                for (int d = 0; d < 6; ++d, ++t) {
                    int nu[4];
                    for (int q = 0; q < 4; ++q) nu[q] = conn[4 * t + q] = n[cubeTets[d][q]];
                    elab[t] = 0;
                    for (int f = 0; f < 4; ++f) {
                        int nf[3] = {nu[nvfaceTet[f][0]], nu[nvfaceTet[f][1]], nu[nvfaceTet[f][2]]};
                        int l = vlab[nf[0]] & vlab[nf[1]] & vlab[nf[2]];
                        for (int kk = 0; kk < 6; ++kk)
                            if (l == (1 << kk)) {
                                if (bconn) {
                                    bconn[3 * kf + 0] = nf[0];
                                    bconn[3 * kf + 1] = nf[1];
                                    bconn[3 * kf + 2] = nf[2];
                                    blab[kf] = nff[kk] + 1;
                                    belem[kf] = t;
                                    bface[kf] = f;
                                }
                                kf++;
                            }
                    }
                }
            }
    free(vlab);
}

void ffo_square_sizes(int nx, int ny, int *nv, int *nt, int *nbe)
{
    *nv = (nx + 1) * (ny + 1);
    *nt = 2 * nx * ny;
    *nbe = 2 * (nx + ny);
}

void ffo_square(int nx, int ny, double *xy, int32_t *conn, int32_t *elab,
                int32_t *bconn, int32_t *blab, int32_t *belem, int32_t *bface)
{
    const int nx1 = nx + 1, ny1 = ny + 1;
    int p = 0;
    for (int j = 0; j < ny1; ++j)
        for (int i = 0; i < nx1; ++i, ++p) {
            xy[2 * p + 0] = (double)i / nx;
            xy[2 * p + 1] = (double)j / ny;
        }
    int t = 0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) { /* flags=0: diagonal i0-i2, direct orientation */
            int i0 = i + j * nx1, i1 = i0 + 1, i2 = i1 + nx1, i3 = i2 - 1;
            conn[3 * t + 0] = i0; conn[3 * t + 1] = i1; conn[3 * t + 2] = i2; elab[t++] = 0;
            conn[3 * t + 0] = i0; conn[3 * t + 1] = i2; conn[3 * t + 2] = i3; elab[t++] = 0;
        }
    if (!bconn) return;
    int e = 0;
    for (int i = 0; i < nx; ++i, ++e) { /* bottom, label 1 */
        bconn[2 * e] = i; bconn[2 * e + 1] = i + 1; blab[e] = 1;
        belem[e] = 2 * i; bface[e] = 2;
    }
    for (int j = 0; j < ny; ++j, ++e) { /* right, label 2 */
        int i1 = nx + j * nx1;
        bconn[2 * e] = i1; bconn[2 * e + 1] = i1 + nx1; blab[e] = 2;
        belem[e] = 2 * ((nx - 1) + j * nx); bface[e] = 0;
    }
    for (int i = 0; i < nx; ++i, ++e) { /* top, label 3 */
        int i1 = i + ny * nx1;
        bconn[2 * e] = i1; bconn[2 * e + 1] = i1 + 1; blab[e] = 3;
        belem[e] = 2 * (i + (ny - 1) * nx) + 1; bface[e] = 0;
    }
    for (int j = 0; j < ny; ++j, ++e) { /* left, label 4 */
        int i1 = j * nx1;
        bconn[2 * e] = i1; bconn[2 * e + 1] = i1 + nx1; blab[e] = 4;
        belem[e] = 2 * (j * nx) + 1; bface[e] = 1;
    }
}

/* ------------------------------------------------------------------ buildlayers --------- */
/* fflib/msh3.cpp:978-1668.  A 2-D vertex i with ni[i] layers becomes the column of 3-D vertices first[i] .. first[i]+ni[i];
 * level s of Nmax (taken from the top, jNmax = Nmax-1 .. 0) maps to the column position (s*ni)/Nmax (integer division),
 * so a column with fewer layers repeats positions and the prism over a triangle degenerates to a pyramid or a
 * tetrahedron.  Quadrilateral faces are cut by the diagonal that holds the largest 3-D vertex number. */
static int bl_lab(int lab, int n, const int32_t *pairs)
{
    int out = lab;
    for (int k = 0; k < n; ++k)
        if (pairs[2 * k] == lab) out = pairs[2 * k + 1];
    return out;
}
static int bl_max(int a, int b) { return a > b ? a : b; }

/* prism over one triangle at one level: P[0..2] lower, P[3..5] upper 3-D vertex numbers; returns the tets (0..3) */
static int bl_prism(const int P[6], int out[3][4])
{
    /* pentahedron cuts of dpent1 (fflib/msh3.cpp:1694-1701), 1-based in the source: data */
    static const int mu[6][12] = {{1, 6, 2, 3, 1, 5, 2, 6, 1, 6, 4, 5}, {1, 6, 2, 3, 1, 4, 2, 6, 2, 6, 4, 5},
                                  {1, 4, 2, 3, 2, 6, 3, 4, 2, 6, 4, 5}, {1, 5, 2, 3, 1, 5, 3, 6, 1, 6, 4, 5},
                                  {1, 5, 2, 3, 1, 5, 3, 4, 3, 6, 4, 5}, {1, 4, 2, 3, 2, 5, 3, 4, 3, 6, 4, 5}};
    static const int pdd[8] = {1, 0, 2, 3, 4, 5, 0, 6};
    int cas = 0;
    for (int j = 0; j < 3; ++j)
        if (P[j] != P[j + 3]) cas += 1 << j;
    if (cas == 0) return 0;
    if (cas == 1 || cas == 2 || cas == 4) {
        const int top = cas == 1 ? 3 : (cas == 2 ? 4 : 5);
        out[0][0] = P[0]; out[0][1] = P[1]; out[0][2] = P[2]; out[0][3] = P[top];
        return 1;
    }
    if (cas != 7) { /* pyramid: the two columns a < b that rise, c merged */
        const int a = cas == 6 ? 1 : 0, b = cas == 3 ? 1 : 2;
        const int first = bl_max(P[a], P[b + 3]) > bl_max(P[b], P[a + 3]);
        out[0][0] = P[0]; out[0][1] = P[1]; out[0][2] = P[2]; out[0][3] = first ? P[b + 3] : P[a + 3];
        out[1][0] = P[5]; out[1][1] = P[4]; out[1][2] = P[3]; out[1][3] = first ? P[a] : P[b];
        return 2;
    }
    const int i1 = bl_max(P[0], P[5]) > bl_max(P[2], P[3]) ? 1 : 2;
    const int i2 = bl_max(P[0], P[4]) > bl_max(P[1], P[3]) ? 1 : 2;
    const int i3 = bl_max(P[1], P[5]) > bl_max(P[2], P[4]) ? 1 : 2;
    const int cut = pdd[(i1 - 1) + 2 * (i2 - 1) + 4 * (i3 - 1)];
    if (cut == 0) return -1; /* cyclic choice: cannot happen with "largest number" diagonals */
    for (int t = 0; t < 3; ++t)
        for (int q = 0; q < 4; ++q) out[t][q] = P[mu[cut - 1][4 * t + q] - 1];
    return 3;
}

/* side quadrilateral over the 2-D edge (i1,i2) at one level: a,b lower (over i1, i2), d,c upper; returns the triangles (0..2) */
static int bl_side(int a, int b, int c, int d, int out[2][3])
{
    const int type = (a != d ? 1 : 0) + (b != c ? 2 : 0);
    if (type == 0) return 0;
    if (type == 1) { out[0][0] = a; out[0][1] = b; out[0][2] = d; return 1; }
    if (type == 2) { out[0][0] = a; out[0][1] = b; out[0][2] = c; return 1; }
    if (bl_max(a, c) > bl_max(b, d)) {
        out[0][0] = a; out[0][1] = b; out[0][2] = c;
        out[1][0] = c; out[1][1] = d; out[1][2] = a;
    } else {
        out[0][0] = a; out[0][1] = b; out[0][2] = d;
        out[1][0] = c; out[1][1] = d; out[1][2] = b;
    }
    return 2;
}

static void bl_edge_vertices(const int32_t *tri, int el, int f, int *i1, int *i2)
{ /* Mesh::VerticesNumberOfEdge femlib/fem.hpp:561-564 */
    *i1 = tri[3 * el + (f + 1) % 3];
    *i2 = tri[3 * el + (f + 2) % 3];
}

void ffo_buildlayers_sizes(int nv2, int nt2, const int32_t *tri, int nbe2, const int32_t *bedge_elem, const int32_t *bedge_face,
                           int nlayer, const int32_t *ni, int *nv, int *nt, int *nbe)
{
    int64_t v = 0, t = 0, b = 2 * (int64_t)nt2;
    for (int i = 0; i < nv2; ++i) v += ni[i] + 1;
    /* the counts of :936-976 are upper bounds when columns degenerate; the exact ones are what the fill produces */
    for (int k = 0; k < nt2; ++k)
        for (int s = nlayer - 1; s >= 0; --s) {
            int P[6], o[3][4];
            for (int j = 0; j < 3; ++j) {
                const int N = ni[tri[3 * k + j]];
                P[j] = (s * N) / nlayer;
                P[j + 3] = ((s + 1) * N) / nlayer;
                /* distinct columns: make numbers of different columns differ */
                P[j] += 1000003 * j; P[j + 3] += 1000003 * j;
            }
            int c = bl_prism(P, o);
            t += c > 0 ? c : 0;
        }
    for (int e = 0; e < nbe2; ++e) {
        int i1, i2, o[2][3];
        bl_edge_vertices(tri, bedge_elem[e], bedge_face[e], &i1, &i2);
        for (int s = nlayer - 1; s >= 0; --s)
            b += bl_side((s * ni[i1]) / nlayer, 1000003 + (s * ni[i2]) / nlayer, 1000003 + ((s + 1) * ni[i2]) / nlayer,
                         ((s + 1) * ni[i1]) / nlayer, o);
    }
    *nv = (int)v; *nt = (int)t; *nbe = (int)b;
}

void ffo_buildlayers(int nv2, const double *xy, int nt2, const int32_t *tri, const int32_t *trilab, int nbe2,
                     const int32_t *bedge_lab, const int32_t *bedge_elem, const int32_t *bedge_face, int nlayer,
                     const int32_t *ni, const double *zmin, const double *zmax, int nreg, const int32_t *regmap, int nmid,
                     const int32_t *midmap, int nup, const int32_t *upmap, int ndown, const int32_t *downmap, double *xyz,
                     int32_t *conn, int32_t *elab, int32_t *bconn, int32_t *blab, int32_t *belem, int32_t *bface)
{
    int *first = (int *)malloc(sizeof(int) * ((size_t)nv2 + 1));
    int nv = 0;
    for (int i = 0; i < nv2; ++i) { /* :1016-1050 */
        const int N = ni[i];
        const double dz = N == 0 ? 0. : (zmax[i] - zmin[i]) / N;
        first[i] = nv;
        for (int j = 0; j <= N; ++j, ++nv) {
            xyz[3 * nv + 0] = xy[2 * i];
            xyz[3 * nv + 1] = xy[2 * i + 1];
            xyz[3 * nv + 2] = zmin[i] + dz * j;
        }
    }
    first[nv2] = nv;
    int nb = 0;
    for (int k = 0; k < nt2; ++k, ++nb) { /* faces at zmax :1111-1128 */
        for (int j = 0; j < 3; ++j) bconn[3 * nb + j] = first[tri[3 * k + j] + 1] - 1;
        blab[nb] = bl_lab(trilab[k], nup, upmap);
    }
    for (int k = 0; k < nt2; ++k, ++nb) { /* faces at zmin, orientation reversed :1132-1150 */
        for (int j = 0; j < 3; ++j) bconn[3 * nb + 2 - j] = first[tri[3 * k + j]];
        blab[nb] = bl_lab(trilab[k], ndown, downmap);
    }
    for (int e = 0; e < nbe2; ++e) { /* lateral faces :1154-1316 */
        int i1, i2, o[2][3];
        bl_edge_vertices(tri, bedge_elem[e], bedge_face[e], &i1, &i2);
        const int lab = bl_lab(bedge_lab[e], nmid, midmap);
        for (int s = nlayer - 1; s >= 0; --s) {
            const int c = bl_side(first[i1] + (s * ni[i1]) / nlayer, first[i2] + (s * ni[i2]) / nlayer,
                                  first[i2] + ((s + 1) * ni[i2]) / nlayer, first[i1] + ((s + 1) * ni[i1]) / nlayer, o);
            for (int t = 0; t < c; ++t, ++nb) {
                for (int j = 0; j < 3; ++j) bconn[3 * nb + j] = o[t][j];
                blab[nb] = lab;
            }
        }
    }
    int nt = 0;
    for (int k = 0; k < nt2; ++k) { /* tetrahedra :1330-1666 */
        const int lab = bl_lab(trilab[k], nreg, regmap);
        for (int s = nlayer - 1; s >= 0; --s) {
            int P[6], o[3][4];
            for (int j = 0; j < 3; ++j) {
                const int v = tri[3 * k + j];
                P[j] = first[v] + (s * ni[v]) / nlayer;
                P[j + 3] = first[v] + ((s + 1) * ni[v]) / nlayer;
            }
            const int c = bl_prism(P, o);
            for (int t = 0; t < c; ++t, ++nt) {
                for (int q = 0; q < 4; ++q) conn[4 * nt + q] = o[t][q];
                elab[nt] = lab;
            }
        }
    }
    free(first);
    if (belem) ffo_boundary_links(nt, conn, elab, nb, bconn, belem, bface);
}

/* The boundary part of GenericMesh::BuildAdj (femlib/GenericMesh.hpp:914-1017) for tetrahedral meshes: every boundary
 * triangle gets its (element, face) — BoundaryElementHeadLink — and the orientation the reference leaves it with.
 * sign of a vertex triple = parity of the permutation that sorts it (SortArray<T,3>, femlib/HashTable.hpp:65-80). */
typedef struct { int32_t v[3]; int32_t id; } bl_face;
static int bl_face_cmp(const void *pa, const void *pb)
{
    const bl_face *a = (const bl_face *)pa, *b = (const bl_face *)pb;
    for (int i = 0; i < 3; ++i)
        if (a->v[i] != b->v[i]) return a->v[i] < b->v[i] ? -1 : 1;
    return a->id < b->id ? -1 : (a->id > b->id ? 1 : 0);
}
static int bl_sort3(int32_t *v)
{
    int32_t t;
    int s = 1;
    if (v[0] > v[1]) { s = -s; t = v[0]; v[0] = v[1]; v[1] = t; }
    if (v[1] > v[2]) {
        s = -s; t = v[1]; v[1] = v[2]; v[2] = t;
        if (v[0] > v[1]) { s = -s; t = v[0]; v[0] = v[1]; v[1] = t; }
    }
    return s;
}
static int bl_face_sign(const int32_t *conn, int id)
{
    int32_t v[3];
    for (int j = 0; j < 3; ++j) v[j] = conn[4 * (id / 4) + nvfaceTet[id % 4][j]];
    return bl_sort3(v);
}

void ffo_boundary_links(int nt, const int32_t *conn, const int32_t *elab, int nbe, int32_t *bconn, int32_t *belem, int32_t *bface)
{
    const size_t nf = (size_t)nt * 4;
    bl_face *F = (bl_face *)malloc(sizeof(bl_face) * (nf ? nf : 1));
    for (size_t id = 0; id < nf; ++id) {
        for (int j = 0; j < 3; ++j) F[id].v[j] = conn[4 * (id / 4) + nvfaceTet[id % 4][j]];
        bl_sort3(F[id].v);
        F[id].id = (int32_t)id;
    }
    qsort(F, nf, sizeof(bl_face), bl_face_cmp);
    /* region pairs of the internal boundary faces: (first, second) counts of :958-965 */
    typedef struct { int a, b, first, second; } pair_t;
    pair_t *pairs = (pair_t *)malloc(sizeof(pair_t) * ((size_t)nbe + 1));
    int npairs = 0, uncorrect = 0;
    for (int step = 0; step < 2; ++step) {
        for (int b = 0; b < nbe; ++b) {
            bl_face key;
            for (int j = 0; j < 3; ++j) key.v[j] = bconn[3 * b + j];
            const int sens = bl_sort3(key.v);
            key.id = -1;
            size_t lo = 0, hi = nf;
            while (lo < hi) {
                size_t mid = (lo + hi) / 2;
                if (bl_face_cmp(&F[mid], &key) < 0) lo = mid + 1; else hi = mid;
            }
            belem[b] = bface[b] = -1;
            if (!(lo < nf && F[lo].v[0] == key.v[0] && F[lo].v[1] == key.v[1] && F[lo].v[2] == key.v[2])) continue;
            const int two = lo + 1 < nf && F[lo + 1].v[0] == key.v[0] && F[lo + 1].v[1] == key.v[1] && F[lo + 1].v[2] == key.v[2];
            int nk = F[lo].id;
            if (!two) { /* true boundary face: same orientation as the face of its element :989-1000 */
                if (bl_face_sign(conn, nk) != sens && step == 0) {
                    int32_t t = bconn[3 * b]; bconn[3 * b] = bconn[3 * b + 1]; bconn[3 * b + 1] = t;
                }
            } else { /* internal face: the element on the side where the face runs the other way :934-986 */
                int nkk = F[lo].id;
                nk = F[lo + 1].id; /* the later of the two elements is looked at first */
                if (sens == bl_face_sign(conn, nk)) { int t = nk; nk = nkk; nkk = t; }
                int regk = elab[nk / 4], regkk = elab[nkk / 4];
                if (regk != regkk) {
                    const int lo_r = regk < regkk ? regk : regkk, hi_r = regk < regkk ? regkk : regk;
                    int q = 0;
                    while (q < npairs && !(pairs[q].a == lo_r && pairs[q].b == hi_r)) ++q;
                    if (q == npairs) { pairs[q].a = lo_r; pairs[q].b = hi_r; pairs[q].first = pairs[q].second = 0; ++npairs; }
                    if (step == 0) {
                        if (regk > regkk) pairs[q].second++; else pairs[q].first++;
                    } else { /* the minority turns round :966-984 */
                        const int sr = regk > regkk ? -1 : 1;
                        if ((pairs[q].first < pairs[q].second && sr == 1) || (pairs[q].first > pairs[q].second && sr == -1)) {
                            int32_t t = bconn[3 * b]; bconn[3 * b] = bconn[3 * b + 1]; bconn[3 * b + 1] = t;
                            nk = nkk;
                        }
                    }
                }
            }
            belem[b] = nk / 4;
            bface[b] = nk % 4;
        }
        uncorrect = 0;
        for (int q = 0; q < npairs; ++q)
            if (pairs[q].first && pairs[q].second) ++uncorrect;
        if (uncorrect == 0) break;
    }
    free(pairs);
    free(F);
}

/* ------------------------------------------------------------------ dof numbering ------- */
int ffo_nloc(int dim, int order)
{
    if (order == 1) return dim + 1;
    return dim == 2 ? 6 : 10;
}

typedef struct { uint64_t key; int32_t val; } hent;

static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

/* insert-or-find; returns pointer to the slot value, *isnew set */
static int32_t *hfind(hent *tab, uint64_t mask, uint64_t key, int *isnew)
{
    uint64_t h = mix64(key) & mask;
    for (;;) {
        if (tab[h].key == key) { *isnew = 0; return &tab[h].val; }
        if (tab[h].key == UINT64_MAX) { tab[h].key = key; *isnew = 1; return &tab[h].val; }
        h = (h + 1) & mask;
    }
}

static hent *hnew(uint64_t want, uint64_t *mask)
{
    uint64_t sz = 16;
    while (sz < 2 * want) sz <<= 1;
    hent *t = (hent *)malloc(sizeof(hent) * sz);
    for (uint64_t i = 0; i < sz; ++i) t[i].key = UINT64_MAX;
    *mask = sz - 1;
    return t;
}

int ffo_p2_nodes_3d(int nv, int nt, const int32_t *conn, int32_t *elem2node)
{
    /* keys: vertex (v,inf) and edge (min,max) in one table; numbered at first encounter walking
     * elements in order, 4 vertices then the 6 edges {01,02,03,12,13,23} (GenericMesh.hpp:1878-1929). */
    uint64_t mask;
    hent *tab = hnew((uint64_t)nt * 3 + (uint64_t)nv + 16, &mask);
    int nn = 0;
    for (int k = 0; k < nt; ++k) {
        const int32_t *K = conn + 4 * (size_t)k;
        for (int a = 0; a < 10; ++a) {
            uint64_t key;
            if (a < 4) key = ((uint64_t)(uint32_t)K[a] << 32) | 0xffffffffu;
            else {
                uint32_t v0 = (uint32_t)K[nvedgeTet[a - 4][0]], v1 = (uint32_t)K[nvedgeTet[a - 4][1]];
                if (v0 > v1) { uint32_t s = v0; v0 = v1; v1 = s; }
                key = ((uint64_t)v0 << 32) | v1;
            }
            int isnew;
            int32_t *pv = hfind(tab, mask, key, &isnew);
            if (isnew) *pv = nn++;
            elem2node[10 * (size_t)k + a] = *pv;
        }
    }
    free(tab);
    return nn;
}

/* ------------------------------------------------------------------ geometry + basis ---- */
/* R3.hpp:90-104 — determinant by Gaussian elimination with partial pivoting on x */
static double det3(const double *a, const double *b, const double *c)
{
    double A[3] = {a[0], a[1], a[2]}, B[3] = {b[0], b[1], b[2]}, C[3] = {c[0], c[1], c[2]}, T[3];
    double s = 1.;
    if (fabs(A[0]) < fabs(B[0])) { memcpy(T, A, 24); memcpy(A, B, 24); memcpy(B, T, 24); s = -s; }
    if (fabs(A[0]) < fabs(C[0])) { memcpy(T, A, 24); memcpy(A, C, 24); memcpy(C, T, 24); s = -s; }
    if (fabs(A[0]) > 1e-50) {
        s *= A[0];
        A[1] /= A[0]; A[2] /= A[0];
        B[1] -= A[1] * B[0]; B[2] -= A[2] * B[0];
        C[1] -= A[1] * C[0]; C[2] -= A[2] * C[0];
        return s * (B[1] * C[2] - B[2] * C[1]);
    }
    return 0.;
}

static void cross3(const double *x, const double *p, double *r)
{ /* R3::operator^ */
    r[0] = x[1] * p[2] - x[2] * p[1];
    r[1] = p[0] * x[2] - x[0] * p[2];
    r[2] = x[0] * p[1] - x[1] * p[0];
}

/* element geometry: measure and grad(lambda_i). X: (dim+1) x dim vertex coords. */
static double geom(int dim, const double *X, double (*G)[3])
{
    if (dim == 3) {
        double V1[3], V2[3], V3[3], c[3];
        for (int d = 0; d < 3; ++d) { V1[d] = X[3 + d] - X[d]; V2[d] = X[6 + d] - X[d]; V3[d] = X[9 + d] - X[d]; }
        double mes = det3(V1, V2, V3) / 6.;
        double det1 = 1. / (6. * mes);
        cross3(V2, V3, c); for (int d = 0; d < 3; ++d) G[1][d] = c[d] * det1;
        cross3(V3, V1, c); for (int d = 0; d < 3; ++d) G[2][d] = c[d] * det1;
        cross3(V1, V2, c); for (int d = 0; d < 3; ++d) G[3][d] = c[d] * det1;
        for (int d = 0; d < 3; ++d) G[0][d] = -G[1][d] - G[2][d] - G[3][d];
        return mes;
    }
    /* triangle: area = ((B-A)^(C-A))*0.5 ; H(i) = (-E.y, E.x)/(2 area), E = v[(i+2)%3]-v[(i+1)%3] */
    double bx = X[2] - X[0], by = X[3] - X[1], cx = X[4] - X[0], cy = X[5] - X[1];
    double area = (bx * cy - by * cx) * 0.5;
    for (int i = 0; i < 3; ++i) {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        double ex = X[2 * i2] - X[2 * i1], ey = X[2 * i2 + 1] - X[2 * i1 + 1];
        G[i][0] = -ey / (2 * area);
        G[i][1] = ex / (2 * area);
        G[i][2] = 0;
    }
    return area;
}

/* basis values at reference point P: val[a][s], s=0 value, s=1..dim gradient (FB of P1/P2) */
static void basis(int dim, int order, const double *P, double (*G)[3], double (*val)[4])
{
    double l[4];
    if (dim == 3) { l[0] = 1. - (P[0] + P[1] + P[2]); l[1] = P[0]; l[2] = P[1]; l[3] = P[2]; }
    else { l[0] = 1 - P[0] - P[1]; l[1] = P[0]; l[2] = P[1]; l[3] = 0; }
    const int nv = dim + 1;
    if (order == 1) {
        for (int a = 0; a < nv; ++a) {
            val[a][0] = l[a];
            for (int d = 0; d < dim; ++d) val[a][1 + d] = G[a][d];
        }
        return;
    }
    double l4[4];
    for (int a = 0; a < nv; ++a) l4[a] = 4 * l[a] - 1;
    for (int a = 0; a < nv; ++a) {
        val[a][0] = l[a] * (2 * l[a] - 1.);
        for (int d = 0; d < dim; ++d) val[a][1 + d] = G[a][d] * l4[a];
    }
    if (dim == 3) {
        for (int e = 0; e < 6; ++e) {
            int i0 = nvedgeTet[e][0], i1 = nvedgeTet[e][1];
            val[4 + e][0] = 4. * l[i0] * l[i1];
            for (int d = 0; d < 3; ++d) val[4 + e][1 + d] = 4 * (G[i1][d] * l[i0] + G[i0][d] * l[i1]);
        }
    } else {
        /* FESpace.cpp:1219-1262: dof 3+e sits on the edge opposite vertex e */
        val[3][0] = 4 * l[1] * l[2]; val[4][0] = 4 * l[0] * l[2]; val[5][0] = 4 * l[1] * l[0];
        for (int d = 0; d < 2; ++d) {
            val[3][1 + d] = 4 * (G[1][d] * l[2] + G[2][d] * l[1]);
            val[4][1 + d] = 4 * (G[2][d] * l[0] + G[0][d] * l[2]);
            val[5][1 + d] = 4 * (G[0][d] * l[1] + G[1][d] * l[0]);
        }
    }
}

static int opslot(int op)
{ /* id,dx,dy,dz -> 0..3 */
    switch (op) { case FFO_OP_ID: return 0; case FFO_OP_DX: return 1; case FFO_OP_DY: return 2; case FFO_OP_DZ: return 3; }
    return -1;
}

static int in_labels(int lab, int nlab, const int32_t *labels)
{
    if (!labels) return 1;
    for (int i = 0; i < nlab; ++i) if (labels[i] == lab) return 1;
    return 0;
}

static void elem_coords(int dim, const double *xyz, const int32_t *K, double *X)
{
    for (int a = 0; a <= dim; ++a)
        for (int d = 0; d < dim; ++d) X[a * dim + d] = xyz[(size_t)K[a] * dim + d];
}

/* ------------------------------------------------------------------ bilinear form ------- */
static int64_t assemble_coo_impl(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                                 int order, int ncomp, const int32_t *elem2node,
                                 int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                                 int nlab, const int32_t *labels, const double *cq,
                                 int32_t *coo_i, int32_t *coo_j, double *coo_a);

int64_t ffo_assemble_coo(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                         int order, int ncomp, const int32_t *elem2node,
                         int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                         int nlab, const int32_t *labels,
                         int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    return assemble_coo_impl(dim, nv, xyz, nt, conn, elab, order, ncomp, elem2node, nterms, terms, nq, qpts, qw, nlab, labels, NULL,
                             coo_i, coo_j, coo_a);
}

/* the same with every term multiplied by a coefficient that depends on the mesh point, given by its values at the
 * quadrature nodes cq[k * nq + q] - what GetAny<R>(ll.second.eval(stack)) returns inside Element_Op (problem.cpp:6407) */
int64_t ffo_assemble_coo_qcoef(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                               int order, int ncomp, const int32_t *elem2node,
                               int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                               const double *cq, int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    return assemble_coo_impl(dim, nv, xyz, nt, conn, elab, order, ncomp, elem2node, nterms, terms, nq, qpts, qw, 0, NULL, cq,
                             coo_i, coo_j, coo_a);
}

static int64_t assemble_coo_impl(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                                 int order, int ncomp, const int32_t *elem2node,
                                 int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                                 int nlab, const int32_t *labels, const double *cq,
                                 int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    (void)nv;
    const int nloc = ffo_nloc(dim, order), nd = nloc * ncomp;
    const int nvk = dim + 1;
    uint64_t mask;
    hent *tab = hnew((uint64_t)nt * (uint64_t)(nd < 8 ? nd * 4 : nd * 8) + 64, &mask);
    double *mat = (double *)malloc(sizeof(double) * (size_t)nd * nd);
    int32_t *gd = (int32_t *)malloc(sizeof(int32_t) * (size_t)nd);
    int64_t nnz = 0;
    for (int k = 0; k < nt; ++k) {
        if (!in_labels(elab ? elab[k] : 0, nlab, labels)) continue;
        const int32_t *K = conn + (size_t)nvk * k;
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * k : K;
        double X[12], G[4][3], val[10][4];
        elem_coords(dim, xyz, K, X);
        double mes = geom(dim, X, G);
        for (int i = 0; i < nd * nd; ++i) mat[i] = 0.;
        /* Element_Op: quadrature point outermost, then terms, then (i,j) */
        for (int q = 0; q < nq; ++q) {
            double coef = mes * qw[q];
            basis(dim, order, qpts + (size_t)q * dim, G, val);
            for (int t = 0; t < nterms; ++t) {
                int so = opslot(terms[t].uop), to = opslot(terms[t].vop);
                double ccc = terms[t].coef;
                if (cq) ccc *= cq[(size_t)k * nq + q];
                ccc *= coef;
                int fi = terms[t].vcomp * nloc, fj = terms[t].ucomp * nloc;
                for (int a = 0; a < nloc; ++a)
                    for (int b = 0; b < nloc; ++b) {
                        double w_i = val[a][to], w_j = val[b][so];
                        mat[(fi + a) * nd + fj + b] += ccc * w_i * w_j;
                    }
            }
        }
        for (int c = 0; c < ncomp; ++c)
            for (int a = 0; a < nloc; ++a) gd[c * nloc + a] = N[a] * ncomp + c;
        /* HashMatrix::operator+=: every (il,jl) couple creates/accumulates an entry */
        for (int il = 0; il < nd; ++il)
            for (int jl = 0; jl < nd; ++jl) {
                uint64_t key = ((uint64_t)(uint32_t)gd[il] << 32) | (uint32_t)gd[jl];
                int isnew;
                int32_t *pv = hfind(tab, mask, key, &isnew);
                if (isnew) {
                    *pv = (int32_t)nnz;
                    coo_i[nnz] = gd[il]; coo_j[nnz] = gd[jl]; coo_a[nnz] = 0.;
                    nnz++;
                }
                coo_a[*pv] += mat[il * nd + jl];
            }
    }
    free(tab); free(mat); free(gd);
    return nnz;
}

/* Rectangular matrices, `matrix B = vb(Uh,Vh)` with two different spaces on the same mesh (rows = dofs of the test space Vh,
 * columns = dofs of the space of the unknown Uh): Element_Op with Ku != Kv (fflib/problem.cpp:6337-6437: `same` false, fu and
 * fv tabulated separately, n = Kv.NbDoF, m = Ku.NbDoF), every (il, jl) couple of the n x m element matrix added to the
 * HashMatrix (femlib/HashMatrix.cpp:1295-1332).  Each side: P1 / P2, 1..3 components; e2n_* may be NULL for P1.
 * coo arrays must hold nt * (nloc_v*ncomp_v) * (nloc_u*ncomp_u) entries.  Returns nnz. */
int64_t ffo_assemble_coo_rect(int dim, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                              int order_v, int ncomp_v, const int32_t *e2n_v, int order_u, int ncomp_u, const int32_t *e2n_u,
                              int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                              int nlab, const int32_t *labels, int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    const int nlv = ffo_nloc(dim, order_v), nlu = ffo_nloc(dim, order_u), ndv = nlv * ncomp_v, ndu = nlu * ncomp_u;
    const int nvk = dim + 1;
    uint64_t mask;
    hent *tab = hnew((uint64_t)nt * (uint64_t)(ndv * ndu < 64 ? 32 : ndv * ndu) + 64, &mask);
    double *mat = (double *)malloc(sizeof(double) * (size_t)ndv * ndu);
    int32_t *gv = (int32_t *)malloc(sizeof(int32_t) * (size_t)ndv), *gu = (int32_t *)malloc(sizeof(int32_t) * (size_t)ndu);
    int64_t nnz = 0;
    for (int k = 0; k < nt; ++k) {
        if (!in_labels(elab ? elab[k] : 0, nlab, labels)) continue;
        const int32_t *K = conn + (size_t)nvk * k;
        const int32_t *Nv = e2n_v ? e2n_v + (size_t)nlv * k : K, *Nu = e2n_u ? e2n_u + (size_t)nlu * k : K;
        double X[12], G[4][3], fv[10][4], fu[10][4];
        elem_coords(dim, xyz, K, X);
        double mes = geom(dim, X, G);
        for (int i = 0; i < ndv * ndu; ++i) mat[i] = 0.;
        for (int q = 0; q < nq; ++q) {
            double coef = mes * qw[q];
            basis(dim, order_u, qpts + (size_t)q * dim, G, fu);
            basis(dim, order_v, qpts + (size_t)q * dim, G, fv);
            for (int t = 0; t < nterms; ++t) {
                int so = opslot(terms[t].uop), to = opslot(terms[t].vop);
                double ccc = terms[t].coef;
                ccc *= coef;
                int fi = terms[t].vcomp * nlv, fj = terms[t].ucomp * nlu;
                for (int a = 0; a < nlv; ++a)
                    for (int b = 0; b < nlu; ++b) {
                        double w_i = fv[a][to], w_j = fu[b][so];
                        mat[(fi + a) * ndu + fj + b] += ccc * w_i * w_j;
                    }
            }
        }
        for (int c = 0; c < ncomp_v; ++c)
            for (int a = 0; a < nlv; ++a) gv[c * nlv + a] = Nv[a] * ncomp_v + c;
        for (int c = 0; c < ncomp_u; ++c)
            for (int a = 0; a < nlu; ++a) gu[c * nlu + a] = Nu[a] * ncomp_u + c;
        for (int il = 0; il < ndv; ++il)
            for (int jl = 0; jl < ndu; ++jl) {
                uint64_t key = ((uint64_t)(uint32_t)gv[il] << 32) | (uint32_t)gu[jl];
                int isnew;
                int32_t *pv = hfind(tab, mask, key, &isnew);
                if (isnew) {
                    *pv = (int32_t)nnz;
                    coo_i[nnz] = gv[il]; coo_j[nnz] = gu[jl]; coo_a[nnz] = 0.;
                    nnz++;
                }
                coo_a[*pv] += mat[il * ndu + jl];
            }
    }
    free(tab); free(mat); free(gv); free(gu);
    return nnz;
}

typedef struct { int32_t i, j; int64_t k; } ijk;
static int cmp_ij(const void *a, const void *b)
{
    const ijk *x = (const ijk *)a, *y = (const ijk *)b;
    if (x->i != y->i) return x->i < y->i ? -1 : 1;
    if (x->j != y->j) return x->j < y->j ? -1 : 1;
    return 0;
}

void ffo_coo_to_csr(int n, int64_t nnz, const int32_t *coo_i, const int32_t *coo_j, const double *coo_a,
                    int32_t *rowptr, int32_t *colind, double *vals)
{
    ijk *s = (ijk *)malloc(sizeof(ijk) * (size_t)(nnz ? nnz : 1));
    for (int64_t k = 0; k < nnz; ++k) { s[k].i = coo_i[k]; s[k].j = coo_j[k]; s[k].k = k; }
    qsort(s, (size_t)nnz, sizeof(ijk), cmp_ij);
    for (int i = 0; i <= n; ++i) rowptr[i] = 0;
    for (int64_t k = 0; k < nnz; ++k) {
        rowptr[s[k].i + 1]++;
        colind[k] = s[k].j;
        if (vals) vals[k] = coo_a[s[k].k];
    }
    for (int i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
    free(s);
}

/* ------------------------------------------------------------------ linear form --------- */
void ffo_assemble_rhs(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                      int order, int ncomp, const int32_t *elem2node, int ndof,
                      int nterms, const ffo_lterm *terms, int nq, const double *qpts, const double *qw,
                      int nlab, const int32_t *labels, double *b)
{
    (void)nv;
    const int nloc = ffo_nloc(dim, order), nvk = dim + 1;
    for (int i = 0; i < ndof; ++i) b[i] = 0.;
    for (int k = 0; k < nt; ++k) {
        if (!in_labels(elab ? elab[k] : 0, nlab, labels)) continue;
        const int32_t *K = conn + (size_t)nvk * k;
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * k : K;
        double X[12], G[4][3], val[10][4];
        elem_coords(dim, xyz, K, X);
        double mes = geom(dim, X, G);
        for (int q = 0; q < nq; ++q) {
            double coef = mes * qw[q];
            basis(dim, order, qpts + (size_t)q * dim, G, val);
            for (int c = 0; c < ncomp; ++c)
                for (int a = 0; a < nloc; ++a)
                    for (int t = 0; t < nterms; ++t) {
                        /* fu(i,comp,op) vanishes unless comp is the component of local dof i */
                        double w_i = (terms[t].vcomp == c) ? val[a][opslot(terms[t].vop)] : 0.;
                        double av = coef * terms[t].coef * w_i;
                        b[N[a] * ncomp + c] += av;
                    }
        }
    }
}

/* Linear form with data depending on the mesh point: Element_rhs (fflib/problem.cpp:7917-7985) with the value of the
 * coefficient expression at node q of element k given as fq[(c * nt + k) * nq + q] (summed over the value terms of
 * component c).  ADDS to b. */
void ffo_assemble_rhs_qvalues(int dim, const double *xyz, int nt, const int32_t *conn, int order, int ncomp,
                              const int32_t *elem2node, int nq, const double *qpts, const double *qw, const double *fq, double *b)
{
    const int nloc = ffo_nloc(dim, order), nvk = dim + 1;
    for (int k = 0; k < nt; ++k) {
        const int32_t *K = conn + (size_t)nvk * k;
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * k : K;
        double X[12], G[4][3], val[10][4];
        elem_coords(dim, xyz, K, X);
        double mes = geom(dim, X, G);
        for (int q = 0; q < nq; ++q) {
            double coef = mes * qw[q];
            basis(dim, order, qpts + (size_t)q * dim, G, val);
            for (int c = 0; c < ncomp; ++c)
                for (int a = 0; a < nloc; ++a) b[N[a] * ncomp + c] += coef * fq[((size_t)c * nt + k) * nq + q] * val[a][0];
        }
    }
}

/* the same with derivatives of the test function: fq[((c * (dim+1) + s) * nt + k) * nq + q] = coefficient of d^s v_c
 * (s = 0 value, 1..dim = dx, dy, dz) at node q of element k - e.g. the residual of a Newton step, int(dx(uk) dx(v) + ...) */
void ffo_assemble_rhs_qterms(int dim, const double *xyz, int nt, const int32_t *conn, int order, int ncomp,
                             const int32_t *elem2node, int nq, const double *qpts, const double *qw, const double *fq, double *b)
{
    const int nloc = ffo_nloc(dim, order), nvk = dim + 1, ns = dim + 1;
    for (int k = 0; k < nt; ++k) {
        const int32_t *K = conn + (size_t)nvk * k;
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * k : K;
        double X[12], G[4][3], val[10][4];
        elem_coords(dim, xyz, K, X);
        double mes = geom(dim, X, G);
        for (int q = 0; q < nq; ++q) {
            double coef = mes * qw[q];
            basis(dim, order, qpts + (size_t)q * dim, G, val);
            for (int c = 0; c < ncomp; ++c)
                for (int a = 0; a < nloc; ++a)
                    for (int sl = 0; sl < ns; ++sl)
                        b[N[a] * ncomp + c] += coef * fq[(((size_t)c * ns + sl) * nt + k) * nq + q] * val[a][sl];
        }
    }
}

/* measure factor and reference point of a border quadrature node: T.N(ie) (cross product of two edges of the face,
 * norm = 2 * area, hence the 0.5) / edge length, and Pt = PBord(ie, pi) (femlib/Mesh3dn.hpp:76, Mesh2dn.hpp:65) */
static double face_measure(int dim, const double *X, int ie)
{
    if (dim == 3) {
        const double *A = X + 3 * nvfaceTet[ie][0], *B = X + 3 * nvfaceTet[ie][1], *Cc = X + 3 * nvfaceTet[ie][2];
        double u[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, v[3] = {Cc[0] - A[0], Cc[1] - A[1], Cc[2] - A[2]}, nn[3];
        cross3(u, v, nn);
        return 0.5 * sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
    }
    const double *A = X + 2 * nvedgeTri[ie][0], *B = X + 2 * nvedgeTri[ie][1];
    return sqrt((B[0] - A[0]) * (B[0] - A[0]) + (B[1] - A[1]) * (B[1] - A[1]));
}
static void face_point(int dim, int ie, const double *qp, double *P)
{
    static const double hat3[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static const double hat2[3][2] = {{0, 0}, {1, 0}, {0, 1}};
    if (dim == 3) {
        const double x = qp[0], y = qp[1];
        for (int d = 0; d < 3; ++d)
            P[d] = hat3[nvfaceTet[ie][0]][d] * (1 - x - y) + hat3[nvfaceTet[ie][1]][d] * x + hat3[nvfaceTet[ie][2]][d] * y;
    } else {
        const double x = qp[0];
        for (int d = 0; d < 2; ++d) P[d] = hat2[nvedgeTri[ie][0]][d] * (1 - x) + hat2[nvedgeTri[ie][1]][d] * x;
    }
}

/* Boundary integrals of a linear form, int2d(Th3,labels)(...) / int1d(Th,labels)(...): Element_rhs on a border element
 * (fflib/problem.cpp:8517-8587 in 3-D, :8439-8513 in 2-D): for every boundary element with a listed label, every
 * quadrature point of the face rule, every dof of the ADJACENT ELEMENT: B[dof] += (face measure * w_q) * c * d^op phi_i(Pt),
 * Pt = PBord(ie, pi).  Adds to b (the caller zeroes it). */
static void rhs_boundary_impl(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                              const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                              const int32_t *bface, int nterms, const ffo_lterm *terms, int nq, const double *qpts,
                              const double *qw, int nlab, const int32_t *labels, const double *gq, double *b);
void ffo_assemble_rhs_boundary(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                               const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                               const int32_t *bface, int nterms, const ffo_lterm *terms, int nq, const double *qpts,
                               const double *qw, int nlab, const int32_t *labels, double *b)
{
    rhs_boundary_impl(dim, xyz, conn, order, ncomp, elem2node, nbe, blab, belem, bface, nterms, terms, nq, qpts, qw, nlab, labels, NULL, b);
}
/* the same with data depending on the mesh point: gq[(c * nbe + ib) * nq + q] = coefficient of the value of v_c at node q
 * of boundary element ib (0 where the label is not listed), as evaluated inside Element_rhs (problem.cpp:8551-8570) */
void ffo_assemble_rhs_boundary_qvalues(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                       const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                       const int32_t *bface, int nq, const double *qpts, const double *qw, const double *gq, double *b)
{
    rhs_boundary_impl(dim, xyz, conn, order, ncomp, elem2node, nbe, blab, belem, bface, 0, NULL, nq, qpts, qw, 0, NULL, gq, b);
}
static void rhs_boundary_impl(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                              const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                              const int32_t *bface, int nterms, const ffo_lterm *terms, int nq, const double *qpts,
                              const double *qw, int nlab, const int32_t *labels, const double *gq, double *b)
{
    const int nloc = ffo_nloc(dim, order), nvk = dim + 1;
    for (int ib = 0; ib < nbe; ++ib) {
        if (!in_labels(blab[ib], nlab, labels)) continue;
        const int it = belem[ib], ie = bface[ib];
        const int32_t *K = conn + (size_t)nvk * it;
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * it : K;
        double X[12], G[4][3], val[10][4];
        elem_coords(dim, xyz, K, X);
        geom(dim, X, G);
        const double le = face_measure(dim, X, ie);
        for (int q = 0; q < nq; ++q) {
            double P[3];
            face_point(dim, ie, qpts + (size_t)q * (dim - 1), P);
            const double coef = le * qw[q];
            basis(dim, order, P, G, val);
            for (int c = 0; c < ncomp; ++c)
                for (int a = 0; a < nloc; ++a) {
                    for (int t = 0; t < nterms; ++t) {
                        double w_i = (terms[t].vcomp == c) ? val[a][opslot(terms[t].vop)] : 0.;
                        b[N[a] * ncomp + c] += coef * terms[t].coef * w_i;
                    }
                    if (gq) b[N[a] * ncomp + c] += coef * gq[((size_t)c * nbe + ib) * nq + q] * val[a][0];
                }
        }
    }
}

/* Boundary integrals of a bilinear form (Robin terms), int2d(Th3,labels)(c u v) / int1d(Th,labels)(c u v):
 * AssembleBilinearForm's loop over the border elements (fflib/problem.cpp:1317-1326 in 3-D, :1030-1040 in 2-D) calls
 * MatriceElementairePleine::call(k, ie, label) on the ADJACENT ELEMENT k, whose Element_Op border branch
 * (:6518-6560 3-D, :6216-6290 2-D) fills the whole n x m element matrix at the face quadrature nodes, and
 * HashMatrix::operator+= then creates/accumulates every (il, jl) couple of that element, zero or not.
 * Output: COO with one entry per distinct couple (first-encounter order); returns their number. */
static int64_t coo_boundary_impl(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                 const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                 const int32_t *bface, int nterms, const ffo_bterm *terms, int nq, const double *qpts,
                                 const double *qw, int nlab, const int32_t *labels, const double *cq,
                                 int32_t *coo_i, int32_t *coo_j, double *coo_a);
int64_t ffo_assemble_coo_boundary(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                  const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                  const int32_t *bface, int nterms, const ffo_bterm *terms, int nq, const double *qpts,
                                  const double *qw, int nlab, const int32_t *labels,
                                  int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    return coo_boundary_impl(dim, xyz, conn, order, ncomp, elem2node, nbe, blab, belem, bface, nterms, terms, nq, qpts, qw, nlab, labels,
                             NULL, coo_i, coo_j, coo_a);
}
/* the same with every term multiplied by a coefficient depending on the mesh point, cq[ib * nq + q]; labels as above */
int64_t ffo_assemble_coo_boundary_qcoef(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                        const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                        const int32_t *bface, int nterms, const ffo_bterm *terms, int nq, const double *qpts,
                                        const double *qw, int nlab, const int32_t *labels, const double *cq,
                                        int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    return coo_boundary_impl(dim, xyz, conn, order, ncomp, elem2node, nbe, blab, belem, bface, nterms, terms, nq, qpts, qw, nlab, labels,
                             cq, coo_i, coo_j, coo_a);
}
static int64_t coo_boundary_impl(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                 const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                 const int32_t *bface, int nterms, const ffo_bterm *terms, int nq, const double *qpts,
                                 const double *qw, int nlab, const int32_t *labels, const double *cq,
                                 int32_t *coo_i, int32_t *coo_j, double *coo_a)
{
    const int nloc = ffo_nloc(dim, order), nd = nloc * ncomp, nvk = dim + 1;
    uint64_t mask;
    hent *tab = hnew((uint64_t)(nbe > 0 ? nbe : 1) * (uint64_t)nd * nd * 2 + 64, &mask);
    double *mat = (double *)malloc(sizeof(double) * (size_t)nd * nd);
    int32_t *gd = (int32_t *)malloc(sizeof(int32_t) * (size_t)nd);
    int64_t nnz = 0;
    for (int ib = 0; ib < nbe; ++ib) {
        if (!in_labels(blab[ib], nlab, labels)) continue;
        const int it = belem[ib], ie = bface[ib];
        const int32_t *K = conn + (size_t)nvk * it;
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * it : K;
        double X[12], G[4][3], val[10][4];
        elem_coords(dim, xyz, K, X);
        geom(dim, X, G);
        const double le = face_measure(dim, X, ie);
        for (int i = 0; i < nd * nd; ++i) mat[i] = 0.;
        for (int q = 0; q < nq; ++q) {
            double P[3];
            face_point(dim, ie, qpts + (size_t)q * (dim - 1), P);
            const double coef = le * qw[q];
            basis(dim, order, P, G, val);
            for (int t = 0; t < nterms; ++t) {
                const int so = opslot(terms[t].uop), to = opslot(terms[t].vop);
                const double ccc = terms[t].coef * (cq ? cq[(size_t)ib * nq + q] : 1.) * coef;
                const int fi = terms[t].vcomp * nloc, fj = terms[t].ucomp * nloc;
                for (int a = 0; a < nloc; ++a)
                    for (int b = 0; b < nloc; ++b) mat[(fi + a) * nd + fj + b] += ccc * val[a][to] * val[b][so];
            }
        }
        for (int c = 0; c < ncomp; ++c)
            for (int a = 0; a < nloc; ++a) gd[c * nloc + a] = N[a] * ncomp + c;
        for (int il = 0; il < nd; ++il)
            for (int jl = 0; jl < nd; ++jl) {
                uint64_t key = ((uint64_t)(uint32_t)gd[il] << 32) | (uint32_t)gd[jl];
                int isnew;
                int32_t *pv = hfind(tab, mask, key, &isnew);
                if (isnew) {
                    *pv = (int32_t)nnz;
                    coo_i[nnz] = gd[il]; coo_j[nnz] = gd[jl]; coo_a[nnz] = 0.;
                    nnz++;
                }
                coo_a[*pv] += mat[il * nd + jl];
            }
    }
    free(tab); free(mat); free(gd);
    return nnz;
}

/* ------------------------------------------------------------------ Dirichlet ----------- */
int ffo_bc_pairs(int dim, int nt, const int32_t *conn, int order, int ncomp, const int32_t *elem2node,
                 int nbe, const int32_t *blab, const int32_t *belem, const int32_t *bface,
                 int nlab, const int32_t *labels, int compmask, const double *values,
                 int32_t *out_dof, double *out_val)
{
    (void)nt;
    const int nloc = ffo_nloc(dim, order), nvk = dim + 1;
    int n = 0;
    for (int ib = 0; ib < nbe; ++ib) {
        if (!in_labels(blab[ib], nlab, labels)) continue;
        int it = belem[ib], ie = bface[ib];
        const int32_t *N = elem2node ? elem2node + (size_t)nloc * it : conn + (size_t)nvk * it;
        for (int c = 0; c < ncomp; ++c) {
            if (!(compmask >> c & 1)) continue;
            for (int a = 0; a < nloc; ++a) {
                int on;
                if (a < nvk) on = (a != ie); /* face ie is opposite vertex ie */
                else if (dim == 2) on = (a - 3 == ie);
                else on = (nvedgeTet[a - 4][0] != ie && nvedgeTet[a - 4][1] != ie);
                if (!on) continue;
                out_dof[n] = N[a] * ncomp + c;
                out_val[n] = values ? values[c] : 0.;
                n++;
            }
        }
    }
    (void)nvedgeTri;
    return n;
}

/* HashMatrix::SetBC (femlib/HashMatrix.cpp:1195-1238).  tgv >= 0: diag = tgv.  tgv < 0 (exact elimination): on the rows
 * of the Dirichlet dofs diag = 1 (0 when tgv < -9), the other entries 0 unless tgv is -3 / -30; for tgv = -2, -20, -3,
 * -30 the columns of those dofs are zeroed too (the diagonal as well when tgv < -19). */
void ffo_bc_matrix_coo(int64_t nnz, const int32_t *coo_i, const int32_t *coo_j, double *coo_a,
                       int n, int nbc, const int32_t *dofs, double tgv)
{
    char *on = (char *)calloc((size_t)n, 1);
    for (int k = 0; k < nbc; ++k) on[dofs[k]] = 1;
    if (tgv >= 0) {
        for (int64_t k = 0; k < nnz; ++k)
            if (coo_i[k] == coo_j[k] && on[coo_i[k]]) coo_a[k] = tgv;
    } else {
        const int keeprow = fabs(tgv + 3.0) <= 1.0e-10 || fabs(tgv + 30.0) <= 1.0e-10;
        const int cols = fabs(tgv + 2.0) < 1.0e-10 || fabs(tgv + 20.0) < 1.0e-10 || fabs(tgv + 3.0) < 1.0e-10 ||
                         fabs(tgv + 30.0) < 1.0e-10;
        for (int64_t k = 0; k < nnz; ++k)
            if (on[coo_i[k]]) {
                if (coo_j[k] == coo_i[k]) coo_a[k] = tgv < -9.0 ? 0.0 : 1.0;
                else if (!keeprow) coo_a[k] = 0.0;
            }
        if (cols)
            for (int64_t k = 0; k < nnz; ++k)
                if (on[coo_j[k]] && (coo_i[k] != coo_j[k] || tgv < -19.0)) coo_a[k] = 0.0;
    }
    free(on);
}

/* AssembleBC: B[dof] = tgv1 * g with tgv1 = tgv >= 0 ? tgv : 1 (fflib/problem.cpp:10099,10176; rank 0 / sequential) */
void ffo_bc_rhs(double *b, int nbc, const int32_t *dofs, const double *vals, double tgv)
{
    const double tgv1 = tgv < 0 ? 1.0 : tgv;
    for (int k = 0; k < nbc; ++k) b[dofs[k]] = tgv1 * vals[k];
}

/* ------------------------------------------------------------------ SpMV + CG ----------- */
void ffo_spmv_coo(int n, int64_t nnz, const int32_t *ai, const int32_t *aj, const double *aa,
                  const double *x, double *y)
{
    for (int i = 0; i < n; ++i) y[i] = 0.; /* CGMatVirt::matmul zeroes, then addMatMul */
    for (int64_t k = 0; k < nnz; ++k) y[ai[k]] += aa[k] * x[aj[k]];
}

static double dotp(int n, const double *x, const double *y)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

int ffo_cg(int n, int64_t nnz, const int32_t *ai, const int32_t *aj, const double *aa,
           const double *b, double *x, double eps, int itmax, double tgv, int *iters, double *gcg_out)
{
    double *diag = (double *)calloc((size_t)n, sizeof(double));
    char *has = (char *)calloc((size_t)n, 1);
    for (int64_t k = 0; k < nnz; ++k)
        if (ai[k] == aj[k]) { diag[ai[k]] = aa[k]; has[ai[k]] = 1; }
    /* gettgv (HashMatrix.cpp:1341-1371), ratio 1e6 */
    double ttgv = 0, max1 = 0; int ntgv = 0;
    for (int i = 0; i < n; ++i)
        if (has[i]) {
            double a = diag[i];
            if (a > ttgv) { max1 = ttgv; ttgv = a; ntgv = 1; }
            else if (a == ttgv) ++ntgv;
            else if (a > max1) max1 = a;
        }
    if (max1 * 1e6 > ttgv) { ttgv = 0; ntgv = 0; }
    /* HMatVirtPrecon ctor: wcl, diag1 */
    double *d1 = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) d1[i] = (diag[i] * diag[i] < 1e-60) ? 1. : 1. / diag[i];
    if (ntgv) { /* SetInitWithBC */
        double tgve = ttgv <= 0 ? 1e200 : ttgv;
        for (int i = 0; i < n; ++i)
            if (diag[i] == tgve) x[i] = b[i] / tgv;
    }
    if (itmax <= 0) itmax = n;
    /* ConjugueGradient (CG.cpp:195-265); AH aliases CG */
    double *G = (double *)malloc(sizeof(double) * (size_t)n);
    double *CG = (double *)malloc(sizeof(double) * (size_t)n);
    double *H = (double *)malloc(sizeof(double) * (size_t)n);
    double *AH = CG;
    double eps2 = eps * eps, gCg, gCgp, rho, gamma;
    int ret = 0, it = 0;
    ffo_spmv_coo(n, nnz, ai, aj, aa, x, G);
    for (int i = 0; i < n; ++i) G[i] += -1. * b[i];
    for (int i = 0; i < n; ++i) { CG[i] = 0.; CG[i] += d1[i] * G[i]; }
    gCg = dotp(n, G, CG);
    for (int i = 0; i < n; ++i) H[i] = CG[i];
    for (int i = 0; i < n; ++i) H[i] *= -1.;
    if (eps > 0) eps2 *= gCg;
    if (gCg < 1e-30) { ret = 2; it = 0; }
    else
        for (int iter = 1; iter <= itmax; ++iter) {
            gCgp = gCg;
            double gh = dotp(n, G, H);
            ffo_spmv_coo(n, nnz, ai, aj, aa, H, AH);
            rho = -gh / dotp(n, H, AH);
            for (int i = 0; i < n; ++i) x[i] += rho * H[i];
            for (int i = 0; i < n; ++i) G[i] += rho * AH[i];
            for (int i = 0; i < n; ++i) { CG[i] = 0.; CG[i] += d1[i] * G[i]; }
            gCg = dotp(n, G, CG);
            gamma = gCg / gCgp;
            for (int i = 0; i < n; ++i) H[i] *= gamma;
            for (int i = 0; i < n; ++i) H[i] += -1. * CG[i];
            it = iter;
            if (gCg < eps2) { ret = 1; break; }
        }
    if (iters) *iters = it;
    if (gcg_out) *gcg_out = gCg;
    free(G); free(CG); free(H); free(d1); free(diag); free(has);
    return ret;
}

/* ------------------------------------------------------------------ GMRES ---------------- */
/* SolverGMRES::dosolver (femlib/VirtualSolverCG.hpp:236-255) = SetInitWithBC, then fgmres (femlib/CG.cpp:347-517) with the
 * Jacobi preconditioner of HMatVirtPrecon (VirtualSolverCG.hpp:24-70) applied on the right (leftC is forced to 0, :374):
 * flexible GMRES(nbkrylov), modified Gram-Schmidt, Givens rotations, stop when |g[it+1]| / normb < |eps| where
 * normb = || rhs with the tgv rows zeroed ||.  x holds the initial guess on entry.  Returns 1 when converged;
 * *iters = the reference's `iter` (completed inner iterations before the one that converged). */
static double nrm2(int n, const double *x)
{
    double s = 0.;
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}
int ffo_gmres(int n, int64_t nnz, const int32_t *ai, const int32_t *aj, const double *aa,
              const double *b, double *x, double eps, int itmax, int nbkrylov, double tgv, int *iters, double *relres_out)
{
    double *diag = (double *)calloc((size_t)n, sizeof(double));
    char *has = (char *)calloc((size_t)n, 1);
    for (int64_t k = 0; k < nnz; ++k)
        if (ai[k] == aj[k]) { diag[ai[k]] = aa[k]; has[ai[k]] = 1; }
    double ttgv = 0, max1 = 0; int ntgv = 0; /* gettgv, as in ffo_cg */
    for (int i = 0; i < n; ++i)
        if (has[i]) {
            double a = diag[i];
            if (a > ttgv) { max1 = ttgv; ttgv = a; ntgv = 1; }
            else if (a == ttgv) ++ntgv;
            else if (a > max1) max1 = a;
        }
    if (max1 * 1e6 > ttgv) { ttgv = 0; ntgv = 0; }
    double *d1 = (double *)malloc(sizeof(double) * (size_t)n);
    char *wbc = (char *)calloc((size_t)n, 1);
    for (int i = 0; i < n; ++i) d1[i] = (diag[i] * diag[i] < 1e-60) ? 1. : 1. / diag[i];
    if (ntgv) {
        double tgve = ttgv <= 0 ? 1e200 : ttgv;
        for (int i = 0; i < n; ++i)
            if (diag[i] == tgve) { wbc[i] = 1; x[i] = b[i] / tgv; } /* wcl + SetInitWithBC */
    }
    if (itmax <= 0) itmax = n;
    if (nbkrylov > itmax + 2) nbkrylov = itmax + 2; /* storage only: the inner loop leaves at it > itmax */
    const int m = nbkrylov;
    double *rot0 = (double *)calloc((size_t)m + 2, sizeof(double)), *rot1 = (double *)calloc((size_t)m + 2, sizeof(double));
    double *g = (double *)calloc((size_t)m + 1, sizeof(double)), *g1 = (double *)calloc((size_t)m + 1, sizeof(double));
    double *y = (double *)calloc((size_t)m + 1, sizeof(double));
    double *Hn = (double *)calloc((size_t)(m + 2) * (m + 1), sizeof(double));
#define HN(i, j) Hn[(size_t)(i) * (m + 1) + (j)]
    double **Vi = (double **)calloc((size_t)m + 1, sizeof(double *)), **Vpi = (double **)calloc((size_t)m + 1, sizeof(double *));
    double *uni = (double *)malloc(sizeof(double) * (size_t)n), *x0 = (double *)malloc(sizeof(double) * (size_t)n);
    double *ri = (double *)malloc(sizeof(double) * (size_t)n), *vi = (double *)malloc(sizeof(double) * (size_t)n);
    double *wi = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) wi[i] = wbc[i] ? 0. : b[i];
    double normb = nrm2(n, wi), relres = 1e100;
    for (int i = 0; i < n; ++i) uni[i] = x[i];
    int noconv = 1, iter = 0;
    while (noconv) {
        for (int i = 0; i < n; ++i) x0[i] = uni[i];
        ffo_spmv_coo(n, nnz, ai, aj, aa, uni, ri);
        for (int i = 0; i < n; ++i) ri[i] += -1. * b[i];
        for (int i = 0; i < n; ++i) ri[i] *= -1.;
        g[0] = nrm2(n, ri); /* zi = Id ri */
        if (normb < 1.e-20 || eps < 0) normb = 1.;
        if (!Vi[0]) Vi[0] = (double *)malloc(sizeof(double) * (size_t)n);
        { const double s = 1. / g[0]; for (int i = 0; i < n; ++i) Vi[0][i] = s * ri[i]; }
        int it;
        for (it = 0; it < m; it++, iter++) {
            if (!Vpi[it]) Vpi[it] = (double *)malloc(sizeof(double) * (size_t)n);
            if (!Vi[it + 1]) Vi[it + 1] = (double *)malloc(sizeof(double) * (size_t)n);
            for (int i = 0; i < n; ++i) { vi[i] = 0.; vi[i] += d1[i] * Vi[it][i]; } /* C.matmul */
            for (int i = 0; i < n; ++i) Vpi[it][i] = vi[i];
            ffo_spmv_coo(n, nnz, ai, aj, aa, vi, wi);
            for (int i = 0; i < it + 1; i++) {
                HN(i, it) = dotp(n, wi, Vi[i]);
                const double a = -HN(i, it);
                for (int k = 0; k < n; ++k) wi[k] += a * Vi[i][k];
            }
            const double aux = HN(it + 1, it) = nrm2(n, wi);
            { const double s = 1. / aux; for (int k = 0; k < n; ++k) Vi[it + 1][k] = s * wi[k]; }
            for (int i = 0; i < it; i++) {
                double aa_ = rot0[i] * HN(i, it) + rot1[i] * HN(i + 1, it);
                double bb_ = -rot1[i] * HN(i, it) + rot0[i] * HN(i + 1, it);
                HN(i, it) = aa_;
                HN(i + 1, it) = bb_;
            }
            const double sq = sqrt(HN(it, it) * HN(it, it) + HN(it + 1, it) * HN(it + 1, it));
            rot0[it] = HN(it, it) / sq;
            rot1[it] = HN(it + 1, it) / sq;
            HN(it, it) = rot0[it] * HN(it, it) + rot1[it] * HN(it + 1, it);
            HN(it + 1, it) = 0.;
            g[it + 1] = -rot1[it] * g[it];
            g[it] = rot0[it] * g[it];
            relres = fabs(g[it + 1]);
            if (relres / normb < fabs(eps)) { noconv = 0; break; }
            if (it > itmax) break;
        }
        if (it > m - 1) it = m - 1;
        for (int i = it; i >= 0; i--) {
            g1[i] = g[i];
            for (int j = i + 1; j < it + 1; j++) g1[i] = g1[i] - HN(i, j) * y[j];
            y[i] = g1[i] / HN(i, i);
        }
        for (int k = 0; k < n; ++k) wi[k] = 0.;
        for (int i = 0; i < it + 1; i++)
            for (int k = 0; k < n; ++k) wi[k] += y[i] * Vpi[i][k];
        for (int k = 0; k < n; ++k) uni[k] = wi[k];
        for (int k = 0; k < n; ++k) uni[k] += x0[k];
        for (int k = 0; k < n; ++k) x[k] = uni[k];
        if (!noconv) break;
        if (iter > itmax) break;
    }
#undef HN
    if (iters) *iters = iter;
    if (relres_out) *relres_out = relres / normb;
    for (int i = 0; i <= m; ++i) { free(Vi[i]); free(Vpi[i]); }
    free(Vi); free(Vpi); free(uni); free(x0); free(ri); free(vi); free(wi); free(Hn); free(rot0); free(rot1); free(g); free(g1);
    free(y); free(d1); free(wbc); free(diag); free(has);
    return !noconv;
}
