/* fforacle.h — CPU ORACLE for the ffcuda hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of FreeFEM 4.15's algorithm for
 *   varf -> element loop -> MatriceMorse (COO/CSR) + right-hand side -> Jacobi-CG
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product (libffcuda_core.so,
 * ffcuda.so) never links or calls it.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks every function here against
 * the fixtures in tests/golden/ (npz files), which were dumped from the unmodified reference
 * built by `make -C oracle ref` (tests/golden/make_golden.py).
 *
 * Reference files restated (paths under /root/reference/src):
 *   fflib/msh3.cpp:7683-7742,7879-8132      BuildCube        -> ffo_cube
 *   fflib/lgmesh.cpp:1229-1384              Carre_           -> ffo_square
 *   femlib/GenericMesh.hpp:1711-1954        BuildDFNumbering -> ffo_p2_nodes_3d
 *   femlib/QuadratureFormular.cpp           rule tables      -> ffo_quadrature
 *   femlib/Mesh3dn.hpp:65-71,126-136, R3.hpp:90-104, fem.hpp:277,321-324   geometry
 *   femlib/P012_3d.cpp:123-300, FESpace.cpp:1100-1136,1219-1262            P1/P2 basis
 *   fflib/problem.cpp:6063-6160,6337-6437   Element_Op       -> ffo_assemble_coo
 *   femlib/HashMatrix.cpp:1295-1332,671-682,993-1030   += / Sortij / Buildp -> COO order, CSR
 *   fflib/problem.cpp:7839-7985             Element_rhs      -> ffo_assemble_rhs
 *   fflib/problem.cpp:9881-10194, HashMatrix.cpp:1195-1238   AssembleBC/SetBC -> ffo_bc_*
 *   femlib/VirtualSolverCG.hpp:13-192, CG.cpp:195-265, HashMatrix.cpp:1087-1154,1341-1371
 *                                            SolverCG         -> ffo_cg
 */
#ifndef FFORACLE_H
#define FFORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* operator codes = FreeFEM's (femlib/FESpacen.hpp:73-82) */
enum { FFO_OP_ID = 0, FFO_OP_DX = 1, FFO_OP_DY = 2, FFO_OP_DZ = 6 };

typedef struct { int32_t ucomp, uop, vcomp, vop; double coef; } ffo_bterm; /* unknown=column, test=row */
typedef struct { int32_t vcomp, vop; double coef; } ffo_lterm;

/* quadrature rules by FreeFEM name: 2-D "qf1pT" "qf1pTlump" "qf2pT" "qf5pT"(default);
 * 3-D "qfV1" "qfV1lump" "qfV2" "qfV5"(default).  pts: npts*dim reference coords. Returns npts (<=16) or -1. */
int ffo_quadrature(int dim, const char *name, double *pts, double *w);

/* meshes.  belem/bface = Th.BoundaryElement(ib, ie). */
void ffo_cube_sizes(int nx, int ny, int nz, int *nv, int *nt, int *nbe);
void ffo_cube(int nx, int ny, int nz, double *xyz, int32_t *conn, int32_t *elab,
              int32_t *bconn, int32_t *blab, int32_t *belem, int32_t *bface);
void ffo_square_sizes(int nx, int ny, int *nv, int *nt, int *nbe);
void ffo_square(int nx, int ny, double *xy, int32_t *conn, int32_t *elab,
                int32_t *bconn, int32_t *blab, int32_t *belem, int32_t *bface);

/* buildlayers (fflib/msh3.cpp:895-934 build_layer, :936-976 sizes, :978-1668 vertices / boundary / tetrahedra,
 * :1670-1757 dpent1_mesh): the layered 3-D mesh over a 2-D mesh.  Inputs: the 2-D mesh (vertices only enter through
 * zmin/zmax/ni; xy, triangles with labels, boundary edges with labels and belem/bface = Mesh::BoundaryElement),
 * nlayer (= Nmax), per 2-D vertex ni (0..nlayer), zmin, zmax, and the label maps as (old,new) pairs (the last pair of a
 * label wins, labels without a pair keep their value: BuildLayeMesh_Op fflib/msh3.cpp:4608-4645).
 * The boundary triangles come out as the reference's mesh holds them after BuildAdj (ffo_boundary_links). */
void ffo_buildlayers_sizes(int nv2, int nt2, const int32_t *tri, int nbe2, const int32_t *bedge_elem, const int32_t *bedge_face,
                           int nlayer, const int32_t *ni, int *nv, int *nt, int *nbe);
void ffo_buildlayers(int nv2, const double *xy, int nt2, const int32_t *tri, const int32_t *trilab, int nbe2,
                     const int32_t *bedge_lab, const int32_t *bedge_elem, const int32_t *bedge_face, int nlayer,
                     const int32_t *ni, const double *zmin, const double *zmax, int nreg, const int32_t *regmap, int nmid,
                     const int32_t *midmap, int nup, const int32_t *upmap, int ndown, const int32_t *downmap, double *xyz,
                     int32_t *conn, int32_t *elab, int32_t *bconn, int32_t *blab, int32_t *belem, int32_t *bface);

/* Boundary part of GenericMesh::BuildAdj (femlib/GenericMesh.hpp:914-1017), tetrahedra: (element, face) of every boundary
 * triangle and its final orientation (bconn is modified in place: a true boundary face takes the orientation of its
 * element's face; internal faces pick the element on the side where the face runs the other way, the minority between
 * two regions is turned round). */
void ffo_boundary_links(int nt, const int32_t *conn, const int32_t *elab, int nbe, int32_t *bconn, int32_t *belem, int32_t *bface);

/* 3-D P2 node numbering (vertices and edges in first-encounter order). elem2node: nt*10. Returns nnodes. */
int ffo_p2_nodes_3d(int nv, int nt, const int32_t *conn, int32_t *elem2node);

/* number of nodes per element for (dim, order) */
int ffo_nloc(int dim, int order);

/* Bilinear form -> COO in HashMatrix insertion order (duplicates merged).  elem2node may be NULL for P1.
 * dof(node,c) = node*ncomp + c ; local dof i = c*nloc + a.  labels==NULL: all regions.
 * coo arrays must hold nt*(nloc*ncomp)^2 entries (upper bound).  Returns nnz. */
int64_t ffo_assemble_coo(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                         int order, int ncomp, const int32_t *elem2node,
                         int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                         int nlab, const int32_t *labels,
                         int32_t *coo_i, int32_t *coo_j, double *coo_a);

/* Rectangular matrices `matrix B = vb(Uh,Vh)`: two different spaces on the same mesh, rows = test space (order_v, ncomp_v,
 * e2n_v), columns = space of the unknown (order_u, ncomp_u, e2n_u); Element_Op with Ku != Kv (fflib/problem.cpp:6337-6437).
 * coo arrays must hold nt * (nloc_v*ncomp_v) * (nloc_u*ncomp_u) entries.  Returns nnz. */
int64_t ffo_assemble_coo_rect(int dim, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                              int order_v, int ncomp_v, const int32_t *e2n_v, int order_u, int ncomp_u, const int32_t *e2n_u,
                              int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                              int nlab, const int32_t *labels, int32_t *coo_i, int32_t *coo_j, double *coo_a);

/* COO -> CSR sorted by (i,j) (Sortij/Buildp). */
void ffo_coo_to_csr(int n, int64_t nnz, const int32_t *coo_i, const int32_t *coo_j, const double *coo_a,
                    int32_t *rowptr, int32_t *colind, double *vals);

/* Linear form: b zeroed then filled (OpArraytoLinearForm + Element_rhs). */
void ffo_assemble_rhs(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                      int order, int ncomp, const int32_t *elem2node, int ndof,
                      int nterms, const ffo_lterm *terms, int nq, const double *qpts, const double *qw,
                      int nlab, const int32_t *labels, double *b);
/* linear form with data depending on the mesh point (Element_rhs, problem.cpp:7917-7985): fq[(c*nt + k)*nq + q] = value of
 * the coefficient of v_c at quadrature node q of element k; ADDS to b */
void ffo_assemble_rhs_qvalues(int dim, const double *xyz, int nt, const int32_t *conn, int order, int ncomp,
                              const int32_t *elem2node, int nq, const double *qpts, const double *qw, const double *fq, double *b);
/* ... and with derivatives of the test function: fq[((c*(dim+1) + s)*nt + k)*nq + q], s = 0 value, 1..dim = dx, dy, dz */
void ffo_assemble_rhs_qterms(int dim, const double *xyz, int nt, const int32_t *conn, int order, int ncomp,
                             const int32_t *elem2node, int nq, const double *qpts, const double *qw, const double *fq, double *b);
/* boundary integrals of a linear form (Element_rhs on border elements, problem.cpp:8439-8587); ADDS to b.
 * qpts: nq x (dim-1) reference coordinates on the face / edge */
void ffo_assemble_rhs_boundary(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                               const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                               const int32_t *bface, int nterms, const ffo_lterm *terms, int nq, const double *qpts,
                               const double *qw, int nlab, const int32_t *labels, double *b);

/* ffo_assemble_coo with every term multiplied by a coefficient depending on the mesh point, given at the quadrature
 * nodes: cq[k * nq + q] (Element_Op evaluates the coefficient expression there, problem.cpp:6407) */
int64_t ffo_assemble_coo_qcoef(int dim, int nv, const double *xyz, int nt, const int32_t *conn, const int32_t *elab,
                               int order, int ncomp, const int32_t *elem2node,
                               int nterms, const ffo_bterm *terms, int nq, const double *qpts, const double *qw,
                               const double *cq, int32_t *coo_i, int32_t *coo_j, double *coo_a);

/* boundary integrals of a bilinear form (Robin terms: AssembleBilinearForm border loop problem.cpp:1317-1326, Element_Op
 * border branch :6518-6560 / :6216-6290): COO with one entry per distinct (il, jl) couple of the elements adjacent to the
 * labelled boundary elements, zero or not.  Arrays sized (labelled boundary elements) * (nloc*ncomp)^2.  Returns count. */
int64_t ffo_assemble_coo_boundary(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                  const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                  const int32_t *bface, int nterms, const ffo_bterm *terms, int nq, const double *qpts,
                                  const double *qw, int nlab, const int32_t *labels,
                                  int32_t *coo_i, int32_t *coo_j, double *coo_a);

/* boundary integrals with data depending on the mesh point, given at the face quadrature nodes (problem.cpp:8551-8570,
 * :6526-6556): gq[(c * nbe + ib) * nq + q] for the linear form (adds to b), cq[ib * nq + q] multiplying every term of the
 * bilinear form */
void ffo_assemble_rhs_boundary_qvalues(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                       const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                       const int32_t *bface, int nq, const double *qpts, const double *qw, const double *gq, double *b);
int64_t ffo_assemble_coo_boundary_qcoef(int dim, const double *xyz, const int32_t *conn, int order, int ncomp,
                                        const int32_t *elem2node, int nbe, const int32_t *blab, const int32_t *belem,
                                        const int32_t *bface, int nterms, const ffo_bterm *terms, int nq, const double *qpts,
                                        const double *qw, int nlab, const int32_t *labels, const double *cq,
                                        int32_t *coo_i, int32_t *coo_j, double *coo_a);

/* Dirichlet dofs as AssembleBC visits them: for each boundary element (in order) whose label is in
 * labels[], for each component c with compmask bit c set, each dof lying on that face -> (dof, value[c]).
 * Later pairs overwrite earlier ones.  out arrays sized nbe*ncomp*nlocface at most.  Returns count. */
int ffo_bc_pairs(int dim, int nt, const int32_t *conn, int order, int ncomp, const int32_t *elem2node,
                 int nbe, const int32_t *blab, const int32_t *belem, const int32_t *bface,
                 int nlab, const int32_t *labels, int compmask, const double *values,
                 int32_t *out_dof, double *out_val);

/* matrix: diag(dof) = tgv (overwrite) on COO or CSR values (diag must exist). rhs: b[dof] = tgv*val. */
void ffo_bc_matrix_coo(int64_t nnz, const int32_t *coo_i, const int32_t *coo_j, double *coo_a,
                       int n, int nbc, const int32_t *dofs, double tgv);
void ffo_bc_rhs(double *b, int nbc, const int32_t *dofs, const double *vals, double tgv);

/* y = A x over entries in the order given (HashMatrix::addMatMul on COO storage order). */
void ffo_spmv_coo(int n, int64_t nnz, const int32_t *ai, const int32_t *aj, const double *aa,
                  const double *x, double *y);

/* SolverCG::dosolver: Jacobi preconditioner, tgv rows via gettgv, SetInitWithBC, ConjugueGradient.
 * x holds the initial guess on entry.  itmax<=0 -> n.  Returns ConjugueGradient's code
 * (1 converged, 2 converged at start, 0 not converged); *iters = iterations done. */
int ffo_cg(int n, int64_t nnz, const int32_t *ai, const int32_t *aj, const double *aa,
           const double *b, double *x, double eps, int itmax, double tgv, int *iters, double *gcg_out);

/* SolverGMRES (femlib/VirtualSolverCG.hpp:196-258) = SetInitWithBC + fgmres (femlib/CG.cpp:347-517): right-preconditioned
 * (Jacobi) flexible GMRES(nbkrylov) with modified Gram-Schmidt.  x = initial guess on entry.  Returns 1 when converged. */
int ffo_gmres(int n, int64_t nnz, const int32_t *ai, const int32_t *aj, const double *aa,
              const double *b, double *x, double eps, int itmax, int nbkrylov, double tgv, int *iters, double *relres_out);

#ifdef __cplusplus
}
#endif
#endif
