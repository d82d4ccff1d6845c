/* ffcuda.h — C ABI of libffcuda_core.so, the B200-native (sm_100a) implementation of FreeFEM's
 * finite-element assembly + CG hot path.
 *
 * This is the drop-in boundary (SURVEY.md §8 b, layer B2): plain C, opaque handles, plain pointers and
 * sizes, no C++/torch/FreeFEM types.  It is called by
 *   - freefem-sources_b200/plugin/ffcuda.cpp  (the FreeFEM plugin, `load "ffcuda"`, layer B1), and
 *   - the standalone harness (tests/, bench.py through ctypes).
 * Every function returns 0 on success, non-zero on failure; ffcuda_last_error() gives the message.
 * No exception crosses the boundary.  Host buffers belong to the caller, device memory to the library.
 * One host thread per context.  There is NO CPU fallback: without a CUDA device every compute entry fails.
 *
 * Reference interfaces replaced (paths under FreeFEM 4.15 src/):
 *   ffcuda_mesh_upload          <- reads Mesh/Mesh3 (femlib/fem.hpp, Mesh3dn.hpp, GenericMesh.hpp:315)
 *   ffcuda_mesh_cube / _square  <- BuildCube fflib/msh3.cpp:7879-8132, Carre_ fflib/lgmesh.cpp:1229-1384
 *   ffcuda_space_create         <- FESpace/FESpace3 dof numbering: GFElement::operator() femlib/FESpacen.hpp:444,656,
 *                                  BuildDFNumbering femlib/GenericMesh.hpp:1711-1954
 *   ffcuda_symbolic             <- HashMatrix::operator+=(MatriceElementaire&) femlib/HashMatrix.cpp:1295-1332
 *                                  (entry creation) + Sortij/Buildp/CSR() :671-682,:993-1030,:859-876
 *   ffcuda_assemble_bilinear    <- AssembleBilinearForm fflib/problem.cpp:803-1111 (2-D), :1117-1417 (3-D),
 *                                  Element_Op :6063-6160, :6337-6437, MatriceElementairePleine::call
 *                                  femlib/MatriceCreuse_tpl.hpp:233-258
 *   ffcuda_assemble_linear      <- AssembleLinearForm fflib/problem.cpp:10555, :10878-11227, Element_rhs :7839-7985
 *   ffcuda_assemble_linear_boundary <- Element_rhs on border elements fflib/problem.cpp:8439-8587
 *   ffcuda_assemble_bilinear_boundary <- AssembleBilinearForm border loop fflib/problem.cpp:1317-1326, Element_Op :6518-6560
 *   ffcuda_assemble_linear_qvalues / _qterms <- coefficient evaluation at the quadrature nodes inside Element_rhs
 *                                  fflib/problem.cpp:7876-7884, :7951-7975 (data depending on the mesh point)
 *   ffcuda_assemble_bilinear_qcoef <- the same inside Element_Op fflib/problem.cpp:6380-6407
 *   ffcuda_assemble_linear_boundary_qvalues / ffcuda_assemble_bilinear_boundary_qcoef <- the same on border elements
 *                                  fflib/problem.cpp:8551-8570, :6526-6556
 *   ffcuda_assemble_bilinear_rect <- `matrix B = vb(Uh,Vh)` with two different spaces: the operator takes both
 *                                  fflib/problem.hpp:1628-1631, Element_Op with Ku != Kv fflib/problem.cpp:6337-6437, :6063-6160
 *   ffcuda_fe_table             <- an FE function used as data of a form, evaluated from its dof array instead of through
 *                                  the interpreter: pfer2R fflib/lgfem.cpp:2053-2088 -> FElement::operator()(PHat,u,comp,op)
 *                                  femlib/FESpace.cpp:1078-1099, :1637-1654, femlib/P012_3d.cpp:98-122
 *   ffcuda_bc_* / *_apply_bc    <- AssembleBC fflib/problem.cpp:9881-10034, :10039-10194, HashMatrix::SetBC
 *                                  femlib/HashMatrix.cpp:1195-1238 (penalty and exact elimination)
 *   ffcuda_gmres                <- SolverGMRES femlib/VirtualSolverCG.hpp:196-258, fgmres femlib/CG.cpp:347-517
 *   ffcuda_quadrature           <- CDomainOfIntegration::FIT/FIV fflib/problem.cpp:14102-14145, QF_Simplex
 *                                  femlib/QuadratureFormular.cpp:73-115 and the rule tables :138-188, :699-743
 *   ffcuda_spmv                 <- HashMatrix::addMatMul femlib/HashMatrix.cpp:1087-1154
 *   ffcuda_cg                   <- SolverCG::dosolver femlib/VirtualSolverCG.hpp:112-192, HMatVirtPrecon :13-111,
 *                                  ConjugueGradient femlib/CG.cpp:195-265, gettgv HashMatrix.cpp:1341-1371
 *   ffcuda_matrix_export_device / _download_coo / _write_morse <- hand-off formats: PETSc through host arrays
 *                                  plugin/mpi/PETSc-code.hpp, `[I,J,C]=A` fflib/lgmat.cpp, `ofstream << A`
 *                                  femlib/HashMatrix.hpp:485-508 (reader femlib/HashMatrix.cpp:137-188)
 *   ffcuda_mesh_adjacency       <- GenericMesh::BuildAdj femlib/GenericMesh.hpp:837-930
 *   ffcuda_mesh_buildlayers     <- build_layer fflib/msh3.cpp:895-1757 (operator BuildLayeMesh_Op :4536-4760) and the
 *                                  boundary part of BuildAdj femlib/GenericMesh.hpp:914-1017
 *   ffcuda_partition_rcb / _local, ffcuda_mesh_upload_distributed, ffcuda_mesh_cube_distributed, ffcuda_comm_*
 *                               <- element-range split fflib/problem.cpp:1133-1138, partition vector plugin/seq/metis.cpp,
 *                                  MPI_Allreduce per dot product plugin/mpi/MPICG.cpp:93-101
 */
#ifndef FFCUDA_H
#define FFCUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ffcuda_ctx ffcuda_ctx;
typedef struct ffcuda_mesh ffcuda_mesh;
typedef struct ffcuda_space ffcuda_space;
typedef struct ffcuda_pattern ffcuda_pattern;
typedef struct ffcuda_matrix ffcuda_matrix;
typedef struct ffcuda_vec ffcuda_vec;
typedef struct ffcuda_bc ffcuda_bc;

/* differential operator codes = FreeFEM's (femlib/FESpacen.hpp:73-82) */
enum { FFCUDA_OP_ID = 0, FFCUDA_OP_DX = 1, FFCUDA_OP_DY = 2, FFCUDA_OP_DZ = 6 };

/* one term of a bilinear form: coef * d^uop(u_ucomp) * d^vop(v_vcomp); unknown -> column, test -> row
 * (BilinearOperator = LinearComb<pair<MGauche,MDroit>,C_F0>, femlib/DOperator.hpp:279-319) */
typedef struct { int32_t ucomp, uop, vcomp, vop; double coef; } ffcuda_bterm;
/* one term of a linear form: coef * d^vop(v_vcomp) */
typedef struct { int32_t vcomp, vop; double coef; } ffcuda_lterm;

/* ---- context ------------------------------------------------------------------------------------ */
int ffcuda_ctx_create(int device, ffcuda_ctx **out);
void ffcuda_ctx_destroy(ffcuda_ctx *ctx);
const char *ffcuda_last_error(ffcuda_ctx *ctx); /* ctx may be NULL: last error of the calling thread */
int ffcuda_ctx_sync(ffcuda_ctx *ctx);
/* run all library work on an externally owned cudaStream_t (e.g. torch's current stream); NULL restores
 * the library-owned stream */
int ffcuda_ctx_set_stream(ffcuda_ctx *ctx, void *cuda_stream);
void *ffcuda_ctx_get_stream(ffcuda_ctx *ctx);
/* tuning knobs (never change results beyond round-off; FreeFEM has no counterpart).  "tile_policy": 0 = scalar P1
 * forms are assembled by the thread-per-row kernel, 1 (default) = by row tiles from the second assembly on the same
 * fespace (the tile set is built once per fespace), 2 = always by row tiles.  "tile_rows": rows per tile (8..256,
 * default 96); takes effect for fespaces whose tile set is not built yet. */
int ffcuda_ctx_set_option(ffcuda_ctx *ctx, const char *name, int value);
/* built-in kernel profiler: when enabled every kernel launch is bracketed by CUDA events on the launch
 * stream.  ffcuda_prof_get returns accumulated milliseconds and launch count of kernels whose name starts
 * with `prefix` ("" = all). */
int ffcuda_prof_enable(ffcuda_ctx *ctx, int on);
int ffcuda_prof_reset(ffcuda_ctx *ctx);
int ffcuda_prof_get(ffcuda_ctx *ctx, const char *prefix, double *ms, int64_t *launches);
/* total number of kernels launched by this context since creation / last reset (always counted) */
int64_t ffcuda_launch_count(ffcuda_ctx *ctx);

/* ---- mesh ---------------------------------------------------------------------------------------- */
/* dim 2: triangles, dim 3: tetrahedra.  xyz: nv*dim (x,y[,z] per vertex).  conn: nt*(dim+1) vertex ids.
 * elab: nt region labels (NULL = 0).  Boundary elements: bconn nbe*dim vertex ids, blab labels,
 * belem/bface = Th.BoundaryElement(ib, ie) (element and local face; face ie is opposite local vertex ie).
 * belem/bface may be NULL (then they are recovered by matching faces on the device). */
int ffcuda_mesh_upload(ffcuda_ctx *ctx, int dim, int nv, const double *xyz, int nt, const int32_t *conn,
                       const int32_t *elab, int nbe, const int32_t *bconn, const int32_t *blab,
                       const int32_t *belem, const int32_t *bface, ffcuda_mesh **out);
/* structured meshes generated ON THE DEVICE with FreeFEM's vertex/element/boundary ordering and labels */
int ffcuda_mesh_cube(ffcuda_ctx *ctx, int nx, int ny, int nz, ffcuda_mesh **out);
int ffcuda_mesh_square(ffcuda_ctx *ctx, int nx, int ny, ffcuda_mesh **out);
int ffcuda_mesh_info(ffcuda_mesh *m, int *dim, int *nv, int *nt, int *nbe);
/* any output pointer may be NULL */
int ffcuda_mesh_download(ffcuda_mesh *m, double *xyz, int32_t *conn, int32_t *elab, int32_t *bconn,
                         int32_t *blab, int32_t *belem, int32_t *bface);
/* element adjacency, GenericMesh::BuildAdj (femlib/GenericMesh.hpp:837-930): adj[(dim+1)*k + i] = (dim+1)*k' + i' when face i
 * of element k (opposite its vertex i) is face i' of element k'; -1 on the boundary; -2 for a face shared by more than two
 * elements.  Built on the device on first use (face hashes, radix sort, match) and kept with the mesh; `adj` (host) and
 * `d_adj` (borrowed device pointer) may be NULL. */
int ffcuda_mesh_adjacency(ffcuda_mesh *m, int32_t *adj, const int32_t **d_adj);
/* buildlayers(Th2, nlayer, zbound=[zmin,zmax], coef=..., region=, labelmid=, labelup=, labeldown=) on the device
 * (fflib/msh3.cpp:895-1757; `mesh3 Th = buildlayers(...)` of idp/Heat3d.idp:16): the layered tetrahedral mesh over the
 * 2-D mesh m2 (uploaded, or ffcuda_mesh_square), same vertex / element / boundary-element order, labels and boundary
 * orientation as FreeFEM's.  Host inputs per 2-D vertex, as the operator derives them (:4566-4657): ni[] layers over the
 * vertex (NULL: nlayer everywhere), zmin[], zmax[] (NULL: 0 and 1).  Label maps are (old,new) pairs, n* = number of pairs:
 * regmap for the tetrahedra (from the triangle labels), midmap for the lateral faces (from the boundary-edge labels), upmap /
 * downmap for the faces at zmax / zmin (from the triangle labels); labels without a pair are kept. */
int ffcuda_mesh_buildlayers(ffcuda_mesh *m2, int nlayer, const int32_t *ni, const double *zmin, const double *zmax,
                            int nreg, const int32_t *regmap, int nmid, const int32_t *midmap, int nup, const int32_t *upmap,
                            int ndown, const int32_t *downmap, ffcuda_mesh **out);
void ffcuda_mesh_destroy(ffcuda_mesh *m);

/* ---- finite-element space ------------------------------------------------------------------------ */
/* order 1|2 Lagrange, ncomp identical components ([P2,P2,P2] -> order 2, ncomp 3).
 * Numbering contract (FESpacen.cpp:196-235, probed): dof(node,c) = node*ncomp + c; local dof i of an
 * element = c*nloc + a, nodes a = vertices then edges ({01,02,03,12,13,23} in 3-D; edge opposite vertex
 * a-3 in 2-D).  elem2node: nt*nloc node ids as FreeFEM numbered them (Vh(k,a)/ncomp); NULL means
 * "number them for me": P1 -> vertex ids; P2 3-D -> first-encounter order of BuildDFNumbering; P2 2-D is
 * rejected (FreeFEM applies its Gibbs renumbering there, FESpace.cpp:991 — pass the table). */
int ffcuda_space_create(ffcuda_mesh *m, int order, int ncomp, const int32_t *elem2node, int nnodes,
                        ffcuda_space **out);
int ffcuda_space_info(ffcuda_space *s, int *ndof, int *ndofK, int *nnodes);
int ffcuda_space_download_dofs(ffcuda_space *s, int32_t *dof /* nt*ndofK, FreeFEM's Vh(k,i) */);
void ffcuda_space_destroy(ffcuda_space *s);

/* ---- symbolic sparsity (kernel 1) ----------------------------------------------------------------- */
/* CSR pattern of the (Vh,Vh) matrix exactly as MatriceMorse would hold it after CSR(): one entry per couple
 * of dofs sharing an element (structural zeros and uncoupled components included), rows and columns sorted
 * ascending, int32, 0-based.  Also builds the node->element incidence used by the row-owner assembly.
 * Lifetimes: the space must be alive whenever a matrix of this pattern is ASSEMBLED (the assembly entries take both).  A
 * matrix and its pattern on an ordinary (non-distributed) mesh may be solved with, multiplied and downloaded after the space
 * and its mesh are destroyed; on a distributed mesh the space and the mesh must outlive them (solves read the halo lists). */
int ffcuda_symbolic(ffcuda_space *s, ffcuda_pattern **out);
int ffcuda_pattern_info(ffcuda_pattern *p, int *n, int64_t *nnz);
int ffcuda_pattern_download(ffcuda_pattern *p, int32_t *rowptr /* n+1 */, int32_t *colind /* nnz */);
/* same copies on a second stream, behind the work enqueued so far: the call returns at once and the numeric assembly
 * that follows overlaps the transfer.  The buffers (pinned host memory, or the copy degrades to a synchronous one) are
 * valid after ffcuda_ctx_sync. */
int ffcuda_pattern_download_async(ffcuda_pattern *p, int32_t *rowptr, int32_t *colind);
/* half storage (`sym=1`): FreeFEM's MatriceMorse then keeps the entries (i, j <= i) only (MatriceElementaireSymetrique,
 * HashMatrix::operator+= femlib/HashMatrix.cpp:1319-1325; addMatMul :1087-1154 mirrors them).  The device matrix stays
 * full; these entries cut the lower triangle out for the host (rowptr n+1, colind / vals ffcuda_pattern_lower_nnz long),
 * which equals FreeFEM's half matrix for symmetric forms. */
int ffcuda_pattern_lower_nnz(ffcuda_pattern *p, int64_t *nnz_lower);
int ffcuda_pattern_download_lower(ffcuda_pattern *p, int32_t *rowptr, int32_t *colind);
void ffcuda_pattern_destroy(ffcuda_pattern *p);

/* ---- matrices and vectors (device resident) ------------------------------------------------------- */
int ffcuda_matrix_create(ffcuda_pattern *p, ffcuda_matrix **out); /* values zeroed */
/* a matrix given by host CSR arrays (solver-only use: `set(A,solver=CG)` on an existing MatriceMorse) */
int ffcuda_matrix_from_csr(ffcuda_ctx *ctx, int n, int64_t nnz, const int32_t *rowptr, const int32_t *colind,
                           const double *vals, ffcuda_matrix **out);
/* the same from a half-stored matrix (sorted rows, entries j <= i): expanded to the full symmetric matrix */
int ffcuda_matrix_from_csr_lower(ffcuda_ctx *ctx, int n, int64_t nnz_lower, const int32_t *rowptr, const int32_t *colind,
                                 const double *vals, ffcuda_matrix **out);
int ffcuda_matrix_info(ffcuda_matrix *A, int *n, int64_t *nnz);
int ffcuda_matrix_shape(ffcuda_matrix *A, int *n, int *ncols, int64_t *nnz); /* rows, columns (rectangular / distributed matrices), nnz */
/* CSR arrays of a matrix that has no pattern object (ffcuda_matrix_from_csr*, ffcuda_assemble_bilinear_rect) -> host;
 * vals may be NULL */
int ffcuda_matrix_download_csr(ffcuda_matrix *A, int32_t *rowptr /* n+1 */, int32_t *colind /* nnz */, double *vals /* nnz */);
int ffcuda_matrix_download(ffcuda_matrix *A, double *vals /* nnz, CSR order */);
int ffcuda_matrix_download_lower(ffcuda_matrix *A, double *vals); /* values of the lower triangle, see ffcuda_pattern_lower_nnz */
int ffcuda_matrix_upload(ffcuda_matrix *A, const double *vals);
/* ---- hand-off formats straight from the device CSR ----
 * ffcuda_matrix_export_device: BORROWED device pointers to the CSR triple (int32 rowptr[n+1], int32 colind[nnz], fp64
 *   vals[nnz]; valid until the matrix is destroyed or re-assembled): what a consumer that stays on the GPU takes, e.g.
 *   PETSc's MatCreateSeqAIJCUSPARSE / MatSeqAIJCUSPARSESetPreallocationCSR-style constructors (the reference reaches
 *   PETSc through host arrays, plugin/mpi/PETSc-code.hpp) - no host round trip.
 * ffcuda_matrix_download_coo: the triple of `[I,J,C] = A` (fflib/lgmat.cpp), row indices expanded on the device, CSR order,
 *   index_base 0 or 1.
 * ffcuda_matrix_write_morse: FreeFEM's Morse text format, the output of `ofstream f; f << A` after `A.CSR`
 *   (femlib/HashMatrix.hpp:485-508: header `n m half  nnz`, then `i j a_ij` 1-based, values with 20 digits), readable by
 *   HashMatrix(istream&) (femlib/HashMatrix.cpp:137-188); half != 0 writes the entries (i, j <= i) only. */
int ffcuda_matrix_export_device(ffcuda_matrix *A, const int32_t **d_rowptr, const int32_t **d_colind, const double **d_vals,
                                int *n, int64_t *nnz);
int ffcuda_matrix_download_coo(ffcuda_matrix *A, int32_t *I, int32_t *J, double *C, int index_base);
int ffcuda_matrix_write_morse(ffcuda_matrix *A, const char *path, int half);
void ffcuda_matrix_destroy(ffcuda_matrix *A);

int ffcuda_vec_create(ffcuda_ctx *ctx, int n, ffcuda_vec **out); /* zeroed */
int ffcuda_vec_upload(ffcuda_vec *v, const double *host);
int ffcuda_vec_download(ffcuda_vec *v, double *host);
int ffcuda_vec_fill(ffcuda_vec *v, double value);
void *ffcuda_vec_ptr(ffcuda_vec *v); /* device pointer (double*) */
void ffcuda_vec_destroy(ffcuda_vec *v);

/* ---- default quadrature ------------------------------------------------------------------------------ */
/* The rule FreeFEM uses for int2d/int3d(Th, qforder=q) (default q = 6): fewest points exact for degree q-1.
 * qpts: nq*dim reference coordinates, qw: weights summing to 1; either may be NULL to query *nq (<= 16).
 * Host-only helper (no device needed); the FreeFEM plugin forwards FreeFEM's own tables instead. */
int ffcuda_quadrature(int dim, int qforder, int *nq, double *qpts, double *qw);

/* ---- numeric assembly (kernels 2+3 fused: row-owner gather with in-register element evaluation) ---- */
/* A (+)= sum over elements whose region label is in labels[] (NULL = all) of the local matrices of the
 * form.  Quadrature: nq points, qpts nq*dim reference coordinates, qw weights summing to 1 (the rule
 * FreeFEM selected through qforder/qft/qfV; default 7-point (2-D) / 14-point (3-D)).  Coefficients are
 * constants (MeshIndependent() terms).  accumulate=0 overwrites A. */
int ffcuda_assemble_bilinear(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms,
                             int nq, const double *qpts, const double *qw,
                             int nlab, const int32_t *labels, int accumulate);
/* A (+)= the same with every term multiplied by ONE coefficient that depends on the mesh point (kappa(x,y,z), a P0 / P1
 * FE function, ...), given by the values Element_Op would compute (fflib/problem.cpp:6407): cq[k * nq + q] (HOST array, or a DEVICE pointer: ffcuda_fe_table) at
 * quadrature node q of element k.  P1: the gradients do not depend on the node, so the element integrals reduce exactly to
 * the moments sum_q w_q c_q, sum_q w_q c_q lambda_a, sum_q w_q c_q lambda_a lambda_b of the coefficient, formed on the
 * device.  P2: the per-pair tensors sum_q c_q w_q d phi_a(q) d phi_b(q) are formed on the fly from the node values.
 * Forms with several coefficient functions: one call per function. */
int ffcuda_assemble_bilinear_qcoef(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms,
                                   int nq, const double *qpts, const double *qw, const double *cq, int accumulate);
/* RECTANGULAR matrices, `matrix B = vb(Uh,Vh)` with two different spaces on the SAME device mesh (blocks of a Stokes / mixed
 * problem assembled one by one, projection matrices between P1 and P2): *out = a new matrix with the dofs of the test
 * space sv as rows and the dofs of the space of the unknown su as columns (Element_Op with Ku != Kv,
 * fflib/problem.cpp:6337-6437), every couple of the element matrices in the pattern (HashMatrix::operator+=), sorted CSR.
 * Terms: ucomp < ncomp(su), vcomp < ncomp(sv), constant coefficients; labels: regions that contribute (NULL = all; the
 * pattern keeps every couple).  Each space P1 / P2 with 1..3 components.  The result has no pattern object: ffcuda_spmv
 * (x of size columns, y of size rows), ffcuda_matrix_shape / _download_csr / _download_coo / _export_device /
 * _write_morse apply; the solvers and ffcuda_matrix_apply_bc refuse it.  General kernel (one thread per row node,
 * quadrature loop per couple): not the square hot path. */
int ffcuda_assemble_bilinear_rect(ffcuda_space *sv, ffcuda_space *su, int nterms, const ffcuda_bterm *terms,
                                  int nq, const double *qpts, const double *qw,
                                  int nlab, const int32_t *labels, ffcuda_matrix **out);
int ffcuda_assemble_linear(ffcuda_vec *b, ffcuda_space *s, int nterms, const ffcuda_lterm *terms,
                           int nq, const double *qpts, const double *qw,
                           int nlab, const int32_t *labels, int accumulate);
/* b (+)= boundary integrals int2d(Th3, labels)(c v) / int1d(Th, labels)(c v) of a linear form (Neumann, traction data):
 * Element_rhs on border elements, fflib/problem.cpp:8439-8513 (2-D), :8517-8587 (3-D).  Value terms only (vop = id),
 * constant c.  qpts: nq x (dim-1) reference coordinates of the face rule FreeFEM selected (default: 7 points on a face,
 * 3 Gauss points on an edge), qw: weights summing to 1; labels NULL = every boundary element.  The mesh must have been
 * uploaded with its boundary elements (belem / bface given or recovered). */
int ffcuda_assemble_linear_boundary(ffcuda_vec *b, ffcuda_space *s, int nterms, const ffcuda_lterm *terms,
                                    int nq, const double *qpts, const double *qw,
                                    int nlab, const int32_t *labels, int accumulate);
/* b (+)= volume integral of a linear form whose data depend on the mesh point (f(x,y,z) v, uold v / dt, ...): Element_rhs
 * (fflib/problem.cpp:7839-7985) evaluates the coefficient at every quadrature node of every element, and the caller hands
 * those very values over: fq[(c * nt + k) * nq + q] (HOST array or DEVICE pointer) = coefficient of the value of v_c at node q of element k,
 * summed over the terms, 0 where the element is outside the integral's region.  Value terms only. */
int ffcuda_assemble_linear_qvalues(ffcuda_vec *b, ffcuda_space *s, int nq, const double *qpts, const double *qw,
                                   const double *fq, int accumulate);
/* the same with derivatives of the test function (the residual of a Newton step, int(dx(uk) dx(v) + ...)):
 * fq[((c * (dim+1) + s) * nt + k) * nq + q] (HOST or DEVICE) = coefficient of d^s v_c, s = 0 value, 1..dim = dx, dy, dz */
int ffcuda_assemble_linear_qterms(ffcuda_vec *b, ffcuda_space *s, int nq, const double *qpts, const double *qw,
                                  const double *fq, int accumulate);
/* A (+)= boundary integrals int2d(Th3, labels)(c u v) / int1d(Th, labels)(c u v) of a bilinear form (Robin terms):
 * the border loop of AssembleBilinearForm, fflib/problem.cpp:1317-1326 (3-D), :1030-1040 (2-D), with Element_Op's border
 * branch :6518-6560 / :6216-6290.  Value terms only (uop = vop = id), constant c.  The couples FreeFEM creates for a border
 * element are those of the adjacent element, all present in the pattern of the space.  qpts / qw / labels as above. */
int ffcuda_assemble_bilinear_boundary(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms,
                                      int nq, const double *qpts, const double *qw,
                                      int nlab, const int32_t *labels, int accumulate);
/* the two boundary integrals with data that depend on the mesh point, given by the values FreeFEM's evaluator returns at
 * the face quadrature nodes (fflib/problem.cpp:8551-8570, :6526-6556): gq[(c * nbe + ib) * nq + q] (HOST or DEVICE) = coefficient of
 * the value of v_c at node q of boundary element ib, 0 where the integral does not go; cq[ib * nq + q] (HOST or DEVICE) = the ONE
 * coefficient function multiplying every listed term (labels as above). */
int ffcuda_assemble_linear_boundary_qvalues(ffcuda_vec *b, ffcuda_space *s, int nq, const double *qpts, const double *qw,
                                            const double *gq, int accumulate);
int ffcuda_assemble_bilinear_boundary_qcoef(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms,
                                            int nq, const double *qpts, const double *qw, const double *cq,
                                            int nlab, const int32_t *labels, int accumulate);

/* An FE function handed over as its DOF ARRAY (f in int3d(Th)(f v), kappa in int3d(Th)(kappa grad u . grad v), uk in the
 * residual int3d(Th)(dx(uk) dx(v) + ...) of a Newton step): its values (op = FFCUDA_OP_ID) or derivatives (DX, DY, DZ) at
 * the quadrature nodes, formed on the device instead of one interpreter call per node (pfer2R, fflib/lgfem.cpp:2053-2088 ->
 * FElement::operator()(PHat,u,comp,op), femlib/FESpace.cpp:1078-1099, :1637-1654, femlib/P012_3d.cpp:98-122):
 *   table[offset + u * nq + q] (+)= scale * d^op f(P_q of unit u),   0 where the label of unit u is not listed
 * border = 0: units = elements, qpts nq x dim reference coordinates, labels = region labels;  border = 1: units = boundary
 * elements, qpts nq x (dim-1) coordinates of the face rule, P_q = PBord(face, q) in the adjacent element, labels = boundary
 * labels (NULL: every unit).  f lives on mesh m: order 0 (P0), 1 or 2 (P1 / P2 Lagrange); e2n = its element -> node table
 * (HOST, nt x nloc, required for P2; NULL: node = element for P0, node = vertex for P1, as FreeFEM numbers them); the dof
 * of node i is dofs[i * dstride + doff] (component c of a [P1,P1,P1] function: dstride 3, doff c).  dofs and table are
 * device vectors; the table has the layout of cq / fq / gq of the entries above, which all accept a DEVICE pointer
 * (ffcuda_vec_ptr(table)) in place of the host array and then use it where it lies. */
int ffcuda_fe_table(ffcuda_mesh *m, int order, const int32_t *e2n, int dstride, int doff, ffcuda_vec *dofs, int op,
                    int border, int nq, const double *qpts, double scale, int nlab, const int32_t *labels,
                    ffcuda_vec *table, int64_t offset, int accumulate);

/* ---- Dirichlet conditions ------------------------------------------------------------------------------
 * tgv >= 0: penalty, A(d,d) = tgv and b[d] = tgv*g(d).  tgv < 0: exact elimination exactly as HashMatrix::SetBC
 * (femlib/HashMatrix.cpp:1195-1238) and AssembleBC (fflib/problem.cpp:10099,10176): rows of the Dirichlet dofs zeroed
 * with A(d,d) = 1 (tgv = -1), rows and columns (-2), columns only (-3), the -10/-20/-30 variants with A(d,d) = 0;
 * b[d] = g(d).  When several ffcuda_bc are applied with tgv = -2/-3 the result equals SetBC on their union as long as
 * every one is applied to the matrix before the right-hand side is used (column zeroing is idempotent). */
/* explicit (dof, value) pairs as AssembleBC produced them on the host (later pairs win) */
int ffcuda_bc_from_pairs(ffcuda_space *s, int n, const int32_t *dofs, const double *vals, ffcuda_bc **out);
/* on(labels..., u_c = values[c]) for the components in compmask, evaluated on the device */
int ffcuda_bc_from_labels(ffcuda_space *s, int nlab, const int32_t *labels, int compmask, const double *values,
                          ffcuda_bc **out);
/* several ffcuda_bc may be applied in sequence (one per on(...) item of the varf) */
int ffcuda_bc_count(ffcuda_bc *bc, int *ndofs);
int ffcuda_matrix_apply_bc(ffcuda_matrix *A, ffcuda_bc *bc, double tgv); /* A(d,d) = tgv | exact elimination */
int ffcuda_vec_apply_bc(ffcuda_vec *b, ffcuda_bc *bc, double tgv);       /* b[d] = tgv*g(d) | g(d)           */
int ffcuda_vec_set_bc_values(ffcuda_vec *x, ffcuda_bc *bc);              /* x[d] = g(d)        */
void ffcuda_bc_destroy(ffcuda_bc *bc);

/* ---- SpMV and CG (kernel 4) ------------------------------------------------------------------------ */
int ffcuda_spmv(ffcuda_matrix *A, ffcuda_vec *x, ffcuda_vec *y); /* y = A x */
/* Jacobi-preconditioned CG with FreeFEM's recurrence, tgv-row handling and stopping rule
 * (gCg < eps^2 * gCg0 for eps > 0, absolute gCg < eps^2 for eps < 0).  x: initial guess in, solution out.
 * itmax <= 0 -> n.  Returns 0 also when not converged; *converged = 1 converged, 2 converged before the
 * first iteration, 0 itmax reached.  *gcg = final <g,Cg>. */
int ffcuda_cg(ffcuda_matrix *A, ffcuda_vec *b, ffcuda_vec *x, double eps, int itmax, double tgv,
              int *iters, int *converged, double *gcg);
/* same with host vectors (the call the FreeFEM solver plugin makes) */
int ffcuda_cg_host(ffcuda_matrix *A, const double *b, double *x, double eps, int itmax, double tgv,
                   int *iters, int *converged, double *gcg);

/* the ABSOLUTE threshold on <g,Cg> the last CG solve of A's context stopped on: eps^2 * <g0,Cg0> for eps > 0 (ConjugueGradient
 * rewrites its eps to the square root of this, femlib/CG.cpp:226, and SolverCG hands it back through `veps=`,
 * femlib/VirtualSolverCG.hpp:186), eps^2 for eps < 0 */
int ffcuda_cg_stop_threshold(ffcuda_matrix *A, double *eps2);

/* GMRES for non-symmetric matrices: SolverGMRES (femlib/VirtualSolverCG.hpp:196-258) = SetInitWithBC + fgmres
 * (femlib/CG.cpp:347-517), flexible GMRES(restart) with the Jacobi preconditioner on the right, modified Gram-Schmidt;
 * stops when |g[it+1]| / ||b (tgv rows zeroed)|| < |eps| (eps < 0: absolute).  restart <= 0: FreeFEM's default 1000
 * (dimKrylov=), itmax <= 0: n.  x = initial guess on entry.  *iters = fgmres's iteration counter, *converged = 0/1,
 * *relres = the last relative residual.  One GPU. */
int ffcuda_gmres(ffcuda_matrix *A, ffcuda_vec *b, ffcuda_vec *x, double eps, int itmax, int restart, double tgv,
                 int *iters, int *converged, double *relres);
int ffcuda_gmres_host(ffcuda_matrix *A, const double *b, double *x, double eps, int itmax, int restart, double tgv,
                      int *iters, int *converged, double *relres);

/* ---- multi-GPU (one process per GPU; the caller's launcher provides rank/size and moves the 128-byte
 *      NCCL id between ranks, e.g. with torch.distributed) ------------------------------------------- */
int ffcuda_comm_unique_id(void *id128);
int ffcuda_comm_init(ffcuda_ctx *ctx, int rank, int nranks, const void *id128);
int ffcuda_comm_finalize(ffcuda_ctx *ctx);
/* slab partition of cube(nx,ny,nz) along z into nranks parts: builds this rank's local mesh (owned
 * elements + one layer of halo elements), local numbering (owned vertices first, then ghosts) and the
 * halo exchange lists.  The space/pattern/matrix/vector calls above then work on the local problem:
 * rows = owned dofs, columns = owned + ghost dofs; ffcuda_spmv and ffcuda_cg exchange ghosts and
 * all-reduce dot products over NCCL. */
int ffcuda_mesh_cube_distributed(ffcuda_ctx *ctx, int nx, int ny, int nz, ffcuda_mesh **out);
/* the partition arithmetic alone (host only, no device needed): out16 = { first owned vertex layer, owned layers,
 * first local cell layer, local cell layers, owned vertices, local vertices (owned + ghost), local tets,
 * lower neighbour rank (-1 none), upper neighbour rank, send offset down, send offset up, recv offset from below,
 * recv offset from above, vertices per exchanged layer, has lower neighbour, has upper neighbour } */
int ffcuda_partition_cube(int nx, int ny, int nz, int rank, int nranks, int64_t *out16);
/* General (unstructured) meshes, host only, no device needed - the arithmetic a distributed upload is built on:
 * recursive coordinate bisection of the vertices into nparts parts of equal size (+-1 per split; deterministic: ties by
 * vertex id), and the local problem of `rank` for any vertex partition: it owns the vertices with part[v] == rank (= its
 * matrix rows), holds every element touching one of them (rows assemble without communication), the other vertices of
 * those elements are its ghosts.  Local numbering: owned vertices (ascending global id), then ghosts grouped by owner
 * rank, ascending id inside - a neighbour's data arrives as one contiguous range, only the sender gathers.
 * sizes8 = { owned, ghosts, local elements, neighbours, total send count, 0, 0, 0 }; call once with the arrays NULL for
 * the sizes, then with l2g[owned+ghosts], elems[local elements], nbr/recv_off/recv_cnt[neighbours],
 * send_ptr[neighbours+1], send_idx[total send count] (owned LOCAL indices, in the receiver's ghost order).
 * (FreeFEM's own counterpart splits the element range and all-reduces the whole matrix, fflib/problem.cpp:1133-1138.) */
int ffcuda_partition_rcb(int dim, int nv, const double *xyz, int nparts, int32_t *part);
int ffcuda_partition_local(int dim, int nv, int nt, const int32_t *conn, const int32_t *part, int rank, int nranks,
                           int64_t *sizes8, int32_t *l2g, int32_t *elems, int32_t *nbr, int32_t *recv_off, int32_t *recv_cnt,
                           int32_t *send_ptr, int32_t *send_idx);
/* The same for any element -> node table (nloc nodes per element) and node partition - P2 spaces on a distributed mesh:
 * elem2node = the GLOBAL table (FreeFEM's numbering), part[node] = rank of the vertex, or of one end point of the edge (then
 * the local elements are those of the vertex partition, in the same order, and every owned row assembles without
 * communication).  Same outputs; l2g = local -> global NODE. */
int ffcuda_partition_local_nodes(int nloc, int nnodes, int nt, const int32_t *elem2node, const int32_t *part, int rank, int nranks,
                                 int64_t *sizes8, int32_t *l2g, int32_t *elems, int32_t *nbr, int32_t *recv_off, int32_t *recv_cnt,
                                 int32_t *send_ptr, int32_t *send_idx);
/* A space on a distributed mesh with its own node table and node-level halo lists (P2, scalar or vector): elem2node =
 * nt_local x nloc LOCAL node ids (owned nodes first, ghosts grouped by owner rank), lists as ffcuda_partition_local_nodes
 * returns them.  Rows = owned nodes, columns = local nodes; ffcuda_spmv / ffcuda_cg / ffcuda_gmres exchange the ghost
 * nodes. */
/* A host CSR matrix (n x n, sorted rows) shared out by contiguous row blocks - rank r owns the rows [n r / nranks, n (r+1) / nranks):
 * the local problem of `rank` (host arithmetic only).  sizes8 = { owned rows, ghosts, local nnz, neighbours, total send count,
 * first owned row, 0, 0 }; call once with the arrays NULL for the sizes, then with l2g[owned + ghosts] (local -> global dof),
 * lrowptr[owned + 1], lcolind[local nnz] (local numbering; the values are the slice vals[rowptr[first] ...] of the caller's
 * array), nbr / recv_off / recv_cnt[neighbours], send_ptr[neighbours + 1], send_idx[total send count]. */
int ffcuda_partition_rows_local(int n, const int32_t *rowptr, const int32_t *colind, int rank, int nranks, int64_t *sizes8,
                                int32_t *l2g, int32_t *lrowptr, int32_t *lcolind, int32_t *nbr, int32_t *recv_off, int32_t *recv_cnt,
                                int32_t *send_ptr, int32_t *send_idx);
/* The rows of one rank of a matrix shared out by rows, from host CSR arrays in LOCAL numbering (owned dofs first: column i is
 * row i; then the ghost dofs grouped by owner rank) with its halo lists in the form ffcuda_partition_local returns them
 * (an empty range in one direction is allowed - non-symmetric structure - but both ranks must list each other).  What the
 * FreeFEM plugin uses to solve on several GPUs a matrix that lives on the host (FFCUDA_NGPU). */
int ffcuda_matrix_from_csr_distributed(ffcuda_ctx *ctx, int n_owned, int ncols, int64_t nnz, const int32_t *rowptr,
                                       const int32_t *colind, const double *vals, int nnbr, const int32_t *nbr,
                                       const int32_t *recv_off, const int32_t *recv_cnt, const int32_t *send_ptr,
                                       const int32_t *send_idx, ffcuda_matrix **out);
int ffcuda_space_create_distributed(ffcuda_mesh *m, int order, int ncomp, const int32_t *elem2node, int nnodes_owned,
                                    int nnodes_local, int nnbr, const int32_t *nbr, const int32_t *recv_off, const int32_t *recv_cnt,
                                    const int32_t *send_ptr, const int32_t *send_idx, ffcuda_space **out);
/* The local problem of this rank for ANY vertex partition (the arrays ffcuda_partition_local returns, local numbering: owned
 * vertices first, then the ghosts grouped by owner rank): xyz[nv_local*dim], conn with LOCAL vertex ids, the boundary
 * elements whose element is local, gid[nv_local] global ids, and the halo description - per neighbour x the contiguous ghost
 * range [recv_off[x], +recv_cnt[x]) and the gather list send_idx[send_ptr[x] .. send_ptr[x+1]) of owned vertices it needs, in
 * the order of ITS ghost range.  Any number of neighbours up to 16; the halo exchange packs and stores into the peers'
 * mailboxes in one kernel (NCCL send/recv after a pack kernel as the fallback).  P1 spaces (scalar or vector). */
int ffcuda_mesh_upload_distributed(ffcuda_ctx *ctx, int dim, int nv_owned, int nv_local, const double *xyz, int nt,
                                   const int32_t *conn, const int32_t *elab, int nbe, const int32_t *bconn, const int32_t *blab,
                                   const int32_t *belem, const int32_t *bface, const int64_t *gid, int nnbr, const int32_t *nbr,
                                   const int32_t *recv_off, const int32_t *recv_cnt, const int32_t *send_ptr,
                                   const int32_t *send_idx, ffcuda_mesh **out);
/* global ids of the local vertices (owned first): for gathering results / parity checks */
int ffcuda_mesh_local_to_global(ffcuda_mesh *m, int *nowned, int *nlocal, int64_t *gid /* nlocal or NULL */);

#ifdef __cplusplus
}
#endif
#endif
