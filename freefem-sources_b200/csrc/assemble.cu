// assemble.cu — KERNELS 2+3 fused: numeric assembly by ROW-OWNER GATHER with in-register element evaluation.
//
// Replaces AssembleBilinearForm / Element_Op / MatriceElementairePleine::call / HashMatrix::operator+= and
// AssembleLinearForm / Element_rhs of the reference (fflib/problem.cpp:803-1417, 6063-6437, 7839-7985,
// 10555-11227; femlib/MatriceCreuse_tpl.hpp:233-258; femlib/HashMatrix.cpp:1295-1332).
//
// Each node row of the matrix is owned by one thread (P1) or one group of lanes (P2).  The owner walks the
// sorted list of (element, local node a) incidences of its node, evaluates the geometry of the element
// (Jacobian, measure, grad lambda) in fp64 registers and the a-th block row of the element matrix, and adds it
// into the row's CSR segment kept in shared memory at positions precomputed by the symbolic phase.  The segment
// is written to HBM once, coalesced.  No fp64 atomics, no colouring: the summation order of every entry is the
// element order (the reference's own order), so results are bit-reproducible from run to run.
//
// Quadrature: the form's coefficients are element-wise constant (MeshIndependent() terms) and simplices are
// affine, so the quadrature sum  sum_q w_q |K| c D^vop(phi_a)(q) D^uop(phi_b)(q)  of Element_Op factorises into
// reference tensors  R[a][b][s][t] = sum_q w_q B_a^s(q) B_b^t(q)  (s,t: 0 = value, r = d/dxhat_r) evaluated ONCE
// on the host from the very quadrature rule FreeFEM selected (any qforder/qft/qfV rule, lumped ones included),
// contracted per element with the inverse Jacobian.  Same mathematics as the reference loop, different
// summation order (differences ~1e-16 relative).
#include "common.cuh"
#include <memory>
#include <algorithm>
#include <cmath>

// ----------------------------------------------------------------------------------------------------
// host side: reference-element tabulation
// ----------------------------------------------------------------------------------------------------
static const int h_edge3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static const int h_edge2[3][2] = {{1, 2}, {0, 2}, {0, 1}}; // dof 3+e lies on the edge opposite vertex e

// B[a][s]: s=0 value, s=1..dim derivative along reference axis r=s (lambda_r = xhat_r, lambda_0 = 1-sum)
static void ref_basis(int dim, int order, const double *P, double B[10][4])
{
    double l[4] = {1, 0, 0, 0};
    for (int d = 0; d < dim; ++d) {
        l[d + 1] = P[d];
        l[0] -= P[d];
    }
    const int nv = dim + 1;
    auto dl = [&](int a, int r) { return a == 0 ? -1.0 : (a == r ? 1.0 : 0.0); }; // d lambda_a / d xhat_r
    if (order == 1) {
        for (int a = 0; a < nv; ++a) {
            B[a][0] = l[a];
            for (int r = 1; r <= dim; ++r) B[a][r] = dl(a, r);
        }
        return;
    }
    for (int a = 0; a < nv; ++a) {
        B[a][0] = l[a] * (2 * l[a] - 1.);
        for (int r = 1; r <= dim; ++r) B[a][r] = (4 * l[a] - 1) * dl(a, r);
    }
    const int ne = dim == 3 ? 6 : 3;
    for (int e = 0; e < ne; ++e) {
        int i0 = dim == 3 ? h_edge3[e][0] : h_edge2[e][0], i1 = dim == 3 ? h_edge3[e][1] : h_edge2[e][1];
        B[nv + e][0] = 4. * l[i0] * l[i1];
        for (int r = 1; r <= dim; ++r) B[nv + e][r] = 4 * (dl(i1, r) * l[i0] + dl(i0, r) * l[i1]);
    }
}

static int op_slot(int dim, int op)
{
    if (op == FFCUDA_OP_ID) return 0;
    if (op == FFCUDA_OP_DX) return 1;
    if (op == FFCUDA_OP_DY) return 2;
    if (op == FFCUDA_OP_DZ && dim == 3) return 3;
    throw FFError("unsupported differential operator code " + std::to_string(op) +
                  " (only id, dx, dy, dz of P1/P2 forms are on the ffcuda path)");
}

static constexpr int MAXLAB = 16;

struct FormParams {
    double C[3][3][4][4]; // [vcomp][ucomp][vslot][uslot] summed coefficients
    double W;             // sum of weights
    double Lh[4];         // sum_q w_q lambda_a
    double Mh[4][4];      // sum_q w_q lambda_a lambda_b
    double fast_cw, fast_md, fast_mo; // FAST P1 path: c*W*RFAC, m*M_diag*RFAC, m*M_offdiag*RFAC
    double iso_a, iso_b, iso_c;       // ISO P2 path (ncomp = dim): C = a d(cv,sv)d(cu,su) + b d(cv,cu)d(sv,su) + c d(cv,su)d(cu,sv)
    uint32_t mask;        // bit (sv*4+su) set when some C[.][.][sv][su] != 0
    int nlab;             // <0: all regions
    int labels[MAXLAB];
};

struct LinParams {
    double CL[3][4]; // [vcomp][slot]
    int nlab;
    int labels[MAXLAB];
};

static void fill_labels(int nlab, const int32_t *labels, int &onlab, int *olabels)
{
    if (!labels) {
        onlab = -1;
        return;
    }
    FF_REQUIRE(nlab <= MAXLAB, "at most 16 region labels per integral");
    onlab = nlab;
    for (int i = 0; i < nlab; ++i) olabels[i] = labels[i];
}

// A table of the caller (cq / fq / gq of the entries below): a HOST array, copied to the device here, or a DEVICE pointer
// (ffcuda_vec_ptr of a table formed by ffcuda_fe_table), used where it lies.
static const double *table_on_device(const double *t, size_t count, DBuf<double> &buf, cudaStream_t st)
{
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, t);
    if (e != cudaSuccess) (void)cudaGetLastError(); // (plain host memory is reported as an error by old drivers)
    else if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return t;
    buf.alloc(count);
    FF_CUDA(cudaMemcpyAsync(buf.p, t, buf.bytes(), cudaMemcpyHostToDevice, st));
    return buf.p;
}

// ----------------------------------------------------------------------------------------------------
// device: geometry
// ----------------------------------------------------------------------------------------------------
template <int DIM>
struct Geom {
    double g[DIM][DIM]; // g[r-1][x] = d lambda_r / d x , r = 1..DIM
    double mes;
};

// one padded vertex (x,y,z,0) = one 32-byte sector, through the read-only path
__device__ __forceinline__ double4 ldg_vertex(const double *__restrict__ xyz4, int v)
{
    double4 r; // one 256-bit load (LDG.E.256 on sm_100a)
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(xyz4 + 4 * (size_t)v));
    return r;
}

__device__ __forceinline__ void load_geom3(const double *__restrict__ xyz4, int v0, int v1, int v2, int v3, Geom<3> &G,
                                           int want_grad = 1)
{
    const double4 p0 = ldg_vertex(xyz4, v0), p1 = ldg_vertex(xyz4, v1), p2 = ldg_vertex(xyz4, v2), p3 = ldg_vertex(xyz4, v3);
    const double ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    const double bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    const double cx = p3.x - p0.x, cy = p3.y - p0.y, cz = p3.z - p0.z;
    // n1 = V2 x V3, n2 = V3 x V1, n3 = V1 x V2 ; det = V1 . n1 ; grad lambda_r = n_r / det (Mesh3dn.hpp:126-136)
    const double n1x = by * cz - bz * cy, n1y = bz * cx - bx * cz, n1z = bx * cy - by * cx;
    const double det = ax * n1x + ay * n1y + az * n1z;
    G.mes = det * (1.0 / 6.0);
    if (!want_grad) return;
    const double n2x = cy * az - cz * ay, n2y = cz * ax - cx * az, n2z = cx * ay - cy * ax;
    const double n3x = ay * bz - az * by, n3y = az * bx - ax * bz, n3z = ax * by - ay * bx;
    const double inv = 1.0 / det;
    G.g[0][0] = n1x * inv; G.g[0][1] = n1y * inv; G.g[0][2] = n1z * inv;
    G.g[1][0] = n2x * inv; G.g[1][1] = n2y * inv; G.g[1][2] = n2z * inv;
    G.g[2][0] = n3x * inv; G.g[2][1] = n3y * inv; G.g[2][2] = n3z * inv;
}

__device__ __forceinline__ void load_geom2(const double *__restrict__ xy, int v0, int v1, int v2, Geom<2> &G)
{
    const double2 *X = reinterpret_cast<const double2 *>(xy);
    const double2 p0 = __ldg(X + v0), p1 = __ldg(X + v1), p2 = __ldg(X + v2);
    const double bx = p1.x - p0.x, by = p1.y - p0.y, cx = p2.x - p0.x, cy = p2.y - p0.y;
    const double det = bx * cy - by * cx; // 2 area
    const double inv = 1.0 / det;
    // grad lambda_i = (-E.y, E.x)/(2 area), E = edge opposite vertex i (fem.hpp:321-324)
    G.g[0][0] = -(p0.y - p2.y) * inv; G.g[0][1] = (p0.x - p2.x) * inv; // lambda_1: E = v0 - v2
    G.g[1][0] = -(p1.y - p0.y) * inv; G.g[1][1] = (p1.x - p0.x) * inv; // lambda_2: E = v1 - v0
    G.mes = det * 0.5;
}

__device__ __forceinline__ bool region_ok(int nlab, const int *labels, const int32_t *__restrict__ elab, int k)
{
    if (nlab < 0) return true;
    int l = elab[k];
    bool ok = false;
    for (int i = 0; i < nlab; ++i) ok |= (labels[i] == l);
    return ok;
}

// ----------------------------------------------------------------------------------------------------
// P1: one thread per node row, incidence records in the ELL-32 layout (a warp reads 32 consecutive records)
// ----------------------------------------------------------------------------------------------------
// Unscaled "normals" N[b][x] = det * d lambda_b / d x  and det = DIM! * |K| (signed), from the vertex coordinates.
//   3-D: N1 = V2 x V3, N2 = V3 x V1, N3 = V1 x V2, det = V1 . N1 (Mesh3dn.hpp:126-136);  2-D: fem.hpp:321-324.
template <int DIM>
__device__ __forceinline__ void p1_normals(const double (&X)[DIM + 1][DIM], double (&N)[DIM + 1][DIM], double &det)
{
    if (DIM == 3) {
        const double ax = X[1][0] - X[0][0], ay = X[1][1] - X[0][1], az = X[1][DIM - 1] - X[0][DIM - 1];
        const double bx = X[2][0] - X[0][0], by = X[2][1] - X[0][1], bz = X[2][DIM - 1] - X[0][DIM - 1];
        const double cx = X[DIM][0] - X[0][0], cy = X[DIM][1] - X[0][1], cz = X[DIM][DIM - 1] - X[0][DIM - 1];
        N[1][0] = by * cz - bz * cy; N[1][1] = bz * cx - bx * cz; N[1][DIM - 1] = bx * cy - by * cx;
        N[2][0] = cy * az - cz * ay; N[2][1] = cz * ax - cx * az; N[2][DIM - 1] = cx * ay - cy * ax;
        N[DIM][0] = ay * bz - az * by; N[DIM][1] = az * bx - ax * bz; N[DIM][DIM - 1] = ax * by - ay * bx;
        det = ax * N[1][0] + ay * N[1][1] + az * N[1][DIM - 1];
    } else {
        const double bx = X[1][0] - X[0][0], by = X[1][1] - X[0][1], cx = X[2][0] - X[0][0], cy = X[2][1] - X[0][1];
        det = bx * cy - by * cx;
        N[1][0] = cy; N[1][1] = -cx;
        N[2][0] = -by; N[2][1] = bx;
    }
#pragma unroll
    for (int x = 0; x < DIM; ++x) {
        double s = N[1][x];
#pragma unroll
        for (int r = 2; r <= DIM; ++r) s += N[r][x];
        N[0][x] = -s;
    }
}

// The warp's ELL block: stage the coordinates of its distinct vertices (blkvert) in shared memory, SoA, once.
// Returns true when the block is staged.
template <int DIM>
__device__ __forceinline__ bool p1_stage(const double *__restrict__ xyz, const int32_t *__restrict__ blkvert,
                                         const int32_t *__restrict__ blkvcnt, int blk, int lane, int SV, double *stage)
{
    const int vcnt = blkvcnt[blk];
    if (vcnt < 0) return false;
    for (int s = lane; s < vcnt; s += 32) {
        const int v = __ldg(blkvert + (size_t)blk * FF_STAGE_MAX + s);
        if (DIM == 3) {
            const double4 p = ldg_vertex(xyz, v);
            stage[s] = p.x; stage[SV + s] = p.y; stage[2 * SV + s] = p.z;
        } else {
            const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + v);
            stage[s] = p.x; stage[SV + s] = p.y;
        }
    }
    __syncwarp();
    return true;
}

// vertices 1..DIM of a record (owner-first order) from the staged coordinates; X[0] (the owner) is already set
template <int DIM>
__device__ __forceinline__ void p1_points_staged(const double *stage, int SV, uint32_t lw, double (&X)[DIM + 1][DIM])
{
#pragma unroll
    for (int b = 1; b <= DIM; ++b) {
        const int s = (lw >> (8 * b)) & 255;
#pragma unroll
        for (int x = 0; x < DIM; ++x) X[b][x] = stage[x * SV + s];
    }
}

// vertices 1..DIM of element k in owner-first order (a = local index of the owner) straight from global memory
template <int DIM>
__device__ __forceinline__ void p1_points_global(const double *__restrict__ xyz, const int32_t *__restrict__ conn, int k, int a,
                                                 double (&X)[DIM + 1][DIM])
{
#pragma unroll
    for (int i = 1; i <= DIM; ++i) {
        const int o = DIM == 3 ? (a ^ i) : (a + i) % 3;
        const int v = __ldg(conn + (size_t)(DIM + 1) * k + o);
        if (DIM == 3) {
            const double4 p = ldg_vertex(xyz, v);
            X[i][0] = p.x; X[i][1] = p.y; X[i][DIM - 1] = p.z;
        } else {
            const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + v);
            X[i][0] = p.x; X[i][1] = p.y;
        }
    }
}

// Thread per row.  Every record lists its element's vertices OWNER-FIRST (the row's vertex, then the others in an even
// permutation of the element's order): the thread evaluates the simplex with its own vertex as origin, so the row of
// the element matrix it needs is always "row 0": N[0] is its own normal, no selection by local index.
// FAST: scalar space, form = c grad u . grad v (+ m u v) with a symmetric quadrature rule -> a handful of constants.
// GG: only gradient-gradient terms (no value of u or v): the coefficient contraction runs without per-term mask tests
// EM: the form is multiplied by a coefficient that depends on the mesh point; its moments against the rule on every element
// (emom, 24 doubles per element: sum_q w_q c_q | sum_q w_q c_q lambda_a | sum_q w_q c_q lambda_a lambda_b, k_emom below) take the
// place of the constants F.W, F.Lh, F.Mh - exact for P1, whose gradients do not depend on the quadrature node
template <int DIM, int NC, bool FAST, bool GG, bool EM = false>
__global__ void __launch_bounds__(128) k_asm_p1(const double *__restrict__ xyz, const int32_t *__restrict__ conn,
                                                const int32_t *__restrict__ elab, int nrows,
                                                const int32_t *__restrict__ nrowptr, const IncView V,
                                                const uint32_t *__restrict__ loc, const int32_t *__restrict__ blkvert,
                                                const int32_t *__restrict__ blkvcnt, const uint32_t *__restrict__ pos,
                                                double *__restrict__ vals, int S, int SV, int accumulate,
                                                const __grid_constant__ FormParams F, const double *__restrict__ emom = nullptr)
{
    extern __shared__ double smem_d[];
    constexpr int NV = DIM + 1;
    constexpr double RFAC = DIM == 3 ? 1.0 / 6.0 : 0.5; // |K| = det * RFAC
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int row = blockIdx.x * blockDim.x + tid;
    double *stage = smem_d + (size_t)warp * DIM * SV;                     // [DIM][SV] per warp
    double *sacc = smem_d + (size_t)nwarp * DIM * SV;                     // [thread][S]
    double *acc = sacc + (size_t)tid * S;
    int L = 0, rb = 0, mycnt = 0;
    double X[NV][DIM];
    if (row < nrows) {
        rb = nrowptr[row];
        L = nrowptr[row + 1] - rb;
        mycnt = V.cnt[row];
        const int nflat = NC * NC * L;
        for (int j = 0; j < nflat; ++j) acc[j] = 0.0;
        if (DIM == 3) { // P1: node id = vertex id
            const double4 p = ldg_vertex(xyz, row);
            X[0][0] = p.x; X[0][1] = p.y; X[0][DIM - 1] = p.z;
        } else {
            const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + row);
            X[0][0] = p.x; X[0][1] = p.y;
        }
    }
    const int blk = row >> 5, nblk = (nrows + 31) >> 5;
    if (blk < nblk) {
        const bool staged = p1_stage<DIM>(xyz, blkvert, blkvcnt, blk, lane, SV, stage);
        const uint32_t base = V.blkoff[blk];
        const int Lb = (int)((V.blkoff[blk + 1] - base) >> 5);
        const uint32_t *rinc = V.inc + base + lane;
        const uint32_t *rpos = pos + base + lane;
        const uint32_t *rloc = loc + base + lane;
        const bool need_inc = EM || !FAST || !staged || F.nlab >= 0;
        const bool gradgrad = GG || (F.mask & 0xEEE0u) != 0, valgrad = !GG && (F.mask & 0x000Eu) != 0,
                   gradval = !GG && (F.mask & 0x1110u) != 0, valval = !GG && (F.mask & 1u) != 0;
        for (int e = 0; e < Lb; ++e) {
            if (e >= mycnt) continue;
            const uint32_t pw = __ldcs(rpos + (size_t)e * 32);
            int k = 0, a = 0;
            if (need_inc) {
                const uint32_t ka = __ldcs(rinc + (size_t)e * 32);
                k = ka >> 4;
                a = ka & 15;
                if (!region_ok(F.nlab, F.labels, elab, k)) continue;
            }
            double N[NV][DIM], det;
            if (staged) p1_points_staged<DIM>(stage, SV, __ldcs(rloc + (size_t)e * 32), X);
            else p1_points_global<DIM>(xyz, conn, k, a, X);
            p1_normals<DIM>(X, N, det);
            if (FAST) {
                // |K| c W g_0 . g_i = (c W RFAC / det) N_0 . N_i ; |K| m M_0i = RFAC det m M_0i
                const double sgg = F.fast_cw * __drcp_rn(det);
                double w[DIM];
#pragma unroll
                for (int x = 0; x < DIM; ++x) w[x] = sgg * N[0][x];
                const double md = F.fast_md * det, mo = F.fast_mo * det;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    double v = i == 0 ? md : mo;
#pragma unroll
                    for (int x = 0; x < DIM; ++x) v = fma(w[x], N[i][x], v);
                    const int pb = (pw >> (8 * i)) & 255;
                    acc[pb] += v;
                }
            } else {
                // |K| W g_0[sv] g_i[su] = (W RFAC / det) N_0[sv] N_i[su] ; |K| L g_i[su] = RFAC L N_i[su] ; |K| M = RFAC det M
                const double *em = EM ? emom + (size_t)24 * k : nullptr;
                const double sgg = gradgrad ? (EM ? em[0] : F.W) * RFAC * __drcp_rn(det) : 0.0;
                const double La = (EM ? em[1 + a] : F.Lh[a]) * RFAC;
#pragma unroll
                for (int cv = 0; cv < NC; ++cv)
#pragma unroll
                    for (int cu = 0; cu < NC; ++cu) {
                        // wa[su] = sgg * sum_sv C[sv][su] N_0[sv] (+ C[0][su] L_a RFAC) ;  ca0 = RFAC * sum_sv C[sv][0] N_0[sv]
                        double wa[DIM], ca0 = 0.0;
#pragma unroll
                        for (int su = 0; su < DIM; ++su) {
                            double t = 0.0;
#pragma unroll
                            for (int sv = 0; sv < DIM; ++sv)
                                if (GG || (F.mask >> ((sv + 1) * 4 + su + 1) & 1u)) t = fma(F.C[cv][cu][sv + 1][su + 1], N[0][sv], t);
                            wa[su] = t * sgg;
                            if (valgrad) wa[su] = fma(F.C[cv][cu][0][su + 1], La, wa[su]);
                        }
                        if (gradval) {
#pragma unroll
                            for (int sv = 0; sv < DIM; ++sv) ca0 = fma(F.C[cv][cu][sv + 1][0], N[0][sv], ca0);
                            ca0 *= RFAC;
                        }
                        const double cm = valval ? F.C[cv][cu][0][0] * RFAC * det : 0.0;
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            const int o = DIM == 3 ? (a ^ i) : (a + i) % 3; // the element's own local index of vertex i
                            double v = wa[0] * N[i][0];
#pragma unroll
                            for (int su = 1; su < DIM; ++su) v = fma(wa[su], N[i][su], v);
                            if (gradval) v = fma(ca0, EM ? em[1 + o] : F.Lh[o], v);
                            if (valval) v = fma(cm, EM ? em[5 + 4 * a + o] : F.Mh[a][o], v);
                            const int pb = (pw >> (8 * i)) & 255;
                            acc[cv * (NC * L) + pb * NC + cu] += v;
                        }
                    }
            }
        }
    }
    __syncwarp();
    // coalesced write-out: the 32 rows of a warp are contiguous in vals; lanes sweep one row segment at a time
    const int wbase = tid & ~31;
    for (int r = 0; r < 32; ++r) {
        const int rrb = __shfl_sync(0xffffffffu, rb, r), rL = __shfl_sync(0xffffffffu, L, r);
        const int nflat = NC * NC * rL;
        const double *src = sacc + (size_t)(wbase + r) * S;
        double *dst = vals + (size_t)NC * NC * rrb;
        for (int j = lane; j < nflat; j += 32) dst[j] = accumulate ? dst[j] + src[j] : src[j];
    }
}

// reciprocal of a double to ~1 ulp without the slow path of a correctly rounded division: hardware seed + two Newton steps
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double t = fma(-d, r, 1.0);
    r = fma(r, t, r);
    t = fma(-d, r, 1.0);
    return fma(r, t, r);
}

// shared-memory accesses through 32-bit shared-space addresses kept in registers (the compiler otherwise re-derives the
// generic bases - thread id, kernel parameters, window base - inside the record loop: ~25 of its 140 instructions)
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// The lean thread-per-row kernel: scalar P1, form = c grad u . grad v (+ m u v), no region filter, every ELL block
// staged.  It serves the assemblies that are not on the tile path (first assembly on a fespace, forms with a mass term).
// Per record it reads 8 bytes (position word + slot word), everything else is shared memory and fp64 registers; records
// are prefetched two iterations ahead; the stage holds the block's vertices as (x, y, z) triples (one address per vertex).
template <int DIM>
__global__ void __launch_bounds__(128) k_asm_p1_lean(const double *__restrict__ xyz, int nrows, const int32_t *__restrict__ nrowptr,
                                                     const int32_t *__restrict__ cnt, const uint32_t *__restrict__ blkoff,
                                                     const uint32_t *__restrict__ loc, const int32_t *__restrict__ blkvert,
                                                     const int32_t *__restrict__ blkvcnt, const uint32_t *__restrict__ pos,
                                                     double *__restrict__ vals, int S, int SV, int accumulate, double cw, double cmd,
                                                     double cmo)
{
    extern __shared__ double smem_d[];
    constexpr int NV = DIM + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int row = blockIdx.x * blockDim.x + tid;
    double *stage = smem_d + (size_t)warp * DIM * SV;               // [slot][DIM]
    double *sacc = smem_d + (size_t)nwarp * DIM * SV;
    double *acc = sacc + (size_t)tid * S;
    const uint32_t stage_a = (uint32_t)__cvta_generic_to_shared(stage), acc_a = (uint32_t)__cvta_generic_to_shared(acc);
    int L = 0, rb = 0, mycnt = 0;
    double X0[DIM];
#pragma unroll
    for (int x = 0; x < DIM; ++x) X0[x] = 0.0;
    if (row < nrows) {
        rb = nrowptr[row];
        L = nrowptr[row + 1] - rb;
        mycnt = cnt[row];
        for (int j = 0; j < L; ++j) acc[j] = 0.0;
        if (DIM == 3) {
            const double4 p = ldg_vertex(xyz, row);
            X0[0] = p.x; X0[1] = p.y; X0[DIM - 1] = p.z;
        } else {
            const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + row);
            X0[0] = p.x; X0[1] = p.y;
        }
    }
    const int blk = row >> 5, nblk = (nrows + 31) >> 5;
    if (blk < nblk) {
        {   // stage the block's distinct vertices
            const int vcnt = blkvcnt[blk];
            for (int s = lane; s < vcnt; s += 32) {
                const int v = __ldg(blkvert + (size_t)blk * FF_STAGE_MAX + s);
                if (DIM == 3) {
                    const double4 p = ldg_vertex(xyz, v);
                    stage[DIM * s] = p.x; stage[DIM * s + 1] = p.y; stage[DIM * s + DIM - 1] = p.z;
                } else {
                    const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + v);
                    stage[DIM * s] = p.x; stage[DIM * s + 1] = p.y;
                }
            }
            __syncwarp();
        }
        const uint32_t base = blkoff[blk];
        const int Lb = (int)((blkoff[blk + 1] - base) >> 5);
        const uint32_t *ppos = pos + base + lane, *ploc = loc + base + lane;
        uint32_t pw0 = 0, lw0 = 0, pw1 = 0, lw1 = 0; // records e and e+1
        if (Lb > 0) {
            pw0 = __ldcs(ppos);
            lw0 = __ldcs(ploc);
        }
        if (Lb > 1) {
            pw1 = __ldcs(ppos + 32);
            lw1 = __ldcs(ploc + 32);
        }
        double sdet = 0.0;
        int dpos = 0;
        for (int e = 0; e < Lb; ++e) {
            const uint32_t pwc = pw0, lwc = lw0;
            pw0 = pw1;
            lw0 = lw1;
            if (e + 2 < Lb) { // prefetch two records ahead (padding records hold unused words)
                pw1 = __ldcs(ppos + (size_t)(e + 2) * 32);
                lw1 = __ldcs(ploc + (size_t)(e + 2) * 32);
            }
            if (e < mycnt) {
                double X[NV][DIM], N[NV][DIM], det;
#pragma unroll
                for (int x = 0; x < DIM; ++x) X[0][x] = X0[x];
#pragma unroll
                for (int b = 1; b <= DIM; ++b) {
                    const uint32_t va = stage_a + ((lwc >> (8 * b)) & 255u) * (DIM * 8);
#pragma unroll
                    for (int x = 0; x < DIM; ++x) X[b][x] = lds_f64(va + 8 * x);
                }
                p1_normals<DIM>(X, N, det);
                const double sgg = cw * fast_rcp(det);
                double w[DIM];
#pragma unroll
                for (int x = 0; x < DIM; ++x) w[x] = sgg * N[0][x];
                const double mo = cmo * det;
                sdet += det;
                dpos = pwc & 255;
                // the DIM off-diagonal entries of the owner's row of the element matrix; their columns are distinct, so
                // the read-modify-writes are independent: all loads first, then all stores
                double v[NV], a[NV];
                uint32_t pa[NV];
#pragma unroll
                for (int i = 1; i < NV; ++i) {
                    v[i] = mo;
#pragma unroll
                    for (int x = 0; x < DIM; ++x) v[i] = fma(w[x], N[i][x], v[i]);
                    pa[i] = acc_a + ((pwc >> (8 * i)) & 255u) * 8;
                    a[i] = lds_f64(pa[i]);
                }
#pragma unroll
                for (int i = 1; i < NV; ++i) sts_f64(pa[i], a[i] + v[i]);
            }
        }
        // The diagonal is not accumulated record by record: the P1 basis is a partition of unity, so the stiffness part
        // of a row sums to zero, K_ii = -sum_{j != i} K_ij; the mass part is m (M_d - DIM M_o) |K| summed over the star.
        if (mycnt > 0) {
            double off = 0.0;
            for (int j = 0; j < L; ++j)
                if (j != dpos) off += acc[j];
            acc[dpos] = (cmd + DIM * cmo) * sdet - off;
        }
    }
    __syncwarp();
    const int wbase = tid & ~31;
    for (int r = 0; r < 32; ++r) {
        const int rrb = __shfl_sync(0xffffffffu, rb, r), rL = __shfl_sync(0xffffffffu, L, r);
        const double *src = sacc + (size_t)(wbase + r) * S;
        double *dst = vals + (size_t)rrb;
        for (int j = lane; j < rL; j += 32) dst[j] = accumulate ? dst[j] + src[j] : src[j];
    }
}

// Geometry of every element once per assembly (P2 rows): gradients of the barycentric coordinates and the measure, GS
// doubles per element.  k_asm_p2 visits an element once per (node row, lane) - 10 rows x 10 lanes on tetrahedra - and used
// to re-derive the inverse Jacobian each time (ncu r01d: FP64 pipe 34 %, a third of it this geometry): now one 80-byte
// broadcast load per visit.
template <int DIM>
__global__ void k_elem_geom(const double *__restrict__ xyz, const int32_t *__restrict__ conn, int nt, double *__restrict__ egeo)
{
    constexpr int GS = DIM == 3 ? 10 : 6;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt) return;
    Geom<DIM> G;
    if (DIM == 3) {
        const int4 K = __ldg(reinterpret_cast<const int4 *>(conn) + k);
        load_geom3(xyz, K.x, K.y, K.z, K.w, reinterpret_cast<Geom<3> &>(G));
    } else {
        load_geom2(xyz, __ldg(conn + 3 * (size_t)k), __ldg(conn + 3 * (size_t)k + 1), __ldg(conn + 3 * (size_t)k + 2), reinterpret_cast<Geom<2> &>(G));
    }
    double *o = egeo + (size_t)k * GS;
#pragma unroll
    for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int x = 0; x < DIM; ++x) o[r * DIM + x] = G.g[r][x];
    o[DIM * DIM] = G.mes;
    if (DIM == 2) o[5] = 0.0;
}

// ----------------------------------------------------------------------------------------------------
// P2: one group of GL lanes per node row, lane b handles the pair (a, b)
// ----------------------------------------------------------------------------------------------------
// GG: the form only couples gradients (d_x u d_y v terms: Laplace, Lame, ...): the value row/column of the per-pair
// tensor is not formed and the coefficient contraction runs over the DIM x DIM gradient block without per-term tests
// ISO (with GG, ncomp = dim): the coefficient tensor is isotropic - lambda div u div v + 2 mu eps(u):eps(v), vector Laplacians,
// div-div terms - so the contraction with the DIM x DIM block M of a node pair collapses from (NC DIM)^2 to 2 NC^2 + DIM
// operations: v[cv][cu] = a M[cv][cu] + c M[cu][cv] + (cv == cu) b tr M
// QC: every term is multiplied by a coefficient that depends on the mesh point, given at the quadrature nodes (cq[k][q]): the
// per-pair tensor of the element, sum_q c_q w_q d^sa phi_a(q) d^sb phi_b(q), is formed on the fly from the constant table
// Tq[q][a][b][sa][sb] (L1/L2 resident, 16 contiguous doubles per lane) instead of being read from the shared reference tensor
template <int DIM, int NC, int GL, typename PosT, bool GG, bool ISO = false, bool QC = false>
__global__ void __launch_bounds__(128) k_asm_p2(const double *__restrict__ xyz, const int32_t *__restrict__ conn,
                                                const int32_t *__restrict__ elab, int nrows,
                                                const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ incptr,
                                                const uint32_t *__restrict__ inc, const PosT *__restrict__ pos,
                                                const double *__restrict__ Rg, double *__restrict__ vals, int S, int accumulate,
                                                const __grid_constant__ FormParams F, const int32_t *__restrict__ rowperm, int row0,
                                                const double *__restrict__ Tq = nullptr, const double *__restrict__ cq = nullptr, int nqc = 0,
                                                const double *__restrict__ egeo = nullptr)
{
    constexpr int NL = DIM == 3 ? 10 : 6;
    constexpr int NS = DIM + 1;
    constexpr int RS = NS * NS + 1; // padded stride of one (a,b) tensor: odd -> conflict-free across b
    extern __shared__ double smem[];
    double *sR = smem;                         // NL*NL*RS
    double *sacc = smem + NL * NL * RS;        // groups * S
    const int tid = threadIdx.x;
    for (int x = tid; x < NL * NL * NS * NS; x += blockDim.x) {
        int ab = x / (NS * NS), st = x - ab * NS * NS;
        sR[ab * RS + st] = Rg[x];
    }
    const int grp = tid / GL, b = tid % GL;
    const int groups = blockDim.x / GL;
    // rows are taken in the order of rowperm (longest first, see launch_p2): the groups of a block then own rows of
    // about the same cost, and the launch over the short rows can run with small accumulators.  nrows = end of the range.
    const int ridx = row0 + blockIdx.x * groups + grp;
    double *acc = sacc + (size_t)grp * S;
    __syncthreads();
    if (ridx >= nrows) return;
    const int row = rowperm ? rowperm[ridx] : ridx;
    const int rb = nrowptr[row], L = nrowptr[row + 1] - rb;
    const int nflat = NC * NC * L;
    for (int j = b; j < nflat; j += GL) acc[j] = 0.0;
    // the mask of the lanes of this group inside the warp
    const unsigned gmask = (GL == 32) ? 0xffffffffu : (((1u << GL) - 1u) << ((tid & 31) / GL * GL));
    __syncwarp(gmask);
    const int ie = incptr[row + 1];
    for (int e = incptr[row]; e < ie; ++e) {
        const uint32_t ka = __ldg(inc + e);
        const int k = ka >> 4, a = ka & 15;
        if (region_ok(F.nlab, F.labels, elab, k) && b < NL) {
            Geom<DIM> G;
            if (egeo) { // computed once per element by k_elem_geom: the lanes of the group read the same 16-byte words (broadcast)
                constexpr int GS = DIM == 3 ? 10 : 6;
                const double2 *gp = reinterpret_cast<const double2 *>(egeo + (size_t)k * GS);
                double t[GS];
#pragma unroll
                for (int i = 0; i < GS / 2; ++i) {
                    const double2 v2 = __ldg(gp + i);
                    t[2 * i] = v2.x;
                    t[2 * i + 1] = v2.y;
                }
#pragma unroll
                for (int r = 0; r < DIM; ++r)
#pragma unroll
                    for (int x = 0; x < DIM; ++x) G.g[r][x] = t[r * DIM + x];
                G.mes = t[DIM * DIM];
            } else if (DIM == 3) {
                const int4 K = __ldg(reinterpret_cast<const int4 *>(conn) + k);
                load_geom3(xyz, K.x, K.y, K.z, K.w, reinterpret_cast<Geom<3> &>(G));
            } else {
                const int v0 = __ldg(conn + 3 * (size_t)k), v1 = __ldg(conn + 3 * (size_t)k + 1), v2 = __ldg(conn + 3 * (size_t)k + 2);
                load_geom2(xyz, v0, v1, v2, reinterpret_cast<Geom<2> &>(G));
            }
            const int pb = pos[(size_t)e * NL + b];
            double Rl[QC ? NS * NS : 1];
            if (QC) {
#pragma unroll
                for (int st = 0; st < NS * NS; ++st) Rl[st] = 0.0;
                const double *c = cq + (size_t)k * nqc;
                const double *T = Tq + ((size_t)a * NL + b) * (NS * NS);
                for (int q = 0; q < nqc; ++q) {
                    const double cv_ = __ldg(c + q);
                    const double *Tp = T + (size_t)q * NL * NL * NS * NS;
#pragma unroll
                    for (int st = GG ? NS + 1 : 0; st < NS * NS; ++st)
                        if (!GG || st % NS) Rl[st] = fma(cv_, __ldg(Tp + st), Rl[st]);
                }
            }
            const double *R = QC ? Rl : sR + (a * NL + b) * RS;
            double M[NS][NS];
            if (!GG) {
                M[0][0] = R[0];
#pragma unroll
                for (int x = 0; x < DIM; ++x) {
                    double s0 = 0, s1 = 0;
#pragma unroll
                    for (int r = 0; r < DIM; ++r) {
                        s0 = fma(R[r + 1], G.g[r][x], s0);          // value(a) * d_x(b)
                        s1 = fma(R[(r + 1) * NS], G.g[r][x], s1);   // d_x(a) * value(b)
                    }
                    M[0][x + 1] = s0;
                    M[x + 1][0] = s1;
                }
            }
            // Y[r][x] = sum_r' R[r][r'] g[r'][x] ; M[sv][su] = sum_r g[r][sv] Y[r][su]
            double Y[DIM][DIM];
#pragma unroll
            for (int r = 0; r < DIM; ++r)
#pragma unroll
                for (int x = 0; x < DIM; ++x) {
                    double s = 0;
#pragma unroll
                    for (int q = 0; q < DIM; ++q) s = fma(R[(r + 1) * NS + q + 1], G.g[q][x], s);
                    Y[r][x] = s;
                }
#pragma unroll
            for (int sv = 0; sv < DIM; ++sv)
#pragma unroll
                for (int su = 0; su < DIM; ++su) {
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < DIM; ++r) s = fma(G.g[r][sv], Y[r][su], s);
                    M[sv + 1][su + 1] = s;
                }
#pragma unroll
            for (int cv = 0; cv < NC; ++cv)
#pragma unroll
                for (int cu = 0; cu < NC; ++cu) {
                    double v = 0.0;
                    if (ISO) {
                        v = F.iso_a * M[cv + 1][cu + 1];
                        v = fma(F.iso_c, M[cu + 1][cv + 1], v);
                        if (cv == cu) {
                            double tr = M[1][1];
#pragma unroll
                            for (int x = 2; x < NS; ++x) tr += M[x][x];
                            v = fma(F.iso_b, tr, v);
                        }
                    } else if (GG) {
#pragma unroll
                        for (int sv = 1; sv < NS; ++sv)
#pragma unroll
                            for (int su = 1; su < NS; ++su) v = fma(F.C[cv][cu][sv][su], M[sv][su], v);
                    } else {
#pragma unroll
                        for (int sv = 0; sv < NS; ++sv)
#pragma unroll
                            for (int su = 0; su < NS; ++su)
                                if (F.mask >> (sv * 4 + su) & 1u) v = fma(F.C[cv][cu][sv][su], M[sv][su], v);
                    }
                    acc[cv * (NC * L) + pb * NC + cu] += G.mes * v;
                }
        }
        __syncwarp(gmask);
    }
    double *dst = vals + (size_t)NC * NC * rb;
    for (int j = b; j < nflat; j += GL) dst[j] = accumulate ? dst[j] + acc[j] : acc[j];
}

// ----------------------------------------------------------------------------------------------------
// right-hand side: one thread per node row (P1: ELL-32 records, P2: CSR lists)
// ----------------------------------------------------------------------------------------------------
template <int DIM, int NC>
__global__ void __launch_bounds__(128) k_rhs(const double *__restrict__ xyz, const int32_t *__restrict__ conn,
                                             const int32_t *__restrict__ elab, int nrows, const IncView V,
                                             const uint32_t *__restrict__ loc, const int32_t *__restrict__ blkvert,
                                             const int32_t *__restrict__ blkvcnt, int SV,
                                             const double *__restrict__ Fh /* nloc*(DIM+1) */, int nloc, int hasgrad,
                                             double *__restrict__ bvec, int accumulate, const __grid_constant__ LinParams Lp)
{
    __shared__ double sF[10 * 4];
    for (int x = threadIdx.x; x < nloc * (DIM + 1); x += blockDim.x) sF[x] = Fh[x];
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    double out[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) out[c] = 0.0;
    // record walk: ELL -> every lane of the warp runs the padded length of its block; CSR -> its own list
    int ne = 0, stride = 1;
    const uint32_t *rinc = V.inc;
    if (V.ell) {
        const int blk = row >> 5, nblk = (nrows + 31) >> 5;
        if (blk < nblk) {
            const uint32_t base = V.blkoff[blk];
            ne = (int)((V.blkoff[blk + 1] - base) >> 5);
            rinc = V.inc + base + lane;
            stride = 32;
        }
    } else if (row < nrows) {
        ne = V.cnt[row];
        rinc = V.inc + V.incptr[row];
    }
    extern __shared__ double stage_all[];
    bool staged = false;
    double *stage = stage_all + (size_t)(threadIdx.x >> 5) * DIM * SV;
    const uint32_t *rloc = nullptr;
    double X[DIM + 1][DIM];
    if (V.ell && ne > 0) {
        const int blk = row >> 5;
        staged = p1_stage<DIM>(xyz, blkvert, blkvcnt, blk, lane, SV, stage);
        rloc = loc + V.blkoff[blk] + lane;
        if (row < nrows) { // P1: node id = vertex id
            if (DIM == 3) {
                const double4 p = ldg_vertex(xyz, row);
                X[0][0] = p.x; X[0][1] = p.y; X[0][DIM - 1] = p.z;
            } else {
                const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + row);
                X[0][0] = p.x; X[0][1] = p.y;
            }
        }
    }
    constexpr double RFAC = DIM == 3 ? 1.0 / 6.0 : 0.5;
    for (int e = 0; e < ne; ++e) {
        const uint32_t ka = __ldcs(rinc + (size_t)e * stride);
        if (ka == FF_NOREC) continue;
        const int k = ka >> 4, a = ka & 15;
        if (!region_ok(Lp.nlab, Lp.labels, elab, k)) continue;
        double Fa[DIM + 1], mes;
        Fa[0] = sF[a * (DIM + 1)];
        if (staged) {
            // P1 from the staged coordinates, owner-first: |K| = det RFAC, grad of the row's own basis function = N_0 / det,
            // and sum_q w_q d(lambda_a)/d(xhat_r) g_r = W grad(lambda_a) with W = -Fh[0][1]
            double N[DIM + 1][DIM], det;
            p1_points_staged<DIM>(stage, SV, __ldcs(rloc + (size_t)e * 32), X);
            p1_normals<DIM>(X, N, det);
            mes = det * RFAC;
            const double winv = hasgrad ? -sF[1] * __drcp_rn(det) : 0.0;
#pragma unroll
            for (int x = 0; x < DIM; ++x) Fa[x + 1] = N[0][x] * winv;
        } else {
            Geom<DIM> G;
            if (DIM == 3) {
                const int4 K = __ldg(reinterpret_cast<const int4 *>(conn) + k);
                load_geom3(xyz, K.x, K.y, K.z, K.w, reinterpret_cast<Geom<3> &>(G), hasgrad);
            } else {
                const int v0 = __ldg(conn + 3 * (size_t)k), v1 = __ldg(conn + 3 * (size_t)k + 1), v2 = __ldg(conn + 3 * (size_t)k + 2);
                load_geom2(xyz, v0, v1, v2, reinterpret_cast<Geom<2> &>(G));
            }
            mes = G.mes;
#pragma unroll
            for (int x = 0; x < DIM; ++x) {
                double s = 0;
                if (hasgrad) {
#pragma unroll
                    for (int r = 0; r < DIM; ++r) s = fma(sF[a * (DIM + 1) + r + 1], G.g[r][x], s);
                }
                Fa[x + 1] = s;
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double v = 0;
#pragma unroll
            for (int s = 0; s <= DIM; ++s) v = fma(Lp.CL[c][s], Fa[s], v);
            out[c] += mes * v;
        }
    }
    if (row < nrows) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            size_t d = (size_t)row * NC + c;
            bvec[d] = accumulate ? bvec[d] + out[c] : out[c];
        }
    }
}

// Lean right-hand side: scalar or vector P1, no region filter, every block staged, symmetric rule (the same
// sum_q w_q lambda_a for every a).  b_i += |K| (cval[c] L + W grad(lambda_i) . cgrad[c]) over the elements around i.
template <int DIM, int NC, bool GRAD>
__global__ void __launch_bounds__(128) k_rhs_p1_lean(const double *__restrict__ xyz, int nrows, const int32_t *__restrict__ cnt,
                                                     const uint32_t *__restrict__ blkoff, const uint32_t *__restrict__ loc,
                                                     const int32_t *__restrict__ blkvert, const int32_t *__restrict__ blkvcnt, int SV,
                                                     double Lval, double Wsum, double *__restrict__ bvec, int accumulate,
                                                     const __grid_constant__ LinParams Lp)
{
    extern __shared__ double stage_all[];
    constexpr int NV = DIM + 1;
    constexpr double RFAC = DIM == 3 ? 1.0 / 6.0 : 0.5;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    double *stage = stage_all + (size_t)(threadIdx.x >> 5) * DIM * SV;
    double X[NV][DIM];
    int mycnt = 0;
    if (row < nrows) {
        mycnt = cnt[row];
        if (DIM == 3) {
            const double4 p = ldg_vertex(xyz, row);
            X[0][0] = p.x; X[0][1] = p.y; X[0][DIM - 1] = p.z;
        } else {
            const double2 p = __ldg(reinterpret_cast<const double2 *>(xyz) + row);
            X[0][0] = p.x; X[0][1] = p.y;
        }
    }
    double sdet = 0.0, sn[DIM];
#pragma unroll
    for (int x = 0; x < DIM; ++x) sn[x] = 0.0;
    const int blk = row >> 5, nblk = (nrows + 31) >> 5;
    if (blk < nblk) {
        p1_stage<DIM>(xyz, blkvert, blkvcnt, blk, lane, SV, stage);
        const uint32_t base = blkoff[blk];
        const int Lb = (int)((blkoff[blk + 1] - base) >> 5);
        const uint32_t *ploc = loc + base + lane;
        uint32_t lw = Lb > 0 ? __ldcs(ploc) : 0;
        for (int e = 0; e < Lb; ++e) {
            const uint32_t lwc = lw;
            ploc += 32;
            if (e + 1 < Lb) lw = __ldcs(ploc);
            if (e < mycnt) {
                p1_points_staged<DIM>(stage, SV, lwc, X);
                if (GRAD) {
                    // |K| grad(lambda_own) = RFAC N_0 : independent of det, so the normals are simply summed
                    double N[NV][DIM], det;
                    p1_normals<DIM>(X, N, det);
                    sdet += det;
#pragma unroll
                    for (int x = 0; x < DIM; ++x) sn[x] += N[0][x];
                } else if (DIM == 3) {
                    const double ax = X[1][0] - X[0][0], ay = X[1][1] - X[0][1], az = X[1][DIM - 1] - X[0][DIM - 1];
                    const double bx = X[2][0] - X[0][0], by = X[2][1] - X[0][1], bz = X[2][DIM - 1] - X[0][DIM - 1];
                    const double cx = X[DIM][0] - X[0][0], cy = X[DIM][1] - X[0][1], cz = X[DIM][DIM - 1] - X[0][DIM - 1];
                    sdet += ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
                } else {
                    sdet += (X[1][0] - X[0][0]) * (X[2][1] - X[0][1]) - (X[1][1] - X[0][1]) * (X[2][0] - X[0][0]);
                }
            }
        }
    }
    if (row < nrows) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double v = Lp.CL[c][0] * Lval * RFAC * sdet;
            if (GRAD) {
#pragma unroll
                for (int x = 0; x < DIM; ++x) v = fma(Lp.CL[c][x + 1] * Wsum * RFAC, sn[x], v);
            }
            const size_t d = (size_t)row * NC + c;
            bvec[d] = accumulate ? bvec[d] + v : v;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// host drivers
// ----------------------------------------------------------------------------------------------------
template <int DIM, int NC>
static void launch_p1(ffcuda_ctx *ctx, ffcuda_matrix *A, ffcuda_space *s, const FormParams &F, int accumulate, bool fast,
                      const double *emom = nullptr)
{
    ffcuda_pattern *P = A->pattern;
    ffcuda_mesh *m = s->mesh;
    if (emom) fast = false; // per-element moments: the general thread-per-row kernel
    // scalar c grad u.grad v + m u v without region filter: row tiles (tiles.cu) when the space has / may build them
    if (NC == 1 && fast && F.nlab < 0 && ff_asm_p1_tiles(ctx, A, s, F.fast_cw, F.fast_md, F.fast_mo, accumulate)) return;
    ff_pattern_ensure_pos(P);
    FF_REQUIRE(P->pos8.p, "internal: P1 pattern without 8-bit positions");
    const Incidence &I = s->incidence;
    int S = NC * NC * P->maxrow_node;
    S |= 1; // odd stride: threads of a warp land in different banks
    const int SV = std::max(4, (I.maxstage + 3) & ~3); // staged vertices per warp (SoA, DIM planes)
    int threads = 128;
    while (threads > 32 && (size_t)threads * S * 8 + (size_t)(threads / 32) * DIM * SV * 8 > 48 * 1024) threads >>= 1;
    size_t shmem = (size_t)threads * S * 8 + (size_t)(threads / 32) * DIM * SV * 8;
    FF_REQUIRE(shmem <= 200 * 1024, "matrix rows too long for the shared-memory row accumulators");
    if (NC == 1 && fast && F.nlab < 0 && I.nunstaged == 0) {
        auto lean = k_asm_p1_lean<DIM>;
        FF_CUDA(cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        const int nw = (P->nrows_node + 31) / 32;
        ff_launch(ctx, "asm_rows_p1", [&] {
            lean<<<ff_blocks((size_t)nw * 32, threads), threads, shmem, ctx->stream>>>(
                m->xyz.p, P->nrows_node, P->nrowptr.p, I.cnt.p, I.blkoff.p, I.loc.p, I.blkvert.p, I.blkvcnt.p,
                reinterpret_cast<const uint32_t *>(P->pos8.p), A->vals.p, S, SV, accumulate, F.fast_cw, F.fast_md, F.fast_mo);
        });
        return;
    }
    const bool gg = (F.mask & 0x111Fu) == 0 && (F.mask & 0xEEE0u) != 0;
    auto kern = emom ? (gg ? k_asm_p1<DIM, NC, false, true, true> : k_asm_p1<DIM, NC, false, false, true>)
                     : (NC == 1 && fast) ? k_asm_p1<DIM, NC, (NC == 1), false>
                                         : (gg ? k_asm_p1<DIM, NC, false, true> : k_asm_p1<DIM, NC, false, false>);
    FF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    // rows are taken in whole ELL blocks of 32: the grid covers ceil(nrows/32) warps
    const int nwarps = (P->nrows_node + 31) / 32;
    int blocks = ff_blocks((size_t)nwarps * 32, threads);
    const IncView V = ff_view(s->incidence);
    ff_launch(ctx, "asm_rows_p1", [&] {
        kern<<<blocks, threads, shmem, ctx->stream>>>(m->xyz.p, m->conn.p, m->elab.p, P->nrows_node, P->nrowptr.p, V, I.loc.p,
                                                      I.blkvert.p, I.blkvcnt.p, reinterpret_cast<const uint32_t *>(P->pos8.p),
                                                      A->vals.p, S, SV, accumulate, F, emom);
    });
}

// rows of a P2 space sorted by decreasing length (tiles.cu: CUB radix sort), built once per space
void ff_p2_row_order(ffcuda_ctx *ctx, ffcuda_space *s, const int32_t *nrowptr, int nrows, int maxrow);

template <int DIM, int NC, typename PosT>
static void launch_p2(ffcuda_ctx *ctx, ffcuda_matrix *A, ffcuda_space *s, const FormParams &F, const double *Rg, const PosT *pos,
                      int accumulate, const double *Tq = nullptr, const double *cq = nullptr, int nqc = 0)
{
    constexpr int GL = DIM == 3 ? 16 : 8;
    constexpr int NL = DIM == 3 ? 10 : 6;
    constexpr int RS = (DIM + 1) * (DIM + 1) + 1;
    ffcuda_pattern *P = A->pattern;
    ffcuda_mesh *m = s->mesh;
    const bool gg = (F.mask & 0x111Fu) == 0; // no term involves the value of u or v
    // isotropic coefficient tensor (ncomp = dim): exact structural test on the summed coefficients
    bool iso = gg && NC == DIM && !cq;
    FormParams Fi = F;
    if (iso) {
        const double a = F.C[0][1][1][2], c = F.C[0][1][2][1], bb = F.C[0][0][2][2];
        for (int cv = 0; cv < NC && iso; ++cv)
            for (int cu = 0; cu < NC && iso; ++cu)
                for (int sv = 1; sv <= DIM && iso; ++sv)
                    for (int su = 1; su <= DIM; ++su) {
                        const double want = (cv + 1 == sv && cu + 1 == su ? a : 0.0) + (cv == cu && sv == su ? bb : 0.0) +
                                            (cv + 1 == su && cu + 1 == sv ? c : 0.0);
                        if (F.C[cv][cu][sv][su] != want && fabs(F.C[cv][cu][sv][su] - want) > 4e-16 * fabs(want)) iso = false;
                    }
        Fi.iso_a = a;
        Fi.iso_b = bb;
        Fi.iso_c = c;
    }
    auto kern = cq    ? (gg ? k_asm_p2<DIM, NC, GL, PosT, true, false, true> : k_asm_p2<DIM, NC, GL, PosT, false, false, true>)
                : iso ? k_asm_p2<DIM, NC, GL, PosT, true, (NC == DIM)>
                : gg  ? k_asm_p2<DIM, NC, GL, PosT, true>
                      : k_asm_p2<DIM, NC, GL, PosT, false>;
    const Incidence &I = s->incidence;
    // Vertex nodes have ~5x the elements and ~2.5x the row length of edge nodes, and the numbering interleaves them: with
    // rows in natural order a block waits for its one vertex row while its other groups idle (13 % achieved occupancy,
    // ncu r01d), and every group pays the shared-memory accumulator of the longest row.  Rows are therefore taken in order
    // of decreasing length, in two launches: the long rows, then the short ones with accumulators sized for them.
    ff_p2_row_order(ctx, s, P->nrowptr.p, P->nrows_node, P->maxrow_node);
    const int nrows = P->nrows_node, nlong = s->p2_nlong;
    // geometry of every element, once per assembly
    constexpr int GS = DIM == 3 ? 10 : 6;
    DBuf<double> egeo;
    const bool pregeo = !(getenv("FFCUDA_P2_PREGEOM") && atoi(getenv("FFCUDA_P2_PREGEOM")) == 0);
    if (pregeo) {
        egeo.alloc((size_t)m->nt * GS);
        ff_launch(ctx, "asm_p2_geom", [&] { k_elem_geom<DIM><<<ff_blocks((size_t)m->nt, 256), 256, 0, ctx->stream>>>(m->xyz.p, m->conn.p, m->nt, egeo.p); });
    }
    auto run = [&](int r0, int r1, int maxL) {
        if (r1 <= r0) return;
        int S = NC * NC * maxL;
        int threads = 128;
        while (threads > GL && (size_t)(threads / GL) * S * 8 > 96 * 1024) threads >>= 1;
        const int groups = threads / GL;
        const size_t shmem = ((size_t)NL * NL * RS + (size_t)groups * S) * 8;
        FF_REQUIRE(shmem <= 220 * 1024, "matrix rows too long for the shared-memory row accumulators");
        FF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        const int blocks = ff_blocks((size_t)(r1 - r0), groups);
        ff_launch(ctx, "asm_rows_p2", [&] {
            kern<<<blocks, threads, shmem, ctx->stream>>>(m->xyz.p, m->conn.p, m->elab.p, r1, P->nrowptr.p, I.incptr.p, I.inc.p, pos, Rg,
                                                          A->vals.p, S, accumulate, Fi, s->p2_rowperm.p, r0, Tq, cq, nqc, egeo.p);
        });
    };
    run(0, nlong, P->maxrow_node);
    run(nlong, nrows, s->p2_short_maxrow);
}

template <int DIM>
static void dispatch_p1(ffcuda_ctx *ctx, ffcuda_matrix *A, ffcuda_space *s, const FormParams &F, int accumulate, bool fast,
                        const double *emom = nullptr)
{
    const int nc = s->ncomp;
    if (nc == 1) launch_p1<DIM, 1>(ctx, A, s, F, accumulate, fast, emom);
    else if (nc == 2) launch_p1<DIM, 2>(ctx, A, s, F, accumulate, false, emom);
    else launch_p1<DIM, 3>(ctx, A, s, F, accumulate, false, emom);
}

// moments of a coefficient given at the quadrature nodes against the rule, per element (P1): em[0] = sum_q w_q c_q,
// em[1+a] = sum_q w_q c_q lambda_a(q), em[5+4a+b] = sum_q w_q c_q lambda_a(q) lambda_b(q); 24 doubles per element
__global__ void k_emom(int nt, int nq, int nv, const double *__restrict__ wl /* [q][0] = w_q, [q][1+a] = lambda_a(q); stride 5 */,
                       const double *__restrict__ cq, double *__restrict__ emom)
{
    extern __shared__ double swl[];
    for (int i = threadIdx.x; i < nq * 5; i += blockDim.x) swl[i] = wl[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt) return;
    double m[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) m[i] = 0.0;
    const double *c = cq + (size_t)k * nq;
    for (int q = 0; q < nq; ++q) {
        const double wc = swl[q * 5] * c[q];
        m[0] += wc;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double la = a < nv ? swl[q * 5 + 1 + a] : 0.0;
            m[1 + a] = fma(wc, la, m[1 + a]);
#pragma unroll
            for (int b = 0; b < 4; ++b) m[5 + 4 * a + b] = fma(wc * la, b < nv ? swl[q * 5 + 1 + b] : 0.0, m[5 + 4 * a + b]);
        }
    }
    double *o = emom + (size_t)24 * k;
#pragma unroll
    for (int i = 0; i < 24; ++i) o[i] = m[i];
}

template <int DIM, typename PosT>
static void dispatch_p2(ffcuda_ctx *ctx, ffcuda_matrix *A, ffcuda_space *s, const FormParams &F, const double *Rg, const PosT *pos,
                        int accumulate, const double *Tq = nullptr, const double *cq = nullptr, int nqc = 0)
{
    const int nc = s->ncomp;
    if (nc == 1) launch_p2<DIM, 1, PosT>(ctx, A, s, F, Rg, pos, accumulate, Tq, cq, nqc);
    else if (nc == 2) launch_p2<DIM, 2, PosT>(ctx, A, s, F, Rg, pos, accumulate, Tq, cq, nqc);
    else launch_p2<DIM, 3, PosT>(ctx, A, s, F, Rg, pos, accumulate, Tq, cq, nqc);
}

static int assemble_bilinear_impl(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq, const double *qpts,
                                  const double *qw, int nlab, const int32_t *labels, int accumulate, const double *cq);

extern "C" int ffcuda_assemble_bilinear(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq,
                                        const double *qpts, const double *qw, int nlab, const int32_t *labels, int accumulate)
{
    return assemble_bilinear_impl(A, s, nterms, terms, nq, qpts, qw, nlab, labels, accumulate, nullptr);
}

extern "C" int ffcuda_assemble_bilinear_qcoef(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq,
                                              const double *qpts, const double *qw, const double *cq, int accumulate)
{
    if (!cq) {
        ff_report_error(s ? s->ctx : nullptr, "ffcuda_assemble_bilinear_qcoef: null coefficient table");
        return 1;
    }
    return assemble_bilinear_impl(A, s, nterms, terms, nq, qpts, qw, 0, nullptr, accumulate, cq);
}

static int assemble_bilinear_impl(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq, const double *qpts,
                                  const double *qw, int nlab, const int32_t *labels, int accumulate, const double *cq)
{
    FF_API_BEGIN
    FF_REQUIRE(A && s && A->pattern && A->pattern->space == s, "ffcuda_assemble_bilinear: matrix was not created on this space");
    FF_REQUIRE(!cq || (!s->mesh->distributed && nq <= 256), "coefficients given at the quadrature nodes: one GPU, at most 256 nodes");
    FF_REQUIRE(nterms >= 0 && (nterms == 0 || terms), "bad term list");
    FF_REQUIRE(nq > 0 && qpts && qw, "quadrature rule missing");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    const int dim = s->mesh->dim, nloc = s->nloc, nc = s->ncomp, ns = dim + 1;
    FormParams F;
    memset(&F, 0, sizeof(F));
    for (int t = 0; t < nterms; ++t) {
        const ffcuda_bterm &T = terms[t];
        FF_REQUIRE(T.ucomp >= 0 && T.ucomp < nc && T.vcomp >= 0 && T.vcomp < nc, "term component out of range");
        int su = op_slot(dim, T.uop), sv = op_slot(dim, T.vop);
        F.C[T.vcomp][T.ucomp][sv][su] += T.coef;
    }
    for (int cv = 0; cv < nc; ++cv)
        for (int cu = 0; cu < nc; ++cu)
            for (int sv = 0; sv < ns; ++sv)
                for (int su = 0; su < ns; ++su)
                    if (F.C[cv][cu][sv][su] != 0.0) F.mask |= 1u << (sv * 4 + su);
    fill_labels(nlab, labels, F.nlab, F.labels);
    // reference tensors from the quadrature rule
    std::vector<double> R((size_t)nloc * nloc * ns * ns, 0.0);
    for (int q = 0; q < nq; ++q) {
        double B[10][4];
        ref_basis(dim, s->order, qpts + (size_t)q * dim, B);
        for (int a = 0; a < nloc; ++a)
            for (int b = 0; b < nloc; ++b)
                for (int sa = 0; sa < ns; ++sa)
                    for (int sb = 0; sb < ns; ++sb) R[(((size_t)a * nloc + b) * ns + sa) * ns + sb] += qw[q] * B[a][sa] * B[b][sb];
        if (s->order == 1) {
            F.W += qw[q];
            for (int a = 0; a < nloc; ++a) {
                F.Lh[a] += qw[q] * B[a][0];
                for (int b = 0; b < nloc; ++b) F.Mh[a][b] += qw[q] * B[a][0] * B[b][0];
            }
        }
    }
    ffcuda_pattern *P = A->pattern;
    if (accumulate) ff_matrix_touch(A);
    A->vals_stale = false;
    A->vals_epoch++;
    DBuf<double> Rg;
    if (s->order == 2) {
        Rg.alloc(R.size());
        FF_CUDA(cudaMemcpyAsync(Rg.p, R.data(), Rg.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (s->order == 1) {
        // FAST path: scalar space, c grad u . grad v (+ m u v), symmetric rule (M_aa all equal, M_ab all equal)
        bool fast = nc == 1 && (F.mask & 0x111Eu) == 0;
        const double c = F.C[0][0][1][1];
        for (int sv = 1; sv <= dim && fast; ++sv)
            for (int su = 1; su <= dim; ++su)
                if (F.C[0][0][sv][su] != (sv == su ? c : 0.0)) fast = false;
        for (int a = 0; a < nloc && fast && (F.mask & 1u); ++a)
            for (int b = 0; b < nloc; ++b) {
                const double ref = a == b ? F.Mh[0][0] : F.Mh[0][1];
                if (fabs(F.Mh[a][b] - ref) > 1e-15 * fabs(F.Mh[0][0])) fast = false;
            }
        if (fast) {
            const double rfac = dim == 3 ? 1.0 / 6.0 : 0.5;
            F.fast_cw = c * F.W * rfac;
            F.fast_md = F.C[0][0][0][0] * F.Mh[0][0] * rfac;
            F.fast_mo = F.C[0][0][0][0] * F.Mh[0][1] * rfac;
        }
        DBuf<double> emom;
        if (cq) { // moments of the coefficient on every element
            const int nt = s->mesh->nt;
            std::vector<double> wl((size_t)nq * 5, 0.0);
            for (int q = 0; q < nq; ++q) {
                double B[10][4];
                ref_basis(dim, 1, qpts + (size_t)q * dim, B);
                wl[(size_t)q * 5] = qw[q];
                for (int a = 0; a <= dim; ++a) wl[(size_t)q * 5 + 1 + a] = B[a][0];
            }
            DBuf<double> dwl, dcq;
            dwl.alloc(wl.size());
            emom.alloc((size_t)nt * 24);
            FF_CUDA(cudaMemcpyAsync(dwl.p, wl.data(), dwl.bytes(), cudaMemcpyHostToDevice, ctx->stream));
            const double *pcq = table_on_device(cq, (size_t)nt * nq, dcq, ctx->stream);
            ff_launch(ctx, "asm_coef_moments", [&] {
                k_emom<<<ff_blocks(nt, 128), 128, wl.size() * sizeof(double), ctx->stream>>>(nt, nq, dim + 1, dwl.p, pcq, emom.p);
            });
            FF_CUDA(cudaStreamSynchronize(ctx->stream)); // cq is the caller's pageable memory
            fast = false;
        }
        if (dim == 3) dispatch_p1<3>(ctx, A, s, F, accumulate, fast, emom.p);
        else dispatch_p1<2>(ctx, A, s, F, accumulate, fast, emom.p);
    } else {
        // P2 with a coefficient at the quadrature nodes: the table w_q d^sa phi_a(q) d^sb phi_b(q) and the values go to the device
        DBuf<double> dT, dcq;
        const double *pcq = nullptr;
        if (cq) {
            std::vector<double> T((size_t)nq * nloc * nloc * ns * ns);
            for (int q = 0; q < nq; ++q) {
                double B[10][4];
                ref_basis(dim, 2, qpts + (size_t)q * dim, B);
                for (int a = 0; a < nloc; ++a)
                    for (int b = 0; b < nloc; ++b)
                        for (int sa = 0; sa < ns; ++sa)
                            for (int sb = 0; sb < ns; ++sb)
                                T[((((size_t)q * nloc + a) * nloc + b) * ns + sa) * ns + sb] = qw[q] * B[a][sa] * B[b][sb];
            }
            dT.alloc(T.size());
            FF_CUDA(cudaMemcpyAsync(dT.p, T.data(), dT.bytes(), cudaMemcpyHostToDevice, ctx->stream));
            pcq = table_on_device(cq, (size_t)s->mesh->nt * nq, dcq, ctx->stream);
            FF_CUDA(cudaStreamSynchronize(ctx->stream)); // both sources are pageable host memory
        }
        if (P->pos8.p) {
            if (dim == 3) dispatch_p2<3, uint8_t>(ctx, A, s, F, Rg.p, P->pos8.p, accumulate, dT.p, pcq, nq);
            else dispatch_p2<2, uint8_t>(ctx, A, s, F, Rg.p, P->pos8.p, accumulate, dT.p, pcq, nq);
        } else {
            if (dim == 3) dispatch_p2<3, uint16_t>(ctx, A, s, F, Rg.p, P->pos16.p, accumulate, dT.p, pcq, nq);
            else dispatch_p2<2, uint16_t>(ctx, A, s, F, Rg.p, P->pos16.p, accumulate, dT.p, pcq, nq);
        }
    }
    // Rg is released through the stream-ordered allocator: no synchronisation needed
    FF_API_END(s ? s->ctx : nullptr)
}

template <int DIM>
static void launch_rhs(ffcuda_ctx *ctx, ffcuda_vec *b, ffcuda_space *s, const LinParams &Lp, const double *Fh, const double *hFh,
                       int hasgrad, bool lean, int accumulate)
{
    ffcuda_mesh *m = s->mesh;
    const int nrows = s->incidence.nrows;
    const IncView V = ff_view(s->incidence);
    int blocks = ff_blocks((size_t)((nrows + 31) / 32) * 32, 128);
    const Incidence &I = s->incidence;
    const int SV = std::max(4, (I.maxstage + 3) & ~3);
    const size_t shmem = I.ell ? (size_t)4 * DIM * SV * 8 : 0;
    if (lean) {
        const double Lval = hFh[0], Wsum = -hFh[1];
        {   // row tiles (tiles.cu) when the space has them: every element evaluated once per tile
            constexpr double RFAC = DIM == 3 ? 1.0 / 6.0 : 0.5;
            double cval[3] = {0, 0, 0}, cgrad[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = 0; c < s->ncomp; ++c) {
                cval[c] = Lp.CL[c][0] * Lval * RFAC;
                for (int x = 0; x < DIM; ++x) cgrad[c * 3 + x] = Lp.CL[c][x + 1] * Wsum * RFAC;
            }
            if (ff_rhs_p1_tiles(ctx, b, s, cval, cgrad, hasgrad, accumulate)) return;
        }
        ff_launch(ctx, "rhs_rows", [&] {
#define FF_RHS_LEAN(NCC)                                                                                                          \
    if (hasgrad)                                                                                                                  \
        k_rhs_p1_lean<DIM, NCC, true><<<blocks, 128, shmem, ctx->stream>>>(m->xyz.p, nrows, I.cnt.p, I.blkoff.p, I.loc.p, I.blkvert.p, \
                                                                           I.blkvcnt.p, SV, Lval, Wsum, b->d.p, accumulate, Lp);  \
    else                                                                                                                          \
        k_rhs_p1_lean<DIM, NCC, false><<<blocks, 128, shmem, ctx->stream>>>(m->xyz.p, nrows, I.cnt.p, I.blkoff.p, I.loc.p, I.blkvert.p, \
                                                                            I.blkvcnt.p, SV, Lval, Wsum, b->d.p, accumulate, Lp);
            if (s->ncomp == 1) { FF_RHS_LEAN(1) }
            else if (s->ncomp == 2) { FF_RHS_LEAN(2) }
            else { FF_RHS_LEAN(3) }
#undef FF_RHS_LEAN
        });
        return;
    }
    ff_launch(ctx, "rhs_rows", [&] {
        if (s->ncomp == 1)
            k_rhs<DIM, 1><<<blocks, 128, shmem, ctx->stream>>>(m->xyz.p, m->conn.p, m->elab.p, nrows, V, I.loc.p, I.blkvert.p, I.blkvcnt.p, SV, Fh, s->nloc, hasgrad, b->d.p, accumulate, Lp);
        else if (s->ncomp == 2)
            k_rhs<DIM, 2><<<blocks, 128, shmem, ctx->stream>>>(m->xyz.p, m->conn.p, m->elab.p, nrows, V, I.loc.p, I.blkvert.p, I.blkvcnt.p, SV, Fh, s->nloc, hasgrad, b->d.p, accumulate, Lp);
        else
            k_rhs<DIM, 3><<<blocks, 128, shmem, ctx->stream>>>(m->xyz.p, m->conn.p, m->elab.p, nrows, V, I.loc.p, I.blkvert.p, I.blkvcnt.p, SV, Fh, s->nloc, hasgrad, b->d.p, accumulate, Lp);
    });
}

extern "C" int ffcuda_assemble_linear(ffcuda_vec *b, ffcuda_space *s, int nterms, const ffcuda_lterm *terms, int nq,
                                      const double *qpts, const double *qw, int nlab, const int32_t *labels, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(b && s, "ffcuda_assemble_linear: null argument");
    FF_REQUIRE(nq > 0 && qpts && qw, "quadrature rule missing");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ff_build_incidence(s); // the row-owner gather walks the node -> element lists of the space
    FF_REQUIRE(b->n >= s->nnodes_owned * s->ncomp, "right-hand side vector too short");
    const int dim = s->mesh->dim, nloc = s->nloc, nc = s->ncomp, ns = dim + 1;
    LinParams Lp;
    memset(&Lp, 0, sizeof(Lp));
    int hasgrad = 0;
    for (int t = 0; t < nterms; ++t) {
        FF_REQUIRE(terms[t].vcomp >= 0 && terms[t].vcomp < nc, "term component out of range");
        const int slot = op_slot(dim, terms[t].vop);
        Lp.CL[terms[t].vcomp][slot] += terms[t].coef;
        if (slot > 0 && terms[t].coef != 0.0) hasgrad = 1;
    }
    const int hasgrad_form = hasgrad;
    if (dim == 2) hasgrad = 1; // the generic 2-D geometry helper always evaluates the gradients
    fill_labels(nlab, labels, Lp.nlab, Lp.labels);
    std::vector<double> Fh((size_t)nloc * ns, 0.0);
    for (int q = 0; q < nq; ++q) {
        double B[10][4];
        ref_basis(dim, s->order, qpts + (size_t)q * dim, B);
        for (int a = 0; a < nloc; ++a)
            for (int sa = 0; sa < ns; ++sa) Fh[(size_t)a * ns + sa] += qw[q] * B[a][sa];
    }
    DBuf<double> dF;
    dF.alloc(Fh.size());
    FF_CUDA(cudaMemcpyAsync(dF.p, Fh.data(), dF.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    // (a pageable source is staged before cudaMemcpyAsync returns: Fh may go out of scope)
    // lean path: P1, no region filter, all blocks staged, symmetric rule (sum_q w_q lambda_a the same for every a)
    bool lean = s->order == 1 && Lp.nlab < 0 && s->incidence.ell && s->incidence.nunstaged == 0;
    for (int a = 1; a < nloc && lean; ++a)
        if (fabs(Fh[(size_t)a * ns] - Fh[0]) > 1e-15 * fabs(Fh[0])) lean = false;
    if (dim == 3) launch_rhs<3>(ctx, b, s, Lp, dF.p, Fh.data(), lean ? hasgrad_form : hasgrad, lean, accumulate);
    else launch_rhs<2>(ctx, b, s, Lp, dF.p, Fh.data(), lean ? hasgrad_form : hasgrad, lean, accumulate);
    FF_API_END(s ? s->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// Boundary integrals of a linear form: int2d(Th3, labels)(c v) / int1d(Th, labels)(c v) - Neumann / traction data.
// Element_rhs on border elements (fflib/problem.cpp:8439-8513 2-D, :8517-8587 3-D): for every boundary element with a
// listed label, B[dof] += measure(face) * c * sum_q w_q phi_dof(PBord(face, q)).  Value terms only (vop = id), for which
// only the nodes lying on the face receive something.
// Deterministic without atomics on doubles: the node -> boundary-element incidence (a property of the space, built
// once: integer count / scan / fill, lists sorted by boundary element) and one thread per node adding its list in order.
// ----------------------------------------------------------------------------------------------------
static constexpr int MAXBL = 32;
struct BndParams {
    double Fb[4][10];  // [face of the element][local node]: sum_q w_q phi(PBord(face, q))
    double coef[3];    // per component
    int nlab;          // < 0: every boundary element
    int labels[MAXBL];
};

// local nodes of the element that lie on its face f (opposite vertex f): P1: the dim vertices; P2: + the edges of the face
template <int DIM>
__device__ __forceinline__ int face_nodes(int order, int f, int (&out)[6])
{
    int n = 0;
    for (int a = 0; a <= DIM; ++a)
        if (a != f) out[n++] = a;
    if (order == 2) {
        if (DIM == 3) {
            const int e3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
            for (int e = 0; e < 6; ++e)
                if (e3[e][0] != f && e3[e][1] != f) out[n++] = 4 + e;
        } else {
            out[n++] = 3 + f; // dof 3+e lies on the edge opposite vertex e
        }
    }
    return n;
}

// PASS 0: count the items of every node; PASS 1: fill (item = boundary element << 4 | local node)
template <int DIM, int PASS>
__global__ void k_bnd_items(int nbe, const int32_t *__restrict__ belem, const int32_t *__restrict__ bface,
                            const int32_t *__restrict__ e2n, int nloc, int order, int nrows, int32_t *__restrict__ cntptr,
                            int32_t *__restrict__ cursor, uint32_t *__restrict__ items, int allnodes)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nbe) return;
    int loc[10];
    int n;
    if (allnodes) { // every node of the adjacent element: terms with derivatives reach the nodes off the face too
        n = nloc;
        for (int j = 0; j < nloc; ++j) loc[j] = j;
    } else {
        int fl[6];
        n = face_nodes<DIM>(order, bface[e], fl);
        for (int j = 0; j < n; ++j) loc[j] = fl[j];
    }
    const int32_t *N = e2n + (size_t)nloc * belem[e];
    for (int j = 0; j < n; ++j) {
        const int node = N[loc[j]];
        if (node >= nrows) continue;
        if (PASS == 0) atomicAdd(&cntptr[node], 1);
        else items[cntptr[node] + atomicAdd(&cursor[node], 1)] = ((uint32_t)e << 4) | (uint32_t)loc[j];
    }
}
__global__ void k_bnd_sort(const int32_t *__restrict__ ptr, uint32_t *__restrict__ items, int nrows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int b = ptr[i], n = ptr[i + 1] - b;
    for (int x = 1; x < n; ++x) {
        const uint32_t v = items[b + x];
        int y = x - 1;
        while (y >= 0 && items[b + y] > v) {
            items[b + y + 1] = items[b + y];
            --y;
        }
        items[b + y + 1] = v;
    }
}
// measure of every boundary element whose label is listed (0 otherwise): |N|/2 of the face, length of the edge
template <int DIM>
__global__ void k_bnd_measure(int nbe, const int32_t *__restrict__ belem, const int32_t *__restrict__ bface,
                              const int32_t *__restrict__ blab, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                              int vstride, const __grid_constant__ BndParams Bp, double *__restrict__ meas)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nbe) return;
    bool ok = Bp.nlab < 0;
    for (int i = 0; i < Bp.nlab; ++i) ok |= (Bp.labels[i] == blab[e]);
    double m = 0.0;
    if (ok) {
        const int32_t *K = conn + (size_t)(DIM + 1) * belem[e];
        const int f = bface[e];
        int v[DIM], n = 0;
        for (int a = 0; a <= DIM; ++a)
            if (a != f) v[n++] = K[a];
        const double *A = xyz + (size_t)v[0] * vstride, *B = xyz + (size_t)v[1] * vstride;
        if (DIM == 3) {
            const double *Cc = xyz + (size_t)v[DIM - 1] * vstride;
            const double ux = B[0] - A[0], uy = B[1] - A[1], uz = B[DIM - 1] - A[DIM - 1];
            const double wx = Cc[0] - A[0], wy = Cc[1] - A[1], wz = Cc[DIM - 1] - A[DIM - 1];
            const double nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
            m = 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
        } else {
            m = sqrt((B[0] - A[0]) * (B[0] - A[0]) + (B[1] - A[1]) * (B[1] - A[1]));
        }
    }
    meas[e] = m;
}
__global__ void k_bnd_gather(int nrows, const int32_t *__restrict__ ptr, const uint32_t *__restrict__ items,
                             const int32_t *__restrict__ bface, const double *__restrict__ meas, int nc,
                             const __grid_constant__ BndParams Bp, double *__restrict__ b, int accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    double s = 0.0;
    for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
        const uint32_t it = items[k];
        const int e = it >> 4, a = it & 15;
        s += meas[e] * Bp.Fb[bface[e]][a];
    }
    if (!accumulate || ptr[i + 1] > ptr[i])
        for (int c = 0; c < nc; ++c) {
            double *dst = b + (size_t)i * nc + c;
            *dst = accumulate ? *dst + Bp.coef[c] * s : Bp.coef[c] * s;
        }
}

template <int DIM>
static void bnd_incidence(ffcuda_ctx *ctx, ffcuda_space *s, bool allnodes = false)
{
    DBuf<int32_t> &bptr = allnodes ? s->bnd2_ptr : s->bnd_ptr;
    DBuf<uint32_t> &bitems = allnodes ? s->bnd2_items : s->bnd_items;
    if (bptr.p) return;
    ffcuda_mesh *m = s->mesh;
    cudaStream_t st = ctx->stream;
    const int nrows = s->nnodes_owned, nbe = m->nbe;
    DBuf<int32_t> cnt, cursor;
    cnt.alloc((size_t)nrows + 1);
    cursor.alloc((size_t)nrows + 1);
    FF_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
    FF_CUDA(cudaMemsetAsync(cursor.p, 0, cursor.bytes(), st));
    ff_launch(ctx, "bnd_count", [&] {
        k_bnd_items<DIM, 0><<<ff_blocks(nbe, 256), 256, 0, st>>>(nbe, m->belem.p, m->bface.p, s->e2n, s->nloc, s->order, nrows, cnt.p, nullptr, nullptr,
                                                                (int)allnodes);
    });
    bptr.alloc((size_t)nrows + 1);
    int64_t tot = 0;
    ff_exclusive_scan_i32(ctx, cnt.p, bptr.p, (size_t)nrows + 1, &tot);
    bitems.alloc((size_t)std::max<int64_t>(tot, 1));
    ff_launch(ctx, "bnd_fill", [&] {
        k_bnd_items<DIM, 1><<<ff_blocks(nbe, 256), 256, 0, st>>>(nbe, m->belem.p, m->bface.p, s->e2n, s->nloc, s->order, nrows, bptr.p, cursor.p,
                                                                bitems.p, (int)allnodes);
    });
    ff_launch(ctx, "bnd_sort", [&] { k_bnd_sort<<<ff_blocks(nrows, 256), 256, 0, st>>>(bptr.p, bitems.p, nrows); });
}

// basis values at the point of face f of the reference element that the face rule's node qp maps to:
// PBord, femlib/Mesh3dn.hpp:76 (faces nvfaceTet), Mesh2dn.hpp:65 (edges nvedgeTria)
static void face_ref_basis(int dim, int order, int f, const double *qp, double B[10][4])
{
    static const int nvface[4][3] = {{3, 2, 1}, {0, 2, 3}, {3, 1, 0}, {0, 1, 2}};
    static const int nvedge[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    static const double hat3[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static const double hat2[3][2] = {{0, 0}, {1, 0}, {0, 1}};
    double P[3] = {0, 0, 0};
    if (dim == 3) {
        const double x = qp[0], y = qp[1];
        for (int d = 0; d < 3; ++d) P[d] = hat3[nvface[f][0]][d] * (1 - x - y) + hat3[nvface[f][1]][d] * x + hat3[nvface[f][2]][d] * y;
    } else {
        const double x = qp[0];
        for (int d = 0; d < 2; ++d) P[d] = hat2[nvedge[f][0]][d] * (1 - x) + hat2[nvedge[f][1]][d] * x;
    }
    ref_basis(dim, order, P, B);
}
static void bnd_fill_labels(BndParams &Bp, int nlab, const int32_t *labels)
{
    if (!labels) Bp.nlab = -1;
    else {
        FF_REQUIRE(nlab <= MAXBL, "at most 32 boundary labels per integral");
        Bp.nlab = nlab;
        for (int i = 0; i < nlab; ++i) Bp.labels[i] = labels[i];
    }
}
static void bnd_measures(ffcuda_ctx *ctx, ffcuda_mesh *m, const BndParams &Bp, double *meas)
{
    cudaStream_t st = ctx->stream;
    ff_launch(ctx, "bnd_measure", [&] {
        if (m->dim == 3)
            k_bnd_measure<3><<<ff_blocks(m->nbe, 256), 256, 0, st>>>(m->nbe, m->belem.p, m->bface.p, m->blab.p, m->conn.p, m->xyz.p, m->vstride, Bp, meas);
        else
            k_bnd_measure<2><<<ff_blocks(m->nbe, 256), 256, 0, st>>>(m->nbe, m->belem.p, m->bface.p, m->blab.p, m->conn.p, m->xyz.p, m->vstride, Bp, meas);
    });
}

// ----------------------------------------------------------------------------------------------------
// Boundary integrals with DERIVATIVES of the unknown or of the test function (int2d(Th,lab)(dx(u) v), Nitsche terms,
// int2d(Th,lab)(c dz(v))): the border branch of Element_Op / Element_rhs evaluates the basis functions of the adjacent
// element at the face quadrature points with whatever operators the terms ask for (fflib/problem.cpp:6518-6560,
// :8517-8587), so every node of that element receives something, not only those lying on the face.  Same ownership as the
// value terms: one thread per node row walks its items — here taken from the incidence over ALL nodes of the adjacent
// elements (bnd2) — in order, evaluates its own basis function and those of the element's nodes at every face point
// (gradients of the barycentric coordinates from the vertices, P2 by the product rule) and adds into its row, columns found
// by bisection.  O(boundary) work: clarity over speed.
// ----------------------------------------------------------------------------------------------------
struct BndGenTerm {
    int ucomp, uslot, vcomp, vslot; // slot: 0 value, 1..3 d/dx, d/dy, d/dz
    double coef;
};
static constexpr int MAXBT = 48;
struct BndGenParams {
    int nterms, nq;
    BndGenTerm t[MAXBT];
    double w[16];        // weights of the face rule
    double lam[4][16][4]; // [face][q][vertex]: barycentric coordinates, in the element, of face point q (PBord)
};

// gradients of the barycentric coordinates of element K
template <int DIM>
__device__ __forceinline__ void bary_gradients(const int32_t *__restrict__ K, const double *__restrict__ xyz, int vstride, double (&G)[DIM + 1][DIM])
{
    double X[DIM + 1][DIM];
    for (int a = 0; a <= DIM; ++a)
        for (int d = 0; d < DIM; ++d) X[a][d] = xyz[(size_t)K[a] * vstride + d];
    if (DIM == 2) {
        const double ax = X[1][0] - X[0][0], ay = X[1][1] - X[0][1], bx = X[2][0] - X[0][0], by = X[2][1] - X[0][1];
        const double det = ax * by - ay * bx;
        G[1][0] = by / det;  G[1][1] = -bx / det;
        G[2][0] = -ay / det; G[2][1] = ax / det;
    } else {
        double e[3][3];
        for (int r = 0; r < 3; ++r)
            for (int d = 0; d < 3; ++d) e[r][d] = X[r + 1][d] - X[0][d];
        double c[3][3]; // c[r] = e[r+1] x e[r+2]
        for (int r = 0; r < 3; ++r) {
            const double *u = e[(r + 1) % 3], *v = e[(r + 2) % 3];
            c[r][0] = u[1] * v[2] - u[2] * v[1];
            c[r][1] = u[2] * v[0] - u[0] * v[2];
            c[r][2] = u[0] * v[1] - u[1] * v[0];
        }
        const double det = e[0][0] * c[0][0] + e[0][1] * c[0][1] + e[0][2] * c[0][2];
        for (int r = 0; r < 3; ++r)
            for (int d = 0; d < DIM; ++d) G[r + 1][d] = c[r][d] / det;
    }
    for (int d = 0; d < DIM; ++d) {
        double sgrad = 0.0;
        for (int a = 1; a <= DIM; ++a) sgrad += G[a][d];
        G[0][d] = -sgrad;
    }
}

// value and gradient of the basis function of local node a at the point with barycentric coordinates l: out[0] value,
// out[1..DIM] derivatives.  P2 edges: {01,02,03,12,13,23} on a tetrahedron, edge opposite vertex e on a triangle
// (femlib/P012_3d.cpp:199-300, femlib/FESpace.cpp:1219)
template <int DIM>
__device__ __forceinline__ void basis_at(int order, int a, const double *__restrict__ l, const double (&G)[DIM + 1][DIM], double (&out)[4])
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
        const int e3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
        p = e3[a - 4][0];
        r = e3[a - 4][1];
    } else {
        p = (a - 3 + 1) % 3;
        r = (a - 3 + 2) % 3;
    }
    out[0] = 4.0 * l[p] * l[r];
    for (int d = 0; d < DIM; ++d) out[1 + d] = 4.0 * (l[p] * G[r][d] + l[r] * G[p][d]);
}

template <int DIM>
__global__ void k_bnd_bilinear_gen(int nrows, const int32_t *__restrict__ ptr, const uint32_t *__restrict__ items,
                                   const int32_t *__restrict__ belem, const int32_t *__restrict__ bface, const double *__restrict__ meas,
                                   const int32_t *__restrict__ conn, const double *__restrict__ xyz, int vstride,
                                   const int32_t *__restrict__ e2n, int nloc, int order, int nc, const int32_t *__restrict__ nrowptr,
                                   const int32_t *__restrict__ ncol, const BndGenParams *__restrict__ Gp, const double *__restrict__ cq,
                                   double *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int k0 = ptr[i], k1 = ptr[i + 1];
    if (k0 == k1) return;
    const int rb = nrowptr[i], L = nrowptr[i + 1] - rb;
    double *row = vals + (size_t)nc * nc * rb;
    const int nq = Gp->nq, nterms = Gp->nterms;
    for (int k = k0; k < k1; ++k) {
        const uint32_t it = items[k];
        const int e = it >> 4, a = it & 15;
        const double m = meas[e];
        if (m == 0.0) continue; // label not listed
        const int f = bface[e], el = belem[e];
        const int32_t *N = e2n + (size_t)nloc * el;
        double G[DIM + 1][DIM];
        bary_gradients<DIM>(conn + (size_t)(DIM + 1) * el, xyz, vstride, G);
        for (int b = 0; b < nloc; ++b) {
            double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int q = 0; q < nq; ++q) {
                double va[4], ub[4];
                basis_at<DIM>(order, a, Gp->lam[f][q], G, va);
                basis_at<DIM>(order, b, Gp->lam[f][q], G, ub);
                const double w = m * Gp->w[q] * (cq ? cq[(size_t)e * nq + q] : 1.0);
                for (int t = 0; t < nterms; ++t) {
                    const BndGenTerm &T = Gp->t[t];
                    acc[T.vcomp][T.ucomp] += (T.coef * w) * va[T.vslot] * ub[T.uslot];
                }
            }
            const int j = N[b];
            int lo = 0, hi = L - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ncol[rb + mid] < j) lo = mid + 1;
                else hi = mid;
            }
            for (int cv = 0; cv < nc; ++cv)
                for (int cu = 0; cu < nc; ++cu)
                    if (acc[cv][cu] != 0.0) row[(size_t)cv * nc * L + (size_t)lo * nc + cu] += acc[cv][cu];
        }
    }
}

template <int DIM>
__global__ void k_bnd_linear_gen(int nrows, const int32_t *__restrict__ ptr, const uint32_t *__restrict__ items,
                                 const int32_t *__restrict__ belem, const int32_t *__restrict__ bface, const double *__restrict__ meas,
                                 const int32_t *__restrict__ conn, const double *__restrict__ xyz, int vstride, int order, int nc,
                                 const BndGenParams *__restrict__ Gp, double *__restrict__ bvec, int accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    double s[3] = {0, 0, 0};
    const int nq = Gp->nq, nterms = Gp->nterms;
    for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
        const uint32_t it = items[k];
        const int e = it >> 4, a = it & 15;
        const double m = meas[e];
        if (m == 0.0) continue;
        const int f = bface[e], el = belem[e];
        double G[DIM + 1][DIM];
        bary_gradients<DIM>(conn + (size_t)(DIM + 1) * el, xyz, vstride, G);
        for (int q = 0; q < nq; ++q) {
            double va[4];
            basis_at<DIM>(order, a, Gp->lam[f][q], G, va);
            const double w = m * Gp->w[q];
            for (int t = 0; t < nterms; ++t) {
                const BndGenTerm &T = Gp->t[t];
                s[T.vcomp] += (T.coef * w) * va[T.vslot];
            }
        }
    }
    if (!accumulate || ptr[i + 1] > ptr[i])
        for (int c = 0; c < nc; ++c) {
            double *dst = bvec + (size_t)i * nc + c;
            *dst = accumulate ? *dst + s[c] : s[c];
        }
}

static int op_slot(int op)
{
    switch (op) {
    case FFCUDA_OP_ID: return 0;
    case FFCUDA_OP_DX: return 1;
    case FFCUDA_OP_DY: return 2;
    case FFCUDA_OP_DZ: return 3;
    }
    throw FFError("boundary integrals: operator code not supported (value and first derivatives only)");
}

// face rule -> barycentric coordinates in the element (PBord: femlib/Mesh3dn.hpp:76 faces nvfaceTet, Mesh2dn.hpp:65 edges)
static void bnd_gen_rule(BndGenParams &Gp, int dim, int nq, const double *qpts, const double *qw)
{
    static const int nvface[4][3] = {{3, 2, 1}, {0, 2, 3}, {3, 1, 0}, {0, 1, 2}};
    static const int nvedge[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    FF_REQUIRE(nq <= 16, "boundary integrals with derivatives: at most 16 quadrature points on a face");
    Gp.nq = nq;
    for (int q = 0; q < nq; ++q) Gp.w[q] = qw[q];
    for (int f = 0; f <= dim; ++f)
        for (int q = 0; q < nq; ++q) {
            for (int a = 0; a < 4; ++a) Gp.lam[f][q][a] = 0.0;
            if (dim == 3) {
                const double x = qpts[2 * q], y = qpts[2 * q + 1];
                Gp.lam[f][q][nvface[f][0]] = 1 - x - y;
                Gp.lam[f][q][nvface[f][1]] = x;
                Gp.lam[f][q][nvface[f][2]] = y;
            } else {
                const double x = qpts[q];
                Gp.lam[f][q][nvedge[f][0]] = 1 - x;
                Gp.lam[f][q][nvedge[f][1]] = x;
            }
        }
}

static bool bnd_has_derivative(int nterms, const ffcuda_bterm *terms)
{
    for (int t = 0; t < nterms; ++t)
        if (terms[t].uop != FFCUDA_OP_ID || terms[t].vop != FFCUDA_OP_ID) return true;
    return false;
}

// the general bilinear path: constant coefficients (cq == nullptr) or one coefficient table cq[e * nq + q] (host)
static void bnd_bilinear_general(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq, const double *qpts,
                                 const double *qw, const double *cq, int nlab, const int32_t *labels, int accumulate)
{
    ffcuda_ctx *ctx = s->ctx;
    ffcuda_mesh *m = s->mesh;
    ffcuda_pattern *P = A->pattern;
    cudaStream_t st = ctx->stream;
    const int dim = m->dim, nc = s->ncomp, nbe = m->nbe;
    FF_REQUIRE(nterms <= MAXBT, "boundary integrals with derivatives: at most 48 terms per integral");
    FF_REQUIRE(dim == 2 || dim == 3, "bad dimension");
    std::unique_ptr<BndGenParams> Gp(new BndGenParams());
    memset(Gp.get(), 0, sizeof(BndGenParams));
    bnd_gen_rule(*Gp, dim, nq, qpts, qw);
    Gp->nterms = nterms;
    for (int t = 0; t < nterms; ++t) {
        const ffcuda_bterm &T = terms[t];
        FF_REQUIRE(T.ucomp >= 0 && T.ucomp < nc && T.vcomp >= 0 && T.vcomp < nc, "term component out of range");
        Gp->t[t] = BndGenTerm{T.ucomp, op_slot(T.uop), T.vcomp, op_slot(T.vop), T.coef};
        FF_REQUIRE(dim == 3 || (Gp->t[t].uslot < 3 && Gp->t[t].vslot < 3), "dz on a 2-D mesh");
    }
    BndParams Bp;
    memset(&Bp, 0, sizeof(Bp));
    bnd_fill_labels(Bp, nlab, labels);
    if (accumulate) ff_matrix_touch(A);
    else FF_CUDA(cudaMemsetAsync(A->vals.p, 0, A->vals.bytes(), st));
    A->vals_stale = false;
    A->vals_epoch++;
    if (dim == 3) bnd_incidence<3>(ctx, s, true);
    else bnd_incidence<2>(ctx, s, true);
    DBuf<double> meas, dC;
    DBuf<BndGenParams> dG;
    meas.alloc((size_t)nbe);
    dG.alloc(1);
    FF_CUDA(cudaMemcpyAsync(dG.p, Gp.get(), sizeof(BndGenParams), cudaMemcpyHostToDevice, st));
    const double *pC = cq ? table_on_device(cq, (size_t)nbe * nq, dC, st) : nullptr;
    bnd_measures(ctx, m, Bp, meas.p);
    const int nrows = s->nnodes_owned;
    ff_launch(ctx, "bnd_bilinear_gen", [&] {
        if (dim == 3)
            k_bnd_bilinear_gen<3><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd2_ptr.p, s->bnd2_items.p, m->belem.p, m->bface.p, meas.p,
                                                                       m->conn.p, m->xyz.p, m->vstride, s->e2n, s->nloc, s->order, nc,
                                                                       P->nrowptr.p, P->ncol.p, dG.p, pC, A->vals.p);
        else
            k_bnd_bilinear_gen<2><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd2_ptr.p, s->bnd2_items.p, m->belem.p, m->bface.p, meas.p,
                                                                       m->conn.p, m->xyz.p, m->vstride, s->e2n, s->nloc, s->order, nc,
                                                                       P->nrowptr.p, P->ncol.p, dG.p, pC, A->vals.p);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // host buffers (Gp, cq) are the caller's / ours on the stack
}

static void bnd_linear_general(ffcuda_vec *b, ffcuda_space *s, int nterms, const ffcuda_lterm *terms, int nq, const double *qpts,
                               const double *qw, int nlab, const int32_t *labels, int accumulate)
{
    ffcuda_ctx *ctx = s->ctx;
    ffcuda_mesh *m = s->mesh;
    cudaStream_t st = ctx->stream;
    const int dim = m->dim, nc = s->ncomp, nbe = m->nbe;
    FF_REQUIRE(nterms <= MAXBT, "boundary integrals with derivatives: at most 48 terms per integral");
    std::unique_ptr<BndGenParams> Gp(new BndGenParams());
    memset(Gp.get(), 0, sizeof(BndGenParams));
    bnd_gen_rule(*Gp, dim, nq, qpts, qw);
    Gp->nterms = nterms;
    for (int t = 0; t < nterms; ++t) {
        FF_REQUIRE(terms[t].vcomp >= 0 && terms[t].vcomp < nc, "term component out of range");
        Gp->t[t] = BndGenTerm{0, 0, terms[t].vcomp, op_slot(terms[t].vop), terms[t].coef};
        FF_REQUIRE(dim == 3 || Gp->t[t].vslot < 3, "dz on a 2-D mesh");
    }
    BndParams Bp;
    memset(&Bp, 0, sizeof(Bp));
    bnd_fill_labels(Bp, nlab, labels);
    if (dim == 3) bnd_incidence<3>(ctx, s, true);
    else bnd_incidence<2>(ctx, s, true);
    DBuf<double> meas;
    DBuf<BndGenParams> dG;
    meas.alloc((size_t)nbe);
    dG.alloc(1);
    FF_CUDA(cudaMemcpyAsync(dG.p, Gp.get(), sizeof(BndGenParams), cudaMemcpyHostToDevice, st));
    bnd_measures(ctx, m, Bp, meas.p);
    const int nrows = s->nnodes_owned;
    ff_launch(ctx, "bnd_linear_gen", [&] {
        if (dim == 3)
            k_bnd_linear_gen<3><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd2_ptr.p, s->bnd2_items.p, m->belem.p, m->bface.p, meas.p,
                                                                     m->conn.p, m->xyz.p, m->vstride, s->order, nc, dG.p, b->d.p, accumulate);
        else
            k_bnd_linear_gen<2><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd2_ptr.p, s->bnd2_items.p, m->belem.p, m->bface.p, meas.p,
                                                                     m->conn.p, m->xyz.p, m->vstride, s->order, nc, dG.p, b->d.p, accumulate);
    });
    FF_CUDA(cudaStreamSynchronize(st));
}

extern "C" int ffcuda_assemble_linear_boundary(ffcuda_vec *b, ffcuda_space *s, int nterms, const ffcuda_lterm *terms, int nq,
                                               const double *qpts, const double *qw, int nlab, const int32_t *labels, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(b && s, "ffcuda_assemble_linear_boundary: null argument");
    FF_REQUIRE(nq > 0 && qpts && qw, "face quadrature rule missing");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ffcuda_mesh *m = s->mesh;
    const int dim = m->dim, nloc = s->nloc, nc = s->ncomp;
    FF_REQUIRE(b->n >= s->nnodes_owned * nc, "right-hand side vector too short");
    FF_REQUIRE(m->nbe > 0 && m->belem.p && m->bface.p, "the mesh has no boundary elements");
    FF_REQUIRE(nterms >= 0 && (nterms == 0 || terms), "bad term list");
    bool derivative = false;
    for (int t = 0; t < nterms; ++t) derivative = derivative || terms[t].vop != FFCUDA_OP_ID;
    if (derivative) { // c * dx(v) ...: every node of the adjacent element receives something
        bnd_linear_general(b, s, nterms, terms, nq, qpts, qw, nlab, labels, accumulate);
        return 0;
    }
    BndParams Bp;
    memset(&Bp, 0, sizeof(Bp));
    for (int t = 0; t < nterms; ++t) {
        FF_REQUIRE(terms[t].vcomp >= 0 && terms[t].vcomp < nc, "term component out of range");
        Bp.coef[terms[t].vcomp] += terms[t].coef;
    }
    bnd_fill_labels(Bp, nlab, labels);
    // Fb[f][a] = sum_q w_q phi_a(PBord(f, q))
    for (int f = 0; f <= dim; ++f)
        for (int q = 0; q < nq; ++q) {
            double B[10][4];
            face_ref_basis(dim, s->order, f, qpts + (size_t)q * (dim - 1), B);
            for (int a = 0; a < nloc; ++a) Bp.Fb[f][a] += qw[q] * B[a][0];
        }
    cudaStream_t st = ctx->stream;
    if (dim == 3) bnd_incidence<3>(ctx, s);
    else bnd_incidence<2>(ctx, s);
    DBuf<double> meas;
    meas.alloc((size_t)m->nbe);
    bnd_measures(ctx, m, Bp, meas.p);
    const int nrows = s->nnodes_owned;
    ff_launch(ctx, "bnd_gather", [&] {
        k_bnd_gather<<<ff_blocks(nrows, 256), 256, 0, st>>>(nrows, s->bnd_ptr.p, s->bnd_items.p, m->bface.p, meas.p, nc, Bp, b->d.p, accumulate);
    });
    FF_API_END(s ? s->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// Boundary integrals of a bilinear form (Robin terms): int2d(Th3, labels)(c u v) / int1d(Th, labels)(c u v).
// AssembleBilinearForm's loop over the border elements (fflib/problem.cpp:1317-1326, 2-D :1030-1040) hands the ADJACENT
// element to Element_Op's border branch (:6518-6560, 2-D :6216-6290) and adds all its couples; every such couple is
// already in the pattern of the space, and for value terms only the couples of nodes lying on the face receive
// something: A(i, j) += measure(face) * c * sum_q w_q phi_i phi_j at PBord(face, q).
// Same ownership as the linear case: one thread per node row walks the row's boundary items in order (sorted by
// boundary element), finds each face node's column by bisection in the sorted node row and adds in place.
// ----------------------------------------------------------------------------------------------------
struct BndBilParams {
    double Mb[4][10][10]; // [face][local node a (row)][local node b (column)]: sum_q w_q phi_a phi_b
    double C[3][3];       // [vcomp][ucomp]
};
template <int DIM>
__global__ void k_bnd_bilinear(int nrows, const int32_t *__restrict__ ptr, const uint32_t *__restrict__ items,
                               const int32_t *__restrict__ belem, const int32_t *__restrict__ bface, const double *__restrict__ meas,
                               const int32_t *__restrict__ e2n, int nloc, int order, int nc, const int32_t *__restrict__ nrowptr,
                               const int32_t *__restrict__ ncol, const __grid_constant__ BndBilParams Bq, double *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int k0 = ptr[i], k1 = ptr[i + 1];
    if (k0 == k1) return;
    const int rb = nrowptr[i], L = nrowptr[i + 1] - rb;
    double *row = vals + (size_t)nc * nc * rb;
    for (int k = k0; k < k1; ++k) {
        const uint32_t it = items[k];
        const int e = it >> 4, a = it & 15;
        const double m = meas[e];
        if (m == 0.0) continue; // label not listed
        const int f = bface[e];
        const int32_t *N = e2n + (size_t)nloc * belem[e];
        int loc[6];
        const int nf = face_nodes<DIM>(order, f, loc);
        for (int x = 0; x < nf; ++x) {
            const int b = loc[x], j = N[b];
            int lo = 0, hi = L - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ncol[rb + mid] < j) lo = mid + 1;
                else hi = mid;
            }
            const double w = m * Bq.Mb[f][a][b];
            for (int cv = 0; cv < nc; ++cv)
                for (int cu = 0; cu < nc; ++cu)
                    if (Bq.C[cv][cu] != 0.0) row[(size_t)cv * nc * L + (size_t)lo * nc + cu] += Bq.C[cv][cu] * w;
        }
    }
}

extern "C" int ffcuda_assemble_bilinear_boundary(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq,
                                                 const double *qpts, const double *qw, int nlab, const int32_t *labels, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(A && s && A->pattern && A->pattern->space == s, "ffcuda_assemble_bilinear_boundary: matrix was not created on this space");
    FF_REQUIRE(nterms >= 0 && (nterms == 0 || terms), "bad term list");
    FF_REQUIRE(nq > 0 && qpts && qw, "face quadrature rule missing");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ffcuda_mesh *m = s->mesh;
    ffcuda_pattern *P = A->pattern;
    const int dim = m->dim, nloc = s->nloc, nc = s->ncomp;
    FF_REQUIRE(m->nbe > 0 && m->belem.p && m->bface.p, "the mesh has no boundary elements");
    if (bnd_has_derivative(nterms, terms)) { // dx(u) v, u dz(v), ...: the general path
        bnd_bilinear_general(A, s, nterms, terms, nq, qpts, qw, nullptr, nlab, labels, accumulate);
        return 0;
    }
    BndBilParams Bq;
    memset(&Bq, 0, sizeof(Bq));
    for (int t = 0; t < nterms; ++t) {
        const ffcuda_bterm &T = terms[t];
        FF_REQUIRE(T.ucomp >= 0 && T.ucomp < nc && T.vcomp >= 0 && T.vcomp < nc, "term component out of range");
        Bq.C[T.vcomp][T.ucomp] += T.coef;
    }
    BndParams Bp;
    memset(&Bp, 0, sizeof(Bp));
    bnd_fill_labels(Bp, nlab, labels);
    for (int f = 0; f <= dim; ++f)
        for (int q = 0; q < nq; ++q) {
            double B[10][4];
            face_ref_basis(dim, s->order, f, qpts + (size_t)q * (dim - 1), B);
            for (int a = 0; a < nloc; ++a)
                for (int b = 0; b < nloc; ++b) Bq.Mb[f][a][b] += qw[q] * B[a][0] * B[b][0];
        }
    cudaStream_t st = ctx->stream;
    if (accumulate) ff_matrix_touch(A);
    else FF_CUDA(cudaMemsetAsync(A->vals.p, 0, A->vals.bytes(), st));
    A->vals_stale = false;
    A->vals_epoch++;
    if (dim == 3) bnd_incidence<3>(ctx, s);
    else bnd_incidence<2>(ctx, s);
    DBuf<double> meas;
    meas.alloc((size_t)m->nbe);
    bnd_measures(ctx, m, Bp, meas.p);
    const int nrows = s->nnodes_owned;
    ff_launch(ctx, "bnd_bilinear", [&] {
        if (dim == 3)
            k_bnd_bilinear<3><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd_ptr.p, s->bnd_items.p, m->belem.p, m->bface.p, meas.p, s->e2n, nloc,
                                                                   s->order, nc, P->nrowptr.p, P->ncol.p, Bq, A->vals.p);
        else
            k_bnd_bilinear<2><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd_ptr.p, s->bnd_items.p, m->belem.p, m->bface.p, meas.p, s->e2n, nloc,
                                                                   s->order, nc, P->nrowptr.p, P->ncol.p, Bq, A->vals.p);
    });
    FF_API_END(s ? s->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// Linear forms with data that depend on the mesh point: int3d(Th)(f(x,y,z) v), int2d(Th)(uold v / dt), ...
// Element_rhs (fflib/problem.cpp:7839-7985) evaluates the coefficient expression at every quadrature node of every
// element; the caller (the plugin, through FreeFEM's own expression evaluator) hands exactly those values over:
// fq[(c * nt + k) * nq + q] = coefficient of v_c at node q of element k (0 for elements outside the integral's region).
// Value terms only:  b[dof(node a of k, c)] += |K| sum_q w_q fq[c][k][q] phi_a(q).
// Two passes, no atomics: per element the nloc x ncomp weighted sums (coalesced over elements), then the row-owner
// gather over the node -> element incidence of the space (fixed order: lists are sorted by element).
// ----------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void k_rhs_qvalues_elem(int nt, const int32_t *__restrict__ conn, const double *__restrict__ xyz, int vstride, int nloc, int nc,
                                   int nq, const double *__restrict__ wphi /* [q][a] = w_q phi_a(q) */, const double *__restrict__ fq,
                                   double *__restrict__ G /* [k][a][c] */)
{
    extern __shared__ double swphi[];
    for (int i = threadIdx.x; i < nq * nloc; i += blockDim.x) swphi[i] = wphi[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt) return;
    double X[DIM + 1][DIM], N[DIM + 1][DIM], det;
    const int32_t *K = conn + (size_t)(DIM + 1) * k;
#pragma unroll
    for (int a = 0; a <= DIM; ++a) {
        const double *P = xyz + (size_t)K[a] * vstride;
#pragma unroll
        for (int d = 0; d < DIM; ++d) X[a][d] = P[d];
    }
    p1_normals<DIM>(X, N, det);
    const double mes = det * (DIM == 3 ? 1.0 / 6.0 : 0.5);
    for (int c = 0; c < nc; ++c) {
        const double *f = fq + ((size_t)c * nt + k) * nq;
        for (int a = 0; a < nloc; ++a) {
            double s = 0.0;
            for (int q = 0; q < nq; ++q) s = fma(swphi[q * nloc + a], f[q], s);
            G[((size_t)k * nloc + a) * nc + c] = mes * s;
        }
    }
}
__global__ void k_rhs_qvalues_gather(int nrows, const IncView V, int nloc, int nc, const double *__restrict__ G, double *__restrict__ b,
                                     int accumulate)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int cnt = V.cnt[row];
    double s[3] = {0.0, 0.0, 0.0};
    for (int e = 0; e < cnt; ++e) {
        const uint32_t rec = V.inc[V.idx(row, e)];
        const double *g = G + ((size_t)(rec >> 4) * nloc + (rec & 15u)) * nc;
        for (int c = 0; c < nc; ++c) s[c] += g[c];
    }
    for (int c = 0; c < nc; ++c) {
        double *dst = b + (size_t)row * nc + c;
        *dst = accumulate ? *dst + s[c] : s[c];
    }
}

extern "C" int ffcuda_assemble_linear_qvalues(ffcuda_vec *b, ffcuda_space *s, int nq, const double *qpts, const double *qw,
                                              const double *fq, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(b && s && fq, "ffcuda_assemble_linear_qvalues: null argument");
    FF_REQUIRE(nq > 0 && nq <= 256 && qpts && qw, "quadrature rule missing (or more than 256 nodes)");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ff_build_incidence(s);
    ffcuda_mesh *m = s->mesh;
    FF_REQUIRE(!m->distributed, "ffcuda_assemble_linear_qvalues: single-GPU meshes only");
    FF_REQUIRE(b->n >= s->nnodes_owned * s->ncomp, "right-hand side vector too short");
    const int dim = m->dim, nloc = s->nloc, nc = s->ncomp, nt = m->nt;
    std::vector<double> wphi((size_t)nq * nloc);
    for (int q = 0; q < nq; ++q) {
        double B[10][4];
        ref_basis(dim, s->order, qpts + (size_t)q * dim, B);
        for (int a = 0; a < nloc; ++a) wphi[(size_t)q * nloc + a] = qw[q] * B[a][0];
    }
    cudaStream_t st = ctx->stream;
    DBuf<double> dW, dF, G;
    dW.alloc(wphi.size());
    G.alloc((size_t)nt * nloc * nc);
    FF_CUDA(cudaMemcpyAsync(dW.p, wphi.data(), dW.bytes(), cudaMemcpyHostToDevice, st));
    const double *pF = table_on_device(fq, (size_t)nc * nt * nq, dF, st);
    const size_t shmem = wphi.size() * sizeof(double);
    ff_launch(ctx, "rhs_qvalues_elem", [&] {
        if (dim == 3)
            k_rhs_qvalues_elem<3><<<ff_blocks(nt, 128), 128, shmem, st>>>(nt, m->conn.p, m->xyz.p, m->vstride, nloc, nc, nq, dW.p, pF, G.p);
        else
            k_rhs_qvalues_elem<2><<<ff_blocks(nt, 128), 128, shmem, st>>>(nt, m->conn.p, m->xyz.p, m->vstride, nloc, nc, nq, dW.p, pF, G.p);
    });
    const int nrows = s->nnodes_owned;
    const IncView V = ff_view(s->incidence);
    ff_launch(ctx, "rhs_qvalues_gather", [&] {
        k_rhs_qvalues_gather<<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, V, nloc, nc, G.p, b->d.p, accumulate);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // fq is the caller's pageable memory
    FF_API_END(s ? s->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// Boundary integrals whose data depend on the mesh point: int2d(Th3, labels)(g(x,y,z) v), int1d(Th, labels)(alpha(x,y) u v).
// Element_rhs / Element_Op evaluate the coefficient at every face quadrature node (fflib/problem.cpp:8551-8570, :6526-6556);
// the caller hands those values over (0 on boundary elements whose label is not listed), the sums over the nodes are
// formed inside the owner gathers of the constant-coefficient kernels above.
// ----------------------------------------------------------------------------------------------------
__global__ void k_bnd_gather_q(int nrows, const int32_t *__restrict__ ptr, const uint32_t *__restrict__ items,
                               const int32_t *__restrict__ bface, const double *__restrict__ meas, int nc, int nq, int nloc, int nbe,
                               const double *__restrict__ wphi /* [f][q][a] = w_q phi_a(PBord(f, q)) */,
                               const double *__restrict__ gq /* [c][e][q] */, double *__restrict__ b, int accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    double s[3] = {0.0, 0.0, 0.0};
    for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
        const uint32_t it = items[k];
        const int e = it >> 4, a = it & 15;
        const double *w = wphi + (size_t)bface[e] * nq * nloc + a;
        for (int c = 0; c < nc; ++c) {
            const double *g = gq + ((size_t)c * nbe + e) * nq;
            double t = 0.0;
            for (int q = 0; q < nq; ++q) t = fma(w[(size_t)q * nloc], g[q], t);
            s[c] += meas[e] * t;
        }
    }
    if (!accumulate || ptr[i + 1] > ptr[i])
        for (int c = 0; c < nc; ++c) {
            double *dst = b + (size_t)i * nc + c;
            *dst = accumulate ? *dst + s[c] : s[c];
        }
}
template <int DIM>
__global__ void k_bnd_bilinear_q(int nrows, const int32_t *__restrict__ ptr, const uint32_t *__restrict__ items,
                                 const int32_t *__restrict__ belem, const int32_t *__restrict__ bface, const double *__restrict__ meas,
                                 const int32_t *__restrict__ e2n, int nloc, int order, int nc, int nq,
                                 const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ ncol,
                                 const double *__restrict__ wpp /* [f][q][a][b] = w_q phi_a phi_b */, const double *__restrict__ cq /* [e][q] */,
                                 const __grid_constant__ BndBilParams Bq, double *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int k0 = ptr[i], k1 = ptr[i + 1];
    if (k0 == k1) return;
    const int rb = nrowptr[i], L = nrowptr[i + 1] - rb;
    double *row = vals + (size_t)nc * nc * rb;
    for (int k = k0; k < k1; ++k) {
        const uint32_t it = items[k];
        const int e = it >> 4, a = it & 15;
        const double m = meas[e];
        if (m == 0.0) continue; // label not listed
        const int f = bface[e];
        const int32_t *N = e2n + (size_t)nloc * belem[e];
        const double *c = cq + (size_t)e * nq;
        int loc[6];
        const int nf = face_nodes<DIM>(order, f, loc);
        for (int x = 0; x < nf; ++x) {
            const int b = loc[x], j = N[b];
            int lo = 0, hi = L - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ncol[rb + mid] < j) lo = mid + 1;
                else hi = mid;
            }
            double t = 0.0;
            for (int q = 0; q < nq; ++q) t = fma(wpp[(((size_t)f * nq + q) * nloc + a) * nloc + b], c[q], t);
            const double w = m * t;
            for (int cv = 0; cv < nc; ++cv)
                for (int cu = 0; cu < nc; ++cu)
                    if (Bq.C[cv][cu] != 0.0) row[(size_t)cv * nc * L + (size_t)lo * nc + cu] += Bq.C[cv][cu] * w;
        }
    }
}

extern "C" int ffcuda_assemble_linear_boundary_qvalues(ffcuda_vec *b, ffcuda_space *s, int nq, const double *qpts, const double *qw,
                                                       const double *gq, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(b && s && gq, "ffcuda_assemble_linear_boundary_qvalues: null argument");
    FF_REQUIRE(nq > 0 && qpts && qw, "face quadrature rule missing");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ffcuda_mesh *m = s->mesh;
    const int dim = m->dim, nloc = s->nloc, nc = s->ncomp, nbe = m->nbe;
    FF_REQUIRE(b->n >= s->nnodes_owned * nc, "right-hand side vector too short");
    FF_REQUIRE(nbe > 0 && m->belem.p && m->bface.p, "the mesh has no boundary elements");
    std::vector<double> wphi((size_t)(dim + 1) * nq * nloc);
    for (int f = 0; f <= dim; ++f)
        for (int q = 0; q < nq; ++q) {
            double B[10][4];
            face_ref_basis(dim, s->order, f, qpts + (size_t)q * (dim - 1), B);
            for (int a = 0; a < nloc; ++a) wphi[((size_t)f * nq + q) * nloc + a] = qw[q] * B[a][0];
        }
    BndParams Bp;
    memset(&Bp, 0, sizeof(Bp));
    Bp.nlab = -1; // every boundary element: the table is 0 where the integral does not go
    cudaStream_t st = ctx->stream;
    if (dim == 3) bnd_incidence<3>(ctx, s);
    else bnd_incidence<2>(ctx, s);
    DBuf<double> meas, dW, dG;
    meas.alloc((size_t)nbe);
    dW.alloc(wphi.size());
    FF_CUDA(cudaMemcpyAsync(dW.p, wphi.data(), dW.bytes(), cudaMemcpyHostToDevice, st));
    const double *pG = table_on_device(gq, (size_t)nc * nbe * nq, dG, st);
    bnd_measures(ctx, m, Bp, meas.p);
    const int nrows = s->nnodes_owned;
    ff_launch(ctx, "bnd_gather_q", [&] {
        k_bnd_gather_q<<<ff_blocks(nrows, 256), 256, 0, st>>>(nrows, s->bnd_ptr.p, s->bnd_items.p, m->bface.p, meas.p, nc, nq, nloc, nbe, dW.p, pG,
                                                            b->d.p, accumulate);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // gq is the caller's pageable memory
    FF_API_END(s ? s->ctx : nullptr)
}

extern "C" int ffcuda_assemble_bilinear_boundary_qcoef(ffcuda_matrix *A, ffcuda_space *s, int nterms, const ffcuda_bterm *terms, int nq,
                                                       const double *qpts, const double *qw, const double *cq, int nlab,
                                                       const int32_t *labels, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(A && s && cq && A->pattern && A->pattern->space == s, "ffcuda_assemble_bilinear_boundary_qcoef: bad arguments");
    FF_REQUIRE(nterms >= 0 && (nterms == 0 || terms), "bad term list");
    FF_REQUIRE(nq > 0 && qpts && qw, "face quadrature rule missing");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ffcuda_mesh *m = s->mesh;
    ffcuda_pattern *P = A->pattern;
    const int dim = m->dim, nloc = s->nloc, nc = s->ncomp, nbe = m->nbe;
    FF_REQUIRE(nbe > 0 && m->belem.p && m->bface.p, "the mesh has no boundary elements");
    if (bnd_has_derivative(nterms, terms)) { // N.x * dx(u) * v, ...: the general path with the coefficient table
        bnd_bilinear_general(A, s, nterms, terms, nq, qpts, qw, cq, nlab, labels, accumulate);
        return 0;
    }
    BndBilParams Bq;
    memset(&Bq, 0, sizeof(Bq));
    for (int t = 0; t < nterms; ++t) {
        const ffcuda_bterm &T = terms[t];
        FF_REQUIRE(T.ucomp >= 0 && T.ucomp < nc && T.vcomp >= 0 && T.vcomp < nc, "term component out of range");
        Bq.C[T.vcomp][T.ucomp] += T.coef;
    }
    BndParams Bp;
    memset(&Bp, 0, sizeof(Bp));
    bnd_fill_labels(Bp, nlab, labels);
    std::vector<double> wpp((size_t)(dim + 1) * nq * nloc * nloc);
    for (int f = 0; f <= dim; ++f)
        for (int q = 0; q < nq; ++q) {
            double B[10][4];
            face_ref_basis(dim, s->order, f, qpts + (size_t)q * (dim - 1), B);
            for (int a = 0; a < nloc; ++a)
                for (int b = 0; b < nloc; ++b) wpp[(((size_t)f * nq + q) * nloc + a) * nloc + b] = qw[q] * B[a][0] * B[b][0];
        }
    cudaStream_t st = ctx->stream;
    if (accumulate) ff_matrix_touch(A);
    else FF_CUDA(cudaMemsetAsync(A->vals.p, 0, A->vals.bytes(), st));
    A->vals_stale = false;
    A->vals_epoch++;
    if (dim == 3) bnd_incidence<3>(ctx, s);
    else bnd_incidence<2>(ctx, s);
    DBuf<double> meas, dW, dC;
    meas.alloc((size_t)nbe);
    dW.alloc(wpp.size());
    FF_CUDA(cudaMemcpyAsync(dW.p, wpp.data(), dW.bytes(), cudaMemcpyHostToDevice, st));
    const double *pC = table_on_device(cq, (size_t)nbe * nq, dC, st);
    bnd_measures(ctx, m, Bp, meas.p);
    const int nrows = s->nnodes_owned;
    ff_launch(ctx, "bnd_bilinear_q", [&] {
        if (dim == 3)
            k_bnd_bilinear_q<3><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd_ptr.p, s->bnd_items.p, m->belem.p, m->bface.p, meas.p, s->e2n,
                                                                     nloc, s->order, nc, nq, P->nrowptr.p, P->ncol.p, dW.p, pC, Bq, A->vals.p);
        else
            k_bnd_bilinear_q<2><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, s->bnd_ptr.p, s->bnd_items.p, m->belem.p, m->bface.p, meas.p, s->e2n,
                                                                     nloc, s->order, nc, nq, P->nrowptr.p, P->ncol.p, dW.p, pC, Bq, A->vals.p);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // cq is the caller's pageable memory
    FF_API_END(s ? s->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// Linear forms with data depending on the mesh point AND derivatives of the test function (the residual of a Newton step,
// int(dx(uk) dx(v) + ...)): fq[((c * (DIM+1) + s) * nt + k) * nq + q] = coefficient of d^s v_c at node q of element k.
// d_x phi_a(q) = sum_r dhat_r phi_a(q) grad(lambda_r)[x], grad(lambda_r) = N_r / det.  Same two passes as the value-only
// entry above: per element the nloc x ncomp weighted sums, then the row-owner gather.
// ----------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void k_rhs_qterms_elem(int nt, const int32_t *__restrict__ conn, const double *__restrict__ xyz, int vstride, int nloc, int nc,
                                  int nq, const double *__restrict__ wB /* [q][a][s] = w_q dhat^s phi_a(q) */,
                                  const double *__restrict__ fq, double *__restrict__ G /* [k][a][c] */)
{
    extern __shared__ double swB[];
    constexpr int NS = DIM + 1;
    for (int i = threadIdx.x; i < nq * nloc * NS; i += blockDim.x) swB[i] = wB[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt) return;
    double X[DIM + 1][DIM], N[DIM + 1][DIM], det;
    const int32_t *K = conn + (size_t)(DIM + 1) * k;
#pragma unroll
    for (int a = 0; a <= DIM; ++a) {
        const double *P = xyz + (size_t)K[a] * vstride;
#pragma unroll
        for (int d = 0; d < DIM; ++d) X[a][d] = P[d];
    }
    p1_normals<DIM>(X, N, det);
    const double mes = det * (DIM == 3 ? 1.0 / 6.0 : 0.5), rdet = 1.0 / det;
    for (int c = 0; c < nc; ++c) {
        double acc[10];
#pragma unroll
        for (int a = 0; a < 10; ++a) acc[a] = 0.0;
        for (int q = 0; q < nq; ++q) {
            double h[NS]; // h[0] = f_0, h[r] = sum_x grad(lambda_r)[x] f_x : coefficients of phi_a and of dhat_r phi_a
            h[0] = fq[(((size_t)c * NS) * nt + k) * nq + q];
            double fx[DIM];
#pragma unroll
            for (int x = 0; x < DIM; ++x) fx[x] = fq[(((size_t)c * NS + 1 + x) * nt + k) * nq + q];
#pragma unroll
            for (int r = 1; r <= DIM; ++r) {
                double t = 0.0;
#pragma unroll
                for (int x = 0; x < DIM; ++x) t = fma(N[r][x], fx[x], t);
                h[r] = t * rdet;
            }
            const double *w = swB + (size_t)q * nloc * NS;
#pragma unroll
            for (int a = 0; a < 10; ++a)
                if (a < nloc) {
#pragma unroll
                    for (int s2 = 0; s2 < NS; ++s2) acc[a] = fma(w[a * NS + s2], h[s2], acc[a]);
                }
        }
#pragma unroll
        for (int a = 0; a < 10; ++a)
            if (a < nloc) G[((size_t)k * nloc + a) * nc + c] = mes * acc[a];
    }
}

extern "C" int ffcuda_assemble_linear_qterms(ffcuda_vec *b, ffcuda_space *s, int nq, const double *qpts, const double *qw,
                                             const double *fq, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(b && s && fq, "ffcuda_assemble_linear_qterms: null argument");
    FF_REQUIRE(nq > 0 && nq <= 256 && qpts && qw, "quadrature rule missing (or more than 256 nodes)");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    ff_build_incidence(s);
    ffcuda_mesh *m = s->mesh;
    FF_REQUIRE(!m->distributed, "ffcuda_assemble_linear_qterms: single-GPU meshes only");
    FF_REQUIRE(b->n >= s->nnodes_owned * s->ncomp, "right-hand side vector too short");
    const int dim = m->dim, nloc = s->nloc, nc = s->ncomp, nt = m->nt, ns = dim + 1;
    std::vector<double> wB((size_t)nq * nloc * ns);
    for (int q = 0; q < nq; ++q) {
        double B[10][4];
        ref_basis(dim, s->order, qpts + (size_t)q * dim, B);
        for (int a = 0; a < nloc; ++a)
            for (int s2 = 0; s2 < ns; ++s2) wB[((size_t)q * nloc + a) * ns + s2] = qw[q] * B[a][s2];
    }
    cudaStream_t st = ctx->stream;
    DBuf<double> dW, dF, G;
    dW.alloc(wB.size());
    G.alloc((size_t)nt * nloc * nc);
    FF_CUDA(cudaMemcpyAsync(dW.p, wB.data(), dW.bytes(), cudaMemcpyHostToDevice, st));
    const double *pF = table_on_device(fq, (size_t)nc * ns * nt * nq, dF, st);
    const size_t shmem = wB.size() * sizeof(double);
    FF_REQUIRE(shmem <= 96 * 1024, "quadrature rule too large for the shared-memory basis table");
    ff_launch(ctx, "rhs_qterms_elem", [&] {
        if (dim == 3) {
            FF_CUDA(cudaFuncSetAttribute(k_rhs_qterms_elem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
            k_rhs_qterms_elem<3><<<ff_blocks(nt, 128), 128, shmem, st>>>(nt, m->conn.p, m->xyz.p, m->vstride, nloc, nc, nq, dW.p, pF, G.p);
        } else {
            FF_CUDA(cudaFuncSetAttribute(k_rhs_qterms_elem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
            k_rhs_qterms_elem<2><<<ff_blocks(nt, 128), 128, shmem, st>>>(nt, m->conn.p, m->xyz.p, m->vstride, nloc, nc, nq, dW.p, pF, G.p);
        }
    });
    const int nrows = s->nnodes_owned;
    const IncView V = ff_view(s->incidence);
    ff_launch(ctx, "rhs_qvalues_gather", [&] {
        k_rhs_qvalues_gather<<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, V, nloc, nc, G.p, b->d.p, accumulate);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // fq is the caller's pageable memory
    FF_API_END(s ? s->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// FE functions as data of a form, handed over as DOF ARRAYS: int3d(Th)(f v), int3d(Th)(kappa grad u . grad v),
// the residual int3d(Th)(dx(uk) dx(v) + ...) of a Newton step with f, kappa, uk P0 / P1 / P2 functions on the mesh of the
// form.  The reference evaluates such a coefficient through the interpreter at every quadrature node of every element
// (Element_Op / Element_rhs, fflib/problem.cpp:6380-6407, :7951-7975, calling pfer2R / pf3r2R -> FElement::operator()(PHat,
// u, comp, op), femlib/FESpace.cpp:1078-1099, :1637-1654, femlib/P012_3d.cpp:98-122): sum_a u[K(a)] d^op phi_a(PHat).  Here
// the same sum is formed on the device from the dof vector, one thread per (unit, node), and lands in the table layout the
// q-table entries above consume (they take it where it lies: table_on_device).
// ----------------------------------------------------------------------------------------------------
struct FeTabLabels {
    int nlab; // < 0: every unit
    int labels[MAXBL];
};
template <int DIM, bool BND>
__global__ void __launch_bounds__(256)
k_fe_table(size_t nitems, int nq, int nloc, int slot /* 0: value, 1..DIM: d/dx_slot */, const int32_t *__restrict__ conn,
           const double *__restrict__ xyz, int vstride, const int32_t *__restrict__ ulab, const int32_t *__restrict__ belem,
           const int32_t *__restrict__ bface, const int32_t *__restrict__ e2n /* nt x nloc or null */, int dstride, int doff,
           const double *__restrict__ dofs, const double *__restrict__ Bq /* [face][q][a][4]: value, dhat_1..3 */, double scale,
           const FeTabLabels L, double *__restrict__ out, int accumulate)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    const int u = (int)(i / (size_t)nq), q = (int)(i - (size_t)u * nq);
    bool in = L.nlab < 0;
    if (!in) {
        const int l = ulab[u];
        for (int j = 0; j < L.nlab; ++j) in |= (L.labels[j] == l);
    }
    double val = 0.0;
    if (in) {
        const int k = BND ? belem[u] : u, f = BND ? bface[u] : 0;
        const int32_t *K = conn + (size_t)(DIM + 1) * k;
        const double *B = Bq + ((size_t)f * nq + q) * nloc * 4;
        double g[DIM]; // g[r-1] = d lambda_r / d x_slot
        if (slot > 0) {
            double X[DIM + 1][DIM], N[DIM + 1][DIM], det;
#pragma unroll
            for (int a = 0; a <= DIM; ++a)
#pragma unroll
                for (int d = 0; d < DIM; ++d) X[a][d] = xyz[(size_t)K[a] * vstride + d];
            p1_normals<DIM>(X, N, det);
            const double rdet = 1.0 / det;
#pragma unroll
            for (int r = 1; r <= DIM; ++r) {
                double nx = N[r][0];
#pragma unroll
                for (int d = 1; d < DIM; ++d)
                    if (slot - 1 == d) nx = N[r][d];
                g[r - 1] = nx * rdet;
            }
        }
        for (int a = 0; a < nloc; ++a) {
            const int node = e2n ? e2n[(size_t)k * nloc + a] : (nloc == 1 ? k : K[a]);
            const double ua = dofs[(size_t)node * dstride + doff];
            double w;
            if (slot == 0) w = B[a * 4];
            else {
                w = 0.0;
#pragma unroll
                for (int r = 1; r <= DIM; ++r) w = fma(B[a * 4 + r], g[r - 1], w);
            }
            val = fma(ua, w, val);
        }
        val *= scale;
    }
    out[i] = accumulate ? out[i] + val : val;
}

extern "C" int ffcuda_fe_table(ffcuda_mesh *m, int order, const int32_t *e2n, int dstride, int doff, ffcuda_vec *dofs, int op,
                               int border, int nq, const double *qpts, double scale, int nlab, const int32_t *labels,
                               ffcuda_vec *table, int64_t offset, int accumulate)
{
    FF_API_BEGIN
    FF_REQUIRE(m && dofs && table && qpts, "ffcuda_fe_table: null argument");
    FF_REQUIRE(!m->distributed, "ffcuda_fe_table: single-GPU meshes only");
    FF_REQUIRE(order >= 0 && order <= 2, "ffcuda_fe_table: the function must be P0, P1 or P2 Lagrange");
    FF_REQUIRE(nq > 0 && nq <= 256, "quadrature rule missing (or more than 256 nodes)");
    FF_REQUIRE(dstride >= 1 && doff >= 0 && doff < dstride, "ffcuda_fe_table: bad dof stride / offset");
    ffcuda_ctx *ctx = m->ctx;
    ff_enter(ctx);
    const int dim = m->dim, nt = m->nt;
    const int nloc = order == 0 ? 1 : order == 1 ? dim + 1 : (dim == 3 ? 10 : 6);
    const int slot = op_slot(dim, op);
    FF_REQUIRE(order != 2 || e2n, "ffcuda_fe_table: a P2 function needs its element -> node table");
    const int nunits = border ? m->nbe : nt;
    FF_REQUIRE(!border || (m->nbe > 0 && m->belem.p && m->bface.p && m->blab.p), "the mesh has no boundary elements");
    FF_REQUIRE(offset >= 0 && (size_t)offset + (size_t)nunits * nq <= (size_t)table->n, "ffcuda_fe_table: table too short");
    // the largest node the table refers to must lie inside the dof vector
    int64_t maxnode = order == 0 ? nt - 1 : m->nv - 1;
    if (e2n) {
        maxnode = -1;
        for (size_t i = 0; i < (size_t)nt * nloc; ++i) {
            FF_REQUIRE(e2n[i] >= 0, "ffcuda_fe_table: negative node in the element -> node table");
            maxnode = std::max<int64_t>(maxnode, e2n[i]);
        }
    }
    FF_REQUIRE(maxnode * dstride + doff < (int64_t)dofs->n, "ffcuda_fe_table: dof vector too short for this space");
    FeTabLabels L;
    memset(&L, 0, sizeof(L));
    if (!labels) L.nlab = -1;
    else {
        FF_REQUIRE(nlab >= 0 && nlab <= MAXBL, "at most 32 labels per integral");
        L.nlab = nlab;
        for (int i = 0; i < nlab; ++i) L.labels[i] = labels[i];
    }
    // reference values and derivatives at the nodes of the rule (volume: one "face"; border: PBord of each face)
    const int nf = border ? dim + 1 : 1;
    std::vector<double> Bq((size_t)nf * nq * nloc * 4, 0.0);
    for (int f = 0; f < nf; ++f)
        for (int q = 0; q < nq; ++q) {
            double B[10][4];
            for (int a = 0; a < 10; ++a)
                for (int s2 = 0; s2 < 4; ++s2) B[a][s2] = 0.0;
            if (order == 0) B[0][0] = 1.0; // constant on the element: derivatives vanish
            else if (border) face_ref_basis(dim, order, f, qpts + (size_t)q * (dim - 1), B);
            else ref_basis(dim, order, qpts + (size_t)q * dim, B);
            for (int a = 0; a < nloc; ++a)
                for (int s2 = 0; s2 <= dim; ++s2) Bq[(((size_t)f * nq + q) * nloc + a) * 4 + s2] = B[a][s2];
        }
    cudaStream_t st = ctx->stream;
    DBuf<double> dB;
    DBuf<int32_t> dE;
    dB.alloc(Bq.size());
    FF_CUDA(cudaMemcpyAsync(dB.p, Bq.data(), dB.bytes(), cudaMemcpyHostToDevice, st));
    if (e2n) {
        dE.alloc((size_t)nt * nloc);
        FF_CUDA(cudaMemcpyAsync(dE.p, e2n, dE.bytes(), cudaMemcpyHostToDevice, st));
    }
    const size_t nitems = (size_t)nunits * nq;
    double *out = table->d.p + offset;
    const int32_t *ulab = border ? m->blab.p : m->elab.p;
    FF_REQUIRE(L.nlab < 0 || ulab, "ffcuda_fe_table: the mesh carries no labels");
    if (nitems)
        ff_launch(ctx, "fe_table", [&] {
            const int nb = ff_blocks(nitems, 256);
            if (dim == 3 && border)
                k_fe_table<3, true><<<nb, 256, 0, st>>>(nitems, nq, nloc, slot, m->conn.p, m->xyz.p, m->vstride, ulab, m->belem.p, m->bface.p, dE.p,
                                                        dstride, doff, dofs->d.p, dB.p, scale, L, out, accumulate);
            else if (dim == 3)
                k_fe_table<3, false><<<nb, 256, 0, st>>>(nitems, nq, nloc, slot, m->conn.p, m->xyz.p, m->vstride, ulab, nullptr, nullptr, dE.p,
                                                         dstride, doff, dofs->d.p, dB.p, scale, L, out, accumulate);
            else if (border)
                k_fe_table<2, true><<<nb, 256, 0, st>>>(nitems, nq, nloc, slot, m->conn.p, m->xyz.p, m->vstride, ulab, m->belem.p, m->bface.p, dE.p,
                                                        dstride, doff, dofs->d.p, dB.p, scale, L, out, accumulate);
            else
                k_fe_table<2, false><<<nb, 256, 0, st>>>(nitems, nq, nloc, slot, m->conn.p, m->xyz.p, m->vstride, ulab, nullptr, nullptr, dE.p,
                                                         dstride, doff, dofs->d.p, dB.p, scale, L, out, accumulate);
            FF_CUDA(cudaGetLastError());
        });
    FF_CUDA(cudaStreamSynchronize(st)); // Bq and e2n are host memory of this call / of the caller
    FF_API_END(m ? m->ctx : nullptr)
}

// ----------------------------------------------------------------------------------------------------
// RECTANGULAR matrices: `matrix B = vb(Uh,Vh)` with two different spaces on one mesh (the blocks of a Stokes / mixed
// problem assembled one by one, interpolation and projection matrices between P1 and P2).  Rows = dofs of the test space,
// columns = dofs of the space of the unknown (Element_Op with Ku != Kv, fflib/problem.cpp:6337-6437, 2-D :6063-6160; the
// operator takes the two spaces, fflib/problem.hpp:1628-1631).  Symbolic phase: ff_rect_node_pattern (symbolic.cu);
// numeric phase: one thread per row node runs rect_row (rect_row.cuh) over the incidence lists of the test space.  The
// result is a matrix without a pattern object (like ffcuda_matrix_from_csr): products and hand-off formats apply, the
// solvers and Dirichlet entries do not.  O(nt nq nloc_v nloc_u) on a general kernel: clarity over speed, this is not the
// square hot path.
// ----------------------------------------------------------------------------------------------------
#include "rect_row.cuh"

struct RecView { // e-th record of row `row` in either incidence layout
    IncView V;
    int row;
    __device__ __forceinline__ uint32_t operator()(int e) const { return V.inc[V.idx(row, e)]; }
};

template <int DIM>
__global__ void __launch_bounds__(128) k_asm_rect(int nrows, const IncView V, const int32_t *__restrict__ conn, const int32_t *__restrict__ elab,
                                                  const double *__restrict__ xyz, int vstride, const int32_t *__restrict__ e2n_u,
                                                  const RectParams *__restrict__ Pp, const int32_t *__restrict__ nrowptr,
                                                  const int32_t *__restrict__ ncol, double *__restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int rb = nrowptr[i], L = nrowptr[i + 1] - rb;
    if (L == 0) return;
    rect_row<DIM>(V.cnt[i], RecView{V, i}, conn, elab, xyz, vstride, e2n_u, *Pp, ncol, rb, L,
                  vals + (size_t)Pp->ncv * Pp->ncu * rb);
}

// node-level CSR -> dof-level CSR of the component-block layout: row (i, cv) holds, for every column node p of row node i
// in order, the ncu components: rowptr[i*ncv + cv] = ncu * (ncv * nrowptr[i] + cv * L), colind = ncol[p] * ncu + cu
__global__ void k_rect_expand(int nrows, int ncv, int ncu, const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ ncol,
                              int32_t *__restrict__ rowptr, int32_t *__restrict__ colind)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nrows) return;
    if (i == nrows) {
        rowptr[(size_t)nrows * ncv] = ncu * ncv * nrowptr[nrows];
        return;
    }
    const int b = nrowptr[i], L = nrowptr[i + 1] - b;
    for (int cv = 0; cv < ncv; ++cv) {
        const int r = ncu * (ncv * b + cv * L);
        rowptr[(size_t)i * ncv + cv] = r;
        for (int p = 0; p < L; ++p)
            for (int cu = 0; cu < ncu; ++cu) colind[(size_t)r + (size_t)p * ncu + cu] = ncol[b + p] * ncu + cu;
    }
}

extern "C" int ffcuda_assemble_bilinear_rect(ffcuda_space *sv, ffcuda_space *su, int nterms, const ffcuda_bterm *terms, int nq,
                                             const double *qpts, const double *qw, int nlab, const int32_t *labels, ffcuda_matrix **out)
{
    ffcuda_matrix *A = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(sv && su && out && terms && qpts && qw, "ffcuda_assemble_bilinear_rect: null argument");
    FF_REQUIRE(sv->mesh == su->mesh, "ffcuda_assemble_bilinear_rect: the two spaces must live on the same device mesh");
    ffcuda_ctx *ctx = sv->ctx;
    ffcuda_mesh *m = sv->mesh;
    FF_REQUIRE(!m->distributed, "rectangular matrices on distributed meshes are not on the ffcuda path");
    FF_REQUIRE(nterms >= 1 && nterms <= RECT_MAXT, "rectangular forms: 1 to 64 terms");
    FF_REQUIRE(nq >= 1 && nq <= RECT_MAXQ, "rectangular forms: at most 32 quadrature points");
    FF_REQUIRE(!labels || (nlab >= 0 && nlab <= RECT_MAXLAB), "at most 16 region labels per integral");
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    const int dim = m->dim, ncv = sv->ncomp, ncu = su->ncomp;
    FF_REQUIRE(dim == 2 || dim == 3, "bad dimension");
    FF_REQUIRE(ncv >= 1 && ncv <= 3 && ncu >= 1 && ncu <= 3, "1 to 3 components per space");
    std::unique_ptr<RectParams> P(new RectParams());
    memset(P.get(), 0, sizeof(RectParams));
    P->nq = nq;
    P->nterms = nterms;
    P->nlab = labels ? nlab : -1;
    for (int i = 0; labels && i < nlab; ++i) P->labels[i] = labels[i];
    P->order_v = sv->order;
    P->order_u = su->order;
    P->ncv = ncv;
    P->ncu = ncu;
    P->nloc_u = su->nloc;
    for (int q = 0; q < nq; ++q) {
        P->w[q] = qw[q];
        double l0 = 1.0;
        for (int d = 0; d < dim; ++d) {
            P->lam[q][d + 1] = qpts[(size_t)q * dim + d];
            l0 -= qpts[(size_t)q * dim + d];
        }
        P->lam[q][0] = l0;
    }
    for (int t = 0; t < nterms; ++t) {
        const ffcuda_bterm &T = terms[t];
        FF_REQUIRE(T.ucomp >= 0 && T.ucomp < ncu && T.vcomp >= 0 && T.vcomp < ncv, "term component out of range");
        P->t[t] = RectTerm{T.coef, T.vcomp, T.ucomp, op_slot(T.vop), op_slot(T.uop)};
        FF_REQUIRE(dim == 3 || (P->t[t].uslot < 3 && P->t[t].vslot < 3), "dz on a 2-D mesh");
    }
    // --- symbolic phase
    DBuf<int32_t> nrowptr, ncol;
    int64_t nnzn = 0;
    int maxrow_node = 0;
    ff_rect_node_pattern(sv, su, nrowptr, ncol, &nnzn, &maxrow_node);
    const int nrows = sv->nnodes_owned;
    const int64_t nnz = nnzn * ncv * ncu;
    FF_REQUIRE(nnz > 0, "rectangular form on a mesh without elements");
    FF_REQUIRE(nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros (int32 CSR, like MatriceMorse)");
    A = new ffcuda_matrix();
    A->ctx = ctx;
    A->ref.set(ctx);
    A->rect = true;
    A->n = nrows * ncv;
    A->ncols = su->nnodes * ncu;
    A->nnz = nnz;
    A->maxrow = maxrow_node * ncu;
    A->rowptr_own.alloc((size_t)A->n + 1);
    A->colind_own.alloc((size_t)std::max<int64_t>(nnz, 1));
    A->vals.alloc((size_t)std::max<int64_t>(nnz, 1));
    A->rowptr = A->rowptr_own.p;
    A->colind = A->colind_own.p;
    A->diagpos = nullptr;
    FF_CUDA(cudaMemsetAsync(A->vals.p, 0, A->vals.bytes(), st));
    ff_launch(ctx, "rect_expand", [&] {
        k_rect_expand<<<ff_blocks((size_t)nrows + 1, 128), 128, 0, st>>>(nrows, ncv, ncu, nrowptr.p, ncol.p, A->rowptr_own.p, A->colind_own.p);
    });
    // --- numeric phase
    DBuf<RectParams> dP;
    dP.alloc(1);
    FF_CUDA(cudaMemcpyAsync(dP.p, P.get(), sizeof(RectParams), cudaMemcpyHostToDevice, st));
    const IncView V = ff_view(sv->incidence);
    ff_launch(ctx, "asm_rect", [&] {
        if (dim == 3)
            k_asm_rect<3><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, V, m->conn.p, m->elab.p, m->xyz.p, m->vstride, su->e2n, dP.p, nrowptr.p,
                                                                 ncol.p, A->vals.p);
        else
            k_asm_rect<2><<<ff_blocks(nrows, 128), 128, 0, st>>>(nrows, V, m->conn.p, m->elab.p, m->xyz.p, m->vstride, su->e2n, dP.p, nrowptr.p,
                                                                 ncol.p, A->vals.p);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // P (host) and the node-level pattern go out of scope
    *out = A;
    A = nullptr;
    FF_API_END((delete A, sv ? sv->ctx : nullptr))
}
