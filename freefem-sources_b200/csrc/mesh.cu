// mesh.cu — mesh handles: upload of a host (FreeFEM) mesh, and device-side generation of the structured
// cube / square meshes with FreeFEM's exact vertex, element and boundary-element ordering
// (BuildCube fflib/msh3.cpp:7879-8132 with kind=6; Carre_ fflib/lgmesh.cpp:1229-1384 with flags=0).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <array>
#include <memory>
#include <unordered_map>

// local vertex tables of the reference simplices (femlib/Mesh3dn.cpp:71-72)
__constant__ int c_nvfaceTet[4][3] = {{3, 2, 1}, {0, 2, 3}, {3, 1, 0}, {0, 1, 2}};
// kind=6 split of a cell into 6 tets around the diagonal 0-7 (corner id = a + 2b + 4c)
__constant__ int c_cubeTets[6][4] = {{4, 0, 6, 7}, {0, 4, 5, 7}, {1, 0, 5, 7}, {0, 1, 3, 7}, {2, 0, 3, 7}, {0, 2, 6, 7}};

__global__ void k_pad_xyz3(const double *__restrict__ in, double *__restrict__ out, int nv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    double4 v = make_double4(in[3 * (size_t)i], in[3 * (size_t)i + 1], in[3 * (size_t)i + 2], 0.0);
    reinterpret_cast<double4 *>(out)[i] = v;
}

__global__ void k_unpad_xyz3(const double *__restrict__ in, double *__restrict__ out, int nv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    double4 v = reinterpret_cast<const double4 *>(in)[i];
    out[3 * (size_t)i] = v.x;
    out[3 * (size_t)i + 1] = v.y;
    out[3 * (size_t)i + 2] = v.z;
}

static void recover_boundary_elements(int dim, int nt, const int32_t *conn, int nbe, const int32_t *bconn,
                                      std::vector<int32_t> &belem, std::vector<int32_t> &bface)
{
    // host-side input preparation: match every boundary element with the (element, local face) that has the
    // same vertex set (what Mesh3::BoundaryElement returns).  First element in element order wins.
    struct Key {
        std::array<int32_t, 3> v;
        bool operator==(const Key &o) const { return v == o.v; }
    };
    struct KeyHash {
        size_t operator()(const Key &k) const
        {
            uint64_t h = 1469598103934665603ull;
            for (int i = 0; i < 3; ++i) h = (h ^ (uint32_t)k.v[i]) * 1099511628211ull;
            return (size_t)h;
        }
    };
    std::unordered_map<Key, int32_t, KeyHash> want;
    want.reserve((size_t)nbe * 2);
    for (int ib = 0; ib < nbe; ++ib) {
        Key k;
        k.v = {-1, -1, -1};
        for (int i = 0; i < dim; ++i) k.v[i] = bconn[(size_t)ib * dim + i];
        std::sort(k.v.begin(), k.v.begin() + dim);
        want.emplace(k, ib);
    }
    belem.assign(nbe, -1);
    bface.assign(nbe, -1);
    const int nvk = dim + 1;
    for (int t = 0; t < nt; ++t)
        for (int f = 0; f < nvk; ++f) {
            Key k;
            k.v = {-1, -1, -1};
            int m = 0;
            for (int a = 0; a < nvk; ++a)
                if (a != f) k.v[m++] = conn[(size_t)t * nvk + a];
            std::sort(k.v.begin(), k.v.begin() + dim);
            auto it = want.find(k);
            if (it != want.end() && belem[it->second] < 0) {
                belem[it->second] = t;
                bface[it->second] = f;
            }
        }
    for (int ib = 0; ib < nbe; ++ib)
        FF_REQUIRE(belem[ib] >= 0, "boundary element " + std::to_string(ib) + " is not a face of any element");
}

extern "C" int ffcuda_mesh_upload(ffcuda_ctx *ctx, int dim, int nv, const double *xyz, int nt, const int32_t *conn,
                                  const int32_t *elab, int nbe, const int32_t *bconn, const int32_t *blab,
                                  const int32_t *belem, const int32_t *bface, ffcuda_mesh **out)
{
    ffcuda_mesh *m = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(ctx && out, "ffcuda_mesh_upload: null context/output");
    FF_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    FF_REQUIRE(nv > 0 && nt > 0 && xyz && conn, "empty mesh");
    FF_REQUIRE(nbe == 0 || (bconn && blab), "boundary arrays missing");
    FF_REQUIRE((int64_t)nt < (int64_t)1 << 27, "too many elements for one device (limit 2^27)");
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    m = new ffcuda_mesh();
    m->ctx = ctx;
    m->ref.set(ctx);
    m->dim = dim; m->nv = nv; m->nt = nt; m->nbe = nbe;
    m->nv_owned = nv;
    const int nvk = dim + 1;
    m->vstride = dim == 3 ? 4 : 2;
    m->xyz.alloc((size_t)nv * m->vstride);
    if (dim == 3) {
        DBuf<double> tmp;
        tmp.alloc((size_t)nv * 3);
        FF_CUDA(cudaMemcpyAsync(tmp.p, xyz, tmp.bytes(), cudaMemcpyHostToDevice, st));
        ff_launch(ctx, "mesh_pad_xyz", [&] { k_pad_xyz3<<<ff_blocks(nv, 256), 256, 0, st>>>(tmp.p, m->xyz.p, nv); });
        FF_CUDA(cudaStreamSynchronize(st));
    } else {
        FF_CUDA(cudaMemcpyAsync(m->xyz.p, xyz, m->xyz.bytes(), cudaMemcpyHostToDevice, st));
    }
    m->conn.alloc((size_t)nt * nvk);
    FF_CUDA(cudaMemcpyAsync(m->conn.p, conn, m->conn.bytes(), cudaMemcpyHostToDevice, st));
    m->elab.alloc((size_t)nt);
    if (elab) FF_CUDA(cudaMemcpyAsync(m->elab.p, elab, m->elab.bytes(), cudaMemcpyHostToDevice, st));
    else FF_CUDA(cudaMemsetAsync(m->elab.p, 0, m->elab.bytes(), st));
    if (nbe) {
        std::vector<int32_t> be, bf;
        if (!belem || !bface) {
            recover_boundary_elements(dim, nt, conn, nbe, bconn, be, bf);
            belem = be.data();
            bface = bf.data();
        }
        m->bconn.alloc((size_t)nbe * dim);
        m->blab.alloc(nbe); m->belem.alloc(nbe); m->bface.alloc(nbe);
        FF_CUDA(cudaMemcpyAsync(m->bconn.p, bconn, m->bconn.bytes(), cudaMemcpyHostToDevice, st));
        FF_CUDA(cudaMemcpyAsync(m->blab.p, blab, m->blab.bytes(), cudaMemcpyHostToDevice, st));
        FF_CUDA(cudaMemcpyAsync(m->belem.p, belem, m->belem.bytes(), cudaMemcpyHostToDevice, st));
        FF_CUDA(cudaMemcpyAsync(m->bface.p, bface, m->bface.bytes(), cudaMemcpyHostToDevice, st));
        FF_CUDA(cudaStreamSynchronize(st));
    }
    FF_CUDA(cudaStreamSynchronize(st));
    *out = m;
    m = nullptr;
    FF_API_END((delete m, ctx))
}

// ---------------------------------------------------------------------------------------------------
// cube(nx,ny,nz) — whole, or the z-slab of one rank (owned vertex layers + one ghost layer each side)
// ---------------------------------------------------------------------------------------------------
struct CubeLayout {
    int nx, ny, nz;
    int c_lo, ncl;          // first local cell layer, number of local cell layers
    int L0, nown;           // first owned vertex layer, number of owned vertex layers
    int has_lower, has_upper;
};

// local id of the vertex (i,j,k): owned layers first, then the lower ghost layer, then the upper one
__device__ __forceinline__ int cube_lv(const CubeLayout &C, int i, int j, int k)
{
    const int nj = C.nx + 1, nk = nj * (C.ny + 1), r = j * nj + i;
    if (k >= C.L0 && k < C.L0 + C.nown) return (k - C.L0) * nk + r;
    if (k < C.L0) return C.nown * nk + r;
    return C.nown * nk + (C.has_lower ? nk : 0) + r;
}

__global__ void k_cube_vertices(double *__restrict__ xyz4, int64_t *__restrict__ gid, const CubeLayout C, int nvloc)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nvloc) return;
    const int nj = C.nx + 1, nk = nj * (C.ny + 1);
    int kl = p / nk, r = p - kl * nk;
    int k;
    if (kl < C.nown) k = C.L0 + kl;
    else if (C.has_lower && kl == C.nown) k = C.L0 - 1;
    else k = C.L0 + C.nown;
    int j = r / nj, i = r - j * nj;
    double xd = 1. / C.nx, yd = 1. / C.ny, zd = 1. / C.nz;
    reinterpret_cast<double4 *>(xyz4)[p] = make_double4(0 + xd * i, 0 + yd * j, 0 + zd * k, 0.0);
    if (gid) gid[p] = (int64_t)k * nk + r;
}

__device__ __forceinline__ int cube_vlab(int i, int j, int k, int nx, int ny, int nz)
{
    return 1 * (i == 0) + 2 * (i == nx) + 4 * (j == 0) + 8 * (j == ny) + 16 * (k == 0) + 32 * (k == nz);
}

// one thread per cell: writes its 6 tets; counts (pass 0) or writes (pass 1) its boundary triangles in the
// reference order (tet d, face f, plane kk).
template <int PASS>
__global__ void k_cube_cells(const CubeLayout C, int32_t *__restrict__ conn, int32_t *__restrict__ bcount,
                             const int32_t *__restrict__ boff, int32_t *__restrict__ bconn, int32_t *__restrict__ blab,
                             int32_t *__restrict__ belem, int32_t *__restrict__ bface)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = C.nx, ny = C.ny, nz = C.nz;
    int nc = nx * ny * C.ncl;
    if (c >= nc) return;
    int kc = c / (nx * ny), r = c - kc * nx * ny;
    int j = r / nx, i = r - j * nx;
    int k = C.c_lo + kc;
    int n[8], lab[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int a = q & 1, b = (q >> 1) & 1, cc = q >> 2;
        n[q] = cube_lv(C, i + a, j + b, k + cc);
        lab[q] = cube_vlab(i + a, j + b, k + cc, nx, ny, nz);
    }
    const int nff[6] = {3, 1, 0, 2, 4, 5};
    int kf = PASS ? boff[c] : 0;
    for (int d = 0; d < 6; ++d) {
        int t = 6 * c + d;
        int lc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) lc[q] = c_cubeTets[d][q];
        if (PASS == 0) reinterpret_cast<int4 *>(conn)[t] = make_int4(n[lc[0]], n[lc[1]], n[lc[2]], n[lc[3]]);
        for (int f = 0; f < 4; ++f) {
            int f0 = lc[c_nvfaceTet[f][0]], f1 = lc[c_nvfaceTet[f][1]], f2 = lc[c_nvfaceTet[f][2]];
            int l = lab[f0] & lab[f1] & lab[f2];
            if (l && (l & (l - 1)) == 0) { // exactly one boundary plane
                if (PASS) {
                    int kk = __ffs(l) - 1;
                    bconn[3 * (size_t)kf] = n[f0];
                    bconn[3 * (size_t)kf + 1] = n[f1];
                    bconn[3 * (size_t)kf + 2] = n[f2];
                    blab[kf] = nff[kk] + 1;
                    belem[kf] = t;
                    bface[kf] = f;
                }
                kf++;
            }
        }
    }
    if (PASS == 0) bcount[c] = kf;
}

// The slab partition of cube(nx,ny,nz) along z (host arithmetic only; exported so that it can be checked without a GPU).
// Vertex layers [L0, L0+nown) are owned by `rank`; it holds the cell layers [c_lo, c_lo+ncl) that touch them (one
// layer of halo cells on each inner side) and, after its owned vertices, one ghost vertex layer per neighbour.
struct CubePart {
    int L0, nown, has_lower, has_upper, c_lo, ncl;
    int64_t nk, nv_owned, nv_local, nt_local;
    int nbr[2], send_off[2], recv_off[2], cnt;
};
static CubePart cube_partition(int nx, int ny, int nz, int rank, int nranks)
{
    FF_REQUIRE(nx > 0 && ny > 0 && nz > 0, "cube sizes must be positive");
    FF_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / number of ranks");
    FF_REQUIRE(nz + 1 >= nranks, "cube has fewer vertex layers than ranks");
    CubePart P;
    const int64_t L0 = (int64_t)rank * (nz + 1) / nranks, L1 = (int64_t)(rank + 1) * (nz + 1) / nranks;
    P.L0 = (int)L0;
    P.nown = (int)(L1 - L0);
    P.has_lower = rank > 0;
    P.has_upper = rank < nranks - 1;
    P.c_lo = std::max(P.L0 - 1, 0);
    const int c_hi = std::min(P.L0 + P.nown - 1, nz - 1); // inclusive
    P.ncl = c_hi - P.c_lo + 1;
    P.nk = (int64_t)(nx + 1) * (ny + 1);
    P.nv_owned = P.nk * P.nown;
    P.nv_local = P.nk * (P.nown + P.has_lower + P.has_upper);
    P.nt_local = (int64_t)6 * nx * ny * P.ncl;
    P.nbr[0] = P.has_lower ? rank - 1 : -1;
    P.nbr[1] = P.has_upper ? rank + 1 : -1;
    P.cnt = (int)P.nk;
    P.send_off[0] = 0;                                   // first owned layer goes down
    P.send_off[1] = (int)(P.nk * (P.nown - 1));          // last owned layer goes up
    P.recv_off[0] = (int)P.nv_owned;                     // ghost layer from below
    P.recv_off[1] = (int)(P.nv_owned + (P.has_lower ? P.nk : 0)); // ghost layer from above
    return P;
}

extern "C" int ffcuda_partition_cube(int nx, int ny, int nz, int rank, int nranks, int64_t *out16)
{
    FF_API_BEGIN
    FF_REQUIRE(out16, "null output");
    const CubePart P = cube_partition(nx, ny, nz, rank, nranks);
    const int64_t v[16] = {P.L0, P.nown, P.c_lo, P.ncl, P.nv_owned, P.nv_local, P.nt_local, P.nbr[0], P.nbr[1],
                           P.send_off[0], P.send_off[1], P.recv_off[0], P.recv_off[1], P.cnt, P.has_lower, P.has_upper};
    for (int i = 0; i < 16; ++i) out16[i] = v[i];
    FF_API_END(nullptr)
}

static void build_cube(ffcuda_ctx *ctx, int nx, int ny, int nz, int rank, int nranks, ffcuda_mesh **out)
{
    const CubePart Pt = cube_partition(nx, ny, nz, rank, nranks);
    CubeLayout C;
    C.nx = nx; C.ny = ny; C.nz = nz;
    C.L0 = Pt.L0; C.nown = Pt.nown; C.has_lower = Pt.has_lower; C.has_upper = Pt.has_upper; C.c_lo = Pt.c_lo; C.ncl = Pt.ncl;
    int64_t nc64 = (int64_t)nx * ny * C.ncl;
    FF_REQUIRE(nc64 * 6 < ((int64_t)1 << 27), "cube (slab) too large for one device (limit 2^27 tets); use more ranks");
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    std::unique_ptr<ffcuda_mesh> m(new ffcuda_mesh());
    m->ctx = ctx;
    m->ref.set(ctx);
    m->dim = 3; m->vstride = 4;
    int nc = (int)nc64;
    m->nv_owned = (int)Pt.nv_owned;
    m->nv = (int)Pt.nv_local;
    m->nt = 6 * nc;
    m->xyz.alloc((size_t)m->nv * 4);
    m->conn.alloc((size_t)m->nt * 4);
    m->elab.alloc(m->nt);
    FF_CUDA(cudaMemsetAsync(m->elab.p, 0, m->elab.bytes(), st));
    if (nranks > 1) {
        m->distributed = true;
        m->gid.alloc(m->nv);
        m->nnbr = 2;
        for (int s = 0; s < 2; ++s) {
            m->nbr[s] = Pt.nbr[s];
            m->send_off[s] = Pt.send_off[s];
            m->send_cnt[s] = Pt.cnt;
            m->recv_off[s] = Pt.recv_off[s];
            m->recv_cnt[s] = Pt.cnt;
        }
    }
    ff_launch(ctx, "mesh_cube_vertices", [&] { k_cube_vertices<<<ff_blocks(m->nv, 256), 256, 0, st>>>(m->xyz.p, m->gid.p, C, m->nv); });
    DBuf<int32_t> bcount, boff;
    bcount.alloc(nc); boff.alloc(nc);
    ff_launch(ctx, "mesh_cube_cells", [&] {
        k_cube_cells<0><<<ff_blocks(nc, 128), 128, 0, st>>>(C, m->conn.p, bcount.p, nullptr, nullptr, nullptr, nullptr, nullptr);
    });
    int64_t tot = 0;
    ff_exclusive_scan_i32(ctx, bcount.p, boff.p, nc, &tot);
    if (nranks == 1) FF_REQUIRE(tot == 4 * ((int64_t)nx * ny + (int64_t)nx * nz + (int64_t)ny * nz), "internal: boundary face count mismatch");
    m->nbe = (int)tot;
    m->bconn.alloc((size_t)m->nbe * 3);
    m->blab.alloc(m->nbe); m->belem.alloc(m->nbe); m->bface.alloc(m->nbe);
    ff_launch(ctx, "mesh_cube_bfaces", [&] {
        k_cube_cells<1><<<ff_blocks(nc, 128), 128, 0, st>>>(C, nullptr, nullptr, boff.p, m->bconn.p, m->blab.p, m->belem.p, m->bface.p);
    });
    FF_CUDA(cudaStreamSynchronize(st));
    *out = m.release();
}

// The local problem of one rank for ANY vertex partition (the output of ffcuda_partition_local, or of METIS / the user's
// own partitioner): owned vertices first, ghosts grouped by owner; local elements = every element touching an owned
// vertex, so the owned rows assemble without communication.  Reference counterpart: the element-range split of
// fflib/problem.cpp:1133-1138 (+ an all-reduce of the whole matrix), plugin/seq/metis.cpp for the partition vector.
extern "C" int ffcuda_mesh_upload_distributed(ffcuda_ctx *ctx, int dim, int nv_owned, int nv_local, const double *xyz, int nt,
                                              const int32_t *conn, const int32_t *elab, int nbe, const int32_t *bconn,
                                              const int32_t *blab, const int32_t *belem, const int32_t *bface, const int64_t *gid,
                                              int nnbr, const int32_t *nbr, const int32_t *recv_off, const int32_t *recv_cnt,
                                              const int32_t *send_ptr, const int32_t *send_idx, ffcuda_mesh **out)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && out, "ffcuda_mesh_upload_distributed: null context/output");
    FF_REQUIRE(ctx->nranks == 1 || ctx->nccl_comm, "ffcuda_mesh_upload_distributed: call ffcuda_comm_init first");
    FF_REQUIRE(nv_owned > 0 && nv_owned <= nv_local, "every rank must own at least one vertex");
    FF_REQUIRE(nnbr >= 0 && nnbr <= ffcuda_mesh::MAXNBR, "at most 16 neighbour ranks");
    FF_REQUIRE(nnbr == 0 || (nbr && recv_off && recv_cnt && send_ptr && (send_idx || send_ptr[nnbr] == 0)), "halo arrays missing");
    int covered = nv_owned;
    for (int x = 0; x < nnbr; ++x) {
        FF_REQUIRE(nbr[x] >= 0 && nbr[x] < ctx->nranks && nbr[x] != ctx->rank, "bad neighbour rank");
        FF_REQUIRE(recv_off[x] == covered && recv_cnt[x] > 0, "ghost ranges must follow the owned vertices, in neighbour order, without gaps");
        covered += recv_cnt[x];
        FF_REQUIRE(send_ptr[x + 1] >= send_ptr[x], "send_ptr must be non-decreasing");
    }
    FF_REQUIRE(covered == nv_local, "ghost ranges do not cover the ghost vertices");
    for (int k = 0; k < (nnbr ? send_ptr[nnbr] : 0); ++k) FF_REQUIRE(send_idx[k] >= 0 && send_idx[k] < nv_owned, "send list entry is not an owned vertex");
    ffcuda_mesh *m = nullptr;
    const int rc = ffcuda_mesh_upload(ctx, dim, nv_local, xyz, nt, conn, elab, nbe, bconn, blab, belem, bface, &m);
    if (rc) throw FFError(ffcuda_last_error(ctx));
    std::unique_ptr<ffcuda_mesh> guard(m);
    ff_enter(ctx);
    m->nv_owned = nv_owned;
    m->distributed = ctx->nranks > 1;
    m->gid.alloc(nv_local);
    if (gid) FF_CUDA(cudaMemcpyAsync(m->gid.p, gid, (size_t)nv_local * 8, cudaMemcpyHostToDevice, ctx->stream));
    else FF_CUDA(cudaMemsetAsync(m->gid.p, 0, (size_t)nv_local * 8, ctx->stream));
    m->nnbr = nnbr;
    for (int x = 0; x < nnbr; ++x) {
        m->nbr[x] = nbr[x];
        m->recv_off[x] = recv_off[x];
        m->recv_cnt[x] = recv_cnt[x];
        m->send_off[x] = send_ptr[x];
        m->send_cnt[x] = send_ptr[x + 1] - send_ptr[x];
    }
    if (nnbr && send_ptr[nnbr] > 0) {
        m->send_idx.alloc((size_t)send_ptr[nnbr]);
        FF_CUDA(cudaMemcpyAsync(m->send_idx.p, send_idx, (size_t)send_ptr[nnbr] * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = guard.release();
    FF_API_END(ctx)
}

extern "C" int ffcuda_mesh_cube(ffcuda_ctx *ctx, int nx, int ny, int nz, ffcuda_mesh **out)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && out, "ffcuda_mesh_cube: null context/output");
    build_cube(ctx, nx, ny, nz, 0, 1, out);
    FF_API_END(ctx)
}

extern "C" int ffcuda_mesh_cube_distributed(ffcuda_ctx *ctx, int nx, int ny, int nz, ffcuda_mesh **out)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && out, "ffcuda_mesh_cube_distributed: null context/output");
    FF_REQUIRE(ctx->nranks == 1 || ctx->nccl_comm, "ffcuda_mesh_cube_distributed: call ffcuda_comm_init first");
    build_cube(ctx, nx, ny, nz, ctx->rank, ctx->nranks, out);
    FF_API_END(ctx)
}

extern "C" int ffcuda_mesh_local_to_global(ffcuda_mesh *m, int *nowned, int *nlocal, int64_t *gid)
{
    FF_API_BEGIN
    FF_REQUIRE(m, "null mesh");
    if (nowned) *nowned = m->nv_owned;
    if (nlocal) *nlocal = m->nv;
    if (gid) {
        ff_enter(m->ctx);
        if (m->gid.p) {
            FF_CUDA(ff_memcpy_sync(m->ctx, gid, m->gid.p, m->gid.bytes(), cudaMemcpyDeviceToHost));
        } else
            for (int i = 0; i < m->nv; ++i) gid[i] = i;
    }
    FF_API_END(m ? m->ctx : nullptr)
}

// ---------------------------------------------------------------------------------------------------
// square(nx,ny)
// ---------------------------------------------------------------------------------------------------
__global__ void k_square_vertices(double *__restrict__ xy, int nx, int ny)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int nx1 = nx + 1;
    if (p >= nx1 * (ny + 1)) return;
    int j = p / nx1, i = p - j * nx1;
    reinterpret_cast<double2 *>(xy)[p] = make_double2((double)i / nx, (double)j / ny);
}

__global__ void k_square_cells(int32_t *__restrict__ conn, int nx, int ny)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nx * ny) return;
    int j = c / nx, i = c - j * nx, nx1 = nx + 1;
    int i0 = i + j * nx1, i1 = i0 + 1, i2 = i1 + nx1, i3 = i2 - 1;
    int32_t *t = conn + 6 * (size_t)c;
    t[0] = i0; t[1] = i1; t[2] = i2;
    t[3] = i0; t[4] = i2; t[5] = i3;
}

__global__ void k_square_bedges(int32_t *__restrict__ bconn, int32_t *__restrict__ blab, int32_t *__restrict__ belem,
                                int32_t *__restrict__ bface, int nx, int ny)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    int nbe = 2 * (nx + ny), nx1 = nx + 1;
    if (e >= nbe) return;
    int v0, v1, lab, el, f;
    if (e < nx) { // bottom
        int i = e;
        v0 = i; v1 = i + 1; lab = 1; el = 2 * i; f = 2;
    } else if (e < nx + ny) { // right
        int j = e - nx;
        v0 = nx + j * nx1; v1 = v0 + nx1; lab = 2; el = 2 * ((nx - 1) + j * nx); f = 0;
    } else if (e < 2 * nx + ny) { // top
        int i = e - nx - ny;
        v0 = i + ny * nx1; v1 = v0 + 1; lab = 3; el = 2 * (i + (ny - 1) * nx) + 1; f = 0;
    } else { // left
        int j = e - 2 * nx - ny;
        v0 = j * nx1; v1 = v0 + nx1; lab = 4; el = 2 * (j * nx) + 1; f = 1;
    }
    bconn[2 * e] = v0; bconn[2 * e + 1] = v1;
    blab[e] = lab; belem[e] = el; bface[e] = f;
}

extern "C" int ffcuda_mesh_square(ffcuda_ctx *ctx, int nx, int ny, ffcuda_mesh **out)
{
    ffcuda_mesh *m = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(ctx && out, "ffcuda_mesh_square: null context/output");
    FF_REQUIRE(nx > 0 && ny > 0, "square sizes must be positive");
    FF_REQUIRE((int64_t)nx * ny * 2 < ((int64_t)1 << 27), "square too large");
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    m = new ffcuda_mesh();
    m->ctx = ctx;
    m->ref.set(ctx);
    m->dim = 2; m->vstride = 2;
    m->nv = (nx + 1) * (ny + 1);
    m->nv_owned = m->nv;
    m->nt = 2 * nx * ny;
    m->nbe = 2 * (nx + ny);
    m->xyz.alloc((size_t)m->nv * 2);
    m->conn.alloc((size_t)m->nt * 3);
    m->elab.alloc(m->nt);
    m->bconn.alloc((size_t)m->nbe * 2);
    m->blab.alloc(m->nbe); m->belem.alloc(m->nbe); m->bface.alloc(m->nbe);
    FF_CUDA(cudaMemsetAsync(m->elab.p, 0, m->elab.bytes(), st));
    ff_launch(ctx, "mesh_square_vertices", [&] { k_square_vertices<<<ff_blocks(m->nv, 256), 256, 0, st>>>(m->xyz.p, nx, ny); });
    ff_launch(ctx, "mesh_square_cells", [&] { k_square_cells<<<ff_blocks((size_t)nx * ny, 256), 256, 0, st>>>(m->conn.p, nx, ny); });
    ff_launch(ctx, "mesh_square_bedges", [&] {
        k_square_bedges<<<ff_blocks(m->nbe, 256), 256, 0, st>>>(m->bconn.p, m->blab.p, m->belem.p, m->bface.p, nx, ny);
    });
    FF_CUDA(cudaStreamSynchronize(st));
    *out = m;
    m = nullptr;
    FF_API_END((delete m, ctx))
}

extern "C" int ffcuda_mesh_info(ffcuda_mesh *m, int *dim, int *nv, int *nt, int *nbe)
{
    FF_API_BEGIN
    FF_REQUIRE(m, "null mesh");
    if (dim) *dim = m->dim;
    if (nv) *nv = m->nv;
    if (nt) *nt = m->nt;
    if (nbe) *nbe = m->nbe;
    FF_API_END(m ? m->ctx : nullptr)
}

extern "C" int ffcuda_mesh_download(ffcuda_mesh *m, double *xyz, int32_t *conn, int32_t *elab, int32_t *bconn,
                                    int32_t *blab, int32_t *belem, int32_t *bface)
{
    FF_API_BEGIN
    FF_REQUIRE(m, "null mesh");
    ffcuda_ctx *ctx = m->ctx;
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    if (xyz) {
        if (m->dim == 3) {
            DBuf<double> tmp;
            tmp.alloc((size_t)m->nv * 3);
            ff_launch(ctx, "mesh_unpad_xyz", [&] { k_unpad_xyz3<<<ff_blocks(m->nv, 256), 256, 0, st>>>(m->xyz.p, tmp.p, m->nv); });
            FF_CUDA(cudaMemcpyAsync(xyz, tmp.p, tmp.bytes(), cudaMemcpyDeviceToHost, st));
            FF_CUDA(cudaStreamSynchronize(st));
        } else
            FF_CUDA(cudaMemcpyAsync(xyz, m->xyz.p, m->xyz.bytes(), cudaMemcpyDeviceToHost, st));
    }
    if (conn) FF_CUDA(cudaMemcpyAsync(conn, m->conn.p, m->conn.bytes(), cudaMemcpyDeviceToHost, st));
    if (elab) FF_CUDA(cudaMemcpyAsync(elab, m->elab.p, m->elab.bytes(), cudaMemcpyDeviceToHost, st));
    if (bconn && m->nbe) FF_CUDA(cudaMemcpyAsync(bconn, m->bconn.p, m->bconn.bytes(), cudaMemcpyDeviceToHost, st));
    if (blab && m->nbe) FF_CUDA(cudaMemcpyAsync(blab, m->blab.p, m->blab.bytes(), cudaMemcpyDeviceToHost, st));
    if (belem && m->nbe) FF_CUDA(cudaMemcpyAsync(belem, m->belem.p, m->belem.bytes(), cudaMemcpyDeviceToHost, st));
    if (bface && m->nbe) FF_CUDA(cudaMemcpyAsync(bface, m->bface.p, m->bface.bytes(), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_API_END(m ? m->ctx : nullptr)
}

// ---------------------------------------------------------------------------------------------------------------
// Element adjacency on the device: GenericMesh::BuildAdj (femlib/GenericMesh.hpp:837-930).  adj[nea*k + i] = nea*k' + i'
// when face i of element k (the face opposite vertex i: nvfaceTet {3,2,1},{0,2,3},{3,1,0},{0,1,2}, femlib/Mesh3dn.cpp:72;
// edges {1,2},{2,0},{0,1} of a triangle) is face i' of element k', -1 on the boundary, -2 for a face shared by more than
// two elements (the reference breaks those lists the same way, :887-911).  The reference inserts every face in a hash
// table one after the other; here: one 64-bit hash per face (of its sorted vertices), a radix sort of (hash, face), and
// one thread per sorted position comparing the vertices themselves with its neighbours in the run (a hash collision
// between different faces only makes the run longer).
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <int NV>
__device__ __forceinline__ void face_verts(const int32_t *__restrict__ conn, int f, int (&v)[NV - 1])
{
    const int k = f / NV, i = f % NV;
    int o = 0;
#pragma unroll
    for (int a = 0; a < NV; ++a)
        if (a != i) v[o++] = conn[(size_t)k * NV + a];
    // sorted
#pragma unroll
    for (int x = 0; x < NV - 1; ++x)
#pragma unroll
        for (int y = x + 1; y < NV - 1; ++y)
            if (v[y] < v[x]) {
                const int t = v[x];
                v[x] = v[y];
                v[y] = t;
            }
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
template <int NV>
__global__ void k_face_keys(const int32_t *__restrict__ conn, int nfaces, unsigned long long *__restrict__ key, int32_t *__restrict__ val)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    int v[NV - 1];
    face_verts<NV>(conn, f, v);
    unsigned long long h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int x = 0; x < NV - 1; ++x) h = mix64(h ^ (unsigned long long)(unsigned)v[x]) + 0x632BE59BD9B4E019ull * (x + 1);
    key[f] = h;
    val[f] = f;
}
template <int NV>
__global__ void k_face_match(const int32_t *__restrict__ conn, int nfaces, const unsigned long long *__restrict__ key,
                             const int32_t *__restrict__ val, int32_t *__restrict__ adj)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nfaces) return;
    const int f = val[x];
    int v[NV - 1];
    face_verts<NV>(conn, f, v);
    const unsigned long long h = key[x];
    int mate = -1, nsame = 0;
    // the run of equal hashes around x (length 1 or 2 on a manifold mesh)
    int lo = x;
    while (lo > 0 && key[lo - 1] == h) --lo;
    for (int y = lo; y < nfaces && key[y] == h; ++y) {
        if (y == x) continue;
        int w[NV - 1];
        face_verts<NV>(conn, val[y], w);
        bool same = true;
#pragma unroll
        for (int c = 0; c < NV - 1; ++c) same = same && (w[c] == v[c]);
        if (same) {
            ++nsame;
            mate = val[y];
        }
    }
    adj[f] = nsame == 0 ? -1 : (nsame == 1 ? mate : -2);
}
} // namespace

// ---------------------------------------------------------------------------------------------------------------
// Boundary links of a tetrahedral mesh on the device: the boundary part of GenericMesh::BuildAdj
// (femlib/GenericMesh.hpp:914-1017).  Every boundary triangle gets its (element, face) and the orientation the reference
// leaves it with: a true boundary face takes the orientation of its element's face; an internal boundary face points to
// the element on the side where the face runs the other way, and between two regions the minority is turned round.
// Same sorted (hash, face) array as the adjacency; one thread per boundary triangle bisects it.
// ---------------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ int sort3_sign(int (&v)[3])
{ // SortArray<T,3> (femlib/HashTable.hpp:65-80): the sorted triple and the parity of the permutation
    int s = 1, t;
    if (v[0] > v[1]) { s = -s; t = v[0]; v[0] = v[1]; v[1] = t; }
    if (v[1] > v[2]) {
        s = -s; t = v[1]; v[1] = v[2]; v[2] = t;
        if (v[0] > v[1]) { s = -s; t = v[0]; v[0] = v[1]; v[1] = t; }
    }
    return s;
}
__device__ __forceinline__ unsigned long long face_hash3(const int (&v)[3])
{
    unsigned long long h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int x = 0; x < 3; ++x) h = mix64(h ^ (unsigned long long)(unsigned)v[x]) + 0x632BE59BD9B4E019ull * (x + 1);
    return h;
}
__device__ __forceinline__ int tet_face_sign(const int32_t *__restrict__ conn, int id, int (&v)[3])
{
    const int k = id >> 2, f = id & 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = conn[4 * (size_t)k + c_nvfaceTet[f][j]];
    return sort3_sign(v);
}

// other[b] = the second element face of an internal boundary face between two regions (else -1); reg[2b], reg[2b+1] its regions
__global__ void k_bface_link(const int32_t *__restrict__ conn, const int32_t *__restrict__ elab, int nfaces,
                             const unsigned long long *__restrict__ key, const int32_t *__restrict__ val, int nbe,
                             int32_t *__restrict__ bconn, int32_t *__restrict__ belem, int32_t *__restrict__ bface,
                             int32_t *__restrict__ other, int32_t *__restrict__ reg, int *__restrict__ flags)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbe) return;
    int v[3] = {bconn[3 * (size_t)b], bconn[3 * (size_t)b + 1], bconn[3 * (size_t)b + 2]};
    const int sens = sort3_sign(v);
    const unsigned long long h = face_hash3(v);
    int lo = 0, hi = nfaces;
    while (lo < hi) {
        const int mid = (int)(((long long)lo + hi) >> 1);
        if (key[mid] < h) lo = mid + 1; else hi = mid;
    }
    int first = -1, last = -1, cnt = 0;
    for (int y = lo; y < nfaces && key[y] == h; ++y) {
        int w[3];
        tet_face_sign(conn, val[y], w);
        if (w[0] == v[0] && w[1] == v[1] && w[2] == v[2]) {
            const int id = val[y];
            first = cnt ? min(first, id) : id;
            last = cnt ? max(last, id) : id;
            ++cnt;
        }
    }
    other[b] = -1;
    if (cnt == 0) { // not a face of the mesh
        belem[b] = bface[b] = -1;
        atomicOr(flags, 1);
        return;
    }
    int w[3], nk = first;
    if (cnt == 1) {
        if (tet_face_sign(conn, nk, w) != sens) { // the orientation of the element's face
            const int t = bconn[3 * (size_t)b];
            bconn[3 * (size_t)b] = bconn[3 * (size_t)b + 1];
            bconn[3 * (size_t)b + 1] = t;
        }
    } else {
        if (cnt > 2) atomicOr(flags, 4); // non-manifold: first and last of the run are taken
        int nkk = first;
        nk = last; // the later element is looked at first
        if (sens == tet_face_sign(conn, nk, w)) { const int t = nk; nk = nkk; nkk = t; }
        const int rk = elab[nk >> 2], rkk = elab[nkk >> 2];
        if (rk != rkk) {
            other[b] = nkk;
            reg[2 * (size_t)b] = rk;
            reg[2 * (size_t)b + 1] = rkk;
            atomicOr(flags, 2);
        }
    }
    belem[b] = nk >> 2;
    bface[b] = nk & 3;
}

__global__ void k_bface_turn(const int32_t *__restrict__ turn, int nturn, const int32_t *__restrict__ other,
                             int32_t *__restrict__ bconn, int32_t *__restrict__ belem, int32_t *__restrict__ bface)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nturn) return;
    const int b = turn[x];
    const int t = bconn[3 * (size_t)b];
    bconn[3 * (size_t)b] = bconn[3 * (size_t)b + 1];
    bconn[3 * (size_t)b + 1] = t;
    belem[b] = other[b] >> 2;
    bface[b] = other[b] & 3;
}

void sorted_face_keys(ffcuda_mesh *m, DBuf<unsigned long long> &k1, DBuf<int32_t> &v1)
{
    ffcuda_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int NV = m->dim + 1;
    FF_REQUIRE((int64_t)m->nt * NV < ((int64_t)1 << 31), "too many faces for 32-bit face ids");
    const int nf = m->nt * NV;
    DBuf<unsigned long long> k0;
    DBuf<int32_t> v0;
    k0.alloc(nf); k1.alloc(nf); v0.alloc(nf); v1.alloc(nf);
    ff_launch(ctx, "adj_face_keys", [&] {
        if (NV == 4) k_face_keys<4><<<ff_blocks(nf, 256), 256, 0, st>>>(m->conn.p, nf, k0.p, v0.p);
        else k_face_keys<3><<<ff_blocks(nf, 256), 256, 0, st>>>(m->conn.p, nf, k0.p, v0.p);
    });
    size_t tb = 0;
    FF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, v0.p, v1.p, nf, 0, 64, st));
    DBuf<unsigned char> tmp;
    tmp.alloc(tb + 16);
    ctx->launches++;
    FF_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k0.p, k1.p, v0.p, v1.p, nf, 0, 64, st));
    FF_CUDA(cudaStreamSynchronize(st)); // the temporaries go out of scope
}

// belem / bface / final orientation of the boundary triangles of a 3-D mesh whose bconn, blab are set
void boundary_links_3d(ffcuda_mesh *m)
{
    ffcuda_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    if (m->nbe == 0) return;
    DBuf<unsigned long long> k1;
    DBuf<int32_t> v1, other, reg;
    sorted_face_keys(m, k1, v1);
    const int nf = m->nt * 4, nbe = m->nbe;
    other.alloc(nbe); reg.alloc((size_t)2 * nbe);
    DBuf<int> flags;
    flags.alloc(1);
    FF_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int), st));
    ff_launch(ctx, "mesh_bface_link", [&] {
        k_bface_link<<<ff_blocks(nbe, 256), 256, 0, st>>>(m->conn.p, m->elab.p, nf, k1.p, v1.p, nbe, m->bconn.p, m->belem.p,
                                                          m->bface.p, other.p, reg.p, flags.p);
    });
    int hf = 0;
    FF_CUDA(ff_memcpy_sync(ctx, &hf, flags.p, sizeof(int), cudaMemcpyDeviceToHost));
    FF_REQUIRE(!(hf & 1), "a boundary element is not a face of the mesh");
    if (!(hf & 2)) return;
    // internal faces between two regions: per pair of regions the minority is turned round (GenericMesh.hpp:958-984);
    // a few faces, counted on the host
    std::vector<int32_t> ho(nbe), hr((size_t)2 * nbe);
    FF_CUDA(ff_memcpy_sync(ctx, ho.data(), other.p, (size_t)nbe * 4, cudaMemcpyDeviceToHost));
    FF_CUDA(ff_memcpy_sync(ctx, hr.data(), reg.p, (size_t)nbe * 8, cudaMemcpyDeviceToHost));
    std::unordered_map<uint64_t, std::pair<int64_t, int64_t>> cnt;
    auto pkey = [](int a, int b) { return ((uint64_t)(uint32_t)std::min(a, b) << 32) | (uint32_t)std::max(a, b); };
    for (int b = 0; b < nbe; ++b)
        if (ho[b] >= 0) {
            auto &c = cnt[pkey(hr[2 * (size_t)b], hr[2 * (size_t)b + 1])];
            (hr[2 * (size_t)b] > hr[2 * (size_t)b + 1] ? c.second : c.first)++;
        }
    bool mixed = false;
    for (auto &kv : cnt) mixed = mixed || (kv.second.first && kv.second.second);
    if (!mixed) return;
    std::vector<int32_t> turn;
    for (int b = 0; b < nbe; ++b)
        if (ho[b] >= 0) {
            const auto &c = cnt[pkey(hr[2 * (size_t)b], hr[2 * (size_t)b + 1])];
            const int sr = hr[2 * (size_t)b] > hr[2 * (size_t)b + 1] ? -1 : 1;
            if ((c.first < c.second && sr == 1) || (c.first > c.second && sr == -1)) turn.push_back(b);
        }
    if (turn.empty()) return;
    DBuf<int32_t> dturn;
    dturn.alloc(turn.size());
    FF_CUDA(cudaMemcpyAsync(dturn.p, turn.data(), turn.size() * 4, cudaMemcpyHostToDevice, st));
    ff_launch(ctx, "mesh_bface_turn", [&] {
        k_bface_turn<<<ff_blocks(turn.size(), 256), 256, 0, st>>>(dturn.p, (int)turn.size(), other.p, m->bconn.p, m->belem.p, m->bface.p);
    });
    FF_CUDA(cudaStreamSynchronize(st));
}
} // namespace

// ---------------------------------------------------------------------------------------------------------------
// buildlayers (fflib/msh3.cpp:895-1757): the layered tetrahedral mesh over a 2-D mesh that is on the device.  2-D vertex i
// with ni[i] layers becomes the column of 3-D vertices first[i] .. first[i] + ni[i]; at level s (taken from the top) a
// column stands at position (s ni)/Nmax, so columns with fewer layers repeat positions and the prism over a triangle
// degenerates to a pyramid (2 tets) or a tetrahedron; quadrilateral faces are cut by the diagonal through the largest 3-D
// vertex number, which makes neighbouring prisms agree.  The reference walks triangles and levels one after the other; here
// every (triangle, level) and (boundary edge, level) is a thread: count, scan, fill — same element and face order.
// ---------------------------------------------------------------------------------------------------------------
namespace {
__constant__ int c_pentaCut[6][12] = {{0, 5, 1, 2, 0, 4, 1, 5, 0, 5, 3, 4}, {0, 5, 1, 2, 0, 3, 1, 5, 1, 5, 3, 4},
                                      {0, 3, 1, 2, 1, 5, 2, 3, 1, 5, 3, 4}, {0, 4, 1, 2, 0, 4, 2, 5, 0, 5, 3, 4},
                                      {0, 4, 1, 2, 0, 4, 2, 3, 2, 5, 3, 4}, {0, 3, 1, 2, 1, 4, 2, 3, 2, 5, 3, 4}}; // dpent1 :1694-1701, 0-based
__constant__ int c_pentaSel[8] = {0, -1, 1, 2, 3, 4, -1, 5};                                                   // pdd :1693

struct LayerMaps { // (old,new) pairs: region, labelmid, labelup, labeldown; the last pair of a label wins
    const int32_t *p[4];
    int n[4];
};
__device__ __forceinline__ int map_label(const LayerMaps &M, int which, int lab)
{
    int out = lab;
    for (int k = 0; k < M.n[which]; ++k)
        if (M.p[which][2 * k] == lab) out = M.p[which][2 * k + 1];
    return out;
}

__global__ void k_layer_vertices(const double *__restrict__ xy, const int32_t *__restrict__ first, const int32_t *__restrict__ ni,
                                 const double *__restrict__ zmin, const double *__restrict__ zmax, int nv2, int nlayer,
                                 double *__restrict__ xyz4)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(t / (size_t)(nlayer + 1)), j = (int)(t - (size_t)i * (nlayer + 1));
    if (i >= nv2) return;
    const int N = ni[i];
    if (j > N) return;
    const double dz = N == 0 ? 0. : __ddiv_rn(__dsub_rn(zmax[i], zmin[i]), (double)N);
    reinterpret_cast<double4 *>(xyz4)[first[i] + j] =
        make_double4(xy[2 * (size_t)i], xy[2 * (size_t)i + 1], __dadd_rn(zmin[i], __dmul_rn(dz, (double)j)), 0.0);
}

// tets of the prism P[0..2] (lower) / P[3..5] (upper); returns their number
__device__ __forceinline__ int prism_tets(const int (&P)[6], int (&o)[3][4])
{
    const int cas = (P[0] != P[3]) + 2 * (P[1] != P[4]) + 4 * (P[2] != P[5]);
    if (cas == 0) return 0;
    o[0][0] = P[0]; o[0][1] = P[1]; o[0][2] = P[2];
    if (cas == 1 || cas == 2 || cas == 4) {
        o[0][3] = P[cas == 1 ? 3 : (cas == 2 ? 4 : 5)];
        return 1;
    }
    if (cas != 7) {
        const int a = cas == 6 ? 1 : 0, b = cas == 3 ? 1 : 2;
        const bool one = max(P[a], P[b + 3]) > max(P[b], P[a + 3]);
        o[0][3] = one ? P[b + 3] : P[a + 3];
        o[1][0] = P[5]; o[1][1] = P[4]; o[1][2] = P[3]; o[1][3] = one ? P[a] : P[b];
        return 2;
    }
    const int i1 = max(P[0], P[5]) > max(P[2], P[3]) ? 0 : 1;
    const int i2 = max(P[0], P[4]) > max(P[1], P[3]) ? 0 : 1;
    const int i3 = max(P[1], P[5]) > max(P[2], P[4]) ? 0 : 1;
    const int cut = c_pentaSel[i1 + 2 * i2 + 4 * i3];
    if (cut < 0) return -1;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) o[t][q] = P[c_pentaCut[cut][4 * t + q]];
    return 3;
}

template <int PASS>
__global__ void k_layer_tets(const int32_t *__restrict__ tri, const int32_t *__restrict__ trilab, const int32_t *__restrict__ first,
                             const int32_t *__restrict__ ni, int nt2, int nlayer, const LayerMaps M, int32_t *__restrict__ cnt,
                             const int32_t *__restrict__ off, int32_t *__restrict__ conn, int32_t *__restrict__ elab, int *__restrict__ bad)
{
    const size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= (size_t)nt2 * nlayer) return;
    const int k = (int)(it / nlayer), s = nlayer - 1 - (int)(it - (size_t)k * nlayer);
    int P[6], o[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int v = tri[3 * (size_t)k + j], N = ni[v], f = first[v];
        P[j] = f + (s * N) / nlayer;
        P[j + 3] = f + ((s + 1) * N) / nlayer;
    }
    const int c = prism_tets(P, o);
    if (c < 0) {
        atomicOr(bad, 1);
        if (PASS == 0) cnt[it] = 0;
        return;
    }
    if (PASS == 0) {
        cnt[it] = c;
        return;
    }
    const int lab = map_label(M, 0, trilab[k]);
    const size_t base = (size_t)off[it];
    for (int t = 0; t < c; ++t) {
        reinterpret_cast<int4 *>(conn)[base + t] = make_int4(o[t][0], o[t][1], o[t][2], o[t][3]);
        elab[base + t] = lab;
    }
}

// faces at zmax (first nt2) and zmin (next nt2, orientation reversed): :1111-1150
__global__ void k_layer_caps(const int32_t *__restrict__ tri, const int32_t *__restrict__ trilab, const int32_t *__restrict__ first,
                             int nt2, const LayerMaps M, int32_t *__restrict__ bconn, int32_t *__restrict__ blab)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nt2) return;
    const int lab = trilab[k];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int v = tri[3 * (size_t)k + j];
        bconn[3 * (size_t)k + j] = first[v + 1] - 1;
        bconn[3 * ((size_t)nt2 + k) + 2 - j] = first[v];
    }
    blab[k] = map_label(M, 2, lab);
    blab[(size_t)nt2 + k] = map_label(M, 3, lab);
}

// lateral faces over the boundary edges: :1154-1316
template <int PASS>
__global__ void k_layer_sides(const int32_t *__restrict__ tri, const int32_t *__restrict__ bedge_lab, const int32_t *__restrict__ bedge_elem,
                              const int32_t *__restrict__ bedge_face, const int32_t *__restrict__ first, const int32_t *__restrict__ ni,
                              int nbe2, int nlayer, const LayerMaps M, int32_t *__restrict__ cnt, const int32_t *__restrict__ off,
                              int base0, int32_t *__restrict__ bconn, int32_t *__restrict__ blab)
{
    const size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= (size_t)nbe2 * nlayer) return;
    const int e = (int)(it / nlayer), s = nlayer - 1 - (int)(it - (size_t)e * nlayer);
    const int el = bedge_elem[e], f = bedge_face[e];
    const int i1 = tri[3 * (size_t)el + (f + 1) % 3], i2 = tri[3 * (size_t)el + (f + 2) % 3]; // VerticesNumberOfEdge femlib/fem.hpp:561
    const int a = first[i1] + (s * ni[i1]) / nlayer, d = first[i1] + ((s + 1) * ni[i1]) / nlayer;
    const int b = first[i2] + (s * ni[i2]) / nlayer, c = first[i2] + ((s + 1) * ni[i2]) / nlayer;
    const int type = (a != d ? 1 : 0) + (b != c ? 2 : 0);
    const int n = type == 0 ? 0 : (type == 3 ? 2 : 1);
    if (PASS == 0) {
        cnt[it] = n;
        return;
    }
    if (n == 0) return;
    const int lab = map_label(M, 1, bedge_lab[e]);
    int32_t *o = bconn + 3 * ((size_t)base0 + off[it]);
    int32_t *l = blab + (size_t)base0 + off[it];
    if (type == 1) { o[0] = a; o[1] = b; o[2] = d; l[0] = lab; }
    else if (type == 2) { o[0] = a; o[1] = b; o[2] = c; l[0] = lab; }
    else {
        const bool one = max(a, c) > max(b, d);
        o[0] = a; o[1] = b; o[2] = one ? c : d;
        o[3] = c; o[4] = d; o[5] = one ? a : b;
        l[0] = l[1] = lab;
    }
}
} // namespace

extern "C" int ffcuda_mesh_buildlayers(ffcuda_mesh *m2, int nlayer, const int32_t *ni, const double *zmin, const double *zmax,
                                       int nreg, const int32_t *regmap, int nmid, const int32_t *midmap, int nup,
                                       const int32_t *upmap, int ndown, const int32_t *downmap, ffcuda_mesh **out)
{
    FF_API_BEGIN
    FF_REQUIRE(m2 && out, "ffcuda_mesh_buildlayers: null mesh/output");
    FF_REQUIRE(m2->dim == 2 && !m2->distributed, "ffcuda_mesh_buildlayers: the base mesh must be a 2-D mesh on one device");
    FF_REQUIRE(nlayer > 0 && nlayer < (1 << 15), "ffcuda_mesh_buildlayers: the number of layers must lie in [1, 32767]");
    FF_REQUIRE(nreg >= 0 && nmid >= 0 && nup >= 0 && ndown >= 0, "negative map size");
    FF_REQUIRE((!nreg || regmap) && (!nmid || midmap) && (!nup || upmap) && (!ndown || downmap), "label map missing");
    FF_REQUIRE(m2->nbe == 0 || (m2->belem.p && m2->bface.p), "the base mesh has no boundary links");
    ffcuda_ctx *ctx = m2->ctx;
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    const int nv2 = m2->nv, nt2 = m2->nt, nbe2 = m2->nbe;
    // columns: first[i] (host prefix sum over the 2-D vertices; fflib/msh3.cpp:1016-1052)
    std::vector<int32_t> hni(nv2), hfirst((size_t)nv2 + 1);
    int64_t acc = 0;
    for (int i = 0; i < nv2; ++i) {
        hni[i] = ni ? ni[i] : nlayer;
        FF_REQUIRE(hni[i] >= 0 && hni[i] <= nlayer, "ffcuda_mesh_buildlayers: ni[] must lie in [0, nlayer]");
        hfirst[i] = (int32_t)acc;
        acc += hni[i] + 1;
        FF_REQUIRE(acc < ((int64_t)1 << 31), "too many vertices");
    }
    hfirst[nv2] = (int32_t)acc;
    FF_REQUIRE((int64_t)nt2 * nlayer * 3 < ((int64_t)1 << 27), "layered mesh too large for one device (limit 2^27 tets)");
    std::vector<double> hz((size_t)2 * nv2);
    for (int i = 0; i < nv2; ++i) {
        hz[i] = zmin ? zmin[i] : 0.;
        hz[(size_t)nv2 + i] = zmax ? zmax[i] : 1.;
    }
    DBuf<int32_t> dni, dfirst, dmaps;
    DBuf<double> dz;
    dni.alloc(nv2); dfirst.alloc((size_t)nv2 + 1); dz.alloc((size_t)2 * nv2);
    FF_CUDA(cudaMemcpyAsync(dni.p, hni.data(), (size_t)nv2 * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(dfirst.p, hfirst.data(), ((size_t)nv2 + 1) * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(dz.p, hz.data(), (size_t)2 * nv2 * 8, cudaMemcpyHostToDevice, st));
    std::vector<int32_t> hmaps;
    const int mn[4] = {nreg, nmid, nup, ndown};
    const int32_t *mp[4] = {regmap, midmap, upmap, downmap};
    size_t moff[4];
    for (int w = 0; w < 4; ++w) {
        moff[w] = hmaps.size();
        hmaps.insert(hmaps.end(), mp[w], mp[w] + (mn[w] ? 2 * (size_t)mn[w] : 0));
    }
    dmaps.alloc(std::max<size_t>(hmaps.size(), 1));
    if (!hmaps.empty()) FF_CUDA(cudaMemcpyAsync(dmaps.p, hmaps.data(), hmaps.size() * 4, cudaMemcpyHostToDevice, st));
    LayerMaps M;
    for (int w = 0; w < 4; ++w) {
        M.p[w] = dmaps.p + moff[w];
        M.n[w] = mn[w];
    }
    std::unique_ptr<ffcuda_mesh> m(new ffcuda_mesh());
    m->ctx = ctx;
    m->ref.set(ctx);
    m->dim = 3; m->vstride = 4;
    m->nv = m->nv_owned = (int)acc;
    m->xyz.alloc((size_t)m->nv * 4);
    ff_launch(ctx, "mesh_layer_vertices", [&] {
        k_layer_vertices<<<ff_blocks((size_t)nv2 * (nlayer + 1), 256), 256, 0, st>>>(m2->xyz.p, dfirst.p, dni.p, dz.p, dz.p + nv2, nv2, nlayer, m->xyz.p);
    });
    // tetrahedra: count per (triangle, level), scan, fill
    const size_t nit = (size_t)nt2 * nlayer;
    DBuf<int32_t> cnt, off;
    cnt.alloc(nit); off.alloc(nit);
    DBuf<int> bad;
    bad.alloc(1);
    FF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    ff_launch(ctx, "mesh_layer_tets_count", [&] {
        k_layer_tets<0><<<ff_blocks(nit, 256), 256, 0, st>>>(m2->conn.p, m2->elab.p, dfirst.p, dni.p, nt2, nlayer, M, cnt.p, nullptr, nullptr, nullptr, bad.p);
    });
    int64_t tot = 0;
    ff_exclusive_scan_i32(ctx, cnt.p, off.p, nit, &tot);
    int hbad = 0;
    FF_CUDA(ff_memcpy_sync(ctx, &hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
    FF_REQUIRE(!hbad, "internal: a prism has no conforming cut");
    FF_REQUIRE(tot > 0, "ffcuda_mesh_buildlayers: no tetrahedron (every column has zero layers)");
    m->nt = (int)tot;
    m->conn.alloc((size_t)m->nt * 4);
    m->elab.alloc(m->nt);
    ff_launch(ctx, "mesh_layer_tets", [&] {
        k_layer_tets<1><<<ff_blocks(nit, 256), 256, 0, st>>>(m2->conn.p, m2->elab.p, dfirst.p, dni.p, nt2, nlayer, M, nullptr, off.p, m->conn.p, m->elab.p, bad.p);
    });
    // boundary: caps, then the lateral faces
    const size_t nis = (size_t)nbe2 * nlayer;
    DBuf<int32_t> scnt, soff;
    int64_t stot = 0;
    if (nis) {
        scnt.alloc(nis); soff.alloc(nis);
        ff_launch(ctx, "mesh_layer_sides_count", [&] {
            k_layer_sides<0><<<ff_blocks(nis, 256), 256, 0, st>>>(m2->conn.p, m2->blab.p, m2->belem.p, m2->bface.p, dfirst.p, dni.p, nbe2, nlayer, M,
                                                                   scnt.p, nullptr, 0, nullptr, nullptr);
        });
        ff_exclusive_scan_i32(ctx, scnt.p, soff.p, nis, &stot);
    }
    m->nbe = 2 * nt2 + (int)stot;
    m->bconn.alloc((size_t)m->nbe * 3);
    m->blab.alloc(m->nbe); m->belem.alloc(m->nbe); m->bface.alloc(m->nbe);
    ff_launch(ctx, "mesh_layer_caps", [&] {
        k_layer_caps<<<ff_blocks(nt2, 256), 256, 0, st>>>(m2->conn.p, m2->elab.p, dfirst.p, nt2, M, m->bconn.p, m->blab.p);
    });
    if (stot)
        ff_launch(ctx, "mesh_layer_sides", [&] {
            k_layer_sides<1><<<ff_blocks(nis, 256), 256, 0, st>>>(m2->conn.p, m2->blab.p, m2->belem.p, m2->bface.p, dfirst.p, dni.p, nbe2, nlayer, M,
                                                                   nullptr, soff.p, 2 * nt2, m->bconn.p, m->blab.p);
        });
    boundary_links_3d(m.get()); // what Mesh3's BuildAdj does to the boundary triangles (orientation, element, face)
    FF_CUDA(cudaStreamSynchronize(st));
    *out = m.release();
    FF_API_END(m2 ? m2->ctx : nullptr)
}

extern "C" int ffcuda_mesh_adjacency(ffcuda_mesh *m, int32_t *adj /* host, (dim+1)*nt, may be NULL */, const int32_t **d_adj /* may be NULL */)
{
    FF_API_BEGIN
    FF_REQUIRE(m, "null mesh");
    ffcuda_ctx *ctx = m->ctx;
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    const int NV = m->dim + 1;
    FF_REQUIRE((int64_t)m->nt * NV < ((int64_t)1 << 31), "too many faces for 32-bit face ids");
    const int nf = m->nt * NV;
    if (!m->adj.p && nf > 0) {
        DBuf<unsigned long long> k1;
        DBuf<int32_t> v1;
        sorted_face_keys(m, k1, v1);
        m->adj.alloc(nf);
        ff_launch(ctx, "adj_face_match", [&] {
            if (NV == 4) k_face_match<4><<<ff_blocks(nf, 256), 256, 0, st>>>(m->conn.p, nf, k1.p, v1.p, m->adj.p);
            else k_face_match<3><<<ff_blocks(nf, 256), 256, 0, st>>>(m->conn.p, nf, k1.p, v1.p, m->adj.p);
        });
        FF_CUDA(cudaStreamSynchronize(st)); // k1, v1 go out of scope
    }
    if (adj && nf > 0) FF_CUDA(cudaMemcpyAsync(adj, m->adj.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    if (d_adj) *d_adj = m->adj.p;
    FF_API_END(m ? m->ctx : nullptr)
}

extern "C" void ffcuda_mesh_destroy(ffcuda_mesh *m)
{
    if (!m) return;
    ff_enter(m->ctx);
    delete m;
}
