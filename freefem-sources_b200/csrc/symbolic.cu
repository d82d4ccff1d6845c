// symbolic.cu — KERNEL 1: symbolic sparsity from the element -> dof map.
//
// Output = the CSR pattern MatriceMorse holds after CSR(): for every element, every couple of its dofs is an
// entry (HashMatrix::operator+=(MatriceElementaire&), femlib/HashMatrix.cpp:1310-1317), rows/columns sorted
// (Sortij :671, Buildp :993).  Built at NODE level (a vector space [P,P,P] has dof = node*ncomp + c, so its
// pattern is the node pattern with every entry replaced by a dense ncomp x ncomp block) and then expanded.
//
// By-products kept for the numeric phase (row-owner gather, assemble.cu):
//   inc    : node -> sorted list of (element, local node) incidences
//   pos    : for every incidence and every node b of that element, the position of b in the node row
//   diagpos: position of A(i,i)
//
// Steps: count incidences (integer atomics: order-independent result) -> scan -> fill -> per-node sort
// (restores a deterministic order) -> per-row sort+unique of the candidate columns in shared memory (bitonic,
// one warp per row) for the row lengths -> scan -> same again writing columns and positions.
#include "common.cuh"
#include <climits>

__global__ void k_count_inc(const int32_t *__restrict__ e2n, size_t nitems, int nrows, int32_t *__restrict__ cnt)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    int node = e2n[i];
    if (node < nrows) atomicAdd(&cnt[node], 1);
}

__global__ void k_fill_inc(const int32_t *__restrict__ e2n, size_t nitems, int nloc, int nrows,
                           const int32_t *__restrict__ incptr, int32_t *__restrict__ cursor, uint32_t *__restrict__ inc)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    int node = e2n[i];
    if (node >= nrows) return;
    uint32_t k = (uint32_t)(i / nloc), a = (uint32_t)(i - (size_t)k * nloc);
    int slot = atomicAdd(&cursor[node], 1);
    inc[(size_t)incptr[node] + slot] = (k << 4) | a;
}

// one thread per node: insertion sort of its incidence list (ascending element index), and max list length
__global__ void k_sort_inc(const int32_t *__restrict__ incptr, uint32_t *__restrict__ inc, int nrows, int32_t *__restrict__ maxinc)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (i < nrows) {
        int b = incptr[i];
        len = incptr[i + 1] - b;
        uint32_t *p = inc + b;
        for (int x = 1; x < len; ++x) {
            uint32_t v = p[x];
            int y = x - 1;
            while (y >= 0 && p[y] > v) {
                p[y + 1] = p[y];
                --y;
            }
            p[y + 1] = v;
        }
    }
    // warp max then one atomic per warp
    for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(maxinc, len);
}

// one warp per node row.  PASS 0: row length.  PASS 1: columns, positions, diagonal position.
template <int PASS, typename PosT>
__global__ void __launch_bounds__(256) k_row_pattern(const int32_t *__restrict__ e2n, int nloc, int nlocp, int nrows, int cap,
                                                     const int32_t *__restrict__ incptr, const uint32_t *__restrict__ inc,
                                                     int32_t *__restrict__ rowlen, const int32_t *__restrict__ nrowptr,
                                                     int32_t *__restrict__ ncol, PosT *__restrict__ pos,
                                                     int32_t *__restrict__ diagnode, int32_t *__restrict__ maxrow)
{
    extern __shared__ int32_t smem[];
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t *buf = smem + (size_t)w * cap;
    int localmax = 0;
    for (int row = blockIdx.x * warps + w; row < nrows; row += gridDim.x * warps) {
        const int ib = incptr[row], ninc = incptr[row + 1] - ib;
        const int ncand = ninc * nloc;
        int m = 32;
        while (m < ncand) m <<= 1;
        for (int x = lane; x < m; x += 32) {
            int v = INT_MAX;
            if (x < ncand) {
                int e = x / nloc, b = x - e * nloc;
                uint32_t ka = inc[ib + e];
                v = e2n[(size_t)(ka >> 4) * nloc + b];
            }
            buf[x] = v;
        }
        __syncwarp();
        // bitonic sort of m keys by one warp
        for (int k = 2; k <= m; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int x = lane; x < m; x += 32) {
                    int y = x ^ j;
                    if (y > x) {
                        int a = buf[x], b = buf[y];
                        bool up = (x & k) == 0;
                        if ((a > b) == up) {
                            buf[x] = b;
                            buf[y] = a;
                        }
                    }
                }
                __syncwarp();
            }
        // unique compaction in place (write index never passes the read index)
        int nu = 0;
        for (int base = 0; base < m; base += 32) {
            int x = base + lane;
            int v = buf[x];
            int prev = x > 0 ? buf[x - 1] : -1;
            bool keep = (v != INT_MAX) && (x == 0 || v != prev);
            unsigned msk = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) buf[nu + __popc(msk & ((1u << lane) - 1))] = v;
            nu += __popc(msk);
            __syncwarp();
        }
        if (PASS == 0) {
            if (lane == 0) rowlen[row] = nu;
            localmax = max(localmax, nu);
        } else {
            const int rb = nrowptr[row];
            for (int x = lane; x < nu; x += 32) {
                int v = buf[x];
                ncol[(size_t)rb + x] = v;
                if (v == row) diagnode[row] = x;
            }
            // positions of the nodes of every incident element inside this row
            for (int x = lane; x < ncand; x += 32) {
                int e = x / nloc, b = x - e * nloc;
                uint32_t ka = inc[ib + e];
                int v = e2n[(size_t)(ka >> 4) * nloc + b];
                int lo = 0, hi = nu - 1;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (buf[mid] < v) lo = mid + 1;
                    else hi = mid;
                }
                pos[(size_t)(ib + e) * nlocp + b] = (PosT)lo;
            }
        }
        __syncwarp();
    }
    if (PASS == 0) {
        for (int o = 16; o; o >>= 1) localmax = max(localmax, __shfl_xor_sync(0xffffffffu, localmax, o));
        if (lane == 0 && localmax > 0) atomicMax(maxrow, localmax);
    }
}

// node-level CSR -> dof-level CSR for ncomp > 1
__global__ void k_expand_rowptr(const int32_t *__restrict__ nrowptr, int nrows, int nc, int32_t *__restrict__ rowptr,
                                const int32_t *__restrict__ diagnode, int32_t *__restrict__ diagpos)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nrows) return;
    if (i == nrows) {
        rowptr[(size_t)nrows * nc] = nc * nc * nrowptr[nrows];
        return;
    }
    int b = nrowptr[i], L = nrowptr[i + 1] - b;
    for (int c = 0; c < nc; ++c) {
        int r = nc * nc * b + c * nc * L;
        rowptr[(size_t)i * nc + c] = r;
        diagpos[(size_t)i * nc + c] = r + diagnode[i] * nc + c;
    }
}

__global__ void k_expand_colind(const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ ncol, int nrows, int nc,
                                int32_t *__restrict__ colind)
{
    // one warp per node row
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= nrows) return;
    int b = nrowptr[row], L = nrowptr[row + 1] - b;
    size_t base = (size_t)nc * nc * b;
    int W = nc * L; // entries per dof row
    for (int x = lane; x < nc * W; x += 32) {
        int c = x / W, r = x - c * W;
        int p = r / nc, d = r - p * nc;
        colind[base + (size_t)c * W + r] = ncol[b + p] * nc + d;
    }
}

__global__ void k_diagpos_scalar(const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ diagnode, int nrows,
                                 int32_t *__restrict__ diagpos)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nrows) diagpos[i] = nrowptr[i] + diagnode[i];
}

template <typename PosT>
static void run_row_pattern(ffcuda_ctx *ctx, ffcuda_pattern *P, const int32_t *e2n, int nloc, int cap, int warps, size_t shmem,
                            int blocks, int32_t *rowlen, int32_t *diagnode, int32_t *d_max, PosT *pos, int pass)
{
    cudaStream_t st = ctx->stream;
    if (pass == 0) {
        FF_CUDA(cudaFuncSetAttribute(k_row_pattern<0, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        ff_launch(ctx, "sym_row_count", [&] {
            k_row_pattern<0, PosT><<<blocks, warps * 32, shmem, st>>>(e2n, nloc, P->nlocp, P->nrows_node, cap, P->incptr.p, P->inc.p,
                                                                      rowlen, nullptr, nullptr, nullptr, nullptr, d_max);
        });
    } else {
        FF_CUDA(cudaFuncSetAttribute(k_row_pattern<1, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        ff_launch(ctx, "sym_row_fill", [&] {
            k_row_pattern<1, PosT><<<blocks, warps * 32, shmem, st>>>(e2n, nloc, P->nlocp, P->nrows_node, cap, P->incptr.p, P->inc.p,
                                                                      nullptr, P->nrowptr.p, P->ncol.p, pos, diagnode, nullptr);
        });
    }
}

extern "C" int ffcuda_symbolic(ffcuda_space *s, ffcuda_pattern **out)
{
    ffcuda_pattern *P = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(s && out, "ffcuda_symbolic: null space/output");
    ffcuda_ctx *ctx = s->ctx;
    FF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int nt = s->mesh->nt, nloc = s->nloc, nc = s->ncomp;
    const int nrows = s->nnodes_owned;
    const size_t nitems = (size_t)nt * nloc;
    P = new ffcuda_pattern();
    P->space = s;
    P->ctx = ctx;
    P->nrows_node = nrows;
    P->ncols_node = s->nnodes;
    P->ncomp = nc;
    P->n = nrows * nc;

    // --- node -> element incidence
    DBuf<int32_t> cnt;
    cnt.alloc((size_t)nrows + 1);
    FF_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
    ff_launch(ctx, "sym_count_inc", [&] { k_count_inc<<<ff_blocks(nitems, 256), 256, 0, st>>>(s->e2n, nitems, nrows, cnt.p); });
    P->incptr.alloc((size_t)nrows + 1);
    int64_t ninc = 0;
    ff_exclusive_scan_i32(ctx, cnt.p, P->incptr.p, (size_t)nrows + 1, &ninc);
    FF_REQUIRE(ninc < ((int64_t)1 << 31), "incidence table exceeds int32");
    P->inc.alloc((size_t)ninc);
    FF_CUDA(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
    ff_launch(ctx, "sym_fill_inc", [&] {
        k_fill_inc<<<ff_blocks(nitems, 256), 256, 0, st>>>(s->e2n, nitems, nloc, nrows, P->incptr.p, cnt.p, P->inc.p);
    });
    DBuf<int32_t> d_max;
    d_max.alloc(2);
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, d_max.bytes(), st));
    ff_launch(ctx, "sym_sort_inc", [&] { k_sort_inc<<<ff_blocks(nrows, 128), 128, 0, st>>>(P->incptr.p, P->inc.p, nrows, d_max.p); });
    int32_t h_max[2] = {0, 0};
    FF_CUDA(cudaMemcpyAsync(h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    const int maxinc = h_max[0];
    FF_REQUIRE(maxinc > 0, "no element touches any owned node");

    // --- row lengths
    int cap = 32;
    while (cap < maxinc * nloc) cap <<= 1;
    FF_REQUIRE((size_t)cap * 4 <= 200 * 1024, "a node has too many incident elements for the shared-memory sort");
    int warps = 8;
    while (warps > 1 && (size_t)warps * cap * 4 > 96 * 1024) warps >>= 1;
    size_t shmem = (size_t)warps * cap * 4;
    int blocks = min(ff_blocks((size_t)nrows, warps), ctx->sm_count * 16);
    DBuf<int32_t> rowlen, diagnode;
    rowlen.alloc((size_t)nrows + 1);
    diagnode.alloc((size_t)nrows);
    FF_CUDA(cudaMemsetAsync(rowlen.p, 0, rowlen.bytes(), st));
    run_row_pattern<uint8_t>(ctx, P, s->e2n, nloc, cap, warps, shmem, blocks, rowlen.p, diagnode.p, d_max.p + 1, nullptr, 0);
    P->nrowptr.alloc((size_t)nrows + 1);
    int64_t nnzn = 0;
    ff_exclusive_scan_i32(ctx, rowlen.p, P->nrowptr.p, (size_t)nrows + 1, &nnzn);
    FF_CUDA(cudaMemcpyAsync(h_max + 1, d_max.p + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    P->maxrow_node = h_max[1];
    P->nnz_node = nnzn;
    P->nnz = nnzn * nc * nc;
    FF_REQUIRE(P->nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros (int32 CSR, like MatriceMorse)");
    rowlen.release();

    // --- columns + positions
    P->ncol.alloc((size_t)nnzn);
    P->nlocp = (s->order == 1) ? 4 : nloc;
    if (P->maxrow_node <= 255) {
        P->pos8.alloc((size_t)ninc * P->nlocp);
        run_row_pattern<uint8_t>(ctx, P, s->e2n, nloc, cap, warps, shmem, blocks, nullptr, diagnode.p, nullptr, P->pos8.p, 1);
    } else {
        P->pos16.alloc((size_t)ninc * P->nlocp);
        run_row_pattern<uint16_t>(ctx, P, s->e2n, nloc, cap, warps, shmem, blocks, nullptr, diagnode.p, nullptr, P->pos16.p, 1);
    }

    // --- dof-level CSR
    P->diagpos.alloc((size_t)P->n);
    if (nc == 1) {
        P->rowptr = P->nrowptr.p;
        P->colind = P->ncol.p;
        ff_launch(ctx, "sym_diagpos", [&] { k_diagpos_scalar<<<ff_blocks(nrows, 256), 256, 0, st>>>(P->nrowptr.p, diagnode.p, nrows, P->diagpos.p); });
    } else {
        P->rowptr_own.alloc((size_t)P->n + 1);
        P->colind_own.alloc((size_t)P->nnz);
        ff_launch(ctx, "sym_expand_rowptr", [&] {
            k_expand_rowptr<<<ff_blocks((size_t)nrows + 1, 256), 256, 0, st>>>(P->nrowptr.p, nrows, nc, P->rowptr_own.p, diagnode.p, P->diagpos.p);
        });
        ff_launch(ctx, "sym_expand_colind", [&] {
            k_expand_colind<<<ff_blocks((size_t)nrows * 32, 256), 256, 0, st>>>(P->nrowptr.p, P->ncol.p, nrows, nc, P->colind_own.p);
        });
        P->rowptr = P->rowptr_own.p;
        P->colind = P->colind_own.p;
    }
    FF_CUDA(cudaStreamSynchronize(st));
    s->last_pattern = P;
    *out = P;
    P = nullptr;
    FF_API_END((delete P, s ? s->ctx : nullptr))
}

extern "C" int ffcuda_pattern_info(ffcuda_pattern *p, int *n, int64_t *nnz)
{
    FF_API_BEGIN
    FF_REQUIRE(p, "null pattern");
    if (n) *n = p->n;
    if (nnz) *nnz = p->nnz;
    FF_API_END(p ? p->ctx : nullptr)
}

extern "C" int ffcuda_pattern_download(ffcuda_pattern *p, int32_t *rowptr, int32_t *colind)
{
    FF_API_BEGIN
    FF_REQUIRE(p, "null pattern");
    FF_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    if (rowptr) FF_CUDA(cudaMemcpyAsync(rowptr, p->rowptr, ((size_t)p->n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (colind) FF_CUDA(cudaMemcpyAsync(colind, p->colind, (size_t)p->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_API_END(p ? p->ctx : nullptr)
}

ffcuda_pattern *ff_space_pattern(ffcuda_space *s) { return s->last_pattern; }

extern "C" void ffcuda_pattern_destroy(ffcuda_pattern *p)
{
    if (!p) return;
    if (p->space && p->space->last_pattern == p) p->space->last_pattern = nullptr;
    cudaSetDevice(p->ctx->device);
    delete p;
}
