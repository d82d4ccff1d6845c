// symbolic.cu — KERNEL 1: symbolic sparsity from the element -> dof map.
//
// Output = the CSR pattern MatriceMorse holds after CSR(): for every element, every couple of its dofs is an
// entry (HashMatrix::operator+=(MatriceElementaire&), femlib/HashMatrix.cpp:1310-1317), rows/columns sorted
// (Sortij :671, Buildp :993).  Built at NODE level (a vector space [P,P,P] has dof = node*ncomp + c, so its
// pattern is the node pattern with every entry replaced by a dense ncomp x ncomp block) and then expanded.
//
// Two stages:
//  (1) ff_build_incidence — the transpose of the element -> node table (node -> sorted (element, local node) lists),
//      a property of the FE space, built once per space: count (integer atomics: order-independent result) -> scan
//      -> fill -> per-node sort (restores a deterministic order).
//  (2) ffcuda_symbolic — one warp per node row: the nodes of the incident elements are de-duplicated in a per-warp
//      shared-memory hash table (atomicCAS), which gives the row length (pass 0); after the scan the same is done
//      again, the <= 32 distinct columns are sorted in registers by a shuffle bitonic network (longer rows: bitonic
//      sort in shared memory), written out, and every (incidence, element node) gets its position in the row by
//      binary search (pass 1).  These positions are what lets the numeric phase run without any search or atomic.
#include "common.cuh"
#include <climits>

// ---------------------------------------------------------------------------------------------------------------
// stage 1: node -> element incidence
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_count_inc(const int32_t *__restrict__ e2n, size_t nitems, int nrows, int32_t *__restrict__ cnt)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    int node = e2n[i];
    if (node < nrows) atomicAdd(&cnt[node], 1);
}

// one warp per block of 32 rows: padded length of the block (32 * longest list) and the global maximum
__global__ void k_blk_len(const int32_t *__restrict__ cnt, int nrows, int nblk, int32_t *__restrict__ blklen, int32_t *__restrict__ maxinc)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nblk) return;
    const int row = b * 32 + lane;
    int m = row < nrows ? cnt[row] : 0;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        blklen[b] = 32 * m;
        if (m > 0) atomicMax(maxinc, m);
    }
}

__global__ void k_max_i32(const int32_t *__restrict__ v, int n, int32_t *__restrict__ out)
{
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

__global__ void k_fill_inc(const int32_t *__restrict__ e2n, size_t nitems, int nloc, int nrows, const IncView V,
                           int32_t *__restrict__ cursor, uint32_t *__restrict__ inc)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nitems) return;
    int node = e2n[i];
    if (node >= nrows) return;
    uint32_t k = (uint32_t)(i / nloc), a = (uint32_t)(i - (size_t)k * nloc);
    int slot = atomicAdd(&cursor[node], 1);
    inc[V.idx(node, slot)] = (k << 4) | a;
}

// one thread per node: insertion sort of its incidence list (ascending element index).  In the ELL layout the 32
// lanes of a warp touch 32 consecutive records at every step.
__global__ void k_sort_inc(const IncView V, uint32_t *__restrict__ inc, int nrows)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int len = V.cnt[i];
    if (len < 2) return;
    uint32_t *p = inc + V.idx(i, 0);
    const int st = V.ell ? 32 : 1;
    for (int x = 1; x < len; ++x) {
        uint32_t v = p[(size_t)x * st];
        int y = x - 1;
        while (y >= 0 && p[(size_t)y * st] > v) {
            p[(size_t)(y + 1) * st] = p[(size_t)y * st];
            --y;
        }
        p[(size_t)(y + 1) * st] = v;
    }
}

void ff_build_incidence(ffcuda_space *s)
{
    Incidence &I = s->incidence;
    if (I.built) return;
    ffcuda_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int nt = s->mesh->nt, nloc = s->nloc;
    const int nrows = s->nnodes_owned;
    const size_t nitems = (size_t)nt * nloc;
    FF_REQUIRE((int64_t)nt < ((int64_t)1 << 28), "too many elements for one device (limit 2^28)");
    I.nrows = nrows;
    I.ell = (s->order == 1);
    I.cnt.alloc((size_t)nrows + 1);
    FF_CUDA(cudaMemsetAsync(I.cnt.p, 0, I.cnt.bytes(), st));
    ff_launch(ctx, "inc_count", [&] { k_count_inc<<<ff_blocks(nitems, 256), 256, 0, st>>>(s->e2n, nitems, nrows, I.cnt.p); });
    DBuf<int32_t> d_max;
    d_max.alloc(1);
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, sizeof(int32_t), st));
    int64_t nrec = 0;
    if (I.ell) {
        const int nblk = (nrows + 31) / 32;
        DBuf<int32_t> blklen;
        blklen.alloc((size_t)nblk + 1);
        FF_CUDA(cudaMemsetAsync(blklen.p, 0, blklen.bytes(), st));
        ff_launch(ctx, "inc_blk_len", [&] { k_blk_len<<<ff_blocks((size_t)nblk * 32, 256), 256, 0, st>>>(I.cnt.p, nrows, nblk, blklen.p, d_max.p); });
        I.blkoff.alloc((size_t)nblk + 1);
        ff_exclusive_scan_i32(ctx, blklen.p, reinterpret_cast<int32_t *>(I.blkoff.p), (size_t)nblk + 1, &nrec);
    } else {
        ff_launch(ctx, "inc_max", [&] { k_max_i32<<<ctx->sm_count * 4, 256, 0, st>>>(I.cnt.p, nrows, d_max.p); });
        I.incptr.alloc((size_t)nrows + 1);
        ff_exclusive_scan_i32(ctx, I.cnt.p, I.incptr.p, (size_t)nrows + 1, &nrec);
    }
    FF_REQUIRE(nrec < ((int64_t)1 << 31), "incidence table exceeds int32");
    int32_t h_max = 0;
    FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    I.maxinc = h_max;
    I.nrec = nrec;
    FF_REQUIRE(I.maxinc > 0, "no element touches any owned node");
    I.inc.alloc((size_t)nrec);
    if (I.ell) FF_CUDA(cudaMemsetAsync(I.inc.p, 0xff, I.inc.bytes(), st)); // padding records = FF_NOREC
    DBuf<int32_t> cursor;
    cursor.alloc((size_t)nrows + 1);
    FF_CUDA(cudaMemsetAsync(cursor.p, 0, cursor.bytes(), st));
    const IncView V = ff_view(I);
    ff_launch(ctx, "inc_fill", [&] { k_fill_inc<<<ff_blocks(nitems, 256), 256, 0, st>>>(s->e2n, nitems, nloc, nrows, V, cursor.p, I.inc.p); });
    ff_launch(ctx, "inc_sort", [&] { k_sort_inc<<<ff_blocks(nrows, 128), 128, 0, st>>>(V, I.inc.p, nrows); });
    I.built = true;
}

// ---------------------------------------------------------------------------------------------------------------
// stage 2: row pattern.  One warp per node row.
// ---------------------------------------------------------------------------------------------------------------
static constexpr uint32_t HT_EMPTY = 0xffffffffu;

__device__ __forceinline__ int warp_sort32(int v, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            v = (lower == up) ? min(v, o) : max(v, o);
        }
    return v;
}

// PASS 0: row length.  PASS 1: columns, positions, diagonal position.
template <int PASS, typename PosT>
__global__ void __launch_bounds__(256) k_row_pattern(const int32_t *__restrict__ e2n, int nloc, int nlocp, int nrows, int TS, int LOG,
                                                     const IncView V, int32_t *__restrict__ rowlen,
                                                     const int32_t *__restrict__ nrowptr, int32_t *__restrict__ ncol,
                                                     PosT *__restrict__ pos, int32_t *__restrict__ diagnode,
                                                     int32_t *__restrict__ maxrow)
{
    extern __shared__ uint32_t smem_u[];
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *tab = smem_u + (size_t)w * (TS + TS / 2);       // hash table, later the sorted distinct columns
    int32_t *cand = reinterpret_cast<int32_t *>(tab + TS);    // the candidates in incidence order (PASS 1)
    const uint32_t lt = (1u << lane) - 1u, hmask = (uint32_t)TS - 1u;
    int localmax = 0;
    for (int row = blockIdx.x * warps + w; row < nrows; row += gridDim.x * warps) {
        const int ninc = V.cnt[row];
        const int ncand = ninc * nloc;
        for (int x = lane; x < TS; x += 32) tab[x] = HT_EMPTY;
        __syncwarp();
        int nu = 0;
        for (int x0 = 0; x0 < ncand; x0 += 32) {
            const int x = x0 + lane;
            bool isnew = false;
            if (x < ncand) {
                const int e = x / nloc, b = x - e * nloc;
                const uint32_t ka = V.inc[V.idx(row, e)];
                const uint32_t v = (uint32_t)e2n[(size_t)(ka >> 4) * nloc + b];
                if (PASS == 1) cand[x] = (int32_t)v;
                uint32_t h = (v * 0x9E3779B1u) >> (32 - LOG);
                while (true) {
                    const uint32_t old = atomicCAS(&tab[h], HT_EMPTY, v);
                    if (old == HT_EMPTY) {
                        isnew = true;
                        break;
                    }
                    if (old == v) break;
                    h = (h + 1) & hmask;
                }
            }
            nu += __popc(__ballot_sync(0xffffffffu, isnew));
        }
        __syncwarp();
        if (PASS == 0) {
            if (lane == 0) rowlen[row] = nu;
            localmax = max(localmax, nu);
        } else {
            // compact the distinct values to the front of the table (write index never passes the read index)
            int nc = 0;
            for (int base = 0; base < TS; base += 32) {
                const uint32_t v = tab[base + lane];
                const bool keep = v != HT_EMPTY;
                const unsigned msk = __ballot_sync(0xffffffffu, keep);
                if (keep) tab[nc + __popc(msk & lt)] = v;
                nc += __popc(msk);
                __syncwarp();
            }
            int32_t *u = reinterpret_cast<int32_t *>(tab);
            if (nu <= 32) {
                int v = lane < nu ? u[lane] : INT_MAX;
                v = warp_sort32(v, lane);
                __syncwarp();
                u[lane] = v;
            } else {
                int m = 64;
                while (m < nu) m <<= 1;
                for (int x = nu + lane; x < m; x += 32) u[x] = INT_MAX;
                __syncwarp();
                for (int k = 2; k <= m; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int x = lane; x < m; x += 32) {
                            const int y = x ^ j;
                            if (y > x) {
                                const int a = u[x], b = u[y];
                                const bool up = (x & k) == 0;
                                if ((a > b) == up) {
                                    u[x] = b;
                                    u[y] = a;
                                }
                            }
                        }
                        __syncwarp();
                    }
            }
            __syncwarp();
            const int rb = nrowptr[row];
            for (int x = lane; x < nu; x += 32) {
                const int v = u[x];
                ncol[(size_t)rb + x] = v;
                if (v == row) diagnode[row] = x;
            }
            // positions of the nodes of every incident element inside this row
            for (int x = lane; x < ncand; x += 32) {
                const int v = cand[x];
                int lo = 0, hi = nu - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (u[mid] < v) lo = mid + 1;
                    else hi = mid;
                }
                const int e = x / nloc, b = x - e * nloc;
                pos[V.idx(row, e) * nlocp + b] = (PosT)lo;
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
static void run_row_pattern(ffcuda_ctx *ctx, ffcuda_pattern *P, const int32_t *e2n, int nloc, int TS, int LOG, int warps, size_t shmem,
                            int blocks, const IncView &V, int32_t *rowlen, int32_t *diagnode, int32_t *d_max, PosT *pos, int pass)
{
    cudaStream_t st = ctx->stream;
    if (pass == 0) {
        FF_CUDA(cudaFuncSetAttribute(k_row_pattern<0, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        ff_launch(ctx, "sym_row_count", [&] {
            k_row_pattern<0, PosT><<<blocks, warps * 32, shmem, st>>>(e2n, nloc, P->nlocp, P->nrows_node, TS, LOG, V, rowlen, nullptr,
                                                                      nullptr, nullptr, nullptr, d_max);
        });
    } else {
        FF_CUDA(cudaFuncSetAttribute(k_row_pattern<1, PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        ff_launch(ctx, "sym_row_fill", [&] {
            k_row_pattern<1, PosT><<<blocks, warps * 32, shmem, st>>>(e2n, nloc, P->nlocp, P->nrows_node, TS, LOG, V, nullptr,
                                                                      P->nrowptr.p, P->ncol.p, pos, diagnode, nullptr);
        });
    }
}

extern "C" int ffcuda_symbolic(ffcuda_space *s, ffcuda_pattern **out)
{
    ffcuda_pattern *P = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(s && out, "ffcuda_symbolic: null space/output");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    const int nloc = s->nloc, nc = s->ncomp;
    const int nrows = s->nnodes_owned;
    ff_build_incidence(s);
    const Incidence &I = s->incidence;
    const IncView V = ff_view(I);
    P = new ffcuda_pattern();
    P->space = s;
    P->ctx = ctx;
    P->nrows_node = nrows;
    P->ncols_node = s->nnodes;
    P->ncomp = nc;
    P->n = nrows * nc;
    P->nlocp = (s->order == 1) ? 4 : nloc;

    // --- row lengths
    int TS = 64, LOG = 6;
    while (TS < 2 * I.maxinc * nloc) {
        TS <<= 1;
        ++LOG;
    }
    const size_t per_warp = ((size_t)TS + TS / 2) * 4;
    FF_REQUIRE(per_warp <= 200 * 1024, "a node has too many incident elements for the shared-memory hash table");
    int warps = 8;
    while (warps > 1 && (size_t)warps * per_warp > 96 * 1024) warps >>= 1;
    const size_t shmem = (size_t)warps * per_warp;
    const int blocks = min(ff_blocks((size_t)nrows, warps), ctx->sm_count * 16);
    DBuf<int32_t> rowlen, diagnode, d_max;
    rowlen.alloc((size_t)nrows + 1);
    diagnode.alloc((size_t)nrows);
    d_max.alloc(1);
    FF_CUDA(cudaMemsetAsync(rowlen.p + nrows, 0, sizeof(int32_t), st));
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, sizeof(int32_t), st));
    run_row_pattern<uint8_t>(ctx, P, s->e2n, nloc, TS, LOG, warps, shmem, blocks, V, rowlen.p, diagnode.p, d_max.p, nullptr, 0);
    P->nrowptr.alloc((size_t)nrows + 1);
    int64_t nnzn = 0;
    int32_t h_max = 0;
    FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ff_exclusive_scan_i32(ctx, rowlen.p, P->nrowptr.p, (size_t)nrows + 1, &nnzn); // synchronises the stream
    P->maxrow_node = h_max;
    P->nnz_node = nnzn;
    P->nnz = nnzn * nc * nc;
    FF_REQUIRE(P->nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros (int32 CSR, like MatriceMorse)");
    rowlen.release();

    // --- columns + positions
    P->ncol.alloc((size_t)nnzn);
    if (P->maxrow_node <= 255) {
        P->pos8.alloc((size_t)I.nrec * P->nlocp);
        run_row_pattern<uint8_t>(ctx, P, s->e2n, nloc, TS, LOG, warps, shmem, blocks, V, nullptr, diagnode.p, nullptr, P->pos8.p, 1);
    } else {
        FF_REQUIRE(s->order == 2, "a P1 node with more than 254 neighbours is not supported");
        P->pos16.alloc((size_t)I.nrec * P->nlocp);
        run_row_pattern<uint16_t>(ctx, P, s->e2n, nloc, TS, LOG, warps, shmem, blocks, V, nullptr, diagnode.p, nullptr, P->pos16.p, 1);
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
    // no synchronisation here: everything downstream is ordered on the same stream (the temporaries above are
    // released through the stream-ordered allocator)
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
    ff_enter(p->ctx);
    cudaStream_t st = p->ctx->stream;
    if (rowptr) FF_CUDA(cudaMemcpyAsync(rowptr, p->rowptr, ((size_t)p->n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (colind) FF_CUDA(cudaMemcpyAsync(colind, p->colind, (size_t)p->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_API_END(p ? p->ctx : nullptr)
}

extern "C" void ffcuda_pattern_destroy(ffcuda_pattern *p)
{
    if (!p) return;
    ff_enter(p->ctx);
    delete p;
}
