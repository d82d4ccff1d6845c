// matrix.cu — device-resident CSR matrices and Dirichlet (penalty) conditions.
// AssembleBC (fflib/problem.cpp:9881-10194) + HashMatrix::SetBC (femlib/HashMatrix.cpp:1195-1238), tgv >= 0.
#include "common.cuh"
#include <memory>
#include <vector>
#include <cmath>
#include <algorithm>

extern "C" int ffcuda_matrix_create(ffcuda_pattern *p, ffcuda_matrix **out)
{
    ffcuda_matrix *A = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(p && out, "ffcuda_matrix_create: null argument");
    ffcuda_ctx *ctx = p->ctx;
    ff_enter(ctx);
    A = new ffcuda_matrix();
    A->ctx = ctx;
    A->ref.set(ctx);
    A->pattern = p;
    A->n = p->n;
    A->ncols = p->ncols_node * p->ncomp;
    A->nnz = p->nnz;
    A->rowptr = p->rowptr;
    A->colind = p->colind;
    A->diagpos = p->diagpos.p;
    A->maxrow = p->maxrow_node * p->ncomp;
    A->vals.alloc((size_t)p->nnz);
    A->vals_stale = true; // zeroed on first read; an overwriting assembly never pays for the memset
    *out = A;
    A = nullptr;
    FF_API_END((delete A, p ? p->ctx : nullptr))
}

void ff_matrix_touch(ffcuda_matrix *A)
{
    if (!A->vals_stale) return;
    FF_CUDA(cudaMemsetAsync(A->vals.p, 0, A->vals.bytes(), A->ctx->stream));
    A->vals_stale = false;
    A->vals_epoch++;
}

__global__ void k_maxrow(const int32_t *__restrict__ rowptr, int n, int32_t *__restrict__ out)
{
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, rowptr[i + 1] - rowptr[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

__global__ void k_find_diag(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, int n, int32_t *__restrict__ diagpos)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = -1;
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
        if (colind[j] == i) d = j;
    diagpos[i] = d;
}

extern "C" int ffcuda_matrix_from_csr(ffcuda_ctx *ctx, int n, int64_t nnz, const int32_t *rowptr, const int32_t *colind,
                                      const double *vals, ffcuda_matrix **out)
{
    ffcuda_matrix *A = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(ctx && out && rowptr && colind && n > 0 && nnz >= 0, "ffcuda_matrix_from_csr: bad arguments");
    FF_REQUIRE(nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros");
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    A = new ffcuda_matrix();
    A->ctx = ctx;
    A->ref.set(ctx);
    A->n = n;
    A->ncols = n;
    A->nnz = nnz;
    A->rowptr_own.alloc((size_t)n + 1);
    A->colind_own.alloc((size_t)nnz);
    A->diagpos_own.alloc((size_t)n);
    A->vals.alloc((size_t)nnz);
    FF_CUDA(cudaMemcpyAsync(A->rowptr_own.p, rowptr, A->rowptr_own.bytes(), cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(A->colind_own.p, colind, A->colind_own.bytes(), cudaMemcpyHostToDevice, st));
    if (vals) FF_CUDA(cudaMemcpyAsync(A->vals.p, vals, A->vals.bytes(), cudaMemcpyHostToDevice, st));
    else FF_CUDA(cudaMemsetAsync(A->vals.p, 0, A->vals.bytes(), st));
    A->rowptr = A->rowptr_own.p;
    A->colind = A->colind_own.p;
    A->diagpos = A->diagpos_own.p;
    ff_launch(ctx, "matrix_find_diag", [&] { k_find_diag<<<ff_blocks(n, 256), 256, 0, st>>>(A->rowptr, ff_matrix_colind(A), n, A->diagpos_own.p); });
    DBuf<int32_t> d_max;
    d_max.alloc(1);
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, sizeof(int32_t), st));
    ff_launch(ctx, "matrix_maxrow", [&] { k_maxrow<<<ctx->sm_count * 4, 256, 0, st>>>(A->rowptr, n, d_max.p); });
    int32_t h_max = 0;
    FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    A->maxrow = h_max;
    *out = A;
    A = nullptr;
    FF_API_END((delete A, ctx))
}

// The rows of ONE rank of a matrix that is shared out by rows (any origin: the FreeFEM plugin splits a host MatriceMorse
// into row blocks, one per GPU): n_owned rows, columns in local numbering - owned dofs first (column i is row i), then the
// ghost dofs grouped by owner rank - and the halo lists in the form ffcuda_partition_local returns them (a neighbour may
// have an empty range in one direction when the structure is not symmetric; both ranks must list each other).  ffcuda_spmv /
// ffcuda_cg / ffcuda_gmres on such a matrix exchange the ghost values and all-reduce the dot products.
extern "C" int ffcuda_matrix_from_csr_distributed(ffcuda_ctx *ctx, int n_owned, int ncols, int64_t nnz, const int32_t *rowptr,
                                                  const int32_t *colind, const double *vals, int nnbr, const int32_t *nbr,
                                                  const int32_t *recv_off, const int32_t *recv_cnt, const int32_t *send_ptr,
                                                  const int32_t *send_idx, ffcuda_matrix **out)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && out && rowptr && colind && n_owned > 0 && ncols >= n_owned, "ffcuda_matrix_from_csr_distributed: bad arguments");
    FF_REQUIRE(ctx->nranks == 1 || ctx->nccl_comm, "ffcuda_matrix_from_csr_distributed: call ffcuda_comm_init first");
    FF_REQUIRE(nnbr >= 0 && nnbr <= ffcuda_mesh::MAXNBR, "at most 16 neighbour ranks");
    FF_REQUIRE(nnbr == 0 || (nbr && recv_off && recv_cnt && send_ptr && (send_idx || send_ptr[nnbr] == 0)), "halo arrays missing");
    int covered = n_owned;
    for (int x = 0; x < nnbr; ++x) {
        FF_REQUIRE(nbr[x] >= 0 && nbr[x] < ctx->nranks && nbr[x] != ctx->rank, "bad neighbour rank");
        FF_REQUIRE(recv_off[x] == covered && recv_cnt[x] >= 0, "ghost ranges must follow the owned dofs, in neighbour order, without gaps");
        covered += recv_cnt[x];
        FF_REQUIRE(send_ptr[x + 1] >= send_ptr[x], "send_ptr must be non-decreasing");
    }
    FF_REQUIRE(covered == ncols, "ghost ranges do not cover the ghost columns");
    for (int k = 0; k < (nnbr ? send_ptr[nnbr] : 0); ++k) FF_REQUIRE(send_idx[k] >= 0 && send_idx[k] < n_owned, "send list entry is not an owned dof");
    for (int64_t k = 0; k < nnz; ++k) FF_REQUIRE(colind[k] >= 0 && colind[k] < ncols, "column index outside the local columns");
    ffcuda_matrix *A = nullptr;
    if (ffcuda_matrix_from_csr(ctx, n_owned, nnz, rowptr, colind, vals, &A) != 0) throw FFError(ffcuda_last_error(ctx));
    std::unique_ptr<ffcuda_matrix> guard(A);
    ff_enter(ctx);
    A->ncols = ncols;
    A->own_halo = true;
    A->nnbr = nnbr;
    for (int x = 0; x < ffcuda_mesh::MAXNBR; ++x) {
        A->nbr[x] = -1;
        A->send_off[x] = A->send_cnt[x] = A->recv_off[x] = A->recv_cnt[x] = 0;
    }
    for (int x = 0; x < nnbr; ++x) {
        A->nbr[x] = nbr[x];
        A->recv_off[x] = recv_off[x];
        A->recv_cnt[x] = recv_cnt[x];
        A->send_off[x] = send_ptr[x];
        A->send_cnt[x] = send_ptr[x + 1] - send_ptr[x];
    }
    if (nnbr && send_ptr[nnbr] > 0) {
        A->send_idx.alloc((size_t)send_ptr[nnbr]);
        FF_CUDA(ff_memcpy_sync(ctx, A->send_idx.p, send_idx, (size_t)send_ptr[nnbr] * 4, cudaMemcpyHostToDevice));
    }
    *out = guard.release();
    FF_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------------
// Half storage (sym=1): MatriceMorse keeps the entries (i, j <= i) only - MatriceElementaireSymetrique /
// HashMatrix::operator+= (femlib/HashMatrix.cpp:1319-1325), mirrored by addMatMul (:1087-1154).  On the device a
// matrix is always stored in full; the lower triangle is cut out when it is handed to the host, and a half-stored host
// matrix is expanded when it comes in.
// ---------------------------------------------------------------------------------------------------
__global__ void k_lower_len(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ diagpos, int n, int32_t *__restrict__ len)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    len[i] = i < n ? diagpos[i] - rowptr[i] + 1 : 0; // sorted rows: everything up to and including the diagonal
}
template <class T>
__global__ void k_lower_copy(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ hrowptr, const T *__restrict__ src,
                             T *__restrict__ dst, int n)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const int b = rowptr[i], hb = hrowptr[i], len = hrowptr[i + 1] - hb;
    for (int k = lane; k < len; k += 32) dst[hb + k] = src[b + k];
}

// rowptr of the lower triangle (device, n+1), built on first use and kept with the pattern
static const int32_t *lower_rowptr(ffcuda_pattern *P)
{
    if (P->lower_rowptr.p) return P->lower_rowptr.p;
    ffcuda_ctx *ctx = P->ctx;
    FF_REQUIRE(P->diagpos.p, "pattern without diagonal index");
    DBuf<int32_t> len;
    len.alloc((size_t)P->n + 1);
    ff_launch(ctx, "lower_len", [&] { k_lower_len<<<ff_blocks((size_t)P->n + 1, 256), 256, 0, ctx->stream>>>(P->rowptr, P->diagpos.p, P->n, len.p); });
    P->lower_rowptr.alloc((size_t)P->n + 1);
    int64_t tot = 0;
    ff_exclusive_scan_i32(ctx, len.p, P->lower_rowptr.p, (size_t)P->n + 1, &tot);
    P->lower_nnz = tot;
    return P->lower_rowptr.p;
}

extern "C" int ffcuda_pattern_lower_nnz(ffcuda_pattern *p, int64_t *nnz_lower)
{
    FF_API_BEGIN
    FF_REQUIRE(p && nnz_lower, "null argument");
    ff_enter(p->ctx);
    lower_rowptr(p);
    *nnz_lower = p->lower_nnz;
    FF_API_END(p ? p->ctx : nullptr)
}

extern "C" int ffcuda_pattern_download_lower(ffcuda_pattern *p, int32_t *rowptr, int32_t *colind)
{
    FF_API_BEGIN
    FF_REQUIRE(p, "null pattern");
    ffcuda_ctx *ctx = p->ctx;
    ff_enter(ctx);
    const int32_t *hrp = lower_rowptr(p);
    if (rowptr) FF_CUDA(cudaMemcpyAsync(rowptr, hrp, ((size_t)p->n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (colind) {
        ff_pattern_ensure_colind(p);
        DBuf<int32_t> tmp;
        tmp.alloc((size_t)p->lower_nnz);
        ff_launch(ctx, "lower_copy", [&] {
            k_lower_copy<int32_t><<<ff_blocks((size_t)p->n * 32, 256), 256, 0, ctx->stream>>>(p->rowptr, hrp, p->colind, tmp.p, p->n);
        });
        FF_CUDA(cudaMemcpyAsync(colind, tmp.p, (size_t)p->lower_nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        FF_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    FF_API_END(p ? p->ctx : nullptr)
}

extern "C" int ffcuda_matrix_download_lower(ffcuda_matrix *A, double *vals)
{
    FF_API_BEGIN
    FF_REQUIRE(A && vals && A->pattern, "ffcuda_matrix_download_lower: needs a matrix created on a pattern");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    ff_matrix_touch(A);
    ffcuda_pattern *p = A->pattern;
    const int32_t *hrp = lower_rowptr(p);
    DBuf<double> tmp;
    tmp.alloc((size_t)p->lower_nnz);
    ff_launch(ctx, "lower_copy", [&] {
        k_lower_copy<double><<<ff_blocks((size_t)p->n * 32, 256), 256, 0, ctx->stream>>>(p->rowptr, hrp, A->vals.p, tmp.p, p->n);
    });
    FF_CUDA(cudaMemcpyAsync(vals, tmp.p, (size_t)p->lower_nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}

// a half-stored host matrix (sorted lower triangle) -> full device matrix; the mirror entries are laid out on the host
// (one pass over the entries: the upper part of row j receives its columns i > j in increasing order)
extern "C" int ffcuda_matrix_from_csr_lower(ffcuda_ctx *ctx, int n, int64_t nnz_lower, const int32_t *rowptr, const int32_t *colind,
                                            const double *vals, ffcuda_matrix **out)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && out && rowptr && colind && vals && n > 0 && nnz_lower >= 0, "ffcuda_matrix_from_csr_lower: bad arguments");
    std::vector<int64_t> cnt((size_t)n + 1, 0);
    int64_t nstrict = 0;
    for (int i = 0; i < n; ++i) {
        cnt[i + 1] += rowptr[i + 1] - rowptr[i];
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int j = colind[k];
            FF_REQUIRE(j >= 0 && j <= i, "ffcuda_matrix_from_csr_lower: an entry lies above the diagonal");
            if (j < i) {
                cnt[j + 1]++;
                nstrict++;
            }
        }
    }
    const int64_t nnz = nnz_lower + nstrict;
    FF_REQUIRE(nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros");
    std::vector<int32_t> frp((size_t)n + 1), fci((size_t)nnz);
    std::vector<double> fv((size_t)nnz);
    frp[0] = 0;
    for (int i = 0; i < n; ++i) frp[i + 1] = (int32_t)(frp[i] + cnt[i + 1]);
    std::vector<int32_t> cur((size_t)n);
    for (int i = 0; i < n; ++i) cur[i] = frp[i] + (rowptr[i + 1] - rowptr[i]); // the upper part starts behind the lower one
    for (int i = 0; i < n; ++i) {
        int o = frp[i];
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k, ++o) {
            const int j = colind[k];
            fci[o] = j;
            fv[o] = vals[k];
            if (j < i) {
                fci[cur[j]] = i;
                fv[cur[j]] = vals[k];
                cur[j]++;
            }
        }
    }
    if (ffcuda_matrix_from_csr(ctx, n, nnz, frp.data(), fci.data(), fv.data(), out) != 0) return 1;
    FF_API_END(ctx)
}

extern "C" int ffcuda_matrix_info(ffcuda_matrix *A, int *n, int64_t *nnz)
{
    FF_API_BEGIN
    FF_REQUIRE(A, "null matrix");
    if (n) *n = A->n;
    if (nnz) *nnz = A->nnz;
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_matrix_shape(ffcuda_matrix *A, int *n, int *ncols, int64_t *nnz)
{
    FF_API_BEGIN
    FF_REQUIRE(A, "null matrix");
    if (n) *n = A->n;
    if (ncols) *ncols = A->ncols;
    if (nnz) *nnz = A->nnz;
    FF_API_END(A ? A->ctx : nullptr)
}

// CSR arrays of a matrix that has no pattern object (ffcuda_matrix_from_csr*, ffcuda_assemble_bilinear_rect) -> host
extern "C" int ffcuda_matrix_download_csr(ffcuda_matrix *A, int32_t *rowptr, int32_t *colind, double *vals)
{
    FF_API_BEGIN
    FF_REQUIRE(A && rowptr && colind, "null argument");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    ff_matrix_touch(A);
    const int32_t *ci = ff_matrix_colind(A);
    FF_CUDA(cudaMemcpyAsync(rowptr, A->rowptr, ((size_t)A->n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (A->nnz > 0) FF_CUDA(cudaMemcpyAsync(colind, ci, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (vals && A->nnz > 0) FF_CUDA(cudaMemcpyAsync(vals, A->vals.p, (size_t)A->nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_matrix_download(ffcuda_matrix *A, double *vals)
{
    FF_API_BEGIN
    FF_REQUIRE(A && vals, "null argument");
    ff_enter(A->ctx);
    ff_matrix_touch(A);
    FF_CUDA(cudaMemcpyAsync(vals, A->vals.p, A->vals.bytes(), cudaMemcpyDeviceToHost, A->ctx->stream));
    FF_CUDA(cudaStreamSynchronize(A->ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_matrix_upload(ffcuda_matrix *A, const double *vals)
{
    FF_API_BEGIN
    FF_REQUIRE(A && vals, "null argument");
    ff_enter(A->ctx);
    A->vals_stale = false;
    A->vals_epoch++;
    FF_CUDA(cudaMemcpyAsync(A->vals.p, vals, A->vals.bytes(), cudaMemcpyHostToDevice, A->ctx->stream));
    FF_CUDA(cudaStreamSynchronize(A->ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}

// ---------------------------------------------------------------------------------------------------------------
// Hand-off formats straight from the device CSR: borrowed device pointers for consumers that
// stay on the GPU (PETSc's MatCreateSeqAIJCUSPARSE-style constructors take exactly these three arrays), the COO triple
// of `[I,J,C] = A` (fflib/lgmat.cpp) with the row indices expanded on the device, and FreeFEM's Morse text format
// (`ofstream << A` after A.CSR, femlib/HashMatrix.hpp:485-508; read back by HashMatrix(istream&), HashMatrix.cpp:137-188).
// ---------------------------------------------------------------------------------------------------------------
namespace {
__global__ void k_expand_rows(const int32_t *__restrict__ rowptr, int n, int32_t *__restrict__ rows, int base)
{
    // one warp per row: rows[rowptr[i] .. rowptr[i+1]) = i
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const int b = rowptr[w], e = rowptr[w + 1];
    for (int k = b + lane; k < e; k += 32) rows[k] = w + base;
}
__global__ void k_shift_i32(const int32_t *__restrict__ in, int64_t n, int base, int32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + base;
}
} // namespace

extern "C" int ffcuda_matrix_export_device(ffcuda_matrix *A, const int32_t **d_rowptr, const int32_t **d_colind, const double **d_vals,
                                           int *n, int64_t *nnz)
{
    FF_API_BEGIN
    FF_REQUIRE(A && d_rowptr && d_colind && d_vals, "null argument");
    ff_enter(A->ctx);
    ff_matrix_touch(A);
    *d_rowptr = A->rowptr;
    *d_colind = ff_matrix_colind(A);
    *d_vals = A->vals.p;
    if (n) *n = A->n;
    if (nnz) *nnz = A->nnz;
    FF_CUDA(cudaStreamSynchronize(A->ctx->stream)); // the arrays are final when the call returns (any stream may read them)
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_matrix_download_coo(ffcuda_matrix *A, int32_t *I, int32_t *J, double *C, int index_base)
{
    FF_API_BEGIN
    FF_REQUIRE(A && I && J && C, "null argument");
    FF_REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    ff_matrix_touch(A);
    const int32_t *colind = ff_matrix_colind(A);
    DBuf<int32_t> rows, cols;
    rows.alloc((size_t)std::max<int64_t>(A->nnz, 1));
    ff_launch(ctx, "export_rows", [&] {
        k_expand_rows<<<ff_blocks((size_t)A->n * 32, 256), 256, 0, ctx->stream>>>(A->rowptr, A->n, rows.p, index_base);
    });
    const int32_t *jsrc = colind;
    if (index_base) {
        cols.alloc((size_t)std::max<int64_t>(A->nnz, 1));
        ff_launch(ctx, "export_cols", [&] { k_shift_i32<<<ff_blocks((size_t)A->nnz, 256), 256, 0, ctx->stream>>>(colind, A->nnz, 1, cols.p); });
        jsrc = cols.p;
    }
    FF_CUDA(cudaMemcpyAsync(I, rows.p, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaMemcpyAsync(J, jsrc, (size_t)A->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaMemcpyAsync(C, A->vals.p, (size_t)A->nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_matrix_write_morse(ffcuda_matrix *A, const char *path, int half)
{
    FF_API_BEGIN
    FF_REQUIRE(A && path, "null argument");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    ff_matrix_touch(A);
    const int32_t *colind = ff_matrix_colind(A);
    FILE *f = fopen(path, "w");
    FF_REQUIRE(f != nullptr, std::string("cannot open ") + path);
    // chunks of rows through pinned staging buffers: the next chunk is copied while this one is formatted
    std::vector<int32_t> rp((size_t)A->n + 1);
    FF_CUDA(ff_memcpy_sync(ctx, rp.data(), A->rowptr, rp.size() * 4, cudaMemcpyDeviceToHost));
    int64_t nnz_out = A->nnz;
    if (half) { // entries (i, j <= i) only, as FreeFEM stores a symmetric matrix
        nnz_out = 0;
    }
    const int64_t CH = (int64_t)1 << 22;
    int32_t *hj[2] = {nullptr, nullptr};
    double *hv[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::string err;
    try {
        for (int b = 0; b < 2; ++b) {
            FF_CUDA(cudaMallocHost((void **)&hj[b], (size_t)CH * 4));
            FF_CUDA(cudaMallocHost((void **)&hv[b], (size_t)CH * 8));
            FF_CUDA(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
        }
        if (half) { // count first (columns are sorted: the prefix of every row up to the diagonal)
            for (int64_t k0 = 0; k0 < A->nnz; k0 += CH) {
                const int64_t m = std::min(CH, A->nnz - k0);
                FF_CUDA(ff_memcpy_sync(ctx, hj[0], colind + k0, (size_t)m * 4, cudaMemcpyDeviceToHost));
                int row = (int)(std::upper_bound(rp.begin(), rp.end(), (int32_t)k0) - rp.begin()) - 1;
                for (int64_t k = 0; k < m; ++k) {
                    while (k0 + k >= rp[(size_t)row + 1]) ++row;
                    if (hj[0][k] <= row) ++nnz_out;
                }
            }
        }
        fprintf(f, "# Sparse Matrix (Morse)  %p\n# first line: n m (is symmetic) nnz \n"
                   "# after for each nonzero coefficient:   i j a_ij where (i,j) \\in  {1,...,n}x{1,...,m} \n",
                (void *)A);
        fprintf(f, "%d %d %d  %lld\n", A->n, A->ncols > 0 && !A->pattern ? A->ncols : A->n, half ? 1 : 0, (long long)nnz_out);
        auto fetch = [&](int b, int64_t k0) {
            const int64_t m = std::min(CH, A->nnz - k0);
            FF_CUDA(cudaMemcpyAsync(hj[b], colind + k0, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
            FF_CUDA(cudaMemcpyAsync(hv[b], A->vals.p + k0, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
            FF_CUDA(cudaEventRecord(ev[b], ctx->stream));
        };
        if (A->nnz > 0) fetch(0, 0);
        int b = 0, row = 0;
        std::vector<char> line((size_t)1 << 20);
        for (int64_t k0 = 0; k0 < A->nnz; k0 += CH, b ^= 1) {
            if (k0 + CH < A->nnz) fetch(b ^ 1, k0 + CH);
            FF_CUDA(cudaEventSynchronize(ev[b]));
            const int64_t m = std::min(CH, A->nnz - k0);
            size_t pos = 0;
            for (int64_t k = 0; k < m; ++k) {
                while (k0 + k >= rp[(size_t)row + 1]) ++row;
                if (half && hj[b][k] > row) continue;
                const double v = std::fabs(hv[b][k]) < 1e-305 ? 0.0 : hv[b][k]; // RNM::removeeps
                pos += (size_t)snprintf(line.data() + pos, 80, "%9d %9d %.20g\n", row + 1, hj[b][k] + 1, v);
                if (pos + 128 > line.size()) {
                    fwrite(line.data(), 1, pos, f);
                    pos = 0;
                }
            }
            fwrite(line.data(), 1, pos, f);
        }
    } catch (const std::exception &e) {
        err = e.what();
    }
    for (int b = 0; b < 2; ++b) {
        if (hj[b]) cudaFreeHost(hj[b]);
        if (hv[b]) cudaFreeHost(hv[b]);
        if (ev[b]) cudaEventDestroy(ev[b]);
    }
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (!err.empty()) throw FFError(err);
    FF_REQUIRE(!bad, std::string("error while writing ") + path);
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" void ffcuda_matrix_destroy(ffcuda_matrix *A)
{
    if (!A) return;
    ff_enter(A->ctx);
    delete A;
}

// ---------------------------------------------------------------------------------------------------
// Dirichlet conditions
// ---------------------------------------------------------------------------------------------------
extern "C" int ffcuda_bc_from_pairs(ffcuda_space *s, int n, const int32_t *dofs, const double *vals, ffcuda_bc **out)
{
    ffcuda_bc *bc = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(s && out && n >= 0 && (n == 0 || (dofs && vals)), "ffcuda_bc_from_pairs: bad arguments");
    ffcuda_ctx *ctx = s->ctx;
    ff_enter(ctx);
    const int ndof = s->nnodes_owned * s->ncomp;
    // later pairs win (AssembleBC overwrites B[ddf] as it walks the boundary elements)
    std::vector<std::pair<int32_t, int>> ord(n);
    for (int i = 0; i < n; ++i) {
        FF_REQUIRE(dofs[i] >= 0 && dofs[i] < ndof, "Dirichlet dof out of range");
        ord[i] = {dofs[i], i};
    }
    std::sort(ord.begin(), ord.end());
    std::vector<int32_t> ud;
    std::vector<double> uv;
    for (int i = 0; i < n; ++i)
        if (i + 1 == n || ord[i + 1].first != ord[i].first) {
            ud.push_back(ord[i].first);
            uv.push_back(vals[ord[i].second]);
        }
    bc = new ffcuda_bc();
    bc->ctx = ctx;
    bc->ref.set(ctx);
    bc->ndofs = (int)ud.size();
    bc->dofs.alloc(ud.size());
    bc->vals.alloc(uv.size());
    if (bc->ndofs) {
        FF_CUDA(ff_memcpy_sync(ctx, bc->dofs.p, ud.data(), bc->dofs.bytes(), cudaMemcpyHostToDevice));
        FF_CUDA(ff_memcpy_sync(ctx, bc->vals.p, uv.data(), bc->vals.bytes(), cudaMemcpyHostToDevice));
    }
    *out = bc;
    bc = nullptr;
    FF_API_END((delete bc, s ? s->ctx : nullptr))
}

struct LabSet {
    int n;
    int lab[32];
};

// mark the dofs lying on boundary elements whose label is selected (Element::onWhatBorder semantics:
// a vertex is on face ie iff it is not the vertex opposite to it; an edge iff both its ends are)
__global__ void k_bc_mark(int dim, int nbe, const int32_t *__restrict__ blab, const int32_t *__restrict__ belem,
                          const int32_t *__restrict__ bface, const int32_t *__restrict__ e2n, int nloc, int ncomp, int compmask,
                          int ndof, LabSet L, int32_t *__restrict__ flag)
{
    int ib = blockIdx.x * blockDim.x + threadIdx.x;
    if (ib >= nbe) return;
    int l = blab[ib];
    bool ok = false;
    for (int i = 0; i < L.n; ++i) ok |= (L.lab[i] == l);
    if (!ok) return;
    const int it = belem[ib], ie = bface[ib], nvk = dim + 1;
    const int e0[6] = {0, 0, 0, 1, 1, 2}, e1[6] = {1, 2, 3, 2, 3, 3};
    for (int a = 0; a < nloc; ++a) {
        bool on;
        if (a < nvk) on = (a != ie);
        else if (dim == 2) on = (a - 3 == ie);
        else on = (e0[a - 4] != ie && e1[a - 4] != ie);
        if (!on) continue;
        int node = e2n[(size_t)it * nloc + a];
        for (int c = 0; c < ncomp; ++c)
            if (compmask >> c & 1) {
                int d = node * ncomp + c;
                if (d < ndof) flag[d] = 1;
            }
    }
}

__global__ void k_bc_compact(const int32_t *__restrict__ flag, const int32_t *__restrict__ off, int ndof, int ncomp,
                             double v0, double v1, double v2, int32_t *__restrict__ dofs, double *__restrict__ vals)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof || !flag[d]) return;
    int c = d % ncomp;
    dofs[off[d]] = d;
    vals[off[d]] = c == 0 ? v0 : (c == 1 ? v1 : v2);
}

extern "C" int ffcuda_bc_from_labels(ffcuda_space *s, int nlab, const int32_t *labels, int compmask, const double *values,
                                     ffcuda_bc **out)
{
    ffcuda_bc *bc = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(s && out && nlab > 0 && labels, "ffcuda_bc_from_labels: bad arguments");
    FF_REQUIRE(nlab <= 32, "at most 32 labels per on(...)");
    ffcuda_ctx *ctx = s->ctx;
    ffcuda_mesh *m = s->mesh;
    ff_enter(ctx);
    cudaStream_t st = ctx->stream;
    const int ndof = s->nnodes_owned * s->ncomp;
    LabSet L;
    L.n = nlab;
    for (int i = 0; i < nlab; ++i) L.lab[i] = labels[i];
    DBuf<int32_t> flag, off;
    flag.alloc((size_t)ndof + 1);
    off.alloc((size_t)ndof + 1);
    FF_CUDA(cudaMemsetAsync(flag.p, 0, flag.bytes(), st));
    if (m->nbe)
        ff_launch(ctx, "bc_mark", [&] {
            k_bc_mark<<<ff_blocks(m->nbe, 128), 128, 0, st>>>(m->dim, m->nbe, m->blab.p, m->belem.p, m->bface.p, s->e2n, s->nloc,
                                                              s->ncomp, compmask, ndof, L, flag.p);
        });
    int64_t cnt = 0;
    ff_exclusive_scan_i32(ctx, flag.p, off.p, (size_t)ndof + 1, &cnt);
    bc = new ffcuda_bc();
    bc->ctx = ctx;
    bc->ref.set(ctx);
    bc->ndofs = (int)cnt;
    bc->dofs.alloc((size_t)cnt);
    bc->vals.alloc((size_t)cnt);
    double v[3] = {0, 0, 0};
    if (values)
        for (int c = 0; c < s->ncomp; ++c) v[c] = values[c];
    if (cnt)
        ff_launch(ctx, "bc_compact", [&] {
            k_bc_compact<<<ff_blocks(ndof, 256), 256, 0, st>>>(flag.p, off.p, ndof, s->ncomp, v[0], v[1], v[2], bc->dofs.p, bc->vals.p);
        });
    FF_CUDA(cudaStreamSynchronize(st));
    *out = bc;
    bc = nullptr;
    FF_API_END((delete bc, s ? s->ctx : nullptr))
}

extern "C" int ffcuda_bc_count(ffcuda_bc *bc, int *ndofs)
{
    FF_API_BEGIN
    FF_REQUIRE(bc && ndofs, "null argument");
    *ndofs = bc->ndofs;
    FF_API_END(bc ? bc->ctx : nullptr)
}

__global__ void k_bc_matrix(const int32_t *__restrict__ dofs, int n, const int32_t *__restrict__ diagpos, double *__restrict__ vals, double tgv)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals[diagpos[dofs[i]]] = tgv;
}

// Exact elimination, HashMatrix::SetBC with tgv < 0 (femlib/HashMatrix.cpp:1195-1238).
// rows: one warp per Dirichlet dof: diagonal = dval (1, or 0 when tgv < -9), the rest of the row 0 (kept for -3 / -30)
__global__ void k_bc_rows_exact(const int32_t *__restrict__ dofs, int n, const int32_t *__restrict__ rowptr,
                                const int32_t *__restrict__ colind, double *__restrict__ vals, double dval, int keeprow)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const int d = dofs[i];
    for (int k = rowptr[d] + lane; k < rowptr[d + 1]; k += 32) {
        if (colind[k] == d) vals[k] = dval;
        else if (!keeprow) vals[k] = 0.0;
    }
}
__global__ void k_bc_mask(const int32_t *__restrict__ dofs, int n, unsigned char *__restrict__ on)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) on[dofs[i]] = 1;
}
// columns (tgv = -2, -20, -3, -30): one warp per row of the matrix; the diagonal goes too when tgv < -19
__global__ void k_bc_cols_exact(int nrows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                const unsigned char *__restrict__ on, double *__restrict__ vals, int withdiag)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= nrows) return;
    for (int k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
        const int c = colind[k];
        if (on[c] && (c != r || withdiag)) vals[k] = 0.0;
    }
}

__global__ void k_bc_vec(const int32_t *__restrict__ dofs, const double *__restrict__ g, int n, double *__restrict__ b, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[dofs[i]] = scale * g[i];
}

extern "C" int ffcuda_matrix_apply_bc(ffcuda_matrix *A, ffcuda_bc *bc, double tgv)
{
    FF_API_BEGIN
    FF_REQUIRE(A && bc, "null argument");
    FF_REQUIRE(tgv == tgv, "tgv is NaN");
    FF_REQUIRE(!A->rect, "ffcuda_matrix_apply_bc is for square matrices (this one is rectangular: products and hand-off only)");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    ff_matrix_touch(A);
    A->vals_epoch++;
    if (tgv >= 0) {
        if (bc->ndofs)
            ff_launch(ctx, "bc_matrix", [&] {
                k_bc_matrix<<<ff_blocks(bc->ndofs, 256), 256, 0, ctx->stream>>>(bc->dofs.p, bc->ndofs, A->diagpos, A->vals.p, tgv);
            });
    } else if (bc->ndofs) {
        FF_REQUIRE(A->n == A->ncols, "exact elimination (tgv < 0) needs a square, non-distributed matrix");
        auto near = [&](double v) { return fabs(tgv - v) < 1.0e-10; };
        const int keeprow = near(-3.0) || near(-30.0);
        const int cols = near(-2.0) || near(-20.0) || near(-3.0) || near(-30.0);
        ff_launch(ctx, "bc_matrix", [&] {
            k_bc_rows_exact<<<ff_blocks((size_t)bc->ndofs * 32, 256), 256, 0, ctx->stream>>>(bc->dofs.p, bc->ndofs, A->rowptr, ff_matrix_colind(A),
                                                                                              A->vals.p, tgv < -9.0 ? 0.0 : 1.0, keeprow);
        });
        if (cols) {
            DBuf<unsigned char> on;
            on.alloc((size_t)A->ncols);
            FF_CUDA(cudaMemsetAsync(on.p, 0, on.bytes(), ctx->stream));
            ff_launch(ctx, "bc_mark", [&] { k_bc_mask<<<ff_blocks(bc->ndofs, 256), 256, 0, ctx->stream>>>(bc->dofs.p, bc->ndofs, on.p); });
            ff_launch(ctx, "bc_matrix", [&] {
                k_bc_cols_exact<<<ff_blocks((size_t)A->n * 32, 256), 256, 0, ctx->stream>>>(A->n, A->rowptr, ff_matrix_colind(A), on.p, A->vals.p,
                                                                                            tgv < -19.0);
            });
        }
    }
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_vec_apply_bc(ffcuda_vec *b, ffcuda_bc *bc, double tgv)
{
    FF_API_BEGIN
    FF_REQUIRE(b && bc, "null argument");
    ffcuda_ctx *ctx = b->ctx;
    ff_enter(ctx);
    if (bc->ndofs)
        ff_launch(ctx, "bc_vec", [&] {
            // AssembleBC: B[dof] = tgv1 * g, tgv1 = tgv for the penalty form, 1 for exact elimination (problem.cpp:10099)
            k_bc_vec<<<ff_blocks(bc->ndofs, 256), 256, 0, ctx->stream>>>(bc->dofs.p, bc->vals.p, bc->ndofs, b->d.p, tgv < 0 ? 1.0 : tgv);
        });
    FF_API_END(b ? b->ctx : nullptr)
}

extern "C" int ffcuda_vec_set_bc_values(ffcuda_vec *x, ffcuda_bc *bc)
{
    FF_API_BEGIN
    FF_REQUIRE(x && bc, "null argument");
    ffcuda_ctx *ctx = x->ctx;
    ff_enter(ctx);
    if (bc->ndofs)
        ff_launch(ctx, "bc_vec", [&] {
            k_bc_vec<<<ff_blocks(bc->ndofs, 256), 256, 0, ctx->stream>>>(bc->dofs.p, bc->vals.p, bc->ndofs, x->d.p, 1.0);
        });
    FF_API_END(x ? x->ctx : nullptr)
}

extern "C" void ffcuda_bc_destroy(ffcuda_bc *bc)
{
    if (!bc) return;
    ff_enter(bc->ctx);
    delete bc;
}
