// solver.cu — KERNEL 4: CSR SpMV (a group of T lanes per row) and the Jacobi-preconditioned CG built on it.
//
// Replaces HashMatrix::addMatMul (femlib/HashMatrix.cpp:1087-1154), HMatVirtPrecon / SolverCG::dosolver
// (femlib/VirtualSolverCG.hpp:13-192), gettgv (HashMatrix.cpp:1341-1371) and ConjugueGradient
// (femlib/CG.cpp:195-265).  The recurrence, the initial guess handling (SetInitWithBC), the preconditioner and the
// stopping rule are the reference's; what differs is the summation order of the SpMV rows and of the dot products
// (fixed-shape trees: bit-reproducible from run to run, ~1e-16 relative away from the sequential CPU sums).
//
// One CG iteration = 3 kernels, every scalar stays on the device (no host round trip inside the loop):
//   K1  AH = A*H           fused with the dot products <G,H> and <H,AH>        reads 12 B/nnz + H,G   writes AH
//   K2  G += rho*AH        fused with <G, D1*G>  (rho = -<G,H>/<H,AH>)          reads G,AH,D1          writes G
//   K3  x += rho*H ; H = gamma*H - D1*G  (gamma = gCg/gCg_prev) ; convergence   reads x,H,G,D1         writes x,H
// Dot products: per-block partial sums in a fixed tree, finished by the last block to retire (fixed order over the
// partials) -> deterministic, no fp64 atomics.  The host enqueues iterations in batches and polls a device flag;
// kernels of iterations past the converged one are no-ops, so x is exactly the iterate of the stopping iteration.
#include <deque>
#include <cooperative_groups.h>
#include "common.cuh"
#include <cstddef>
#include <algorithm>
#include <cmath>

// d_scal layout (doubles)
enum { S_GH = 0, S_HAH = 1, S_GCG0 = 2, S_GCG1 = 3, S_EPS2 = 4, S_TMP0 = 8, S_TMP1 = 9, S_TMP2 = 10 };
// d_flag layout (ints)
enum { F_CONV_ITER = 0, F_ITBASE = 1, F_COUNTER = 2 };

static constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fixed-shape block sum (all threads must call); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *sh /* >= 32 doubles */)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (w == 0) {
        r = l < nw ? sh[l] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// Every block deposits NV partial sums; the last block to retire adds the partials of all blocks in a fixed order
// and stores the totals in out[0..NV).  counter must be 0 on entry and is reset to 0 on exit.
template <int NV>
__device__ __forceinline__ void grid_sum_finish(const double (&v)[NV], double *__restrict__ partial, int *counter,
                                                double *__restrict__ out, double *sh, bool xrank = false)
{
    __shared__ bool last;
    double r[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) r[k] = block_sum(v[k], sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partial[(size_t)k * gridDim.x + blockIdx.x] = r[k];
        __threadfence();
        int t = atomicAdd(counter, 1);
        last = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partial + (size_t)k * gridDim.x + i);
        tot[k] = block_sum(s, sh); // valid in thread 0
    }
    if (threadIdx.x < 32) {
        // xrank: the sums of a distributed CG are all-reduced right here, by the last block of the kernel that formed
        // them, through the peer mailboxes (comm.cu) - no separate collective on the stream.  counter = flags + F_COUNTER,
        // the descriptor sits FF_P2P_DESC_OFF - 32 doubles behind the flags.
        P2PDesc *D = reinterpret_cast<P2PDesc *>(reinterpret_cast<double *>(counter - F_COUNTER) + (FF_P2P_DESC_OFF - 32));
        if (xrank && D->fused) ff_p2p_allreduce_warp<NV>(D, tot, false);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) out[k] = tot[k];
            *counter = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// row product: T lanes per row, lane l takes entries l, l+T, ... ; xor tree over the T lanes
// ---------------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ double row_product(const int32_t *__restrict__ colind, const double *__restrict__ vals,
                                              const double *__restrict__ x, int rb, int re, int l)
{
    double s = 0;
    for (int j = rb + l; j < re; j += T) s = fma(__ldcs(vals + j), __ldg(x + __ldcs(colind + j)), s);
#pragma unroll
    for (int o = T >> 1; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// ---------------------------------------------------------------------------------------------------
// CSR-stream SpMV (the default): a CTA owns a contiguous run of rows holding about TILE non-zeros (row-block table
// built once per matrix).  Phase 1: the 256 threads sweep the run's (value, column) pairs in perfectly coalesced
// order, STREAM_ITEMS independent loads per thread in flight, gather x and park the products in shared memory.
// Phase 2: a group of T lanes per row adds the row's products (T = 1: left to right, the CPU's own order).
// MODE 0: y = A x          MODE 1: y = A x - sub, partial <y, d1 y>          MODE 2: y = A x, partials <g,x>, <x,y>
// ---------------------------------------------------------------------------------------------------
static constexpr int STREAM_TILE = 2048;
static constexpr int STREAM_ITEMS = STREAM_TILE / RED_THREADS;

__global__ void k_stream_blocks(const int32_t *__restrict__ rowptr, int n, int nblk, int32_t *__restrict__ rb)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > nblk) return;
    if (j == nblk) {
        rb[j] = n;
        return;
    }
    // smallest row r with rowptr[r] >= j * TILE
    const long long target = (long long)j * STREAM_TILE;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rowptr[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    rb[j] = lo;
}

template <int T, int MODE>
__global__ void __launch_bounds__(RED_THREADS) k_spmv_stream(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                                             const double *__restrict__ vals, const double *__restrict__ x,
                                                             const int32_t *__restrict__ rb, int nblk, const double *__restrict__ aux0,
                                                             const double *__restrict__ aux1, double *__restrict__ y, int iter,
                                                             double *__restrict__ partial, int *__restrict__ flags,
                                                             double *__restrict__ out)
{
    extern __shared__ double sprod[];
    __shared__ double sh[32];
    if (MODE == 2) {
        iter += flags[F_ITBASE]; // batches replayed as a CUDA graph pass the index inside the batch, the base advances on the device
        const int ci = flags[F_CONV_ITER]; // 0: running, -1/-2: stopped before the first iteration, k>0: converged at k
        if (ci != 0 && iter > ci) return;
    }
    double acc0 = 0.0, acc1 = 0.0;
    const int l = threadIdx.x & (T - 1);
    // each CTA takes a contiguous share of the row blocks (the grid is sized to the machine, not to the matrix)
    const int t0 = (int)((long long)nblk * blockIdx.x / gridDim.x), t1 = (int)((long long)nblk * (blockIdx.x + 1) / gridDim.x);
    for (int t = t0; t < t1; ++t) {
        const int r0 = rb[t], r1 = rb[t + 1];
        const int nz0 = __ldg(rowptr + r0), nz1 = __ldg(rowptr + r1);
        const int cnt = nz1 - nz0;
        // phase 1: products into shared memory
        for (int base = 0; base < cnt; base += STREAM_TILE) {
            double v[STREAM_ITEMS];
            int c[STREAM_ITEMS];
#pragma unroll
            for (int i = 0; i < STREAM_ITEMS; ++i) {
                const int j = base + i * RED_THREADS + threadIdx.x;
                const bool ok = j < cnt;
                c[i] = ok ? __ldcs(colind + nz0 + j) : 0;
                v[i] = ok ? __ldcs(vals + nz0 + j) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < STREAM_ITEMS; ++i) {
                const int j = base + i * RED_THREADS + threadIdx.x;
                if (j < cnt) sprod[j] = v[i] * __ldg(x + c[i]);
            }
        }
        __syncthreads();
        // phase 2: row sums
        const int nrow = r1 - r0;
        const int nround = (nrow + RED_THREADS / T - 1) / (RED_THREADS / T); // uniform trip count: shuffles stay convergent
        for (int it = 0; it < nround; ++it) {
            const int rl = it * (RED_THREADS / T) + threadIdx.x / T;
            const bool ok = rl < nrow;
            const int row = r0 + rl;
            double s = 0.0;
            if (ok) {
                const int b = __ldg(rowptr + row) - nz0, e = __ldg(rowptr + row + 1) - nz0;
                for (int j = b + l; j < e; j += T) s += sprod[j];
            }
#pragma unroll
            for (int o = T >> 1; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (ok && l == 0) {
                if (MODE == 0) y[row] = s;
                if (MODE == 1) {
                    const double g = s - aux0[row];
                    y[row] = g;
                    acc0 = fma(g, aux1[row] * g, acc0);
                }
                if (MODE == 2) {
                    const double h = x[row];
                    y[row] = s;
                    acc0 = fma(aux0[row], h, acc0);
                    acc1 = fma(h, s, acc1);
                }
            }
        }
        __syncthreads(); // sprod is reused by the next row block
    }
    if (MODE == 1) {
        double a[1] = {acc0};
        grid_sum_finish<1>(a, partial, flags + F_COUNTER, out, sh, true);
    }
    if (MODE == 2) {
        double a[2] = {acc0, acc1};
        grid_sum_finish<2>(a, partial, flags + F_COUNTER, out, sh, true);
    }
}

// ---------------------------------------------------------------------------------------------------
// SELL-32 SpMV (the default where padding stays small): slices of 32 consecutive rows, one lane per row, entry k of
// lane l at off + k*32 + l.  Every load of column indices and values is one fully coalesced line per warp, the x
// gather of a warp reads 32 neighbouring rows' k-th neighbours (near-contiguous on FE numberings: a few L1 wavefronts
// instead of ~one per lane with CSR), and a row's products are added left to right in a register (the CPU's own
// order; no shared memory, no shuffles).  Same MODEs as k_spmv_stream.
// ---------------------------------------------------------------------------------------------------
__global__ void k_sell_len(const int32_t *__restrict__ rowptr, int n, int nslices, int32_t *__restrict__ slen)
{
    const int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (sl >= nslices) return;
    const int row = sl * 32 + lane;
    int m = row < n ? rowptr[row + 1] - rowptr[row] : 0;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) slen[sl] = 32 * m;
}

// WHAT 0: column indices (padding: the row's own index, value 0), WHAT 1: values
template <int WHAT>
__global__ void k_sell_fill(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, const double *__restrict__ vals,
                            int n, int nslices, const int32_t *__restrict__ soff, int32_t *__restrict__ scol, double *__restrict__ sval)
{
    const int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (sl >= nslices) return;
    const int row = sl * 32 + lane;
    const int off = soff[sl], len = (soff[sl + 1] - off) >> 5;
    const int rb = row < n ? rowptr[row] : 0, rl = row < n ? rowptr[row + 1] - rb : 0;
    for (int k = 0; k < len; ++k) {
        const size_t d = (size_t)off + (size_t)k * 32 + lane;
        if (WHAT == 0) scol[d] = k < rl ? __ldg(colind + rb + k) : (row < n ? row : 0);
        else sval[d] = k < rl ? __ldg(vals + rb + k) : 0.0;
    }
}

template <int MODE>
__global__ void __launch_bounds__(RED_THREADS) k_spmv_sell(const int32_t *__restrict__ soff, const int32_t *__restrict__ scol,
                                                           const double *__restrict__ sval, const double *__restrict__ x, int n,
                                                           int nslices, const double *__restrict__ aux0, const double *__restrict__ aux1,
                                                           double *__restrict__ y, int iter, double *__restrict__ partial,
                                                           int *__restrict__ flags, double *__restrict__ out)
{
    __shared__ double sh[32];
    if (MODE == 2) {
        iter += flags[F_ITBASE]; // batches replayed as a CUDA graph pass the index inside the batch, the base advances on the device
        const int ci = flags[F_CONV_ITER]; // 0: running, -1/-2: stopped before the first iteration, k>0: converged at k
        if (ci != 0 && iter > ci) return;
    }
    double acc0 = 0.0, acc1 = 0.0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int sl = warp; sl < nslices; sl += nwarps) {
        const int off = __ldg(soff + sl), len = (__ldg(soff + sl + 1) - off) >> 5;
        const int32_t *pc = scol + off + lane;
        const double *pv = sval + off + lane;
        double s = 0.0;
        int k = 0;
        for (; k + 4 <= len; k += 4) {
            int c[4];
            double v[4], xv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c[i] = __ldcs(pc + (size_t)(k + i) * 32);
                v[i] = __ldcs(pv + (size_t)(k + i) * 32);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = __ldg(x + c[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) s = __dadd_rn(s, __dmul_rn(v[i], xv[i]));
        }
        for (; k < len; ++k) s = __dadd_rn(s, __dmul_rn(__ldcs(pv + (size_t)k * 32), __ldg(x + __ldcs(pc + (size_t)k * 32))));
        const int row = sl * 32 + lane;
        if (row < n) {
            if (MODE == 0) y[row] = s;
            if (MODE == 1) {
                const double g = s - aux0[row];
                y[row] = g;
                acc0 = fma(g, aux1[row] * g, acc0);
            }
            if (MODE == 2) {
                const double h = x[row];
                y[row] = s;
                acc0 = fma(aux0[row], h, acc0);
                acc1 = fma(h, s, acc1);
            }
        }
    }
    if (MODE == 1) {
        double a[1] = {acc0};
        grid_sum_finish<1>(a, partial, flags + F_COUNTER, out, sh, true);
    }
    if (MODE == 2) {
        double a[2] = {acc0, acc1};
        grid_sum_finish<2>(a, partial, flags + F_COUNTER, out, sh, true);
    }
}

// ---------------------------------------------------------------------------------------------------
// Node-block SpMV for vector spaces ([P,P,P]: dof = node*NC + c).  The NC dof rows of a node share one column
// structure (the node row of the pattern, every node column expanded to NC consecutive dofs), so the kernel walks the
// NODE-level CSR: one warp per node row, one 4-byte column index per NC x NC block instead of one per entry, the
// values read in place from the dof-level CSR (NC coalesced streams), x gathered NC contiguous doubles at a time.
// Bytes per non-zero: 8 + 4/NC^2 instead of 12.  Same MODEs as k_spmv_stream.
// ---------------------------------------------------------------------------------------------------
template <int NC, int MODE>
__global__ void __launch_bounds__(RED_THREADS) k_spmv_nodeblock(const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ ncol,
                                                                const int32_t *__restrict__ rowptr, const double *__restrict__ vals,
                                                                const double *__restrict__ x, int nnode, const double *__restrict__ aux0,
                                                                const double *__restrict__ aux1, double *__restrict__ y, int iter,
                                                                double *__restrict__ partial, int *__restrict__ flags,
                                                                double *__restrict__ out)
{
    __shared__ double sh[32];
    if (MODE == 2) {
        iter += flags[F_ITBASE]; // batches replayed as a CUDA graph pass the index inside the batch, the base advances on the device
        const int ci = flags[F_CONV_ITER]; // 0: running, -1/-2: stopped before the first iteration, k>0: converged at k
        if (ci != 0 && iter > ci) return;
    }
    double acc0 = 0.0, acc1 = 0.0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    // the header of the next node row (node-level and dof-level row pointers) is fetched while the current one is summed
    int i = warp;
    int nb = 0, ne = 0;
    size_t vo[NC];
    if (i < nnode) {
        nb = __ldg(nrowptr + i);
        ne = __ldg(nrowptr + i + 1);
#pragma unroll
        for (int c = 0; c < NC; ++c) vo[c] = (size_t)__ldg(rowptr + NC * i + c);
    }
    for (; i < nnode; i += nwarps) {
        const int len = NC * (ne - nb), nbc = nb;
        const double *v[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] = vals + vo[c];
        const int inext = i + nwarps;
        if (inext < nnode) {
            nb = __ldg(nrowptr + inext);
            ne = __ldg(nrowptr + inext + 1);
#pragma unroll
            for (int c = 0; c < NC; ++c) vo[c] = (size_t)__ldg(rowptr + NC * inext + c);
        }
        double s[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) s[c] = 0.0;
#pragma unroll 4
        for (int k = lane; k < len; k += 32) {
            const int jn = NC == 3 ? (int)(((unsigned)k * 0xAAABu) >> 17) : (NC == 2 ? k >> 1 : k); // k / NC for k < 2^15
            const double xv = __ldg(x + NC * __ldg(ncol + nbc + jn) + (k - NC * jn));
#pragma unroll
            for (int c = 0; c < NC; ++c) s[c] = fma(__ldcs(v[c] + k), xv, s[c]);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) s[c] = warp_sum(s[c]);
        if (lane < NC) {
            double sv = s[0];
#pragma unroll
            for (int c = 1; c < NC; ++c)
                if (lane == c) sv = s[c];
            const int row = NC * i + lane;
            if (MODE == 0) y[row] = sv;
            if (MODE == 1) {
                const double g = sv - aux0[row];
                y[row] = g;
                acc0 = fma(g, aux1[row] * g, acc0);
            }
            if (MODE == 2) {
                const double h = x[row];
                y[row] = sv;
                acc0 = fma(aux0[row], h, acc0);
                acc1 = fma(h, sv, acc1);
            }
        }
    }
    if (MODE == 1) {
        double a[1] = {acc0};
        grid_sum_finish<1>(a, partial, flags + F_COUNTER, out, sh, true);
    }
    if (MODE == 2) {
        double a[2] = {acc0, acc1};
        grid_sum_finish<2>(a, partial, flags + F_COUNTER, out, sh, true);
    }
}

// y = A x  (optionally y = A x - b)
template <int T>
__global__ void __launch_bounds__(RED_THREADS) k_spmv(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                                      const double *__restrict__ vals, const double *__restrict__ x,
                                                      const double *__restrict__ sub, double *__restrict__ y, int n)
{
    const int l = threadIdx.x & (T - 1);
    const int ngroups = gridDim.x * (RED_THREADS / T);
    const int nround = (n + ngroups - 1) / ngroups; // uniform trip count: shuffles stay convergent
    for (int it = 0; it < nround; ++it) {
        const int row = it * ngroups + blockIdx.x * (RED_THREADS / T) + threadIdx.x / T;
        const bool ok = row < n;
        const int rb = ok ? __ldg(rowptr + row) : 0, re = ok ? __ldg(rowptr + row + 1) : 0;
        double s = row_product<T>(colind, vals, x, rb, re, l);
        if (ok && l == 0) y[row] = sub ? s - sub[row] : s;
    }
}

// CG start: G = A x - b ; partial <G, D1 G>
template <int T>
__global__ void __launch_bounds__(RED_THREADS) k_cg_init1(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                                          const double *__restrict__ vals, const double *__restrict__ x,
                                                          const double *__restrict__ b, const double *__restrict__ d1,
                                                          double *__restrict__ G, int n, double *__restrict__ partial,
                                                          int *__restrict__ flags, double *__restrict__ scal)
{
    __shared__ double sh[32];
    const int l = threadIdx.x & (T - 1);
    const int ngroups = gridDim.x * (RED_THREADS / T);
    double acc[1] = {0.0};
    const int nround = (n + ngroups - 1) / ngroups; // uniform trip count: shuffles stay convergent
    for (int it = 0; it < nround; ++it) {
        const int row = it * ngroups + blockIdx.x * (RED_THREADS / T) + threadIdx.x / T;
        const bool ok = row < n;
        const int rb = ok ? __ldg(rowptr + row) : 0, re = ok ? __ldg(rowptr + row + 1) : 0;
        double s = row_product<T>(colind, vals, x, rb, re, l);
        if (ok && l == 0) {
            const double g = s - b[row];
            G[row] = g;
            acc[0] = fma(g, d1[row] * g, acc[0]);
        }
    }
    grid_sum_finish<1>(acc, partial, flags + F_COUNTER, scal + S_GCG0, sh, true);
}

// CG start, second half: H = -D1 G ; eps2 ; "converged before the first iteration"
__global__ void __launch_bounds__(RED_THREADS) k_cg_init2(const double *__restrict__ G, const double *__restrict__ d1,
                                                          double *__restrict__ H, int n, double eps, int *__restrict__ flags,
                                                          double *__restrict__ scal)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) H[i] = -(d1[i] * G[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double gcg = scal[S_GCG0];
        double eps2 = eps * eps;
        if (eps > 0) eps2 *= gcg;
        scal[S_EPS2] = eps2;
        if (gcg != gcg) flags[F_CONV_ITER] = -2; // NaN: bad matrix (the reference asserts)
        else if (gcg < 1e-30) flags[F_CONV_ITER] = -1;
    }
}

// K1: AH = A H ; <G,H>, <H,AH>
template <int T>
__global__ void __launch_bounds__(RED_THREADS) k_cg_spmv(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                                         const double *__restrict__ vals, const double *__restrict__ H,
                                                         const double *__restrict__ G, double *__restrict__ AH, int n, int iter,
                                                         double *__restrict__ partial, int *__restrict__ flags,
                                                         double *__restrict__ scal)
{
    __shared__ double sh[32];
    {
        iter += flags[F_ITBASE]; // batches replayed as a CUDA graph pass the index inside the batch, the base advances on the device
        const int ci = flags[F_CONV_ITER]; // 0: running, -1/-2: stopped before the first iteration, k>0: converged at k
        if (ci != 0 && iter > ci) return;
    }
    const int l = threadIdx.x & (T - 1);
    const int ngroups = gridDim.x * (RED_THREADS / T);
    double acc[2] = {0.0, 0.0};
    const int nround = (n + ngroups - 1) / ngroups;
    for (int it = 0; it < nround; ++it) {
        const int row = it * ngroups + blockIdx.x * (RED_THREADS / T) + threadIdx.x / T;
        const bool ok = row < n;
        const int rb = ok ? __ldg(rowptr + row) : 0, re = ok ? __ldg(rowptr + row + 1) : 0;
        double s = row_product<T>(colind, vals, H, rb, re, l);
        if (ok && l == 0) {
            const double h = H[row];
            AH[row] = s;
            acc[0] = fma(G[row], h, acc[0]);
            acc[1] = fma(h, s, acc[1]);
        }
    }
    grid_sum_finish<2>(acc, partial, flags + F_COUNTER, scal + S_GH, sh, true);
}

// K2: G += rho AH ; <G, D1 G> -> scal[S_GCG0 + (iter & 1)]
__global__ void __launch_bounds__(RED_THREADS) k_cg_update_g(double *__restrict__ G, const double *__restrict__ AH,
                                                             const double *__restrict__ d1, int n, int iter,
                                                             double *__restrict__ partial, int *__restrict__ flags,
                                                             double *__restrict__ scal)
{
    __shared__ double sh[32];
    {
        iter += flags[F_ITBASE]; // batches replayed as a CUDA graph pass the index inside the batch, the base advances on the device
        const int ci = flags[F_CONV_ITER]; // 0: running, -1/-2: stopped before the first iteration, k>0: converged at k
        if (ci != 0 && iter > ci) return;
    }
    const double rho = -scal[S_GH] / scal[S_HAH];
    double acc[1] = {0.0};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        const double g = fma(rho, AH[i], G[i]);
        G[i] = g;
        acc[0] = fma(g, d1[i] * g, acc[0]);
    }
    grid_sum_finish<1>(acc, partial, flags + F_COUNTER, scal + S_GCG0 + (iter & 1), sh, true);
}

// K3: x += rho H ; H = gamma H - D1 G ; convergence test
__global__ void __launch_bounds__(RED_THREADS) k_cg_update_xh(double *__restrict__ x, double *__restrict__ H,
                                                              const double *__restrict__ G, const double *__restrict__ d1, int n,
                                                              int iter, int *__restrict__ flags, const double *__restrict__ scal)
{
    {
        iter += flags[F_ITBASE]; // batches replayed as a CUDA graph pass the index inside the batch, the base advances on the device
        const int ci = flags[F_CONV_ITER]; // 0: running, -1/-2: stopped before the first iteration, k>0: converged at k
        if (ci != 0 && iter > ci) return;
    }
    const double rho = -scal[S_GH] / scal[S_HAH];
    const double gcg = scal[S_GCG0 + (iter & 1)], gcgp = scal[S_GCG0 + ((iter - 1) & 1)];
    const double gamma = gcg / gcgp;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        const double h = H[i];
        x[i] = fma(rho, h, x[i]);
        H[i] = gamma * h - d1[i] * G[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && gcg < scal[S_EPS2]) {
        // other blocks of this launch test `iter > conv_iter`, which stays false for them
        flags[F_CONV_ITER] = iter;
    }
}

// ---------------------------------------------------------------------------------------------------
// preconditioner set-up: diag, gettgv statistics, D1, SetInitWithBC
// ---------------------------------------------------------------------------------------------------
// pass 0: max of the diagonal.  pass 1: largest diagonal value < ref, and number of entries == ref
__global__ void __launch_bounds__(RED_THREADS) k_diag_stats(const int32_t *__restrict__ diagpos, const double *__restrict__ vals, int n,
                                                            int pass, double ref, double *__restrict__ partial,
                                                            int *__restrict__ flags, double *__restrict__ out)
{
    __shared__ double sh[32];
    __shared__ bool last;
    double m = -INFINITY, cnt = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        const int p = diagpos[i];
        if (p < 0) continue;
        const double a = vals[p];
        if (pass == 0) m = fmax(m, a);
        else if (a == ref) cnt += 1;
        else if (a < ref) m = fmax(m, a);
    }
    // block max
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = m;
    __syncthreads();
    if (w == 0) {
        m = l < (int)(blockDim.x >> 5) ? sh[l] : -INFINITY;
        for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    double c = block_sum(cnt, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = m;
        partial[gridDim.x + blockIdx.x] = c;
        __threadfence();
        last = (atomicAdd(flags + F_COUNTER, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    {   // the last block to retire folds the per-block results, all threads taking part
        double mm = -INFINITY, cc = 0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            mm = fmax(mm, __ldcg(partial + i));
            cc += __ldcg(partial + gridDim.x + i);
        }
        for (int o = 16; o; o >>= 1) mm = fmax(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        __syncthreads();
        if (l == 0) sh[w] = mm;
        __syncthreads();
        if (w == 0) {
            mm = l < (int)(blockDim.x >> 5) ? sh[l] : -INFINITY;
            for (int o = 16; o; o >>= 1) mm = fmax(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        }
        cc = block_sum(cc, sh);
        if (threadIdx.x == 0) {
            out[0] = mm;
            out[1] = cc;
            flags[F_COUNTER] = 0;
        }
    }
}

__global__ void k_precond(const int32_t *__restrict__ diagpos, const double *__restrict__ vals, int n, double *__restrict__ d1,
                          int has_tgv, double ttgv, double tgv, const double *__restrict__ b, double *__restrict__ x)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = diagpos[i];
    const double a = p >= 0 ? vals[p] : 0.0;
    d1[i] = (a * a < 1e-60) ? 1.0 : 1.0 / a;
    if (has_tgv && a == ttgv) x[i] = b[i] / tgv; // SetInitWithBC
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
bool ff_is_distributed(ffcuda_matrix *A);                              // comm.cu
void ff_halo_exchange(ffcuda_matrix *A, double *v);                    // comm.cu (no-op on one GPU)
void ff_allreduce(ffcuda_matrix *A, double *d, int count, int op_max); // comm.cu (no-op on one GPU)

static int *ctx_flags(ffcuda_ctx *ctx) { return reinterpret_cast<int *>(ctx->d_scal + 32); }
__global__ void k_cg_advance(int *flags, int by) { flags[F_ITBASE] += by; }

static void ensure_partial(ffcuda_ctx *ctx, size_t ndoubles)
{
    if (ctx->partial_cap >= ndoubles) return;
    if (ctx->d_partial) cudaFree(ctx->d_partial);
    ctx->d_partial = nullptr;
    FF_CUDA(cudaMalloc((void **)&ctx->d_partial, ndoubles * sizeof(double)));
    ctx->partial_cap = ndoubles;
}

static int pick_T(const ffcuda_matrix *A)
{
    const double avg = A->n ? (double)A->nnz / A->n : 1.0;
    if (avg <= 6) return 4;
    if (avg <= 12) return 8;
    if (avg <= 40) return 16;
    return 32;
}

static int grid_for(const ffcuda_ctx *ctx, size_t work_threads)
{
    size_t need = (work_threads + RED_THREADS - 1) / RED_THREADS;
    size_t cap = (size_t)ctx->sm_count * 8;
    return (int)std::max<size_t>(1, std::min(need, cap));
}

#define FF_DISPATCH_T(T, ...)                  \
    switch (T) {                               \
    case 4: { constexpr int TT = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int TT = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int TT = 16; __VA_ARGS__; } break; \
    default: { constexpr int TT = 32; __VA_ARGS__; } break; \
    }

// ---- CSR-stream set-up (once per matrix): row-block table; T lanes per row in the reduction phase
static bool stream_prepare(ffcuda_matrix *A)
{
    if (A->stream_state) return A->stream_state > 0;
    ffcuda_ctx *ctx = A->ctx;
    const size_t shmem = ((size_t)STREAM_TILE + (size_t)A->maxrow) * sizeof(double);
    if (A->maxrow <= 0 || shmem > 160 * 1024 || A->nnz == 0) {
        A->stream_state = -1; // rows too long for the shared-memory tile: lane-group kernel
        return false;
    }
    A->stream_nblk = (int)((A->nnz + STREAM_TILE - 1) / STREAM_TILE);
    A->stream_rb.alloc((size_t)A->stream_nblk + 1);
    ff_launch(ctx, "spmv_row_blocks", [&] {
        k_stream_blocks<<<ff_blocks((size_t)A->stream_nblk + 1, 256), 256, 0, ctx->stream>>>(A->rowptr, A->n, A->stream_nblk, A->stream_rb.p);
    });
    const double rows_per_blk = (double)A->n / A->stream_nblk;
    int T = 1;
    while (T < 32 && rows_per_blk * T * 2 <= RED_THREADS) T <<= 1;
    A->stream_T = T;
    A->stream_shmem = shmem;
    {   // persistent grid: as many CTAs as the machine holds at once (threads and shared memory), at most one per block
        int per_sm = 2048 / RED_THREADS;
        const int by_smem = (int)((200 * 1024) / (shmem + 1024));
        per_sm = std::max(1, std::min(per_sm, by_smem));
        A->stream_grid = std::max(1, std::min(A->stream_nblk, ctx->sm_count * per_sm));
    }
    A->stream_state = 1;
    return true;
}

// ---- SELL-32 set-up: structure once per matrix, values whenever they changed since the last packing
static constexpr double SELL_MAX_PADDING = 1.15; // padded entries / nnz above which the CSR-stream kernel is used instead

static bool sell_prepare(ffcuda_matrix *A)
{
    if (A->sell_state < 0) return false;
    ffcuda_ctx *ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const int ns = (A->n + 31) / 32;
    if (A->sell_state == 0) {
        if (A->nnz == 0 || A->n == 0) {
            A->sell_state = -1;
            return false;
        }
        DBuf<int32_t> slen;
        slen.alloc((size_t)ns + 1);
        FF_CUDA(cudaMemsetAsync(slen.p + ns, 0, sizeof(int32_t), st));
        ff_launch(ctx, "spmv_sell_len", [&] { k_sell_len<<<ff_blocks((size_t)ns * 32, 256), 256, 0, st>>>(A->rowptr, A->n, ns, slen.p); });
        A->sell_off.alloc((size_t)ns + 1);
        int64_t total = 0;
        ff_exclusive_scan_i32(ctx, slen.p, A->sell_off.p, (size_t)ns + 1, &total); // synchronises the stream
        // (a padded total beyond int32 wraps to something far from nnz and is rejected by the same test)
        if (total < A->nnz || (double)total > SELL_MAX_PADDING * (double)A->nnz) {
            A->sell_off.release();
            A->sell_state = -1;
            return false;
        }
        A->sell_nslices = ns;
        A->sell_entries = total;
        A->sell_col.alloc((size_t)total);
        A->sell_val.alloc((size_t)total);
        ff_launch(ctx, "spmv_sell_cols", [&] {
            k_sell_fill<0><<<ff_blocks((size_t)ns * 32, 256), 256, 0, st>>>(A->rowptr, ff_matrix_colind(A), nullptr, A->n, ns, A->sell_off.p, A->sell_col.p, nullptr);
        });
        A->sell_state = 1;
        A->sell_epoch = 0;
    }
    if (A->sell_epoch != A->vals_epoch) {
        ff_launch(ctx, "spmv_sell_vals", [&] {
            k_sell_fill<1><<<ff_blocks((size_t)ns * 32, 256), 256, 0, st>>>(A->rowptr, nullptr, A->vals.p, A->n, ns, A->sell_off.p, nullptr, A->sell_val.p);
        });
        A->sell_epoch = A->vals_epoch;
    }
    return true;
}

static int sell_grid(const ffcuda_matrix *A)
{
    const int warps_per_block = RED_THREADS / 32;
    const int need = (A->sell_nslices + warps_per_block - 1) / warps_per_block;
    return std::max(1, std::min(need, A->ctx->sm_count * (2048 / RED_THREADS)));
}

template <int MODE>
static void sell_launch(ffcuda_matrix *A, const char *name, const double *x, const double *aux0, const double *aux1, double *y, int iter,
                        double *partial, int *flags, double *out)
{
    ffcuda_ctx *ctx = A->ctx;
    ff_launch(ctx, name, [&] {
        k_spmv_sell<MODE><<<sell_grid(A), RED_THREADS, 0, ctx->stream>>>(A->sell_off.p, A->sell_col.p, A->sell_val.p, x, A->n,
                                                                         A->sell_nslices, aux0, aux1, y, iter, partial, flags, out);
    });
}

#define FF_DISPATCH_ST(T, ...)                                \
    switch (T) {                                              \
    case 1: { constexpr int TT = 1; __VA_ARGS__; } break;     \
    case 2: { constexpr int TT = 2; __VA_ARGS__; } break;     \
    case 4: { constexpr int TT = 4; __VA_ARGS__; } break;     \
    case 8: { constexpr int TT = 8; __VA_ARGS__; } break;     \
    case 16: { constexpr int TT = 16; __VA_ARGS__; } break;   \
    default: { constexpr int TT = 32; __VA_ARGS__; } break;   \
    }

template <int MODE>
static void stream_launch(ffcuda_matrix *A, const char *name, const double *x, const double *aux0, const double *aux1, double *y,
                          int iter, double *partial, int *flags, double *out)
{
    ffcuda_ctx *ctx = A->ctx;
    FF_DISPATCH_ST(A->stream_T, {
        auto kern = k_spmv_stream<TT, MODE>;
        if (A->stream_shmem > 48 * 1024)
            FF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A->stream_shmem));
        ff_launch(ctx, name, [&] {
            kern<<<A->stream_grid, RED_THREADS, A->stream_shmem, ctx->stream>>>(A->rowptr, ff_matrix_colind(A), A->vals.p, x, A->stream_rb.p,
                                                                                 A->stream_nblk, aux0, aux1, y, iter, partial, flags, out);
        });
    });
}

// ---- node-block path: matrices assembled on a vector space (values in the pattern's dof-level CSR layout)
static bool nodeblock_ok(const ffcuda_matrix *A)
{
    const ffcuda_pattern *P = A->pattern;
    return P && (P->ncomp == 2 || P->ncomp == 3) && A->rowptr == P->rowptr && P->maxrow_node * P->ncomp < 32768 && A->nnz > 0;
}
static int nodeblock_grid(const ffcuda_matrix *A)
{
    const int warps_per_block = RED_THREADS / 32;
    const int need = (A->pattern->nrows_node + warps_per_block - 1) / warps_per_block;
    return std::max(1, std::min(need, A->ctx->sm_count * (2048 / RED_THREADS)));
}
template <int MODE>
static void nodeblock_launch(ffcuda_matrix *A, const char *name, const double *x, const double *aux0, const double *aux1, double *y,
                             int iter, double *partial, int *flags, double *out)
{
    ffcuda_ctx *ctx = A->ctx;
    const ffcuda_pattern *P = A->pattern;
    ff_launch(ctx, name, [&] {
        if (P->ncomp == 3)
            k_spmv_nodeblock<3, MODE><<<nodeblock_grid(A), RED_THREADS, 0, ctx->stream>>>(P->nrowptr.p, P->ncol.p, A->rowptr, A->vals.p, x,
                                                                                          P->nrows_node, aux0, aux1, y, iter, partial, flags, out);
        else
            k_spmv_nodeblock<2, MODE><<<nodeblock_grid(A), RED_THREADS, 0, ctx->stream>>>(P->nrowptr.p, P->ncol.p, A->rowptr, A->vals.p, x,
                                                                                          P->nrows_node, aux0, aux1, y, iter, partial, flags, out);
    });
}

static void spmv_launch(ffcuda_matrix *A, const double *x, const double *sub, double *y)
{
    ffcuda_ctx *ctx = A->ctx;
    if (!sub && nodeblock_ok(A)) {
        nodeblock_launch<0>(A, "spmv", x, nullptr, nullptr, y, 0, nullptr, nullptr, nullptr);
        return;
    }
    if (!sub && sell_prepare(A)) {
        sell_launch<0>(A, "spmv", x, nullptr, nullptr, y, 0, nullptr, nullptr, nullptr);
        return;
    }
    if (!sub && stream_prepare(A)) {
        stream_launch<0>(A, "spmv", x, nullptr, nullptr, y, 0, nullptr, nullptr, nullptr);
        return;
    }
    const int T = pick_T(A);
    const int grid = grid_for(ctx, (size_t)A->n * T);
    FF_DISPATCH_T(T, ff_launch(ctx, "spmv", [&] {
                      k_spmv<TT><<<grid, RED_THREADS, 0, ctx->stream>>>(A->rowptr, ff_matrix_colind(A), A->vals.p, x, sub, y, A->n);
                  }));
}

extern "C" int ffcuda_spmv(ffcuda_matrix *A, ffcuda_vec *x, ffcuda_vec *y)
{
    FF_API_BEGIN
    FF_REQUIRE(A && x && y, "ffcuda_spmv: null argument");
    FF_REQUIRE(x->n >= A->ncols, "ffcuda_spmv: x is shorter than the number of (owned + ghost) columns");
    FF_REQUIRE(y->n >= A->n, "ffcuda_spmv: y is shorter than the number of rows");
    FF_REQUIRE(x->d.p != y->d.p, "ffcuda_spmv: x and y must be different vectors");
    ff_enter(A->ctx);
    ff_matrix_touch(A);
    ff_halo_exchange(A, x->d.p);
    spmv_launch(A, x->d.p, nullptr, y->d.p);
    if (A->ctx->p2p && ff_is_distributed(A)) { // a neighbour that never answered must not pass as a result
        int to = 0;
        FF_CUDA(ff_memcpy_sync(A->ctx, &to, reinterpret_cast<char *>(A->ctx->d_scal + FF_P2P_DESC_OFF) + offsetof(P2PDesc, timed_out),
                               sizeof(int), cudaMemcpyDeviceToHost));
        FF_REQUIRE(to == 0, "ffcuda_spmv: a peer rank did not answer within the spin limit (halo exchange timed out)");
    }
    FF_API_END(A ? A->ctx : nullptr)
}

// gettgv (femlib/HashMatrix.cpp:1341-1371): largest diagonal value, its multiplicity, and the next one (ratio 1e6)
static void detect_tgv(ffcuda_matrix *A, double *partial, double *ttgv_out, long *ntgv_out)
{
    ffcuda_ctx *ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const int n = A->n;
    const int grid_v = grid_for(ctx, (size_t)n);
    double *scal = ctx->d_scal;
    int *flags = ctx_flags(ctx);
    double *hs = ctx->h_scal;
    ff_launch(ctx, "cg_diag_stats", [&] { k_diag_stats<<<grid_v, RED_THREADS, 0, st>>>(A->diagpos, A->vals.p, n, 0, 0.0, partial, flags, scal + S_TMP0); });
    ff_allreduce(A, scal + S_TMP0, 1, 1);
    FF_CUDA(cudaMemcpyAsync(hs, scal + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    double ttgv = hs[0];
    ff_launch(ctx, "cg_diag_stats", [&] { k_diag_stats<<<grid_v, RED_THREADS, 0, st>>>(A->diagpos, A->vals.p, n, 1, ttgv, partial, flags, scal + S_TMP0); });
    ff_allreduce(A, scal + S_TMP0, 1, 1);
    ff_allreduce(A, scal + S_TMP1, 1, 0);
    FF_CUDA(cudaMemcpyAsync(hs, scal + S_TMP0, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    double max1 = hs[0];
    long ntgv = (long)hs[1];
    if (!(ttgv > 0)) { // the reference starts its scan from ttgv = max1 = 0
        ttgv = 0;
        ntgv = 0;
    }
    if (!(max1 > 0)) max1 = 0;
    if (max1 * 1e6 > ttgv) {
        ttgv = 0;
        ntgv = 0;
    }
    *ttgv_out = ttgv;
    *ntgv_out = ntgv;
}

static void cg_device(ffcuda_matrix *A, const double *b, double *x, double eps, int itmax, double tgv, int *iters, int *converged,
                      double *gcg_out)
{
    ffcuda_ctx *ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const int n = A->n, ncols = A->ncols;
    FF_REQUIRE(A->diagpos, "matrix has no diagonal index");
    ff_matrix_touch(A);
    if (itmax <= 0) itmax = n;
    if (!A->wG.p) {
        A->wG.alloc(n);
        A->wAH.alloc(n);
        A->wD1.alloc(n);
        A->wH.alloc(ncols);
        if (ncols > n) A->wX.alloc(ncols);
    }
    double *G = A->wG.p, *H = A->wH.p, *AH = A->wAH.p, *D1 = A->wD1.p;
    double *scal = ctx->d_scal;
    int *flags = ctx_flags(ctx);
    const int T = pick_T(A);
    const bool nblock = nodeblock_ok(A);
    const bool sell = !nblock && sell_prepare(A);
    const bool streamed = !nblock && !sell && stream_prepare(A);
    const int grid_s = nblock ? nodeblock_grid(A) : sell ? sell_grid(A) : streamed ? A->stream_grid : grid_for(ctx, (size_t)n * T),
              grid_v = grid_for(ctx, (size_t)n);
    ensure_partial(ctx, 2 * (size_t)std::max(grid_s, grid_v) + 16);
    double *partial = ctx->d_partial;
    FF_CUDA(cudaMemsetAsync(scal, 0, 64 * sizeof(double), st));
    // distributed solve with peer mailboxes: the kernels that form <G,H>, <H,AH>, <G,CG> all-reduce them themselves
    const int fused = (ctx->p2p && ff_is_distributed(A)) ? 1 : 0;
    FF_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(scal + FF_P2P_DESC_OFF) + offsetof(P2PDesc, fused), &fused, sizeof(int),
                            cudaMemcpyHostToDevice, st));
    if (ctx->p2p) // a time-out of an earlier solve is that solve's error, not this one's
        FF_CUDA(cudaMemsetAsync(reinterpret_cast<char *>(scal + FF_P2P_DESC_OFF) + offsetof(P2PDesc, timed_out), 0, sizeof(int), st));

    double *hs = ctx->h_scal;
    double ttgv = 0;
    long ntgv = 0;
    detect_tgv(A, partial, &ttgv, &ntgv);
    ff_launch(ctx, "cg_precond", [&] {
        k_precond<<<ff_blocks(n, 256), 256, 0, st>>>(A->diagpos, A->vals.p, n, D1, ntgv > 0, ttgv, tgv, b, x);
    });

    // --- G = A x - b, H = -D1 G
    const double *xin = x;
    if (ncols > n) {
        FF_CUDA(cudaMemcpyAsync(A->wX.p, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        ff_halo_exchange(A, A->wX.p);
        xin = A->wX.p;
    }
    if (nblock)
        nodeblock_launch<1>(A, "cg_init_spmv", xin, b, D1, G, 0, partial, flags, scal + S_GCG0);
    else if (sell)
        sell_launch<1>(A, "cg_init_spmv", xin, b, D1, G, 0, partial, flags, scal + S_GCG0);
    else if (streamed)
        stream_launch<1>(A, "cg_init_spmv", xin, b, D1, G, 0, partial, flags, scal + S_GCG0);
    else
        FF_DISPATCH_T(T, ff_launch(ctx, "cg_init_spmv", [&] {
                          k_cg_init1<TT><<<grid_s, RED_THREADS, 0, st>>>(A->rowptr, ff_matrix_colind(A), A->vals.p, xin, b, D1, G, n, partial, flags, scal);
                      }));
    if (!fused) ff_allreduce(A, scal + S_GCG0, 1, 0);
    ff_launch(ctx, "cg_init_h", [&] { k_cg_init2<<<grid_v, RED_THREADS, 0, st>>>(G, D1, H, n, eps, flags, scal); });

    // --- iterations, enqueued in batches; the flag of batch k is inspected while batch k+1 runs
    int *hflags = reinterpret_cast<int *>(ctx->h_scal + 32); // 2 slots of 4 ints
    cudaEvent_t ev[2];
    FF_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    FF_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int it = 0, slot = 0, batch = 4;
    bool pending[2] = {false, false}, done = false;
    try {
        FF_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        FF_CUDA(cudaStreamSynchronize(st));
        done = hflags[F_CONV_ITER] != 0;
        // Single-GPU solves: a batch of GB iterations (3 kernels each + the kernel that advances the iteration base) is
        // captured ONCE as a CUDA graph and replayed - the kernels take the index inside the batch and add the base kept
        // on the device.  The gaps between dependent launches shrink (square(1000): the kernels of an iteration take
        // 46 us, the iteration took 70 us), the polling of the convergence flag is unchanged.
        constexpr int GB = 16;
        cudaGraphExec_t gexec = nullptr;
        const bool use_graph = !ff_is_distributed(A) && !ctx->prof && itmax >= GB && !(getenv("FFCUDA_CG_GRAPH") && atoi(getenv("FFCUDA_CG_GRAPH")) == 0);
        auto enqueue_iteration = [&](int itk) {
            if (nblock)
                nodeblock_launch<2>(A, "cg_spmv_dots", H, G, nullptr, AH, itk, partial, flags, scal + S_GH);
            else if (sell)
                sell_launch<2>(A, "cg_spmv_dots", H, G, nullptr, AH, itk, partial, flags, scal + S_GH);
            else if (streamed)
                stream_launch<2>(A, "cg_spmv_dots", H, G, nullptr, AH, itk, partial, flags, scal + S_GH);
            else
                FF_DISPATCH_T(T, ff_launch(ctx, "cg_spmv_dots", [&] {
                                  k_cg_spmv<TT><<<grid_s, RED_THREADS, 0, st>>>(A->rowptr, ff_matrix_colind(A), A->vals.p, H, G, AH, n, itk, partial, flags, scal);
                              }));
            ff_launch(ctx, "cg_update_g", [&] { k_cg_update_g<<<grid_v, RED_THREADS, 0, st>>>(G, AH, D1, n, itk, partial, flags, scal); });
            ff_launch(ctx, "cg_update_xh", [&] { k_cg_update_xh<<<grid_v, RED_THREADS, 0, st>>>(x, H, G, D1, n, itk, flags, scal); });
        };
        if (use_graph) {
            cudaGraph_t graph = nullptr;
            const int64_t l0 = ctx->launches;
            FF_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                for (int k = 1; k <= GB; ++k) enqueue_iteration(k);
                k_cg_advance<<<1, 1, 0, st>>>(flags, GB);
            } catch (...) {
                cudaStreamEndCapture(st, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            FF_CUDA(cudaStreamEndCapture(st, &graph));
            ctx->launches = l0; // (counted when the graph is launched)
            const cudaError_t ge = cudaGraphInstantiate(&gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (ge != cudaSuccess) {
                cudaGetLastError();
                gexec = nullptr;
            }
        }
        int itbase = 0; // host mirror of flags[F_ITBASE]
        while (!done && it < itmax) {
            if (gexec && itmax - it >= GB) {
                if (cudaGraphLaunch(gexec, st) != cudaSuccess) {
                    cudaGraphExecDestroy(gexec);
                    throw FFError("ffcuda: cudaGraphLaunch failed in the CG");
                }
                ctx->launches += 3 * GB + 1;
                it += GB;
                itbase += GB;
                FF_CUDA(cudaMemcpyAsync(hflags + 4 * slot, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
                FF_CUDA(cudaEventRecord(ev[slot], st));
                pending[slot] = true;
                const int prevg = slot ^ 1;
                if (pending[prevg]) {
                    FF_CUDA(cudaEventSynchronize(ev[prevg]));
                    pending[prevg] = false;
                    if (hflags[4 * prevg + F_CONV_ITER] != 0) done = true;
                }
                slot ^= 1;
                continue;
            }
            const int nb = std::min(batch, itmax - it);
            for (int k = 0; k < nb; ++k) {
                ++it;
                if (itbase) { // the tail after graph batches: index relative to the base on the device
                    enqueue_iteration(it - itbase);
                    continue;
                }
                ff_halo_exchange(A, H);
                if (nblock)
                    nodeblock_launch<2>(A, "cg_spmv_dots", H, G, nullptr, AH, it, partial, flags, scal + S_GH);
                else if (sell)
                    sell_launch<2>(A, "cg_spmv_dots", H, G, nullptr, AH, it, partial, flags, scal + S_GH);
                else if (streamed)
                    stream_launch<2>(A, "cg_spmv_dots", H, G, nullptr, AH, it, partial, flags, scal + S_GH);
                else
                    FF_DISPATCH_T(T, ff_launch(ctx, "cg_spmv_dots", [&] {
                                      k_cg_spmv<TT><<<grid_s, RED_THREADS, 0, st>>>(A->rowptr, ff_matrix_colind(A), A->vals.p, H, G, AH, n, it, partial, flags, scal);
                                  }));
                if (!fused) ff_allreduce(A, scal + S_GH, 2, 0);
                ff_launch(ctx, "cg_update_g", [&] { k_cg_update_g<<<grid_v, RED_THREADS, 0, st>>>(G, AH, D1, n, it, partial, flags, scal); });
                if (!fused) ff_allreduce(A, scal + S_GCG0 + (it & 1), 1, 0);
                ff_launch(ctx, "cg_update_xh", [&] { k_cg_update_xh<<<grid_v, RED_THREADS, 0, st>>>(x, H, G, D1, n, it, flags, scal); });
            }
            FF_CUDA(cudaMemcpyAsync(hflags + 4 * slot, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
            FF_CUDA(cudaEventRecord(ev[slot], st));
            pending[slot] = true;
            const int prev = slot ^ 1;
            if (pending[prev]) {
                FF_CUDA(cudaEventSynchronize(ev[prev]));
                pending[prev] = false;
                if (hflags[4 * prev + F_CONV_ITER] != 0) done = true;
            }
            slot ^= 1;
            if (batch < 16) batch *= 2;
        }
        FF_CUDA(cudaStreamSynchronize(st));
        if (gexec) cudaGraphExecDestroy(gexec);
    } catch (...) {
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
        throw;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    FF_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaMemcpyAsync(hs, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    if (ctx->p2p && ff_is_distributed(A)) {
        int to = 0;
        FF_CUDA(ff_memcpy_sync(ctx, &to, reinterpret_cast<char *>(scal + FF_P2P_DESC_OFF) + offsetof(P2PDesc, timed_out), sizeof(int),
                               cudaMemcpyDeviceToHost));
        FF_REQUIRE(to == 0, "CG: a peer rank did not answer within the spin limit (multi-GPU exchange timed out)");
    }
    const int ci = hflags[F_CONV_ITER];
    FF_REQUIRE(ci != -2, "CG: <g,Cg> is NaN (bad matrix)");
    const int ret = ci == -1 ? 2 : (ci > 0 ? 1 : 0);
    const int nit = ret == 2 ? 0 : (ret == 1 ? ci : it);
    if (iters) *iters = nit;
    if (converged) *converged = ret;
    if (gcg_out) *gcg_out = hs[S_GCG0 + (nit & 1)];
    ctx->last_cg_eps2 = hs[S_EPS2]; // the absolute threshold the solve stopped on (eps^2 * gCg0 for eps > 0)
}

extern "C" int ffcuda_cg_stop_threshold(ffcuda_matrix *A, double *eps2)
{
    FF_API_BEGIN
    FF_REQUIRE(A && eps2, "ffcuda_cg_stop_threshold: null argument");
    *eps2 = A->ctx->last_cg_eps2;
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_cg(ffcuda_matrix *A, ffcuda_vec *b, ffcuda_vec *x, double eps, int itmax, double tgv, int *iters,
                         int *converged, double *gcg)
{
    FF_API_BEGIN
    FF_REQUIRE(A && b && x, "ffcuda_cg: null argument");
    FF_REQUIRE(!A->rect, "ffcuda_cg is for square matrices (this one is rectangular: products and hand-off only)");
    FF_REQUIRE(b->n >= A->n && x->n >= A->n, "ffcuda_cg: vectors shorter than the matrix");
    ff_enter(A->ctx);
    cg_device(A, b->d.p, x->d.p, eps, itmax, tgv, iters, converged, gcg);
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_cg_host(ffcuda_matrix *A, const double *b, double *x, double eps, int itmax, double tgv, int *iters,
                              int *converged, double *gcg)
{
    FF_API_BEGIN
    FF_REQUIRE(A && b && x, "ffcuda_cg_host: null argument");
    FF_REQUIRE(!A->rect, "ffcuda_cg_host is for square matrices (this one is rectangular: products and hand-off only)");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    DBuf<double> db, dx;
    db.alloc(A->n);
    dx.alloc(A->n);
    FF_CUDA(cudaMemcpyAsync(db.p, b, db.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    FF_CUDA(cudaMemcpyAsync(dx.p, x, dx.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    cg_device(A, db.p, dx.p, eps, itmax, tgv, iters, converged, gcg);
    FF_CUDA(cudaMemcpyAsync(x, dx.p, dx.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}

// ---------------------------------------------------------------------------------------------------
// GMRES: SolverGMRES::dosolver (femlib/VirtualSolverCG.hpp:236-255) = SetInitWithBC + fgmres (femlib/CG.cpp:347-517):
// flexible GMRES(m) with the Jacobi preconditioner on the right (leftC is forced to 0 there), modified Gram-Schmidt,
// Givens rotations, stop when |g[it+1]| / normb < |eps|, normb = norm of the right-hand side with the tgv rows zeroed.
// Like the CG, every scalar stays on the device: the Hessenberg column, the rotations and the convergence test are
// done by the last block to retire of the kernel that forms the last norm; the Gram-Schmidt chain is one kernel per
// basis vector (subtract the previous projection, form the next dot product in the same pass: 2 reads + 1 write of n
// per step); kernels enqueued after convergence are no-ops, so the host polls a flag every few iterations only.
// gs layout (doubles): H (m+2)x(m+1) row-major | rot0 m+2 | rot1 m+2 | g m+1 | y m+1 | normb, aux, relres
// ---------------------------------------------------------------------------------------------------
struct GmresLayout {
    int m;
    __host__ __device__ size_t H(int i, int j) const { return (size_t)i * (m + 1) + j; }
    __host__ __device__ size_t rot0() const { return (size_t)(m + 2) * (m + 1); }
    __host__ __device__ size_t rot1() const { return rot0() + m + 2; }
    __host__ __device__ size_t g() const { return rot1() + m + 2; }
    __host__ __device__ size_t y() const { return g() + m + 1; }
    __host__ __device__ size_t normb() const { return y() + m + 1; }
    __host__ __device__ size_t aux() const { return normb() + 1; }
    __host__ __device__ size_t relres() const { return normb() + 2; }
    __host__ __device__ size_t size() const { return normb() + 4; }
};

// totals of NV per-thread values over the whole grid, fixed summation shape; returns true in thread 0 of the last block to
// retire (tot[] valid there); that thread must reset *counter to 0 when it is done
template <int NV>
__device__ __forceinline__ bool grid_sum_last(const double (&v)[NV], double *__restrict__ partial, int *counter, double (&tot)[NV], double *sh)
{
    __shared__ bool last;
    double r[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) r[k] = block_sum(v[k], sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partial[(size_t)k * gridDim.x + blockIdx.x] = r[k];
        __threadfence();
        last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partial + (size_t)k * gridDim.x + i);
        tot[k] = block_sum(s, sh);
    }
    return threadIdx.x == 0;
}

// what the last block does with a finished sum; on several GPUs the sums are all-reduced first and the k_gm_fin_* kernels
// (one thread) do the same
__device__ __forceinline__ void gm_cycle_start(double sum, double eps, const GmresLayout &L, double *__restrict__ gs)
{
    gs[L.g()] = sqrt(sum);
    if (gs[L.normb()] < 1.e-20 || eps < 0) gs[L.normb()] = 1.0;
}
// the new Hessenberg column goes through the rotations, the new rotation is formed, g is updated and the stopping test is
// taken (CG.cpp:436-471); flags[F_CONV_ITER] = it + 1 on convergence
__device__ __forceinline__ void gm_hessenberg_step(double sum, int it, double eps, const GmresLayout &L, double *__restrict__ gs, int *__restrict__ flags)
{
    double *H = gs, *rot0 = gs + L.rot0(), *rot1 = gs + L.rot1(), *g = gs + L.g();
    const double aux = sqrt(sum);
    gs[L.aux()] = aux;
    H[L.H(it + 1, it)] = aux;
    for (int i = 0; i < it; i++) {
        const double aa = rot0[i] * H[L.H(i, it)] + rot1[i] * H[L.H(i + 1, it)];
        const double bb = -rot1[i] * H[L.H(i, it)] + rot0[i] * H[L.H(i + 1, it)];
        H[L.H(i, it)] = aa;
        H[L.H(i + 1, it)] = bb;
    }
    const double hii = H[L.H(it, it)], hi1 = H[L.H(it + 1, it)];
    const double sq = sqrt(hii * hii + hi1 * hi1);
    rot0[it] = hii / sq;
    rot1[it] = hi1 / sq;
    H[L.H(it, it)] = rot0[it] * hii + rot1[it] * hi1;
    H[L.H(it + 1, it)] = 0.0;
    g[it + 1] = -rot1[it] * g[it];
    g[it] = rot0[it] * g[it];
    const double relres = fabs(g[it + 1]);
    gs[L.relres()] = relres;
    __threadfence();
    if (relres / gs[L.normb()] < fabs(eps)) flags[F_CONV_ITER] = it + 1;
}
__global__ void k_gm_fin_normb(const double *__restrict__ raw, double *__restrict__ out) { *out = sqrt(*raw); }
__global__ void k_gm_fin_resid(const double *__restrict__ raw, double eps, GmresLayout L, double *__restrict__ gs) { gm_cycle_start(*raw, eps, L, gs); }
__global__ void k_gm_fin_mgs(const double *__restrict__ raw, double *__restrict__ hout, const int *__restrict__ flags)
{
    if (flags[F_CONV_ITER] == 0) *hout = *raw;
}
__global__ void k_gm_fin_last(const double *__restrict__ raw, int it, double eps, GmresLayout L, double *__restrict__ gs, int *__restrict__ flags)
{
    if (flags[F_CONV_ITER] == 0) gm_hessenberg_step(*raw, it, eps, L, gs, flags);
}

// normb = || b with the tgv rows zeroed ||  (fgmres: wi = rhs; wi[wbc] = 0; normb = nrm2(Cl wi), Cl = Id)
__global__ void __launch_bounds__(RED_THREADS) k_gm_normb(const double *__restrict__ b, const int32_t *__restrict__ diagpos,
                                                          const double *__restrict__ vals, int has_tgv, double ttgv, int n,
                                                          double *__restrict__ partial, int *__restrict__ flags, double *__restrict__ out,
                                                          double *__restrict__ raw /* distributed: the local sum goes here, see k_gm_fin_* */)
{
    __shared__ double sh[32];
    double v[1] = {0.0}, tot[1];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        const int p = diagpos[i];
        const bool bc = has_tgv && p >= 0 && vals[p] == ttgv;
        const double x = bc ? 0.0 : b[i];
        v[0] += x * x;
    }
    if (grid_sum_last<1>(v, partial, flags + F_COUNTER, tot, sh)) {
        if (raw) *raw = tot[0];
        else *out = sqrt(tot[0]);
        flags[F_COUNTER] = 0;
    }
}

// start of a cycle: W = -(A x - b), g[0] = || W ||, all other cycle scalars restart
__global__ void __launch_bounds__(RED_THREADS) k_gm_resid(const double *__restrict__ b, const double *__restrict__ Ax, double *__restrict__ W,
                                                          int n, double eps, GmresLayout L, double *__restrict__ gs,
                                                          double *__restrict__ partial, int *__restrict__ flags, double *__restrict__ raw)
{
    __shared__ double sh[32];
    double v[1] = {0.0}, tot[1];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        const double r = -(Ax[i] + -1.0 * b[i]);
        W[i] = r;
        v[0] += r * r;
    }
    if (grid_sum_last<1>(v, partial, flags + F_COUNTER, tot, sh)) {
        if (raw) *raw = tot[0];
        else gm_cycle_start(tot[0], eps, L, gs);
        flags[F_COUNTER] = 0;
    }
}

// dst = (1 / *s) * src  (Vi[0] = (1./g[0])*zi, Vi[it+1] = (1./aux)*wi)
__global__ void k_gm_scale(double *__restrict__ dst, const double *__restrict__ src, const double *__restrict__ s, int n,
                           const int *__restrict__ flags)
{
    if (flags[F_CONV_ITER] != 0) return;
    const double a = 1.0 / *s;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) dst[i] = a * src[i];
}

// Vp = C V, C = diag1 (HMatVirtPrecon::addmatmul on a zeroed vector)
__global__ void k_gm_precond(double *__restrict__ Vp, const double *__restrict__ V, const double *__restrict__ d1, int n,
                             const int *__restrict__ flags)
{
    if (flags[F_CONV_ITER] != 0) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) Vp[i] = d1[i] * V[i];
}

// one step of the modified Gram-Schmidt chain: W -= h_prev * Vprev (when there is a previous step), then h = <W, Vcur>
__global__ void __launch_bounds__(RED_THREADS) k_gm_mgs(double *__restrict__ W, const double *__restrict__ Vprev, const double *__restrict__ Vcur,
                                                        const double *__restrict__ hprev, double *__restrict__ hout, int n,
                                                        double *__restrict__ partial, int *__restrict__ flags, double *__restrict__ raw)
{
    if (flags[F_CONV_ITER] != 0) return;
    __shared__ double sh[32];
    double v[1] = {0.0}, tot[1];
    const double a = Vprev ? -*hprev : 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        double w = W[i];
        if (Vprev) {
            w += a * Vprev[i];
            W[i] = w;
        }
        v[0] += w * Vcur[i];
    }
    if (grid_sum_last<1>(v, partial, flags + F_COUNTER, tot, sh)) {
        *(raw ? raw : hout) = tot[0];
        flags[F_COUNTER] = 0;
    }
}

// end of the chain: W -= H(it,it) * V_it, aux = || W ||, then the Hessenberg column goes through the rotations, the new
// rotation is formed, g is updated and the stopping test is taken (CG.cpp:436-471); flags[F_CONV_ITER] = it + 1 on convergence
__global__ void __launch_bounds__(RED_THREADS) k_gm_mgs_last(double *__restrict__ W, const double *__restrict__ Vprev, int n, int it, double eps,
                                                             GmresLayout L, double *__restrict__ gs, double *__restrict__ partial,
                                                             int *__restrict__ flags, double *__restrict__ raw)
{
    if (flags[F_CONV_ITER] != 0) return;
    __shared__ double sh[32];
    double v[1] = {0.0}, tot[1];
    const double a = -gs[L.H(it, it)];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
        const double w = W[i] + a * Vprev[i];
        W[i] = w;
        v[0] += w * w;
    }
    if (grid_sum_last<1>(v, partial, flags + F_COUNTER, tot, sh)) {
        flags[F_COUNTER] = 0;
        if (raw) *raw = tot[0];
        else gm_hessenberg_step(tot[0], it, eps, L, gs, flags);
    }
}

// One Arnoldi step in ONE cooperative kernel (grid = co-resident CTAs, grid-wide barriers): the whole modified
// Gram-Schmidt chain, the norm, the Givens update and the next basis vector.  Every thread keeps its entries of w in
// registers for the whole chain (KW per thread), so a step of the chain reads one basis vector (the one before comes from
// L2) and costs one grid barrier; the dot products are summed in a fixed shape (per-block partials, every block adds
// them in the same order), so the result is reproducible and every block holds the same h.
namespace cg = cooperative_groups;
static constexpr int ARN_THREADS = 512; // one CTA per SM: half as many partial sums and barrier participants as 2 x 256
template <int KW>
__global__ void __launch_bounds__(ARN_THREADS, 1)
    k_gm_arnoldi(double *__restrict__ W, const double *const *__restrict__ Vtab, int n, int it, double eps, GmresLayout L,
                 double *__restrict__ gs, double *__restrict__ partial /* 2 x gridDim */, int *__restrict__ flags)
{
    if (flags[F_CONV_ITER] != 0) return; // (uniform over the grid: no barrier is left waiting)
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[32];
    __shared__ double sbc;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x; // (n < 2^31 / KW entries)
    double w[KW];
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const int idx = gtid + j * gstride;
        w[j] = idx < n ? W[idx] : 0.0;
    }
    double hprev = 0.0;
    int par = 0;
    // grid-wide sum of one value per thread: partials of every block, barrier, every block adds them in the same order
    auto grid_total = [&](double v) -> double {
        const double r = block_sum(v, sh);
        if (threadIdx.x == 0) partial[(size_t)par * gridDim.x + blockIdx.x] = r;
        grid.sync();
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partial + (size_t)par * gridDim.x + i);
        const double t = block_sum(s, sh);
        if (threadIdx.x == 0) sbc = t;
        __syncthreads();
        par ^= 1;
        return sbc;
    };
    for (int i = 0; i <= it; ++i) {
        const double *Vcur = Vtab[i], *Vprev = i ? Vtab[i - 1] : nullptr;
        const double a = -hprev;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int idx = gtid + j * gstride;
            if (idx < n) {
                if (i) w[j] += a * Vprev[idx];
                acc += w[j] * Vcur[idx];
            }
        }
        if (i < it) { // the next basis vector does not depend on h: pull it into L2 while the barrier is crossed
            const double *Vn = Vtab[i + 1];
#pragma unroll
            for (int j = 0; j < KW; ++j) {
                const int idx = gtid + j * gstride;
                if (idx < n && (threadIdx.x & 3) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(Vn + idx));
            }
        }
        hprev = grid_total(acc);
        if (blockIdx.x == 0 && threadIdx.x == 0) gs[L.H(i, it)] = hprev;
    }
    {
        const double *Vit = Vtab[it];
        const double a = -hprev;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int idx = gtid + j * gstride;
            if (idx < n) {
                w[j] += a * Vit[idx];
                acc += w[j] * w[j];
            }
        }
        const double aux = sqrt(grid_total(acc));
        const double s = 1.0 / aux;
        double *Vnext = const_cast<double *>(Vtab[it + 1]);
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const int idx = gtid + j * gstride;
            if (idx < n) {
                W[idx] = w[j];
                Vnext[idx] = s * w[j];
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { // Hessenberg column through the rotations, new rotation, g, stopping test
            double *H = gs, *rot0 = gs + L.rot0(), *rot1 = gs + L.rot1(), *g = gs + L.g();
            gs[L.aux()] = aux;
            H[L.H(it + 1, it)] = aux;
            for (int i = 0; i < it; i++) {
                const double aa = rot0[i] * H[L.H(i, it)] + rot1[i] * H[L.H(i + 1, it)];
                const double bb = -rot1[i] * H[L.H(i, it)] + rot0[i] * H[L.H(i + 1, it)];
                H[L.H(i, it)] = aa;
                H[L.H(i + 1, it)] = bb;
            }
            const double hii = H[L.H(it, it)], hi1 = H[L.H(it + 1, it)];
            const double sq = sqrt(hii * hii + hi1 * hi1);
            rot0[it] = hii / sq;
            rot1[it] = hi1 / sq;
            H[L.H(it, it)] = rot0[it] * hii + rot1[it] * hi1;
            H[L.H(it + 1, it)] = 0.0;
            g[it + 1] = -rot1[it] * g[it];
            g[it] = rot0[it] * g[it];
            const double relres = fabs(g[it + 1]);
            gs[L.relres()] = relres;
            __threadfence();
            if (relres / gs[L.normb()] < fabs(eps)) flags[F_CONV_ITER] = it + 1;
        }
    }
}

// y by back substitution on the rotated Hessenberg matrix (CG.cpp:477-483)
__global__ void k_gm_backsolve(int it, GmresLayout L, double *__restrict__ gs)
{
    if (threadIdx.x || blockIdx.x) return;
    const double *H = gs, *g = gs + L.g();
    double *y = gs + L.y();
    for (int i = it; i >= 0; i--) {
        double g1 = g[i];
        for (int j = i + 1; j < it + 1; j++) g1 = g1 - H[L.H(i, j)] * y[j];
        y[i] = g1 / H[L.H(i, i)];
    }
}

// x = (sum_i y_i Vp_i) + x0, x0 = x (CG.cpp:485-495)
__global__ void k_gm_update_x(double *__restrict__ x, const double *const *__restrict__ Vp, const double *__restrict__ y, int it, int n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < (size_t)n; k += stride) {
        double w = 0.0;
        for (int i = 0; i < it + 1; ++i) w += y[i] * Vp[i][k];
        x[k] = w + x[k];
    }
}

static void gmres_device(ffcuda_matrix *A, const double *b, double *x, double eps, int itmax, int restart, double tgv, int *iters,
                         int *converged, double *relres_out)
{
    ffcuda_ctx *ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const int n = A->n;
    FF_REQUIRE(A->diagpos, "matrix has no diagonal index");
    // several GPUs (a matrix on a distributed mesh): the rows are shared out, so every dot product / norm is a local sum, an
    // all-reduce (peer mailboxes, else NCCL: ff_allreduce) and a one-thread kernel that does what the last block does on one
    // GPU; the vectors A is applied to carry the ghost columns and get them by a halo exchange first.  The sums are formed in
    // rank order on every rank: every rank holds the same Hessenberg matrix and takes the same decisions.  Reference role:
    // MPI_Allreduce per dot product, plugin/mpi/MPICG.cpp:93-101 (MPILinearGMRES there), idp/MPIGMRESmacro.idp.
    const bool dist = ff_is_distributed(A);
    const int ncols = A->ncols;
    FF_REQUIRE(dist || ncols == n, "GMRES: the matrix has ghost columns but no distributed mesh");
    ff_matrix_touch(A);
    if (itmax <= 0) itmax = n;
    if (restart <= 0) restart = 1000; // Data_Sparse_Solver::NbSpace default (femlib/VirtualSolver.hpp:79)
    const int m = (int)std::min<int64_t>(restart, (int64_t)itmax + 2); // storage only: the inner loop leaves at it > itmax
    GmresLayout L{m};
    const int grid_v = grid_for(ctx, (size_t)n);
    ensure_partial(ctx, 2 * (size_t)std::max(grid_v, 1024) + 16);
    double *partial = ctx->d_partial;
    int *flags = ctx_flags(ctx);
    FF_CUDA(cudaMemsetAsync(ctx->d_scal, 0, 64 * sizeof(double), st));
    DBuf<double> gs, D1, W, R, XG;
    gs.alloc(L.size());
    D1.alloc(n);
    W.alloc(n);
    R.alloc(n);
    if (dist) XG.alloc(ncols); // x with its ghost columns
    double *raw = dist ? ctx->d_scal + S_TMP2 : nullptr; // local sums on their way through the all-reduce
    if (dist && ctx->p2p) // a time-out of an earlier solve is that solve's error, not this one's
        FF_CUDA(cudaMemsetAsync(reinterpret_cast<char *>(ctx->d_scal + FF_P2P_DESC_OFF) + offsetof(P2PDesc, timed_out), 0, sizeof(int), st));
    FF_CUDA(cudaMemsetAsync(gs.p, 0, gs.bytes(), st));
    double ttgv = 0;
    long ntgv = 0;
    detect_tgv(A, partial, &ttgv, &ntgv);
    ff_launch(ctx, "gmres_precond", [&] {
        k_precond<<<ff_blocks(n, 256), 256, 0, st>>>(A->diagpos, A->vals.p, n, D1.p, ntgv > 0, ttgv, tgv, b, x);
    });
    ff_launch(ctx, "gmres_normb", [&] {
        k_gm_normb<<<grid_v, RED_THREADS, 0, st>>>(b, A->diagpos, A->vals.p, ntgv > 0, ttgv, n, partial, flags, gs.p + L.normb(), raw);
    });
    if (dist) {
        ff_allreduce(A, raw, 1, 0);
        ff_launch(ctx, "gmres_fin", [&] { k_gm_fin_normb<<<1, 1, 0, st>>>(raw, gs.p + L.normb()); });
    }
    // Krylov vectors in chunks of 16, allocated as the basis grows (the default dimension is 1000)
    constexpr int CH = 16;
    std::deque<DBuf<double>> cV, cP;
    std::vector<double *> hP((size_t)m + 1, nullptr), hV((size_t)m + 2, nullptr);
    DBuf<double *> dP, dV;
    dP.alloc((size_t)m + 1);
    dV.alloc((size_t)m + 2);
    auto vecV = [&](int i) -> double * {
        while ((int)cV.size() * CH <= i) {
            cV.emplace_back();
            cV.back().alloc((size_t)CH * n);
            for (int k = 0; k < CH && (cV.size() - 1) * CH + k <= (size_t)m + 1; ++k) hV[(cV.size() - 1) * CH + k] = cV.back().p + (size_t)k * n;
            FF_CUDA(ff_memcpy_sync(ctx, dV.p, hV.data(), hV.size() * sizeof(double *), cudaMemcpyHostToDevice));
        }
        return cV[i / CH].p + (size_t)(i % CH) * n;
    };
    // the cooperative Arnoldi kernel: as many CTAs as are co-resident, the entries of w in registers (KW per thread)
    int coop_grid = 0, coop_kw = 0;
    if (ctx->gmres_coop && !dist) {
        int dev = 0, can = 0;
        FF_CUDA(cudaGetDevice(&dev));
        FF_CUDA(cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev));
        if (can) {
            for (int kw : {4, 8, 16, 32}) {
                int per_sm = 0;
                cudaError_t e = kw == 4    ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gm_arnoldi<4>, ARN_THREADS, 0)
                                : kw == 8  ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gm_arnoldi<8>, ARN_THREADS, 0)
                                : kw == 16 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gm_arnoldi<16>, ARN_THREADS, 0)
                                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gm_arnoldi<32>, ARN_THREADS, 0);
                FF_CUDA(e);
                const int g = std::min(per_sm, 1) * ctx->sm_count;
                if (g > 0 && (size_t)g * ARN_THREADS * kw >= (size_t)n) {
                    coop_grid = (int)std::min<size_t>((size_t)g, ((size_t)n + ARN_THREADS - 1) / ARN_THREADS); // every SM, as long as a thread has an entry
                    coop_kw = kw;
                    break;
                }
            }
        }
    }
    if (coop_grid) ensure_partial(ctx, 2 * (size_t)std::max({grid_v, 1024, coop_grid}) + 16);
    partial = ctx->d_partial;
    auto vecP = [&](int i) -> double * {
        while ((int)cP.size() * CH <= i) {
            cP.emplace_back();
            cP.back().alloc((size_t)CH * ncols); // A is applied to these: room for the ghost columns
            for (int k = 0; k < CH && (cP.size() - 1) * CH + k <= (size_t)m; ++k) hP[(cP.size() - 1) * CH + k] = cP.back().p + (size_t)k * ncols;
            FF_CUDA(ff_memcpy_sync(ctx, dP.p, hP.data(), hP.size() * sizeof(double *), cudaMemcpyHostToDevice));
        }
        return hP[i];
    };
    int *hflags = reinterpret_cast<int *>(ctx->h_scal + 32);
    const int poll = 8;
    bool conv = false;
    int iter = 0;
    while (true) {
        const double *xin = x;
        if (dist) {
            FF_CUDA(cudaMemcpyAsync(XG.p, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
            ff_halo_exchange(A, XG.p);
            xin = XG.p;
        }
        spmv_launch(A, xin, nullptr, R.p);
        ff_launch(ctx, "gmres_resid", [&] { k_gm_resid<<<grid_v, RED_THREADS, 0, st>>>(b, R.p, W.p, n, eps, L, gs.p, partial, flags, raw); });
        if (dist) {
            ff_allreduce(A, raw, 1, 0);
            ff_launch(ctx, "gmres_fin", [&] { k_gm_fin_resid<<<1, 1, 0, st>>>(raw, eps, L, gs.p); });
        }
        double *V0 = vecV(0);
        ff_launch(ctx, "gmres_scale", [&] { k_gm_scale<<<grid_v, RED_THREADS, 0, st>>>(V0, W.p, gs.p + L.g(), n, flags); });
        int it = 0, it_used = m;
        for (; it < m; ++it) {
            double *Vit = vecV(it), *Vnext = vecV(it + 1), *Pit = vecP(it);
            ff_launch(ctx, "gmres_precond_apply", [&] { k_gm_precond<<<grid_v, RED_THREADS, 0, st>>>(Pit, Vit, D1.p, n, flags); });
            if (dist) ff_halo_exchange(A, Pit);
            spmv_launch(A, Pit, nullptr, W.p);
            if (coop_grid) {
                double *Wp = W.p, *gsp = gs.p;
                const double *const *vt = dV.p;
                int nn = n, iit = it;
                double e = eps;
                void *args[] = {&Wp, &vt, &nn, &iit, &e, &L, &gsp, &partial, &flags};
                const void *fn = coop_kw == 4    ? (const void *)k_gm_arnoldi<4>
                                 : coop_kw == 8  ? (const void *)k_gm_arnoldi<8>
                                 : coop_kw == 16 ? (const void *)k_gm_arnoldi<16>
                                                 : (const void *)k_gm_arnoldi<32>;
                ff_launch(ctx, "gmres_arnoldi", [&] { FF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(coop_grid), dim3(ARN_THREADS), args, 0, st)); });
            } else {
                for (int i = 0; i <= it; ++i) {
                    const double *Vprev = i ? vecV(i - 1) : nullptr;
                    const double *Vcur = vecV(i);
                    ff_launch(ctx, "gmres_mgs", [&] {
                        k_gm_mgs<<<grid_v, RED_THREADS, 0, st>>>(W.p, Vprev, Vcur, gs.p + L.H(i ? i - 1 : 0, it), gs.p + L.H(i, it), n, partial, flags, raw);
                    });
                    if (dist) {
                        ff_allreduce(A, raw, 1, 0);
                        ff_launch(ctx, "gmres_fin", [&] { k_gm_fin_mgs<<<1, 1, 0, st>>>(raw, gs.p + L.H(i, it), flags); });
                    }
                }
                ff_launch(ctx, "gmres_mgs_last", [&] { k_gm_mgs_last<<<grid_v, RED_THREADS, 0, st>>>(W.p, Vit, n, it, eps, L, gs.p, partial, flags, raw); });
                if (dist) {
                    ff_allreduce(A, raw, 1, 0);
                    ff_launch(ctx, "gmres_fin", [&] { k_gm_fin_last<<<1, 1, 0, st>>>(raw, it, eps, L, gs.p, flags); });
                }
                ff_launch(ctx, "gmres_scale", [&] { k_gm_scale<<<grid_v, RED_THREADS, 0, st>>>(Vnext, W.p, gs.p + L.aux(), n, flags); });
            }
            const bool leave = it > itmax; // `if( it > nbitermx) break;`
            if (leave || it == m - 1 || it % poll == poll - 1) {
                FF_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
                FF_CUDA(cudaStreamSynchronize(st));
                if (hflags[F_CONV_ITER] != 0) {
                    conv = true;
                    it_used = hflags[F_CONV_ITER] - 1;
                    break;
                }
            }
            if (leave) {
                it_used = it;
                break;
            }
        }
        iter += it_used; // neither a convergence nor a forced exit counts the iteration they happen in
        const int ity = std::min(it_used, m - 1);
        ff_launch(ctx, "gmres_backsolve", [&] { k_gm_backsolve<<<1, 32, 0, st>>>(ity, L, gs.p); });
        ff_launch(ctx, "gmres_update_x", [&] { k_gm_update_x<<<grid_v, RED_THREADS, 0, st>>>(x, dP.p, gs.p + L.y(), ity, n); });
        if (conv || iter > itmax) break;
    }
    double hr[4];
    FF_CUDA(ff_memcpy_sync(ctx, hr, gs.p + L.normb(), 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (dist && ctx->p2p) {
        int to = 0;
        FF_CUDA(ff_memcpy_sync(ctx, &to, reinterpret_cast<char *>(ctx->d_scal + FF_P2P_DESC_OFF) + offsetof(P2PDesc, timed_out), sizeof(int),
                               cudaMemcpyDeviceToHost));
        FF_REQUIRE(to == 0, "GMRES: a peer rank did not answer within the spin limit (multi-GPU exchange timed out)");
    }
    FF_REQUIRE(hr[2] == hr[2], "GMRES: the residual is NaN (bad matrix)");
    if (iters) *iters = iter;
    if (converged) *converged = conv ? 1 : 0;
    if (relres_out) *relres_out = hr[2] / hr[0];
}

extern "C" int ffcuda_gmres(ffcuda_matrix *A, ffcuda_vec *b, ffcuda_vec *x, double eps, int itmax, int restart, double tgv, int *iters,
                            int *converged, double *relres)
{
    FF_API_BEGIN
    FF_REQUIRE(A && b && x, "ffcuda_gmres: null argument");
    FF_REQUIRE(!A->rect, "ffcuda_gmres is for square matrices (this one is rectangular: products and hand-off only)");
    FF_REQUIRE(b->n >= A->n && x->n >= A->n, "ffcuda_gmres: vectors shorter than the matrix");
    ff_enter(A->ctx);
    gmres_device(A, b->d.p, x->d.p, eps, itmax, restart, tgv, iters, converged, relres);
    FF_API_END(A ? A->ctx : nullptr)
}

extern "C" int ffcuda_gmres_host(ffcuda_matrix *A, const double *b, double *x, double eps, int itmax, int restart, double tgv,
                                 int *iters, int *converged, double *relres)
{
    FF_API_BEGIN
    FF_REQUIRE(A && b && x, "ffcuda_gmres_host: null argument");
    FF_REQUIRE(!A->rect, "ffcuda_gmres_host is for square matrices (this one is rectangular: products and hand-off only)");
    ffcuda_ctx *ctx = A->ctx;
    ff_enter(ctx);
    DBuf<double> db, dx;
    db.alloc(A->n);
    dx.alloc(A->n);
    FF_CUDA(cudaMemcpyAsync(db.p, b, db.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    FF_CUDA(cudaMemcpyAsync(dx.p, x, dx.bytes(), cudaMemcpyHostToDevice, ctx->stream));
    gmres_device(A, db.p, dx.p, eps, itmax, restart, tgv, iters, converged, relres);
    FF_CUDA(cudaMemcpyAsync(x, dx.p, dx.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    FF_API_END(A ? A->ctx : nullptr)
}
