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
#include <algorithm>

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
    const unsigned empty = __ballot_sync(0xffffffffu, row < nrows && m == 0);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        blklen[b] = 32 * m;
        if (m > 0) atomicMax(maxinc, m);
        if (empty) atomicAdd(maxinc + 1, __popc(empty)); // rows without any element
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


// ---------------------------------------------------------------------------------------------------------------
// P1 (ELL layout): block-level kernels.  One warp owns the 32 rows of an ELL block; the block's records and the
// vertex ids of their elements are first brought into shared memory with coalesced / fully independent loads, after
// which all the set logic runs out of shared memory.
// ---------------------------------------------------------------------------------------------------------------
static constexpr uint32_t HT_EMPTY = 0xffffffffu;

__device__ __forceinline__ uint32_t ht_hash(uint32_t v, int log) { return (v * 0x9E3779B1u) >> (32 - log); }

// records of block b -> recs[e*32 + lane] ; vertex ids of every record's element -> cand[lane*cstride + e*4 + i]
// in OWNER-FIRST order: i = 0 is the row's own vertex (local vertex a of the element), i = 1.. the others in an order
// that is an even permutation of the element's own (a^i on tetrahedra, (a+i)%3 on triangles), so that the re-ordered
// simplex has the same orientation.  -1 where there is no record / no 4th vertex.
template <int NLOC>
__device__ __forceinline__ void blk_load(const int32_t *__restrict__ conn, const IncView &V, uint32_t base, int Lb, int lane,
                                         int cstride, uint32_t *recs, int32_t *cand)
{
    for (int e = 0; e < Lb; ++e) recs[e * 32 + lane] = __ldcs(V.inc + base + (size_t)e * 32 + lane);
    int4 *crow = reinterpret_cast<int4 *>(cand + (size_t)lane * cstride);
#pragma unroll 4
    for (int e = 0; e < Lb; ++e) {
        const uint32_t r = recs[e * 32 + lane];
        int4 c = make_int4(-1, -1, -1, -1);
        if (r != FF_NOREC) {
            const size_t k = r >> 4;
            const int a = r & 15;
            if (NLOC == 4) {
                const int4 q = __ldg(reinterpret_cast<const int4 *>(conn) + k);
                // c[i] = q[a ^ i]
                c.x = a == 0 ? q.x : a == 1 ? q.y : a == 2 ? q.z : q.w;
                c.y = a == 0 ? q.y : a == 1 ? q.x : a == 2 ? q.w : q.z;
                c.z = a == 0 ? q.z : a == 1 ? q.w : a == 2 ? q.x : q.y;
                c.w = a == 0 ? q.w : a == 1 ? q.z : a == 2 ? q.y : q.x;
            } else {
                const int q0 = __ldg(conn + 3 * k), q1 = __ldg(conn + 3 * k + 1), q2 = __ldg(conn + 3 * k + 2);
                // c[i] = q[(a + i) % 3]
                c.x = a == 0 ? q0 : a == 1 ? q1 : q2;
                c.y = a == 0 ? q1 : a == 1 ? q2 : q0;
                c.z = a == 0 ? q2 : a == 1 ? q0 : q1;
            }
        }
        crow[e] = c;
    }
    __syncwarp();
}

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

// bitonic sort of u[0..m) (m a power of two) by one warp in shared memory
__device__ __forceinline__ void warp_sort_smem(int32_t *u, int m, int lane)
{
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

// stage tables of the incidence: distinct vertices of the block (slot numbering) and per-record slot bytes.
// One warp per block of 32 rows; shared memory holds only the warp's hash table and vertex list (5 KB), the records and
// the connectivity are read twice (second time from L1/L2): 3x the occupancy of the version that parked them in
// shared memory, which was latency bound (2.8 ms on cube(128)).
template <int NLOC>
__device__ __forceinline__ void owner_first(const int32_t *__restrict__ conn, uint32_t r, int (&c)[4])
{
    const size_t k = r >> 4;
    const int a = r & 15;
    if (NLOC == 4) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(conn) + k);
        // c[i] = q[a ^ i]
        c[0] = a == 0 ? q.x : a == 1 ? q.y : a == 2 ? q.z : q.w;
        c[1] = a == 0 ? q.y : a == 1 ? q.x : a == 2 ? q.w : q.z;
        c[2] = a == 0 ? q.z : a == 1 ? q.w : a == 2 ? q.x : q.y;
        c[3] = a == 0 ? q.w : a == 1 ? q.z : a == 2 ? q.y : q.x;
    } else {
        const int q0 = __ldg(conn + 3 * k), q1 = __ldg(conn + 3 * k + 1), q2 = __ldg(conn + 3 * k + 2);
        // c[i] = q[(a + i) % 3]
        c[0] = a == 0 ? q0 : a == 1 ? q1 : q2;
        c[1] = a == 0 ? q1 : a == 1 ? q2 : q0;
        c[2] = a == 0 ? q2 : a == 1 ? q0 : q1;
        c[3] = -1;
    }
}

static constexpr int STAGE_BT = 512, STAGE_BLOG = 9; // hash table of a block: at most FF_STAGE_MAX = 256 keys
static constexpr int STAGE_WORDS = 2 * STAGE_BT + FF_STAGE_MAX + 4;

template <int NLOC>
__global__ void __launch_bounds__(256) k_block_stage(const int32_t *__restrict__ conn, int nrows, const IncView V,
                                                     uint32_t *__restrict__ loc, int32_t *__restrict__ blkvert,
                                                     int32_t *__restrict__ blkvcnt, int32_t *__restrict__ maxstage)
{
    extern __shared__ uint32_t smem_u[];
    constexpr int BT = STAGE_BT, BLOG = STAGE_BLOG;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *bkey = smem_u + (size_t)w * STAGE_WORDS;
    uint32_t *bslot = bkey + BT;
    int32_t *vlist = reinterpret_cast<int32_t *>(bslot + BT);
    int *bcnt = vlist + FF_STAGE_MAX;
    const int blk = blockIdx.x * (blockDim.x >> 5) + w, nblk = (nrows + 31) >> 5;
    if (blk >= nblk) return;
    const uint32_t base = V.blkoff[blk];
    const int Lb = (int)((V.blkoff[blk + 1] - base) >> 5);
    for (int x = lane; x < BT; x += 32) bkey[x] = HT_EMPTY;
    if (lane == 0) *bcnt = 0;
    __syncwarp();
    // pass 1: the distinct vertices of the block's records
    for (int e = 0; e < Lb; ++e) {
        const uint32_t r = __ldg(V.inc + base + (size_t)e * 32 + lane);
        if (r == FF_NOREC) continue;
        int c[4];
        owner_first<NLOC>(conn, r, c);
#pragma unroll
        for (int b = 0; b < NLOC; ++b) {
            // more than FF_STAGE_MAX distinct vertices: the block will be marked unstaged, stop filling the table
            // (at most 32 lanes overshoot by one insertion each: the 512-slot table cannot fill up)
            if (*reinterpret_cast<volatile int *>(bcnt) > FF_STAGE_MAX) continue;
            const uint32_t v = (uint32_t)c[b];
            uint32_t h = ht_hash(v, BLOG);
            while (true) {
                uint32_t k = bkey[h];
                if (k == v) break;
                if (k == HT_EMPTY) {
                    k = atomicCAS(&bkey[h], HT_EMPTY, v);
                    if (k == HT_EMPTY) {
                        const int slot = atomicAdd(bcnt, 1);
                        bslot[h] = (uint32_t)slot;
                        if (slot < FF_STAGE_MAX) vlist[slot] = (int32_t)v;
                        break;
                    }
                    if (k == v) break;
                }
                h = (h + 1) & (BT - 1);
            }
        }
    }
    __syncwarp();
    const int nv = *bcnt;
    const bool staged = nv <= FF_STAGE_MAX;
    if (staged) {
        // slots are numbered in ASCENDING VERTEX ORDER: the rank of a column inside a row is then a population count
        // over the row's slot bitmap (k_sym_p1_rows), no search and no hashing in the per-assembly symbolic phase
        int m = 32;
        while (m < nv) m <<= 1;
        for (int x = nv + lane; x < m; x += 32) vlist[x] = INT_MAX;
        __syncwarp();
        warp_sort_smem(vlist, m, lane);
        for (int x = lane; x < nv; x += 32) {
            const uint32_t v = (uint32_t)vlist[x];
            uint32_t h = ht_hash(v, BLOG);
            while (bkey[h] != v) h = (h + 1) & (BT - 1);
            bslot[h] = (uint32_t)x;
        }
        __syncwarp();
    }
    // pass 2: slot words of the records (owner-first order)
    for (int e = 0; e < Lb; ++e) {
        const uint32_t r = __ldg(V.inc + base + (size_t)e * 32 + lane);
        if (r == FF_NOREC) continue;
        uint32_t word = 0;
        if (staged) {
            int c[4];
            owner_first<NLOC>(conn, r, c);
#pragma unroll
            for (int b = 0; b < NLOC; ++b) {
                const uint32_t v = (uint32_t)c[b];
                uint32_t h = ht_hash(v, BLOG);
                while (bkey[h] != v) h = (h + 1) & (BT - 1);
                word |= bslot[h] << (8 * b);
            }
        }
        loc[base + (size_t)e * 32 + lane] = word;
    }
    if (staged)
        for (int x = lane; x < nv; x += 32) blkvert[(size_t)blk * FF_STAGE_MAX + x] = vlist[x];
    if (lane == 0) {
        blkvcnt[blk] = staged ? nv : -1;
        if (staged) atomicMax(maxstage, nv);
        else atomicAdd(maxstage + 1, 1);
    }
}

// row patterns of the 32 rows of a block in ONE pass: sorted distinct columns into a fixed-stride scratch
// (tmpcol[row*cap + x], compacted after the scan of the row lengths), positions of every record's vertices in the row
template <int NLOC>
__global__ void __launch_bounds__(128) k_block_pattern(const int32_t *__restrict__ conn, int nrows, const IncView V, int Lmax, int cap,
                                                       int TSR, int RLOG, int32_t *__restrict__ tmpcol, int32_t *__restrict__ rowlen,
                                                       int32_t *__restrict__ diagnode, uint32_t *__restrict__ pos,
                                                       int32_t *__restrict__ maxrow)
{
    extern __shared__ uint32_t smem_u[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cstride = Lmax * 4 + 4;
    int UL = 64;
    while (UL < cap) UL <<= 1; // room for the shared-memory bitonic sort
    const int RW = (Lmax * 33 + 3) & ~3;
    const size_t per_warp = (size_t)RW + (size_t)32 * cstride + 2 * (size_t)TSR + UL;
    uint32_t *recs = smem_u + (size_t)w * per_warp; // records, later the position words [e*33 + row]
    int32_t *cand = reinterpret_cast<int32_t *>(recs + RW);
    uint32_t *rkey = reinterpret_cast<uint32_t *>(cand + (size_t)32 * cstride);
    uint32_t *rrank = rkey + TSR;
    int32_t *ulist = reinterpret_cast<int32_t *>(rrank + TSR);
    const uint32_t lt = (1u << lane) - 1u, hmask = (uint32_t)TSR - 1u;
    const int blk = blockIdx.x * (blockDim.x >> 5) + w, nblk = (nrows + 31) >> 5;
    if (blk >= nblk) return;
    const uint32_t base = V.blkoff[blk];
    const int Lb = (int)((V.blkoff[blk + 1] - base) >> 5);
    blk_load<NLOC>(conn, V, base, Lb, lane, cstride, recs, cand);
    const int myrow = blk * 32 + lane;
    const int mycnt = myrow < nrows ? V.cnt[myrow] : 0;
    __syncwarp();
    int localmax = 0;
    for (int x = lane; x < TSR; x += 32) rkey[x] = HT_EMPTY; // cleared once; every row removes its own keys when done
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
        const int row = blk * 32 + r;
        if (row >= nrows) break;
        const int cnt = __shfl_sync(0xffffffffu, mycnt, r);
        const int ncand = cnt * 4;
        const int32_t *crow = cand + (size_t)r * cstride;
        if (lane == 0) { // the row's own node is always a column (it is a vertex of every incident element)
            rkey[ht_hash((uint32_t)row, RLOG)] = (uint32_t)row;
            ulist[0] = row;
        }
        __syncwarp();
        int nu = 1;
        for (int x0 = 0; x0 < ncand; x0 += 32) {
            const int x = x0 + lane;
            const int vi = x < ncand ? crow[x] : -1;
            bool isnew = false;
            if (vi >= 0 && vi != row) {
                const uint32_t v = (uint32_t)vi;
                uint32_t h = ht_hash(v, RLOG);
                while (true) {
                    uint32_t k = rkey[h];
                    if (k == v) break;
                    if (k == HT_EMPTY) {
                        k = atomicCAS(&rkey[h], HT_EMPTY, v);
                        if (k == HT_EMPTY) {
                            isnew = true;
                            break;
                        }
                        if (k == v) break;
                    }
                    h = (h + 1) & hmask;
                }
            }
            const unsigned msk = __ballot_sync(0xffffffffu, isnew);
            if (isnew) ulist[nu + __popc(msk & lt)] = vi;
            nu += __popc(msk);
        }
        __syncwarp();
        if (nu <= 32) {
            // rank by counting (all distinct): no sort needed, lane x owns ulist[x]
            const int v = lane < nu ? ulist[lane] : INT_MAX;
            int rank = 0;
            for (int j = 0; j < nu; ++j) rank += (__shfl_sync(0xffffffffu, v, j) < v) ? 1 : 0;
            if (lane < nu) {
                uint32_t h = ht_hash((uint32_t)v, RLOG);
                while (rkey[h] != (uint32_t)v) h = (h + 1) & hmask;
                rrank[h] = (uint32_t)rank;
                tmpcol[(size_t)row * cap + rank] = v;
                if (v == row) diagnode[row] = rank;
            }
        } else {
            int m = 64;
            while (m < nu) m <<= 1;
            for (int x = nu + lane; x < m; x += 32) ulist[x] = INT_MAX;
            __syncwarp();
            warp_sort_smem(ulist, m, lane);
            for (int x = lane; x < nu; x += 32) {
                const int v = ulist[x];
                uint32_t h = ht_hash((uint32_t)v, RLOG);
                while (rkey[h] != (uint32_t)v) h = (h + 1) & hmask;
                rrank[h] = (uint32_t)x;
                tmpcol[(size_t)row * cap + x] = v;
                if (v == row) diagnode[row] = x;
            }
        }
        if (lane == 0) rowlen[row] = nu;
        localmax = max(localmax, nu);
        __syncwarp();
        // position words: lane e takes record e of this row
        for (int e = lane; e < cnt; e += 32) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < NLOC; ++b) {
                const uint32_t v = (uint32_t)crow[e * 4 + b];
                uint32_t h = ht_hash(v, RLOG);
                while (rkey[h] != v) h = (h + 1) & hmask;
                word |= (rrank[h] & 255u) << (8 * b);
            }
            recs[e * 33 + r] = word;
        }
        __syncwarp();
        // remove this row's keys (linear probing without deletions in between: every key is still reachable)
        for (int x = lane; x < nu; x += 32) {
            const uint32_t v = (uint32_t)ulist[x];
            uint32_t h = ht_hash(v, RLOG);
            while (rkey[h] != v) h = (h + 1) & hmask;
            ulist[x] = (int32_t)h; // remember the slot, clear after everybody has found theirs
        }
        __syncwarp();
        for (int x = lane; x < nu; x += 32) rkey[ulist[x]] = HT_EMPTY;
        __syncwarp();
    }
    for (int e = 0; e < mycnt; ++e) pos[base + (size_t)e * 32 + lane] = recs[e * 33 + lane];
    if (lane == 0 && localmax > 0) atomicMax(maxrow, localmax);
}

// ---------------------------------------------------------------------------------------------------------------
// P1, every ELL block staged (the common case): the per-assembly symbolic phase as pure bit arithmetic.
// The slot bytes of the incidence records (Incidence::loc) number the distinct vertices of a block in ascending vertex
// order, so a row's column set is a 256-bit bitmap over the block's slots, its length a population count, and the
// position of a column inside the row the population count of the bits below its slot.  One thread per row, no
// hashing, no sorting, no atomics on the data path; reads 4 B and writes 4 B per incidence record.
// ---------------------------------------------------------------------------------------------------------------
static constexpr int SYM_THREADS = 128;
static constexpr int SYM_WORDS = FF_STAGE_MAX / 32;

template <int NLOC, bool WITH_POS>
__global__ void __launch_bounds__(SYM_THREADS) k_sym_p1_rows(int nrows, int nrows_pad, const int32_t *__restrict__ cnt,
                                                             const uint32_t *__restrict__ blkoff, const uint32_t *__restrict__ loc,
                                                             uint32_t *__restrict__ pos, uint32_t *__restrict__ bitmaps,
                                                             int32_t *__restrict__ rowlen, int32_t *__restrict__ diagnode,
                                                             int32_t *__restrict__ maxrow)
{
    __shared__ uint32_t sbm[SYM_WORDS][SYM_THREADS];  // [word][thread]: conflict-free
    __shared__ uint32_t spre[SYM_WORDS][SYM_THREADS]; // population count of the words below
    const int tid = threadIdx.x, lane = tid & 31;
    const int row = blockIdx.x * SYM_THREADS + tid;
    const int blk = row >> 5, nblk = (nrows + 31) >> 5;
    if (blk >= nblk) return; // whole warps only
    const int mycnt = row < nrows ? cnt[row] : 0;
    const uint32_t base = blkoff[blk];
    const int Lb = (int)((blkoff[blk + 1] - base) >> 5);
    const uint32_t *ploc = loc + base + lane;
#pragma unroll
    for (int w = 0; w < SYM_WORDS; ++w) sbm[w][tid] = 0u;
    for (int e = 0; e < Lb; ++e) {
        const uint32_t lw = __ldg(ploc + (size_t)e * 32);
        if (e < mycnt) {
#pragma unroll
            for (int b = 0; b < NLOC; ++b) {
                const uint32_t sl = (lw >> (8 * b)) & 255u;
                sbm[sl >> 5][tid] |= 1u << (sl & 31u);
            }
        }
    }
    int nu = 0;
#pragma unroll
    for (int w = 0; w < SYM_WORDS; ++w) {
        const uint32_t bits = sbm[w][tid];
        spre[w][tid] = (uint32_t)nu;
        nu += __popc(bits);
        if (bitmaps) bitmaps[(size_t)w * nrows_pad + row] = bits;
    }
    uint32_t *ppos = pos + base + lane;
    int diag = 0;
    // WITH_POS = false: only the position of the diagonal (byte 0 of the first record = the row's own vertex); the
    // per-record positions are produced later, on demand, by the same kernel (ff_pattern_ensure_pos)
    const int epos = WITH_POS ? mycnt : min(mycnt, 1);
    for (int e = 0; e < epos; ++e) {
        const uint32_t lw = __ldg(ploc + (size_t)e * 32);
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < NLOC; ++b) {
            const uint32_t sl = (lw >> (8 * b)) & 255u, w = sl >> 5;
            const uint32_t r = spre[w][tid] + __popc(sbm[w][tid] & ((1u << (sl & 31u)) - 1u));
            word |= r << (8 * b);
        }
        if (e == 0) diag = (int)(word & 255u); // byte 0 = the row's own vertex
        if (WITH_POS) __stcs(ppos + (size_t)e * 32, word);
    }
    if (row < nrows && rowlen) {
        rowlen[row] = nu;
        diagnode[row] = diag;
    }
    if (maxrow) {
        int m = nu;
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0) atomicMax(maxrow, m);
    }
}

// columns of every row from its slot bitmap and the block's (ascending) vertex list; diagonal positions.  One thread
// per row expands its bitmap into the warp's shared-memory stage (the 32 rows of a warp are contiguous in ncol), then
// the warp writes the whole span with coalesced stores.  Rows too long for the stage (span > cap) are written directly.
__global__ void __launch_bounds__(SYM_THREADS) k_sym_p1_cols(int nrows, int nrows_pad, const uint32_t *__restrict__ bitmaps,
                                                             const int32_t *__restrict__ nrowptr, const int32_t *__restrict__ blkvert,
                                                             const int32_t *__restrict__ diagnode, int32_t *__restrict__ ncol,
                                                             int32_t *__restrict__ diagpos, int cap)
{
    extern __shared__ int32_t scol[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x * SYM_THREADS + tid;
    const int r0 = row & ~31;
    if (r0 >= nrows) return; // whole warps only
    const int rlast = min(r0 + 32, nrows);
    const int o0 = nrowptr[r0], span = nrowptr[rlast] - o0;
    const bool staged = span <= cap;
    int32_t *st = scol + (size_t)warp * cap;
    if (row < nrows) {
        const int32_t *bv = blkvert + (size_t)(row >> 5) * FF_STAGE_MAX;
        int o = nrowptr[row];
        diagpos[row] = o + diagnode[row];
#pragma unroll
        for (int w = 0; w < SYM_WORDS; ++w) {
            uint32_t bits = __ldcs(bitmaps + (size_t)w * nrows_pad + row);
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                const int32_t c = __ldg(bv + w * 32 + b);
                if (staged) st[o - o0] = c;
                else ncol[o] = c;
                ++o;
            }
        }
    }
    __syncwarp();
    if (staged)
        for (int j = lane; j < span; j += 32) ncol[o0 + j] = st[j];
}

// The whole symbolic phase of a staged P1 space in ONE kernel (used once the space's pattern size is known, i.e. from
// the second `matrix A = va(Vh,Vh)` on a fespace): per CTA of 128 rows the slot bitmaps are built in shared memory, the
// row lengths are scanned in the block, the block's offset comes from a decoupled look-back over the preceding blocks
// (tickets, (flag, sum) words as in k_scan_lookback), and the columns are expanded from the bitmaps still in shared
// memory and written with coalesced stores: no bitmaps through HBM, no separate scan, one launch.
//   st[0..ntiles): status words, st[ntiles]: total, st[ntiles+1]: ticket counter; all zero on entry.
template <int NLOC>
__global__ void __launch_bounds__(SYM_THREADS) k_sym_p1_fused(int nrows, const int32_t *__restrict__ cnt,
                                                              const uint32_t *__restrict__ blkoff, const uint32_t *__restrict__ loc,
                                                              const int32_t *__restrict__ blkvert, int32_t *__restrict__ nrowptr,
                                                              int32_t *__restrict__ ncol, int32_t *__restrict__ diagpos,
                                                              int32_t *__restrict__ maxrow, unsigned long long *__restrict__ st,
                                                              int ntiles, int cap)
{
    extern __shared__ int32_t scol[];
    // [word][thread]: conflict-free.  Measured alternatives, both slower (symbolic phase 0.26 ms with this layout): bitmap
    // words in registers selected by comparisons (0.41 ms), one array per vertex of the record to break the chain of
    // dependent read-modify-writes (0.39 ms).  The bits are set with shared-memory atomics whose result is not used (ATOMS.OR RZ):
    // one LSU operation per bit instead of a load / or / store chain (symbolic phase 0.263 -> 0.253 ms).
    __shared__ uint32_t sbm[SYM_WORDS][SYM_THREADS];
    __shared__ int s_tile, s_wsum[SYM_THREADS / 32];
    __shared__ long long s_prefix;
    constexpr unsigned long long MASK = (1ull << 62) - 1ull;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int)atomicAdd(st + ntiles + 1, 1ull);
    __syncthreads();
    const int tile = s_tile;
    const int row = tile * SYM_THREADS + tid;
    const int blk = row >> 5, nblk = (nrows + 31) >> 5;
    int nu = 0, diag = 0;
#pragma unroll
    for (int w = 0; w < SYM_WORDS; ++w) sbm[w][tid] = 0u;
    // the block's ascending vertex list (the columns) is requested now and used at the very end: the expansion of the
    // bitmaps into columns was a chain of dependent global loads (a third of the stall samples in the r02 capture)
    __shared__ int32_t sbv[SYM_THREADS / 32][FF_STAGE_MAX];
    if (blk < nblk) {
        const int32_t *bvg = blkvert + (size_t)blk * FF_STAGE_MAX;
#pragma unroll
        for (int j = 0; j < FF_STAGE_MAX / 32; ++j) sbv[warp][j * 32 + lane] = __ldg(bvg + j * 32 + lane);
    }
    if (blk < nblk) {
        const int mycnt = row < nrows ? cnt[row] : 0;
        const uint32_t base = blkoff[blk];
        const int Lb = (int)((blkoff[blk + 1] - base) >> 5);
        const uint32_t *ploc = loc + base + lane;
        uint32_t own = 0;
        // the slot words of 8 records are requested together (the loop is bound by the latency of these loads: 18.8 stall
        // cycles per issued instruction on the long scoreboard in the r02a capture with one load in flight per thread)
        constexpr int UNR = 8;
        for (int e0 = 0; e0 < Lb; e0 += UNR) {
            uint32_t lw[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) lw[u] = e0 + u < Lb ? __ldg(ploc + (size_t)(e0 + u) * 32) : 0u;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int e = e0 + u;
                if (e < mycnt) {
                    if (e == 0) { // byte 0 of a record = the row's own vertex: the same in every record, marked once
                        own = lw[u] & 255u;
                        atomicOr(&sbm[own >> 5][tid], 1u << (own & 31u));
                    }
#pragma unroll
                    for (int b = 1; b < NLOC; ++b) {
                        const uint32_t sl = (lw[u] >> (8 * b)) & 255u;
                        atomicOr(&sbm[sl >> 5][tid], 1u << (sl & 31u)); // result unused: fire-and-forget, no dependent read-modify-write chain
                    }
                }
            }
        }
#pragma unroll
        for (int w = 0; w < SYM_WORDS; ++w) {
            const uint32_t bits = sbm[w][tid];
            if (mycnt > 0 && w < (int)(own >> 5)) diag += __popc(bits);
            if (mycnt > 0 && w == (int)(own >> 5)) diag += __popc(bits & ((1u << (own & 31u)) - 1u));
            nu += __popc(bits);
        }
    }
    {
        int m = nu;
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0) atomicMax(maxrow, m);
    }
    // block scan of the row lengths
    int inc = nu;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SYM_THREADS / 32; ++w) {
        if (w < warp) wbase += s_wsum[w];
        tot += s_wsum[w];
    }
    // look-back
    if (tid == 0 && tile > 0) atomicExch(st + tile, (1ull << 62) | (unsigned long long)tot);
    if (tid < 32) {
        long long prefix = 0;
        for (int j = tile - 1; tile > 0;) {
            const int idx = j - lane;
            unsigned long long w = 2ull << 62;
            if (idx >= 0) {
                do {
                    w = *reinterpret_cast<volatile unsigned long long *>(st + idx);
                } while ((w >> 62) == 0);
            }
            const unsigned incl = __ballot_sync(0xffffffffu, (w >> 62) == 2);
            const int first = incl ? __ffs(incl) - 1 : 31;
            long long val = lane <= first ? (long long)(w & MASK) : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            prefix += val;
            if (incl) break;
            j -= 32;
        }
        if (lane == 0) {
            __threadfence();
            atomicExch(st + tile, (2ull << 62) | (unsigned long long)(prefix + tot));
            if (tile == ntiles - 1) {
                st[ntiles] = (unsigned long long)(prefix + tot);
                nrowptr[nrows] = (int32_t)(prefix + tot);
            }
            s_prefix = prefix;
        }
    }
    __syncthreads();
    const int o = (int)s_prefix + wbase + inc - nu; // first entry of this row
    if (row < nrows) {
        nrowptr[row] = o;
        diagpos[row] = o + diag;
    }
    // columns: expand the bitmaps into the warp's stage, then coalesced stores of the warp's span
    const int o0 = __shfl_sync(0xffffffffu, o, 0), span = __shfl_sync(0xffffffffu, o + nu, 31) - o0;
    const bool staged = span <= cap;
    int32_t *stg = scol + (size_t)warp * cap;
    if (row < nrows && nu > 0) {
        const int32_t *bv = sbv[warp];
        int oo = o;
#pragma unroll
        for (int w = 0; w < SYM_WORDS; ++w) {
            uint32_t bits = sbm[w][tid];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1u;
                const int32_t c = bv[w * 32 + b];
                if (staged) stg[oo - o0] = c;
                else ncol[oo] = c;
                ++oo;
            }
        }
    }
    __syncwarp();
    if (staged)
        for (int j = lane; j < span; j += 32) ncol[o0 + j] = stg[j];
}

// tmpcol (fixed stride) -> ncol (CSR): one warp per 32 rows, a row segment at a time
__global__ void k_compact_cols(const int32_t *__restrict__ tmpcol, int cap, const int32_t *__restrict__ nrowptr, int nrows,
                               int32_t *__restrict__ ncol)
{
    const int blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int row0 = blk * 32;
    if (row0 >= nrows) return;
    const int myrow = row0 + lane;
    const int rb = myrow <= nrows ? nrowptr[min(myrow, nrows)] : 0;
    const int rbn = myrow < nrows ? nrowptr[myrow + 1] : rb;
    for (int r = 0; r < 32 && row0 + r < nrows; ++r) {
        const int b = __shfl_sync(0xffffffffu, rb, r), L = __shfl_sync(0xffffffffu, rbn, r) - b;
        for (int x = lane; x < L; x += 32) ncol[(size_t)b + x] = tmpcol[(size_t)(row0 + r) * cap + x];
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
    d_max.alloc(2);
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, 2 * sizeof(int32_t), st));
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
    int32_t h_max[2] = {0, 0};
    FF_CUDA(cudaMemcpyAsync(h_max, d_max.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    I.maxinc = h_max[0];
    I.nempty = h_max[1];
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
    if (I.ell) {
        // vertex staging tables for the thread-per-row numeric kernels
        const int nblk = (nrows + 31) / 32;
        const int warps = 8;
        I.loc.alloc((size_t)nrec);
        I.blkvert.alloc((size_t)nblk * FF_STAGE_MAX);
        I.blkvcnt.alloc((size_t)nblk);
        DBuf<int32_t> d_st;
        d_st.alloc(2);
        FF_CUDA(cudaMemsetAsync(d_st.p, 0, 2 * sizeof(int32_t), st));
        const size_t shmem = (size_t)warps * STAGE_WORDS * 4;
        const int blocks = ff_blocks((size_t)nblk, warps);
        if (nloc == 4) {
            FF_CUDA(cudaFuncSetAttribute(k_block_stage<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
            ff_launch(ctx, "inc_stage", [&] {
                k_block_stage<4><<<blocks, warps * 32, shmem, st>>>(s->e2n, nrows, V, I.loc.p, I.blkvert.p, I.blkvcnt.p, d_st.p);
            });
        } else {
            FF_CUDA(cudaFuncSetAttribute(k_block_stage<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
            ff_launch(ctx, "inc_stage", [&] {
                k_block_stage<3><<<blocks, warps * 32, shmem, st>>>(s->e2n, nrows, V, I.loc.p, I.blkvert.p, I.blkvcnt.p, d_st.p);
            });
        }
        int32_t h_st[2] = {0, 0};
        FF_CUDA(cudaMemcpyAsync(h_st, d_st.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FF_CUDA(cudaStreamSynchronize(st));
        I.maxstage = h_st[0];
        I.nunstaged = h_st[1];
    }
    I.built = true;
}

// ---------------------------------------------------------------------------------------------------------------
// stage 2: row pattern.  One warp per node row.
// ---------------------------------------------------------------------------------------------------------------
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
            // positions of the nodes of every incident element inside this row (not asked for by rectangular patterns)
            for (int x = lane; pos != nullptr && x < ncand; x += 32) {
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
    P->dist = s->mesh && s->mesh->distributed;
    P->ctx = ctx;
    P->ref.set(ctx);
    P->nrows_node = nrows;
    P->ncols_node = s->nnodes;
    P->ncomp = nc;
    P->n = nrows * nc;
    P->nlocp = (s->order == 1) ? 4 : nloc;

    DBuf<int32_t> rowlen, diagnode, d_max;
    rowlen.alloc((size_t)nrows + 1);
    diagnode.alloc((size_t)nrows);
    d_max.alloc(1);
    FF_CUDA(cudaMemsetAsync(rowlen.p + nrows, 0, sizeof(int32_t), st));
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, sizeof(int32_t), st));
    P->nrowptr.alloc((size_t)nrows + 1);
    int64_t nnzn = 0;
    int32_t h_max = 0;
    bool diag_done = false;
    const bool lazy_pos_space = s->order == 1 && I.nunstaged == 0 && I.nempty == 0 && nc == 1 && s->tiles.state == 1 && ctx->tile_policy != 0;
    if (lazy_pos_space && s->sym_nnz_node > 0) {
        // --- P1 scalar space seen before (row tiles built, pattern size known): the fused single-kernel symbolic phase
        const int ntl = ff_blocks((size_t)((nrows + 31) / 32) * 32, SYM_THREADS);
        DBuf<unsigned long long> stw;
        stw.alloc((size_t)ntl + 2);
        FF_CUDA(cudaMemsetAsync(stw.p, 0, stw.bytes(), st));
        P->ncol.alloc((size_t)s->sym_nnz_node);
        P->diagpos.alloc((size_t)P->n);
        const int ccap = std::min(32 * std::max(1, s->sym_maxrow), 1536);
        const size_t shm = (size_t)(SYM_THREADS / 32) * ccap * 4;
        ff_launch(ctx, "sym_p1_fused", [&] {
            if (nloc == 4)
                k_sym_p1_fused<4><<<ntl, SYM_THREADS, shm, st>>>(nrows, I.cnt.p, I.blkoff.p, I.loc.p, I.blkvert.p, P->nrowptr.p, P->ncol.p,
                                                                  P->diagpos.p, d_max.p, stw.p, ntl, ccap);
            else
                k_sym_p1_fused<3><<<ntl, SYM_THREADS, shm, st>>>(nrows, I.cnt.p, I.blkoff.p, I.loc.p, I.blkvert.p, P->nrowptr.p, P->ncol.p,
                                                                  P->diagpos.p, d_max.p, stw.p, ntl, ccap);
        });
        unsigned long long tot = 0;
        FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FF_CUDA(cudaMemcpyAsync(&tot, stw.p + ntl, sizeof(tot), cudaMemcpyDeviceToHost, st));
        FF_CUDA(cudaStreamSynchronize(st));
        nnzn = (int64_t)tot;
        FF_REQUIRE(nnzn == s->sym_nnz_node, "internal: the pattern of a fespace changed size between two symbolic phases");
        P->maxrow_node = h_max;
        P->nnz_node = nnzn;
        P->nnz = nnzn;
        diag_done = true;
    } else if (s->order == 1 && I.nunstaged == 0 && I.nempty == 0) {
        // --- P1, all blocks staged: slot bitmaps (k_sym_p1_rows) -> scan -> columns (k_sym_p1_cols)
        const int nblk = (nrows + 31) / 32, nrows_pad = nblk * 32;
        DBuf<uint32_t> bitmaps;
        bitmaps.alloc((size_t)SYM_WORDS * nrows_pad);
        // The per-record positions (201 MB written at cube(128)) only serve the thread-per-row numeric kernels.  On a
        // scalar space whose row tiles are built they are produced on demand instead (ff_pattern_ensure_pos).
        const bool lazy_pos = nc == 1 && s->tiles.state == 1 && ctx->tile_policy != 0;
        uint32_t *posw = nullptr;
        if (!lazy_pos) {
            P->pos8.alloc((size_t)I.nrec * P->nlocp);
            posw = reinterpret_cast<uint32_t *>(P->pos8.p);
        }
        const int blocks = ff_blocks((size_t)nrows_pad, SYM_THREADS);
        ff_launch(ctx, "sym_p1_rows", [&] {
            if (lazy_pos) {
                if (nloc == 4)
                    k_sym_p1_rows<4, false><<<blocks, SYM_THREADS, 0, st>>>(nrows, nrows_pad, I.cnt.p, I.blkoff.p, I.loc.p, nullptr, bitmaps.p,
                                                                            rowlen.p, diagnode.p, d_max.p);
                else
                    k_sym_p1_rows<3, false><<<blocks, SYM_THREADS, 0, st>>>(nrows, nrows_pad, I.cnt.p, I.blkoff.p, I.loc.p, nullptr, bitmaps.p,
                                                                            rowlen.p, diagnode.p, d_max.p);
            } else if (nloc == 4)
                k_sym_p1_rows<4, true><<<blocks, SYM_THREADS, 0, st>>>(nrows, nrows_pad, I.cnt.p, I.blkoff.p, I.loc.p, posw, bitmaps.p, rowlen.p,
                                                                       diagnode.p, d_max.p);
            else
                k_sym_p1_rows<3, true><<<blocks, SYM_THREADS, 0, st>>>(nrows, nrows_pad, I.cnt.p, I.blkoff.p, I.loc.p, posw, bitmaps.p, rowlen.p,
                                                                       diagnode.p, d_max.p);
        });
        FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        ff_exclusive_scan_i32(ctx, rowlen.p, P->nrowptr.p, (size_t)nrows + 1, &nnzn); // synchronises the stream
        P->maxrow_node = h_max;
        FF_REQUIRE(P->maxrow_node <= 255, "a P1 node with more than 254 neighbours is not supported");
        P->nnz_node = nnzn;
        P->nnz = nnzn * nc * nc;
        FF_REQUIRE(P->nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros (int32 CSR, like MatriceMorse)");
        P->ncol.alloc((size_t)nnzn);
        P->diagpos.alloc((size_t)P->n);
        // scalar spaces: diagpos is final; vector spaces overwrite it in k_expand_rowptr below
        // stage of the column kernel: the 32 rows of a warp hold at most 32 * maxrow entries; capped so that 8 CTAs fit
        const int ccap = std::min(32 * std::max(1, P->maxrow_node), 1536);
        ff_launch(ctx, "sym_p1_cols", [&] {
            k_sym_p1_cols<<<blocks, SYM_THREADS, (size_t)(SYM_THREADS / 32) * ccap * 4, st>>>(nrows, nrows_pad, bitmaps.p, P->nrowptr.p,
                                                                                              I.blkvert.p, diagnode.p, P->ncol.p, P->diagpos.p,
                                                                                              ccap);
        });
        diag_done = (nc == 1);
        if (nc == 1) { // remembered for the fused kernel of the later symbolic phases on this space
            s->sym_nnz_node = nnzn;
            s->sym_maxrow = P->maxrow_node;
        }
    } else if (s->order == 1) {
        // --- P1, general: one pass, one warp per block of 32 rows (k_block_pattern), columns through a fixed-stride scratch
        const int Lmax = I.maxinc, cstride = Lmax * 4 + 4;
        const int cap = Lmax * (nloc - 1) + 1; // a row has at most this many distinct columns
        int TSR = 64, RLOG = 6, UL = 64;
        while (TSR < 2 * cap) {
            TSR <<= 1;
            ++RLOG;
        }
        while (UL < cap) UL <<= 1;
        const size_t per_warp = ((size_t)((Lmax * 33 + 3) & ~3) + (size_t)32 * cstride + 2 * (size_t)TSR + UL) * 4;
        FF_REQUIRE(per_warp <= 200 * 1024, "a vertex has too many incident elements for the block pattern kernel");
        int warps = 4;
        while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
        const size_t shmem = warps * per_warp;
        const int nblk = (nrows + 31) / 32, blocks = ff_blocks((size_t)nblk, warps);
        DBuf<int32_t> tmpcol;
        tmpcol.alloc((size_t)nrows * cap);
        P->pos8.alloc((size_t)I.nrec * P->nlocp);
        uint32_t *posw = reinterpret_cast<uint32_t *>(P->pos8.p);
        if (nloc == 4) {
            FF_CUDA(cudaFuncSetAttribute(k_block_pattern<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
            ff_launch(ctx, "sym_block_pattern", [&] {
                k_block_pattern<4><<<blocks, warps * 32, shmem, st>>>(s->e2n, nrows, V, Lmax, cap, TSR, RLOG, tmpcol.p, rowlen.p, diagnode.p,
                                                                       posw, d_max.p);
            });
        } else {
            FF_CUDA(cudaFuncSetAttribute(k_block_pattern<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
            ff_launch(ctx, "sym_block_pattern", [&] {
                k_block_pattern<3><<<blocks, warps * 32, shmem, st>>>(s->e2n, nrows, V, Lmax, cap, TSR, RLOG, tmpcol.p, rowlen.p, diagnode.p,
                                                                       posw, d_max.p);
            });
        }
        FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        ff_exclusive_scan_i32(ctx, rowlen.p, P->nrowptr.p, (size_t)nrows + 1, &nnzn); // synchronises the stream
        P->maxrow_node = h_max;
        FF_REQUIRE(P->maxrow_node <= 255, "a P1 node with more than 254 neighbours is not supported");
        P->nnz_node = nnzn;
        P->nnz = nnzn * nc * nc;
        FF_REQUIRE(P->nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros (int32 CSR, like MatriceMorse)");
        P->ncol.alloc((size_t)nnzn);
        ff_launch(ctx, "sym_compact_cols", [&] {
            k_compact_cols<<<ff_blocks((size_t)nblk * 32, 256), 256, 0, st>>>(tmpcol.p, cap, P->nrowptr.p, nrows, P->ncol.p);
        });
    } else {
        // --- P2: one warp per node row, two passes (row lengths, then columns + positions)
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
        run_row_pattern<uint8_t>(ctx, P, s->e2n, nloc, TS, LOG, warps, shmem, blocks, V, rowlen.p, diagnode.p, d_max.p, nullptr, 0);
        FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        ff_exclusive_scan_i32(ctx, rowlen.p, P->nrowptr.p, (size_t)nrows + 1, &nnzn); // synchronises the stream
        P->maxrow_node = h_max;
        P->nnz_node = nnzn;
        P->nnz = nnzn * nc * nc;
        FF_REQUIRE(P->nnz < ((int64_t)1 << 31), "matrix exceeds 2^31 nonzeros (int32 CSR, like MatriceMorse)");
        P->ncol.alloc((size_t)nnzn);
        if (P->maxrow_node <= 255) {
            P->pos8.alloc((size_t)I.nrec * P->nlocp);
            run_row_pattern<uint8_t>(ctx, P, s->e2n, nloc, TS, LOG, warps, shmem, blocks, V, nullptr, diagnode.p, nullptr, P->pos8.p, 1);
        } else {
            P->pos16.alloc((size_t)I.nrec * P->nlocp);
            run_row_pattern<uint16_t>(ctx, P, s->e2n, nloc, TS, LOG, warps, shmem, blocks, V, nullptr, diagnode.p, nullptr, P->pos16.p, 1);
        }
    }
    rowlen.release();

    // --- dof-level CSR
    if (!P->diagpos.p) P->diagpos.alloc((size_t)P->n);
    if (nc == 1) {
        P->rowptr = P->nrowptr.p;
        P->colind = P->ncol.p;
        if (!diag_done) ff_launch(ctx, "sym_diagpos", [&] { k_diagpos_scalar<<<ff_blocks(nrows, 256), 256, 0, st>>>(P->nrowptr.p, diagnode.p, nrows, P->diagpos.p); });
    } else {
        P->rowptr_own.alloc((size_t)P->n + 1);
        ff_launch(ctx, "sym_expand_rowptr", [&] {
            k_expand_rowptr<<<ff_blocks((size_t)nrows + 1, 256), 256, 0, st>>>(P->nrowptr.p, nrows, nc, P->rowptr_own.p, diagnode.p, P->diagpos.p);
        });
        P->rowptr = P->rowptr_own.p;
        // the dof-level column indices (4 bytes per entry: 2.2 GB at config 3) are only materialised when somebody reads
        // them (download to the host, a kernel that is not the node-block SpMV): ff_pattern_ensure_colind
        P->colind = nullptr;
    }
    // no synchronisation here: everything downstream is ordered on the same stream (the temporaries above are
    // released through the stream-ordered allocator)
    *out = P;
    P = nullptr;
    FF_API_END((delete P, s ? s->ctx : nullptr))
}

// Node-level pattern of a RECTANGULAR matrix (`matrix B = vb(Uh,Vh)`, assemble.cu: ffcuda_assemble_bilinear_rect): row
// node i of the test space sv is coupled with every node of the space of the unknown su carried by an element around i.
// The two passes of k_row_pattern with the incidence lists of sv and the element -> node table of su.
void ff_rect_node_pattern(ffcuda_space *sv, ffcuda_space *su, DBuf<int32_t> &nrowptr, DBuf<int32_t> &ncol, int64_t *nnz_node,
                          int *maxrow_node)
{
    ffcuda_ctx *ctx = sv->ctx;
    cudaStream_t st = ctx->stream;
    ff_build_incidence(sv);
    const Incidence &I = sv->incidence;
    const IncView V = ff_view(I);
    const int nrows = sv->nnodes_owned, nloc = su->nloc;
    DBuf<int32_t> rowlen, diagnode, d_max;
    rowlen.alloc((size_t)nrows + 1);
    diagnode.alloc((size_t)nrows); // written where a column node has the number of the row node; not used
    d_max.alloc(1);
    FF_CUDA(cudaMemsetAsync(rowlen.p, 0, rowlen.bytes(), st));
    FF_CUDA(cudaMemsetAsync(d_max.p, 0, sizeof(int32_t), st));
    nrowptr.alloc((size_t)nrows + 1);
    int TS = 64, LOG = 6;
    while (TS < 2 * std::max(I.maxinc, 1) * nloc) {
        TS <<= 1;
        ++LOG;
    }
    const size_t per_warp = ((size_t)TS + TS / 2) * 4;
    FF_REQUIRE(per_warp <= 200 * 1024, "a node has too many incident elements for the shared-memory hash table");
    int warps = 8;
    while (warps > 1 && (size_t)warps * per_warp > 96 * 1024) warps >>= 1;
    const size_t shmem = (size_t)warps * per_warp;
    const int blocks = std::max(1, min(ff_blocks((size_t)nrows, warps), ctx->sm_count * 16));
    FF_CUDA(cudaFuncSetAttribute(k_row_pattern<0, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    ff_launch(ctx, "sym_rect_count", [&] {
        k_row_pattern<0, uint8_t><<<blocks, warps * 32, shmem, st>>>(su->e2n, nloc, nloc, nrows, TS, LOG, V, rowlen.p, nullptr, nullptr, nullptr,
                                                                     nullptr, d_max.p);
    });
    int32_t h_max = 0;
    int64_t nnzn = 0;
    FF_CUDA(cudaMemcpyAsync(&h_max, d_max.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ff_exclusive_scan_i32(ctx, rowlen.p, nrowptr.p, (size_t)nrows + 1, &nnzn); // synchronises the stream
    ncol.alloc((size_t)std::max<int64_t>(nnzn, 1));
    FF_CUDA(cudaFuncSetAttribute(k_row_pattern<1, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    ff_launch(ctx, "sym_rect_fill", [&] {
        k_row_pattern<1, uint8_t><<<blocks, warps * 32, shmem, st>>>(su->e2n, nloc, nloc, nrows, TS, LOG, V, nullptr, nrowptr.p, ncol.p, nullptr,
                                                                     diagnode.p, nullptr);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // rowlen / diagnode / d_max go out of scope
    *nnz_node = nnzn;
    *maxrow_node = h_max;
}

// dof-level column indices of a vector-space pattern, expanded from the node-level ones on first use
void ff_pattern_ensure_colind(ffcuda_pattern *P)
{
    if (P->colind) return;
    ffcuda_ctx *ctx = P->ctx;
    P->colind_own.alloc((size_t)P->nnz);
    ff_launch(ctx, "sym_expand_colind", [&] {
        k_expand_colind<<<ff_blocks((size_t)P->nrows_node * 32, 256), 256, 0, ctx->stream>>>(P->nrowptr.p, P->ncol.p, P->nrows_node,
                                                                                             P->ncomp, P->colind_own.p);
    });
    P->colind = P->colind_own.p;
}

// per-record positions of a P1 pattern whose symbolic phase skipped them (see lazy_pos above)
void ff_pattern_ensure_pos(ffcuda_pattern *P)
{
    if (P->pos8.p || P->pos16.p) return;
    ffcuda_space *s = P->space;
    ffcuda_ctx *ctx = P->ctx;
    const Incidence &I = s->incidence;
    FF_REQUIRE(s->order == 1 && I.built && I.ell && I.nunstaged == 0, "internal: pattern without positions");
    const int nrows = P->nrows_node, nblk = (nrows + 31) / 32, nrows_pad = nblk * 32;
    P->pos8.alloc((size_t)I.nrec * P->nlocp);
    uint32_t *posw = reinterpret_cast<uint32_t *>(P->pos8.p);
    const int blocks = ff_blocks((size_t)nrows_pad, SYM_THREADS);
    cudaStream_t st = ctx->stream;
    ff_launch(ctx, "sym_p1_positions", [&] {
        if (s->nloc == 4)
            k_sym_p1_rows<4, true><<<blocks, SYM_THREADS, 0, st>>>(nrows, nrows_pad, I.cnt.p, I.blkoff.p, I.loc.p, posw, nullptr, nullptr, nullptr,
                                                                   nullptr);
        else
            k_sym_p1_rows<3, true><<<blocks, SYM_THREADS, 0, st>>>(nrows, nrows_pad, I.cnt.p, I.blkoff.p, I.loc.p, posw, nullptr, nullptr, nullptr,
                                                                   nullptr);
    });
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
    if (colind) ff_pattern_ensure_colind(p);
    if (rowptr) FF_CUDA(cudaMemcpyAsync(rowptr, p->rowptr, ((size_t)p->n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (colind) FF_CUDA(cudaMemcpyAsync(colind, p->colind, (size_t)p->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    FF_API_END(p ? p->ctx : nullptr)
}

// The copies run on a second stream behind an event of the context's stream: the numeric assembly that follows the
// symbolic phase overlaps the transfer of the pattern.  Complete after ffcuda_ctx_sync.
extern "C" int ffcuda_pattern_download_async(ffcuda_pattern *p, int32_t *rowptr, int32_t *colind)
{
    FF_API_BEGIN
    FF_REQUIRE(p, "null pattern");
    ffcuda_ctx *ctx = p->ctx;
    ff_enter(ctx);
    if (!ctx->copy_stream) {
        FF_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        FF_CUDA(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
    }
    if (colind) ff_pattern_ensure_colind(p);
    FF_CUDA(cudaEventRecord(ctx->copy_event, ctx->stream));
    FF_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0));
    if (rowptr)
        FF_CUDA(cudaMemcpyAsync(rowptr, p->rowptr, ((size_t)p->n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (colind) FF_CUDA(cudaMemcpyAsync(colind, p->colind, (size_t)p->nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    p->copy_pending = true;
    FF_API_END(p ? p->ctx : nullptr)
}

extern "C" void ffcuda_pattern_destroy(ffcuda_pattern *p)
{
    if (!p) return;
    if (p->copy_pending && p->ctx && p->ctx->copy_stream) cudaStreamSynchronize(p->ctx->copy_stream);
    ff_enter(p->ctx);
    delete p;
}
