// tiles.cu — numeric assembly of scalar P1 forms (c grad u . grad v + m u v) and their right-hand sides by ROW TILES.
//
// Replaces, for the headline configurations, the element loop of AssembleBilinearForm (fflib/problem.cpp:1096-1103,
// :1398-1405) + Element_Op (:6063-6160, :6337-6437) + HashMatrix::operator+=(MatriceElementaire&)
// (femlib/HashMatrix.cpp:1295-1332), and AssembleLinearForm / Element_rhs (:10878-11227, :7917-7985).  Same mathematics
// as the thread-per-row kernels of assemble.cu; different work decomposition:
//
//   * the rows of the matrix are grouped in TILES: compact clusters of <= TR vertices, consecutive in the Morton order
//     of the vertex coordinates, the axes scaled by the mean edge extent (the clustering is internal: the CSR that is
//     produced is FreeFEM's, row by row);
//   * one CTA per tile (persistent).  Every element touching a row of the tile is evaluated ONCE per tile into shared
//     memory — the thread-per-row kernel evaluates every element once per vertex (4x on tetrahedra);
//   * every matrix entry (i, j) of the tile's rows is OWNED by one thread, which sums the contributions of the elements
//     around the edge ij in a register, in a fixed order, and stores the entry: no atomics, no read-modify-write,
//     bit-reproducible;
//   * the diagonal follows from the partition of unity: K_ii = - sum_j K_ij (+ the mass part).
//
// Two generations of kernels live here:
//   round 1  k_tile_build / k_asm_tiles / k_rhs_tiles: element by element, one descriptor blob per tile (2-D spaces, and
//            3-D with tile_fans = 0 or tile_policy = 2);
//   round 2  k_fan_build / k_asm_fans / k_rhs_fans (3-D): the elements of a tile in FANS around their longest edge — one
//            new vertex per element, shared face normals, contributions pre-summed along the fan — entries sorted by list
//            length with transposed code lists, descriptor in parts with their own bulk copies.  See the comment above
//            k_fan_build and DESIGN.md section 3.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr int TB_THREADS = 256;   // build kernel
constexpr int SORT_CAP = 8192;    // incidence records of a tile's rows / vertex candidates (sort buffer)
constexpr int NE_CAP = 2048;      // elements per tile
constexpr int NV_CAP = 256;       // distinct vertices per tile (8-bit slots)
constexpr int NQ_CAP = 4080;      // matrix entries per tile (2*(NQ_CAP+1) ints share the sort buffer)
constexpr int NC_CAP = 16384;     // contribution codes per tile
constexpr int TR_CAP = 256;       // rows per tile
constexpr int BMW = NV_CAP / 32;  // bitmap words per row
constexpr int HDR = 16;           // header words
// header: 0 nr, 1 nvt, 2 nelem, 3 nq, 4 ncodes (16-bit units, lists padded to even), 5 o_gbase, 6 o_rinfo, 7 o_coord,
//         8 o_telem, 9 o_einfo, 10 o_codes, 11 words
// sections: gbase[nr]   int32  index into vals of the first entry of the row (the pattern's row pointer)
//           rinfo[nr+1] u32    first entry of the row | position of the diagonal << 16 | row length << 24
//           coord[nvt]  DIM doubles per slot
//           telem[nelem] u32   4 slot bytes, element-local vertex order
//           einfo[nq]   u32    offset of the entry's list in code PAIRS | list length << 15 | local row << 20
//           codes       u16    pair * nes + element: index into the numeric kernel's value table (nes: TileSet::nes)

// words of the build kernels' common scratch (BuildScratch)
__host__ __device__ constexpr int build_words()
{
    return SORT_CAP + SORT_CAP + NE_CAP + NE_CAP + NV_CAP + TR_CAP * BMW + (TR_CAP + 1) + (NQ_CAP + 1) + (TB_THREADS + 1) + TR_CAP + NV_CAP / 4 + 16;
}
__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }
#define DIM_PAIRS(d) ((d) * ((d) + 1) / 2)

// ---------------------------------------------------------------------------------------------------------------
// Morton keys
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ord64(double d)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static double unord64(unsigned long long u)
{
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

__global__ void k_bbox(const double *__restrict__ xyz, int vstride, int dim, int n, unsigned long long *__restrict__ box)
{
    unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int x = 0; x < dim; ++x) {
            const unsigned long long o = ord64(xyz[(size_t)i * vstride + x]);
            lo[x] = min(lo[x], o);
            hi[x] = max(hi[x], o);
        }
    for (int x = 0; x < dim; ++x) {
        for (int o = 16; o; o >>= 1) {
            lo[x] = min(lo[x], __shfl_xor_sync(0xffffffffu, lo[x], o));
            hi[x] = max(hi[x], __shfl_xor_sync(0xffffffffu, hi[x], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(box + x, lo[x]);
            atomicMax(box + 3 + x, hi[x]);
        }
    }
}

__device__ __forceinline__ uint32_t spread3(uint32_t v) // 10 bits -> every third bit
{
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t spread2(uint32_t v) // 15 bits -> every second bit
{
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// mean extent of the element edges along every axis (the metric of the mesh): the Morton quantisation is scaled by it so
// that tiles are compact in INDEX space - a cube(128,128,256) slab has vertices twice as dense along z
template <int NV>
__global__ void k_edge_extents(const double *__restrict__ xyz, int vstride, const int32_t *__restrict__ conn, int nt, double fix,
                               unsigned long long *__restrict__ acc)
{
    constexpr int DIM = NV - 1;
    double s[3] = {0.0, 0.0, 0.0};
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nt; k += gridDim.x * blockDim.x) {
        double X[NV][3];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int c = 0; c < DIM; ++c) X[a][c] = xyz[(size_t)conn[(size_t)k * NV + a] * vstride + c];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int b = a + 1; b < NV; ++b)
#pragma unroll
                for (int c = 0; c < DIM; ++c) s[c] += fabs(X[a][c] - X[b][c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
        // fixed point, integer atomics: the sum does not depend on the order of arrival (the tiles, hence the summation
        // order of the assembly, are the same in every run)
        if ((threadIdx.x & 31) == 0) atomicAdd(acc + c, (unsigned long long)llrint(s[c] * fix));
    }
}

struct BoxScale {
    double lo[3], sc[3];
};

__global__ void k_morton(const double *__restrict__ xyz, int vstride, int dim, int n, const BoxScale B, uint32_t *__restrict__ key,
                         int32_t *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t qmax = dim == 3 ? 1023u : 32767u;
    uint32_t q[3] = {0, 0, 0};
    for (int x = 0; x < dim; ++x) {
        const double t = (xyz[(size_t)i * vstride + x] - B.lo[x]) * B.sc[x] + 0.5;
        q[x] = (uint32_t)min((double)qmax, max(0.0, t));
    }
    key[i] = dim == 3 ? (spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2)) : (spread2(q[0]) | (spread2(q[1]) << 1));
    val[i] = i;
}

// ---------------------------------------------------------------------------------------------------------------
// block-wide helpers of the build kernel (TB_THREADS threads, shared-memory arrays)
// ---------------------------------------------------------------------------------------------------------------
__device__ void blk_sort(uint32_t *a, int m) // bitonic, m a power of two
{
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int x = threadIdx.x; x < m; x += TB_THREADS) {
                const int y = x ^ j;
                if (y > x) {
                    const uint32_t u = a[x], v = a[y];
                    if ((u > v) == ((x & k) == 0)) {
                        a[x] = v;
                        a[y] = u;
                    }
                }
            }
            __syncthreads();
        }
}

// in-place exclusive scan of a[0..n), returns the total; part: TB_THREADS+1 ints
__device__ int blk_scan(int *a, int n, int *part)
{
    const int per = (n + TB_THREADS - 1) / TB_THREADS;
    const int b = min(n, (int)threadIdx.x * per), e = min(n, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += a[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) { // one warp scans the TB_THREADS partial sums
        int carry = 0;
        for (int c = 0; c < TB_THREADS; c += 32) {
            const int v = part[c + threadIdx.x];
            int inc = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += u;
            }
            part[c + threadIdx.x] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) part[TB_THREADS] = carry;
    }
    __syncthreads();
    int r = part[threadIdx.x];
    for (int i = b; i < e; ++i) {
        const int v = a[i];
        a[i] = r;
        r += v;
    }
    __syncthreads();
    return part[TB_THREADS];
}

__device__ __forceinline__ int bsearch_u32(const uint32_t *a, int n, uint32_t v)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <int NV>
__device__ __forceinline__ int pair_id(int a, int b) // a < b
{
    if (NV == 4) return a == 0 ? b - 1 : (a == 1 ? b + 1 : 5);
    return a == 0 ? b - 1 : 2;
}

// Scratch of the build kernels (dynamic shared memory, build_shmem() bytes)
struct BuildScratch {
    uint32_t *sbuf;  // SORT_CAP   sort buffer; later entry offsets / cursors
    int *tmp;        // SORT_CAP   flags; later the codes (16 bit) / fan records
    uint32_t *elist; // NE_CAP     element ids, ascending
    uint32_t *telem; // NE_CAP     slot words
    uint32_t *vlist; // NV_CAP     vertex ids, ascending
    uint32_t *bm;    // TR_CAP*BMW row bitmaps over the slots
    int *rowq;       // TR_CAP+1   first entry of every row
    int *cntq;       // NQ_CAP+1   contributions per entry
    int *part;       // TB_THREADS+1
    int *rslot;      // TR_CAP     slot of every row's own vertex
    uint8_t *s2r;    // NV_CAP     row of a slot (255: not a row of the tile)
};
__device__ __forceinline__ BuildScratch build_scratch(uint32_t *sm)
{
    BuildScratch S;
    S.sbuf = sm;
    S.tmp = reinterpret_cast<int *>(S.sbuf + SORT_CAP);
    S.elist = reinterpret_cast<uint32_t *>(S.tmp + SORT_CAP);
    S.telem = S.elist + NE_CAP;
    S.vlist = S.telem + NE_CAP;
    S.bm = S.vlist + NV_CAP;
    S.rowq = reinterpret_cast<int *>(S.bm + TR_CAP * BMW);
    S.cntq = S.rowq + TR_CAP + 1;
    S.part = S.cntq + NQ_CAP + 1;
    S.rslot = S.part + TB_THREADS + 1;
    S.s2r = reinterpret_cast<uint8_t *>(S.rslot + TR_CAP);
    return S;
}

// Steps 1-4 of a tile's construction, shared by the build kernels: the elements touching a row of the tile (ascending),
// the distinct vertices they touch (ascending: slot order = column order), the slot word of every element, the column
// bitmap of every row and the first entry of every row.  Returns whether the tile fits the capacities.
template <int NV>
__device__ int tile_topology(const BuildScratch &S, const int32_t *__restrict__ rord, int r0, int nr, const int32_t *__restrict__ conn,
                             const IncView &V, int &nelem, int &nvt, int &nq, int &nrec)
{
    uint32_t *sbuf = S.sbuf, *elist = S.elist, *telem = S.telem, *vlist = S.vlist, *bm = S.bm;
    int *rowq = S.rowq, *part = S.part, *rslot = S.rslot;
    uint8_t *s2r = S.s2r;
    __shared__ int s_n, s_bad, s_ne, s_nv;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_n = 0;
        s_bad = 0;
        s_ne = s_nv = 0;
    }
    // hash tables (open addressing, 1024 slots, in the sort buffer's tail): vertex id -> local row, later vertex id set
    constexpr int HT = 1024;
    uint32_t *hkey = sbuf + SORT_CAP - 2 * HT, *hval = hkey + HT;
    for (int x = tid; x < HT; x += TB_THREADS) hkey[x] = 0xffffffffu;
    __syncthreads();
    for (int l = tid; l < nr; l += TB_THREADS) {
        const uint32_t v = (uint32_t)rord[r0 + l];
        uint32_t h = (v * 0x9E3779B1u) >> 22;
        while (atomicCAS(&hkey[h], 0xffffffffu, v) != 0xffffffffu) h = (h + 1) & (HT - 1);
        hval[h] = (uint32_t)l;
        atomicAdd(&s_n, V.cnt[rord[r0 + l]]);
    }
    __syncthreads();
    nrec = s_n;
    // 1. elements of the tile.  A record (row l, element k) contributes k when l is the smallest local row among the
    // element's vertices that are rows of the tile: every element exactly once, no sort over all the records
    int fit = (nr <= TR_CAP && s_n <= 65535) ? 1 : 0;
    nelem = 0; nvt = 0; nq = 0;
    if (fit) {
        for (int l = tid; l < nr; l += TB_THREADS) {
            const int row = rord[r0 + l];
            const int c = V.cnt[row];
            for (int e = 0; e < c; ++e) {
                const uint32_t ka = V.inc[V.idx(row, e)];
                const uint32_t k = ka >> 4;
                const int a = ka & 15;
                bool first = true;
                for (int bb = 0; bb < NV && first; ++bb) {
                    if (bb == a) continue;
                    const uint32_t v = (uint32_t)conn[(size_t)k * NV + bb];
                    uint32_t h = (v * 0x9E3779B1u) >> 22;
                    while (hkey[h] != 0xffffffffu) {
                        if (hkey[h] == v) {
                            if ((int)hval[h] < l) first = false;
                            break;
                        }
                        h = (h + 1) & (HT - 1);
                    }
                }
                if (first) {
                    const int o = atomicAdd(&s_ne, 1);
                    if (o < NE_CAP) elist[o] = k;
                }
            }
        }
        __syncthreads();
        nelem = s_ne;
        if (nelem > NE_CAP) fit = 0;
    }
    if (fit) {
        int m = 32;
        while (m < nelem) m <<= 1;
        for (int x = nelem + tid; x < m; x += TB_THREADS) elist[x] = 0xffffffffu;
        __syncthreads();
        blk_sort(elist, m); // ascending element ids
        // 2. distinct vertices: hash set, then sorted
        for (int x = tid; x < HT; x += TB_THREADS) hkey[x] = 0xffffffffu;
        __syncthreads();
        for (int x = tid; x < nelem * NV; x += TB_THREADS) {
            const uint32_t v = (uint32_t)conn[(size_t)elist[x / NV] * NV + (x % NV)];
            uint32_t h = (v * 0x9E3779B1u) >> 22;
            while (true) {
                const uint32_t k = hkey[h];
                if (k == v) break;
                if (k == 0xffffffffu) {
                    // more than NV_CAP distinct vertices: the tile is unfit, stop filling (the table cannot overflow:
                    // at most TB_THREADS insertions are in flight past the test)
                    if (*reinterpret_cast<volatile int *>(&s_nv) > NV_CAP) break;
                    const uint32_t old = atomicCAS(&hkey[h], 0xffffffffu, v);
                    if (old == 0xffffffffu) {
                        const int o = atomicAdd(&s_nv, 1);
                        if (o < NV_CAP) vlist[o] = v;
                        break;
                    }
                    if (old == v) break;
                }
                h = (h + 1) & (HT - 1);
            }
        }
        __syncthreads();
        nvt = s_nv;
        if (nvt > NV_CAP) fit = 0;
    }
    if (fit) {
        int m = 32;
        while (m < nvt) m <<= 1;
        for (int x = nvt + tid; x < m; x += TB_THREADS) vlist[x] = 0xffffffffu;
        __syncthreads();
        blk_sort(vlist, m);
    }
    if (fit) {
        // 3. slots of every element's vertices; rows <-> slots
        for (int e = tid; e < nelem; e += TB_THREADS) {
            uint32_t w = 0;
            for (int a = 0; a < NV; ++a) w |= (uint32_t)bsearch_u32(vlist, nvt, (uint32_t)conn[(size_t)elist[e] * NV + a]) << (8 * a);
            telem[e] = w;
        }
        for (int x = tid; x < NV_CAP; x += TB_THREADS) s2r[x] = 255;
        for (int x = tid; x < TR_CAP * BMW; x += TB_THREADS) bm[x] = 0u;
        __syncthreads();
        for (int l = tid; l < nr; l += TB_THREADS) {
            const int row = rord[r0 + l];
            // a row without any element is not in vlist: it gets no slot and an empty row
            const int sl = nvt > 0 ? bsearch_u32(vlist, nvt, (uint32_t)row) : 0;
            const bool has = nvt > 0 && vlist[sl] == (uint32_t)row;
            rslot[l] = has ? sl : -1;
            if (has) s2r[sl] = (uint8_t)l;
        }
        __syncthreads();
        // 4. column bitmaps of the rows
        for (int e = tid; e < nelem; e += TB_THREADS) {
            const uint32_t w = telem[e];
            for (int a = 0; a < NV; ++a) {
                const int l = s2r[(w >> (8 * a)) & 255u];
                if (l == 255) continue;
                for (int b = 0; b < NV; ++b) {
                    const uint32_t sb = (w >> (8 * b)) & 255u;
                    atomicOr(&bm[l * BMW + (sb >> 5)], 1u << (sb & 31u));
                }
            }
        }
        __syncthreads();
        for (int l = tid; l <= nr; l += TB_THREADS) {
            int L = 0;
            if (l < nr)
                for (int w = 0; w < BMW; ++w) L += __popc(bm[l * BMW + w]);
            rowq[l] = L;
            if (L > 255) s_bad = 1;
        }
        __syncthreads();
        nq = blk_scan(rowq, nr + 1, part);
        if (nq > NQ_CAP || s_bad) fit = 0;
    }
    return fit;
}

// position of slot sb inside row l (number of set bits of the row's bitmap below it)
__device__ __forceinline__ int row_pos(const uint32_t *bm, int l, uint32_t sb)
{
    int p = __popc(bm[l * BMW + (sb >> 5)] & ((1u << (sb & 31u)) - 1u));
    for (int ww = 0; ww < (int)(sb >> 5); ++ww) p += __popc(bm[l * BMW + ww]);
    return p;
}

// One CTA per tile.  WRITE = 0: sizes only (stats[t*8 ..] = nvt, nelem, nq, ncodes, fit, nr).  WRITE = 1: the blob.
template <int NV, int WRITE>
__global__ void __launch_bounds__(TB_THREADS) k_tile_build(const int32_t *__restrict__ rord, const int32_t *__restrict__ tstart,
                                                           const int32_t *__restrict__ conn, const IncView V,
                                                           const double *__restrict__ xyz, int vstride,
                                                           const int32_t *__restrict__ nrowptr, int nes,
                                                           int32_t *__restrict__ stats, const uint32_t *__restrict__ toff,
                                                           uint32_t *__restrict__ blob, const uint32_t *__restrict__ roff,
                                                           uint32_t *__restrict__ rblob)
{
    extern __shared__ uint32_t sm[];
    const BuildScratch S = build_scratch(sm);
    uint32_t *sbuf = S.sbuf, *elist = S.elist, *telem = S.telem, *vlist = S.vlist, *bm = S.bm;
    int *tmp = S.tmp, *rowq = S.rowq, *cntq = S.cntq, *part = S.part, *rslot = S.rslot;
    uint8_t *s2r = S.s2r;
    __shared__ int s_bad;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int r0 = tstart[t], nr = tstart[t + 1] - r0;
    if (tid == 0) s_bad = 0;
    int nelem = 0, nvt = 0, nq = 0, ncodes = 0, s_nrec = 0;
    int fit = tile_topology<NV>(S, rord, r0, nr, conn, V, nelem, nvt, nq, s_nrec);
    if (fit) {
        // 5. contributions per entry: every ordered vertex pair (a, b), a != b, of every element whose vertex a is a row
        for (int x = tid; x <= nq; x += TB_THREADS) cntq[x] = 0;
        __syncthreads();
        for (int e = tid; e < nelem; e += TB_THREADS) {
            const uint32_t w = telem[e];
            for (int a = 0; a < NV; ++a) {
                const int l = s2r[(w >> (8 * a)) & 255u];
                if (l == 255) continue;
                for (int b = 0; b < NV; ++b) {
                    if (b == a) continue;
                    const uint32_t sb = (w >> (8 * b)) & 255u;
                    int p = __popc(bm[l * BMW + (sb >> 5)] & ((1u << (sb & 31u)) - 1u));
                    for (int ww = 0; ww < (int)(sb >> 5); ++ww) p += __popc(bm[l * BMW + ww]);
                    atomicAdd(&cntq[rowq[l] + p], 1);
                }
            }
        }
        __syncthreads();
        int *ncnt = reinterpret_cast<int *>(sbuf); // true list lengths (the scanned array holds the padded ones)
        for (int x = tid; x <= nq; x += TB_THREADS) {
            const int c = x < nq ? cntq[x] : 0;
            if (c > 31) s_bad = 1;
            ncnt[x] = c;
            cntq[x] = (c + 1) & ~1; // every list starts on a 32-bit boundary
        }
        __syncthreads();
        ncodes = blk_scan(cntq, nq + 1, part); // cntq[q] = offset of entry q's list, cntq[nq] = ncodes
        if (ncodes > NC_CAP || ncodes > 65534 || s_bad) fit = 0;
    }
    if (!WRITE) {
        if (tid == 0) {
            int32_t *st = stats + (size_t)t * 8;
            st[0] = nvt; st[1] = nelem; st[2] = nq; st[3] = ncodes; st[4] = fit; st[5] = nr; st[6] = s_nrec;
        }
        return;
    }
    if (!fit) return; // cannot happen: the host only writes tile sets whose tiles all fit
    // 6. fill the lists (cursor = sbuf), then sort every list: ascending code = ascending element
    uint16_t *codes = reinterpret_cast<uint16_t *>(tmp);
    int *ncnt = reinterpret_cast<int *>(sbuf);
    int *cursor = ncnt + NQ_CAP + 1;
    for (int x = tid; x < nq; x += TB_THREADS) cursor[x] = cntq[x];
    for (int x = tid; x < (ncodes + 1) / 2 * 2; x += TB_THREADS) codes[x] = 0;
    __syncthreads();
    for (int e = tid; e < nelem; e += TB_THREADS) {
        const uint32_t w = telem[e];
        for (int a = 0; a < NV; ++a) {
            const int l = s2r[(w >> (8 * a)) & 255u];
            if (l == 255) continue;
            for (int b = 0; b < NV; ++b) {
                if (b == a) continue;
                const uint32_t sb = (w >> (8 * b)) & 255u;
                int p = __popc(bm[l * BMW + (sb >> 5)] & ((1u << (sb & 31u)) - 1u));
                for (int ww = 0; ww < (int)(sb >> 5); ++ww) p += __popc(bm[l * BMW + ww]);
                const int o = atomicAdd(&cursor[rowq[l] + p], 1);
                codes[o] = (uint16_t)((e << 3) | pair_id<NV>(min(a, b), max(a, b)));
            }
        }
    }
    __syncthreads();
    for (int q = tid; q < nq; q += TB_THREADS) {
        const int o = cntq[q], n = ncnt[q];
        for (int x = 1; x < n; ++x) {
            const uint16_t v = codes[o + x];
            int y = x - 1;
            while (y >= 0 && codes[o + y] > v) {
                codes[o + y + 1] = codes[o + y];
                --y;
            }
            codes[o + y + 1] = v;
        }
    }
    __syncthreads();
    // The order inside a list is free (any fixed order is reproducible).  The numeric kernel reads the k-th value of 16
    // consecutive entries in one shared-memory wavefront (64-bit accesses: half a warp at a time), so the lists of every
    // group of 16 entries are re-ordered greedily to put the k-th values of the group in distinct banks.
    for (int gq = tid * 16; gq < nq; gq += TB_THREADS * 16) {
        int maxn = 0;
        for (int j = 0; j < 16 && gq + j < nq; ++j) maxn = max(maxn, ncnt[gq + j]);
        for (int k = 0; k < maxn; ++k) {
            uint32_t used = 0;
            for (int j = 0; j < 16 && gq + j < nq; ++j) {
                const int q = gq + j, n = ncnt[q], o = cntq[q];
                if (k >= n) continue;
                int best = k;
                for (int x = k; x < n; ++x) {
                    const uint32_t c = codes[o + x];
                    const uint32_t bank = ((c & 7u) * nes + (c >> 3)) & 15u;
                    if (!((used >> bank) & 1u)) {
                        best = x;
                        break;
                    }
                }
                const uint16_t c = codes[o + best];
                codes[o + best] = codes[o + k];
                codes[o + k] = c;
                used |= 1u << ((((uint32_t)c & 7u) * nes + ((uint32_t)c >> 3)) & 15u);
            }
        }
    }
    __syncthreads();
    // (element << 3 | pair) -> index of the value in the numeric kernel's [pair][nes] table
    for (int q = tid; q < nq; q += TB_THREADS) {
        const int o = cntq[q], n = ncnt[q];
        for (int x = 0; x < n; ++x) {
            const uint32_t c = codes[o + x];
            codes[o + x] = (uint16_t)((c & 7u) * nes + (c >> 3));
        }
    }
    __syncthreads();
    // 7. the blob
    uint32_t *g = blob + toff[t];
    constexpr int DIM = NV - 1;
    const int o_gbase = HDR, o_grow = o_gbase + pad4(nr), o_rinfo = o_grow + pad4(nr), o_coord = o_rinfo + pad4(nr + 1),
              o_telem = o_coord + pad4(2 * DIM * nvt), o_einfo = o_telem + pad4(nelem), o_codes = o_einfo + pad4(nq + 1),
              words = o_codes + pad4((ncodes + 1) / 2);
    if (tid == 0) {
        g[0] = nr; g[1] = nvt; g[2] = nelem; g[3] = nq; g[4] = ncodes; g[5] = o_gbase; g[6] = o_rinfo; g[7] = o_coord;
        g[8] = o_telem; g[9] = o_einfo; g[10] = o_codes; g[11] = words; g[12] = o_grow; g[13] = g[14] = g[15] = 0;
    }
    for (int l = tid; l < pad4(nr); l += TB_THREADS) {
        g[o_gbase + l] = l < nr ? (uint32_t)nrowptr[rord[r0 + l]] : 0u;
        g[o_grow + l] = l < nr ? (uint32_t)rord[r0 + l] : 0u;
    }
    for (int l = tid; l < pad4(nr + 1); l += TB_THREADS) {
        uint32_t w = 0;
        if (l < nr) {
            const int L = rowq[l + 1] - rowq[l], sl = rslot[l];
            int pd = 0;
            if (sl >= 0) {
                pd = __popc(bm[l * BMW + (sl >> 5)] & ((1u << (sl & 31)) - 1u));
                for (int ww = 0; ww < (sl >> 5); ++ww) pd += __popc(bm[l * BMW + ww]);
            }
            w = (uint32_t)rowq[l] | ((uint32_t)pd << 16) | ((uint32_t)L << 24);
        } else if (l == nr)
            w = (uint32_t)nq;
        g[o_rinfo + l] = w;
    }
    for (int x = tid; x < pad4(2 * DIM * nvt); x += TB_THREADS) {
        uint32_t w = 0;
        if (x < 2 * DIM * nvt) {
            const int v = x / (2 * DIM), c = (x % (2 * DIM)) >> 1;
            const unsigned long long b = (unsigned long long)__double_as_longlong(xyz[(size_t)vlist[v] * vstride + c]);
            w = (x & 1) ? (uint32_t)(b >> 32) : (uint32_t)b;
        }
        g[o_coord + x] = w;
    }
    for (int x = tid; x < pad4(nelem); x += TB_THREADS) g[o_telem + x] = x < nelem ? telem[x] : 0u;
    // entry words: offset of the list in code pairs | list length << 15 | local row << 20
    for (int l = tid; l < nr; l += TB_THREADS)
        for (int q = rowq[l]; q < rowq[l + 1]; ++q) g[o_einfo + q] = (uint32_t)(cntq[q] >> 1) | ((uint32_t)ncnt[q] << 15) | ((uint32_t)l << 20);
    for (int x = nq + tid; x < pad4(nq + 1); x += TB_THREADS) g[o_einfo + x] = 0u;
    const int cw = pad4((ncodes + 1) / 2);
    for (int x = tid; x < cw; x += TB_THREADS) {
        const uint32_t lo = 2 * x < ncodes ? codes[2 * x] : 0u, hi = 2 * x + 1 < ncodes ? codes[2 * x + 1] : 0u;
        g[o_codes + x] = lo | (hi << 16);
    }
    // 8. the record lists of the rows (second blob: right-hand sides, mass diagonals): for every row the elements of its
    // star as (element << 2 | local vertex), ascending element.  header: 0 nrec, 1 o_rroff, 2 o_rrec, 3 words
    __syncthreads();
    int *rc = rowq; // the entry offsets are written out: re-used for the record offsets
    for (int l = tid; l <= nr; l += TB_THREADS) rc[l] = l < nr ? V.cnt[rord[r0 + l]] : 0;
    __syncthreads();
    const int nrec = blk_scan(rc, nr + 1, part);
    uint32_t *rg = rblob + roff[t];
    const int o_rroff = 4, o_rrec = o_rroff + pad4((nr + 2) / 2), rwords = o_rrec + pad4((nrec + 1) / 2);
    if (tid == 0) {
        rg[0] = nrec; rg[1] = o_rroff; rg[2] = o_rrec; rg[3] = rwords;
    }
    uint16_t *rr16 = reinterpret_cast<uint16_t *>(rg + o_rroff);
    for (int l = tid; l < 2 * pad4((nr + 2) / 2); l += TB_THREADS) rr16[l] = (uint16_t)(l <= nr ? rc[l] : nrec);
    uint16_t *rec16 = reinterpret_cast<uint16_t *>(rg + o_rrec);
    for (int l = tid; l < nr; l += TB_THREADS) {
        const int row = rord[r0 + l], c = rc[l + 1] - rc[l];
        for (int e = 0; e < c; ++e) {
            const uint32_t ka = V.inc[V.idx(row, e)];
            rec16[rc[l] + e] = (uint16_t)((bsearch_u32(elist, nelem, ka >> 4) << 2) | (ka & 3u));
        }
    }
    for (int x = nrec + tid; x < 2 * pad4((nrec + 1) / 2); x += TB_THREADS) rec16[x] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// the numeric kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tile_rcp(double d) // ~1 ulp reciprocal: hardware seed + two Newton steps
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double t = fma(-d, r, 1.0);
    r = fma(r, t, r);
    t = fma(-d, r, 1.0);
    return fma(r, t, r);
}

struct TileSmem { // byte offsets of the shared-memory regions
    int buf1;      // second descriptor buffer (the first one is at 0)
    int rb0, rb1;  // record-list buffers (mass forms, right-hand sides)
    int ent, vals, nes;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tile_mbar_init(unsigned long long *mbar)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void tile_bulk(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tile_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_wait(uint32_t bar, uint32_t phase)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(bar), "r"(phase)
                     : "memory");
}

// geometry of one P1 element from the tile's coordinate table: unscaled normals N[a][x] = det * d lambda_a / d x, det
template <int DIM>
__device__ __forceinline__ void tile_geom(const double *coord, uint32_t w, double (&N)[DIM + 1][DIM], double &det)
{
    const double *p0 = coord + DIM * (w & 255u), *p1 = coord + DIM * ((w >> 8) & 255u), *p2 = coord + DIM * ((w >> 16) & 255u);
    if (DIM == 3) {
        const double *p3 = coord + DIM * (w >> 24);
        const double x0 = p0[0], y0 = p0[1], z0 = p0[DIM - 1];
        const double ax = p1[0] - x0, ay = p1[1] - y0, az = p1[DIM - 1] - z0;
        const double bx = p2[0] - x0, by = p2[1] - y0, bz = p2[DIM - 1] - z0;
        const double cx = p3[0] - x0, cy = p3[1] - y0, cz = p3[DIM - 1] - z0;
        // N1 = b x c, N2 = c x a, N3 = a x b, det = a . N1, N0 = -(N1 + N2 + N3)   (Mesh3dn.hpp:126-136)
        N[1][0] = by * cz - bz * cy; N[1][1] = bz * cx - bx * cz; N[1][DIM - 1] = bx * cy - by * cx;
        N[2][0] = cy * az - cz * ay; N[2][1] = cz * ax - cx * az; N[2][DIM - 1] = cx * ay - cy * ax;
        N[DIM][0] = ay * bz - az * by; N[DIM][1] = az * bx - ax * bz; N[DIM][DIM - 1] = ax * by - ay * bx;
        det = ax * N[1][0] + ay * N[1][1] + az * N[1][DIM - 1];
    } else {
        const double x0 = p0[0], y0 = p0[1];
        const double bx = p1[0] - x0, by = p1[1] - y0, cx = p2[0] - x0, cy = p2[1] - y0;
        det = bx * cy - by * cx; // N1 = (cy, -cx), N2 = (-by, bx), N0 = -(N1 + N2)   (fem.hpp:321-324)
        N[1][0] = cy; N[1][1] = -cx;
        N[2][0] = -by; N[2][1] = bx;
    }
#pragma unroll
    for (int x = 0; x < DIM; ++x) {
        double sx = N[1][x];
#pragma unroll
        for (int r = 2; r <= DIM; ++r) sx += N[r][x];
        N[0][x] = -sx;
    }
}

// Persistent CTAs: tile t, t + grid, ...  The descriptor blob of the NEXT tile is brought in by one TMA bulk copy
// (cp.async.bulk, completion on an mbarrier) while the current tile is computed: no thread ever waits on a global load
// in the steady state.  MASS: the form has a mass term; the record lists of the rows come along (second bulk copy) and
// give the measure of every row's star for the diagonal.
template <int DIM, bool MASS>
__global__ void __launch_bounds__(512) k_asm_tiles(const uint32_t *__restrict__ toff, const uint32_t *__restrict__ blob,
                                                   const uint32_t *__restrict__ roff, const uint32_t *__restrict__ rblob, int ntiles,
                                                   double *__restrict__ out, int accumulate, double cw, double cmd, double cmo,
                                                   const TileSmem S)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    constexpr int NV = DIM + 1, NP = DIM * (DIM + 1) / 2;
    double *sE = reinterpret_cast<double *>(smem_raw + S.ent);  // off-diagonal sums of the tile's entries
    double *sV = reinterpret_cast<double *>(smem_raw + S.vals); // [pair][NES] (+ [NP][NES] = det when MASS)
    const int NES = S.nes;
    const int tid = threadIdx.x, nthr = blockDim.x;
    tile_mbar_init(mbar);
    auto issue = [&](int b, int t) { // one thread: bulk copies of tile t's descriptors into buffer b
        const uint32_t w0 = __ldg(toff + t), bytes = (__ldg(toff + t + 1) - w0) * 4u;
        const uint32_t bar = smem_u32(&mbar[b]);
        uint32_t r0 = 0, rbytes = 0;
        if (MASS) {
            r0 = __ldg(roff + t);
            rbytes = (__ldg(roff + t + 1) - r0) * 4u;
        }
        tile_expect(bar, bytes + rbytes);
        tile_bulk(smem_u32(smem_raw + (b ? S.buf1 : 0)), blob + w0, bytes, bar);
        if (MASS) tile_bulk(smem_u32(smem_raw + (b ? S.rb1 : S.rb0)), rblob + r0, rbytes, bar);
    };
    int t = blockIdx.x, b = 0;
    uint32_t ph0 = 0, ph1 = 0;
    if (tid == 0 && t < ntiles) issue(0, t);
    for (; t < ntiles; t += gridDim.x, b ^= 1) {
        if (tid == 0 && t + (int)gridDim.x < ntiles) issue(b ^ 1, t + gridDim.x); // that buffer was released by the last barrier
        tile_wait(smem_u32(&mbar[b]), b ? ph1 : ph0);
        if (b) ph1 ^= 1;
        else ph0 ^= 1;
        const uint32_t *sb = reinterpret_cast<const uint32_t *>(smem_raw + (b ? S.buf1 : 0));
        const int nr = sb[0], nelem = sb[2], nq = sb[3];
        const int32_t *gbase = reinterpret_cast<const int32_t *>(sb + sb[5]);
        const uint32_t *rinfo = sb + sb[6];
        const double *coord = reinterpret_cast<const double *>(sb + sb[7]);
        const uint32_t *telem = sb + sb[8];
        const uint32_t *einfo = sb + sb[9];
        const uint32_t *codes2 = sb + sb[10];
        // ---- every element of the tile once: off-diagonal entries of its element matrix (+ its determinant)
        for (int e = tid; e < nelem; e += nthr) {
            double N[NV][DIM], det;
            tile_geom<DIM>(coord, telem[e], N, det);
            const double s = cw * tile_rcp(det), mo = MASS ? cmo * det : 0.0;
            int k = 0;
#pragma unroll
            for (int a2 = 0; a2 < NV; ++a2)
#pragma unroll
                for (int b2 = a2 + 1; b2 < NV; ++b2) {
                    double d = N[a2][0] * N[b2][0];
#pragma unroll
                    for (int x = 1; x < DIM; ++x) d = fma(N[a2][x], N[b2][x], d);
                    sV[k * NES + e] = fma(d, s, mo); // pair order (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) = pair_id
                    ++k;
                }
            if (MASS) sV[NP * NES + e] = det;
        }
        __syncthreads();
        // ---- every entry of the tile's rows: sum of the contributions of the elements around its edge, in a register
        for (int q = tid; q < nq; q += nthr) {
            const uint32_t info = einfo[q];
            const int n = (info >> 15) & 31u;
            const uint32_t *cp = codes2 + (info & 0x7fffu);
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k < n; k += 2) {
                const uint32_t c2 = *cp++;
                acc += sV[c2 & 0xffffu];
                if (k + 1 < n) acc += sV[c2 >> 16];
            }
            sE[q] = acc;
            if (n > 0) {
                const int l = (info >> 20) & 255u;
                double *dst = out + (size_t)gbase[l] + (q - (int)(rinfo[l] & 0xffffu));
                *dst = accumulate ? *dst + acc : acc;
            }
        }
        __syncthreads();
        // ---- diagonals: K_ii = -sum_{j != i} K_ij (partition of unity: the off-diagonal entries carry m_o |K| each, DIM
        // per element); mass: + (m_d + DIM m_o) sum of |K| over the star (record lists, 4 lanes per row)
        if (!MASS) {
            for (int l = tid; l < nr; l += nthr) {
                const uint32_t ri = rinfo[l];
                const int q0 = ri & 0xffffu, L = ri >> 24;
                if (L == 0) continue;
                double sx = 0.0;
                for (int k = 0; k < L; ++k) sx += sE[q0 + k];
                double *dst = out + (size_t)gbase[l] + ((ri >> 16) & 255u);
                *dst = accumulate ? *dst - sx : -sx;
            }
        } else {
            for (int i0 = 0; i0 < 4 * nr; i0 += nthr) {
                const int idx = i0 + tid, l = idx >> 2, j = idx & 3;
                double sx = 0.0, sd = 0.0;
                uint32_t ri = 0;
                if (l < nr) {
                    ri = rinfo[l];
                    const int q0 = ri & 0xffffu, L = ri >> 24;
                    for (int k = j; k < L; k += 4) sx += sE[q0 + k];
                    const uint32_t *rb = reinterpret_cast<const uint32_t *>(smem_raw + (b ? S.rb1 : S.rb0));
                    const uint16_t *rro = reinterpret_cast<const uint16_t *>(rb + rb[1]), *rec = reinterpret_cast<const uint16_t *>(rb + rb[2]);
                    const int o1 = rro[l + 1];
                    for (int r = rro[l] + j; r < o1; r += 4) sd += sV[NP * NES + (rec[r] >> 2)];
                }
                sx += __shfl_xor_sync(0xffffffffu, sx, 1);
                sx += __shfl_xor_sync(0xffffffffu, sx, 2);
                sd += __shfl_xor_sync(0xffffffffu, sd, 1);
                sd += __shfl_xor_sync(0xffffffffu, sd, 2);
                if (l < nr && j == 0 && (ri >> 24) != 0) {
                    const double d = (cmd + DIM * cmo) * sd - sx;
                    double *dst = out + (size_t)gbase[l] + ((ri >> 16) & 255u);
                    *dst = accumulate ? *dst + d : d;
                }
            }
        }
        __syncthreads(); // buffer b, sV and sE are free again
    }
}

// Right-hand side of a P1 space by tiles: b_i = sum over the star of i of |K| (c_0 <lambda> + sum_x c_x d_x lambda_i),
// i.e. (RFAC c_0 <lambda>) sum det + (RFAC W) sum_x c_x sum N_i[x]  (AssembleLinearForm, fflib/problem.cpp:10878-11227,
// Element_rhs :7917-7985).  Every element is evaluated once per tile (its determinant, with gradient terms its
// normals); 4 lanes per row add the star's values through the row's record list.  Only the head of the tile descriptor
// (row ids, coordinates, element words) is copied, plus the record lists.
struct RhsCoef {
    double cval[3];     // per component: RFAC * c_0 * <lambda>
    double cgrad[3][3]; // per component: RFAC * W * c_x
};

template <int DIM, bool GRAD>
__global__ void __launch_bounds__(512) k_rhs_tiles(const uint32_t *__restrict__ toff, const uint32_t *__restrict__ tpre,
                                                   const uint32_t *__restrict__ blob, const uint32_t *__restrict__ roff,
                                                   const uint32_t *__restrict__ rblob, int ntiles, double *__restrict__ bvec, int nc,
                                                   int accumulate, const __grid_constant__ RhsCoef C, const TileSmem S)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    constexpr int NV = DIM + 1;
    double *sV = reinterpret_cast<double *>(smem_raw + S.vals); // det[NES] (+ N[a][x][NES] when GRAD)
    const int NES = S.nes;
    const int tid = threadIdx.x, nthr = blockDim.x;
    tile_mbar_init(mbar);
    auto issue = [&](int b, int t) {
        const uint32_t w0 = __ldg(toff + t), bytes = __ldg(tpre + t) * 4u;
        const uint32_t r0 = __ldg(roff + t), rbytes = (__ldg(roff + t + 1) - r0) * 4u;
        const uint32_t bar = smem_u32(&mbar[b]);
        tile_expect(bar, bytes + rbytes);
        tile_bulk(smem_u32(smem_raw + (b ? S.buf1 : 0)), blob + w0, bytes, bar);
        tile_bulk(smem_u32(smem_raw + (b ? S.rb1 : S.rb0)), rblob + r0, rbytes, bar);
    };
    int t = blockIdx.x, b = 0;
    uint32_t ph0 = 0, ph1 = 0;
    if (tid == 0 && t < ntiles) issue(0, t);
    for (; t < ntiles; t += gridDim.x, b ^= 1) {
        if (tid == 0 && t + (int)gridDim.x < ntiles) issue(b ^ 1, t + gridDim.x);
        tile_wait(smem_u32(&mbar[b]), b ? ph1 : ph0);
        if (b) ph1 ^= 1;
        else ph0 ^= 1;
        const uint32_t *sb = reinterpret_cast<const uint32_t *>(smem_raw + (b ? S.buf1 : 0));
        const uint32_t *rb = reinterpret_cast<const uint32_t *>(smem_raw + (b ? S.rb1 : S.rb0));
        const int nr = sb[0], nelem = sb[2];
        const int32_t *grow = reinterpret_cast<const int32_t *>(sb + sb[12]);
        const double *coord = reinterpret_cast<const double *>(sb + sb[7]);
        const uint32_t *telem = sb + sb[8];
        const uint16_t *rro = reinterpret_cast<const uint16_t *>(rb + rb[1]), *rec = reinterpret_cast<const uint16_t *>(rb + rb[2]);
        for (int e = tid; e < nelem; e += nthr) {
            double N[NV][DIM], det;
            tile_geom<DIM>(coord, telem[e], N, det);
            sV[e] = det;
            if (GRAD) {
#pragma unroll
                for (int a = 0; a < NV; ++a)
#pragma unroll
                    for (int x = 0; x < DIM; ++x) sV[(1 + a * DIM + x) * NES + e] = N[a][x];
            }
        }
        __syncthreads();
        for (int i0 = 0; i0 < 4 * nr; i0 += nthr) {
            const int idx = i0 + tid, l = idx >> 2, j = idx & 3;
            double sd = 0.0, sn[DIM];
#pragma unroll
            for (int x = 0; x < DIM; ++x) sn[x] = 0.0;
            if (l < nr) {
                const int o1 = rro[l + 1];
                for (int r = rro[l] + j; r < o1; r += 4) {
                    const uint32_t c = rec[r];
                    sd += sV[c >> 2];
                    if (GRAD) {
#pragma unroll
                        for (int x = 0; x < DIM; ++x) sn[x] += sV[(1 + (c & 3u) * DIM + x) * NES + (c >> 2)];
                    }
                }
            }
            sd += __shfl_xor_sync(0xffffffffu, sd, 1);
            sd += __shfl_xor_sync(0xffffffffu, sd, 2);
            if (GRAD) {
#pragma unroll
                for (int x = 0; x < DIM; ++x) {
                    sn[x] += __shfl_xor_sync(0xffffffffu, sn[x], 1);
                    sn[x] += __shfl_xor_sync(0xffffffffu, sn[x], 2);
                }
            }
            if (l < nr && j == 0) {
                for (int c = 0; c < nc; ++c) {
                    double v = C.cval[c] * sd;
                    if (GRAD) {
#pragma unroll
                        for (int x = 0; x < DIM; ++x) v = fma(C.cgrad[c][x], sn[x], v);
                    }
                    double *dst = bvec + (size_t)grow[l] * nc + c;
                    *dst = accumulate ? *dst + v : v;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// FANS (3-D scalar P1 stiffness forms, the headline kernel of round 2).
//
// ncu on the round-1 tile kernel: 72 % of the shared-memory pipe, 25 % of the FP64 pipe - per evaluated element 12 LDS.64
// of coordinates, 6 STS.64 of values, and 12 LDS.64 + 6 code reads on the gather side.  A warp LDS.64 costs two cycles of
// the 128 B/clk pipe whatever the addresses (tools/micro/pipes.cu), and so does a shuffle, so the only way down is fewer
// bytes through shared memory per element.  Elements are therefore taken in FANS: runs of <= KMAX elements around a
// common edge (the AXIS, the longest edge of each element: the main diagonal of the cell in BuildCube's meshes), ordered
// around it so that consecutive elements share a face.  One thread walks one fan:
//   * axis end points p, q loaded once, one new ring vertex r_{t+1} per element (3 LDS.64 instead of 12);
//   * a x (r_{t+1} - p) is the normal of a face of element t AND (negated) of element t+1: computed once;
//   * the contributions to the axis edge are summed over the whole fan in a register, those to the spokes p-r_t, q-r_t
//     over the two elements that share them: 3k+3 values stored for k elements instead of 6k, and the entry gather
//     reads as many fewer.
// Everything that is used once (fan records, entry rows, contribution codes) is read straight from global memory with
// coalesced loads; only the coordinates of the tile's vertices, the row tables and the group tables go through shared
// memory (one TMA bulk copy per tile, double-buffered).  Contribution codes are stored TRANSPOSED per group of 32
// consecutive entries (code word kk of lane l at kk*32 + l, lists padded with the index of a zero) : no per-entry
// offsets, 128-byte loads.  Same owner-gather as before: no atomics on doubles, fixed order, bit-reproducible.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KMAX = 7;          // elements per fan (8 ring vertices)
constexpr int NVAL_CAP = 20000;  // value slots per tile
constexpr int FHDR = 16;
// tile blob v2 (32-bit words): header [0 nr, 1 nvt, 2 nfan, 3 nq, 4 ngf (fan groups), 5 nge (entry groups), 6 nvals,
//   7 o_gbase, 8 o_rinfo, 9 o_coord, 10 o_fans, 11 o_fgrp, 12 o_egrp, 13 o_erow, 14 o_codes, 15 words]
//   HEAD (copied to shared memory): header, gbase[nr], rinfo[nr], fgrp[ngf] (first value slot | kmax << 16),
//        egrp[nge+1] (first code word of the group), coord[3 nvt] doubles
//   then, read from global memory: fans[nfan] uint4, erow[nq] bytes, codes

// elements of a run (all share the axis) -> arcs of <= KMAX elements, consecutive elements share a ring vertex
struct FanOut {
    uint32_t *rec; // 4 words per fan
    int *count;
};
__device__ void emit_fan(const FanOut &F, uint32_t p, uint32_t q, const uint32_t *ring, int k, uint32_t key16)
{
    const int idx = atomicAdd(F.count, 1);
    if (idx >= NE_CAP) return;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = i <= k ? ring[i] : 0u;
    uint32_t *w = F.rec + 4 * idx;
    w[0] = p | (q << 8) | (r[0] << 16) | (r[1] << 24);
    w[1] = r[2] | (r[3] << 8) | (r[4] << 16) | (r[5] << 24);
    w[2] = r[6] | (r[7] << 8) | ((uint32_t)k << 16);
    w[3] = ((uint32_t)(KMAX - k) << 24) | (key16 << 8) | r[0]; // sort key: longest first, then axis, then first ring vertex
}

template <int WRITE>
__global__ void __launch_bounds__(TB_THREADS) k_fan_build(const int32_t *__restrict__ rord, const int32_t *__restrict__ tstart,
                                                          const int32_t *__restrict__ conn, const IncView V,
                                                          const double *__restrict__ xyz, int vstride,
                                                          const int32_t *__restrict__ nrowptr, int32_t *__restrict__ stats,
                                                          const uint32_t *__restrict__ foff, uint32_t *__restrict__ fblob,
                                                          const uint32_t *__restrict__ fcoff, uint32_t *__restrict__ fcblob)
{
    extern __shared__ uint32_t sm[];
    const BuildScratch S = build_scratch(sm);
    uint32_t *sbuf = S.sbuf, *telem = S.telem, *vlist = S.vlist, *bm = S.bm;
    int *rowq = S.rowq, *cntq = S.cntq, *part = S.part, *rslot = S.rslot;
    uint8_t *s2r = S.s2r;
    uint32_t *frec = reinterpret_cast<uint32_t *>(S.tmp); // 4 words per fan (NE_CAP fans at most)
    __shared__ int s_nf, s_bad, s_nvals;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int r0 = tstart[t], nr = tstart[t + 1] - r0;
    if (tid == 0) {
        s_nf = 0;
        s_bad = 0;
        s_nvals = 0;
    }
    int nelem = 0, nvt = 0, nq = 0, nrec = 0, nfan = 0, ngf = 0, nge = 0, nvals = 0, ncw = 0;
    int fit = tile_topology<4>(S, rord, r0, nr, conn, V, nelem, nvt, nq, nrec);
    // the sort buffer (SORT_CAP = 8192 words), by use:
    uint32_t *ekey = sbuf;          // [0, 2048): element sort keys; from F3 on: fan order (order[rank] = fan index)
    uint32_t *fkey = sbuf + NE_CAP; // [2048, 4096): fan sort keys (F3 only)
    uint32_t *skey = sbuf + NE_CAP; // [2048, 6144): entries sorted by decreasing list length: (63 - length) << 12 | entry (from F6 on)
    int *egl = reinterpret_cast<int *>(sbuf + 6144);      // [6144, 6400): first code word of every entry group
    int *fgt = reinterpret_cast<int *>(sbuf + 6400);      // [6400, 6464): fan groups: first value slot | kmax << 16
    int *epos = reinterpret_cast<int *>(sm) + build_words(); // NQ_CAP+16: sorted position of every entry (extra scratch of this kernel)
    int *cursor = epos + 4096;                               // NQ_CAP+16: fill cursors of the entries (F7 only)
    int *rcnt = cursor + 4096;                               // TR_CAP: fans around every row's vertex (right-hand sides, part C)
    uint32_t *rkey = reinterpret_cast<uint32_t *>(rcnt + TR_CAP); // TR_CAP: rows sorted by decreasing count
    int *rpos = reinterpret_cast<int *>(rkey + TR_CAP);      // TR_CAP: sorted position of every row
    int *rgl = rpos + TR_CAP;                                // TR_CAP/32 + 1: first code word of every row group
    int *rfg = rgl + 16;                                     // 64: fan groups of the right-hand side: first value slot
    uint32_t *scw = reinterpret_cast<uint32_t *>(rfg + 64);  // 4096 words: the code lists while they are filled and ordered
    static_assert(NE_CAP == 2048 && NQ_CAP + 1 <= 4096 && SORT_CAP >= 6464, "scratch layout of k_fan_build");
    if (fit) {
        // F1. axis of every element: its longest edge (ties: the smallest pair of global vertex ids, the same choice in
        // every element around the edge)
        for (int e = tid; e < nelem; e += TB_THREADS) {
            const uint32_t w = telem[e];
            uint32_t sl[4], gv[4];
            double X[4][3];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                sl[a] = (w >> (8 * a)) & 255u;
                gv[a] = vlist[sl[a]];
#pragma unroll
                for (int c = 0; c < 3; ++c) X[a][c] = xyz[(size_t)gv[a] * vstride + c];
            }
            double best = -1.0;
            uint32_t blo = 0, bhi = 0, bsl = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a + 1; b < 4; ++b) {
                    const int i0 = gv[a] < gv[b] ? a : b, i1 = gv[a] < gv[b] ? b : a; // canonical order: same rounding everywhere
                    const double dx = X[i1][0] - X[i0][0], dy = X[i1][1] - X[i0][1], dz = X[i1][2] - X[i0][2];
                    const double l2 = dx * dx + dy * dy + dz * dz;
                    const uint32_t lo = gv[i0], hi = gv[i1];
                    if (l2 > best || (l2 == best && (lo < blo || (lo == blo && hi < bhi)))) {
                        best = l2;
                        blo = lo;
                        bhi = hi;
                        bsl = (min(sl[a], sl[b]) << 8) | max(sl[a], sl[b]);
                    }
                }
            ekey[e] = (bsl << 11) | (uint32_t)e;
        }
        int m = 32;
        while (m < nelem) m <<= 1;
        for (int x = nelem + tid; x < m; x += TB_THREADS) ekey[x] = 0xffffffffu;
        __syncthreads();
        blk_sort(ekey, m);
        // F2. one thread per run of elements with the same axis: chains around the axis
        const FanOut F{frec, &s_nf};
        for (int x = tid; x < nelem; x += TB_THREADS) {
            const uint32_t key16 = ekey[x] >> 11;
            if (x > 0 && (ekey[x - 1] >> 11) == key16) continue;
            const uint32_t p = key16 >> 8, q = key16 & 255u;
            int y = x;
            while (y < nelem && (ekey[y] >> 11) == key16) {
                // chunk of at most 32 elements of the run
                uint32_t ru[32], rv[32];
                int mrun = 0;
                while (y < nelem && mrun < 32 && (ekey[y] >> 11) == key16) {
                    const uint32_t w = telem[ekey[y] & 2047u];
                    uint32_t o[2];
                    int no = 0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const uint32_t sa = (w >> (8 * a)) & 255u;
                        if (sa != p && sa != q && no < 2) o[no++] = sa;
                    }
                    ru[mrun] = o[0];
                    rv[mrun] = o[1];
                    ++mrun;
                    ++y;
                }
                uint32_t used = 0;
                for (int i = 0; i < mrun; ++i) {
                    if ((used >> i) & 1u) continue;
                    // arc through element i: ring[h .. h+k], grown at both ends
                    uint32_t ring[2 * KMAX + 2];
                    int h = KMAX, k = 1;
                    ring[h] = ru[i];
                    ring[h + 1] = rv[i];
                    used |= 1u << i;
                    bool grown = true;
                    while (grown && k < KMAX) {
                        grown = false;
                        for (int j = 0; j < mrun && k < KMAX; ++j) {
                            if ((used >> j) & 1u) continue;
                            const uint32_t tail = ring[h + k], head = ring[h];
                            if (ru[j] == tail || rv[j] == tail) {
                                ring[h + k + 1] = ru[j] == tail ? rv[j] : ru[j];
                                ++k;
                                used |= 1u << j;
                                grown = true;
                            } else if (ru[j] == head || rv[j] == head) {
                                --h;
                                ring[h] = ru[j] == head ? rv[j] : ru[j];
                                ++k;
                                used |= 1u << j;
                                grown = true;
                            }
                        }
                    }
                    // canonical direction (the walk above depends on nothing but the sorted run, this is for the key)
                    emit_fan(F, p, q, ring + h, k, key16);
                }
            }
        }
        __syncthreads();
        nfan = s_nf;
        if (nfan > NE_CAP) fit = 0;
    }
    if (fit) {
        // F3. fans sorted by decreasing length (keys are unique: a vertex starts one arc of an axis at most ... if not,
        // equal keys still land on distinct ranks below)
        int m = 32;
        while (m < nfan) m <<= 1;
        for (int x = tid; x < m; x += TB_THREADS) fkey[x] = x < nfan ? frec[4 * x + 3] : 0xffffffffu;
        __syncthreads();
        blk_sort(fkey, m);
        for (int x = tid; x < nfan; x += TB_THREADS) ekey[x] = 0xffffffffu; // order[rank] = fan index
        __syncthreads();
        for (int x = tid; x < nfan; x += TB_THREADS) {
            int rk = bsearch_u32(fkey, nfan, frec[4 * x + 3]);
            while (atomicCAS(&ekey[rk], 0xffffffffu, (uint32_t)x) != 0xffffffffu) ++rk; // equal keys (non-manifold input): next rank
        }
        __syncthreads();
        // F4. value slots: fan of rank f = 32 G + lane, value j at vb[G] + 32 j + lane, 3 kmax(G) + 3 values per lane
        ngf = (nfan + 31) >> 5;
        if (tid == 0) {
            int vb = 1; // slot 0 holds a zero (padding of the code lists)
            for (int G = 0; G < ngf && G < 64; ++G) {
                const int kmax = (int)((frec[4 * ekey[G * 32] + 2] >> 16) & 15u);
                fgt[G] = vb | (kmax << 16);
                vb += 32 * (3 * kmax + 3);
                if (vb > 65535) break;
            }
            s_nvals = vb;
        }
        __syncthreads();
        nvals = s_nvals;
        if (nvals > NVAL_CAP || ngf > 64) fit = 0;
    }
    // every value of every fan contributes to the entries (row x, column y) and (row y, column x) of its edge x-y
    auto for_contribs = [&](auto &&add) {
        for (int f = tid; f < nfan; f += TB_THREADS) {
            const uint32_t *w = frec + 4 * ekey[f];
            const int G = f >> 5, lane = f & 31;
            const int vb = fgt[G] & 0xffff, kmax = fgt[G] >> 16, K1 = kmax + 1;
            const int k = (int)((w[2] >> 16) & 15u);
            const uint32_t p = w[0] & 255u, q = (w[0] >> 8) & 255u;
            const unsigned long long rr = (unsigned long long)(w[0] >> 16) | ((unsigned long long)w[1] << 16) |
                                          ((unsigned long long)(w[2] & 0xffffu) << 48);
            // value j of the fan in lane l sits at vb + 32 j + ((l + j) & 31): a warp store is a permutation of 32 consecutive
            // doubles (no bank conflict), and the values of ONE fan fall in different banks (the entries of a row that
            // read them together do not collide)
            auto slot = [&](int j) { return vb + 32 * j + ((lane + j) & 31); };
            add(p, q, slot(0));
            for (int tt = 0; tt <= k; ++tt) {
                const uint32_t r = (uint32_t)(rr >> (8 * tt)) & 255u;
                add(p, r, slot(1 + tt));
                add(q, r, slot(1 + K1 + tt));
                if (tt < k) add(r, (uint32_t)(rr >> (8 * (tt + 1))) & 255u, slot(1 + 2 * K1 + tt));
            }
        }
    };
    if (fit) {
        // F5. contributions per entry
        for (int x = tid; x <= nq; x += TB_THREADS) cntq[x] = 0;
        __syncthreads();
        for_contribs([&](uint32_t x, uint32_t y, int) {
            int l = s2r[x];
            if (l != 255) atomicAdd(&cntq[rowq[l] + row_pos(bm, l, y)], 1);
            l = s2r[y];
            if (l != 255) atomicAdd(&cntq[rowq[l] + row_pos(bm, l, x)], 1);
        });
        __syncthreads();
        // F6. entries sorted by decreasing list length, taken in groups of 32 (one warp): the group's lists are padded to
        // the longest = the first one, rounded up to even (two 16-bit codes per word) - next to no padding
        nge = (nq + 31) >> 5;
        int m4 = 32; // (the sort works on the next power of two, not on the capacity: 1024 for a typical tile)
        while (m4 < nq) m4 <<= 1;
        for (int x = tid; x < m4; x += TB_THREADS) {
            uint32_t key = 0xffffffffu;
            if (x < nq) {
                if (cntq[x] > 62) s_bad = 1;
                key = ((uint32_t)(63 - min(cntq[x], 62)) << 12) | (uint32_t)x;
            }
            skey[x] = key;
        }
        __syncthreads();
        blk_sort(skey, m4);
        for (int e = tid; e < nq; e += TB_THREADS) epos[skey[e] & 0xfffu] = e;
        for (int g = tid; g <= nge; g += TB_THREADS) egl[g] = g < nge ? ((cntq[skey[g * 32] & 0xfffu] + 1) >> 1) * 32 : 0; // words
        __syncthreads();
        if (nge + 1 > 256) fit = 0;
        else ncw = blk_scan(egl, nge + 1, part);
        if (s_bad) fit = 0;
    }
    // F8. right-hand sides (part C, its own blob): a fan gives the sum of |det| over its elements to its axis vertices
    // (value 0) and det_{t-1} + det_t to its ring vertex t (value 1 + t); a row adds one value per fan around its vertex.
    // Rows sorted by decreasing number of fans, groups of 32 rows, transposed code lists as for the entries.
    int nvals_r = 0, ncw_r = 0, ngr = 0;
    auto for_vertex_sums = [&](auto &&add) {
        for (int f = tid; f < nfan; f += TB_THREADS) {
            const uint32_t *w = frec + 4 * ekey[f];
            const int G = f >> 5, lane = f & 31;
            const int vb = rfg[G];
            const int k = (int)((w[2] >> 16) & 15u);
            const unsigned long long rr = (unsigned long long)(w[0] >> 16) | ((unsigned long long)w[1] << 16) |
                                          ((unsigned long long)(w[2] & 0xffffu) << 48);
            auto slot = [&](int j) { return vb + 32 * j + ((lane + j) & 31); };
            add(w[0] & 255u, slot(0));
            add((w[0] >> 8) & 255u, slot(0));
            for (int tt = 0; tt <= k; ++tt) add((uint32_t)(rr >> (8 * tt)) & 255u, slot(1 + tt));
        }
    };
    if (fit) {
        if (tid == 0) {
            int vb = 1;
            for (int G = 0; G < ngf; ++G) {
                rfg[G] = vb;
                vb += 32 * ((fgt[G] >> 16) + 2);
            }
            s_nvals = vb;
        }
        for (int l = tid; l < TR_CAP; l += TB_THREADS) rcnt[l] = 0;
        __syncthreads();
        nvals_r = s_nvals;
        for_vertex_sums([&](uint32_t x, int) {
            const int l = s2r[x];
            if (l != 255) atomicAdd(&rcnt[l], 1);
        });
        __syncthreads();
        ngr = (nr + 31) >> 5;
        int mr = 32;
        while (mr < nr) mr <<= 1;
        for (int l = tid; l < mr; l += TB_THREADS) {
            uint32_t key = 0xffffffffu;
            if (l < nr) {
                if (rcnt[l] > 126) s_bad = 1;
                key = ((uint32_t)(127 - min(rcnt[l], 126)) << 12) | (uint32_t)l;
            }
            rkey[l] = key;
        }
        __syncthreads();
        blk_sort(rkey, mr);
        for (int e = tid; e < nr; e += TB_THREADS) rpos[rkey[e] & 0xfffu] = e;
        for (int g = tid; g <= ngr; g += TB_THREADS) rgl[g] = g < ngr ? ((rcnt[rkey[g * 32] & 0xfffu] + 1) >> 1) * 32 : 0;
        __syncthreads();
        if (tid == 0) { // (at most 9 groups)
            int o = 0;
            for (int g = 0; g <= ngr; ++g) {
                const int w = rgl[g];
                rgl[g] = o;
                o += w;
            }
        }
        __syncthreads();
        ncw_r = rgl[ngr];
        if (s_bad) fit = 0;
    }
    if (!WRITE) {
        if (tid == 0) {
            int32_t *st = stats + (size_t)t * 12;
            st[0] = nvt; st[1] = nelem; st[2] = nq; st[3] = nfan; st[4] = fit; st[5] = nr; st[6] = nvals; st[7] = ncw; st[8] = ngf; st[9] = nge;
            st[10] = nvals_r; st[11] = ncw_r;
        }
        return;
    }
    if (!fit) return; // (the host only writes fan sets whose tiles all fit)
    // F7. the blob
    // part A (what the fans need): header, fan groups, coordinates, fan records.  part B (what the gather needs): header,
    // row bases, row words, entry groups, entry words, codes.  Two bulk copies with their own barriers: A of the next tile
    // is fetched while this tile's entries are summed, B of the next tile while its fans are evaluated.
    uint32_t *gA = fblob + foff[t];
    const int o_fgrp = FHDR, o_coord = o_fgrp + pad4(ngf), o_fans = o_coord + pad4(6 * nvt), wordsA = o_fans + 4 * nfan;
    uint32_t *g = gA + wordsA; // part B: offsets below are relative to it
    const int o_gbase = 8, o_rinfo = o_gbase + pad4(nr), o_egrp = o_rinfo + pad4(nr), o_einfo = o_egrp + pad4(nge + 1),
              o_codes = o_einfo + pad4(nq), wordsB = o_codes + pad4(ncw);
    if (tid == 0) {
        gA[0] = nr; gA[1] = nvt; gA[2] = nfan; gA[3] = nq; gA[4] = ngf; gA[5] = nge; gA[6] = nvals; gA[7] = o_fgrp; gA[8] = o_coord;
        gA[9] = o_fans; gA[10] = wordsA; gA[11] = wordsB; gA[12] = gA[13] = gA[14] = gA[15] = 0;
        g[0] = nr; g[1] = nq; g[2] = nge; g[3] = o_gbase; g[4] = o_rinfo; g[5] = o_egrp; g[6] = o_einfo; g[7] = o_codes;
    }
    for (int l = tid; l < pad4(nr); l += TB_THREADS) {
        g[o_gbase + l] = l < nr ? (uint32_t)nrowptr[rord[r0 + l]] : 0u;
        uint32_t w = 0;
        if (l < nr) {
            const int L = rowq[l + 1] - rowq[l], sl = rslot[l];
            const int pd = sl >= 0 ? row_pos(bm, l, (uint32_t)sl) : 0;
            w = (uint32_t)rowq[l] | ((uint32_t)pd << 16) | ((uint32_t)L << 24);
            // entry words, in sorted order: position in the row | local row << 8 | entry (CSR order) << 16 | diagonal << 31
            for (int qq = rowq[l]; qq < rowq[l + 1]; ++qq) {
                const int pos = qq - rowq[l];
                g[o_einfo + epos[qq]] = (uint32_t)pos | ((uint32_t)l << 8) | ((uint32_t)qq << 16) | ((sl >= 0 && pos == pd) ? 0x80000000u : 0u);
            }
        }
        g[o_rinfo + l] = w;
    }
    for (int x = nq + tid; x < pad4(nq); x += TB_THREADS) g[o_einfo + x] = 0x80000000u;
    for (int x = tid; x < pad4(ngf); x += TB_THREADS) gA[o_fgrp + x] = x < ngf ? (uint32_t)fgt[x] : 0u;
    // entry groups: first code word | code words per lane << 24
    for (int x = tid; x < pad4(nge + 1); x += TB_THREADS)
        g[o_egrp + x] = x < nge ? ((uint32_t)egl[x] | ((uint32_t)((egl[x + 1] - egl[x]) >> 5) << 24)) : (uint32_t)ncw;
    for (int x = tid; x < pad4(6 * nvt); x += TB_THREADS) {
        uint32_t w = 0;
        if (x < 6 * nvt) {
            const int v = x / 6, c = (x % 6) >> 1;
            const unsigned long long b = (unsigned long long)__double_as_longlong(xyz[(size_t)vlist[v] * vstride + c]);
            w = (x & 1) ? (uint32_t)(b >> 32) : (uint32_t)b;
        }
        gA[o_coord + x] = w;
    }
    for (int f = tid; f < nfan; f += TB_THREADS) {
        const uint32_t *w = frec + 4 * ekey[f];
        gA[o_fans + 4 * f] = w[0];
        gA[o_fans + 4 * f + 1] = w[1];
        gA[o_fans + 4 * f + 2] = w[2];
        gA[o_fans + 4 * f + 3] = 0u;
    }
    // the lists are filled, sorted and re-ordered in shared memory when they fit (4096 words: always on the tile sizes in
    // use) and written out with coalesced stores; straight in global memory otherwise
    const bool cw_sh = pad4(ncw) <= 4096;
    uint32_t *cwbuf = cw_sh ? scw : g + o_codes;
    for (int x = tid; x < pad4(ncw); x += TB_THREADS) cwbuf[x] = 0u;
    for (int x = tid; x <= nq; x += TB_THREADS) cursor[x] = 0;
    __syncthreads();
    uint16_t *gc = reinterpret_cast<uint16_t *>(cwbuf);
    auto code_at = [&](int qq, int c) -> uint16_t & {
        const int e = epos[qq];
        return gc[2 * (egl[e >> 5] + (c >> 1) * 32 + (e & 31)) + (c & 1)];
    };
    for_contribs([&](uint32_t x, uint32_t y, int slot) {
        int l = s2r[x];
        if (l != 255) {
            const int qq = rowq[l] + row_pos(bm, l, y);
            code_at(qq, atomicAdd(&cursor[qq], 1)) = (uint16_t)slot;
        }
        l = s2r[y];
        if (l != 255) {
            const int qq = rowq[l] + row_pos(bm, l, x);
            code_at(qq, atomicAdd(&cursor[qq], 1)) = (uint16_t)slot;
        }
    });
    __syncthreads();
    // every list in ascending order of the value slots: the summation order does not depend on the order of arrival
    for (int qq = tid; qq < nq; qq += TB_THREADS) {
        const int n = cntq[qq];
        for (int x = 1; x < n; ++x) {
            const uint16_t v = code_at(qq, x);
            int y = x - 1;
            while (y >= 0 && code_at(qq, y) > v) {
                code_at(qq, y + 1) = code_at(qq, y);
                --y;
            }
            code_at(qq, y + 1) = v;
        }
    }
    __syncthreads();
    // ... then, inside every half-warp of 16 (sorted) entries, re-ordered greedily so that the values read together (the
    // k-th of each list) sit in distinct banks: bank of a 64-bit value = slot mod 16
    for (int ge = tid * 16; ge < nq; ge += TB_THREADS * 16) {
        const int maxn = cntq[skey[ge] & 0xfffu];
        for (int k = 0; k < maxn; ++k) {
            uint32_t usedb = 0;
            for (int j = 0; j < 16 && ge + j < nq; ++j) {
                const int qq = skey[ge + j] & 0xfffu, n = cntq[qq];
                if (k >= n) continue;
                int best = k;
                for (int x = k; x < n; ++x) {
                    if (!((usedb >> (code_at(qq, x) & 15u)) & 1u)) {
                        best = x;
                        break;
                    }
                }
                const uint16_t c = code_at(qq, best);
                code_at(qq, best) = code_at(qq, k);
                code_at(qq, k) = c;
                usedb |= 1u << (c & 15u);
            }
        }
    }
    __syncthreads();
    if (cw_sh)
        for (int x = tid; x < pad4(ncw); x += TB_THREADS) g[o_codes + x] = scw[x];
    __syncthreads();
    // part C: header [0 nr, 1 ngr, 2 nvals, 3 o_grow, 4 o_rgrp, 5 o_codes, 6 o_rloc, 7 o_fgrp], fan groups (first value slot |
    // kmax << 16), global row ids in sorted order, row groups (first code word | words per lane << 24), codes
    {
        uint32_t *gC = fcblob + fcoff[t];
        const int o_fg = 8, o_grow = o_fg + pad4(ngf), o_rgrp = o_grow + pad4(nr), o_rc = o_rgrp + pad4(ngr + 1), o_rloc = o_rc + pad4(ncw_r);
        if (tid == 0) {
            gC[0] = nr; gC[1] = ngr; gC[2] = nvals_r; gC[3] = o_grow; gC[4] = o_rgrp; gC[5] = o_rc; gC[6] = o_rloc; gC[7] = o_fg;
        }
        // local row of every sorted row (mass forms: the diagonal needs the row's entries and its star's measure together)
        for (int e = tid; e < pad4(nr); e += TB_THREADS) gC[o_rloc + e] = e < nr ? (rkey[e] & 0xfffu) : 0u;
        for (int x = tid; x < pad4(ngf); x += TB_THREADS) gC[o_fg + x] = x < ngf ? ((uint32_t)rfg[x] | ((uint32_t)(fgt[x] >> 16) << 16)) : 0u;
        for (int e = tid; e < pad4(nr); e += TB_THREADS) gC[o_grow + e] = e < nr ? (uint32_t)rord[r0 + (rkey[e] & 0xfffu)] : 0u;
        for (int x = tid; x < pad4(ngr + 1); x += TB_THREADS)
            gC[o_rgrp + x] = x < ngr ? ((uint32_t)rgl[x] | ((uint32_t)((rgl[x + 1] - rgl[x]) >> 5) << 24)) : (uint32_t)ncw_r;
        const bool cr_sh = pad4(ncw_r) <= 4096;
        uint32_t *crbuf = cr_sh ? scw : gC + o_rc;
        for (int x = tid; x < pad4(ncw_r); x += TB_THREADS) crbuf[x] = 0u;
        for (int l = tid; l < TR_CAP; l += TB_THREADS) cursor[l] = 0;
        __syncthreads();
        uint16_t *gcr = reinterpret_cast<uint16_t *>(crbuf);
        auto rcode_at = [&](int l, int c) -> uint16_t & {
            const int e = rpos[l];
            return gcr[2 * (rgl[e >> 5] + (c >> 1) * 32 + (e & 31)) + (c & 1)];
        };
        for_vertex_sums([&](uint32_t x, int slot) {
            const int l = s2r[x];
            if (l != 255) rcode_at(l, atomicAdd(&cursor[l], 1)) = (uint16_t)slot;
        });
        __syncthreads();
        for (int l = tid; l < nr; l += TB_THREADS) { // ascending value slots: a fixed summation order
            const int n = rcnt[l];
            for (int x = 1; x < n; ++x) {
                const uint16_t v = rcode_at(l, x);
                int y = x - 1;
                while (y >= 0 && rcode_at(l, y) > v) {
                    rcode_at(l, y + 1) = rcode_at(l, y);
                    --y;
                }
                rcode_at(l, y + 1) = v;
            }
        }
        __syncthreads();
        if (cr_sh)
            for (int x = tid; x < pad4(ncw_r); x += TB_THREADS) gC[o_rc + x] = scw[x];
    }
}

struct FanSmem { // byte offsets of the shared-memory regions (part A of the descriptor at 0)
    int bufB, vals, ent;
    int bufC, wvals; // mass forms: part C of the descriptor and the vertex sums of |det|
};

// MASS: the form has a mass term m u v (heat, config 4): every off-diagonal entry of an element also gets m_o |K| (cmo * det),
// the diagonal (m_d + 3 m_o) times the measure of the row's star - the vertex sums of |det| of the right-hand-side
// kernel (part C of the descriptor), gathered per row next to the row sum of its entries.
template <int THREADS, int MINB, bool MASS>
__global__ void __launch_bounds__(THREADS, MINB) k_asm_fans(const uint32_t *__restrict__ foff, const uint32_t *__restrict__ fhead,
                                                            const uint32_t *__restrict__ fblob, const uint32_t *__restrict__ fcoff,
                                                            const uint32_t *__restrict__ fcblob, int ntiles, double *__restrict__ out,
                                                            int accumulate, double cw, double cmd, double cmo, const FanSmem S)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2]; // [0]: part A arrived, [1]: part B arrived
    double *sV = reinterpret_cast<double *>(smem_raw + S.vals); // values of the fans (slot 0 = 0)
    double *sE = reinterpret_cast<double *>(smem_raw + S.ent);  // off-diagonal sums of the tile's entries, CSR order
    const uint32_t *sA = reinterpret_cast<const uint32_t *>(smem_raw);
    const uint32_t *sB = reinterpret_cast<const uint32_t *>(smem_raw + S.bufB);
    const uint32_t *sC = reinterpret_cast<const uint32_t *>(smem_raw + S.bufC);
    double *sW = reinterpret_cast<double *>(smem_raw + S.wvals);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    tile_mbar_init(mbar);
    auto issueA = [&](int t) {
        const uint32_t bytes = __ldg(fhead + t) * 4u;
        tile_expect(smem_u32(&mbar[0]), bytes);
        tile_bulk(smem_u32(smem_raw), fblob + __ldg(foff + t), bytes, smem_u32(&mbar[0]));
    };
    auto issueB = [&](int t) {
        const uint32_t w0 = __ldg(foff + t), wa = __ldg(fhead + t), bytes = (__ldg(foff + t + 1) - w0 - wa) * 4u;
        uint32_t c0 = 0, cbytes = 0;
        if (MASS) {
            c0 = __ldg(fcoff + t);
            cbytes = (__ldg(fcoff + t + 1) - c0) * 4u;
        }
        tile_expect(smem_u32(&mbar[1]), bytes + cbytes);
        tile_bulk(smem_u32(smem_raw + S.bufB), fblob + w0 + wa, bytes, smem_u32(&mbar[1]));
        if (MASS) tile_bulk(smem_u32(smem_raw + S.bufC), fcblob + c0, cbytes, smem_u32(&mbar[1]));
    };
    int t = blockIdx.x;
    uint32_t ph = 0;
    if (tid == 0) {
        sV[0] = 0.0;
        if (MASS) sW[0] = 0.0;
        if (t < ntiles) {
            issueA(t);
            issueB(t);
        }
    }
    for (; t < ntiles; t += gridDim.x, ph ^= 1) {
        tile_wait(smem_u32(&mbar[0]), ph);
        if (MASS) tile_wait(smem_u32(&mbar[1]), ph); // (the fan groups of the vertex sums are in part C)
        {
            const int nfan = sA[2], ngf = sA[4];
            const uint32_t *fgrp = sA + sA[7];
            const double *coord = reinterpret_cast<const double *>(sA + sA[8]);
            const uint4 *fans = reinterpret_cast<const uint4 *>(sA + sA[9]);
            // ---- fans: every element of the tile once
            for (int G = warp; G < ngf; G += NW) {
                const int f = G * 32 + lane;
                uint4 fw = make_uint4(0u, 0u, 0u, 0u);
                if (f < nfan) fw = fans[f];
                const uint32_t fg = fgrp[G];
                const int kmax = (int)(fg >> 16), K1 = kmax + 1;
                const int k = (int)((fw.z >> 16) & 15u);
                double *v = sV + (fg & 0xffffu); // value j of this lane at v[32 j + ((lane + j) & 31)]
                const unsigned long long rr = (unsigned long long)(fw.x >> 16) | ((unsigned long long)fw.y << 16) |
                                              ((unsigned long long)(fw.z & 0xffffu) << 48);
                const double *P = coord + 3 * (fw.x & 255u), *Q = coord + 3 * ((fw.x >> 8) & 255u), *R = coord + 3 * ((uint32_t)rr & 255u);
                const double px = P[0], py = P[1], pz = P[2];
                const double ax = Q[0] - px, ay = Q[1] - py, az = Q[2] - pz;
                double bx = R[0] - px, by = R[1] - py, bz = R[2] - pz;
                double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx; // a x b
                double accA = 0.0, carp = 0.0, carq = 0.0;
                double wtot = 0.0, wprev = 0.0; // MASS: sum of |det| over the fan, |det| of the previous element
                double *wv = MASS ? sW + ((sC + sC[7])[G] & 0xffffu) : nullptr;
                int jp = 1, jq = 1 + K1, jr = 1 + 2 * K1;
                // the coordinates of the next ring vertex are fetched one step ahead (slot 0 past the end: harmless)
                const double *R1 = coord + 3 * ((uint32_t)(rr >> 8) & 255u);
                double nx = R1[0], ny = R1[1], nz = R1[2];
#pragma unroll 2
                for (int tt = 0; tt < kmax; ++tt) {
                    double K01 = 0.0, K02 = 0.0, K03 = 0.0, K12 = 0.0, K13 = 0.0, K23 = 0.0;
                    const double ex = nx - px, ey = ny - py, ez = nz - pz; // the new ring vertex
                    {
                        const double *R2 = coord + 3 * ((uint32_t)(rr >> (8 * min(tt + 2, 7))) & 255u);
                        nx = R2[0]; ny = R2[1]; nz = R2[2];
                    }
                    if (tt < k) {
                        const double fx = ay * ez - az * ey, fy = az * ex - ax * ez, fz = ax * ey - ay * ex; // a x e
                        // element (p, q, r_t, r_t+1): N1 = b x e, N2 = e x a = -(a x e), N3 = a x b, N0 = -(N1 + N2 + N3)
                        const double n1x = by * ez - bz * ey, n1y = bz * ex - bx * ez, n1z = bx * ey - by * ex;
                        const double det = ax * n1x + ay * n1y + az * n1z;
                        const double n0x = fx - n1x - cx, n0y = fy - n1y - cy, n0z = fz - n1z - cz;
                        const double adet = fabs(det);
                        const double s = cw * tile_rcp(adet);
                        K01 = s * (n0x * n1x + n0y * n1y + n0z * n1z);
                        K02 = -s * (n0x * fx + n0y * fy + n0z * fz);
                        K03 = s * (n0x * cx + n0y * cy + n0z * cz);
                        K12 = -s * (n1x * fx + n1y * fy + n1z * fz);
                        K13 = s * (n1x * cx + n1y * cy + n1z * cz);
                        K23 = -s * (fx * cx + fy * cy + fz * cz);
                        bx = ex; by = ey; bz = ez;
                        cx = fx; cy = fy; cz = fz;
                        if (MASS) {
                            const double mo = cmo * adet;
                            K01 += mo; K02 += mo; K03 += mo; K12 += mo; K13 += mo; K23 += mo;
                            wtot += adet;
                            wv[32 * (1 + tt) + ((lane + 1 + tt) & 31)] = wprev + adet; // ring vertex t: elements t-1 and t
                            wprev = adet;
                        }
                    } else if (MASS) {
                        wv[32 * (1 + tt) + ((lane + 1 + tt) & 31)] = wprev;
                        wprev = 0.0;
                    }
                    accA += K01;
                    v[32 * jp + ((lane + jp) & 31)] = carp + K02; // spoke p - r_t: elements t-1 and t
                    carp = K03;
                    v[32 * jq + ((lane + jq) & 31)] = carq + K12; // spoke q - r_t
                    carq = K13;
                    v[32 * jr + ((lane + jr) & 31)] = K23;        // ring edge r_t - r_t+1
                    ++jp; ++jq; ++jr;
                }
                v[32 * jp + ((lane + jp) & 31)] = carp;
                v[32 * jq + ((lane + jq) & 31)] = carq;
                v[lane] = accA;
                if (MASS) {
                    wv[32 * (1 + kmax) + ((lane + 1 + kmax) & 31)] = wprev;
                    wv[lane] = wtot;
                }
            }
        }
        __syncthreads(); // the values are complete; part A is free
        if (tid == 0 && t + (int)gridDim.x < ntiles) issueA(t + gridDim.x);
        tile_wait(smem_u32(&mbar[1]), ph);
        const int nr = sB[0], nq = sB[1], nge = sB[2];
        const int32_t *gbase = reinterpret_cast<const int32_t *>(sB + sB[3]);
        const uint32_t *rinfo = sB + sB[4];
        const uint32_t *egrp = sB + sB[5];
        const uint32_t *einfo = sB + sB[6];
        const uint32_t *codes = sB + sB[7];
        // ---- entries, sorted by list length: 32 per warp, transposed code lists; the sums go to their place in the CSR
        // rows (global memory) and to sE in CSR order (row sums for the diagonals)
        for (int g = warp; g < nge; g += NW) {
            const int e = g * 32 + lane;
            const uint32_t eg = egrp[g], nk = eg >> 24;
            const uint32_t *cp = codes + (eg & 0xffffffu) + lane;
            double acc = 0.0;
            // (lists are short - 1 to 3 words for most groups: the loads of a group are issued together, then added in
            // list order)
            if (nk == 1) {
                const uint32_t c0 = cp[0];
                const double v0 = sV[c0 & 0xffffu], v1 = sV[c0 >> 16];
                acc = v0 + v1;
            } else if (nk == 2) {
                const uint32_t c0 = cp[0], c1 = cp[32];
                const double v0 = sV[c0 & 0xffffu], v1 = sV[c0 >> 16], v2 = sV[c1 & 0xffffu], v3 = sV[c1 >> 16];
                acc = ((v0 + v1) + v2) + v3;
            } else if (nk == 3) {
                const uint32_t c0 = cp[0], c1 = cp[32], c2 = cp[64];
                const double v0 = sV[c0 & 0xffffu], v1 = sV[c0 >> 16], v2 = sV[c1 & 0xffffu], v3 = sV[c1 >> 16], v4 = sV[c2 & 0xffffu],
                             v5 = sV[c2 >> 16];
                acc = ((((v0 + v1) + v2) + v3) + v4) + v5;
            } else {
#pragma unroll 1
                for (uint32_t kk = 0; kk < nk; ++kk) {
                    const uint32_t c = cp[kk * 32];
                    acc += sV[c & 0xffffu];
                    acc += sV[c >> 16];
                }
            }
            if (e < nq) {
                const uint32_t info = einfo[e];
                sE[(info >> 16) & 0xfffu] = acc;
                if (!(info >> 31)) {
                    double *dst = out + (size_t)gbase[(info >> 8) & 255u] + (info & 255u);
                    *dst = accumulate ? *dst + acc : acc;
                }
            }
        }
        __syncthreads();
        // ---- diagonals: K_ii = -sum_{j != i} K_ij (partition of unity: with a mass term every off-diagonal entry carries
        // m_o |K| per element, 3 per element in the row sum) + (m_d + 3 m_o) * measure of the star
        if (!MASS) {
            for (int l = tid; l < nr; l += THREADS) {
                const uint32_t ri = rinfo[l];
                const int q0 = ri & 0xffffu, L = ri >> 24;
                if (L == 0) continue;
                double sx = 0.0;
                for (int kq = 0; kq < L; ++kq) sx += sE[q0 + kq];
                double *dst = out + (size_t)gbase[l] + ((ri >> 16) & 255u);
                *dst = accumulate ? *dst - sx : -sx;
            }
        } else {
            const int ngr = sC[1];
            const uint32_t *rgrp = sC + sC[4], *rcodes = sC + sC[5], *rloc = sC + sC[6];
            for (int g = warp; g < ngr; g += NW) { // rows in part C's order (sorted by the number of fans around them)
                const int e = g * 32 + lane;
                const uint32_t eg = rgrp[g], nk = eg >> 24;
                const uint32_t *cp = rcodes + (eg & 0xffffffu) + lane;
                double sd = 0.0;
#pragma unroll 2
                for (uint32_t kk = 0; kk < nk; ++kk) {
                    const uint32_t c = cp[kk * 32];
                    sd += sW[c & 0xffffu];
                    sd += sW[c >> 16];
                }
                if (e < nr) {
                    const int l = (int)rloc[e];
                    const uint32_t ri = rinfo[l];
                    const int q0 = ri & 0xffffu, L = ri >> 24;
                    if (L != 0) {
                        double sx = 0.0;
                        for (int kq = 0; kq < L; ++kq) sx += sE[q0 + kq];
                        const double d = (cmd + 3.0 * cmo) * sd - sx;
                        double *dst = out + (size_t)gbase[l] + ((ri >> 16) & 255u);
                        *dst = accumulate ? *dst + d : d;
                    }
                }
            }
        }
        __syncthreads(); // part B (and C), sV, sW and sE are free again
        if (tid == 0 && t + (int)gridDim.x < ntiles) issueB(t + gridDim.x);
    }
}

// Right-hand side of value-only linear forms on the fans: b_i = cval * sum of |det K| over the star of i.  Part A of the
// tile descriptor (fans, coordinates) and part C (rows sorted by the number of fans around them, code lists).
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_rhs_fans(const uint32_t *__restrict__ foff, const uint32_t *__restrict__ fhead,
                                                            const uint32_t *__restrict__ fblob, const uint32_t *__restrict__ fcoff,
                                                            const uint32_t *__restrict__ fcblob, int ntiles, double *__restrict__ bvec,
                                                            int accumulate, double cval, const FanSmem S)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2]; // [0]: part A arrived, [1]: part C arrived
    double *sV = reinterpret_cast<double *>(smem_raw + S.vals);
    const uint32_t *sA = reinterpret_cast<const uint32_t *>(smem_raw);
    const uint32_t *sC = reinterpret_cast<const uint32_t *>(smem_raw + S.bufB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    tile_mbar_init(mbar);
    auto issueA = [&](int t) {
        const uint32_t bytes = __ldg(fhead + t) * 4u;
        tile_expect(smem_u32(&mbar[0]), bytes);
        tile_bulk(smem_u32(smem_raw), fblob + __ldg(foff + t), bytes, smem_u32(&mbar[0]));
    };
    auto issueC = [&](int t) {
        const uint32_t w0 = __ldg(fcoff + t), bytes = (__ldg(fcoff + t + 1) - w0) * 4u;
        tile_expect(smem_u32(&mbar[1]), bytes);
        tile_bulk(smem_u32(smem_raw + S.bufB), fcblob + w0, bytes, smem_u32(&mbar[1]));
    };
    int t = blockIdx.x;
    uint32_t ph = 0;
    if (tid == 0) {
        sV[0] = 0.0;
        if (t < ntiles) {
            issueA(t);
            issueC(t);
        }
    }
    for (; t < ntiles; t += gridDim.x, ph ^= 1) {
        tile_wait(smem_u32(&mbar[0]), ph);
        tile_wait(smem_u32(&mbar[1]), ph); // (the fan groups of the right-hand side are in part C)
        {
            const int nfan = sA[2], ngf = sA[4];
            const double *coord = reinterpret_cast<const double *>(sA + sA[8]);
            const uint4 *fans = reinterpret_cast<const uint4 *>(sA + sA[9]);
            const uint32_t *fgrp = sC + sC[7];
            for (int G = warp; G < ngf; G += NW) {
                const int f = G * 32 + lane;
                uint4 fw = make_uint4(0u, 0u, 0u, 0u);
                if (f < nfan) fw = fans[f];
                const uint32_t fg = fgrp[G];
                const int kmax = (int)(fg >> 16);
                const int k = (int)((fw.z >> 16) & 15u);
                double *v = sV + (fg & 0xffffu);
                const unsigned long long rr = (unsigned long long)(fw.x >> 16) | ((unsigned long long)fw.y << 16) |
                                              ((unsigned long long)(fw.z & 0xffffu) << 48);
                const double *P = coord + 3 * (fw.x & 255u), *Q = coord + 3 * ((fw.x >> 8) & 255u), *R = coord + 3 * ((uint32_t)rr & 255u);
                const double px = P[0], py = P[1], pz = P[2];
                const double ax = Q[0] - px, ay = Q[1] - py, az = Q[2] - pz;
                double bx = R[0] - px, by = R[1] - py, bz = R[2] - pz;
                double tot = 0.0, prev = 0.0;
#pragma unroll 2
                for (int tt = 0; tt < kmax; ++tt) {
                    double d = 0.0;
                    if (tt < k) {
                        const double *R1 = coord + 3 * ((uint32_t)(rr >> (8 * (tt + 1))) & 255u);
                        const double ex = R1[0] - px, ey = R1[1] - py, ez = R1[2] - pz;
                        // det of (a, b, e) = a . (b x e)
                        d = fabs(ax * (by * ez - bz * ey) + ay * (bz * ex - bx * ez) + az * (bx * ey - by * ex));
                        bx = ex; by = ey; bz = ez;
                    }
                    tot += d;
                    v[32 * (1 + tt) + ((lane + 1 + tt) & 31)] = prev + d; // ring vertex t: elements t-1 and t
                    prev = d;
                }
                v[32 * (1 + kmax) + ((lane + 1 + kmax) & 31)] = prev;
                v[lane] = tot;
            }
        }
        __syncthreads();
        if (tid == 0 && t + (int)gridDim.x < ntiles) issueA(t + gridDim.x);
        {
            const int nr = sC[0], ngr = sC[1];
            const int32_t *grow = reinterpret_cast<const int32_t *>(sC + sC[3]);
            const uint32_t *rgrp = sC + sC[4];
            const uint32_t *codes = sC + sC[5];
            for (int g = warp; g < ngr; g += NW) {
                const int e = g * 32 + lane;
                const uint32_t eg = rgrp[g], nk = eg >> 24;
                const uint32_t *cp = codes + (eg & 0xffffffu) + lane;
                double acc = 0.0;
#pragma unroll 2
                for (uint32_t kk = 0; kk < nk; ++kk) {
                    const uint32_t c = cp[kk * 32];
                    acc += sV[c & 0xffffu];
                    acc += sV[c >> 16];
                }
                if (e < nr) {
                    double *dst = bvec + grow[e];
                    const double val = cval * acc;
                    *dst = accumulate ? *dst + val : val;
                }
            }
        }
        __syncthreads();
        if (tid == 0 && t + (int)gridDim.x < ntiles) issueC(t + gridDim.x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host: tile set construction
// ---------------------------------------------------------------------------------------------------------------
size_t build_shmem()
{
    return (size_t)4 * (SORT_CAP + SORT_CAP + NE_CAP + NE_CAP + NV_CAP + TR_CAP * BMW + (TR_CAP + 1) + (NQ_CAP + 1) + (TB_THREADS + 1) +
                        TR_CAP) + NV_CAP + 64;
}

// consecutive Morton blocks packed into tiles of <= tr rows; a block larger than tr is cut
void chunk_rows(const std::vector<uint32_t> &key, int dim, int tr, std::vector<int32_t> &tstart)
{
    const int n = (int)key.size();
    const int nlev = dim == 3 ? 10 : 15;
    std::vector<int64_t> hist(nlev + 1, 0);
    for (int i = 1; i < n; ++i) {
        const uint32_t d = key[i] ^ key[i - 1];
        if (d) hist[(31 - __builtin_clz(d)) / dim]++;
    }
    // nb[l] = number of non-empty blocks at level l (block id = key >> dim*l); the largest l with n/nb <= 1.3 tr
    int lev = 0;
    int64_t nb = 1;
    for (int l = nlev; l >= 0; --l) {
        if (l < nlev) nb += hist[l];
        if ((double)n / (double)nb <= 1.3 * tr) {
            lev = l;
            break;
        }
    }
    const int sh = dim * lev;
    tstart.clear();
    tstart.push_back(0);
    int cur = 0; // rows in the open tile
    int i = 0;
    while (i < n) {
        int j = i + 1;
        const uint32_t b = sh >= 32 ? 0u : key[i] >> sh;
        while (j < n && (sh >= 32 ? 0u : key[j] >> sh) == b) ++j;
        int len = j - i;
        if (cur > 0 && cur + len > tr) { // close the open tile
            tstart.push_back(i);
            cur = 0;
        }
        while (len > tr) { // cut an oversized block
            i += tr;
            len -= tr;
            tstart.push_back(i);
        }
        cur += len;
        i = j;
    }
    if (tstart.back() != n) tstart.push_back(n);
}

void build_fans(ffcuda_ctx *ctx, ffcuda_space *s, const int32_t *nrowptr, const int32_t *d_rord, const int32_t *d_tstart, int ntiles);

void build_tiles(ffcuda_ctx *ctx, ffcuda_space *s, const int32_t *nrowptr)
{
    TileSet &T = s->tiles;
    T.state = -1;
    ffcuda_mesh *m = s->mesh;
    const Incidence &I = s->incidence;
    if (!I.built || !I.ell || s->order != 1 || s->ncomp != 1) return;
    cudaStream_t st = ctx->stream;
    const int dim = m->dim, nrows = s->nnodes_owned;
    if (nrows <= 0) return;
    const int tr = std::max(8, std::min(TR_CAP, ctx->tile_rows));
    // ---- Morton order of the rows' vertices
    DBuf<unsigned long long> box;
    box.alloc(6);
    unsigned long long hbox[6] = {~0ull, ~0ull, ~0ull, 0, 0, 0};
    FF_CUDA(cudaMemcpyAsync(box.p, hbox, sizeof(hbox), cudaMemcpyHostToDevice, st));
    ff_launch(ctx, "tile_bbox", [&] { k_bbox<<<ctx->sm_count * 2, 256, 0, st>>>(m->xyz.p, m->vstride, dim, nrows, box.p); });
    FF_CUDA(ff_memcpy_sync(ctx, hbox, box.p, sizeof(hbox), cudaMemcpyDeviceToHost));
    BoxScale B;
    const double qmax = dim == 3 ? 1024.0 : 32768.0;
    // Morton cells should hold the same number of vertices along every axis: the axes are scaled by the inverse of the
    // mean edge extent along them (a sample of the elements), then one common factor maps the largest scaled extent to the
    // key range.  Isotropic meshes get one scale for all directions (tiles = cubes in space, also on a thin slab of a
    // partitioned cube); cube(128,128,256) gets z stretched by 2 (r01: 0.443 ms instead of 0.380 ms for the same 128^3 cells).
    double hmean[3] = {1.0, 1.0, 1.0};
    {
        DBuf<unsigned long long> acc;
        acc.alloc(3);
        FF_CUDA(cudaMemsetAsync(acc.p, 0, 3 * sizeof(unsigned long long), st));
        const int nsample = std::min(m->nt, 1 << 20);
        double bext = 0.0;
        for (int x = 0; x < dim; ++x) bext = std::max(bext, unord64(hbox[3 + x]) - unord64(hbox[x]));
        const double fix = bext > 0.0 ? 68719476736.0 / bext : 1.0; // 2^36 per box extent: 6 * 2^20 edges stay below 2^63
        ff_launch(ctx, "tile_metric", [&] {
            if (dim == 3) k_edge_extents<4><<<ctx->sm_count * 2, 256, 0, st>>>(m->xyz.p, m->vstride, m->conn.p, nsample, fix, acc.p);
            else k_edge_extents<3><<<ctx->sm_count * 2, 256, 0, st>>>(m->xyz.p, m->vstride, m->conn.p, nsample, fix, acc.p);
        });
        unsigned long long hi[3];
        FF_CUDA(ff_memcpy_sync(ctx, hi, acc.p, sizeof(hi), cudaMemcpyDeviceToHost));
        const double h[3] = {(double)hi[0], (double)hi[1], (double)hi[2]};
        double hmax = 0.0;
        for (int x = 0; x < dim; ++x) hmax = std::max(hmax, h[x]);
        for (int x = 0; x < dim; ++x) hmean[x] = (h[x] > 1e-3 * hmax && hmax > 0.0) ? h[x] / hmax : 1.0; // (degenerate direction: left alone)
    }
    double ext = 0.0;
    for (int x = 0; x < dim; ++x) ext = std::max(ext, (unord64(hbox[3 + x]) - unord64(hbox[x])) / hmean[x]);
    for (int x = 0; x < 3; ++x) {
        B.lo[x] = x < dim ? unord64(hbox[x]) : 0.0;
        B.sc[x] = (x < dim && ext > 0.0) ? qmax * (1.0 - 1e-9) / (ext * hmean[x]) : 0.0;
    }
    DBuf<uint32_t> k0, k1;
    DBuf<int32_t> v0, rord;
    k0.alloc(nrows); k1.alloc(nrows); v0.alloc(nrows); rord.alloc(nrows);
    ff_launch(ctx, "tile_morton", [&] { k_morton<<<ff_blocks(nrows, 256), 256, 0, st>>>(m->xyz.p, m->vstride, dim, nrows, B, k0.p, v0.p); });
    size_t tb = 0;
    FF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, v0.p, rord.p, nrows, 0, 30, st));
    DBuf<unsigned char> tmpbuf;
    tmpbuf.alloc(tb + 16);
    ctx->launches++;
    FF_CUDA(cub::DeviceRadixSort::SortPairs(tmpbuf.p, tb, k0.p, k1.p, v0.p, rord.p, nrows, 0, 30, st));
    std::vector<uint32_t> hkey(nrows);
    FF_CUDA(ff_memcpy_sync(ctx, hkey.data(), k1.p, (size_t)nrows * 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> tstart;
    chunk_rows(hkey, dim, tr, tstart);
    // ---- sizes of every tile; tiles that do not fit are halved
    const size_t shmem = build_shmem();
    const int NV = dim + 1;
    auto kstat = NV == 4 ? k_tile_build<4, 0> : k_tile_build<3, 0>;
    auto kwrite = NV == 4 ? k_tile_build<4, 1> : k_tile_build<3, 1>;
    FF_CUDA(cudaFuncSetAttribute(kstat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    FF_CUDA(cudaFuncSetAttribute(kwrite, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    const IncView V = ff_view(I);
    DBuf<int32_t> d_tstart, d_stats;
    std::vector<int32_t> hst;
    int ntiles = 0;
    for (int round = 0;; ++round) {
        ntiles = (int)tstart.size() - 1;
        d_tstart.alloc(tstart.size());
        d_stats.alloc((size_t)ntiles * 8);
        FF_CUDA(cudaMemcpyAsync(d_tstart.p, tstart.data(), tstart.size() * 4, cudaMemcpyHostToDevice, st));
        ff_launch(ctx, "tile_sizes", [&] {
            kstat<<<ntiles, TB_THREADS, shmem, st>>>(rord.p, d_tstart.p, m->conn.p, V, m->xyz.p, m->vstride, nrowptr, 1, d_stats.p,
                                                     nullptr, nullptr, nullptr, nullptr);
        });
        hst.resize((size_t)ntiles * 8);
        FF_CUDA(ff_memcpy_sync(ctx, hst.data(), d_stats.p, hst.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<int32_t> ns;
        bool split = false;
        for (int t = 0; t < ntiles; ++t) {
            ns.push_back(tstart[t]);
            if (!hst[(size_t)t * 8 + 4]) {
                const int nr = tstart[t + 1] - tstart[t];
                if (nr <= 1 || round >= 8) return; // a single row that does not fit: the thread-per-row kernel keeps the space
                ns.push_back(tstart[t] + nr / 2);
                split = true;
            }
        }
        ns.push_back(tstart[ntiles]);
        if (!split) break;
        tstart.swap(ns);
    }
    // ---- offsets, maxima
    std::vector<uint32_t> htoff((size_t)ntiles + 1), hroff((size_t)ntiles + 1), hpre((size_t)ntiles + 1, 0);
    uint64_t off = 0, roffs = 0;
    T.max_pre = T.max_rwords = 0;
    T.max_rows = T.max_nvt = T.max_nelem = T.max_nq = T.max_ncodes = T.max_words = 0;
    T.sum_nelem = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int32_t *h = &hst[(size_t)t * 8];
        const int nvt = h[0], nelem = h[1], nq = h[2], ncodes = h[3], nr = h[5];
        const int nrec = h[6];
        const int pre = HDR + 2 * pad4(nr) + pad4(nr + 1) + pad4(2 * dim * nvt) + pad4(nelem); // up to the entry words
        const int words = pre + pad4(nq + 1) + pad4((ncodes + 1) / 2);
        const int rwords = 4 + pad4((nr + 2) / 2) + pad4((nrec + 1) / 2);
        htoff[t] = (uint32_t)off;
        hroff[t] = (uint32_t)roffs;
        hpre[t] = (uint32_t)pre;
        off += (uint64_t)words;
        roffs += (uint64_t)rwords;
        T.max_pre = std::max(T.max_pre, pre);
        T.max_rwords = std::max(T.max_rwords, rwords);
        T.max_rows = std::max(T.max_rows, nr); T.max_nvt = std::max(T.max_nvt, nvt); T.max_nelem = std::max(T.max_nelem, nelem);
        T.max_nq = std::max(T.max_nq, nq); T.max_ncodes = std::max(T.max_ncodes, ncodes); T.max_words = std::max(T.max_words, words);
        T.sum_nelem += nelem;
    }
    if (off >= ((uint64_t)1 << 32) || roffs >= ((uint64_t)1 << 32)) return;
    // 3-D spaces with fans: the stiffness / heat matrices and the value-only right-hand sides run on the fan descriptors;
    // the element-by-element descriptors (0.8 GB and 17 ms at cube(128)) are not built
    T.has_blob = true;
    if (dim == 3 && ctx->tile_fans != 0 && ctx->tile_policy != 2) {
        build_fans(ctx, s, nrowptr, rord.p, d_tstart.p, ntiles);
        if (T.fan_state == 1) {
            T.has_blob = false;
            T.nes = T.max_nelem | 1;
            T.tr = tr;
            T.ntiles = ntiles;
            T.state = 1;
            return;
        }
    }
    hroff[ntiles] = (uint32_t)roffs;
    T.nes = T.max_nelem | 1; // odd stride of the value table
    if ((DIM_PAIRS(dim) + 1) * T.nes > 65535) return;
    htoff[ntiles] = (uint32_t)off;
    T.toff.alloc((size_t)ntiles + 1);
    T.roff.alloc((size_t)ntiles + 1);
    T.tpre.alloc((size_t)ntiles + 1);
    T.blob.alloc((size_t)off + 4);
    T.rblob.alloc((size_t)roffs + 4);
    FF_CUDA(cudaMemcpyAsync(T.toff.p, htoff.data(), htoff.size() * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(T.roff.p, hroff.data(), hroff.size() * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(T.tpre.p, hpre.data(), hpre.size() * 4, cudaMemcpyHostToDevice, st));
    ff_launch(ctx, "tile_build", [&] {
        kwrite<<<ntiles, TB_THREADS, shmem, st>>>(rord.p, d_tstart.p, m->conn.p, V, m->xyz.p, m->vstride, nrowptr, T.nes, nullptr,
                                                  T.toff.p, T.blob.p, T.roff.p, T.rblob.p);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // htoff / tstart are host vectors
    T.tr = tr;
    T.ntiles = ntiles;
    T.state = 1;
    if (T.fan_state == 0) build_fans(ctx, s, nrowptr, rord.p, d_tstart.p, ntiles);
    if (getenv("FFCUDA_VERBOSE"))
        fprintf(stderr, "ffcuda tiles: %d tiles of <= %d rows, max rows %d vertices %d elements %d entries %d codes %d, "
                        "%.2f evaluations per element, blobs %.1f + %.1f MB\n",
                ntiles, tr, T.max_rows, T.max_nvt, T.max_nelem, T.max_nq, T.max_ncodes, (double)T.sum_nelem / std::max(1, m->nt),
                off * 4.0 / 1e6, roffs * 4.0 / 1e6);
}

// fan set of a 3-D scalar P1 space on the tiles of the tile set (same rows per tile); T.fan_state = 1 when every tile fits
void build_fans(ffcuda_ctx *ctx, ffcuda_space *s, const int32_t *nrowptr, const int32_t *d_rord, const int32_t *d_tstart, int ntiles)
{
    TileSet &T = s->tiles;
    T.fan_state = -1;
    ffcuda_mesh *m = s->mesh;
    if (m->dim != 3 || ctx->tile_fans == 0) return;
    cudaStream_t st = ctx->stream;
    const size_t shmem = (size_t)4 * (build_words() + 2 * 4096 + 3 * TR_CAP + 16 + 64 + 4096);
    FF_CUDA(cudaFuncSetAttribute(k_fan_build<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    FF_CUDA(cudaFuncSetAttribute(k_fan_build<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    const IncView V = ff_view(s->incidence);
    DBuf<int32_t> d_stats;
    d_stats.alloc((size_t)ntiles * 12);
    ff_launch(ctx, "fan_sizes", [&] {
        k_fan_build<0><<<ntiles, TB_THREADS, shmem, st>>>(d_rord, d_tstart, m->conn.p, V, m->xyz.p, m->vstride, nrowptr, d_stats.p, nullptr, nullptr,
                                                          nullptr, nullptr);
    });
    std::vector<int32_t> hst((size_t)ntiles * 12);
    FF_CUDA(ff_memcpy_sync(ctx, hst.data(), d_stats.p, hst.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> hoff((size_t)ntiles + 1), hhead((size_t)ntiles + 1, 0), hcoff((size_t)ntiles + 1);
    uint64_t off = 0, coff = 0;
    T.fan_max_c = T.fan_max_rvals = 0;
    T.fan_max_head = T.fan_max_b = T.fan_max_nvals = T.fan_max_nq = 0;
    T.fan_sum_fans = 0;
    int64_t sum_vals = 0, sum_cw = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int32_t *h = &hst[(size_t)t * 12];
        const int nvt = h[0], nq = h[2], nfan = h[3], fit = h[4], nr = h[5], nvals = h[6], ncw = h[7], ngf = h[8], nge = h[9];
        if (!fit) return; // some tile exceeds a capacity: the round-1 tile kernel keeps the space
        const int head = FHDR + pad4(ngf) + pad4(6 * nvt) + 4 * nfan;                               // part A
        const int wordsB = 8 + 2 * pad4(nr) + pad4(nge + 1) + pad4(nq) + pad4(ncw);                     // part B
        const int words = head + wordsB;
        T.fan_max_b = std::max(T.fan_max_b, wordsB);
        hoff[t] = (uint32_t)off;
        hhead[t] = (uint32_t)head;
        off += (uint64_t)words;
        const int wordsC = 8 + pad4(ngf) + 2 * pad4(nr) + pad4(((nr + 31) >> 5) + 1) + pad4(h[11]);
        hcoff[t] = (uint32_t)coff;
        coff += (uint64_t)wordsC;
        T.fan_max_c = std::max(T.fan_max_c, wordsC);
        T.fan_max_rvals = std::max(T.fan_max_rvals, (int)h[10]);
        T.fan_max_head = std::max(T.fan_max_head, head);
        T.fan_max_nvals = std::max(T.fan_max_nvals, nvals);
        T.fan_max_nq = std::max(T.fan_max_nq, nq);
        T.fan_sum_fans += nfan;
        sum_vals += nvals;
        sum_cw += ncw;
    }
    if (off >= ((uint64_t)1 << 32) || coff >= ((uint64_t)1 << 32)) return;
    hoff[ntiles] = (uint32_t)off;
    hcoff[ntiles] = (uint32_t)coff;
    T.fcoff.alloc((size_t)ntiles + 1);
    T.fcblob.alloc((size_t)coff + 4);
    FF_CUDA(cudaMemcpyAsync(T.fcoff.p, hcoff.data(), hcoff.size() * 4, cudaMemcpyHostToDevice, st));
    T.foff.alloc((size_t)ntiles + 1);
    T.fhead.alloc((size_t)ntiles + 1);
    T.fblob.alloc((size_t)off + 4);
    FF_CUDA(cudaMemcpyAsync(T.foff.p, hoff.data(), hoff.size() * 4, cudaMemcpyHostToDevice, st));
    FF_CUDA(cudaMemcpyAsync(T.fhead.p, hhead.data(), hhead.size() * 4, cudaMemcpyHostToDevice, st));
    ff_launch(ctx, "fan_build", [&] {
        k_fan_build<1><<<ntiles, TB_THREADS, shmem, st>>>(d_rord, d_tstart, m->conn.p, V, m->xyz.p, m->vstride, nrowptr, nullptr, T.foff.p, T.fblob.p,
                                                          T.fcoff.p, T.fcblob.p);
    });
    FF_CUDA(cudaStreamSynchronize(st)); // hoff / hhead are host vectors
    T.fan_state = 1;
    if (getenv("FFCUDA_VERBOSE"))
        fprintf(stderr, "ffcuda fans: %lld fans for %lld element evaluations (%.2f per fan), %.2f values and %.2f code words per row, "
                        "max part A %d words, part B %d words, max values %d, blob %.1f MB\n",
                (long long)T.fan_sum_fans, (long long)T.sum_nelem, (double)T.sum_nelem / std::max<int64_t>(1, T.fan_sum_fans),
                (double)sum_vals / std::max(1, s->nnodes_owned), (double)sum_cw / std::max(1, s->nnodes_owned), T.fan_max_head, T.fan_max_b,
                T.fan_max_nvals, off * 4.0 / 1e6);
}

} // namespace

namespace {
// tile set of the space, built on demand; false when the tile path does not apply (policy, capacities)
bool tiles_ready(ffcuda_ctx *ctx, ffcuda_space *s, ffcuda_pattern *P)
{
    TileSet &T = s->tiles;
    if (ctx->tile_policy == 0 || T.state < 0) return false;
    if (T.state == 0) {
        if (!P) return false; // the row pointers of a pattern are baked in: the first matrix assembly builds the set
        if (ctx->tile_policy == 1 && s->lean_assemblies < 2) return false;
        build_tiles(ctx, s, P->nrowptr.p); // every pattern of a fespace has the same row pointers
        if (T.state != 1) return false;
        T.nnz_node = P->nnz_node;
    }
    return true;
}

template <class K>
void tile_launch(ffcuda_ctx *ctx, const char *name, K kern, int threads, size_t shmem, int ntiles, const std::function<void(int)> &go)
{
    FF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    int per_sm = 1;
    FF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, shmem));
    const int grid = std::max(1, std::min(ntiles, std::max(1, per_sm) * ctx->sm_count));
    ff_launch(ctx, name, [&] { go(grid); });
}
} // namespace

bool ff_asm_p1_tiles(ffcuda_ctx *ctx, ffcuda_matrix *A, ffcuda_space *s, double cw, double cmd, double cmo, int accumulate)
{
    TileSet &T = s->tiles;
    ffcuda_pattern *P = A->pattern;
    s->lean_assemblies++;
    // (forms with a mass term: the diagonal needs the measure of every row's star, taken from the record lists of the
    // second descriptor blob: 0.45 ms by tiles against 0.48 ms by rows on cube(128) - and the symbolic phase of a space
    // with tiles is the cheap fused one)
    if (!tiles_ready(ctx, s, P)) return false;
    FF_REQUIRE(T.nnz_node == P->nnz_node, "internal: tile set and pattern disagree");
    ffcuda_mesh *m = s->mesh;
    const int dim = m->dim;
    const bool mass = (cmd != 0.0 || cmo != 0.0);
    if (dim == 3 && T.fan_state == 1 && ctx->tile_fans != 0) {
        FanSmem FS;
        size_t o = ((size_t)T.fan_max_head * 4 + 127) & ~(size_t)127;
        FS.bufB = (int)o;
        o += ((size_t)T.fan_max_b * 4 + 127) & ~(size_t)127;
        FS.vals = (int)o; // 128-byte aligned: the bank of a value is its slot mod 16 (the build kernel orders the lists by it)
        o += ((size_t)T.fan_max_nvals * 8 + 127) & ~(size_t)127;
        FS.ent = (int)o;
        o += ((size_t)(T.fan_max_nq + 1) * 8 + 127) & ~(size_t)127;
        FS.bufC = FS.wvals = 0;
        if (mass) {
            FS.bufC = (int)o;
            o += ((size_t)T.fan_max_c * 4 + 127) & ~(size_t)127;
            FS.wvals = (int)o;
            o += ((size_t)T.fan_max_rvals * 8 + 127) & ~(size_t)127;
        }
        const size_t shmem = o;
        if (shmem <= 200 * 1024) {
            int threads = 128;
            if (const char *e = getenv("FFCUDA_FAN_THREADS")) threads = atoi(e);
            auto runf = [&](auto kern, int thr) {
                tile_launch(ctx, "asm_rows_p1", kern, thr, shmem, T.ntiles, [&](int grid) {
                    kern<<<grid, thr, shmem, ctx->stream>>>(T.foff.p, T.fhead.p, T.fblob.p, T.fcoff.p, T.fcblob.p, T.ntiles, A->vals.p, accumulate,
                                                            cw, cmd, cmo, FS);
                });
            };
            if (mass) runf(k_asm_fans<128, 3, true>, 128);
            else if (threads == 256) runf(k_asm_fans<256, 2, false>, 256);
            else if (threads == 64) runf(k_asm_fans<64, 8, false>, 64);
            else runf(k_asm_fans<128, 5, false>, 128);
            return true;
        }
    }
    if (!T.has_blob) return false; // (no element-by-element descriptors on this space: the thread-per-row kernel takes the form)
    const int NP = dim * (dim + 1) / 2;
    TileSmem S;
    size_t o = ((size_t)T.max_words * 4 + 127) & ~(size_t)127;
    S.buf1 = (int)o;
    o *= 2;
    S.rb0 = S.rb1 = (int)o;
    if (mass) {
        const size_t rb = ((size_t)T.max_rwords * 4 + 127) & ~(size_t)127;
        S.rb1 = (int)(o + rb);
        o += 2 * rb;
    }
    S.ent = (int)o;
    o += (size_t)(T.max_nq + 1) * 8;
    S.nes = T.nes;
    o = (o + 127) & ~(size_t)127; // the bank of a value is its index mod 16 (the build kernel orders the lists by it)
    S.vals = (int)o;
    o += (size_t)(NP + (mass ? 1 : 0)) * S.nes * 8;
    const size_t shmem = o;
    if (shmem > 200 * 1024) return false;
    int threads = 512;
    if (const char *e = getenv("FFCUDA_TILE_THREADS")) threads = std::max(32, std::min(512, atoi(e) & ~31));
    auto run = [&](auto kern) {
        tile_launch(ctx, "asm_rows_p1", kern, threads, shmem, T.ntiles, [&](int grid) {
            kern<<<grid, threads, shmem, ctx->stream>>>(T.toff.p, T.blob.p, T.roff.p, T.rblob.p, T.ntiles, A->vals.p, accumulate, cw, cmd,
                                                        cmo, S);
        });
    };
    if (dim == 3) {
        if (mass) run(k_asm_tiles<3, true>);
        else run(k_asm_tiles<3, false>);
    } else {
        if (mass) run(k_asm_tiles<2, true>);
        else run(k_asm_tiles<2, false>);
    }
    return true;
}

// b (+)= sum over the stars: cval[c] * sum det + sum_x cgrad[c][x] * sum N_i[x]; only on spaces whose tile set exists
bool ff_rhs_p1_tiles(ffcuda_ctx *ctx, ffcuda_vec *b, ffcuda_space *s, const double *cval, const double *cgrad /* [nc][3] */, int hasgrad,
                     int accumulate)
{
    TileSet &T = s->tiles;
    if (ctx->tile_policy == 0 || T.state != 1) return false;
    if (!hasgrad && s->ncomp == 1 && s->mesh->dim == 3 && T.fan_state == 1 && ctx->tile_fans != 0) {
        FanSmem FS;
        size_t o = ((size_t)T.fan_max_head * 4 + 127) & ~(size_t)127;
        FS.bufB = (int)o;
        o += ((size_t)T.fan_max_c * 4 + 127) & ~(size_t)127;
        FS.vals = (int)o;
        o += ((size_t)T.fan_max_rvals * 8 + 127) & ~(size_t)127;
        FS.ent = (int)o;
        const size_t shmem = o;
        if (shmem <= 200 * 1024) {
            auto kern = k_rhs_fans<128, 6>;
            tile_launch(ctx, "rhs_rows", kern, 128, shmem, T.ntiles, [&](int grid) {
                kern<<<grid, 128, shmem, ctx->stream>>>(T.foff.p, T.fhead.p, T.fblob.p, T.fcoff.p, T.fcblob.p, T.ntiles, b->d.p, accumulate,
                                                        cval[0], FS);
            });
            return true;
        }
    }
    // gradient terms move 12 more values per element through shared memory: measured slower than the thread-per-row
    // kernel (350 vs 231 us on cube(128)); value-only forms: 162 vs 217 us
    if (hasgrad && ctx->tile_policy != 2) return false;
    if (!T.has_blob) return false;
    const int dim = s->mesh->dim, nc = s->ncomp;
    RhsCoef C;
    memset(&C, 0, sizeof(C));
    for (int c = 0; c < nc; ++c) {
        C.cval[c] = cval[c];
        for (int x = 0; x < 3; ++x) C.cgrad[c][x] = cgrad[c * 3 + x];
    }
    TileSmem S;
    size_t o = ((size_t)T.max_pre * 4 + 127) & ~(size_t)127;
    S.buf1 = (int)o;
    o *= 2;
    const size_t rb = ((size_t)T.max_rwords * 4 + 127) & ~(size_t)127;
    S.rb0 = (int)o;
    S.rb1 = (int)(o + rb);
    o += 2 * rb;
    S.ent = (int)o;
    S.nes = T.nes;
    S.vals = (int)o;
    o += (size_t)(1 + (hasgrad ? (dim + 1) * dim : 0)) * S.nes * 8;
    const size_t shmem = o;
    if (shmem > 200 * 1024) return false;
    int threads = 128;
    if (const char *e = getenv("FFCUDA_RHS_THREADS")) threads = std::max(32, std::min(512, atoi(e) & ~31));
    auto run = [&](auto kern) {
        tile_launch(ctx, "rhs_rows", kern, threads, shmem, T.ntiles, [&](int grid) {
            kern<<<grid, threads, shmem, ctx->stream>>>(T.toff.p, T.tpre.p, T.blob.p, T.roff.p, T.rblob.p, T.ntiles, b->d.p, nc, accumulate,
                                                        C, S);
        });
    };
    if (dim == 3) {
        if (hasgrad) run(k_rhs_tiles<3, true>);
        else run(k_rhs_tiles<3, false>);
    } else {
        if (hasgrad) run(k_rhs_tiles<2, true>);
        else run(k_rhs_tiles<2, false>);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// P2 row order (used by launch_p2 in assemble.cu): node rows by decreasing length, split into long / short
// ---------------------------------------------------------------------------------------------------------------
namespace {
__global__ void k_row_len_keys(const int32_t *__restrict__ nrowptr, int n, uint32_t *__restrict__ key, int32_t *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (uint32_t)(nrowptr[i + 1] - nrowptr[i]);
    val[i] = i;
}
// keys sorted in decreasing order: out[0] = number of keys > T, out[1] = the largest key <= T (0 if none)
__global__ void k_split_desc(const uint32_t *__restrict__ key, int n, uint32_t T, int32_t *__restrict__ out)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key[mid] > T) lo = mid + 1;
        else hi = mid;
    }
    out[0] = lo;
    out[1] = lo < n ? (int32_t)key[lo] : 0;
}
} // namespace

void ff_p2_row_order(ffcuda_ctx *ctx, ffcuda_space *s, const int32_t *nrowptr, int nrows, int maxrow)
{
    if (s->p2_rowperm.p) return;
    cudaStream_t st = ctx->stream;
    DBuf<uint32_t> k0, k1;
    DBuf<int32_t> v0, d_out;
    k0.alloc(nrows); k1.alloc(nrows); v0.alloc(nrows); d_out.alloc(2);
    s->p2_rowperm.alloc(nrows);
    ff_launch(ctx, "p2_row_keys", [&] { k_row_len_keys<<<ff_blocks(nrows, 256), 256, 0, st>>>(nrowptr, nrows, k0.p, v0.p); });
    size_t tb = 0;
    FF_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, k0.p, k1.p, v0.p, s->p2_rowperm.p, nrows, 0, 32, st));
    DBuf<unsigned char> tmpbuf;
    tmpbuf.alloc(tb + 16);
    ctx->launches++;
    FF_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmpbuf.p, tb, k0.p, k1.p, v0.p, s->p2_rowperm.p, nrows, 0, 32, st));
    // long rows: more than 60 % of the longest (vertex nodes against edge nodes on tetrahedra / triangles)
    ff_launch(ctx, "p2_row_split", [&] { k_split_desc<<<1, 1, 0, st>>>(k1.p, nrows, (uint32_t)(0.6 * maxrow), d_out.p); });
    int32_t h[2] = {0, 0};
    FF_CUDA(ff_memcpy_sync(ctx, h, d_out.p, sizeof(h), cudaMemcpyDeviceToHost));
    s->p2_nlong = h[0];
    s->p2_short_maxrow = std::max(1, h[1]);
}
