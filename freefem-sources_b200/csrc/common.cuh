// common.cuh — internal definitions shared by the CUDA translation units of libffcuda_core.so.
// Nothing here is part of the public ABI (include/ffcuda.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/ffcuda.h"

struct FFError : std::runtime_error {
    explicit FFError(const std::string &s) : std::runtime_error(s) {}
};

#define FF_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            throw FFError(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ +   \
                          ":" + std::to_string(__LINE__) + " (" #call ")");                             \
    } while (0)

#define FF_REQUIRE(cond, msg)                                                                           \
    do {                                                                                                \
        if (!(cond)) throw FFError(std::string("ffcuda: ") + (msg));                                    \
    } while (0)

void ff_set_thread_error(const std::string &s);

// every extern "C" entry point wraps its body with these: no exception crosses the ABI
#define FF_API_BEGIN try {
#define FF_API_END(ctxexpr)                                                                             \
    }                                                                                                   \
    catch (const std::exception &e)                                                                     \
    {                                                                                                   \
        ff_report_error((ctxexpr), e.what());                                                           \
        return 1;                                                                                       \
    }                                                                                                   \
    catch (...)                                                                                         \
    {                                                                                                   \
        ff_report_error((ctxexpr), "unknown exception");                                                \
        return 1;                                                                                       \
    }                                                                                                   \
    return 0;

struct ProfEntry {
    double ms = 0;
    int64_t count = 0;
};
struct ProfPending {
    std::string name;
    cudaEvent_t e0, e1;
};

struct ffcuda_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr; // device -> host copies that overlap the work of `stream` (ffcuda_pattern_download_async)
    cudaEvent_t copy_event = nullptr;
    std::string err;
    bool prof = false;
    std::map<std::string, ProfEntry> prof_acc;
    std::vector<ProfPending> prof_pending;
    int64_t launches = 0;
    int sm_count = 148;
    int tile_policy = 1;        // 0: never use row tiles, 1: from the second assembly on a fespace, 2: always
    int tile_rows = 96;         // rows per tile
    int tile_fans = 1;          // 3-D stiffness forms: elements of a tile evaluated in fans around their longest edge (0: element by element)
    double last_cg_eps2 = 0.0;  // stopping threshold of the last CG solve on this context (ffcuda_cg_stop_threshold)
    int gmres_coop = 1;         // 1: one cooperative kernel per Arnoldi step when the vectors fit its registers, 0: one kernel per basis vector
    // reduction scratch (device) + pinned host mirror
    double *d_scal = nullptr;   // small array of device scalars
    double *h_scal = nullptr;   // pinned
    double *d_partial = nullptr;
    size_t partial_cap = 0;
    // multi-GPU
    int rank = 0, nranks = 1;
    void *nccl_comm = nullptr;
    // peer mailboxes (comm.cu): one buffer per rank, mapped into every process of the box through CUDA IPC; the
    // all-reduces and halo exchanges of CG are plain stores into the peers' mailboxes over NVLink, no NCCL call
    bool p2p = false;
    void *p2p_peer[16] = {};              // p2p_peer[r]: rank r's mailbox in this process' address space
    bool p2p_inproc[16] = {};             // rank r lives in this process (one thread per GPU): its pointer is used as it is
    size_t p2p_halo_cap = 0;              // doubles per (direction, parity) halo region
    unsigned long long p2p_seq_halo = 0;
    // lifetime: every handle created on the context holds a reference; ffcuda_ctx_destroy only marks the context
    // closed and the last handle to go tears it down
    int refs = 0;
    bool closed = false;
    // caching device allocator (ctx.cu): freed blocks are kept and handed out again, cudaMalloc/cudaFree (which
    // synchronise the device) are off the hot path.  Safe because all the work of a context is ordered on one stream.
    std::multimap<size_t, void *> pool_free;
    std::map<void *, size_t> pool_live;
    size_t pool_cached = 0, pool_total = 0;
};
// a blocking copy ORDERED ON THE CONTEXT'S STREAM (plain cudaMemcpy runs on the legacy stream, which does not
// synchronise with the non-blocking stream of the context: with the caching allocator a block can be handed out again
// while earlier kernels that used it are still in flight on the context's stream)
static inline cudaError_t ff_memcpy_sync(ffcuda_ctx *ctx, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ctx->stream);
}
void *ff_pool_alloc(ffcuda_ctx *ctx, size_t bytes);
void ff_pool_free(ffcuda_ctx *ctx, void *p);
void ff_pool_trim(ffcuda_ctx *ctx);
void ff_ctx_unref(ffcuda_ctx *ctx);
struct CtxRef { // first member of every handle struct: released after the handle's device buffers
    ffcuda_ctx *c = nullptr;
    void set(ffcuda_ctx *x)
    {
        c = x;
        if (c) c->refs++;
    }
    ~CtxRef()
    {
        if (c) ff_ctx_unref(c);
    }
};

// ---- peer mailboxes (comm.cu): layout and the device-side descriptor -------------------------------------
//   [RFLAG] u64[2][16]     sequence numbers of the all-reduce contributions (parity, rank)
//   [RDATA] f64[2][16][4]  contributions
//   [HFLAG] u64[16][2]     sequence numbers of the halo layers (source rank, parity); [HCNT] block counter
//   [HDATA] f64[16][2][cap] halo layers (source rank, parity)
constexpr size_t FF_P2P_RFLAG = 0, FF_P2P_RDATA = 4096, FF_P2P_HFLAG = 8192, FF_P2P_HCNT = 8192 + 256, FF_P2P_HDATA = 16384;
constexpr int FF_P2P_MAXR = 16;
// lives in device memory at ctx->d_scal + FF_P2P_DESC_OFF (doubles), i.e. FF_P2P_DESC_OFF - 32 doubles behind the CG flags
constexpr int FF_P2P_DESC_OFF = 64;
constexpr long long FF_P2P_SPIN_LIMIT = 60000000000ll; // clock cycles (~30 s: ranks may legitimately arrive seconds apart)
struct P2PDesc {
    unsigned char *peer[FF_P2P_MAXR]; // rank r's mailbox in this process' address space
    int rank, nranks;
    int fused;                        // the dot products of the running CG are all-reduced by their own kernels
    int timed_out;                    // a spin on a peer's sequence number gave up: the solve reports an error
    unsigned long long seq;           // number of all-reduces so far: advanced on the device, in lockstep on all ranks
};
#ifdef __CUDACC__
__device__ __forceinline__ void ff_st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ff_ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ff_ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// One warp (all 32 lanes): all-reduce of t[0..NV) (lane 0's values are used) over the ranks through the mailboxes;
// the result is returned in every lane.  Sum in rank order: bit-identical on every rank.
template <int NV>
__device__ __forceinline__ void ff_p2p_allreduce_warp(P2PDesc *D, double (&t)[NV], bool op_max)
{
    const int lane = threadIdx.x & 31;
    unsigned long long seq = 0;
    if (lane == 0) seq = ++D->seq;
    seq = __shfl_sync(0xffffffffu, seq, 0);
#pragma unroll
    for (int k = 0; k < NV; ++k) t[k] = __shfl_sync(0xffffffffu, t[k], 0);
    const int par = (int)(seq & 1ull), rank = D->rank, nranks = D->nranks;
    if (lane < nranks) {
        double *dst = reinterpret_cast<double *>(D->peer[lane] + FF_P2P_RDATA) + (size_t)(par * FF_P2P_MAXR + rank) * 4;
#pragma unroll
        for (int k = 0; k < NV; ++k) dst[k] = t[k];
        __threadfence_system();
        ff_st_release_sys(reinterpret_cast<unsigned long long *>(D->peer[lane] + FF_P2P_RFLAG) + par * FF_P2P_MAXR + rank, seq);
        const unsigned long long *mine =
            reinterpret_cast<const unsigned long long *>(D->peer[rank] + FF_P2P_RFLAG) + par * FF_P2P_MAXR + lane;
        // a peer that never arrives (crashed rank) must not hang the device: give up after ~30 s and flag it
        const long long t0 = clock64();
        while (ff_ld_acquire_sys(mine) != seq) {
            if (clock64() - t0 > FF_P2P_SPIN_LIMIT) {
                D->timed_out = 1;
                break;
            }
        }
    }
    __syncwarp();
    const double *src = reinterpret_cast<const double *>(D->peer[rank] + FF_P2P_RDATA) + (size_t)par * FF_P2P_MAXR * 4;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = ff_ld_relaxed_sys(src + k);
        for (int r = 1; r < nranks; ++r) {
            const double v = ff_ld_relaxed_sys(src + (size_t)r * 4 + k);
            s = op_max ? fmax(s, v) : s + v;
        }
        t[k] = s;
    }
}
#endif

void ff_report_error(ffcuda_ctx *ctx, const char *msg);
// every entry point starts with this: selects the device and makes ctx the thread's current context, whose stream
// orders the device allocations of DBuf (stream-ordered allocator; the pool keeps its memory between calls)
void ff_enter(ffcuda_ctx *ctx);
ffcuda_ctx *ff_current_ctx();
void ff_prof_flush(ffcuda_ctx *ctx);
void ff_comm_release(ffcuda_ctx *ctx); // comm.cu

// RAII device buffer
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    ffcuda_ctx *owner = nullptr; // context whose caching allocator owns the block (nullptr: plain cudaMalloc)
    void alloc(size_t count)
    {
        release();
        n = count;
        if (!count) return;
        owner = ff_current_ctx();
        if (owner) p = static_cast<T *>(ff_pool_alloc(owner, count * sizeof(T)));
        else FF_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void release()
    {
        if (p) {
            if (owner) ff_pool_free(owner, p);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// kernel launch wrapper: counts the launch, optionally brackets it with events on the launch stream
template <class F>
inline void ff_launch(ffcuda_ctx *ctx, const char *name, F &&f)
{
    ctx->launches++;
    if (ctx->prof) {
        ProfPending pp;
        pp.name = name;
        FF_CUDA(cudaEventCreate(&pp.e0));
        FF_CUDA(cudaEventCreate(&pp.e1));
        FF_CUDA(cudaEventRecord(pp.e0, ctx->stream));
        f();
        FF_CUDA(cudaEventRecord(pp.e1, ctx->stream));
        ctx->prof_pending.push_back(pp);
        if (ctx->prof_pending.size() > 4096) ff_prof_flush(ctx);
    } else {
        f();
    }
    FF_CUDA(cudaGetLastError());
}

static inline int ff_blocks(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

// ---- handle structs --------------------------------------------------------------------------------
struct ffcuda_mesh {
    CtxRef ref; // keep first
    ffcuda_ctx *ctx = nullptr;
    int dim = 0, nv = 0, nt = 0, nbe = 0;
    int vstride = 0;          // doubles per vertex on the device: 2 (2-D) or 4 (3-D, padded: one 32 B sector)
    DBuf<double> xyz;         // nv*vstride
    DBuf<int32_t> conn;       // nt*(dim+1)
    DBuf<int32_t> elab;       // nt
    DBuf<int32_t> bconn, blab, belem, bface;
    DBuf<int32_t> adj;        // element adjacency (ffcuda_mesh_adjacency), built on first use
    // distributed: local vertices [0,nv_owned) are owned, the rest are ghosts, grouped by owner rank.  Per neighbour x:
    // the ghosts owned by nbr[x] are the contiguous range [recv_off, recv_off + recv_cnt); what nbr[x] needs from me is
    // either the contiguous range [send_off, send_off + send_cnt) of my owned vertices (slab partition of a cube) or the
    // gather list send_idx[send_off .. send_off + send_cnt) (any partition: ffcuda_mesh_upload_distributed), in the
    // order of ITS ghost range
    int nv_owned = 0;
    DBuf<int64_t> gid;        // global vertex id of each local vertex
    static constexpr int MAXNBR = 16;
    int nnbr = 0;
    int nbr[MAXNBR];          // neighbour ranks (-1: no neighbour in this slot)
    int send_off[MAXNBR], send_cnt[MAXNBR], recv_off[MAXNBR], recv_cnt[MAXNBR];
    DBuf<int32_t> send_idx;   // gather lists (null: contiguous ranges)
    ffcuda_mesh()
    {
        for (int x = 0; x < MAXNBR; ++x) {
            nbr[x] = -1;
            send_off[x] = send_cnt[x] = recv_off[x] = recv_cnt[x] = 0;
        }
    }
    bool distributed = false;
};

// node -> (element, local node) incidence lists = the transpose of the element -> node table.  A property of the FE
// space: built on first use and kept with it (re-assemblies on the same fespace reuse it).
//   order 1 (thread-per-row kernels): ELL-32 layout - rows are taken in blocks of 32 (one warp), record e of lane l of
//     block b sits at blkoff[b] + e*32 + l, so a warp walking its 32 rows in lockstep reads 128 contiguous bytes;
//     blocks are padded with FF_NOREC to the longest list of the block.
//   order 2 (lane-group-per-row kernels): plain CSR lists, record e of row r at incptr[r] + e.
static constexpr uint32_t FF_NOREC = 0xffffffffu;
struct Incidence {
    bool built = false;
    int ell = 0;
    int nrows = 0, maxinc = 0;
    int64_t nrec = 0;         // records allocated (padding included)
    DBuf<int32_t> cnt;        // nrows+1 list lengths (cnt[nrows] = 0)
    DBuf<int32_t> incptr;     // nrows+1 (CSR layout only)
    DBuf<uint32_t> blkoff;    // nblk+1  (ELL layout only)
    DBuf<uint32_t> inc;       // nrec records: (element << 4) | local node, each list sorted by element
    // ELL layout only - vertex staging tables for the thread-per-row kernels: the distinct vertices touched by the 32
    // rows of a block IN ASCENDING ORDER (at most FF_STAGE_MAX, else the block is marked unstaged and read through
    // global memory), and for every record the 4 block-local slots (= ranks in that list) of its element's vertices
    DBuf<uint32_t> loc;       // nrec words, byte i = slot of the record's i-th vertex in OWNER-FIRST order (see blk_load)
    DBuf<int32_t> blkvert;    // nblk * FF_STAGE_MAX global vertex ids
    DBuf<int32_t> blkvcnt;    // nblk: number of distinct vertices, -1 = unstaged
    int maxstage = 0;         // largest blkvcnt
    int nunstaged = 0;        // number of blocks that could not be staged
    int nempty = 0;           // rows that no element touches (ELL layout only)
};
static constexpr int FF_STAGE_MAX = 256;
struct IncView {
    const int32_t *cnt, *incptr;
    const uint32_t *blkoff, *inc;
    int ell;
    __device__ __forceinline__ size_t idx(int row, int e) const
    {
        return ell ? (size_t)blkoff[row >> 5] + (size_t)e * 32 + (row & 31) : (size_t)incptr[row] + e;
    }
};
static inline IncView ff_view(const Incidence &I) { return IncView{I.cnt.p, I.incptr.p, I.blkoff.p, I.inc.p, I.ell}; }

// Row tiles of a scalar P1 space (tiles.cu): compact clusters of at most TR rows (consecutive in the Morton order of the
// vertex coordinates) with everything the tile kernel needs packed in one blob per tile: the distinct vertices the
// tile touches, every element touching one of its rows ONCE (4 block-local vertex slots), and for every matrix entry
// of its rows the list of (element, local vertex pair) contributions.  A property of the fespace, built once.
struct TileSet {
    int state = 0;            // 0 not built, 1 ready, -1 not applicable (some tile exceeds the kernel's capacities)
    int tr = 0, ntiles = 0;
    int nes = 0;              // stride of the numeric kernel's value table (max_nelem | 1), baked into the codes
    int max_rows = 0, max_nvt = 0, max_nelem = 0, max_nq = 0, max_ncodes = 0, max_words = 0;
    int max_pre = 0, max_rwords = 0; // longest descriptor head (up to the entry words) / record-list blob, in words
    int64_t sum_nelem = 0;    // element evaluations per assembly (diagnostics: redundancy = sum_nelem / nt)
    int64_t nnz_node = 0;     // of the pattern whose row pointers are baked into the blobs
    DBuf<uint32_t> blob;      // tile descriptors, 16-byte aligned each
    DBuf<uint32_t> toff;      // ntiles+1 offsets into blob in 32-bit words
    DBuf<uint32_t> tpre;      // words of every descriptor's head (row ids, coordinates, element words): all a rhs needs
    DBuf<uint32_t> rblob;     // record lists of the rows, one blob per tile: (element << 2 | local vertex) of every star
    DBuf<uint32_t> roff;      // ntiles+1 offsets into rblob
    // fan set (tiles.cu, 3-D): the same tiles with their elements grouped in fans around a common edge
    bool has_blob = true;     // the element-by-element descriptors (blob / rblob) exist (not built on 3-D spaces that run on the fans)
    int fan_state = 0;        // 1 ready, -1 not applicable
    int fan_max_head = 0, fan_max_b = 0, fan_max_nvals = 0, fan_max_nq = 0; // largest part A / part B (words), value table, entries
    int64_t fan_sum_fans = 0;
    int fan_max_c = 0, fan_max_rvals = 0; // part C (right-hand sides): largest descriptor (words) and value table
    DBuf<uint32_t> fcblob, fcoff;         // part C descriptors and their offsets
    DBuf<uint32_t> fblob, foff, fhead; // descriptors, ntiles+1 offsets (words), words of every descriptor's head (the part copied to shared memory)
};

struct ffcuda_space {
    CtxRef ref; // keep first
    ffcuda_mesh *mesh = nullptr;
    ffcuda_ctx *ctx = nullptr;
    int order = 1, ncomp = 1, nloc = 0, nnodes = 0, nnodes_owned = 0;
    DBuf<int32_t> e2n_own;    // nt*nloc when order 2
    const int32_t *e2n = nullptr;   // = conn for P1
    Incidence incidence;
    TileSet tiles;
    int lean_assemblies = 0;  // scalar P1 assemblies seen on this space (the tile set is built from the second one on)
    int64_t sym_nnz_node = 0; // pattern size found by the first symbolic phase on this space (scalar P1, staged)
    int sym_maxrow = 0;
    // node -> boundary-element incidence (assemble.cu, boundary integrals of linear forms), built on first use
    DBuf<int32_t> bnd_ptr;
    DBuf<uint32_t> bnd_items;
    DBuf<int32_t> bnd2_ptr;   // the same with ALL nodes of the element adjacent to a boundary element (terms with derivatives)
    DBuf<uint32_t> bnd2_items;
    // P2: node rows sorted by decreasing length (assemble.cu launch_p2), rows [0, p2_nlong) are the long ones
    DBuf<int32_t> p2_rowperm;
    int p2_nlong = 0, p2_short_maxrow = 0;
    // a space on a distributed mesh whose nodes are not the vertices (P2: ffcuda_space_create_distributed) carries its own
    // node-level halo lists, same meaning as the vertex-level ones of the mesh; own_halo false: the mesh's lists apply
    bool own_halo = false;
    int nnbr = 0;
    int nbr[ffcuda_mesh::MAXNBR], send_off[ffcuda_mesh::MAXNBR], send_cnt[ffcuda_mesh::MAXNBR], recv_off[ffcuda_mesh::MAXNBR],
        recv_cnt[ffcuda_mesh::MAXNBR];
    DBuf<int32_t> send_idx;
};
// tiles.cu: numeric assembly of c grad u.grad v + m u v on a scalar P1 space by row tiles; returns false when the tile
// path does not apply (the caller then runs the thread-per-row kernel)
bool ff_asm_p1_tiles(ffcuda_ctx *ctx, ffcuda_matrix *A, ffcuda_space *s, double cw, double cmd, double cmo, int accumulate);
bool ff_rhs_p1_tiles(ffcuda_ctx *ctx, ffcuda_vec *b, ffcuda_space *s, const double *cval, const double *cgrad, int hasgrad, int accumulate);
void ff_build_incidence(ffcuda_space *s); // symbolic.cu; no-op when already built

struct ffcuda_pattern {
    CtxRef ref; // keep first
    ffcuda_space *space = nullptr;
    ffcuda_ctx *ctx = nullptr;
    int nrows_node = 0;       // owned nodes = block rows
    int ncols_node = 0;       // local nodes (owned + ghost)
    int ncomp = 1;
    int n = 0;                // dof rows
    int64_t nnz = 0;          // dof-level
    int64_t nnz_node = 0;
    int maxrow_node = 0;      // longest node row
    DBuf<int32_t> nrowptr, ncol;          // node-level CSR
    DBuf<int32_t> rowptr_own, colind_own; // dof-level CSR (only when ncomp > 1)
    const int32_t *rowptr = nullptr, *colind = nullptr;
    // per incidence record (same indexing as space->incidence.inc), nlocp entries: position of each node of that
    // element inside the node row.  P1: in OWNER-FIRST order (entry 0 = the diagonal), see blk_load in symbolic.cu
    DBuf<uint8_t> pos8;
    DBuf<uint16_t> pos16;     // used instead when maxrow_node > 255
    int nlocp = 0;            // padded nloc in the pos table (4 for P1, nloc for P2)
    DBuf<int32_t> diagpos;    // n: index into vals of A(i,i)
    DBuf<int32_t> lower_rowptr; // row pointers of the lower triangle (sym=1 hand-off), built on first use
    int64_t lower_nnz = 0;
    bool copy_pending = false; // an asynchronous download of rowptr / colind may still be in flight on ctx->copy_stream
    bool dist = false;         // created on a distributed mesh (then, and only then, the space and its mesh must outlive the pattern:
                               // solves read their halo lists; otherwise a matrix and its pattern may be used after the space is gone)
};

struct ffcuda_matrix {
    CtxRef ref; // keep first
    ffcuda_ctx *ctx = nullptr;
    ffcuda_pattern *pattern = nullptr;    // null for from_csr matrices
    int n = 0, ncols = 0;
    int64_t nnz = 0;
    DBuf<int32_t> rowptr_own, colind_own, diagpos_own;
    const int32_t *rowptr = nullptr, *colind = nullptr, *diagpos = nullptr;
    DBuf<double> vals;
    bool vals_stale = false;  // allocated but not yet zeroed/written (ffcuda_matrix_create defers the memset)
    bool rect = false;        // rectangular (ffcuda_assemble_bilinear_rect): n rows, ncols columns, no diagonal; products and hand-off only
    int maxrow = 0;           // longest dof row
    // CSR-stream SpMV set-up (lazily built, once per matrix): row-block table
    int stream_state = 0;     // 0 not prepared, 1 ready, -1 not applicable
    int stream_nblk = 0, stream_T = 1, stream_grid = 1;
    size_t stream_shmem = 0;
    DBuf<int32_t> stream_rb;
    // SELL-32 copy for the SpMV (sliced ELLPACK, slices of 32 consecutive rows, entry k of lane l at off + k*32 + l):
    // structure built once per matrix, values re-packed when vals_epoch moved on
    int sell_state = 0;       // 0 not prepared, 1 ready, -1 not applicable (too much padding)
    int sell_nslices = 0;
    int64_t sell_entries = 0; // padded entries
    DBuf<int32_t> sell_off, sell_col;
    DBuf<double> sell_val;
    uint64_t vals_epoch = 1, sell_epoch = 0; // vals_epoch: bumped by everything that writes vals
    // CG workspace (lazily allocated)
    DBuf<double> wG, wH, wAH, wD1, wX;
    DBuf<int32_t> wcl;
    // a matrix handed over as the rows of one rank (ffcuda_matrix_from_csr_distributed) carries its own halo lists: rows =
    // owned dofs, columns = owned then ghost dofs; same meaning as the lists of a distributed mesh / space
    bool own_halo = false;
    int nnbr = 0;
    int nbr[ffcuda_mesh::MAXNBR], send_off[ffcuda_mesh::MAXNBR], send_cnt[ffcuda_mesh::MAXNBR], recv_off[ffcuda_mesh::MAXNBR],
        recv_cnt[ffcuda_mesh::MAXNBR];
    DBuf<int32_t> send_idx;
};

struct ffcuda_vec {
    CtxRef ref; // keep first
    ffcuda_ctx *ctx = nullptr;
    int n = 0;
    DBuf<double> d;
};

struct ffcuda_bc {
    CtxRef ref; // keep first
    ffcuda_ctx *ctx = nullptr;
    int ndofs = 0;
    DBuf<int32_t> dofs;
    DBuf<double> vals;
};

// symbolic.cu: node-level pattern (rows = nodes of sv, columns = nodes of su) of a rectangular matrix
void ff_rect_node_pattern(ffcuda_space *sv, ffcuda_space *su, DBuf<int32_t> &nrowptr, DBuf<int32_t> &ncol, int64_t *nnz_node,
                          int *maxrow_node);
void ff_pattern_ensure_colind(ffcuda_pattern *P); // symbolic.cu: dof-level column indices of a vector-space pattern, on demand
void ff_pattern_ensure_pos(ffcuda_pattern *P); // symbolic.cu: per-record positions of a P1 pattern, on demand
void ff_matrix_touch(ffcuda_matrix *A); // matrix.cu: zero the values if nothing has written them yet
// column indices of a matrix (those of its pattern, materialised on first use for vector spaces)
static inline const int32_t *ff_matrix_colind(ffcuda_matrix *A)
{
    if (!A->colind && A->pattern) {
        ff_pattern_ensure_colind(A->pattern);
        A->colind = A->pattern->colind;
    }
    return A->colind;
}

// ---- shared device helpers ----------------------------------------------------------------------
void ff_exclusive_scan_i32(ffcuda_ctx *ctx, const int32_t *in, int32_t *out, size_t n, int64_t *total);
// out[n] receives the total as well when with_total
int ff_nloc(int dim, int order);
