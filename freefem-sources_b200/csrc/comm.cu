// comm.cu — multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Counterpart of the reference's MPI use on this path (one MPI_Allreduce per dot product in plugin/mpi/MPICG.cpp:93-101,
// element-range split of the assembly loop in fflib/problem.cpp:1133-1138): here assembly needs NO communication
// (every rank owns whole rows and holds a one-element-deep halo of elements), the SpMV inside CG exchanges one layer
// of ghost values with at most two neighbours (ncclSend/ncclRecv in one group, contiguous ranges, no packing), and the
// dot products are all-reduced as device scalars, stream-ordered, without host synchronisation.
//
// NCCL is bound at run time with dlopen (libnccl.so.2): the single-GPU product and the FreeFEM plugin do not need it.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

template <class F>
void bind(F &fn, const char *name)
{
    fn = reinterpret_cast<F>(dlsym(g_nccl.handle, name));
    if (!fn) throw FFError(std::string("NCCL symbol missing: ") + name);
}

NcclApi &nccl()
{
    if (g_nccl.handle) return g_nccl;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) throw FFError(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
    bind(g_nccl.GetUniqueId, "ncclGetUniqueId");
    bind(g_nccl.CommInitRank, "ncclCommInitRank");
    bind(g_nccl.CommDestroy, "ncclCommDestroy");
    bind(g_nccl.AllReduce, "ncclAllReduce");
    bind(g_nccl.AllGather, "ncclAllGather");
    bind(g_nccl.Send, "ncclSend");
    bind(g_nccl.Recv, "ncclRecv");
    bind(g_nccl.GroupStart, "ncclGroupStart");
    bind(g_nccl.GroupEnd, "ncclGroupEnd");
    bind(g_nccl.GetErrorString, "ncclGetErrorString");
    return g_nccl;
}
} // namespace

#define FF_NCCL(call)                                                                                        \
    do {                                                                                                     \
        ncclResult_t r__ = (call);                                                                           \
        if (r__ != ncclSuccess)                                                                              \
            throw FFError(std::string("NCCL error: ") + nccl().GetErrorString(r__) + " at " + __FILE__ + ":" + \
                          std::to_string(__LINE__));                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Peer mailboxes.  Every rank owns one device buffer ("mailbox"); all of them are mapped into every process with CUDA
// IPC at ffcuda_comm_init.  A collective is then a kernel that STORES into the peers' mailboxes over NVLink/NVSwitch,
// publishes a sequence number behind a system-scope fence and spins on its own mailbox until every peer's number has
// arrived — no NCCL call, no proxy thread, ~5 us instead of ~25 us for the 8- and 16-byte all-reduces of a CG
// iteration.  Two parities of every region: a rank can be at most one collective ahead of its slowest peer.
//   mailbox layout: [RFLAG] u64[2][16]  sequence numbers of the all-reduce contributions
//                   [RDATA] f64[2][16][4] contributions
//                   [HFLAG] u64[2][2]   sequence numbers of the halo layers (direction, parity); [HCNT] block counter
//                   [HDATA] f64[2][2][cap] halo layers
// Sums are formed in rank order on every rank: bit-identical everywhere, reproducible.
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t P2P_HFLAG = FF_P2P_HFLAG, P2P_HCNT = FF_P2P_HCNT, P2P_HDATA = FF_P2P_HDATA;
constexpr int P2P_MAXR = FF_P2P_MAXR;
constexpr size_t P2P_HALO_CAP = (size_t)1 << 20; // doubles per (source rank, parity) region (8 MB): interfaces up to 1 M dofs per pair of ranks

struct PeerPtrs {
    unsigned char *p[P2P_MAXR];
};
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) { ff_st_release_sys(p, v); }
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) { return ff_ld_acquire_sys(p); }
__device__ __forceinline__ double ld_sys_f64(const double *p) { return ff_ld_relaxed_sys(p); }

// stand-alone all-reduce of d[0..count) (count <= 4): one warp
__global__ void k_p2p_allreduce(P2PDesc *D, double *__restrict__ d, int count, int op_max)
{
    double t[4];
    const int lane = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = (lane == 0 && k < count) ? d[k] : 0.0;
    ff_p2p_allreduce_warp<4>(D, t, op_max != 0);
    if (lane == 0)
        for (int k = 0; k < count; ++k) d[k] = t[k];
}

struct HaloArgs {
    int n, nc;                     // neighbours, components per node
    int nbr[P2P_MAXR];
    int send_off[P2P_MAXR], send_cnt[P2P_MAXR], recv_off[P2P_MAXR], recv_cnt[P2P_MAXR]; // in nodes
    const int32_t *send_idx;       // gather lists (null: send_off is a contiguous range of v)
};

// value i of the layer for neighbour x: node-major, nc components per node
__device__ __forceinline__ double halo_pick(const double *__restrict__ v, const HaloArgs &H, int x, size_t i)
{
    const size_t node = i / (size_t)H.nc, c = i - node * (size_t)H.nc;
    const size_t src = H.send_idx ? (size_t)H.send_idx[H.send_off[x] + node] : (size_t)H.send_off[x] + node;
    return v[src * H.nc + c];
}

// all blocks co-resident (grid <= number of SMs): pack + send phase (my values go straight into the neighbours' mailboxes,
// region [my rank][parity]: the gather IS the peer store), grid-wide "last block publishes", receive phase
__global__ void __launch_bounds__(256) k_p2p_halo(const PeerPtrs P, int rank, double *__restrict__ v, const HaloArgs H, size_t cap,
                                                  unsigned long long seq, P2PDesc *D)
{
    const int par = (int)(seq & 1ull);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (int x = 0; x < H.n; ++x) {
        if (H.nbr[x] < 0) continue;
        double *dst = reinterpret_cast<double *>(P.p[H.nbr[x]] + P2P_HDATA) + (size_t)(rank * 2 + par) * cap;
        const size_t cnt = (size_t)H.send_cnt[x] * H.nc;
        for (size_t i = tid; i < cnt; i += nthr) dst[i] = halo_pick(v, H, x, i);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    unsigned int *cnt = reinterpret_cast<unsigned int *>(P.p[rank] + P2P_HCNT);
    if (threadIdx.x == 0) last = atomicAdd(cnt, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < H.n && H.nbr[threadIdx.x] >= 0) {
        __threadfence_system();
        st_sys_u64(reinterpret_cast<unsigned long long *>(P.p[H.nbr[threadIdx.x]] + P2P_HFLAG) + rank * 2 + par, seq);
    }
    if (last) {
        __syncthreads();
        if (threadIdx.x == 0) *cnt = 0;
    }
    for (int x = 0; x < H.n; ++x) {
        if (H.nbr[x] < 0) continue;
        if (threadIdx.x == 0) {
            const unsigned long long *f = reinterpret_cast<const unsigned long long *>(P.p[rank] + P2P_HFLAG) + H.nbr[x] * 2 + par;
            const long long t0 = clock64();
            while (ld_sys_u64(f) != seq) {
                if (clock64() - t0 > FF_P2P_SPIN_LIMIT) { // crashed neighbour: do not hang the device
                    D->timed_out = 1;
                    break;
                }
            }
        }
        __syncthreads();
        const double *src = reinterpret_cast<const double *>(P.p[rank] + P2P_HDATA) + (size_t)(H.nbr[x] * 2 + par) * cap;
        double *dst = v + (size_t)H.recv_off[x] * H.nc;
        const size_t cnt2 = (size_t)H.recv_cnt[x] * H.nc;
        for (size_t i = tid; i < cnt2; i += nthr) dst[i] = ld_sys_f64(src + i);
    }
}

// NCCL route with gather lists: the layers are packed into a staging buffer first
__global__ void k_halo_pack(const double *__restrict__ v, const HaloArgs H, int x, double *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)H.send_cnt[x] * H.nc) out[i] = halo_pick(v, H, x, i);
}

PeerPtrs peer_ptrs(const ffcuda_ctx *ctx)
{
    PeerPtrs P;
    for (int r = 0; r < P2P_MAXR; ++r) P.p[r] = static_cast<unsigned char *>(ctx->p2p_peer[r]);
    return P;
}

// mailboxes of all ranks: allocate mine, all-gather the IPC handles through NCCL, map the peers'
void p2p_setup(ffcuda_ctx *ctx)
{
    if (ctx->nranks < 2 || ctx->nranks > P2P_MAXR) return;
    if (const char *e = getenv("FFCUDA_P2P"))
        if (atoi(e) == 0) return;
    cudaStream_t st = ctx->stream;
    const int rank = ctx->rank, n = ctx->nranks;
    const size_t bytes = P2P_HDATA + (size_t)2 * P2P_MAXR * P2P_HALO_CAP * sizeof(double);
    void *mine = nullptr;
    // every rank must take the same decision: the outcome of each step is all-reduced (min) before going on
    int ok = cudaMalloc(&mine, bytes) == cudaSuccess ? 1 : 0;
    if (ok) ok = cudaMemsetAsync(mine, 0, bytes, st) == cudaSuccess;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (ok) ok = cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    DBuf<unsigned char> d_all;
    d_all.alloc((size_t)(n + 1) * 80);
    std::vector<unsigned char> rec(80, 0), all((size_t)n * 80, 0);
    memcpy(rec.data(), &h, 64);
    rec[64] = (unsigned char)ok;
    // ranks that are threads of ONE process (the FreeFEM plugin driving several GPUs) cannot open each other's IPC handles:
    // they exchange the pointer itself and enable peer access between the two devices
    const int32_t mypid = (int32_t)getpid();
    const uint64_t myptr = (uint64_t)(uintptr_t)mine;
    memcpy(rec.data() + 65, &mypid, 4);
    memcpy(rec.data() + 69, &myptr, 8);
    rec[77] = (unsigned char)ctx->device;
    FF_CUDA(cudaMemcpyAsync(d_all.p + (size_t)n * 80, rec.data(), 80, cudaMemcpyHostToDevice, st));
    FF_NCCL(nccl().AllGather(d_all.p + (size_t)n * 80, d_all.p, 80, ncclChar, (ncclComm_t)ctx->nccl_comm, st));
    FF_CUDA(ff_memcpy_sync(ctx, all.data(), d_all.p, (size_t)n * 80, cudaMemcpyDeviceToHost)); // also: every memset is done
    for (int r = 0; r < n; ++r) ok = ok && all[(size_t)r * 80 + 64];
    int opened = 0;
    if (ok) {
        for (int r = 0; r < n && ok; ++r) {
            if (r == rank) {
                ctx->p2p_peer[r] = mine;
                continue;
            }
            int32_t rpid = 0;
            uint64_t rptr = 0;
            memcpy(&rpid, &all[(size_t)r * 80 + 65], 4);
            memcpy(&rptr, &all[(size_t)r * 80 + 69], 8);
            if (rpid == mypid) {
                const int rdev = all[(size_t)r * 80 + 77];
                int can = 0;
                cudaError_t e = cudaDeviceCanAccessPeer(&can, ctx->device, rdev);
                if (e == cudaSuccess && can) {
                    e = cudaDeviceEnablePeerAccess(rdev, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) e = cudaSuccess;
                }
                cudaGetLastError();
                if (e != cudaSuccess || !can) {
                    ok = 0;
                    break;
                }
                ctx->p2p_peer[r] = (void *)(uintptr_t)rptr;
                ctx->p2p_inproc[r] = true;
                ++opened;
                continue;
            }
            cudaIpcMemHandle_t hr;
            memcpy(&hr, &all[(size_t)r * 80], 64);
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            ctx->p2p_peer[r] = p;
            ++opened;
        }
    }
    // second agreement round: did every rank map every peer?
    DBuf<double> flag;
    flag.alloc(1);
    double hv = ok ? 1.0 : 0.0;
    FF_CUDA(cudaMemcpyAsync(flag.p, &hv, sizeof(double), cudaMemcpyHostToDevice, st));
    FF_NCCL(nccl().AllReduce(flag.p, flag.p, 1, ncclDouble, ncclMin, (ncclComm_t)ctx->nccl_comm, st));
    FF_CUDA(ff_memcpy_sync(ctx, &hv, flag.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (hv < 0.5) {
        for (int r = 0; r < n; ++r) {
            if (r != rank && ctx->p2p_peer[r] && !ctx->p2p_inproc[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
            ctx->p2p_peer[r] = nullptr;
            ctx->p2p_inproc[r] = false;
        }
        if (mine) cudaFree(mine);
        cudaGetLastError();
        return;
    }
    ctx->p2p_halo_cap = P2P_HALO_CAP;
    ctx->p2p_seq_halo = 0;
    P2PDesc D;
    memset(&D, 0, sizeof(D));
    for (int r = 0; r < n; ++r) D.peer[r] = static_cast<unsigned char *>(ctx->p2p_peer[r]);
    D.rank = rank;
    D.nranks = n;
    FF_CUDA(ff_memcpy_sync(ctx, ctx->d_scal + FF_P2P_DESC_OFF, &D, sizeof(D), cudaMemcpyHostToDevice));
    ctx->p2p = true;
    if (getenv("FFCUDA_VERBOSE")) fprintf(stderr, "ffcuda rank %d: peer mailboxes of %d ranks mapped (%d opened)\n", rank, n, opened);
}

void p2p_teardown(ffcuda_ctx *ctx)
{
    if (!ctx->p2p) return;
    for (int r = 0; r < ctx->nranks; ++r) {
        if (!ctx->p2p_peer[r]) continue;
        if (r == ctx->rank) cudaFree(ctx->p2p_peer[r]);
        else if (!ctx->p2p_inproc[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
        ctx->p2p_peer[r] = nullptr;
        ctx->p2p_inproc[r] = false;
    }
    ctx->p2p = false;
}
} // namespace

extern "C" int ffcuda_comm_unique_id(void *id128)
{
    FF_API_BEGIN
    FF_REQUIRE(id128, "null id buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    FF_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    FF_API_END(nullptr)
}

extern "C" int ffcuda_comm_init(ffcuda_ctx *ctx, int rank, int nranks, const void *id128)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && nranks >= 1 && rank >= 0 && rank < nranks, "ffcuda_comm_init: bad arguments");
    FF_REQUIRE(!ctx->nccl_comm, "communicator already initialised");
    ff_enter(ctx);
    if (nranks > 1) {
        FF_REQUIRE(id128, "ffcuda_comm_init: null NCCL id");
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclComm_t comm;
        FF_NCCL(nccl().CommInitRank(&comm, nranks, id, rank));
        ctx->nccl_comm = comm;
    }
    ctx->rank = rank;
    ctx->nranks = nranks;
    p2p_setup(ctx); // collective: every rank maps every rank's mailbox, or none does (then NCCL carries the CG traffic)
    FF_API_END(ctx)
}

void ff_comm_release(ffcuda_ctx *ctx)
{
    p2p_teardown(ctx);
    if (ctx->nccl_comm) {
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
}

extern "C" int ffcuda_comm_finalize(ffcuda_ctx *ctx)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx, "null context");
    ff_enter(ctx);
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    ff_comm_release(ctx);
    FF_API_END(ctx)
}

static ffcuda_mesh *dist_mesh(ffcuda_matrix *A)
{
    // (the space of a pattern on an ordinary mesh may be gone by now - the FreeFEM plugin keeps a matrix and its pattern with
    //  the solver while its cache of fespaces moves on - so it is only looked at when the pattern says the mesh is distributed)
    if (!A->pattern || !A->pattern->dist || !A->pattern->space) return nullptr;
    ffcuda_mesh *m = A->pattern->space->mesh;
    return (m && m->distributed) ? m : nullptr;
}

// rows of one rank: a matrix on a distributed mesh, or one handed over with its own halo lists
bool ff_is_distributed(ffcuda_matrix *A) { return (A->own_halo && A->ctx->nranks > 1) || dist_mesh(A) != nullptr; }

// the halo lists that apply to a matrix: those of its space when it has its own (P2 on a distributed mesh), else the mesh's
struct HaloView {
    int nnbr = 0;
    const int *nbr = nullptr, *send_off = nullptr, *send_cnt = nullptr, *recv_off = nullptr, *recv_cnt = nullptr;
    const int32_t *send_idx = nullptr;
};
static HaloView halo_view(ffcuda_matrix *A, ffcuda_mesh *m)
{
    HaloView H;
    if (A->own_halo) {
        H.nnbr = A->nnbr; H.nbr = A->nbr; H.send_off = A->send_off; H.send_cnt = A->send_cnt; H.recv_off = A->recv_off; H.recv_cnt = A->recv_cnt;
        H.send_idx = A->send_idx.p;
        return H;
    }
    ffcuda_space *s = A->pattern->space;
    if (s->own_halo) {
        H.nnbr = s->nnbr; H.nbr = s->nbr; H.send_off = s->send_off; H.send_cnt = s->send_cnt; H.recv_off = s->recv_off; H.recv_cnt = s->recv_cnt;
        H.send_idx = s->send_idx.p;
    } else {
        H.nnbr = m->nnbr; H.nbr = m->nbr; H.send_off = m->send_off; H.send_cnt = m->send_cnt; H.recv_off = m->recv_off; H.recv_cnt = m->recv_cnt;
        H.send_idx = m->send_idx.p;
    }
    return H;
}

// v holds owned values in [0, n) ; fills the ghost ranges [n, ncols) from the neighbours' owned boundary layers
void ff_halo_exchange(ffcuda_matrix *A, double *v)
{
    if (!ff_is_distributed(A)) return;
    ffcuda_mesh *m = dist_mesh(A);
    ffcuda_ctx *ctx = A->ctx;
    FF_REQUIRE(ctx->nccl_comm, "distributed matrix without communicator");
    const HaloView V = halo_view(A, m);
    const int nc = A->pattern ? A->pattern->ncomp : 1;
    HaloArgs H;
    memset(&H, 0, sizeof(H));
    H.n = V.nnbr;
    H.nc = nc;
    H.send_idx = V.send_idx;
    // layers that fit the mailbox regions go by peer stores (one kernel for all neighbours), the others by NCCL; both
    // ends of a pair see the same counts, so they take the same route
    bool via_p2p[P2P_MAXR], any_p2p = false, any_nccl = false;
    size_t most = 0;
    for (int x = 0; x < V.nnbr; ++x) {
        via_p2p[x] = false;
        H.nbr[x] = -1;
        H.send_off[x] = V.send_off[x]; H.send_cnt[x] = V.send_cnt[x];
        H.recv_off[x] = V.recv_off[x]; H.recv_cnt[x] = V.recv_cnt[x];
        if (V.nbr[x] < 0) continue;
        const size_t big = (size_t)std::max(V.send_cnt[x], V.recv_cnt[x]) * nc;
        via_p2p[x] = ctx->p2p && big <= ctx->p2p_halo_cap;
        (via_p2p[x] ? any_p2p : any_nccl) = true;
        if (via_p2p[x]) {
            H.nbr[x] = V.nbr[x];
            most = std::max(most, big);
        }
    }
    if (any_p2p) {
        const unsigned long long seq = ++ctx->p2p_seq_halo;
        const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count / 2, (most + 1023) / 1024));
        ff_launch(ctx, "halo_p2p", [&] {
            k_p2p_halo<<<grid, 256, 0, ctx->stream>>>(peer_ptrs(ctx), ctx->rank, v, H, ctx->p2p_halo_cap, seq,
                                                      reinterpret_cast<P2PDesc *>(ctx->d_scal + FF_P2P_DESC_OFF));
        });
    }
    if (!any_nccl) return;
    NcclApi &N = nccl();
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    DBuf<double> stage;
    std::vector<size_t> soff((size_t)V.nnbr + 1, 0);
    if (V.send_idx) { // gather lists: pack first
        for (int x = 0; x < V.nnbr; ++x) soff[(size_t)x + 1] = soff[x] + ((V.nbr[x] >= 0 && !via_p2p[x]) ? (size_t)V.send_cnt[x] * nc : 0);
        stage.alloc(std::max<size_t>(soff[V.nnbr], 1));
        for (int x = 0; x < V.nnbr; ++x) {
            if (V.nbr[x] < 0 || via_p2p[x] || V.send_cnt[x] == 0) continue;
            H.nbr[x] = V.nbr[x];
            ff_launch(ctx, "halo_pack", [&] {
                k_halo_pack<<<ff_blocks((size_t)V.send_cnt[x] * nc, 256), 256, 0, ctx->stream>>>(v, H, x, stage.p + soff[x]);
            });
        }
    }
    FF_NCCL(N.GroupStart());
    for (int x = 0; x < V.nnbr; ++x) {
        if (V.nbr[x] < 0 || via_p2p[x]) continue;
        const double *src = V.send_idx ? stage.p + soff[x] : v + (size_t)V.send_off[x] * nc;
        // (a layer of length zero - a matrix whose structure is not symmetric - is skipped on both ends: counts mirror each other)
        if (V.send_cnt[x] > 0) FF_NCCL(N.Send(src, (size_t)V.send_cnt[x] * nc, ncclDouble, V.nbr[x], comm, ctx->stream));
        if (V.recv_cnt[x] > 0)
            FF_NCCL(N.Recv(v + (size_t)V.recv_off[x] * nc, (size_t)V.recv_cnt[x] * nc, ncclDouble, V.nbr[x], comm, ctx->stream));
    }
    FF_NCCL(N.GroupEnd());
    ctx->launches++;
}

void ff_allreduce(ffcuda_matrix *A, double *d, int count, int op_max)
{
    if (!ff_is_distributed(A)) return;
    ffcuda_ctx *ctx = A->ctx;
    FF_REQUIRE(ctx->nccl_comm, "distributed matrix without communicator");
    if (ctx->p2p && count <= 4) {
        ff_launch(ctx, "allreduce_p2p", [&] {
            k_p2p_allreduce<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<P2PDesc *>(ctx->d_scal + FF_P2P_DESC_OFF), d, count, op_max);
        });
        return;
    }
    FF_NCCL(nccl().AllReduce(d, d, (size_t)count, ncclDouble, op_max ? ncclMax : ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->launches++;
}
