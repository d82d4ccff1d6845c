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
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
    FF_API_END(ctx)
}

void ff_comm_release(ffcuda_ctx *ctx)
{
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
    if (!A->pattern || !A->pattern->space) return nullptr;
    ffcuda_mesh *m = A->pattern->space->mesh;
    return (m && m->distributed) ? m : nullptr;
}

bool ff_is_distributed(ffcuda_matrix *A) { return dist_mesh(A) != nullptr; }

// v holds owned values in [0, n) ; fills the ghost ranges [n, ncols) from the neighbours' owned boundary layers
void ff_halo_exchange(ffcuda_matrix *A, double *v)
{
    ffcuda_mesh *m = dist_mesh(A);
    if (!m) return;
    ffcuda_ctx *ctx = A->ctx;
    FF_REQUIRE(ctx->nccl_comm, "distributed matrix without communicator");
    const int nc = A->pattern->ncomp;
    NcclApi &N = nccl();
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    FF_NCCL(N.GroupStart());
    for (int s = 0; s < 2; ++s) {
        if (m->nbr[s] < 0) continue;
        FF_NCCL(N.Send(v + (size_t)m->send_off[s] * nc, (size_t)m->send_cnt[s] * nc, ncclDouble, m->nbr[s], comm, ctx->stream));
        FF_NCCL(N.Recv(v + (size_t)m->recv_off[s] * nc, (size_t)m->recv_cnt[s] * nc, ncclDouble, m->nbr[s], comm, ctx->stream));
    }
    FF_NCCL(N.GroupEnd());
    ctx->launches++;
}

void ff_allreduce(ffcuda_matrix *A, double *d, int count, int op_max)
{
    if (!dist_mesh(A)) return;
    ffcuda_ctx *ctx = A->ctx;
    FF_REQUIRE(ctx->nccl_comm, "distributed matrix without communicator");
    FF_NCCL(nccl().AllReduce(d, d, (size_t)count, ncclDouble, op_max ? ncclMax : ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->launches++;
}
