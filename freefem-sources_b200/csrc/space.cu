// space.cu — finite-element space handle: element -> node table (the dof numbering FreeFEM's FESpace holds).
#include "common.cuh"
#include <unordered_map>

// 3-D P2 nodes numbered like GenericMesh::BuildDFNumbering (femlib/GenericMesh.hpp:1878-1929): walking the
// elements in order, the 4 vertices then the 6 edges {01,02,03,12,13,23}; a node gets the next free number
// the first time its key (vertex id, or sorted vertex pair) is met.  Host-side, like the reference (it is
// done once per fespace); input preparation, not part of the timed hot path.
static int number_p2_nodes_3d(int nt, const std::vector<int32_t> &conn, std::vector<int32_t> &e2n)
{
    static const int edge[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    std::unordered_map<uint64_t, int32_t> seen;
    seen.reserve((size_t)nt * 2 + 1024);
    e2n.resize((size_t)nt * 10);
    int32_t next = 0;
    for (int k = 0; k < nt; ++k) {
        const int32_t *K = &conn[(size_t)k * 4];
        for (int a = 0; a < 10; ++a) {
            uint64_t key;
            if (a < 4)
                key = ((uint64_t)(uint32_t)K[a] << 32) | 0xffffffffull;
            else {
                uint32_t p = (uint32_t)K[edge[a - 4][0]], q = (uint32_t)K[edge[a - 4][1]];
                key = p < q ? ((uint64_t)p << 32) | q : ((uint64_t)q << 32) | p;
            }
            auto ins = seen.emplace(key, next);
            if (ins.second) ++next;
            e2n[(size_t)k * 10 + a] = ins.first->second;
        }
    }
    return next;
}

extern "C" int ffcuda_space_create(ffcuda_mesh *m, int order, int ncomp, const int32_t *elem2node, int nnodes,
                                   ffcuda_space **out)
{
    ffcuda_space *s = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(m && out, "ffcuda_space_create: null mesh/output");
    ffcuda_ctx *ctx = m->ctx;
    FF_REQUIRE(order == 1 || order == 2, "only P1 and P2 Lagrange spaces are supported");
    FF_REQUIRE(ncomp >= 1 && ncomp <= 3, "1 to 3 components supported");
    ff_enter(ctx);
    s = new ffcuda_space();
    s->mesh = m;
    s->ctx = ctx;
    s->ref.set(ctx);
    s->order = order;
    s->ncomp = ncomp;
    s->nloc = ff_nloc(m->dim, order);
    if (order == 1 && !elem2node) {
        s->e2n = m->conn.p;
        s->nnodes = m->nv;
        s->nnodes_owned = m->nv_owned;
    } else {
        FF_REQUIRE(!m->distributed, "a P2 space on a distributed mesh needs its node-level lists: ffcuda_space_create_distributed");
        std::vector<int32_t> tab;
        if (!elem2node) {
            FF_REQUIRE(m->dim == 3,
                       "2-D P2: FreeFEM renumbers the nodes (Gibbs, FESpace.cpp:991); pass the element->node table");
            std::vector<int32_t> conn((size_t)m->nt * 4);
            FF_CUDA(ff_memcpy_sync(ctx, conn.data(), m->conn.p, m->conn.bytes(), cudaMemcpyDeviceToHost));
            nnodes = number_p2_nodes_3d(m->nt, conn, tab);
            elem2node = tab.data();
        }
        FF_REQUIRE(nnodes > 0, "nnodes must be positive when an element->node table is given");
        s->e2n_own.alloc((size_t)m->nt * s->nloc);
        FF_CUDA(ff_memcpy_sync(ctx, s->e2n_own.p, elem2node, s->e2n_own.bytes(), cudaMemcpyHostToDevice));
        s->e2n = s->e2n_own.p;
        s->nnodes = nnodes;
        s->nnodes_owned = nnodes;
    }
    FF_REQUIRE((int64_t)s->nnodes * ncomp < ((int64_t)1 << 31), "too many dofs for int32 indices");
    *out = s;
    s = nullptr;
    FF_API_END((delete s, m ? m->ctx : nullptr))
}

// A space on a distributed mesh whose nodes are NOT the vertices (P2: vertices and edges): the caller hands the local node
// table (owned nodes first, ghosts grouped by owner rank - the arrays ffcuda_partition_local_nodes derives from the global
// element -> node table and a node partition) and the node-level halo lists; rows = owned nodes, columns = local nodes.
// An edge node belongs to the rank that owns one of its end points, so that the elements around it are all local there and
// its row assembles without communication (reference counterpart: element-range split + all-reduce of the whole matrix,
// fflib/problem.cpp:1133-1138).
extern "C" int ffcuda_space_create_distributed(ffcuda_mesh *m, int order, int ncomp, const int32_t *elem2node, int nnodes_owned,
                                               int nnodes_local, int nnbr, const int32_t *nbr,
                                               const int32_t *recv_off, const int32_t *recv_cnt, const int32_t *send_ptr,
                                               const int32_t *send_idx, ffcuda_space **out)
{
    ffcuda_space *s = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(m && out && elem2node, "ffcuda_space_create_distributed: null mesh / table / output");
    ffcuda_ctx *ctx = m->ctx;
    FF_REQUIRE(order == 1 || order == 2, "only P1 and P2 Lagrange spaces are supported");
    FF_REQUIRE(ncomp >= 1 && ncomp <= 3, "1 to 3 components supported");
    FF_REQUIRE(nnodes_owned > 0 && nnodes_owned <= nnodes_local, "every rank must own at least one node");
    FF_REQUIRE(nnbr >= 0 && nnbr <= ffcuda_mesh::MAXNBR, "at most 16 neighbour ranks");
    FF_REQUIRE(nnbr == 0 || (nbr && recv_off && recv_cnt && send_ptr && (send_idx || send_ptr[nnbr] == 0)), "halo arrays missing");
    FF_REQUIRE(ctx->nranks == 1 || ctx->nccl_comm, "ffcuda_space_create_distributed: call ffcuda_comm_init first");
    int covered = nnodes_owned;
    for (int x = 0; x < nnbr; ++x) {
        FF_REQUIRE(nbr[x] >= 0 && nbr[x] < ctx->nranks && nbr[x] != ctx->rank, "bad neighbour rank");
        FF_REQUIRE(recv_off[x] == covered && recv_cnt[x] > 0, "ghost ranges must follow the owned nodes, in neighbour order, without gaps");
        covered += recv_cnt[x];
        FF_REQUIRE(send_ptr[x + 1] >= send_ptr[x], "send_ptr must be non-decreasing");
    }
    FF_REQUIRE(covered == nnodes_local, "ghost ranges do not cover the ghost nodes");
    for (int k = 0; k < (nnbr ? send_ptr[nnbr] : 0); ++k) FF_REQUIRE(send_idx[k] >= 0 && send_idx[k] < nnodes_owned, "send list entry is not an owned node");
    const int nloc = ff_nloc(m->dim, order);
    for (size_t i = 0; i < (size_t)m->nt * nloc; ++i)
        FF_REQUIRE(elem2node[i] >= 0 && elem2node[i] < nnodes_local, "element -> node table entry outside the local nodes");
    ff_enter(ctx);
    s = new ffcuda_space();
    s->mesh = m;
    s->ctx = ctx;
    s->ref.set(ctx);
    s->order = order;
    s->ncomp = ncomp;
    s->nloc = nloc;
    s->e2n_own.alloc((size_t)m->nt * nloc);
    FF_CUDA(ff_memcpy_sync(ctx, s->e2n_own.p, elem2node, s->e2n_own.bytes(), cudaMemcpyHostToDevice));
    s->e2n = s->e2n_own.p;
    s->nnodes = nnodes_local;
    s->nnodes_owned = nnodes_owned;
    s->own_halo = true;
    s->nnbr = nnbr;
    for (int x = 0; x < ffcuda_mesh::MAXNBR; ++x) {
        s->nbr[x] = -1;
        s->send_off[x] = s->send_cnt[x] = s->recv_off[x] = s->recv_cnt[x] = 0;
    }
    for (int x = 0; x < nnbr; ++x) {
        s->nbr[x] = nbr[x];
        s->recv_off[x] = recv_off[x];
        s->recv_cnt[x] = recv_cnt[x];
        s->send_off[x] = send_ptr[x];
        s->send_cnt[x] = send_ptr[x + 1] - send_ptr[x];
    }
    if (nnbr && send_ptr[nnbr] > 0) {
        s->send_idx.alloc((size_t)send_ptr[nnbr]);
        FF_CUDA(ff_memcpy_sync(ctx, s->send_idx.p, send_idx, (size_t)send_ptr[nnbr] * 4, cudaMemcpyHostToDevice));
    }
    FF_REQUIRE((int64_t)s->nnodes * ncomp < ((int64_t)1 << 31), "too many dofs for int32 indices");
    *out = s;
    s = nullptr;
    FF_API_END((delete s, m ? m->ctx : nullptr))
}

extern "C" int ffcuda_space_info(ffcuda_space *s, int *ndof, int *ndofK, int *nnodes)
{
    FF_API_BEGIN
    FF_REQUIRE(s, "null space");
    if (ndof) *ndof = s->nnodes * s->ncomp;
    if (ndofK) *ndofK = s->nloc * s->ncomp;
    if (nnodes) *nnodes = s->nnodes;
    FF_API_END(s ? s->ctx : nullptr)
}

extern "C" int ffcuda_space_download_dofs(ffcuda_space *s, int32_t *dof)
{
    FF_API_BEGIN
    FF_REQUIRE(s && dof, "null argument");
    const int nt = s->mesh->nt, nloc = s->nloc, nc = s->ncomp;
    std::vector<int32_t> e2n((size_t)nt * nloc);
    ff_enter(s->ctx);
    FF_CUDA(ff_memcpy_sync(s->ctx, e2n.data(), s->e2n, e2n.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < nt; ++k)
        for (int c = 0; c < nc; ++c)
            for (int a = 0; a < nloc; ++a)
                dof[(size_t)k * nloc * nc + c * nloc + a] = e2n[(size_t)k * nloc + a] * nc + c;
    FF_API_END(s ? s->ctx : nullptr)
}

extern "C" void ffcuda_space_destroy(ffcuda_space *s)
{
    if (!s) return;
    ff_enter(s->ctx);
    delete s;
}
