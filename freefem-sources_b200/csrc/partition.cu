// partition.cu — vertex partition of a general (unstructured) mesh and the bookkeeping of a rank's local problem.
// HOST ONLY (no device needed; CPU-tested under gloo, tests/test_partition_gloo.py): the arithmetic the distributed path
// will be built on for meshes that are not the synthetic cube (SURVEY.md §8 e: FreeFEM itself splits the ELEMENT RANGE,
// fflib/problem.cpp:1133-1138, and all-reduces the whole matrix; METIS is only a download recipe, 3rdparty/getall:118).
//
// Model (the same as the z-slab partition of mesh.cu): a rank OWNS vertices = matrix rows; it holds every element that
// touches an owned vertex, so its rows assemble without communication; the other vertices of those elements are its
// GHOSTS (columns only), refreshed before every SpMV.  Local numbering: owned vertices first (ascending global id),
// then the ghosts grouped by owner rank (ascending), ascending global id inside a group - so what a neighbour sends
// arrives as ONE contiguous range, and only the sender gathers.
#include "common.cuh"
#include <algorithm>
#include <numeric>

namespace {

// recursive coordinate bisection: ids[lo, hi) go to parts [p0, p0 + np); split across the longest extent of the bounding
// box at the rank that gives the lower side floor(np / 2) / np of the vertices; ties broken by vertex id (deterministic)
void rcb(int dim, const double *xyz, std::vector<int32_t> &ids, size_t lo, size_t hi, int p0, int np, int32_t *part)
{
    if (np == 1) {
        for (size_t i = lo; i < hi; ++i) part[ids[i]] = p0;
        return;
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (size_t i = lo; i < hi; ++i)
        for (int d = 0; d < dim; ++d) {
            const double x = xyz[(size_t)ids[i] * dim + d];
            mn[d] = std::min(mn[d], x);
            mx[d] = std::max(mx[d], x);
        }
    int ax = 0;
    for (int d = 1; d < dim; ++d)
        if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
    const int npl = np / 2;
    const size_t mid = lo + (size_t)(((hi - lo) * (uint64_t)npl) / (uint64_t)np);
    auto less = [&](int32_t a, int32_t b) {
        const double xa = xyz[(size_t)a * dim + ax], xb = xyz[(size_t)b * dim + ax];
        return xa < xb || (xa == xb && a < b);
    };
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, less);
    rcb(dim, xyz, ids, lo, mid, p0, npl, part);
    rcb(dim, xyz, ids, mid, hi, p0 + npl, np - npl, part);
}

struct LocalLists {
    std::vector<int32_t> l2g;      // local -> global vertex: owned (ascending), then ghosts by (owner, id)
    std::vector<int32_t> elems;    // global ids of the local elements, ascending
    std::vector<int32_t> nbr;      // neighbour ranks, ascending
    std::vector<int32_t> recv_off; // per neighbour: first local index of the ghosts it owns
    std::vector<int32_t> recv_cnt;
    std::vector<int32_t> send_ptr; // per neighbour: range in send_idx
    std::vector<int32_t> send_idx; // owned LOCAL indices to send, in the receiver's ghost order (ascending global id)
    int nowned = 0;
};

LocalLists local_lists(int nvk, int nv, int nt, const int32_t *conn, const int32_t *part, int rank, int nranks)
{
    LocalLists L;
    std::vector<uint8_t> ghost((size_t)nv, 0);
    for (int k = 0; k < nt; ++k) {
        const int32_t *K = conn + (size_t)nvk * k;
        bool mine = false;
        for (int a = 0; a < nvk; ++a) mine = mine || part[K[a]] == rank;
        if (!mine) continue;
        L.elems.push_back(k);
        for (int a = 0; a < nvk; ++a)
            if (part[K[a]] != rank) ghost[K[a]] = 1;
    }
    for (int v = 0; v < nv; ++v)
        if (part[v] == rank) L.l2g.push_back(v);
    L.nowned = (int)L.l2g.size();
    std::vector<int32_t> gh;
    for (int v = 0; v < nv; ++v)
        if (ghost[v]) gh.push_back(v);
    std::stable_sort(gh.begin(), gh.end(), [&](int32_t a, int32_t b) { return part[a] < part[b]; }); // (ids stay ascending inside)
    for (size_t i = 0; i < gh.size(); ++i) {
        const int o = part[gh[i]];
        if (L.nbr.empty() || L.nbr.back() != o) {
            L.nbr.push_back(o);
            L.recv_off.push_back(L.nowned + (int)i);
            L.recv_cnt.push_back(0);
        }
        L.recv_cnt.back()++;
        L.l2g.push_back(gh[i]);
    }
    // what the neighbours need from me: my owned vertices that are ghosts of theirs = vertices of mine in an element that
    // touches a vertex of theirs.  The relation is symmetric (an element with vertices of both ranks is local to both), so
    // the neighbour sets coincide; the list for neighbour r, ascending global id, is r's ghost range owned by me.
    std::vector<std::vector<int32_t>> need((size_t)nranks);
    {
        std::vector<int32_t> stamp((size_t)nv, -1); // last neighbour a vertex was listed for (lists are built rank by rank)
        for (size_t x = 0; x < L.nbr.size(); ++x) {
            const int r = L.nbr[x];
            for (size_t e = 0; e < L.elems.size(); ++e) {
                const int32_t *K = conn + (size_t)nvk * L.elems[e];
                bool theirs = false;
                for (int a = 0; a < nvk; ++a) theirs = theirs || part[K[a]] == r;
                if (!theirs) continue;
                for (int a = 0; a < nvk; ++a)
                    if (part[K[a]] == rank && stamp[K[a]] != r) {
                        stamp[K[a]] = r;
                        need[r].push_back(K[a]);
                    }
            }
            std::sort(need[r].begin(), need[r].end());
        }
    }
    // global -> owned local index: owned vertices are ascending, so a binary search does
    L.send_ptr.push_back(0);
    for (size_t x = 0; x < L.nbr.size(); ++x) {
        for (int32_t g : need[L.nbr[x]])
            L.send_idx.push_back((int32_t)(std::lower_bound(L.l2g.begin(), L.l2g.begin() + L.nowned, g) - L.l2g.begin()));
        L.send_ptr.push_back((int32_t)L.send_idx.size());
    }
    return L;
}

} // namespace

extern "C" int ffcuda_partition_rcb(int dim, int nv, const double *xyz, int nparts, int32_t *part)
{
    FF_API_BEGIN
    FF_REQUIRE((dim == 2 || dim == 3) && nv >= 0 && nparts >= 1 && (nv == 0 || (xyz && part)), "ffcuda_partition_rcb: bad arguments");
    // every part must own a vertex (the distributed mesh rejects empty ranks: ffcuda_mesh_upload_distributed)
    FF_REQUIRE(nv == 0 || nparts <= nv, "ffcuda_partition_rcb: more parts than vertices");
    std::vector<int32_t> ids((size_t)nv);
    std::iota(ids.begin(), ids.end(), 0);
    rcb(dim, xyz, ids, 0, (size_t)nv, 0, nparts, part);
    FF_API_END(nullptr)
}

extern "C" int ffcuda_partition_local(int dim, int nv, int nt, const int32_t *conn, const int32_t *part, int rank, int nranks,
                                      int64_t *sizes8, int32_t *l2g, int32_t *elems, int32_t *nbr, int32_t *recv_off, int32_t *recv_cnt,
                                      int32_t *send_ptr, int32_t *send_idx)
{
    FF_API_BEGIN
    FF_REQUIRE((dim == 2 || dim == 3) && nv >= 0 && nt >= 0 && sizes8 && rank >= 0 && rank < nranks && (nt == 0 || conn) && (nv == 0 || part),
               "ffcuda_partition_local: bad arguments");
    // the inputs index host arrays: a partition vector made for another number of parts, or a connectivity of another mesh,
    // must come back as an error through the ABI, not as an out-of-bounds write
    for (int v = 0; v < nv; ++v) FF_REQUIRE(part[v] >= 0 && part[v] < nranks, "ffcuda_partition_local: part[] entry outside [0, nranks)");
    for (size_t i = 0; i < (size_t)nt * (dim + 1); ++i) FF_REQUIRE(conn[i] >= 0 && conn[i] < nv, "ffcuda_partition_local: conn[] entry outside [0, nv)");
    const LocalLists L = local_lists(dim + 1, nv, nt, conn, part, rank, nranks);
    const int64_t s[8] = {L.nowned, (int64_t)L.l2g.size() - L.nowned, (int64_t)L.elems.size(), (int64_t)L.nbr.size(),
                          (int64_t)L.send_idx.size(), 0, 0, 0};
    for (int i = 0; i < 8; ++i) sizes8[i] = s[i];
    auto put = [](int32_t *dst, const std::vector<int32_t> &v) {
        if (dst) std::copy(v.begin(), v.end(), dst);
    };
    put(l2g, L.l2g);
    put(elems, L.elems);
    put(nbr, L.nbr);
    put(recv_off, L.recv_off);
    put(recv_cnt, L.recv_cnt);
    put(send_ptr, L.send_ptr);
    put(send_idx, L.send_idx);
    FF_API_END(nullptr)
}

// The same for ANY element -> node table and node partition (P2: nloc = 10 nodes per tetrahedron, 6 per triangle; an edge
// node goes to the rank of one of its end points, so the local elements are those of the vertex partition).
extern "C" int ffcuda_partition_local_nodes(int nloc, int nnodes, int nt, const int32_t *elem2node, const int32_t *part, int rank, int nranks,
                                            int64_t *sizes8, int32_t *l2g, int32_t *elems, int32_t *nbr, int32_t *recv_off,
                                            int32_t *recv_cnt, int32_t *send_ptr, int32_t *send_idx)
{
    FF_API_BEGIN
    FF_REQUIRE(nloc >= 2 && nloc <= 10 && nnodes >= 0 && nt >= 0 && sizes8 && rank >= 0 && rank < nranks && (nt == 0 || elem2node) &&
                   (nnodes == 0 || part),
               "ffcuda_partition_local_nodes: bad arguments");
    for (int v = 0; v < nnodes; ++v) FF_REQUIRE(part[v] >= 0 && part[v] < nranks, "ffcuda_partition_local_nodes: part[] entry outside [0, nranks)");
    for (size_t i = 0; i < (size_t)nt * nloc; ++i)
        FF_REQUIRE(elem2node[i] >= 0 && elem2node[i] < nnodes, "ffcuda_partition_local_nodes: table entry outside [0, nnodes)");
    const LocalLists L = local_lists(nloc, nnodes, nt, elem2node, part, rank, nranks);
    const int64_t s[8] = {L.nowned, (int64_t)L.l2g.size() - L.nowned, (int64_t)L.elems.size(), (int64_t)L.nbr.size(),
                          (int64_t)L.send_idx.size(), 0, 0, 0};
    for (int i = 0; i < 8; ++i) sizes8[i] = s[i];
    auto put = [](int32_t *dst, const std::vector<int32_t> &v) {
        if (dst) std::copy(v.begin(), v.end(), dst);
    };
    put(l2g, L.l2g);
    put(elems, L.elems);
    put(nbr, L.nbr);
    put(recv_off, L.recv_off);
    put(recv_cnt, L.recv_cnt);
    put(send_ptr, L.send_ptr);
    put(send_idx, L.send_idx);
    FF_API_END(nullptr)
}

// A host CSR matrix shared out by contiguous row blocks (what the FreeFEM plugin does with a MatriceMorse to solve it on
// several GPUs; for cube / square meshes in FreeFEM's numbering the blocks are slabs): rank r owns the rows
// [n r / nranks, n (r+1) / nranks).  Its local problem: ghost columns = the columns of its rows outside its block, ascending
// (hence grouped by owner); what it sends to rank o = its columns that appear in o's rows, ascending (= o's ghost order).
// A rank is a neighbour when something goes in either direction (structure need not be symmetric).  Host arithmetic only.
// sizes8 = { owned rows, ghosts, local nnz, neighbours, total send count, first owned row, 0, 0 }; call once with the arrays
// NULL for the sizes.  lrowptr[owned+1] and lcolind[local nnz] are the local CSR structure (the values are the slice
// vals[rowptr[first] .. rowptr[first + owned]) of the caller's array, untouched).
extern "C" int ffcuda_partition_rows_local(int n, const int32_t *rowptr, const int32_t *colind, int rank, int nranks, int64_t *sizes8,
                                           int32_t *l2g, int32_t *lrowptr, int32_t *lcolind, int32_t *nbr, int32_t *recv_off,
                                           int32_t *recv_cnt, int32_t *send_ptr, int32_t *send_idx)
{
    FF_API_BEGIN
    FF_REQUIRE(n > 0 && rowptr && colind && sizes8 && nranks >= 1 && rank >= 0 && rank < nranks, "ffcuda_partition_rows_local: bad arguments");
    FF_REQUIRE(nranks <= n, "ffcuda_partition_rows_local: more ranks than rows");
    auto first = [&](int r) { return (int)((int64_t)n * r / nranks); };
    const int lo = first(rank), hi = first(rank + 1), nown = hi - lo;
    std::vector<int32_t> ghosts;
    for (int64_t k = rowptr[lo]; k < rowptr[hi]; ++k) {
        const int32_t j = colind[k];
        FF_REQUIRE(j >= 0 && j < n, "ffcuda_partition_rows_local: column index outside [0, n)");
        if (j < lo || j >= hi) ghosts.push_back(j);
    }
    std::sort(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
    std::vector<std::vector<int32_t>> send((size_t)nranks);
    {
        std::vector<int32_t> stamp((size_t)nown, -1);
        for (int o = 0; o < nranks; ++o) {
            if (o == rank) continue;
            for (int64_t k = rowptr[first(o)]; k < rowptr[first(o + 1)]; ++k) {
                const int32_t j = colind[k];
                if (j >= lo && j < hi && stamp[j - lo] != o) {
                    stamp[j - lo] = o;
                    send[o].push_back(j - lo);
                }
            }
            std::sort(send[o].begin(), send[o].end());
        }
    }
    std::vector<int32_t> vnbr, voff, vcnt, vsp(1, 0), vsi;
    size_t g = 0;
    for (int o = 0; o < nranks; ++o) {
        if (o == rank) continue;
        size_t g1 = g;
        while (g1 < ghosts.size() && ghosts[g1] < first(o + 1)) ++g1;
        if (g1 > g || !send[o].empty()) {
            vnbr.push_back(o);
            voff.push_back(nown + (int32_t)g);
            vcnt.push_back((int32_t)(g1 - g));
            vsi.insert(vsi.end(), send[o].begin(), send[o].end());
            vsp.push_back((int32_t)vsi.size());
        }
        g = g1;
    }
    FF_REQUIRE((int)vnbr.size() <= 16, "ffcuda_partition_rows_local: a row block has more than 16 neighbour blocks");
    const int64_t lnnz = (int64_t)rowptr[hi] - rowptr[lo];
    const int64_t sz[8] = {nown, (int64_t)ghosts.size(), lnnz, (int64_t)vnbr.size(), (int64_t)vsi.size(), lo, 0, 0};
    for (int i = 0; i < 8; ++i) sizes8[i] = sz[i];
    if (l2g) {
        for (int i = 0; i < nown; ++i) l2g[i] = lo + i;
        std::copy(ghosts.begin(), ghosts.end(), l2g + nown);
    }
    if (lrowptr)
        for (int i = 0; i <= nown; ++i) lrowptr[i] = rowptr[lo + i] - rowptr[lo];
    if (lcolind)
        for (int64_t k = 0; k < lnnz; ++k) {
            const int32_t j = colind[rowptr[lo] + k];
            lcolind[k] = (j >= lo && j < hi) ? j - lo : nown + (int32_t)(std::lower_bound(ghosts.begin(), ghosts.end(), j) - ghosts.begin());
        }
    auto put = [](int32_t *dst, const std::vector<int32_t> &v) {
        if (dst) std::copy(v.begin(), v.end(), dst);
    };
    put(nbr, vnbr);
    put(recv_off, voff);
    put(recv_cnt, vcnt);
    put(send_ptr, vsp);
    put(send_idx, vsi);
    FF_API_END(nullptr)
}
