// ffcuda.cpp — the FreeFEM side of the drop-in (layer B1): `load "ffcuda"`.
//
// A thin shim with no arithmetic of its own.  It is compiled against FreeFEM's headers (where they lie; nothing is
// copied) and does three things at load time:
//   1. registers, with preference 100, operators for   matrix A = va(Vh,Vh,...)  /  A = va(Vh,Vh,...)
//      (built-ins: OpMatrixtoBilinearForm, fflib/lgfem.cpp:6669,6673,6823,6826, pref 0);
//   2. the same for   real[int] b = va(0,Vh)  /  b = va(0,Vh)   (OpArraytoLinearForm, lgfem.cpp:6668,6672,6686,6688);
//   3. adds the solver "FFCUDACG" and re-points the name "CG" to it (TheFFSolver::ChangeSolver,
//      femlib/SparseLinearSolver.hpp:59-72), so that `solver=CG` and `A^-1*b` run on the GPU.
// At run time an intercepted call walks the varf exactly like AssembleVarForm (fflib/problem.cpp:9744-9855), flattens
// mesh / dof table / term lists / quadrature rule / Dirichlet sets into plain arrays and calls the C ABI of
// libffcuda_core.so (include/ffcuda.h).  What the GPU path does not cover (other elements, x-dependent coefficients,
// boundary integrals, level sets, complex, ...) is NOT claimed: the call is handed, untouched, to the
// built-in operator it derives from, with a notice at verbosity >= 1 (FFCUDA_STRICT=1 turns that into an error).
// There is no CPU re-implementation here: without a CUDA device every claimed call throws ErrorExec.
//
// Environment: FFCUDA_DEVICE (default 0), FFCUDA_VERBOSE=1 (say which path every call took), FFCUDA_STRICT=1,
//              FFCUDA_DISABLE=1 (register nothing).
#include "ff++.hpp"
#include <chrono>
#include <climits>
#include <sys/mman.h>
#include <thread>
#include "AFunction_ext.hpp"
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <typeinfo>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <vector>
#include "ffcuda.h"

using namespace Fem2D;

namespace {

// ------------------------------------------------------------------------------------------------------------
// context, errors
// ------------------------------------------------------------------------------------------------------------
ffcuda_ctx *g_ctx = nullptr;
bool env_on(const char *name)
{
    const char *v = getenv(name);
    return v && *v && *v != '0';
}
bool g_verbose = false, g_strict = false, g_check = false;
int g_sample_min = 2048, g_sample_n = 512; // grouping of coefficient tables: sampled above g_sample_min units, on ~g_sample_n of them

[[noreturn]] void fail(const std::string &what)
{
    const char *e = ffcuda_last_error(g_ctx);
    std::string msg = "ffcuda: " + what + ((e && *e) ? std::string(" : ") + e : std::string());
    ExecError(msg.c_str());
    throw ErrorExec(msg.c_str(), 1); // not reached (ExecError throws)
}
#define FFC(call)                        \
    do {                                 \
        if ((call) != 0) fail(#call);    \
    } while (0)

ffcuda_ctx *context()
{
    if (!g_ctx) {
        const char *d = getenv("FFCUDA_DEVICE");
        if (ffcuda_ctx_create(d ? atoi(d) : 0, &g_ctx) != 0) {
            g_ctx = nullptr;
            fail("cannot create a CUDA context (the ffcuda path has no CPU fallback)");
        }
    }
    return g_ctx;
}

// ------------------------------------------------------------------------------------------------------------
// host helpers: wall-clock marks (printed with FFCUDA_VERBOSE) and a chunked parallel loop for the O(mesh) / O(nnz) host
// passes of the hand-over (mesh flattening, row expansion, hash chains) - FreeFEM's interpreter is one thread, these
// loops only read FreeFEM's structures or write our own arrays
// ------------------------------------------------------------------------------------------------------------
struct Marks {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    std::string line;
    void mark(const char *what)
    {
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof buf, " %s %.1f ms", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        line += buf;
        t0 = t1;
    }
};
int host_threads()
{
    static int n = 0;
    if (!n) {
        const char *e = getenv("FFCUDA_HOST_THREADS");
        n = e ? atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (n < 1) n = 1;
    }
    return n;
}
template <class F>
void par_for(size_t n, F f) // f(begin, end) on disjoint chunks
{
    const int nt = (int)std::min<size_t>((size_t)host_threads(), std::max<size_t>(1, n / 65536));
    if (nt <= 1) {
        f((size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const size_t b = std::min(n, (size_t)t * chunk), e = std::min(n, b + chunk);
        if (b < e) th.emplace_back([=] { f(b, e); });
    }
    for (size_t t = 0; t < th.size(); ++t) th[t].join();
}

void notice(const char *what, const std::string &why)
{
    if (g_strict) ExecError(("ffcuda (FFCUDA_STRICT): " + std::string(what) + " not on the GPU path: " + why).c_str());
    if (g_verbose || verbosity > 0) cout << "  -- ffcuda: " << what << " left to FreeFEM (" << why << ")" << endl;
}

// ------------------------------------------------------------------------------------------------------------
// several GPUs from the one FreeFEM process (FFCUDA_NGPU=N): one context per device, one host thread per device while a
// distributed call runs (the library is one rank per context; ranks that are threads of one process share their peer
// mailboxes by pointer, csrc/comm.cu).  Used by the solvers: a MatriceMorse is shared out by contiguous row blocks
// (ffcuda_partition_rows_local), every GPU holds its rows, the CG / GMRES iterations exchange ghost values and all-reduce
// their dot products (reference role: MPI_Allreduce per dot product, plugin/mpi/MPICG.cpp:93-101).  The assembly statements
// stay on one GPU: their result has to end in FreeFEM's host matrix anyway.
// ------------------------------------------------------------------------------------------------------------
int g_ngpu = 1;          // FFCUDA_NGPU
long g_ngpu_min_n = 0;   // FFCUDA_NGPU_MIN_N: matrices with fewer rows are solved on one GPU
std::vector<ffcuda_ctx *> g_gang;

// f(rank) for every rank, rank 0 on the calling thread; the first error (if any) is raised after all of them returned
template <class F>
void on_ranks(int n, F f)
{
    std::vector<std::string> err((size_t)n);
    std::vector<std::thread> th;
    auto run = [&](int r) {
        try {
            f(r);
        } catch (const std::exception &e) { // (FreeFEM's ErrorExec is one)
            err[r] = e.what();
        } catch (...) {
            err[r] = "unknown error in rank thread";
        }
    };
    for (int r = 1; r < n; ++r) th.emplace_back(run, r);
    run(0);
    for (size_t t = 0; t < th.size(); ++t) th[t].join();
    for (int r = 0; r < n; ++r)
        if (!err[r].empty()) ExecError(("ffcuda (GPU " + std::to_string(r) + "): " + err[r]).c_str());
}

struct RankError : std::runtime_error {
    explicit RankError(const std::string &s) : std::runtime_error(s) {}
};
inline void rank_check(int rc, ffcuda_ctx *c, const char *what)
{
    if (rc == 0) return;
    const char *e = ffcuda_last_error(c);
    throw RankError(std::string(what) + ((e && *e) ? std::string(": ") + e : std::string()));
}

// the contexts of the gang, with their communicator; false (and a notice) when fewer devices answer
bool gang_ready()
{
    if (g_ngpu < 2) return false;
    if ((int)g_gang.size() == g_ngpu) return true;
    if (!g_gang.empty()) return false; // an earlier attempt failed: one GPU from then on
    const char *d = getenv("FFCUDA_DEVICE");
    const int base = d ? atoi(d) : 0;
    std::vector<ffcuda_ctx *> gang((size_t)g_ngpu, nullptr);
    gang[0] = context();
    bool ok = true;
    Marks mk;
    for (int r = 1; r < g_ngpu && ok; ++r) ok = ffcuda_ctx_create(base + r, &gang[r]) == 0;
    mk.mark("contexts");
    unsigned char id[128];
    ok = ok && ffcuda_comm_unique_id(id) == 0;
    if (ok) {
        try {
            on_ranks(g_ngpu, [&](int r) { rank_check(ffcuda_comm_init(gang[r], r, g_ngpu, id), gang[r], "ffcuda_comm_init"); });
        } catch (...) {
            ok = false;
        }
        mk.mark("communicator + peer mailboxes");
    }
    if (!ok) {
        for (int r = 1; r < g_ngpu; ++r)
            if (gang[r]) ffcuda_ctx_destroy(gang[r]);
        g_gang.assign(1, gang[0]); // marks the failed attempt
        g_ngpu = 1;
        notice("FFCUDA_NGPU", "the devices or their communicator are not available: one GPU");
        return false;
    }
    g_gang = gang;
    if (g_verbose)
        cout << "  -- ffcuda: " << g_ngpu << " GPUs driven from this process (one host thread per GPU during distributed solves):" << mk.line << endl;
    return true;
}

struct Unsupported {
    std::string why;
};

// ------------------------------------------------------------------------------------------------------------
// FE space on the device, cached per FESpace object (identity = address + UniqueffId, cf. problem.hpp:1678)
// ------------------------------------------------------------------------------------------------------------
struct DevSpace {
    const void *key = nullptr;
    UniqueffId uid;
    ffcuda_mesh *mesh = nullptr;
    ffcuda_space *space = nullptr;
    int dim = 0, order = 0, ncomp = 0, ndof = 0;
    ~DevSpace()
    {
        if (space) ffcuda_space_destroy(space);
        if (mesh) ffcuda_mesh_destroy(mesh);
    }
};
std::vector<std::unique_ptr<DevSpace>> g_spaces; // small most-recently-used list
const size_t kMaxSpaces = 4;

template <class MeshT>
struct MeshDim;
template <>
struct MeshDim<Mesh> {
    static const int d = 2;
};
template <>
struct MeshDim<Mesh3> {
    static const int d = 3;
};

inline void coords(const Mesh &Th, int i, double *p)
{
    p[0] = Th(i).x;
    p[1] = Th(i).y;
}
inline void coords(const Mesh3 &Th, int i, double *p)
{
    p[0] = Th(i).x;
    p[1] = Th(i).y;
    p[2] = Th(i).z;
}
inline int belem_of(const Mesh &Th, int ib, int &ie) { return Th.BoundaryElement(ib, ie); }
inline int belem_of(const Mesh3 &Th, int ib, int &ie) { return Th.BoundaryElement(ib, ie); }
inline int blabel(const Mesh &Th, int ib) { return Th.bedges[ib].lab; }
inline int blabel(const Mesh3 &Th, int ib) { return Th.be(ib).lab; }
inline int bvertex(const Mesh &Th, int ib, int j) { return Th(Th.bedges[ib][j]); }
inline int bvertex(const Mesh3 &Th, int ib, int j) { return Th(Th.be(ib)[j]); }
inline int nbe_of(const Mesh &Th) { return Th.neb; }
inline int nbe_of(const Mesh3 &Th) { return Th.nbe; }
inline int elabel(const Mesh &Th, int k) { return Th[k].lab; }
inline int elabel(const Mesh3 &Th, int k) { return Th[k].lab; }

// ------------------------------------------------------------------------------------------------------------
// meshes generated on the device (cube / buildlayers below): the device copy is kept until a fespace on the mesh adopts
// it, so that such a mesh is never flattened and uploaded.  Identity = address + sizes + a sample of coordinates and
// connectivity (a mesh object freed and another allocated at the same address does not pass).
// ------------------------------------------------------------------------------------------------------------
struct Generated {
    const void *th = nullptr;
    int nv = 0, nt = 0, nbe = 0;
    std::vector<double> sx;
    std::vector<int32_t> sc;
    ffcuda_mesh *dm = nullptr;
};
std::vector<Generated> g_generated;
const size_t kMaxGenerated = 2;
const int kSample = 64;

template <class MeshT>
void sample_mesh(const MeshT &Th, std::vector<double> &sx, std::vector<int32_t> &sc)
{
    const int dim = MeshDim<MeshT>::d, nvk = dim + 1;
    sx.clear();
    sc.clear();
    for (int q = 0; q < kSample; ++q) {
        const int i = (int)((long long)q * Th.nv / kSample), k = (int)((long long)q * Th.nt / kSample);
        double p[3];
        coords(Th, i, p);
        for (int d = 0; d < dim; ++d) sx.push_back(p[d]);
        for (int j = 0; j < nvk; ++j) sc.push_back(Th(k, j));
    }
}

template <class MeshT>
void keep_generated(const MeshT &Th, ffcuda_mesh *dm)
{
    Generated G;
    G.th = &Th;
    G.nv = Th.nv;
    G.nt = Th.nt;
    G.nbe = Th.nbe;
    sample_mesh(Th, G.sx, G.sc);
    G.dm = dm;
    g_generated.insert(g_generated.begin(), G);
    while (g_generated.size() > kMaxGenerated) {
        ffcuda_mesh_destroy(g_generated.back().dm);
        g_generated.pop_back();
    }
}

template <class MeshT>
ffcuda_mesh *take_generated(const MeshT &Th, int nbe)
{
    for (size_t x = 0; x < g_generated.size(); ++x) {
        Generated &G = g_generated[x];
        if (G.th != (const void *)&Th || G.nv != Th.nv || G.nt != Th.nt || G.nbe != nbe) continue;
        std::vector<double> sx;
        std::vector<int32_t> sc;
        sample_mesh(Th, sx, sc);
        if (sx != G.sx || sc != G.sc) continue;
        ffcuda_mesh *dm = G.dm;
        g_generated.erase(g_generated.begin() + x);
        return dm;
    }
    return nullptr;
}
inline ffcuda_mesh *take_generated(const Mesh &, int) { return nullptr; } // only 3-D meshes are generated here

// reference basis of P1/P2 Lagrange at a point (value only), in FreeFEM's local dof order: vertices, then edges
// ({01,02,03,12,13,23} on tetrahedra, edge opposite to vertex e on triangles)
void lagrange_values(int dim, int order, const double *l, double *phi)
{
    const int nv = dim + 1;
    if (order == 1) {
        for (int a = 0; a < nv; ++a) phi[a] = l[a];
        return;
    }
    for (int a = 0; a < nv; ++a) phi[a] = l[a] * (2 * l[a] - 1);
    static const int e3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}}, e2[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    const int ne = dim == 3 ? 6 : 3;
    for (int e = 0; e < ne; ++e) phi[nv + e] = 4 * l[dim == 3 ? e3[e][0] : e2[e][0]] * l[dim == 3 ? e3[e][1] : e2[e][1]];
}

inline void basis_values(const FElement &K, const R2 &P, KNMK<double> &val) // 2-D: one flag per operator (FESpace.hpp:840)
{
    bool whatd[last_operatortype];
    for (int i = 0; i < (int)last_operatortype; ++i) whatd[i] = false;
    whatd[op_id] = true;
    K.BF(whatd, P, val);
}
inline void basis_values(const FElement3 &K, const R3 &P, KNMK<double> &val) // 3-D: bit mask (FESpacen.hpp:451)
{
    K.BF(Fop_D0, P, val);
}
inline R2 hat_point(const Mesh *, const double *l) { return R2(l[1], l[2]); }
inline R3 hat_point(const Mesh3 *, const double *l) { return R3(l[1], l[2], l[3]); }

// Is Vh = [Pk]^N with Pk the P1 or P2 Lagrange element, numbered dof = node*N + c, local dof = c*nloc + a ?
// Decided from what the space does, not from its name: sizes, the dof table, and the values of its basis functions at
// an interior point of element 0 (a P1nc / P1b / P1dc space has other sizes or other values).
template <class FESpaceT>
void classify_space(const FESpaceT &Vh, int dim, int &order, int &ncomp, int &nloc)
{
    typedef typename FESpaceT::FElement FElementT;
    typedef typename FESpaceT::Mesh MeshT;
    ncomp = Vh.N;
    if (ncomp < 1 || ncomp > 3) throw Unsupported{"more than 3 components"};
    if (Vh.NbOfElements <= 0) throw Unsupported{"empty space"};
    const FElementT K0(Vh[0]);
    const int nd = K0.NbDoF();
    if (nd % ncomp) throw Unsupported{"element is not a product of identical components"};
    nloc = nd / ncomp;
    const int nv = dim + 1, n2 = dim == 3 ? 10 : 6;
    if (nloc == nv) order = 1;
    else if (nloc == n2) order = 2;
    else throw Unsupported{"finite element is neither P1 nor P2 Lagrange"};
    if ((long)Vh.NbOfNodes * ncomp != (long)Vh.NbOfDF) throw Unsupported{"dofs are not node*N + component"};
    // basis fingerprint
    double l[4] = {0.1, 0.2, 0.3, 0.4};
    if (dim == 2) {
        l[0] = 0.2;
        l[1] = 0.3;
        l[2] = 0.5;
    }
    double phi[10];
    lagrange_values(dim, order, l, phi);
    KNMK<double> val(nd, ncomp, (int)last_operatortype);
    val = 0.;
    basis_values(K0, hat_point((const MeshT *)0, l), val);
    for (int c = 0; c < ncomp; ++c)
        for (int a = 0; a < nloc; ++a)
            for (int c2 = 0; c2 < ncomp; ++c2) {
                const double expect = (c == c2) ? phi[a] : 0.0;
                if (fabs(val(c * nloc + a, c2, (int)op_id) - expect) > 1e-12) throw Unsupported{"basis functions are not [Pk Lagrange]^N, component-major"};
            }
    // dof table layout on a sample of elements
    const int nt = Vh.NbOfElements, step = std::max(1, nt / 64);
    for (int k = 0; k < nt; k += step) {
        const FElementT K(Vh[k]);
        if (K.NbDoF() != nd) throw Unsupported{"variable number of dofs per element"};
        for (int a = 0; a < nloc; ++a) {
            const int d0 = K(a);
            if (d0 % ncomp) throw Unsupported{"dof numbering is not node*N + component"};
            for (int c = 1; c < ncomp; ++c)
                if (K(c * nloc + a) != d0 + c) throw Unsupported{"dof numbering is not node*N + component"};
        }
    }
}

template <class FESpaceT>
DevSpace &device_space(const FESpaceT &Vh)
{
    typedef typename FESpaceT::FElement FElementT;
    typedef typename FESpaceT::Mesh MeshT;
    const int dim = MeshDim<MeshT>::d;
    const MeshT &Th = Vh.Th;
    for (size_t i = 0; i < g_spaces.size(); ++i)
        if (g_spaces[i]->key == (const void *)&Vh && g_spaces[i]->uid == (const UniqueffId &)Vh) {
            if (i) std::swap(g_spaces[0], g_spaces[i]);
            return *g_spaces[0];
        }
    int order, ncomp, nloc;
    classify_space(Vh, dim, order, ncomp, nloc);
    ffcuda_ctx *ctx = context();
    const int nv = Th.nv, nt = Th.nt, nbe = nbe_of(Th), nvk = dim + 1;
    Marks mk;
    std::unique_ptr<DevSpace> D(new DevSpace());
    D->mesh = take_generated(Th, nbe); // generated on the device by this plugin: already there
    if (D->mesh) mk.mark("mesh already on the device");
    else {
    std::vector<double> xyz((size_t)nv * dim);
    par_for((size_t)nv, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) coords(Th, (int)i, &xyz[i * dim]);
    });
    std::vector<int32_t> conn((size_t)nt * nvk), elab(nt), bconn((size_t)nbe * dim), blab(nbe), belem(nbe), bface(nbe);
    par_for((size_t)nt, [&](size_t b, size_t e) {
        for (size_t k = b; k < e; ++k) {
            for (int j = 0; j < nvk; ++j) conn[k * nvk + j] = Th((int)k, j);
            elab[k] = elabel(Th, (int)k);
        }
    });
    for (int ib = 0; ib < nbe; ++ib) {
        int ie;
        belem[ib] = belem_of(Th, ib, ie);
        bface[ib] = ie;
        blab[ib] = blabel(Th, ib);
        for (int j = 0; j < dim; ++j) bconn[(size_t)ib * dim + j] = bvertex(Th, ib, j);
    }
    mk.mark("mesh flattened");
    FFC(ffcuda_mesh_upload(ctx, dim, nv, xyz.data(), nt, conn.data(), elab.data(), nbe, bconn.data(), blab.data(), belem.data(),
                           bface.data(), &D->mesh));
    mk.mark("uploaded");
    }
    D->key = &Vh;
    D->uid = (const UniqueffId &)Vh;
    D->dim = dim;
    D->order = order;
    D->ncomp = ncomp;
    D->ndof = Vh.NbOfDF;
    // the node table as FreeFEM numbered it (2-D P2 is renumbered by FreeFEM: never guessed, always read)
    std::vector<int32_t> e2n;
    const int32_t *pe2n = nullptr;
    bool p1_vertex_numbering = (order == 1);
    if (order == 1) {
        for (int k = 0; k < nt && p1_vertex_numbering; k += std::max(1, nt / 256)) {
            const FElementT K(Vh[k]);
            for (int a = 0; a < nloc; ++a)
                if (K(a) / ncomp != Th(k, a)) p1_vertex_numbering = false;
        }
    }
    if (!p1_vertex_numbering) {
        e2n.resize((size_t)nt * nloc);
        for (int k = 0; k < nt; ++k) {
            const FElementT K(Vh[k]);
            for (int a = 0; a < nloc; ++a) e2n[(size_t)k * nloc + a] = K(a) / ncomp;
        }
        pe2n = e2n.data();
    }
    FFC(ffcuda_space_create(D->mesh, order, ncomp, pe2n, Vh.NbOfNodes, &D->space));
    int ndof = 0;
    FFC(ffcuda_space_info(D->space, &ndof, nullptr, nullptr));
    if (ndof != Vh.NbOfDF) fail("internal: device space has another number of dofs than the fespace");
    mk.mark("space");
    if (g_verbose) cout << "  -- ffcuda: fespace on the device (" << host_threads() << " host threads):" << mk.line << endl;
    g_spaces.insert(g_spaces.begin(), std::move(D));
    if (g_spaces.size() > kMaxSpaces) g_spaces.pop_back();
    return *g_spaces[0];
}

// ------------------------------------------------------------------------------------------------------------
// reading a varf
// ------------------------------------------------------------------------------------------------------------
struct Quad {
    std::vector<double> pts, w;
};
inline void hat_coords(const R2 &P, double *c)
{
    c[0] = P.x;
    c[1] = P.y;
}
inline void hat_coords(const R3 &P, double *c)
{
    c[0] = P.x;
    c[1] = P.y;
    c[2] = P.z;
}
template <class QF>
Quad flat_quadrature(const QF &q, int dim)
{
    Quad Q;
    for (int i = 0; i < q.n; ++i) {
        const typename QF::QuadraturePoint &p = q[i];
        Q.w.push_back(p.a);
        double c[3];
        hat_coords((const typename QF::Rd &)p, c); // (x,y[,z]) of the reference point
        for (int d = 0; d < dim; ++d) Q.pts.push_back(c[d]);
    }
    return Q;
}
inline Quad volume_rule(Stack s, const CDomainOfIntegration &di, const Mesh *) { return flat_quadrature(di.FIT(s), 2); }
inline Quad volume_rule(Stack s, const CDomainOfIntegration &di, const Mesh3 *) { return flat_quadrature(di.FIV(s), 3); }
// rule of a boundary integral in the face coordinates ffcuda_assemble_*_boundary expects (P = A(1-x-y) + Bx + Cy on a face,
// P = A(1-x) + Bx on an edge).  3-D: T.PBord(ie, pi) is exactly that; 2-D: Element_rhs / Element_Op take
// Pt = PA*pi.x + PB*(1-pi.x) (fflib/problem.cpp:8646-8649, :6229-6232), hence 1 - x
inline Quad border_rule(Stack s, const CDomainOfIntegration &di, const Mesh3 *) { return flat_quadrature(di.FIT(s), 2); }
inline Quad border_rule(Stack s, const CDomainOfIntegration &di, const Mesh *)
{
    const QuadratureFormular1d &q = di.FIE(s);
    Quad Q;
    for (int i = 0; i < q.n; ++i) {
        Q.w.push_back(q[i].a);
        Q.pts.push_back(1.0 - q[i].x);
    }
    return Q;
}

struct Region {
    bool all = true;
    std::vector<int32_t> labels;
};
Region region_of(Stack stack, const CDomainOfIntegration &di, size_t maxlab = 16) // Expandsetoflab, fflib/lgfem.cpp:7809
{
    Region R;
    std::set<int> s;
    for (size_t i = 0; i < di.what.size(); ++i) {
        R.all = false;
        if (di.whatis[i] == 0) s.insert((int)GetAny<long>((*di.what[i])(stack)));
        else {
            KN<long> labs(GetAny<KN_<long>>((*di.what[i])(stack)));
            for (long j = 0; j < labs.N(); ++j) s.insert((int)labs[j]);
        }
    }
    R.labels.assign(s.begin(), s.end());
    if (R.labels.size() > maxlab) throw Unsupported{"too many region / boundary labels in one integral"};
    return R;
}

template <class MeshT>
void check_domain_options(Stack stack, const CDomainOfIntegration &di, const MeshT &Th);
// returns true for an integral over boundary elements (int2d on a mesh3, int1d on a mesh), false for a volume integral
template <class MeshT>
bool check_domain(Stack stack, const CDomainOfIntegration &di, const MeshT &Th)
{
    const int dim = MeshDim<MeshT>::d;
    const CDomainOfIntegration::typeofkind volume = dim == 3 ? CDomainOfIntegration::int3d : CDomainOfIntegration::int2d;
    const CDomainOfIntegration::typeofkind border = dim == 3 ? CDomainOfIntegration::int2d : CDomainOfIntegration::int1d;
    if (di.d != dim || di.dHat != dim || (di.kind != volume && di.kind != border))
        throw Unsupported{"neither a volume nor a boundary integral on the mesh of the space"};
    check_domain_options(stack, di, Th);
    return di.kind == border;
}
template <class MeshT>
void check_domain_options(Stack stack, const CDomainOfIntegration &di, const MeshT &Th)
{
    if (di.islevelset()) throw Unsupported{"level-set integral"};
    if (di.withmap()) throw Unsupported{"mapped integration points"};
    typedef const MeshT *pm;
    if (GetAny<pm>((*di.Th)(stack)) != &Th) throw Unsupported{"integral on another mesh"};
}

inline int check_op(int op, int dim)
{
    if (op == op_id || op == op_dx || op == op_dy || (op == op_dz && dim == 3)) return op;
    throw Unsupported{"differential operator other than id, dx, dy, dz"};
}

double constant_coef(Stack stack, const C_F0 &c)
{
    if (!c.LeftValue()->MeshIndependent()) throw Unsupported{"coefficient depends on the mesh point"};
    if (c.left() != atype<double>()) {
        if (c.left() == atype<long>()) return (double)GetAny<long>(c.eval(stack));
        throw Unsupported{"coefficient is not real"};
    }
    return GetAny<double>(c.eval(stack));
}

struct QBTerm { // term of a bilinear form whose coefficient depends on the mesh point
    ffcuda_bterm t; // (coef unused)
    C_F0 coef;
};
struct BilinearItem {
    std::vector<ffcuda_bterm> terms;
    std::vector<QBTerm> qterms;
    Quad q;
    Region reg;
    bool border = false; // integral over the boundary elements with the labels of reg (Robin terms)
};
struct QTerm { // term whose coefficient depends on the mesh point: evaluated at the quadrature nodes by FreeFEM's evaluator
    int vcomp;
    int slot; // 0 = value of the test function, 1..dim = dx, dy, dz
    C_F0 coef;
};
struct LinearItem {
    std::vector<ffcuda_lterm> terms;
    std::vector<QTerm> qterms;
    Quad q;
    Region reg;
    bool border = false; // Neumann / traction data
};
struct BCItem {
    std::vector<int32_t> labels;
    int compmask = 0;
    double values[3] = {0, 0, 0};
    // on(..., u = g(x,y,z)): the (dof, value) pairs AssembleBC would produce, in its order (later pairs win)
    bool pairs = false;
    std::vector<int32_t> pdofs;
    std::vector<double> pvals;
};
ffcuda_bc *make_bc(ffcuda_space *space, const BCItem &B)
{
    ffcuda_bc *bc = nullptr;
    if (B.pairs) FFC(ffcuda_bc_from_pairs(space, (int)B.pdofs.size(), B.pdofs.data(), B.pvals.data(), &bc));
    else FFC(ffcuda_bc_from_labels(space, (int)B.labels.size(), B.labels.data(), B.compmask, B.values, &bc));
    return bc;
}
struct Varf {
    std::vector<BilinearItem> bil;
    std::vector<LinearItem> lin;
    std::vector<BCItem> bc;
    bool other_rhs_items = false; // arrays, A*x, ... (only meaningful for the right-hand side)
};

// Dirichlet data depending on the mesh point, on(labels, u = g(x,y,z)): AssembleBC (fflib/problem.cpp:9881-10034 2-D,
// :10039-10194 3-D) visits the boundary elements in order and, on each one whose label is listed, evaluates g at the
// interpolation point of every dof lying on that face (for P1 / P2 Lagrange: the node itself) with the mesh point set to
// that boundary element (label, unit normal); a later boundary element overwrites an earlier one.  The same here, through
// FreeFEM's evaluator; the pairs go to ffcuda_bc_from_pairs.  O(boundary) evaluations.
inline R3 unit_normal(const Mesh3 &, const Tet &T, int ie)
{
    R3 NN = T.N(ie);
    NN /= NN.norme();
    return NN;
}
inline R2 unit_normal(const Mesh &, const Triangle &T, int ie)
{
    R2 E = T.Edge(ie);
    const double le = sqrt((E, E));
    return R2(E.y, -E.x) / le;
}
template <class FESpaceT>
void dirichlet_pairs(Stack stack, const FESpaceT &Vh, const BC_set *bc, BCItem &B)
{
    typedef typename FESpaceT::Mesh MeshT;
    typedef typename FESpaceT::FElement FElementT;
    typedef typename MeshT::RdHat RdHat;
    const MeshT &Th = Vh.Th;
    const int dim = MeshDim<MeshT>::d;
    int order, ncomp, nloc;
    classify_space(Vh, dim, order, ncomp, nloc);
    static const int edge3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}}; // Mesh3dn.cpp:73
    static const int edge2[3][2] = {{1, 2}, {2, 0}, {0, 1}};                       // dof 3+e on the edge opposite vertex e
    std::vector<Expression> ex((size_t)ncomp, (Expression)0);
    for (size_t k = 0; k < bc->bc.size(); ++k) ex[bc->bc[k].first] = bc->bc[k].second;
    std::set<int> on(B.labels.begin(), B.labels.end());
    MeshPoint *mps = MeshPointStack(stack), mp = *mps;
    B.pairs = true;
    try {
        const int nbe = nbe_of(Th);
        for (int ib = 0; ib < nbe; ++ib) {
            int ie;
            const int it = belem_of(Th, ib, ie);
            const int r = blabel(Th, ib);
            if (!on.count(r)) continue;
            const FElementT K(Vh[it]);
            const typename MeshT::Rd NN = unit_normal(Th, K.T, ie);
            for (int a = 0; a < nloc; ++a) {
                double l[4] = {0, 0, 0, 0}; // barycentric coordinates of local node a
                bool onface;
                if (a <= dim) {
                    l[a] = 1.0;
                    onface = a != ie;
                } else {
                    const int e = a - (dim + 1);
                    const int i0 = dim == 3 ? edge3[e][0] : edge2[e][0], i1 = dim == 3 ? edge3[e][1] : edge2[e][1];
                    l[i0] = l[i1] = 0.5;
                    onface = i0 != ie && i1 != ie;
                }
                if (!onface) continue;
                const RdHat PtHat(hat_point((const MeshT *)0, l));
                mps->set(K.T(PtHat), PtHat, K, r, NN, ie);
                for (int c = 0; c < ncomp; ++c)
                    if (B.compmask >> c & 1) {
                        B.pdofs.push_back((int32_t)K(c * nloc + a));
                        B.pvals.push_back(GetAny<double>((*ex[c])(stack)));
                    }
            }
        }
    } catch (...) {
        *mps = mp;
        throw;
    }
    *mps = mp;
}

template <class MeshT, class FESpaceT>
Varf read_varf(Stack stack, const list<C_F0> &largs, const MeshT &Th, int ncomp, bool want_matrix, const FESpaceT &Vh)
{
    const int dim = MeshDim<MeshT>::d;
    Varf V;
    for (list<C_F0>::const_iterator ii = largs.begin(); ii != largs.end(); ++ii) {
        Expression e = ii->LeftValue();
        aType r = ii->left();
        if (r == atype<const FormBilinear *>()) {
            if (!want_matrix) continue; // ignored when a right-hand side is assembled (problem.cpp:9761-9779)
            const FormBilinear *bf = dynamic_cast<const FormBilinear *>(e);
            if (bf->VF()) throw Unsupported{"discontinuous-Galerkin operators"};
            BilinearItem B;
            B.border = check_domain(stack, *bf->di, Th);
            B.q = B.border ? border_rule(stack, *bf->di, &Th) : volume_rule(stack, *bf->di, &Th);
            B.reg = region_of(stack, *bf->di, B.border ? 32 : 16);
            const Foperator &op = *bf->b;
            for (size_t k = 0; k < op.v.size(); ++k) {
                const pair<MGauche, MDroit> &id = op.v[k].first; // (unknown, test)
                ffcuda_bterm t;
                t.ucomp = id.first.first;
                t.uop = check_op(id.first.second, dim);
                t.vcomp = id.second.first;
                t.vop = check_op(id.second.second, dim);
                if (t.ucomp < 0 || t.ucomp >= ncomp || t.vcomp < 0 || t.vcomp >= ncomp) throw Unsupported{"component out of range"};
                // (derivatives in a boundary integral - dx(u) v, N.x dx(u) v, ... - reach every node of the adjacent element: the
                //  library takes its general path for them, fflib/problem.cpp:6518-6560)
                if (!op.v[k].second.LeftValue()->MeshIndependent()) {
                    // kappa(x,y,z) grad u . grad v, rho(x) u v, a P0 / P1 function as coefficient: evaluated at the quadrature
                    // nodes by FreeFEM's evaluator, integrated on the device (P1 spaces; checked in gpu_matrix)
                    if (op.v[k].second.left() != atype<double>() && op.v[k].second.left() != atype<long>())
                        throw Unsupported{"coefficient is not real"};
                    t.coef = 0.0;
                    B.qterms.push_back(QBTerm{t, op.v[k].second});
                    continue;
                }
                t.coef = constant_coef(stack, op.v[k].second);
                B.terms.push_back(t);
            }
            V.bil.push_back(B);
        } else if (r == atype<const FormLinear *>()) {
            if (want_matrix) continue;
            const FormLinear *lf = dynamic_cast<const FormLinear *>(e);
            if (lf->VF()) throw Unsupported{"discontinuous-Galerkin operators"};
            LinearItem L;
            L.border = check_domain(stack, *lf->di, Th);
            L.q = L.border ? border_rule(stack, *lf->di, &Th) : volume_rule(stack, *lf->di, &Th);
            L.reg = region_of(stack, *lf->di, L.border ? 32 : 16);
            const Ftest &op = *lf->l;
            for (size_t k = 0; k < op.v.size(); ++k) {
                ffcuda_lterm t;
                t.vcomp = op.v[k].first.first;
                t.vop = check_op(op.v[k].first.second, dim);
                if (t.vcomp < 0 || t.vcomp >= ncomp) throw Unsupported{"component out of range"};
                if (!op.v[k].second.LeftValue()->MeshIndependent()) {
                    if (L.border && t.vop != op_id) throw Unsupported{"derivative of the test function times mesh-dependent data in a boundary integral"};
                    // f(x,y,z) v, uold v / dt, ...: the values Element_rhs would compute go to the device as a table
                    if (op.v[k].second.left() != atype<double>() && op.v[k].second.left() != atype<long>())
                        throw Unsupported{"coefficient is not real"};
                    const int slot = t.vop == op_id ? 0 : t.vop == op_dx ? 1 : t.vop == op_dy ? 2 : 3;
                    L.qterms.push_back(QTerm{t.vcomp, slot, op.v[k].second});
                    continue;
                }
                t.coef = constant_coef(stack, op.v[k].second);
                L.terms.push_back(t);
            }
            V.lin.push_back(L);
        } else if (r == atype<const BC_set *>()) {
            const BC_set *bc = dynamic_cast<const BC_set *>(e);
            if (bc->complextype) throw Unsupported{"complex boundary value"};
            BCItem B;
            std::set<long> on; // Expandsetoflab, fflib/lgfem.cpp:7792
            for (size_t i = 0; i < bc->on.size(); ++i)
                if (bc->onis[i] == 0) on.insert(GetAny<long>((*bc->on[i])(stack)));
                else {
                    KN<long> labs(GetAny<KN_<long>>((*bc->on[i])(stack)));
                    for (long j = 0; j < labs.N(); ++j) on.insert(labs[j]);
                }
            if (on.size() > 32) throw Unsupported{"more than 32 labels in one on(...)"};
            for (std::set<long>::const_iterator it = on.begin(); it != on.end(); ++it) B.labels.push_back((int32_t)*it);
            bool varying = false;
            for (size_t k = 0; k < bc->bc.size(); ++k) {
                const int comp = bc->bc[k].first;
                if (comp < 0 || comp >= ncomp) throw Unsupported{"boundary condition on a component out of range"};
                B.compmask |= 1 << comp;
                if (!bc->bc[k].second->MeshIndependent()) varying = true;
            }
            if (ncomp > 1 && B.compmask != (1 << ncomp) - 1)
                throw Unsupported{"vector space with a boundary condition on some components only"};
            if (varying) {
                if (!B.labels.empty()) dirichlet_pairs(stack, Vh, bc, B);
            } else
                for (size_t k = 0; k < bc->bc.size(); ++k) B.values[bc->bc[k].first] = GetAny<double>((*bc->bc[k].second)(stack));
            if (!B.labels.empty()) V.bc.push_back(B);
        } else {
            if (want_matrix) throw Unsupported{"varf item other than integrals and on(...)"};
            V.other_rhs_items = true;
        }
    }
    return V;
}

void apply_bcs(DevSpace &D, const Varf &V, ffcuda_matrix *A, ffcuda_vec *b, double tgv)
{
    for (size_t i = 0; i < V.bc.size(); ++i) {
        const BCItem &B = V.bc[i];
        ffcuda_bc *bc = make_bc(D.space, B);
        int rc = 0;
        if (A) rc |= ffcuda_matrix_apply_bc(A, bc, tgv);
        if (b) rc |= ffcuda_vec_apply_bc(b, bc, tgv);
        ffcuda_bc_destroy(bc);
        if (rc) fail("applying the Dirichlet conditions");
    }
}

// ------------------------------------------------------------------------------------------------------------
// device copies of matrices, waiting for the solver that will be attached to them
// ------------------------------------------------------------------------------------------------------------
struct Resident {
    ffcuda_matrix *A;
    ffcuda_pattern *P; // the matrix refers to the pattern's arrays: destroyed after it
};
void release_resident(Resident &r)
{
    if (r.A) ffcuda_matrix_destroy(r.A);
    if (r.P) ffcuda_pattern_destroy(r.P);
    r.A = nullptr;
    r.P = nullptr;
}
std::map<const void *, Resident> g_resident; // HashMatrix* -> device matrix just assembled
void drop_resident()
{
    for (std::map<const void *, Resident>::iterator it = g_resident.begin(); it != g_resident.end(); ++it) release_resident(it->second);
    g_resident.clear();
}

// bilinear terms whose coefficient depends on the mesh point cost a pass of FreeFEM's evaluator over the quadrature nodes:
// the space is classified first, on the host, so that a refusal costs nothing
template <class FESpaceT>
void check_qterms_supported(const FESpaceT &Vh, const Varf &V)
{
    bool any = false;
    for (size_t i = 0; i < V.bil.size(); ++i) any = any || (!V.bil[i].qterms.empty() && !V.bil[i].border);
    if (!any) return;
    int order, ncomp, nloc;
    classify_space(Vh, MeshDim<typename FESpaceT::Mesh>::d, order, ncomp, nloc); // refuses what is not P1 / P2 Lagrange
}

// the pattern of the device matrix is that of the whole space: FreeFEM's is the same only when the volume integrals
// visit every element (HashMatrix creates the couples of the visited elements only)
template <class MeshT>
void check_full_pattern(const Varf &V, const MeshT &Th)
{
    bool full = false;
    std::set<int> labs;
    for (size_t i = 0; i < V.bil.size(); ++i)
        if (!V.bil[i].border) {
            full = full || V.bil[i].reg.all;
            labs.insert(V.bil[i].reg.labels.begin(), V.bil[i].reg.labels.end());
        }
    for (int k = 0; k < Th.nt && !full; ++k)
        if (!labs.count(Th[k].lab)) throw Unsupported{"the volume integrals do not visit every element (sub-pattern)"};
}

// ------------------------------------------------------------------------------------------------------------
// FE functions as data of a form, shipped as DOF ARRAYS: f in int3d(Th)(f*v), kappa in
// int3d(Th)(kappa*(dx(u)*dx(v)+...)), uk in the Newton residual int3d(Th)(dx(uk)*dx(v)+...).  A coefficient that IS the
// value or a first derivative of a P0 / P1 / P2 Lagrange function living on the mesh of the form compiles to the node
// E_F0_Func1(pfer2R<R,op> | pf3r2R<R,op,v_fes3>, <the FE function>) (fflib/lgfem.cpp:2053-2088, 6993-7000,
// fflib/lgmesh3.cpp:2169, 3094, 3136-3138; `f*v` keeps the node as it is: operator*(C_F0,C_F0) drops the constant 1,
// fflib/AFunction.hpp:3159).  Such a term needs no interpreter call per quadrature node: the dof array goes to the device
// and ffcuda_fe_table forms the very sums FElement::operator()(PHat,u,comp,op) forms.  The functions behind the nodes are
// not exported by name, so their addresses are read, at load time, from the nodes FreeFEM itself compiles for a dummy FE
// function.  Anything else (uold/dt, f*g, functions on another mesh, other elements) stays on the interpreter path above.
// FFCUDA_NO_FE_DOFS=1 switches the recognition off.
// ------------------------------------------------------------------------------------------------------------
template <class MeshT>
struct FeTypesOf;
template <>
struct FeTypesOf<Mesh> {
    typedef v_fes vfes;
};
template <>
struct FeTypesOf<Mesh3> {
    typedef v_fes3 vfes;
};
struct FeNodeFunctions {
    Function1 f[4]; // value, dx, dy, dz (0: not found)
};
FeNodeFunctions g_fe2 = {{0, 0, 0, 0}}, g_fe3 = {{0, 0, 0, 0}};
bool g_fe_dofs = true, g_explain = false;
bool g_rect = false; // FFCUDA_RECT=1: `matrix B = vb(Uh,Vh)` with two different spaces goes to the device as well (see rect_form)
template <class PF>
void find_fe_node_functions(FeNodeFunctions &F, int dim)
{
    static const char *names[4] = {nullptr, "dx", "dy", "dz"};
    for (int k = 0; k <= dim; ++k) {
        F.f[k] = 0;
        try {
            C_F0 dummy(CConstant<PF>(PF((typename PF::first_type)0, 0))); // never evaluated
            C_F0 r;
            if (k == 0) r = atype<double>()->CastTo(dummy);
            else {
                C_F0 g = Global.Find(names[k]);
                const Polymorphic *pop = dynamic_cast<const Polymorphic *>(g.LeftValue());
                if (!pop) continue;
                r = C_F0(pop, "(", dummy);
            }
            if (r.left() != atype<double>()) continue;
            const E_F0_Func1 *n = dynamic_cast<const E_F0_Func1 *>(r.LeftValue());
            if (n) F.f[k] = n->f;
        } catch (...) { // (no such cast / overload in this FreeFEM: the term stays on the interpreter path)
        }
    }
}
inline const FeNodeFunctions &fe_node_functions(const Mesh *) { return g_fe2; }
inline const FeNodeFunctions &fe_node_functions(const Mesh3 *) { return g_fe3; }

// the FE functions met in one statement: dof array and node table go to the device once per function
template <class FESpaceT>
struct FeFunctions {
    struct Fun {
        const KN<double> *x;
        const FESpaceT *Wh;
        int order, ncomp, nloc;
        bool default_numbering;       // P0: node = element, P1: node = vertex (no table needed)
        std::vector<int32_t> e2n;     // nt x nloc otherwise
        ffcuda_vec *dofs;
    };
    std::vector<Fun> funs;
    ~FeFunctions()
    {
        for (size_t i = 0; i < funs.size(); ++i)
            if (funs[i].dofs) ffcuda_vec_destroy(funs[i].dofs);
    }
    ffcuda_vec *on_device(int i)
    {
        Fun &F = funs[i];
        if (!F.dofs) {
            FFC(ffcuda_vec_create(context(), (int)F.x->N(), &F.dofs));
            const double *src = (const double *)*F.x;
            std::vector<double> packed;
            if (F.x->step != 1) { // (a KN owns contiguous storage; kept for safety)
                packed.resize((size_t)F.x->N());
                for (long j = 0; j < F.x->N(); ++j) packed[j] = (*F.x)[j];
                src = packed.data();
            }
            if (ffcuda_vec_upload(F.dofs, src) != 0) fail("uploading the dofs of an FE function");
        }
        return F.dofs;
    }
};
struct FeRef {
    int fun, comp, op; // op: FFCUDA_OP_*
    bool operator==(const FeRef &o) const { return fun == o.fun && comp == o.comp && op == o.op; }
};

// Is Wh a P0 space as FreeFEM numbers it (one dof per element, dof = element, basis = 1) ?
template <class FESpaceT>
bool is_p0_space(const FESpaceT &Wh)
{
    typedef typename FESpaceT::FElement FElementT;
    typedef typename FESpaceT::Mesh MeshT;
    if (Wh.N != 1 || Wh.NbOfElements <= 0 || (long)Wh.NbOfDF != (long)Wh.Th.nt) return false;
    const int nt = Wh.Th.nt, step = std::max(1, nt / 256);
    for (int k = 0; k < nt; k += step) {
        const FElementT K(Wh[k]);
        if (K.NbDoF() != 1 || K(0) != k) return false;
    }
    const FElementT K0(Wh[0]);
    KNMK<double> val(1, 1, (int)last_operatortype);
    val = 0.;
    double l[4] = {0.1, 0.2, 0.3, 0.4};
    if (MeshDim<MeshT>::d == 2) l[0] = 0.2, l[1] = 0.3, l[2] = 0.5;
    basis_values(K0, hat_point((const MeshT *)0, l), val);
    return fabs(val(0, 0, (int)op_id) - 1.0) <= 1e-14;
}

// c = value / dx / dy / dz of an FE function we can evaluate on the device ?  (no device call here)
template <class FESpaceT>
bool fe_reference(Stack stack, const C_F0 &c, const FESpaceT &Vh, FeFunctions<FESpaceT> &cache, FeRef &out)
{
    typedef typename FESpaceT::Mesh MeshT;
    typedef typename FESpaceT::FElement FElementT;
    typedef typename FeTypesOf<MeshT>::vfes vfes;
    typedef pair<FEbase<double, vfes> *, int> PF;
    const int dim = MeshDim<MeshT>::d;
    if (!g_fe_dofs || c.left() != atype<double>()) return false;
    const E_F0_Func1 *node = dynamic_cast<const E_F0_Func1 *>(c.LeftValue());
    if (!node || !node->f || !node->a) return false;
    const FeNodeFunctions &F = fe_node_functions((const MeshT *)0);
    static const int ops[4] = {FFCUDA_OP_ID, FFCUDA_OP_DX, FFCUDA_OP_DY, FFCUDA_OP_DZ};
    int k = -1;
    for (int j = 0; j <= dim; ++j)
        if (F.f[j] && node->f == F.f[j]) k = j;
    if (k < 0) return false;
    PF pp = GetAny<PF>((*node->a)(stack));
    FEbase<double, vfes> *fe = pp.first;
    if (!fe || !fe->x() || !fe->Vh) return false;
    const FESpaceT *Wh = &*fe->Vh;
    if (&Wh->Th != &Vh.Th) return false; // a function on another mesh is interpolated by FreeFEM: interpreter path
    if ((long)fe->x()->N() != (long)Wh->NbOfDF) return false;
    const KN<double> *x = fe->x();
    int fi = -1;
    for (size_t i = 0; i < cache.funs.size(); ++i)
        if (cache.funs[i].x == x && cache.funs[i].Wh == Wh) fi = (int)i;
    if (fi < 0) {
        typename FeFunctions<FESpaceT>::Fun N;
        N.x = x;
        N.Wh = Wh;
        N.dofs = nullptr;
        if (is_p0_space(*Wh)) {
            N.order = 0;
            N.ncomp = 1;
            N.nloc = 1;
            N.default_numbering = true;
        } else {
            try {
                classify_space(*Wh, dim, N.order, N.ncomp, N.nloc);
            } catch (const Unsupported &) {
                return false;
            }
            const int nt = Wh->Th.nt;
            N.default_numbering = N.order == 1;
            N.e2n.resize((size_t)nt * N.nloc);
            par_for((size_t)nt, [&](size_t b, size_t e) {
                for (size_t kk = b; kk < e; ++kk) {
                    const FElementT K((*Wh)[(int)kk]);
                    for (int a = 0; a < N.nloc; ++a) N.e2n[kk * N.nloc + a] = K(a) / N.ncomp;
                }
            });
            for (int kk = 0; kk < nt && N.default_numbering; ++kk)
                for (int a = 0; a < N.nloc; ++a)
                    if (N.e2n[(size_t)kk * N.nloc + a] != Wh->Th(kk, a)) N.default_numbering = false;
            if (N.default_numbering) std::vector<int32_t>().swap(N.e2n);
        }
        cache.funs.push_back(N);
        fi = (int)cache.funs.size() - 1;
    }
    if (pp.second < 0 || pp.second >= cache.funs[fi].ncomp) return false;
    out.fun = fi;
    out.comp = pp.second;
    out.op = ops[k];
    return true;
}

// table[offset + unit*nq + q] += scale * (datum at node q of the unit), on the device
template <class FESpaceT>
void fe_table_add(DevSpace &D, FeFunctions<FESpaceT> &cache, const FeRef &r, bool border, const Quad &Q, const Region &reg, double scale,
                  ffcuda_vec *table, int64_t offset)
{
    typename FeFunctions<FESpaceT>::Fun &F = cache.funs[r.fun];
    ffcuda_vec *dofs = cache.on_device(r.fun);
    FFC(ffcuda_fe_table(D.mesh, F.order, F.default_numbering ? nullptr : F.e2n.data(), F.ncomp, r.comp, dofs, r.op, border ? 1 : 0,
                        (int)Q.w.size(), Q.pts.data(), scale, (int)reg.labels.size(), reg.all ? nullptr : reg.labels.data(), table, offset, 1));
}

// values of mesh-point dependent coefficient expressions at every quadrature node of every element, obtained the way
// Element_rhs / Element_Op obtain them (fflib/problem.cpp:7876-7884, :7951-7960, :6380-6407): MeshPointStack set to the
// node, expression evaluated.  out[e][k * nq + q] for expression e; elements outside the region stay 0.
inline R2 ref_point(const Mesh *, const double *p) { return R2(p[0], p[1]); }
inline R3 ref_point(const Mesh3 *, const double *p) { return R3(p[0], p[1], p[2]); }
template <class FESpaceT>
std::vector<std::vector<double>> eval_at_nodes(Stack stack, const FESpaceT &Vh, const std::vector<const C_F0 *> &exprs, const Quad &Q,
                                               const Region &reg, const std::vector<int> *units = nullptr)
{
    typedef typename FESpaceT::Mesh MeshT;
    typedef typename FESpaceT::FElement FElementT;
    const MeshT &Th = Vh.Th;
    const int dim = MeshDim<MeshT>::d, nq = (int)Q.w.size(), nt = units ? (int)units->size() : Th.nt;
    std::vector<std::vector<double>> out(exprs.size(), std::vector<double>((size_t)nt * nq, 0.0));
    std::set<int> labs(reg.labels.begin(), reg.labels.end());
    MeshPoint *mps = MeshPointStack(stack), mp = *mps;
    try {
        for (int ku = 0; ku < nt; ++ku) {
            const int k = units ? (*units)[ku] : ku; // (out is indexed by the position in the list)
            if (!reg.all && !labs.count(Th[k].lab)) continue;
            const FElementT Kv(Vh[k]);
            const typename MeshT::Element &T = Kv.T;
            for (int q = 0; q < nq; ++q) {
                typename MeshT::RdHat Pt(ref_point(&Th, Q.pts.data() + (size_t)q * dim));
                mps->set(T(Pt), Pt, Kv);
                for (size_t e = 0; e < exprs.size(); ++e) {
                    const C_F0 &c = *exprs[e];
                    out[e][(size_t)ku * nq + q] = c.left() == atype<long>() ? (double)GetAny<long>(c.eval(stack)) : GetAny<double>(c.eval(stack));
                }
            }
        }
    } catch (...) {
        *mps = mp;
        throw;
    }
    *mps = mp;
    return out;
}
// the same at the face quadrature nodes of the boundary elements whose label is listed (Element_rhs / Element_Op on a border
// element, fflib/problem.cpp:8540-8551, :8646-8652, :6520-6528): Pt = PBord(ie, q), mesh point set with the label and the
// unit normal of the boundary element.  out[e][ib * nq + q]; Q holds the face rule in the library's coordinates
// (P = A(1-x-y) + Bx + Cy on a face, P = A(1-x) + Bx on an edge, A, B, C the vertices of the face in FreeFEM's order).
inline R3 bord_point(const Tet &T, int ie, const double *p) { return T.PBord(ie, R2(p[0], p[1])); }
inline R2 bord_point(const Triangle &, int ie, const double *p)
{
    const R2 PA(TriangleHat[VerticesOfTriangularEdge[ie][0]]), PB(TriangleHat[VerticesOfTriangularEdge[ie][1]]);
    return PA * (1.0 - p[0]) + PB * p[0];
}
template <class FESpaceT>
std::vector<std::vector<double>> eval_at_bnodes(Stack stack, const FESpaceT &Vh, const std::vector<const C_F0 *> &exprs, const Quad &Q,
                                                const Region &reg, const std::vector<int> *units = nullptr)
{
    typedef typename FESpaceT::Mesh MeshT;
    typedef typename FESpaceT::FElement FElementT;
    const MeshT &Th = Vh.Th;
    const int dim = MeshDim<MeshT>::d, nq = (int)Q.w.size(), nbe = units ? (int)units->size() : nbe_of(Th);
    std::vector<std::vector<double>> out(exprs.size(), std::vector<double>((size_t)nbe * nq, 0.0));
    std::set<int> labs(reg.labels.begin(), reg.labels.end());
    MeshPoint *mps = MeshPointStack(stack), mp = *mps;
    try {
        for (int iu = 0; iu < nbe; ++iu) {
            const int ib = units ? (*units)[iu] : iu;
            const int r = blabel(Th, ib);
            if (!reg.all && !labs.count(r)) continue;
            int ie;
            const int it = belem_of(Th, ib, ie);
            const FElementT K(Vh[it]);
            const typename MeshT::Rd NN = unit_normal(Th, K.T, ie);
            for (int q = 0; q < nq; ++q) {
                const typename MeshT::RdHat Pt(bord_point(K.T, ie, Q.pts.data() + (size_t)q * (dim - 1)));
                mps->set(K.T(Pt), Pt, K, r, NN, ie);
                for (size_t e = 0; e < exprs.size(); ++e) {
                    const C_F0 &c = *exprs[e];
                    out[e][(size_t)iu * nq + q] = c.left() == atype<long>() ? (double)GetAny<long>(c.eval(stack)) : GetAny<double>(c.eval(stack));
                }
            }
        }
    } catch (...) {
        *mps = mp;
        throw;
    }
    *mps = mp;
    return out;
}
// A coefficient that is an AFFINE combination of FE data with mesh-independent factors — uold/dt, -f (what `- int3d(Th)(f*v)`
// of a problem turns into: (-1)*f), f1 + 2*f2, 2 + kappa + rho/dt (LinearComb::add merges the coefficients of equal terms into
// one sum, femlib/DOperator.hpp:127-131) — is taken apart without reading FreeFEM's private operator nodes:
//  1. E_F0::Optimize (the public pass FieldOfForm itself runs on every coefficient, fflib/AFunction2.cpp:839,
//     AFunction.hpp:2607-2614) flattens the expression into a program: leaves, then operator nodes that read their operands
//     from stack slots.  Every entry must be (i) an FE node we can evaluate on the device = a candidate L_i, (ii) a leaf that
//     does not depend on the mesh point (constants, script variables), or (iii) an operator node (E_F_F0_Opt, E_F_F0F0_Opt,
//     the nested Opt classes of the operator templates, AFunction.hpp:981,1053,2526,2630,2715).  Anything else — x, y, N.x,
//     region, a function on another mesh — and the term stays on the interpreter path: c is then more than a function of
//     the FE data.
//  2. Otherwise c = F(L_1..L_m) with F the program itself, which we can RUN on a scratch stack with any values in the slots
//     of the candidates: beta and alpha_i are read off F by central differences about the middle of the box the L_i range in
//     (dof extrema and the values at the sampled nodes, widened), and F is accepted as affine only if
//     F(p) = beta + sum alpha_i p_i to 1e-13 at the corners, the face centres and 96 random points of that box — so f*g, f^2,
//     sin(f), max(f,c), f > c ? a : b are refused wherever they bend inside the range of the data.
//  3. End-to-end check: the decomposition must reproduce the values the interpreter itself gives for c at the quadrature
//     nodes of a sample of the units (~g_sample_n evenly spread, one per label, every unit of a boundary integral).
// FFCUDA_CHECK=1 compares the whole statement with FreeFEM's own result; FFCUDA_NO_FE_DOFS=1 switches all of this off.
struct FeAffine {
    double beta = 0.0;
    std::vector<std::pair<FeRef, double>> parts;
};
struct FeProgram {
    deque<pair<Expression, int>> ll; // (sub-expression, stack offset) in evaluation order
    size_t top;
    int root;                        // offset of the value of the whole expression (0: the expression could not be flattened)
};
std::map<const E_F0 *, FeProgram> g_fe_programs;
inline bool is_operator_node(const E_F0 *e)
{
    const char *nm = typeid(*e).name(); // Itanium ABI: nested class ...::Opt ends in "3OptE"; the two free templates by prefix
    const size_t len = strlen(nm);
    if (len >= 5 && strcmp(nm + len - 5, "3OptE") == 0) return true;
    return strncmp(nm, "10E_F_F0_Opt", 12) == 0 || strncmp(nm, "12E_F_F0F0_Opt", 14) == 0;
}
template <class FESpaceT>
bool fe_affine(Stack stack, const C_F0 &c, const FESpaceT &Vh, FeFunctions<FESpaceT> &cache, const Quad &Q, const Region &reg, bool border,
               FeAffine &out)
{
    typedef typename FESpaceT::Mesh MeshT;
    const MeshT &Th = Vh.Th;
    out = FeAffine();
    FeRef r0;
    if (fe_reference(stack, c, Vh, cache, r0)) {
        out.parts.push_back(std::make_pair(r0, 1.0));
        return true;
    }
    if (!g_fe_dofs || c.left() != atype<double>()) return false;
    // --- 1. the program (flattened once per expression: Optimize allocates its operator nodes, and a statement inside a time
    //        loop comes back thousands of times)
    std::map<const E_F0 *, FeProgram>::iterator pit = g_fe_programs.find(c.LeftValue());
    if (pit == g_fe_programs.end()) {
        FeProgram P;
        E_F0::MapOfE_F0 mm;
        P.top = 64; // (offset 0 means "not found" in E_F0::find)
        try {
            P.root = c.LeftValue()->Optimize(P.ll, mm, P.top);
        } catch (...) {
            P.root = 0;
        }
        pit = g_fe_programs.insert(std::make_pair((const E_F0 *)c.LeftValue(), P)).first;
    }
    const deque<pair<Expression, int>> &ll = pit->second.ll;
    const size_t top = pit->second.top;
    const int root = pit->second.root;
    if (root <= 0) return false;
    enum Kind { CAND, LEAF, OPER };
    std::vector<Kind> kind(ll.size());
    std::vector<int> cand_of(ll.size(), -1);
    std::vector<C_F0> leaves;
    std::vector<FeRef> refs;
    for (size_t i = 0; i < ll.size(); ++i) {
        C_F0 e(Type_Expr(atype<double>(), ll[i].first));
        FeRef rr;
        if (dynamic_cast<const E_F0_Func1 *>(ll[i].first) && fe_reference(stack, e, Vh, cache, rr)) {
            kind[i] = CAND;
            size_t j = 0;
            while (j < refs.size() && !(refs[j] == rr)) ++j;
            if (j == refs.size()) {
                leaves.push_back(e);
                refs.push_back(rr);
            }
            cand_of[i] = (int)j;
        } else if (is_operator_node(ll[i].first)) kind[i] = OPER;
        else if (ll[i].first->MeshIndependent()) kind[i] = LEAF;
        else return false;
    }
    const size_t m = refs.size();
    if (m == 0 || m > 6 || root <= 0 || (size_t)root + sizeof(AnyType) > top) return false;
    std::vector<char> scratch_mem(top + 2 * sizeof(AnyType) + 64, 0);
    char *scratch = scratch_mem.data();
    scratch += (64 - (reinterpret_cast<uintptr_t>(scratch) & 63)) & 63;
    for (size_t i = 0; i < ll.size(); ++i)
        if (kind[i] == LEAF) *reinterpret_cast<AnyType *>(scratch + ll[i].second) = (*ll[i].first)(stack); // on the REAL stack
    auto F = [&](const double *p) -> double {
        for (size_t i = 0; i < ll.size(); ++i) {
            if (kind[i] == CAND) *reinterpret_cast<AnyType *>(scratch + ll[i].second) = SetAny<double>(p[cand_of[i]]);
            else if (kind[i] == OPER) *reinterpret_cast<AnyType *>(scratch + ll[i].second) = (*ll[i].first)((Stack)scratch);
        }
        return GetAny<double>(*reinterpret_cast<AnyType *>(scratch + root));
    };
    // --- the sample of units (used for the ranges of the data and for the end-to-end check)
    const int nunits = border ? nbe_of(Th) : Th.nt;
    std::set<int> labs(reg.labels.begin(), reg.labels.end()), met;
    std::vector<int> inside, sample;
    for (int k = 0; k < nunits; ++k) {
        const int lab = border ? blabel(Th, k) : elabel(Th, k);
        if (!reg.all && !labs.count(lab)) continue;
        inside.push_back(k);
        if (met.insert(lab).second) sample.push_back(k);
    }
    if (inside.empty()) return false;
    const size_t step = (border && inside.size() <= 262144) ? 1 : std::max<size_t>(1, inside.size() / (size_t)std::max(1, g_sample_n));
    for (size_t i = 0; i < inside.size(); i += step) sample.push_back(inside[i]);
    std::sort(sample.begin(), sample.end());
    sample.erase(std::unique(sample.begin(), sample.end()), sample.end());
    std::vector<const C_F0 *> ex(1, &c);
    for (size_t j = 0; j < m; ++j) ex.push_back(&leaves[j]);
    const std::vector<std::vector<double>> v = border ? eval_at_bnodes(stack, Vh, ex, Q, reg, &sample) : eval_at_nodes(stack, Vh, ex, Q, reg, &sample);
    const size_t rows = v[0].size();
    // --- 2. the box the data range in, and F on it
    std::vector<double> lo(m), hi(m), mid(m), h(m);
    for (size_t j = 0; j < m; ++j) {
        lo[j] = hi[j] = rows ? v[j + 1][0] : 0.0;
        for (size_t k = 0; k < rows; ++k) {
            lo[j] = std::min(lo[j], v[j + 1][k]);
            hi[j] = std::max(hi[j], v[j + 1][k]);
        }
        if (refs[j].op == FFCUDA_OP_ID) { // values: the dofs bound a P0 / P1 function, nearly a P2 one
            const typename FeFunctions<FESpaceT>::Fun &Fn = cache.funs[refs[j].fun];
            const KN<double> &x = *Fn.x;
            for (long q = refs[j].comp; q < x.N(); q += Fn.ncomp) {
                lo[j] = std::min(lo[j], x[q]);
                hi[j] = std::max(hi[j], x[q]);
            }
        }
        const double w = hi[j] - lo[j], pad = refs[j].op == FFCUDA_OP_ID ? 0.25 * w : w; // (derivatives: seen on the sample only)
        lo[j] -= pad;
        hi[j] += pad;
        mid[j] = 0.5 * (lo[j] + hi[j]);
        h[j] = 0.5 * (hi[j] - lo[j]);
        if (!(h[j] > 0.0)) h[j] = std::max(1.0, std::abs(mid[j])); // a constant function: any step will do
        if (!(std::abs(mid[j]) < 1e300) || !(h[j] < 1e300)) return false;
    }
    std::vector<double> p(mid), alpha(m);
    double beta, fmax;
    try {
        const double f0 = F(p.data());
        fmax = std::abs(f0);
        for (size_t j = 0; j < m; ++j) {
            p[j] = mid[j] + h[j];
            const double fp = F(p.data());
            p[j] = mid[j] - h[j];
            const double fm = F(p.data());
            p[j] = mid[j];
            alpha[j] = (fp - fm) / (2.0 * h[j]);
            fmax = std::max(fmax, std::max(std::abs(fp), std::abs(fm)));
        }
        beta = f0;
        for (size_t j = 0; j < m; ++j) beta -= alpha[j] * mid[j];
        double mag = std::abs(beta);
        for (size_t j = 0; j < m; ++j) mag += std::abs(alpha[j]) * (std::abs(mid[j]) + h[j]);
        if (!(mag < 1e300)) return false;
        auto affine_at = [&](const double *q) {
            double fit = beta;
            for (size_t j = 0; j < m; ++j) fit += alpha[j] * q[j];
            const double f = F(q);
            return std::abs(f - fit) <= 1e-13 * std::max(mag, std::abs(f));
        };
        for (unsigned corner = 0; corner < (1u << m); ++corner) { // corners of the box
            for (size_t j = 0; j < m; ++j) p[j] = (corner >> j & 1) ? hi[j] : lo[j];
            if (!affine_at(p.data())) return false;
        }
        uint64_t rng = 0x9E3779B97F4A7C15ull; // deterministic points of the box (xorshift)
        for (int t = 0; t < 96; ++t) {
            for (size_t j = 0; j < m; ++j) {
                rng ^= rng << 13;
                rng ^= rng >> 7;
                rng ^= rng << 17;
                p[j] = lo[j] + (hi[j] - lo[j]) * ((rng >> 11) * (1.0 / 9007199254740992.0));
            }
            if (!affine_at(p.data())) return false;
        }
        // --- 3. end to end: the interpreter's own values of c at the sampled nodes
        for (size_t k = 0; k < rows; ++k) {
            double fit = beta;
            for (size_t j = 0; j < m; ++j) fit += alpha[j] * v[j + 1][k];
            if (std::abs(fit - v[0][k]) > 1e-13 * std::max(mag, std::abs(v[0][k]))) return false;
        }
        if (std::abs(beta) <= 1e-14 * mag) beta = 0.0;
    } catch (...) { // an operator that refuses a value of the box (sqrt of a negative number, ...): not affine for us
        return false;
    }
    out.beta = beta;
    for (size_t j = 0; j < m; ++j)
        if (alpha[j] != 0.0) out.parts.push_back(std::make_pair(refs[j], alpha[j]));
    return true;
}

// FFCUDA_EXPLAIN=1: how every term with mesh-dependent data of a statement will be treated (printed before any device call)
template <class FESpaceT>
void explain_varf(Stack stack, const FESpaceT &Vh, const Varf &V)
{
    FeFunctions<FESpaceT> cache;
    auto say = [&](const char *kind, size_t i, size_t t, const C_F0 &c, const Quad &Q, const Region &reg, bool border) {
        FeAffine A;
        cout << "  -- ffcuda explain: " << kind << " item " << i << " term " << t << ": ";
        if (fe_affine(stack, c, Vh, cache, Q, reg, border, A)) {
            cout << "FE data on the device: " << A.beta;
            for (size_t j = 0; j < A.parts.size(); ++j) {
                const FeRef &r = A.parts[j].first;
                cout << " + " << A.parts[j].second << " * [function #" << r.fun << " (P" << cache.funs[r.fun].order << ", "
                     << cache.funs[r.fun].ncomp << " comp., " << cache.funs[r.fun].x->N() << " dofs"
                     << (cache.funs[r.fun].default_numbering ? "" : ", own node table") << ") comp. " << r.comp << " op " << r.op << "]";
            }
            cout << endl;
        } else cout << "evaluated by the interpreter at the quadrature nodes" << endl;
    };
    for (size_t i = 0; i < V.bil.size(); ++i)
        for (size_t t = 0; t < V.bil[i].qterms.size(); ++t)
            say(V.bil[i].border ? "boundary bilinear" : "bilinear", i, t, V.bil[i].qterms[t].coef, V.bil[i].q, V.bil[i].reg, V.bil[i].border);
    for (size_t i = 0; i < V.lin.size(); ++i)
        for (size_t t = 0; t < V.lin[i].qterms.size(); ++t)
            say(V.lin[i].border ? "boundary linear" : "linear", i, t, V.lin[i].qterms[t].coef, V.lin[i].q, V.lin[i].reg, V.lin[i].border);
}

// table of a linear item: fq[(c * nt + k) * nq + q], summed over the terms of component c; with derivatives of the test
// function (grad = true): fq[((c * (dim+1) + slot) * nt + k) * nq + q]
template <class FESpaceT>
std::vector<double> eval_qvalues(Stack stack, const FESpaceT &Vh, const LinearItem &L, bool negate, bool grad = false)
{
    const size_t per = (size_t)(L.border ? nbe_of(Vh.Th) : Vh.Th.nt) * L.q.w.size();
    const int ns = grad ? MeshDim<typename FESpaceT::Mesh>::d + 1 : 1;
    std::vector<const C_F0 *> ex;
    for (size_t t = 0; t < L.qterms.size(); ++t) ex.push_back(&L.qterms[t].coef);
    std::vector<std::vector<double>> v = L.border ? eval_at_bnodes(stack, Vh, ex, L.q, L.reg) : eval_at_nodes(stack, Vh, ex, L.q, L.reg);
    std::vector<double> fq((size_t)Vh.N * ns * per, 0.0);
    const double sgn = negate ? -1.0 : 1.0;
    for (size_t t = 0; t < L.qterms.size(); ++t) {
        double *dst = fq.data() + ((size_t)L.qterms[t].vcomp * ns + (grad ? L.qterms[t].slot : 0)) * per;
        for (size_t i = 0; i < per; ++i) dst[i] += sgn * v[t][i];
    }
    return fq;
}

// the GPU path proper for a matrix: symbolic, numeric assembly of every bilinear item, Dirichlet conditions, then the
// CSR arrays come back and become a MatriceMorse of FreeFEM's own (HashMatrix::set copies and rebuilds the hash); the
// device copy is returned in res for the solver that will be attached
template <class FESpaceT>
MatriceMorse<double> *gpu_matrix(Stack stack, const FESpaceT &Vh, DevSpace &D, const Varf &V, const Data_Sparse_Solver &ds, Resident &res,
                                 int &n, int64_t &nnz)
{
    // terms whose coefficient depends on the mesh point: values at the quadrature nodes first (this is what may be refused).
    // Terms whose tables are proportional share one coefficient function ((1+x)*lambda, (1+x)*mu, ...): one device pass per
    // group.  The groups are found on a sample of the elements, then one chunked pass over the mesh evaluates every
    // expression, keeps ONE table per group and checks every other term against its group (host memory: groups, not
    // terms); if the sample misled (a coefficient that vanishes or is proportional on the sample only) the exact
    // grouping on full tables is taken instead.
    struct QGroup {
        size_t item;
        std::vector<double> cq;
        std::vector<ffcuda_bterm> terms;
    };
    std::vector<QGroup> groups;
    struct Place {
        int group;     // index into the item's own groups, -1: vanishes
        double alpha;  // multiple of the group's table
        double amax;
    };
    auto place_terms = [](const std::vector<std::vector<double>> &v, std::vector<Place> &pl, std::vector<int> &reps) {
        pl.assign(v.size(), Place{-1, 0.0, 0.0});
        reps.clear();
        for (size_t t = 0; t < v.size(); ++t) {
            const std::vector<double> &a = v[t];
            size_t imax = 0;
            for (size_t k = 1; k < a.size(); ++k)
                if (std::abs(a[k]) > std::abs(a[imax])) imax = k;
            if (a.empty() || a[imax] == 0.0) continue; // the coefficient vanishes at every node
            pl[t].amax = std::abs(a[imax]);
            for (size_t gi = 0; gi < reps.size() && pl[t].group < 0; ++gi) {
                const std::vector<double> &r = v[reps[gi]];
                if (r[imax] == 0.0) continue;
                const double alpha = a[imax] / r[imax];
                double err = 0.0;
                for (size_t k = 0; k < a.size(); ++k) err = std::max(err, std::abs(a[k] - alpha * r[k]));
                if (err <= 1e-14 * std::abs(a[imax])) { // (the products are rounded separately: a few ulp)
                    pl[t].group = (int)gi;
                    pl[t].alpha = alpha;
                }
            }
            if (pl[t].group < 0) {
                pl[t].group = (int)reps.size();
                pl[t].alpha = 1.0;
                reps.push_back((int)t);
            }
        }
    };
    // coefficients that are FE functions on this mesh (kappa, a P0 / P1 / P2 function): one device table per function, formed
    // from its dof array; the terms it multiplies keep the coefficient 1
    struct FeGroup {
        size_t item;
        FeRef ref;
        std::vector<ffcuda_bterm> terms;
    };
    std::vector<FeGroup> fegroups;
    FeFunctions<FESpaceT> fefuns;
    std::vector<std::vector<ffcuda_bterm>> cterms(V.bil.size()); // constant terms of every item (+ constant parts of FE data)
    for (size_t i = 0; i < V.bil.size(); ++i) cterms[i] = V.bil[i].terms;
    for (size_t i = 0; i < V.bil.size(); ++i) {
        const BilinearItem &B = V.bil[i];
        if (B.qterms.empty()) continue;
        std::vector<const C_F0 *> ex;
        std::vector<const QBTerm *> rest; // the terms left to the interpreter
        for (size_t t = 0; t < B.qterms.size(); ++t) {
            FeAffine af;
            if (fe_affine(stack, B.qterms[t].coef, Vh, fefuns, B.q, B.reg, B.border, af)) {
                for (size_t j = 0; j < af.parts.size(); ++j) {
                    const FeRef &r = af.parts[j].first;
                    size_t gi = 0;
                    while (gi < fegroups.size() && !(fegroups[gi].item == i && fegroups[gi].ref == r)) ++gi;
                    if (gi == fegroups.size()) fegroups.push_back(FeGroup{i, r, std::vector<ffcuda_bterm>()});
                    ffcuda_bterm bt = B.qterms[t].t;
                    bt.coef = af.parts[j].second;
                    fegroups[gi].terms.push_back(bt);
                }
                if (af.beta != 0.0) {
                    ffcuda_bterm bt = B.qterms[t].t;
                    bt.coef = af.beta;
                    cterms[i].push_back(bt);
                }
                continue;
            }
            rest.push_back(&B.qterms[t]);
            ex.push_back(&B.qterms[t].coef);
        }
        if (rest.empty()) continue;
        const int nunits = B.border ? nbe_of(Vh.Th) : Vh.Th.nt, nq = (int)B.q.w.size();
        auto eval = [&](const std::vector<int> *units) {
            return B.border ? eval_at_bnodes(stack, Vh, ex, B.q, B.reg, units) : eval_at_nodes(stack, Vh, ex, B.q, B.reg, units);
        };
        std::vector<Place> pl;
        std::vector<int> reps;
        std::vector<std::vector<double>> tables; // one per group
        bool ok = false;
        if (nunits > g_sample_min && ex.size() > 1) {
            std::vector<int> sample;
            const int step = std::max(1, nunits / std::max(1, g_sample_n));
            for (int k = 0; k < nunits; k += step) sample.push_back(k);
            place_terms(eval(&sample), pl, reps);
            tables.assign(reps.size(), std::vector<double>((size_t)nunits * nq, 0.0));
            ok = true;
            const int CH = 8192;
            std::vector<int> chunk;
            for (int k0 = 0; k0 < nunits && ok; k0 += CH) {
                chunk.clear();
                for (int k = k0; k < std::min(nunits, k0 + CH); ++k) chunk.push_back(k);
                const std::vector<std::vector<double>> vc = eval(&chunk);
                const size_t len = chunk.size() * (size_t)nq, off = (size_t)k0 * nq;
                for (size_t gi = 0; gi < reps.size(); ++gi) std::copy(vc[reps[gi]].begin(), vc[reps[gi]].begin() + len, tables[gi].begin() + off);
                for (size_t t = 0; t < ex.size() && ok; ++t) {
                    if (pl[t].group >= 0 && reps[pl[t].group] == (int)t) continue;
                    const double *r = pl[t].group >= 0 ? vc[reps[pl[t].group]].data() : nullptr;
                    for (size_t k = 0; k < len; ++k) {
                        const double want = r ? pl[t].alpha * r[k] : 0.0;
                        if (std::abs(vc[t][k] - want) > 1e-13 * std::max(pl[t].amax, std::abs(want))) {
                            ok = false; // the sample misled: exact grouping below
                            break;
                        }
                    }
                }
            }
        }
        if (!ok) {
            std::vector<std::vector<double>> v = eval(nullptr);
            place_terms(v, pl, reps);
            tables.clear();
            for (size_t gi = 0; gi < reps.size(); ++gi) tables.push_back(std::move(v[reps[gi]]));
        }
        const size_t first_group = groups.size();
        for (size_t gi = 0; gi < reps.size(); ++gi) groups.push_back(QGroup{i, std::move(tables[gi]), std::vector<ffcuda_bterm>()});
        for (size_t t = 0; t < ex.size(); ++t) {
            if (pl[t].group < 0) continue;
            ffcuda_bterm bt = rest[t]->t;
            bt.coef = pl[t].alpha;
            groups[first_group + pl[t].group].terms.push_back(bt);
        }
    }

    ffcuda_pattern *P = nullptr;
    ffcuda_matrix *dA = nullptr;
    FFC(ffcuda_symbolic(D.space, &P));
    if (ffcuda_matrix_create(P, &dA) != 0) {
        ffcuda_pattern_destroy(P);
        fail("ffcuda_matrix_create");
    }
    int rc = ffcuda_pattern_info(P, &n, &nnz);
    // volume integrals first, then the boundary ones (Robin terms) on top; the sum does not depend on the order the
    // varf lists them in beyond round-off
    bool first = true;
    for (int border = 0; border < 2; ++border)
        for (size_t i = 0; i < V.bil.size() && !rc; ++i) {
            const BilinearItem &B = V.bil[i];
            if ((int)B.border != border) continue;
            if (cterms[i].empty() && !B.qterms.empty()) continue; // nothing constant in this item
            rc = (border ? ffcuda_assemble_bilinear_boundary : ffcuda_assemble_bilinear)(
                dA, D.space, (int)cterms[i].size(), cterms[i].data(), (int)B.q.w.size(), B.q.pts.data(), B.q.w.data(),
                (int)B.reg.labels.size(), B.reg.all ? nullptr : B.reg.labels.data(), first ? 0 : 1);
            first = false;
        }
    for (int border = 0; border < 2; ++border)
        for (size_t gi = 0; gi < groups.size() && !rc; ++gi) {
            const BilinearItem &B = V.bil[groups[gi].item];
            if ((int)B.border != border) continue;
            if (border)
                rc = ffcuda_assemble_bilinear_boundary_qcoef(dA, D.space, (int)groups[gi].terms.size(), groups[gi].terms.data(),
                                                             (int)B.q.w.size(), B.q.pts.data(), B.q.w.data(), groups[gi].cq.data(),
                                                             (int)B.reg.labels.size(), B.reg.all ? nullptr : B.reg.labels.data(), first ? 0 : 1);
            else
                rc = ffcuda_assemble_bilinear_qcoef(dA, D.space, (int)groups[gi].terms.size(), groups[gi].terms.data(), (int)B.q.w.size(),
                                                    B.q.pts.data(), B.q.w.data(), groups[gi].cq.data(), first ? 0 : 1);
            first = false;
        }
    for (int border = 0; border < 2; ++border)
        for (size_t gi = 0; gi < fegroups.size() && !rc; ++gi) {
            const BilinearItem &B = V.bil[fegroups[gi].item];
            if ((int)B.border != border) continue;
            const size_t per = (size_t)(border ? nbe_of(Vh.Th) : Vh.Th.nt) * B.q.w.size();
            ffcuda_vec *tab = nullptr;
            try {
                if (per > (size_t)INT_MAX) throw Unsupported{"table of FE data larger than 2^31 entries"};
                FFC(ffcuda_vec_create(context(), (int)per, &tab));
                fe_table_add(D, fefuns, fegroups[gi].ref, border != 0, B.q, B.reg, 1.0, tab, 0);
            } catch (...) {
                if (tab) ffcuda_vec_destroy(tab);
                ffcuda_matrix_destroy(dA);
                ffcuda_pattern_destroy(P);
                throw;
            }
            const double *cq = (const double *)ffcuda_vec_ptr(tab);
            if (border)
                rc = ffcuda_assemble_bilinear_boundary_qcoef(dA, D.space, (int)fegroups[gi].terms.size(), fegroups[gi].terms.data(),
                                                             (int)B.q.w.size(), B.q.pts.data(), B.q.w.data(), cq, (int)B.reg.labels.size(),
                                                             B.reg.all ? nullptr : B.reg.labels.data(), first ? 0 : 1);
            else
                rc = ffcuda_assemble_bilinear_qcoef(dA, D.space, (int)fegroups[gi].terms.size(), fegroups[gi].terms.data(), (int)B.q.w.size(),
                                                    B.q.pts.data(), B.q.w.data(), cq, first ? 0 : 1);
            ffcuda_vec_destroy(tab);
            first = false;
        }
    if (g_verbose && !groups.empty())
        cout << "  -- ffcuda: " << groups.size() << " coefficient function(s) depending on the mesh point, evaluated at the quadrature nodes" << endl;
    if (g_verbose && !fegroups.empty())
        cout << "  -- ffcuda: " << fegroups.size() << " coefficient(s) that are FE functions (" << fefuns.funs.size()
             << " dof array(s) sent), evaluated at the quadrature nodes on the device" << endl;
    // --- hand-off: FreeFEM's own MatriceMorse, its arrays filled straight from the device.  HashMatrix::set
    // (femlib/HashMatrix.cpp:698-728) would copy the three arrays once more and leave the matrix marked `unsorted`, so
    // that the first A.CSR (solver upload, `ofstream << A`, UMFPACK...) heap-sorts nnz entries on one core
    // (Sortij, :671-682).  The device CSR is already sorted by (i, j): the arrays i, j, aij, p are downloaded in place,
    // the hash is rebuilt once (ReHash, :631-642: A(i,j) look-ups work) and the matrix is marked sorted_ij / type_CSR.
    MatriceMorse<double> *M = nullptr;
    if (!rc) {
        try {
            apply_bcs(D, V, dA, nullptr, ds.tgv);
        } catch (...) {
            ffcuda_matrix_destroy(dA);
            ffcuda_pattern_destroy(P);
            throw;
        }
        if (ds.sym) rc = ffcuda_pattern_lower_nnz(P, &nnz); // half storage: FreeFEM keeps the entries (i, j <= i); the device matrix stays full
        if (!rc) {
            Marks hmk;
            M = new MatriceMorse<double>(n, n, 0, 0);
            HashMatrix<int, double> *H = M;
            H->clear();
            H->half = ds.sym ? 1 : 0;
            // the arrays of HashMatrix::Increaze (femlib/HashMatrix.cpp:606-628: i, j, aij, next of nnz entries, head of
            // max(n,m) * min(max(1, nnz/max(n,m)), 50) hash heads), allocated here so that their pages can be asked for as
            // huge pages and touched by several threads before the driver copies into them (first touch of 1.1 GB by one
            // thread, head[] filled by ReHash on one thread: 0.34 s at cube(128) when Increaze did it)
            {
                const size_t mnx = (size_t)n, nzzx = std::max<size_t>((size_t)nnz, 1);
                const double nnzl = std::min(std::max(1., double(nzzx) / double(mnx)), 50.);
                const size_t nh = (size_t)(mnx * nnzl);
                delete[] H->i; delete[] H->j; delete[] H->aij; delete[] H->next; delete[] H->head;
                H->i = new int[nzzx];
                H->j = new int[nzzx];
                H->aij = new double[nzzx];
                H->next = new size_t[nzzx];
                H->head = new size_t[nh];
                H->nnzmax = nzzx;
                H->nhash = nh;
                H->nnz = (size_t)nnz;
                H->setp(n + 1);
                auto touch = [&](void *ptr, size_t bytes) {
                    char *c = static_cast<char *>(ptr);
#ifdef MADV_HUGEPAGE
                    const uintptr_t a0 = ((uintptr_t)c + 4095) & ~(uintptr_t)4095, a1 = ((uintptr_t)c + bytes) & ~(uintptr_t)4095;
                    if (a1 > a0) madvise((void *)a0, a1 - a0, MADV_HUGEPAGE);
#endif
                    par_for((bytes + 4095) / 4096, [&](size_t b, size_t e) {
                        for (size_t pg = b; pg < e; ++pg) c[pg * 4096] = 0;
                    });
                };
                touch(H->i, nzzx * sizeof(int));
                touch(H->j, nzzx * sizeof(int));
                touch(H->aij, nzzx * sizeof(double));
                touch(H->next, nzzx * sizeof(size_t));
                touch(H->head, nh * sizeof(size_t));
                hmk.mark("arrays allocated and touched");
            }
            if (ds.sym) rc = ffcuda_pattern_download_lower(P, H->p, H->j) || ffcuda_matrix_download_lower(dA, H->aij);
            else rc = ffcuda_pattern_download(P, H->p, H->j) || ffcuda_matrix_download(dA, H->aij);
            if (!rc) {
                hmk.mark("downloaded");
                par_for((size_t)n, [&](size_t b, size_t e) {
                    for (size_t i = b; i < e; ++i)
                        for (int k = H->p[i]; k < H->p[i + 1]; ++k) H->i[k] = (int)i;
                });
                // the hash chains of ReHash (femlib/HashMatrix.cpp:631-642: next[k] = head[h]; head[h] = k), built by several
                // threads with an atomic exchange on the heads: the same chains up to the order inside a chain, which only
                // the look-up walks
                {
                    size_t *head = H->head, *next = H->next;
                    const size_t nhash = H->nhash;
                    par_for(nhash, [&](size_t b, size_t e) {
                        for (size_t h = b; h < e; ++h) head[h] = HashMatrix<int, double>::empty;
                    });
                    const int *pi = H->i, *pj = H->j;
                    const size_t nn = (size_t)H->n;
                    par_for((size_t)nnz, [&](size_t b, size_t e) {
                        for (size_t k = b; k < e; ++k) {
                            const size_t h = ((size_t)pi[k] + (size_t)pj[k] * nn) % nhash; // HashMatrix::hash, fortran = 0
                            next[k] = __atomic_exchange_n(&head[h], k, __ATOMIC_RELAXED);
                        }
                    });
                }
                H->state = HashMatrix<int, double>::sorted_ij;
                H->type_state = HashMatrix<int, double>::type_CSR;
                hmk.mark("rows expanded, hash chains built");
                if (g_verbose) cout << "  -- ffcuda: hand-over of the matrix:" << hmk.line << endl;
            }
        }
    }
    if (rc) {
        delete M;
        ffcuda_matrix_destroy(dA);
        ffcuda_pattern_destroy(P);
        fail("assembling the matrix");
    }
    res = Resident{dA, P};
    return M;
}

// the GPU path proper for a right-hand side: every linear item (volume integrals, then boundary integrals), an optional
// change of sign (problem/solve: a(u,v) - l(v) = 0), Dirichlet values; x0 (optional, size n) gets x0[d] = g(d) as
// AssembleBC does for the initial guess
template <class FESpaceT>
void gpu_rhs(Stack stack, const FESpaceT &Vh, DevSpace &D, const Varf &V, double tgv, bool negate, long n, std::vector<double> &host,
             double *x0)
{
    ffcuda_vec *db = nullptr;
    FFC(ffcuda_vec_create(context(), (int)n, &db));
    int rc = 0, nfe = 0;
    bool first = true;
    FeFunctions<FESpaceT> fefuns;
    for (int border = 0; border < 2; ++border)
        for (size_t i = 0; i < V.lin.size() && !rc; ++i) {
            const LinearItem &L = V.lin[i];
            if ((int)L.border != border) continue;
            std::vector<ffcuda_lterm> terms(L.terms);
            if (negate)
                for (size_t k = 0; k < terms.size(); ++k) terms[k].coef = -terms[k].coef;
            // data that are FE functions on this mesh: their dof arrays go to the device, the table is formed there
            std::vector<FeAffine> aff(L.qterms.size());
            bool all_fe = !L.qterms.empty() && !rc;
            try {
                for (size_t t = 0; t < L.qterms.size() && all_fe; ++t)
                    all_fe = fe_affine(stack, L.qterms[t].coef, Vh, fefuns, L.q, L.reg, border != 0, aff[t]);
            } catch (...) {
                ffcuda_vec_destroy(db);
                throw;
            }
            if (all_fe) { // the constant parts (2 + kappa) join the constant terms, assembled first
                for (size_t t = 0; t < L.qterms.size(); ++t)
                    if (aff[t].beta != 0.0) {
                        static const int slot_op[4] = {FFCUDA_OP_ID, FFCUDA_OP_DX, FFCUDA_OP_DY, FFCUDA_OP_DZ};
                        ffcuda_lterm lt;
                        lt.vcomp = L.qterms[t].vcomp;
                        lt.vop = slot_op[L.qterms[t].slot];
                        lt.coef = negate ? -aff[t].beta : aff[t].beta;
                        terms.push_back(lt);
                    }
            }
            if (!terms.empty() || L.qterms.empty()) {
                rc = (border ? ffcuda_assemble_linear_boundary : ffcuda_assemble_linear)(
                    db, D.space, (int)terms.size(), terms.data(), (int)L.q.w.size(), L.q.pts.data(), L.q.w.data(),
                    (int)L.reg.labels.size(), L.reg.all ? nullptr : L.reg.labels.data(), first ? 0 : 1);
                first = false;
            }
            if (all_fe && !rc) {
                bool grad = false;
                for (size_t t = 0; t < L.qterms.size(); ++t) grad = grad || L.qterms[t].slot != 0;
                const int ns = (grad && !border) ? MeshDim<typename FESpaceT::Mesh>::d + 1 : 1;
                const size_t per = (size_t)(border ? nbe_of(Vh.Th) : Vh.Th.nt) * L.q.w.size();
                ffcuda_vec *tab = nullptr;
                try {
                    if ((size_t)Vh.N * ns * per > (size_t)INT_MAX) throw Unsupported{"table of FE data larger than 2^31 entries"};
                    FFC(ffcuda_vec_create(context(), (int)((size_t)Vh.N * ns * per), &tab));
                    for (size_t t = 0; t < L.qterms.size(); ++t)
                        for (size_t j = 0; j < aff[t].parts.size(); ++j)
                            fe_table_add(D, fefuns, aff[t].parts[j].first, border != 0, L.q, L.reg,
                                         negate ? -aff[t].parts[j].second : aff[t].parts[j].second, tab,
                                         (int64_t)(((size_t)L.qterms[t].vcomp * ns + (ns > 1 ? L.qterms[t].slot : 0)) * per));
                } catch (...) {
                    if (tab) ffcuda_vec_destroy(tab);
                    ffcuda_vec_destroy(db);
                    throw;
                }
                rc = (border ? ffcuda_assemble_linear_boundary_qvalues : grad ? ffcuda_assemble_linear_qterms : ffcuda_assemble_linear_qvalues)(
                    db, D.space, (int)L.q.w.size(), L.q.pts.data(), L.q.w.data(), (const double *)ffcuda_vec_ptr(tab), first ? 0 : 1);
                ffcuda_vec_destroy(tab);
                first = false;
                nfe += (int)L.qterms.size();
            } else if (!L.qterms.empty() && !rc) {
                std::vector<double> fq;
                bool grad = false;
                for (size_t t = 0; t < L.qterms.size(); ++t) grad = grad || L.qterms[t].slot != 0;
                try {
                    fq = eval_qvalues(stack, Vh, L, negate, grad);
                } catch (...) {
                    ffcuda_vec_destroy(db);
                    throw;
                }
                rc = (border ? ffcuda_assemble_linear_boundary_qvalues : grad ? ffcuda_assemble_linear_qterms : ffcuda_assemble_linear_qvalues)(
                    db, D.space, (int)L.q.w.size(), L.q.pts.data(), L.q.w.data(), fq.data(), first ? 0 : 1);
                first = false;
            }
        }
    if (first && !rc) rc = ffcuda_vec_fill(db, 0.0);
    if (g_verbose && nfe)
        cout << "  -- ffcuda: " << nfe << " term(s) whose data are FE functions (" << fefuns.funs.size()
             << " dof array(s) sent), evaluated at the quadrature nodes on the device" << endl;
    host.resize((size_t)n);
    ffcuda_vec *dx = nullptr;
    if (!rc) {
        try {
            apply_bcs(D, V, nullptr, db, tgv);
            if (x0 && !V.bc.empty()) {
                FFC(ffcuda_vec_create(context(), (int)n, &dx));
                if (ffcuda_vec_upload(dx, x0) != 0) fail("uploading the initial guess");
                for (size_t i = 0; i < V.bc.size(); ++i) {
                    const BCItem &B = V.bc[i];
                    ffcuda_bc *bc = make_bc(D.space, B);
                    int r2 = ffcuda_vec_set_bc_values(dx, bc);
                    ffcuda_bc_destroy(bc);
                    if (r2) fail("setting the Dirichlet values of the initial guess");
                }
                if (ffcuda_vec_download(dx, x0) != 0) fail("downloading the initial guess");
            }
        } catch (...) {
            ffcuda_vec_destroy(db);
            if (dx) ffcuda_vec_destroy(dx);
            throw;
        }
        rc = ffcuda_vec_download(db, host.data());
    }
    ffcuda_vec_destroy(db);
    if (dx) ffcuda_vec_destroy(dx);
    if (rc) fail("assembling the right-hand side");
}

// ------------------------------------------------------------------------------------------------------------
// 0. matrix B = vb(Uh,Vh) with two DIFFERENT spaces on one mesh (FFCUDA_RECT=1): the blocks of a Stokes / mixed problem
//    assembled one by one, projection matrices between P1 and P2.  The built-in operator takes the two spaces
//    (fflib/problem.hpp:1628-1631, creationBlockOfMatrixToBilinearForm :1664-1700: a MatriceMorse with Vh.NbOfDF rows and
//    Uh.NbOfDF columns, "lines corresponding to test functions").  Claimed here: volume integrals with constant coefficients,
//    one quadrature rule and one set of regions for the whole form, no on(...), no sym=1.  Everything else goes to FreeFEM.
//    Off by default: the library entry behind it (ffcuda_assemble_bilinear_rect) is pinned against the reference on the host
//    only (tests/test_rect_row_host.py) until a GPU run is on record.
// ------------------------------------------------------------------------------------------------------------
struct RectForm {
    std::vector<ffcuda_bterm> terms;
    Quad q;
    Region reg;
    bool have = false;
};
template <class MeshT>
RectForm read_rect_form(Stack stack, const list<C_F0> &largs, const MeshT &Th, int ncomp_u, int ncomp_v, bool *has_bc = nullptr)
{
    const int dim = MeshDim<MeshT>::d;
    RectForm F;
    if (has_bc) *has_bc = false;
    for (list<C_F0>::const_iterator ii = largs.begin(); ii != largs.end(); ++ii) {
        Expression e = ii->LeftValue();
        aType r = ii->left();
        if (r == atype<const FormLinear *>()) continue; // ignored when a matrix is assembled
        if (r == atype<const BC_set *>()) {
            if (!has_bc) throw Unsupported{"on(...) in a form with two different spaces"};
            *has_bc = true; // (one space: the caller runs FreeFEM's own AssembleBC on the assembled matrix)
            continue;
        }
        if (r != atype<const FormBilinear *>()) throw Unsupported{"varf item other than integrals"};
        const FormBilinear *bf = dynamic_cast<const FormBilinear *>(e);
        if (bf->VF()) throw Unsupported{"discontinuous-Galerkin operators"};
        if (check_domain(stack, *bf->di, Th)) throw Unsupported{"boundary integral in a form with two different spaces"};
        Quad q = volume_rule(stack, *bf->di, &Th);
        Region reg = region_of(stack, *bf->di, 16);
        if (!F.have) {
            F.q = q;
            F.reg = reg;
            F.have = true;
        } else if (q.pts != F.q.pts || q.w != F.q.w || reg.all != F.reg.all || reg.labels != F.reg.labels)
            throw Unsupported{"integrals with different quadrature rules or regions in a form with two different spaces"};
        const Foperator &op = *bf->b;
        for (size_t k = 0; k < op.v.size(); ++k) {
            const pair<MGauche, MDroit> &id = op.v[k].first; // (unknown, test)
            ffcuda_bterm t;
            t.ucomp = id.first.first;
            t.uop = check_op(id.first.second, dim);
            t.vcomp = id.second.first;
            t.vop = check_op(id.second.second, dim);
            if (t.ucomp < 0 || t.ucomp >= ncomp_u || t.vcomp < 0 || t.vcomp >= ncomp_v) throw Unsupported{"component out of range"};
            t.coef = constant_coef(stack, op.v[k].second); // (mesh-dependent coefficients: left to FreeFEM here)
            F.terms.push_back(t);
        }
    }
    if (!F.have || F.terms.empty()) throw Unsupported{"no bilinear term"};
    // the device pattern holds the couples of EVERY element; FreeFEM's only those of the elements the integral visits
    // (HashMatrix creates them as it goes): the same thing only when the regions cover the mesh (cf. check_full_pattern)
    if (!F.reg.all) {
        const std::set<int> labs(F.reg.labels.begin(), F.reg.labels.end());
        for (int k = 0; k < Th.nt; ++k)
            if (!labs.count(Th[k].lab)) throw Unsupported{"the volume integrals do not visit every element (sub-pattern)"};
    }
    if (F.terms.size() > 64) throw Unsupported{"more than 64 terms"};
    if (F.q.w.size() > 32) throw Unsupported{"more than 32 quadrature points"};
    return F;
}

// the device part: B as host CSR triple (rows = dofs of Vh, columns = dofs of Uh)
template <class FESpaceT>
void gpu_rect_matrix(const FESpaceT &Uh, const FESpaceT &Vh, const RectForm &F, std::vector<int> &I, std::vector<int> &J,
                     std::vector<double> &C)
{
    typedef typename FESpaceT::FElement FElementT;
    typedef typename FESpaceT::Mesh MeshT;
    const int dim = MeshDim<MeshT>::d;
    int order_u, ncomp_u, nloc_u;
    classify_space(Uh, dim, order_u, ncomp_u, nloc_u);
    DevSpace &DV = device_space(Vh); // the test space and its copy of the mesh; the space of the unknown is put on the same mesh
    const int nt = Uh.NbOfElements;
    std::vector<int32_t> e2n((size_t)nt * nloc_u);
    for (int k = 0; k < nt; ++k) {
        const FElementT K(Uh[k]);
        for (int a = 0; a < nloc_u; ++a) e2n[(size_t)k * nloc_u + a] = K(a) / ncomp_u;
    }
    ffcuda_space *su = nullptr;
    FFC(ffcuda_space_create(DV.mesh, order_u, ncomp_u, e2n.data(), Uh.NbOfNodes, &su));
    ffcuda_matrix *dB = nullptr;
    int rc = ffcuda_assemble_bilinear_rect(DV.space, su, (int)F.terms.size(), F.terms.data(), (int)F.q.w.size(), F.q.pts.data(), F.q.w.data(),
                                           (int)F.reg.labels.size(), F.reg.all ? nullptr : F.reg.labels.data(), &dB);
    int n = 0, m = 0;
    int64_t nnz = 0;
    if (!rc) rc = ffcuda_matrix_shape(dB, &n, &m, &nnz);
    if (!rc && (n != Vh.NbOfDF || m != Uh.NbOfDF)) rc = -1;
    if (!rc) {
        I.resize((size_t)nnz);
        J.resize((size_t)nnz);
        C.resize((size_t)nnz);
        rc = ffcuda_matrix_download_coo(dB, I.data(), J.data(), C.data(), 0);
    }
    if (dB) ffcuda_matrix_destroy(dB);
    ffcuda_space_destroy(su);
    if (rc) fail("assembling the rectangular matrix");
}

// ------------------------------------------------------------------------------------------------------------
// 0b. matrix A = va(Xh,Xh) on a MIXED-ORDER product space in one fespace (FFCUDA_RECT=1): Taylor-Hood `[P2,P2,P1]`,
//     `[P2,P2,P2,P1]`.  The local dofs of such an element are component-major (begin_dfcomp / end_dfcomp,
//     femlib/FESpacen.hpp:454-455, FESpace.hpp:516-517: what Element_Op loops over, fflib/problem.cpp:6414-6417), the global
//     numbering is whatever FreeFEM made it.  Every couple of components (cv, cu) is one scalar block: the rectangular
//     entry is called with two scalar spaces whose "nodes" are the GLOBAL dofs of the component, so each block comes back
//     in the numbering of the whole space and the matrix is their union (the blocks are disjoint; a couple of components
//     without a term still contributes its structural zeros, as HashMatrix::operator+= keeps them).  on(...) items: FreeFEM's
//     own AssembleBC runs on the assembled matrix (O(boundary) work).  Same status and guard as section 0.
// ------------------------------------------------------------------------------------------------------------
struct MixedSpace {
    int ncomp = 0, nd = 0;
    int order[4], begin[5];
};
// is Vh a product of P1 / P2 Lagrange components that are NOT all of the same order ?  (throws nothing; false = not ours)
template <class FESpaceT>
bool classify_mixed(const FESpaceT &Vh, MixedSpace &X)
{
    typedef typename FESpaceT::FElement FElementT;
    typedef typename FESpaceT::Mesh MeshT;
    const int dim = MeshDim<MeshT>::d, nv = dim + 1, n2 = dim == 3 ? 10 : 6;
    X.ncomp = Vh.N;
    if (X.ncomp < 2 || X.ncomp > 4 || Vh.NbOfElements <= 0) return false;
    const FElementT K0(Vh[0]);
    X.nd = K0.NbDoF();
    bool differ = false;
    int at = 0;
    for (int c = 0; c < X.ncomp; ++c) {
        if (K0.dfcbegin(c) != at) return false; // component-major, no gaps
        const int nl = K0.dfcend(c) - K0.dfcbegin(c);
        X.begin[c] = at;
        if (nl == nv) X.order[c] = 1;
        else if (nl == n2) X.order[c] = 2;
        else return false;
        at += nl;
        differ = differ || X.order[c] != X.order[0];
    }
    X.begin[X.ncomp] = at;
    if (at != X.nd || !differ) return false;
    // basis fingerprint: component c of local dof begin[c] + a is the Lagrange function a of that order, the others vanish
    double l[4] = {0.1, 0.2, 0.3, 0.4};
    if (dim == 2) l[0] = 0.2, l[1] = 0.3, l[2] = 0.5;
    KNMK<double> val(X.nd, X.ncomp, (int)last_operatortype);
    val = 0.;
    basis_values(K0, hat_point((const MeshT *)0, l), val);
    for (int c = 0; c < X.ncomp; ++c) {
        double phi[10];
        lagrange_values(dim, X.order[c], l, phi);
        for (int a = 0; a < X.begin[c + 1] - X.begin[c]; ++a)
            for (int c2 = 0; c2 < X.ncomp; ++c2)
                if (fabs(val(X.begin[c] + a, c2, (int)op_id) - (c == c2 ? phi[a] : 0.0)) > 1e-12) return false;
    }
    const int nt = Vh.NbOfElements, step = std::max(1, nt / 64);
    for (int k = 0; k < nt; k += step)
        if (FElementT(Vh[k]).NbDoF() != X.nd) return false;
    return true;
}

// the mesh of a space on the device, for this statement only (no fespace of ours is attached to it: see device_space for
// the cached path of the square operator)
template <class MeshT>
ffcuda_mesh *upload_mesh_plain(const MeshT &Th)
{
    const int dim = MeshDim<MeshT>::d, nv = Th.nv, nt = Th.nt, nvk = dim + 1;
    std::vector<double> xyz((size_t)nv * dim);
    for (int i = 0; i < nv; ++i) coords(Th, i, &xyz[(size_t)i * dim]);
    std::vector<int32_t> conn((size_t)nt * nvk), elab(nt);
    for (int k = 0; k < nt; ++k) {
        for (int j = 0; j < nvk; ++j) conn[(size_t)k * nvk + j] = Th(k, j);
        elab[k] = elabel(Th, k);
    }
    ffcuda_mesh *m = nullptr;
    FFC(ffcuda_mesh_upload(context(), dim, nv, xyz.data(), nt, conn.data(), elab.data(), 0, nullptr, nullptr, nullptr, nullptr, &m));
    return m;
}

// the device part: the union of the ncomp^2 scalar blocks as a COO triple in the numbering of Vh
template <class FESpaceT>
void gpu_mixed_matrix(const FESpaceT &Vh, const MixedSpace &X, const RectForm &F, std::vector<int> &I, std::vector<int> &J,
                      std::vector<double> &C)
{
    typedef typename FESpaceT::FElement FElementT;
    const int nt = Vh.NbOfElements, nc = X.ncomp, ndof = Vh.NbOfDF;
    ffcuda_mesh *mesh = upload_mesh_plain(Vh.Th);
    std::vector<ffcuda_space *> sp((size_t)nc, nullptr);
    ffcuda_matrix *dB = nullptr;
    int rc = 0;
    for (int c = 0; c < nc && !rc; ++c) { // scalar space of component c: its nodes are the global dofs of the component
        const int nl = X.begin[c + 1] - X.begin[c];
        std::vector<int32_t> tab((size_t)nt * nl);
        for (int k = 0; k < nt; ++k) {
            const FElementT K(Vh[k]);
            for (int a = 0; a < nl; ++a) tab[(size_t)k * nl + a] = K(X.begin[c] + a);
        }
        rc = ffcuda_space_create(mesh, X.order[c], 1, tab.data(), ndof, &sp[c]);
    }
    for (int cv = 0; cv < nc && !rc; ++cv)
        for (int cu = 0; cu < nc && !rc; ++cu) {
            std::vector<ffcuda_bterm> bt;
            for (size_t t = 0; t < F.terms.size(); ++t)
                if (F.terms[t].vcomp == cv && F.terms[t].ucomp == cu) {
                    ffcuda_bterm b = F.terms[t];
                    b.vcomp = b.ucomp = 0;
                    bt.push_back(b);
                }
            if (bt.empty()) { // no term couples these components: their couples are in the matrix all the same, with zeros
                ffcuda_bterm z;
                z.ucomp = z.vcomp = 0;
                z.uop = z.vop = FFCUDA_OP_ID;
                z.coef = 0.0;
                bt.push_back(z);
            }
            rc = ffcuda_assemble_bilinear_rect(sp[cv], sp[cu], (int)bt.size(), bt.data(), (int)F.q.w.size(), F.q.pts.data(), F.q.w.data(),
                                               (int)F.reg.labels.size(), F.reg.all ? nullptr : F.reg.labels.data(), &dB);
            int n = 0, m = 0;
            int64_t nnz = 0;
            if (!rc) rc = ffcuda_matrix_shape(dB, &n, &m, &nnz);
            if (!rc && (n != ndof || m != ndof)) rc = -1;
            if (!rc) {
                const size_t at = I.size();
                I.resize(at + (size_t)nnz);
                J.resize(at + (size_t)nnz);
                C.resize(at + (size_t)nnz);
                rc = ffcuda_matrix_download_coo(dB, I.data() + at, J.data() + at, C.data() + at, 0);
            }
            if (dB) ffcuda_matrix_destroy(dB);
            dB = nullptr;
        }
    for (int c = 0; c < nc; ++c)
        if (sp[c]) ffcuda_space_destroy(sp[c]);
    ffcuda_mesh_destroy(mesh);
    if (rc) fail("assembling the blocks of the mixed-order matrix");
}

// ------------------------------------------------------------------------------------------------------------
// 1. matrix A = va(Vh,Vh,...)
// ------------------------------------------------------------------------------------------------------------
template <class MMesh, class v_fes>
struct CudaMatrixOp : public OpMatrixtoBilinearForm<double, MMesh, v_fes, v_fes> {
    typedef OpMatrixtoBilinearForm<double, MMesh, v_fes, v_fes> Base;
    struct Op : public Base::Op {
        Op(Expression aa, Expression bb, int initt) : Base::Op(aa, bb, initt) {}
        // matrix B = vb(Uh,Vh), two different spaces (section 0 above); throws Unsupported for what it does not claim
        AnyType rectangular(Stack stack, const typename v_fes::FESpace &Uh, const typename v_fes::FESpace &Vh) const
        {
            if ((const void *)&Uh.Th != (const void *)&Vh.Th) throw Unsupported{"the two spaces live on different meshes"};
            Data_Sparse_Solver ds;
            ds.factorize = 0;
            ds.initmat = true;
            int np = OpCall_FormBilinear_np::n_name_param - NB_NAME_PARM_HMAT;
            SetEnd_Data_Sparse_Solver<double>(stack, ds, this->b->nargs, np);
            if (ds.sym) throw Unsupported{"sym=1 with two different spaces"};
            const MMesh &Th = Vh.Th;
            if (!isSameMesh(this->b->largs, &Uh.Th, &Vh.Th, stack)) throw Unsupported{"integrals on different meshes"};
            if (Uh.N < 1 || Uh.N > 3 || Vh.N < 1 || Vh.N > 3) throw Unsupported{"more than 3 components"};
            const RectForm F = read_rect_form(stack, this->b->largs, Th, Uh.N, Vh.N);
            if (g_explain)
                cout << "  -- ffcuda explain: rectangular matrix " << Vh.NbOfDF << " x " << Uh.NbOfDF << ": " << F.terms.size() << " term(s), "
                     << F.q.w.size() << " quadrature point(s), " << (F.reg.all ? std::string("all regions") : std::to_string(F.reg.labels.size()) + " region label(s)")
                     << endl;
            std::vector<int> I, J;
            std::vector<double> C;
            gpu_rect_matrix(Uh, Vh, F, I, J, C);
            const int n = Vh.NbOfDF, m = Uh.NbOfDF;
            std::unique_ptr<MatriceMorse<double>> guard(new MatriceMorse<double>(n, m, 0, 0));
            guard->set(n, m, 0, I.size(), I.data(), J.data(), C.data(), 0, 0); // COO triple sorted by (i, j): HashMatrix::set copies it
            if (g_check) { // FFCUDA_CHECK=1: FreeFEM's own operator runs as well, the two matrices are compared, FreeFEM's is kept
                guard.reset();
                AnyType r = Base::Op::operator()(stack);
                Matrice_Creuse<double> &Af(*GetAny<Matrice_Creuse<double> *>((*this->a)(stack)));
                HashMatrix<int, double> *H = Af.pHM();
                if (!H) ExecError("ffcuda check: FreeFEM's operator did not produce a sparse matrix");
                H->CSR();
                bool same = H->n == n && H->m == m && (size_t)H->nnz == J.size();
                for (int r = 0; same && r < n; ++r)
                    for (int k = H->p[r]; same && k < H->p[r + 1]; ++k) same = I[k] == r && H->j[k] == J[k];
                if (!same) ExecError("ffcuda check: the sparsity pattern of the rectangular matrix differs from FreeFEM's");
                double amax = 0, dmax = 0;
                for (size_t k = 0; k < C.size(); ++k) {
                    amax = std::max(amax, std::abs(H->aij[k]));
                    dmax = std::max(dmax, std::abs(H->aij[k] - C[k]));
                }
                cout << "  -- ffcuda check: rectangular matrix " << n << " x " << m << ", nnz " << J.size()
                     << ": pattern identical, max |dB| / max |B| = " << (amax > 0 ? dmax / amax : dmax) << endl;
                if (dmax > 1e-12 * amax) ExecError("ffcuda check: values of the rectangular matrix differ from FreeFEM's by more than 1e-12");
                return r;
            }
            // --- hand-over as the built-in operator does it (problem.hpp:1640-1660)
            WhereStackOfPtr2Free(stack) = new StackOfPtr2Free(stack);
            Matrice_Creuse<double> &A(*GetAny<Matrice_Creuse<double> *>((*this->a)(stack)));
            if (this->init) A.init();
            A.A = 0;
            A.Uh = Uh;
            A.Vh = Vh;
            A.A.master(guard.release());
            A.pHM()->half = 0;
            if (Uh.NbOfDF == Vh.NbOfDF) SetSolver(stack, false, *A.A, ds); // (square by its sizes: the built-in operator sets the solver then)
            if (g_verbose) cout << "  -- ffcuda: rectangular matrix " << n << " x " << m << ", nnz " << J.size() << " assembled on the GPU" << endl;
            return SetAny<Matrice_Creuse<double> *>(&A);
        }
        // matrix A = va(Xh,Xh) on a mixed-order product space (section 0b above); throws Unsupported for what it does not claim
        AnyType mixed(Stack stack, const typename v_fes::FESpace &Vh, const MixedSpace &X) const
        {
            typedef typename v_fes::FESpace FESpaceT;
            Data_Sparse_Solver ds;
            ds.factorize = 0;
            ds.initmat = true;
            int np = OpCall_FormBilinear_np::n_name_param - NB_NAME_PARM_HMAT;
            SetEnd_Data_Sparse_Solver<double>(stack, ds, this->b->nargs, np);
            if (ds.sym) throw Unsupported{"sym=1 on a mixed-order space"};
            if (ds.tgv != ds.tgv) throw Unsupported{"tgv is NaN"};
            const MMesh &Th = Vh.Th;
            if (!isSameMesh(this->b->largs, &Vh.Th, &Vh.Th, stack)) throw Unsupported{"integrals on different meshes"};
            bool has_bc = false;
            const RectForm F = read_rect_form(stack, this->b->largs, Th, X.ncomp, X.ncomp, &has_bc);
            if (g_explain) {
                cout << "  -- ffcuda explain: mixed-order space [";
                for (int c = 0; c < X.ncomp; ++c) cout << (c ? ",P" : "P") << X.order[c];
                cout << "], " << Vh.NbOfDF << " dofs: " << X.ncomp * X.ncomp << " scalar blocks, " << F.terms.size() << " term(s), " << F.q.w.size()
                     << " quadrature point(s)" << (has_bc ? ", on(...) by FreeFEM's AssembleBC" : "") << endl;
            }
            std::vector<int> I, J;
            std::vector<double> C;
            gpu_mixed_matrix(Vh, X, F, I, J, C);
            const int n = Vh.NbOfDF;
            std::unique_ptr<MatriceMorse<double>> guard(new MatriceMorse<double>(n, n, 0, 0)); // (AssembleBC may raise an ExecError)
            MatriceMorse<double> *M = guard.get();
            M->set(n, n, 0, I.size(), I.data(), J.data(), C.data(), 0, 0); // COO, block after block: HashMatrix::set copies it
            if (has_bc) AssembleBC<double, MMesh, FESpaceT, FESpaceT>(stack, Th, Vh, Vh, false, M, 0, 0, this->b->largs, ds.tgv);
            if (g_check) { // FFCUDA_CHECK=1: FreeFEM's own operator runs as well, the two matrices are compared, FreeFEM's is kept
                M->CSR();
                const std::vector<int> p0(M->p, M->p + n + 1), j0(M->j, M->j + M->nnz);
                const std::vector<double> a0(M->aij, M->aij + M->nnz);
                guard.reset();
                AnyType r = Base::Op::operator()(stack);
                Matrice_Creuse<double> &Af(*GetAny<Matrice_Creuse<double> *>((*this->a)(stack)));
                HashMatrix<int, double> *H = Af.pHM();
                if (!H) ExecError("ffcuda check: FreeFEM's operator did not produce a sparse matrix");
                H->CSR();
                bool same = H->n == n && H->m == n && (size_t)H->nnz == j0.size();
                for (int i = 0; same && i <= n; ++i) same = H->p[i] == p0[i];
                for (size_t k = 0; same && k < j0.size(); ++k) same = H->j[k] == j0[k];
                if (!same) ExecError("ffcuda check: the sparsity pattern of the mixed-order matrix differs from FreeFEM's");
                double amax = 0, dmax = 0;
                for (size_t k = 0; k < a0.size(); ++k) {
                    const double a = H->aij[k];
                    if (std::abs(a) > 1e29 || std::abs(a0[k]) > 1e29) {
                        if (a != a0[k]) dmax = 1e300;
                        continue;
                    }
                    amax = std::max(amax, std::abs(a));
                    dmax = std::max(dmax, std::abs(a - a0[k]));
                }
                cout << "  -- ffcuda check: mixed-order matrix " << n << " x " << n << ", nnz " << j0.size()
                     << ": pattern identical, max |dA| / max |A| = " << (amax > 0 ? dmax / amax : dmax) << endl;
                if (dmax > 1e-12 * amax) ExecError("ffcuda check: values of the mixed-order matrix differ from FreeFEM's by more than 1e-12");
                return r;
            }
            WhereStackOfPtr2Free(stack) = new StackOfPtr2Free(stack);
            Matrice_Creuse<double> &A(*GetAny<Matrice_Creuse<double> *>((*this->a)(stack)));
            if (this->init) A.init();
            A.A = 0;
            A.Uh = Vh;
            A.Vh = Vh;
            A.A.master(guard.release());
            A.pHM()->half = 0;
            SetSolver(stack, false, *A.A, ds);
            if (g_verbose) cout << "  -- ffcuda: mixed-order matrix " << n << " x " << n << ", nnz " << I.size() << " assembled on the GPU" << endl;
            return SetAny<Matrice_Creuse<double> *>(&A);
        }
        AnyType operator()(Stack stack) const
        {
            typedef typename v_fes::pfes pfes;
            typedef typename v_fes::FESpace FESpaceT;
            pfes *pUh = GetAny<pfes *>((*this->b->euh)(stack));
            pfes *pVh = GetAny<pfes *>((*this->b->evh)(stack));
            const FESpaceT *PUh = (FESpaceT *)**pUh, *PVh = (FESpaceT *)**pVh;
            try {
                if (!PUh || !PVh) throw Unsupported{"null fespace"};
                if (PUh != PVh) {
                    if (!g_rect) throw Unsupported{"test and unknown spaces differ (FFCUDA_RECT=1 takes such forms to the device)"};
                    return rectangular(stack, *PUh, *PVh);
                }
                if (g_rect) { // one mixed-order product space ([P2,P2,P1] ...): scalar blocks through the rectangular entry
                    MixedSpace X;
                    if (classify_mixed(*PVh, X)) return mixed(stack, *PVh, X);
                }
                Data_Sparse_Solver ds;
                ds.factorize = 0;
                ds.initmat = true;
                int np = OpCall_FormBilinear_np::n_name_param - NB_NAME_PARM_HMAT;
                SetEnd_Data_Sparse_Solver<double>(stack, ds, this->b->nargs, np);
                if (ds.tgv != ds.tgv) throw Unsupported{"tgv is NaN"};
                const FESpaceT &Vh = *PVh;
                const MMesh &Th = Vh.Th;
                if (!isSameMesh(this->b->largs, &Vh.Th, &Vh.Th, stack)) throw Unsupported{"integrals on different meshes"};
                Varf V = read_varf(stack, this->b->largs, Th, Vh.N, true, Vh);
                if (g_explain) explain_varf(stack, Vh, V);
                check_full_pattern(V, Th);
                check_qterms_supported(Vh, V);
                DevSpace &D = device_space(Vh);
                int n = 0;
                int64_t nnz = 0;
                Resident res{nullptr, nullptr};
                MatriceMorse<double> *M = gpu_matrix(stack, Vh, D, V, ds, res, n, nnz);
                if (g_check) { // FFCUDA_CHECK=1: FreeFEM's own operator runs as well, the two matrices are compared, FreeFEM's is kept
                    M->CSR();
                    const std::vector<int> p0(M->p, M->p + n + 1), j0(M->j, M->j + M->nnz);
                    const std::vector<double> a0(M->aij, M->aij + M->nnz);
                    delete M;
                    release_resident(res);
                    AnyType r = Base::Op::operator()(stack);
                    Matrice_Creuse<double> &Af(*GetAny<Matrice_Creuse<double> *>((*this->a)(stack)));
                    HashMatrix<int, double> *H = Af.pHM();
                    if (!H) ExecError("ffcuda check: FreeFEM's operator did not produce a sparse matrix");
                    H->CSR();
                    bool same = H->n == n && (size_t)H->nnz == j0.size();
                    for (int i = 0; same && i <= n; ++i) same = H->p[i] == p0[i];
                    for (size_t k = 0; same && k < j0.size(); ++k) same = H->j[k] == j0[k];
                    if (!same) ExecError("ffcuda check: the sparsity pattern differs from FreeFEM's");
                    double amax = 0, dmax = 0;
                    for (size_t k = 0; k < a0.size(); ++k) {
                        const double a = H->aij[k];
                        if (std::abs(a) > 1e29 || std::abs(a0[k]) > 1e29) {
                            if (a != a0[k]) dmax = 1e300;
                            continue;
                        }
                        amax = std::max(amax, std::abs(a));
                        dmax = std::max(dmax, std::abs(a - a0[k]));
                    }
                    cout << "  -- ffcuda check: matrix " << n << " x " << n << ", nnz " << j0.size() << ": pattern identical, max |dA| / max |A| = "
                         << (amax > 0 ? dmax / amax : dmax) << endl;
                    if (dmax > 1e-12 * amax) ExecError("ffcuda check: matrix values differ from FreeFEM's by more than 1e-12");
                    return r;
                }
                // --- hand the result to FreeFEM as its own MatriceMorse (problem.hpp:1678-1693)
                WhereStackOfPtr2Free(stack) = new StackOfPtr2Free(stack);
                Matrice_Creuse<double> &A(*GetAny<Matrice_Creuse<double> *>((*this->a)(stack)));
                if (this->init) A.init();
                A.A = 0;
                A.Uh = Vh;
                A.Vh = Vh;
                A.A.master(M);
                drop_resident(); // at most one matrix waits for its solver
                g_resident[(const void *)static_cast<HashMatrix<int, double> *>(M)] = res; // stays on the device for the solver
                A.pHM()->half = ds.sym;
                SetSolver(stack, false, *A.A, ds);
                drop_resident(); // not adopted (the solver is not ours): a later set(A,solver=CG) uploads the matrix as it is THEN
                if (g_verbose) cout << "  -- ffcuda: matrix " << n << " x " << n << ", nnz " << nnz << " assembled on the GPU" << endl;
                return SetAny<Matrice_Creuse<double> *>(&A);
            } catch (const Unsupported &u) {
                drop_resident();
                notice("matrix = varf(Vh,Vh)", u.why);
                return Base::Op::operator()(stack);
            }
        }
    };
    E_F0 *code(const basicAC_F0 &args) const { return new Op(to<Matrice_Creuse<double> *>(args[0]), args[1], this->init); }
    CudaMatrixOp(int initt) : Base(initt) { this->pref = 100; }
};
// ------------------------------------------------------------------------------------------------------------
// 2. real[int] b = va(0,Vh)
// ------------------------------------------------------------------------------------------------------------
template <class MMesh, class v_fes>
struct CudaRhsOp : public OpArraytoLinearForm<double, MMesh, v_fes> {
    typedef OpArraytoLinearForm<double, MMesh, v_fes> Base;
    struct Op : public Base::Op {
        Op(Expression xx, Expression ll, bool isptrr, bool initt, bool zzero) : Base::Op(xx, ll, isptrr, initt, zzero) {}
        AnyType operator()(Stack stack) const
        {
            typedef v_fes *pfes;
            typedef typename v_fes::FESpace FESpaceT;
            pfes &pp = *GetAny<pfes *>((*this->l->ppfes)(stack));
            FESpaceT *pVh = *pp;
            try {
                if (!pVh) throw Unsupported{"null fespace"};
                if (!this->zero) throw Unsupported{"b += varf(0,Vh)"};
                FESpaceT &Vh = *pVh;
                double tgv = ff_tgv;
                if (this->l->nargs[0]) tgv = GetAny<double>((*this->l->nargs[0])(stack));
                if (tgv != tgv) throw Unsupported{"tgv is NaN"};
                Varf V = read_varf(stack, this->l->largs, Vh.Th, Vh.N, false, Vh);
                if (g_explain) explain_varf(stack, Vh, V);
                if (V.other_rhs_items) throw Unsupported{"right-hand side with array / matrix-vector items"};
                DevSpace &D = device_space(Vh);
                const long n = Vh.NbOfDF;
                // the array, sized as the built-in operator does (problem.hpp:1363-1383)
                KN<double> *px = 0;
                if (this->isptr) {
                    px = GetAny<KN<double> *>((*this->x)(stack));
                    if (this->init) px->init(n);
                    if (px->N() != n) px->resize(n);
                }
                KN_<double> xx(px ? *(KN_<double> *)px : GetAny<KN_<double>>((*this->x)(stack)));
                if (xx.N() != n) ExecError("ffcuda: array and fespace sizes differ in b = varf(0,Vh)");
                std::vector<double> host;
                gpu_rhs(stack, Vh, D, V, tgv, false, n, host, nullptr);
                if (g_check) { // FFCUDA_CHECK=1: FreeFEM's own operator runs as well, the two vectors are compared, FreeFEM's is kept
                    AnyType r = Base::Op::operator()(stack);
                    KN_<double> xf(px ? *(KN_<double> *)px : GetAny<KN_<double>>((*this->x)(stack)));
                    double bmax = 0, dmax = 0;
                    for (long i = 0; i < n; ++i) {
                        if (std::abs(xf[i]) > 1e20 || std::abs(host[i]) > 1e20) {
                            if (std::abs(xf[i] - host[i]) > 1e-14 * std::abs(xf[i])) dmax = 1e300;
                            continue;
                        }
                        bmax = std::max(bmax, std::abs(xf[i]));
                        dmax = std::max(dmax, std::abs(xf[i] - host[i]));
                    }
                    cout << "  -- ffcuda check: right-hand side of size " << n << ": max |db| / max |b| = " << (bmax > 0 ? dmax / bmax : dmax) << endl;
                    if (dmax > 1e-12 * bmax) ExecError("ffcuda check: right-hand side differs from FreeFEM's by more than 1e-12");
                    return r;
                }
                for (long i = 0; i < n; ++i) xx[i] = host[i]; // KN_ may be strided
                if (g_verbose) cout << "  -- ffcuda: right-hand side of size " << n << " assembled on the GPU" << endl;
                return SetAny<KN_<double>>(xx);
            } catch (const Unsupported &u) {
                notice("array = varf(0,Vh)", u.why);
                return Base::Op::operator()(stack);
            }
        }
    };
    E_F0 *code(const basicAC_F0 &args) const
    {
        if (this->isptr) return new Op(to<KN<double> *>(args[0]), args[1], this->isptr, this->init, this->zero);
        return new Op(to<KN_<double>>(args[0]), args[1], this->isptr, this->init, this->zero);
    }
    CudaRhsOp(const basicForEachType *tt, bool isptrr, bool initt, bool zzero = 1) : Base(tt, isptrr, initt, zzero) { this->pref = 100; }
};

// ------------------------------------------------------------------------------------------------------------
// 3. solver=CG  ->  Jacobi-CG on the device (contract: VirtualSolver<int,double>, femlib/VirtualSolver.hpp:190-243;
//    reference: SolverCG, femlib/VirtualSolverCG.hpp:112-192)
// ------------------------------------------------------------------------------------------------------------
class SolverCudaCG : public VirtualSolver<int, double> {
  public:
    // 1 unsym, 2 herm, 4 sym, 8 pos, 16 nopos, 32 seq : what the reference CG declares
    static const int orTypeSol = 1 | 2 | 4 | 8 | 32;
    typedef HashMatrix<int, double> HMat;
    HMat *A;
    Resident dev;
    int verb, itermax;
    double eps, tgv;
    double *veps;
    long *getnbiter;
    VirtualSolver<int, double> *hostcg; // FreeFEM's own SolverCG when the script gives a preconditioner (precon=)
    std::vector<ffcuda_matrix *> dmat;  // FFCUDA_NGPU > 1: the row block of every GPU
    std::vector<int> dfirst;            // first row of every block (+ n)

    SolverCudaCG(HMat &AA, const Data_Sparse_Solver &ds, Stack stack, bool is_cg = true)
        : A(&AA), dev{nullptr, nullptr}, verb(ds.verb), itermax(ds.itmax > 0 ? ds.itmax : AA.n), eps(ds.epsilon), tgv(ds.tgv),
          veps(ds.veps), getnbiter(ds.getnbiter), hostcg(nullptr)
    {
        if (AA.n != AA.m) ExecError("ffcuda: CG needs a square matrix");
        std::map<const void *, Resident>::iterator it = g_resident.find((const void *)A);
        if (it != g_resident.end()) { // assembled by the statement that attaches this solver: already on the device
            int dn = 0;
            int64_t dnnz = 0;
            // (the entry only lives during that statement; the size test guards against a stale entry all the same)
            if (ffcuda_matrix_info(it->second.A, &dn, &dnnz) == 0 && dn == AA.n && !(is_cg && ds.precon)) {
                dev = it->second;
                A->GetReDoNumerics();
                A->GetReDoSymbolic();
            } else
                release_resident(it->second);
            g_resident.erase(it);
        }
        // solver=CG, precon=P: the preconditioner is a FreeFEM expression evaluated by the interpreter on host vectors
        // (HMatVirtPrecon, femlib/VirtualSolverCG.hpp:27-63, with its tgv-row fix :50-62); FreeFEM's own CG runs it, as
        // its GMRES does for SolverCudaGMRES.  Never silently replaced by Jacobi.
        if (is_cg && ds.precon) {
            notice("solver=CG", "user preconditioner (precon=)");
            hostcg = new SolverCG<int, double>(AA, ds, stack);
        }
    }
    // several GPUs: full storage, enough rows, the gang answers
    bool want_gang() const { return g_ngpu > 1 && !A->half && A->n >= g_ngpu_min_n && A->n >= 64 * g_ngpu && gang_ready(); }
    void release_dist()
    {
        for (size_t r = 0; r < dmat.size(); ++r)
            if (dmat[r]) ffcuda_matrix_destroy(dmat[r]);
        dmat.clear();
        dfirst.clear();
    }
    // the MatriceMorse shared out by contiguous row blocks, one per GPU (values: slices of FreeFEM's own array)
    void upload_dist()
    {
        release_dist();
        release_resident(dev);
        A->CSR();
        const int n = A->n, N = g_ngpu;
        dmat.assign((size_t)N, nullptr);
        dfirst.assign((size_t)N + 1, n);
        Marks mk;
        on_ranks(N, [&](int r) {
            int64_t sz[8];
            rank_check(ffcuda_partition_rows_local(n, A->p, A->j, r, N, sz, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
                       nullptr, "ffcuda_partition_rows_local");
            const int no = (int)sz[0], ng = (int)sz[1], nn = (int)sz[3];
            std::vector<int32_t> l2g((size_t)no + ng), rp((size_t)no + 1), ci((size_t)sz[2]), nbr((size_t)nn), ro((size_t)nn), rc((size_t)nn),
                sp((size_t)nn + 1), si((size_t)sz[4]);
            rank_check(ffcuda_partition_rows_local(n, A->p, A->j, r, N, sz, l2g.data(), rp.data(), ci.data(), nbr.data(), ro.data(), rc.data(),
                                                   sp.data(), si.data()),
                       nullptr, "ffcuda_partition_rows_local");
            dfirst[r] = (int)sz[5];
            rank_check(ffcuda_matrix_from_csr_distributed(g_gang[r], no, no + ng, sz[2], rp.data(), ci.data(), A->aij + A->p[sz[5]], nn,
                                                          nbr.data(), ro.data(), rc.data(), sp.data(), si.data(), &dmat[r]),
                       g_gang[r], "ffcuda_matrix_from_csr_distributed");
        });
        mk.mark("row blocks on the GPUs");
        if (g_verbose) cout << "  -- ffcuda: matrix " << n << " x " << n << " shared out over " << N << " GPUs:" << mk.line << endl;
    }
    void upload()
    {
        release_dist();
        release_resident(dev);
        A->CSR(); // sorted, p[] built (HashMatrix.cpp:859-876)
        if (A->half) { // sym=1: entries (i, j <= i); expanded to the full symmetric matrix on the way in
            if (ffcuda_matrix_from_csr_lower(context(), A->n, (int64_t)A->nnz, A->p, A->j, A->aij, &dev.A) != 0) fail("uploading the matrix");
        } else if (ffcuda_matrix_from_csr(context(), A->n, (int64_t)A->nnz, A->p, A->j, A->aij, &dev.A) != 0)
            fail("uploading the matrix");
    }
    void UpdateState()
    {
        if (hostcg) return; // works on the host matrix
        const bool num = A->GetReDoNumerics(), sym = A->GetReDoSymbolic();
        if (want_gang()) {
            if (dmat.empty() || num || sym) upload_dist();
            return;
        }
        if (!dev.A || num || sym) upload(); // the script changed the matrix after it was assembled
    }
    void dosolver(double *x, double *b, int N, int trans)
    {
        if (hostcg) {
            hostcg->dosolver(x, b, N, trans);
            return;
        }
        (void)trans; // A'^-1 with CG: the matrix is symmetric by the user's contract (the reference multiplies by A^T, the same)
        const bool gang = want_gang();
        if (gang && dmat.empty()) upload_dist();
        if (!gang && !dev.A) upload();
        if (getnbiter) *getnbiter = 0;
        int err = 0;
        for (int k = 0, oo = 0; k < N; ++k, oo += A->n) {
            int iters = 0, conv = 0;
            double gcg = 0;
            if (gang) { // every GPU iterates on its rows; the scalars are the same on all of them
                std::vector<int> its((size_t)g_ngpu, 0), cv((size_t)g_ngpu, 0);
                std::vector<double> gg((size_t)g_ngpu, 0.);
                Marks mk;
                on_ranks(g_ngpu, [&](int r) {
                    rank_check(ffcuda_cg_host(dmat[r], b + oo + dfirst[r], x + oo + dfirst[r], eps, itermax, tgv, &its[r], &cv[r], &gg[r]), g_gang[r],
                               "ffcuda_cg_host");
                });
                mk.mark("CG");
                if (g_verbose) cout << "  -- ffcuda: solve on " << g_ngpu << " GPUs:" << mk.line << endl;
                iters = its[0];
                conv = cv[0];
                gcg = gg[0];
            } else if (ffcuda_cg_host(dev.A, b + oo, x + oo, eps, itermax, tgv, &iters, &conv, &gcg) != 0) fail("ffcuda_cg_host");
            if (verb || g_verbose)
                cout << " GC (ffcuda" << (gang ? ", " + std::to_string(g_ngpu) + " GPUs" : std::string()) << "): "
                     << (conv ? "converge" : "NO convergence") << " after " << iters << " g=" << gcg << endl;
            if (!conv) err++;
            else if (getnbiter) *getnbiter += iters;
            if (veps) { // what FreeFEM's SolverCG hands back: the absolute threshold ConjugueGradient stopped on (CG.cpp:226)
                double eps2 = 0;
                if (eps > 0 && ffcuda_cg_stop_threshold(gang ? dmat[0] : dev.A, &eps2) == 0) *veps = sqrt(eps2);
                else *veps = eps;
            }
        }
        if (err) {
            std::cerr << "Error: ConjugueGradient (ffcuda) do not converge nb end =" << err << std::endl;
            ffassert(0);
        }
    }
    ~SolverCudaCG()
    {
        release_dist();
        release_resident(dev);
        delete hostcg;
    }
};

// ------------------------------------------------------------------------------------------------------------
// 4. solver=GMRES  ->  right-preconditioned (Jacobi) flexible GMRES on the device (reference: SolverGMRES,
//    femlib/VirtualSolverCG.hpp:196-258, fgmres femlib/CG.cpp:347-517).  A user preconditioner (precon=) or a transposed
//    solve (A'^-1) are not on the path: FreeFEM's own SolverGMRES runs them.
// ------------------------------------------------------------------------------------------------------------
class SolverCudaGMRES : public SolverCudaCG {
  public:
    static const int orTypeSol = 1 | 2 | 4 | 8 | 16 | 32; // what the reference GMRES declares
    int restart;
    SolverGMRES<int, double> *host; // FreeFEM's own solver for what is left to it
    bool precon;
    SolverCudaGMRES(HMat &AA, const Data_Sparse_Solver &ds, Stack stack)
        : SolverCudaCG(AA, ds, stack, false), restart(ds.NbSpace), host(nullptr), precon(ds.precon != 0)
    {
        if (precon) host = new SolverGMRES<int, double>(AA, ds, stack); // (Data_Sparse_Solver cannot be kept: built now)
    }
    void dosolver(double *x, double *b, int N, int trans)
    {
        if (trans || precon) {
            notice("GMRES", trans ? "transposed solve" : "user preconditioner");
            if (!host) {
                Data_Sparse_Solver ds; // the parameters this solver was created with
                ds.epsilon = eps;
                ds.itmax = itermax;
                ds.NbSpace = restart;
                ds.tgv = tgv;
                ds.verb = verb;
                host = new SolverGMRES<int, double>(*A, ds, nullptr);
            }
            host->dosolver(x, b, N, trans);
            return;
        }
        const bool gang = want_gang();
        if (gang && dmat.empty()) upload_dist();
        if (!gang && !dev.A) upload();
        if (getnbiter) *getnbiter = 0;
        int err = 0;
        for (int k = 0, oo = 0; k < N; ++k, oo += A->n) {
            int iters = 0, conv = 0;
            double rel = 0;
            if (gang) {
                std::vector<int> its((size_t)g_ngpu, 0), cv((size_t)g_ngpu, 0);
                std::vector<double> rr((size_t)g_ngpu, 0.);
                on_ranks(g_ngpu, [&](int r) {
                    rank_check(ffcuda_gmres_host(dmat[r], b + oo + dfirst[r], x + oo + dfirst[r], eps, itermax, restart, tgv, &its[r], &cv[r], &rr[r]),
                               g_gang[r], "ffcuda_gmres_host");
                });
                iters = its[0];
                conv = cv[0];
                rel = rr[0];
            } else if (ffcuda_gmres_host(dev.A, b + oo, x + oo, eps, itermax, restart, tgv, &iters, &conv, &rel) != 0) fail("ffcuda_gmres_host");
            if (verb || g_verbose)
                cout << "  **  fgmres (ffcuda" << (gang ? ", " + std::to_string(g_ngpu) + " GPUs" : std::string()) << ") "
                     << (conv ? "has converged in " : "has not converged in ") << iters
                     << " iterations The relative residual is " << rel << endl;
            if (!conv) err++;
            if (getnbiter) *getnbiter = iters;
            if (veps) *veps = eps;
        }
        if (err) {
            std::cerr << "Error: fgmres (ffcuda) do not converge nb end =" << err << std::endl;
            ffassert(0);
        }
    }
    ~SolverCudaGMRES() { delete host; }
};

// ------------------------------------------------------------------------------------------------------------
// 5. problem / solve  (Problem::eval, fflib/problem.cpp:12198-12450; types registered at fflib/lgfem.cpp:6308-6309,
//    keywords at :6541-6542).  `solve Poisson(u,v,solver=CG) = int3d(Th)(...) - int3d(Th)(f*v) + on(...)` builds a
//    Problem object at parse time through TypeSolve::SetParam and runs Problem::operator() at execution time.  The
//    plugin makes SetParam build a subclass whose operator() takes the GPU path for what it claims (one real P1/P2
//    space, same unknown and test space, items read_varf accepts) and calls Problem::operator() for everything else.
//    Steps of eval kept here: solver parameters, spaces of the unknowns, Data<FESpace> bookkeeping on the stack
//    (matrix kept across calls when init= says so), X initialised from the previous solution (InitProblem :11826-11902,
//    Nb = 1), A and B assembled (on the device), B = -B, |B_i| < 1e-60 -> 0, Dirichlet rows (AssembleBC), DefSolver,
//    A.Solve(X, B), solution handed to the FE function (DispatchSolution :12086-12101, Nb = 1).
// ------------------------------------------------------------------------------------------------------------
template <class T>
struct FeTypes;
template <>
struct FeTypes<Mesh> {
    typedef FESpace FES;
    typedef v_fes vfes;
};
template <>
struct FeTypes<Mesh3> {
    typedef FESpace3 FES;
    typedef v_fes3 vfes;
};

// dimension of a real problem as dimProblem finds it (fflib/problem.cpp:13118): 2, 3, or 0 for what is not claimed here
int claimed_dim(const ListOfId &l)
{
    typedef pair<FEbase<double, v_fes> *, int> pfer_;
    typedef pair<FEbase<double, v_fes3> *, int> pf3r_;
    int dim = 0;
    bool other = false;
    auto look = [&](const UnId &idi) {
        C_F0 c = ::Find(idi.id);
        if (BCastTo<pfer_>(c)) {
            if (dim == 3) other = true;
            dim = 2;
        } else if (BCastTo<pf3r_>(c)) {
            if (dim == 2) other = true;
            dim = 3;
        } else
            other = true;
    };
    for (size_t i = 0; i < l.size(); ++i)
        if (l[i].e == 0) { // (named parameters carry an expression)
            if (l[i].array) {
                const ListOfId &a = *l[i].array;
                for (size_t j = 0; j < a.size(); ++j)
                    if (a[j].r == 0 && a[j].re == 0 && a[j].array == 0) look(a[j]);
                    else other = true;
            } else
                look(l[i]);
        }
    return other ? 0 : dim;
}

template <class Base>
struct CudaProblem : public Base {
    int cdim;
    CudaProblem(const C_args *ca, const ListOfId &l, size_t &top) : Base(ca, l, top), cdim(claimed_dim(l)) {}

    template <class MeshT>
    AnyType run(Stack stack, Problem::Data<typename FeTypes<MeshT>::FES> *data) const
    {
        typedef typename FeTypes<MeshT>::FES FES;
        typedef typename FeTypes<MeshT>::vfes vfes;
        typedef pair<FEbase<double, vfes> *, int> pfer_;
        if (this->nargs[0]) throw Unsupported{"save= parameter"};
        if (this->nargs[1]) throw Unsupported{"cadna= parameter"};
        const int nvar = (int)this->var.size();
        if (nvar < 2 || nvar % 2) throw Unsupported{"odd number of unknown / test functions"};
        std::vector<pfer_> u_hh((size_t)nvar);
        for (int i = 0; i < nvar; ++i) u_hh[i] = GetAny<pfer_>((*(this->var[i]))(stack));
        // one fespace for the unknowns, one for the test functions: components 0..N-1 of the same FE function each
        const int N = nvar / 2;
        for (int i = 0; i < nvar; ++i) {
            if (u_hh[i].second != i % N) throw Unsupported{"unknowns from several fespaces"};
            if (u_hh[i].first != u_hh[i - i % N].first) throw Unsupported{"unknowns from several FE functions"};
        }
        FEbase<double, vfes> *uh = u_hh[0].first, *vh = u_hh[N].first;
        const FES *Uhp = uh->newVh(), *Vhp = vh->newVh();
        if (!Uhp || !Vhp) throw Unsupported{"null fespace"};
        if (Uhp != Vhp) throw Unsupported{"test and unknown spaces differ"};
        if (Uhp->N != N) throw Unsupported{"number of unknowns differs from the components of the fespace"};
        const MeshT &Th = Uhp->Th;
        if (!isSameMesh(this->op->largs, &Th, &Th, stack)) throw Unsupported{"integrals on different meshes"};
        // everything that may be refused is read before anything is changed
        Varf VA = read_varf(stack, this->op->largs, Th, N, true, *Uhp);
        Varf VB = read_varf(stack, this->op->largs, Th, N, false, *Uhp);
        if (VB.other_rhs_items) throw Unsupported{"array / matrix-vector items in the problem"};
        check_full_pattern(VA, Th);
        check_qterms_supported(*Uhp, VA);
        if (data->pTh == &Th && (const FES *)data->Uh != Uhp) throw Unsupported{"the problem was set up on another fespace of this mesh"};
        DevSpace &D = device_space(*Uhp);

        MeshPoint *mps = MeshPointStack(stack), mp = *mps;
        Data_Sparse_Solver ds;
        const int np = 3 + NB_NAME_PARM_MAT; // Problem::n_name_param - NB_NAME_PARM_HMAT (fflib/problem.hpp:501)
        SetEnd_Data_Sparse_Solver<double>(stack, ds, this->nargs, np);
        if (ds.tgv != ds.tgv) throw Unsupported{"tgv is NaN"};
        WhereStackOfPtr2Free(stack) = new StackOfPtr2Free(stack);
        if (&Th != data->pTh) {
            ds.initmat = true;
            data->pTh = &Th;
            data->Uh = Uhp;
            data->Vh = Vhp;
        }
        const FES &Uh(*data->Uh);
        const long n = Uh.NbOfDF;
        // X: the previous solution when the FE function lives on this mesh (InitProblem, Nb = 1)
        KN<double> *X = new KN<double>(n);
        if (!(const FES *)uh->Vh || &uh->Vh->Th != &Th || !uh->x() || uh->x()->N() != n) *X = 0.;
        else *X = *uh->x();
        KN<double> *B = nullptr;
        try {
            if (ds.initmat) {
                int nn = 0;
                int64_t nnz = 0;
                Resident res{nullptr, nullptr};
                MatriceMorse<double> *M = gpu_matrix(stack, Uh, D, VA, ds, res, nn, nnz);
                data->AR.master(M);
                drop_resident();
                g_resident[(const void *)static_cast<HashMatrix<int, double> *>(M)] = res;
                if (g_verbose) cout << "  -- ffcuda: problem matrix " << nn << " x " << nn << ", nnz " << nnz << " assembled on the GPU" << endl;
            }
            if (!data->AR) throw Unsupported{"init= without a matrix"};
            MatriceCreuse<double> &A(*data->AR);
            std::vector<double> hb;
            gpu_rhs(stack, Uh, D, VB, ds.tgv, true, n, hb, VB.bc.empty() ? nullptr : (double *)*X);
            B = new KN<double>(n);
            for (long i = 0; i < n; ++i) (*B)[i] = std::abs(hb[i]) < 1.e-60 ? 0. : hb[i];
            dynamic_cast<HashMatrix<int, double> *>(&A)->half = ds.sym;
            if (ds.initmat) {
                DefSolver(stack, A, ds);
                drop_resident(); // see CudaMatrixOp: a device copy nobody adopted must not outlive the statement
            }
            if (g_verbose) cout << "  -- ffcuda: problem right-hand side of size " << n << " assembled on the GPU" << endl;
            A.Solve(*X, *B);
        } catch (const Unsupported &) {
            delete X;
            delete B;
            throw;
        } catch (...) {
            if (verbosity) cout << " catch an erreur in  solve  =>  set  sol = 0 !!!!!!! " << endl;
            *X = 0.;
            *uh = X;
            delete B;
            throw;
        }
        *uh = X; // DispatchSolution, Nb = 1: the FE function owns X now
        delete B;
        if (verbosity) cout << "  -- Solve : \n          min " << uh->x()->min() << "  max " << uh->x()->max() << endl;
        *mps = mp;
        return SetAny<const Problem *>(this);
    }

    AnyType operator()(Stack stack) const
    {
        if (this->complextype || this->VF || !cdim) {
            notice("problem / solve", this->complextype ? "complex problem" : this->VF ? "discontinuous-Galerkin operators" : "not a real 2-D / 3-D problem");
            return Problem::operator()(stack);
        }
        try {
            if (cdim == 2) return run<Mesh>(stack, this->dataptr(stack));
            return run<Mesh3>(stack, this->dataptr3(stack));
        } catch (const Unsupported &u) {
            notice("problem / solve", u.why);
            return Problem::operator()(stack);
        }
    }
};

// TypeSolve (fflib/problem.hpp:1040-1094) with SetParam building the subclass above.  The `solve` and `problem` keywords
// hold pointers to the type objects created at start-up (zzzfff->AddF, lgfem.cpp:6541-6542) and mylex refuses a second
// registration of a keyword, so those very objects are re-pointed to this class: it adds no data member and overrides
// one virtual function, the objects keep their addresses and contents.
template <bool exec_init, class P>
struct CudaTypeSolve : public TypeSolve<exec_init, P> {
    Type_Expr SetParam(const C_F0 &c, const ListOfId *l, size_t &top) const
    {
        if (c.left() != atype<const C_args *>()) CompileError(" Problem  a(...) = invalid type ", c.left());
        const C_args *ca = dynamic_cast<const C_args *>(c.LeftValue());
        P *pb = new CudaProblem<P>(ca, *l, top);
        return Type_Expr(this, pb);
    }
};
template <bool exec_init, class P>
void repoint_solve_type()
{
    static_assert(sizeof(CudaTypeSolve<exec_init, P>) == sizeof(TypeSolve<exec_init, P>), "CudaTypeSolve must not add data");
    basicForEachType *t = map_type[typeid(const P *).name()];
    if (!t || !dynamic_cast<TypeSolve<exec_init, P> *>(t)) {
        cerr << " ffcuda: the type of problem / solve is not the one expected; they stay with FreeFEM" << endl;
        return;
    }
    static CudaTypeSolve<exec_init, P> model;
    // The keywords `problem` / `solve` hold pointers to the type objects created at start-up and the lexer refuses a second
    // registration, so the object is given the dynamic type of the subclass (no data member, one overridden virtual
    // function).  This relies on the Itanium C++ ABI (vptr = first word of a polymorphic object with a polymorphic primary
    // base).  Self-test at load time: the first word of the model must be its vptr (two models share it and differ from the
    // base's), and after the switch the object must answer as a CudaTypeSolve through RTTI; otherwise everything is put
    // back and problem / solve stay with FreeFEM.
    static CudaTypeSolve<exec_init, P> model2;
    void *const vp = *reinterpret_cast<void **>(static_cast<basicForEachType *>(&model));
    void *const vp2 = *reinterpret_cast<void **>(static_cast<basicForEachType *>(&model2));
    void *const old = *reinterpret_cast<void **>(t);
    if (vp != vp2 || vp == old || static_cast<void *>(static_cast<basicForEachType *>(&model)) != static_cast<void *>(&model)) {
        cerr << " ffcuda: unexpected object layout; problem / solve stay with FreeFEM" << endl;
        return;
    }
    *reinterpret_cast<void **>(t) = vp;
    if (!dynamic_cast<CudaTypeSolve<exec_init, P> *>(t) || typeid(*t) != typeid(model)) {
        *reinterpret_cast<void **>(t) = old;
        cerr << " ffcuda: the type switch of problem / solve did not take; they stay with FreeFEM" << endl;
    }
}

// ------------------------------------------------------------------------------------------------------------
// cube(nx,ny,nz) and buildlayers(Th2,n,...) with the mesh generated on the device.
// FreeFEM builds the arrays in 0.15 s at cube(64) and then spends 1.8 s in GenericMesh::BuildAdj (one hash-table
// insertion per face, femlib/GenericMesh.hpp:837-1046) — 2 minutes at cube(256).  Here the vertices, tetrahedra, boundary
// triangles (final orientation), the adjacency and the boundary links come from the device (csrc/mesh.cu) and are put
// into a Mesh3 the way build_layer does it (fflib/msh3.cpp:895-934): set(), fill, BuildBound, [adjacency arrays filled
// instead of BuildAdj], Buildbnormalv, BuildjElementConteningVertex, BuildGTree.  The device mesh stays resident for the
// first fespace on it.  Same operator signatures and named parameters as the built-ins, preference 100; what is not
// covered (label= / flags= of cube, transfo= / facemerge= / ptmerge= of buildlayers) runs the built-in expression,
// which is always compiled alongside.
// ------------------------------------------------------------------------------------------------------------
// a CUDA context if there is a device, without raising (the mesh generators fall back to FreeFEM's own code)
bool device_available()
{
    if (g_ctx) return true;
    const char *d = getenv("FFCUDA_DEVICE");
    ffcuda_ctx *c = nullptr;
    if (ffcuda_ctx_create(d ? atoi(d) : 0, &c) != 0) return false;
    g_ctx = c;
    return true;
}

template <class LabelOfVertex>
Mesh3 *mesh3_from_device(ffcuda_mesh *dm, LabelOfVertex vlab, Marks &mk)
{
    int dim = 0, nv = 0, nt = 0, nbe = 0;
    FFC(ffcuda_mesh_info(dm, &dim, &nv, &nt, &nbe));
    if (dim != 3) fail("internal: the generated mesh is not 3-D");
    std::vector<double> xyz((size_t)nv * 3);
    std::vector<int32_t> conn((size_t)nt * 4), elab(nt), bconn((size_t)nbe * 3), blab(nbe), belem(nbe), bface(nbe);
    FFC(ffcuda_mesh_download(dm, xyz.data(), conn.data(), elab.data(), bconn.data(), blab.data(), belem.data(), bface.data()));
    Mesh3 *Th = new Mesh3;
    Th->set(nv, nt, nbe);
    int *adj = new int[(size_t)4 * nt], *head = new int[std::max(nbe, 1)];
    static_assert(sizeof(int) == sizeof(int32_t), "int is expected to be 32 bits");
    FFC(ffcuda_mesh_adjacency(dm, reinterpret_cast<int32_t *>(adj), nullptr));
    mk.mark("generated + downloaded");
    par_for((size_t)nv, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            Vertex3 &V = Th->vertices[i];
            V.x = xyz[3 * i];
            V.y = xyz[3 * i + 1];
            V.z = xyz[3 * i + 2];
            V.lab = vlab((int)i);
        }
    });
    par_for((size_t)nt, [&](size_t b, size_t e) {
        for (size_t k = b; k < e; ++k) {
            int iv[4] = {conn[4 * k], conn[4 * k + 1], conn[4 * k + 2], conn[4 * k + 3]};
            Th->elements[k].set(Th->vertices, iv, elab[k]);
        }
    });
    for (int ib = 0; ib < nbe; ++ib) {
        int iv[3] = {bconn[3 * (size_t)ib], bconn[3 * (size_t)ib + 1], bconn[3 * (size_t)ib + 2]};
        Th->borderelements[ib].set(Th->vertices, iv, blab[ib]);
        head[ib] = 4 * belem[ib] + bface[ib];
    }
    Th->BuildBound();
    // what BuildAdj leaves behind (femlib/GenericMesh.hpp:837-1046): k*4+i -> k'*4+i', -1 on the boundary; the boundary
    // triangles already have their final orientation
    Th->TheAdjacencesLink = adj;
    Th->BoundaryElementHeadLink = head;
    Th->nadjnomanifold = 0;
    Th->Buildbnormalv();
    Th->BuildjElementConteningVertex();
    Th->BuildGTree();
    mk.mark("Mesh3");
    return Th;
}

struct CudaCubeOp : public E_F0mps {
    Expression builtin, enx, eny, enz;
    static const int n_name_param = 3;
    static basicAC_F0::name_and_type name_param[];
    Expression nargs[n_name_param];
    CudaCubeOp(const basicAC_F0 &args, Expression b, Expression nx, Expression ny, Expression nz) : builtin(b), enx(nx), eny(ny), enz(nz)
    {
        args.SetNameParam(n_name_param, name_param, nargs);
    }
    AnyType operator()(Stack stack) const
    {
        const long nx = GetAny<long>((*enx)(stack)), ny = GetAny<long>((*eny)(stack)), nz = GetAny<long>((*enz)(stack));
        std::string why;
        if (nargs[1]) why = "label=";
        else if (nargs[2] && GetAny<long>((*nargs[2])(stack)) != 6) why = "flags=";
        else if (nx < 1 || ny < 1 || nz < 1 || (double)nx * ny * nz * 6 >= (double)(1 << 27)) why = "size";
        else if (!device_available()) why = "no CUDA device";
        if (!why.empty()) {
            notice("cube", why);
            return (*builtin)(stack);
        }
        const long region = nargs[0] ? GetAny<long>((*nargs[0])(stack)) : 0;
        Marks mk;
        ffcuda_mesh *dm = nullptr;
        FFC(ffcuda_mesh_cube(context(), (int)nx, (int)ny, (int)nz, &dm));
        const int n1 = (int)nx + 1, n2 = (int)ny + 1;
        Mesh3 *Th = nullptr;
        try {
            // vertex labels of BuildCube: one bit per face the vertex lies on (fflib/msh3.cpp:8010-8016)
            Th = mesh3_from_device(dm, [=](int p) {
                const int i = p % n1, j = (p / n1) % n2, k = p / (n1 * n2);
                return 1 * (i == 0) + 2 * (i == nx) + 4 * (j == 0) + 8 * (j == ny) + 16 * (k == 0) + 32 * (k == nz);
            }, mk);
        } catch (...) {
            ffcuda_mesh_destroy(dm);
            throw;
        }
        if (region != 0) { // the device copy carries region 0: not kept
            for (int k = 0; k < Th->nt; ++k) Th->elements[k].lab = (int)region;
            ffcuda_mesh_destroy(dm);
        } else
            keep_generated(*Th, dm);
        if (g_verbose) cout << "  -- ffcuda: cube(" << nx << "," << ny << "," << nz << ") on the device:" << mk.line << endl;
        Add2StackOfPtr2FreeRC(stack, Th);
        return Th;
    }
};
basicAC_F0::name_and_type CudaCubeOp::name_param[] = {{"region", &typeid(long)}, {"label", &typeid(KN_<long>)}, {"flags", &typeid(long)}};

struct CudaCube : public OneOperator {
    const OneOperator *builtin;
    explicit CudaCube(const OneOperator *b) : OneOperator(atype<pmesh3>(), atype<long>(), atype<long>(), atype<long>()), builtin(b) { pref = 100; }
    E_F0 *code(const basicAC_F0 &args) const
    {
        return new CudaCubeOp(args, builtin->code(args), t[0]->CastTo(args[0]), t[1]->CastTo(args[1]), t[2]->CastTo(args[2]));
    }
};

struct CudaLayersOp : public E_F0mps {
    Expression builtin, eTh, enmax, ezmin, ezmax;
    static const int n_name_param = 13;
    static basicAC_F0::name_and_type name_param[];
    Expression nargs[n_name_param];
    CudaLayersOp(const basicAC_F0 &args, Expression b, Expression th, Expression nm) : builtin(b), eTh(th), enmax(nm), ezmin(0), ezmax(0)
    {
        args.SetNameParam(n_name_param, name_param, nargs);
        if (nargs[0]) { // zbound=[zmin,zmax]; the built-in expression reports a malformed array
            const E_Array *a = dynamic_cast<const E_Array *>(nargs[0]);
            if (a && a->size() == 2) {
                ezmin = to<double>((*a)[0]);
                ezmax = to<double>((*a)[1]);
            }
        }
    }
    std::vector<int32_t> pairs(Stack stack, int a, int b) const
    { // region= / reftet= and friends: (old,new) pairs
        std::vector<int32_t> out;
        Expression e = nargs[a] ? nargs[a] : nargs[b];
        if (e) {
            KN_<long> v = GetAny<KN_<long>>((*e)(stack));
            if (v.N() % 2) ExecError("buildlayers: a label array needs an even number of entries");
            for (int i = 0; i < v.N(); ++i) out.push_back((int32_t)v[i]);
        }
        return out;
    }
    AnyType operator()(Stack stack) const
    {
        std::string why;
        if (nargs[1]) why = "transfo=";
        else if (nargs[7] || nargs[8]) why = "facemerge= / ptmerge=";
        else if (nargs[0] && !(ezmin && ezmax)) why = "zbound";
        else if (!device_available()) why = "no CUDA device";
        if (!why.empty()) {
            notice("buildlayers", why);
            return (*builtin)(stack);
        }
        MeshPoint *mp(MeshPointStack(stack)), mps = *mp;
        const Mesh *pTh = GetAny<const Mesh *>((*eTh)(stack));
        const int nlayer = (int)GetAny<long>((*enmax)(stack));
        ffassert(pTh && nlayer > 0);
        const Mesh &Th = *pTh;
        const int nbv = Th.nv, nbt = Th.nt, neb = Th.neb;
        // zmin, zmax, coef at the vertices, evaluated like BuildLayeMesh_Op does (fflib/msh3.cpp:4556-4585)
        std::vector<double> zmin(nbv, 0.), zmax(nbv, 1.), clayer(nbv, -1.);
        double maxdz = 0;
        for (int it = 0; it < nbt; ++it)
            for (int iv = 0; iv < 3; ++iv) {
                const int i = Th(it, iv);
                if (clayer[i] < 0) {
                    mp->setP(&Th, it, iv);
                    if (ezmin) zmin[i] = GetAny<double>((*ezmin)(stack));
                    if (ezmax) zmax[i] = GetAny<double>((*ezmax)(stack));
                    maxdz = std::max(maxdz, std::abs(zmin[i] - zmax[i]));
                    const double c = nargs[2] ? GetAny<double>((*nargs[2])(stack)) : 1.;
                    clayer[i] = std::max(0., std::min(1., c));
                }
            }
        *mp = mps;
        std::vector<int32_t> ni(nbv);
        const double epsz = maxdz * 1e-6;
        for (int i = 0; i < nbv; ++i) { // :4647-4657
            ni[i] = std::max(0, std::min(nlayer, (int)lrint(nlayer * std::max(clayer[i], 0.))));
            if (std::abs(zmin[i] - zmax[i]) < epsz) ni[i] = 0;
        }
        bool empty_prism = nlayer >= (1 << 15) || (double)nbt * nlayer * 3 >= (double)(1 << 27);
        for (int i = 0; i < nbv; ++i) empty_prism = empty_prism || clayer[i] < 0; // a vertex of no triangle: FreeFEM asserts (:4587)
        for (int it = 0; it < nbt && !empty_prism; ++it)
            empty_prism = ni[Th(it, 0)] == 0 && ni[Th(it, 1)] == 0 && ni[Th(it, 2)] == 0;
        if (empty_prism) { // FreeFEM stops the run there (:4662-4673), or the mesh is too large for one device: its business
            notice("buildlayers", "a triangle without layers, or size");
            return (*builtin)(stack);
        }
        const std::vector<int32_t> reg = pairs(stack, 3, 9), mid = pairs(stack, 4, 10), up = pairs(stack, 5, 11), down = pairs(stack, 6, 12);
        // the 2-D mesh as the device takes it
        Marks mk;
        std::vector<double> xy((size_t)nbv * 2);
        std::vector<int32_t> tri((size_t)nbt * 3), trilab(nbt), bconn((size_t)neb * 2), blab(neb), belem(neb), bface(neb), vlab2(nbv);
        for (int i = 0; i < nbv; ++i) {
            xy[2 * (size_t)i] = Th(i).x;
            xy[2 * (size_t)i + 1] = Th(i).y;
            vlab2[i] = Th(i).lab;
        }
        for (int k = 0; k < nbt; ++k) {
            for (int j = 0; j < 3; ++j) tri[3 * (size_t)k + j] = Th(k, j);
            trilab[k] = Th[k].lab;
        }
        for (int ib = 0; ib < neb; ++ib) {
            int ie;
            belem[ib] = Th.BoundaryElement(ib, ie);
            bface[ib] = ie;
            blab[ib] = Th.bedges[ib].lab;
            for (int j = 0; j < 2; ++j) bconn[2 * (size_t)ib + j] = Th(Th.bedges[ib][j]);
        }
        ffcuda_mesh *d2 = nullptr, *dm = nullptr;
        FFC(ffcuda_mesh_upload(context(), 2, nbv, xy.data(), nbt, tri.data(), trilab.data(), neb, bconn.data(), blab.data(), belem.data(),
                               bface.data(), &d2));
        const int rc = ffcuda_mesh_buildlayers(d2, nlayer, ni.data(), zmin.data(), zmax.data(), (int)reg.size() / 2, reg.data(),
                                               (int)mid.size() / 2, mid.data(), (int)up.size() / 2, up.data(), (int)down.size() / 2,
                                               down.data(), &dm);
        ffcuda_mesh_destroy(d2);
        if (rc) fail("ffcuda_mesh_buildlayers");
        // 3-D vertex -> its 2-D vertex: the columns follow each other (fflib/msh3.cpp:1016-1050); label of the 2-D vertex
        std::vector<int32_t> col;
        col.reserve((size_t)nbv * (nlayer + 1));
        for (int i = 0; i < nbv; ++i) col.insert(col.end(), (size_t)ni[i] + 1, vlab2[i]);
        Mesh3 *Th3 = nullptr;
        try {
            Th3 = mesh3_from_device(dm, [&](int p) { return col[p]; }, mk);
        } catch (...) {
            ffcuda_mesh_destroy(dm);
            throw;
        }
        keep_generated(*Th3, dm);
        if (g_verbose) cout << "  -- ffcuda: buildlayers(" << nbt << " triangles, " << nlayer << " layers) on the device:" << mk.line << endl;
        Add2StackOfPtr2FreeRC(stack, Th3);
        return Th3;
    }
};
basicAC_F0::name_and_type CudaLayersOp::name_param[] = { // BuildLayeMesh_Op::name_param, fflib/msh3.cpp:4254-4268
    {"zbound", &typeid(E_Array)},          {"transfo", &typeid(E_Array)},        {"coef", &typeid(double)},
    {"reftet", &typeid(KN_<long>)},        {"reffacemid", &typeid(KN_<long>)},   {"reffaceup", &typeid(KN_<long>)},
    {"reffacelow", &typeid(KN_<long>)},    {"facemerge", &typeid(long)},         {"ptmerge", &typeid(double)},
    {"region", &typeid(KN_<long>)},        {"labelmid", &typeid(KN_<long>)},     {"labelup", &typeid(KN_<long>)},
    {"labeldown", &typeid(KN_<long>)}};

struct CudaLayers : public OneOperator {
    const OneOperator *builtin;
    explicit CudaLayers(const OneOperator *b) : OneOperator(atype<pmesh3>(), atype<pmesh>(), atype<long>()), builtin(b) { pref = 100; }
    E_F0 *code(const basicAC_F0 &args) const
    {
        return new CudaLayersOp(args, builtin->code(args), t[0]->CastTo(args[0]), t[1]->CastTo(args[1]));
    }
};

// the built-in operator of a global function for exact argument types (before ours is added)
const OneOperator *builtin_operator(const char *name, const ArrayOfaType &at)
{
    C_F0 f = Global.Find(name); // (a Polymorphic answers Empty(): not asked)
    const Polymorphic *p = dynamic_cast<const Polymorphic *>(f.LeftValue());
    return p ? p->FindWithOutCast("(", at) : nullptr;
}

} // namespace

static void Load_Init()
{
    g_verbose = env_on("FFCUDA_VERBOSE");
    g_strict = env_on("FFCUDA_STRICT");
    g_check = env_on("FFCUDA_CHECK");
    if (const char *e = getenv("FFCUDA_SAMPLE_MIN")) g_sample_min = atoi(e); // (testing knobs of the coefficient grouping)
    if (const char *e = getenv("FFCUDA_SAMPLE_N")) g_sample_n = atoi(e);
    g_fe_dofs = !env_on("FFCUDA_NO_FE_DOFS");
    g_explain = env_on("FFCUDA_EXPLAIN");
    g_rect = env_on("FFCUDA_RECT");
    if (g_fe_dofs) {
        find_fe_node_functions<pfer>(g_fe2, 2);
        find_fe_node_functions<pf3r>(g_fe3, 3);
    }
    if (const char *e = getenv("FFCUDA_NGPU")) g_ngpu = std::max(1, std::min(16, atoi(e)));
    if (const char *e = getenv("FFCUDA_NGPU_MIN_N")) g_ngpu_min_n = atol(e);
    if (env_on("FFCUDA_DISABLE")) {
        if (verbosity) cout << " load: ffcuda disabled by FFCUDA_DISABLE" << endl;
        return;
    }
    if (verbosity) cout << " load: ffcuda (GPU assembly of P1/P2 varf + Jacobi-CG / GMRES; FreeFEM keeps everything else)" << endl;
    // The CUDA context (driver initialisation, module load: about a second in a fresh process) is created here, while the
    // script is being parsed, not inside the first `matrix A = ...` statement.  Without a device nothing happens now: the
    // first claimed form raises the "no CPU fallback" error as before.  FFCUDA_LAZY_INIT=1 keeps the old behaviour.
    if (!env_on("FFCUDA_LAZY_INIT") && !g_ctx) {
        const char *d = getenv("FFCUDA_DEVICE");
        ffcuda_ctx *c = nullptr;
        if (ffcuda_ctx_create(d ? atoi(d) : 0, &c) == 0) g_ctx = c;
        // ... and with FFCUDA_NGPU > 1 the other contexts and the communicator (seconds of NCCL start-up), for the same reason
        if (g_ctx && g_ngpu > 1) {
            try {
                gang_ready();
            } catch (...) { // (FFCUDA_STRICT turns the notice into an error: raised again by the first solve)
            }
        }
    }
    // 1. matrices: "<-" constructs (init = 1), "=" assigns (init = 0)  (fflib/lgfem.cpp:6669,6673,6823,6826)
    TheOperators->Add("<-", new CudaMatrixOp<Mesh, v_fes>(1), new CudaMatrixOp<Mesh3, v_fes3>(1));
    TheOperators->Add("=", new CudaMatrixOp<Mesh, v_fes>(0), new CudaMatrixOp<Mesh3, v_fes3>(0));
    // 2. right-hand sides (fflib/lgfem.cpp:6668,6672,6686,6688)
    TheOperators->Add("=", new CudaRhsOp<Mesh, v_fes>(atype<KN_<double>>(), false, false),
                      new CudaRhsOp<Mesh3, v_fes3>(atype<KN_<double>>(), false, false));
    TheOperators->Add("<-", new CudaRhsOp<Mesh, v_fes>(atype<KN<double> *>(), true, true),
                      new CudaRhsOp<Mesh3, v_fes3>(atype<KN<double> *>(), true, true));
    // 3. solver
    addsolver<SolverCudaCG>("FFCUDACG", 10, 0);
    TheFFSolver<int, double>::ChangeSolver("CG", "FFCUDACG");
    addsolver<SolverCudaGMRES>("FFCUDAGMRES", 10, 0);
    TheFFSolver<int, double>::ChangeSolver("GMRES", "FFCUDAGMRES");
    // 4. problem / solve
    if (!env_on("FFCUDA_NO_PROBLEM")) {
        repoint_solve_type<false, Problem>();
        repoint_solve_type<true, Solve>();
    }
    // 5. mesh generators on the device (cube without transformation, buildlayers)
    if (!env_on("FFCUDA_NO_MESH")) {
        if (const OneOperator *b = builtin_operator("cube", ArrayOfaType(atype<long>(), atype<long>(), atype<long>())))
            Global.Add("cube", "(", new CudaCube(b));
        if (const OneOperator *b = builtin_operator("buildlayers", ArrayOfaType(atype<pmesh>(), atype<long>())))
            Global.Add("buildlayers", "(", new CudaLayers(b));
    }
}

LOADFUNC(Load_Init)
