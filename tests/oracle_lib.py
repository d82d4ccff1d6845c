"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference algorithm).
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

OP_ID, OP_DX, OP_DY, OP_DZ = 0, 1, 2, 6


class BTerm(C.Structure):
    _fields_ = [("ucomp", C.c_int32), ("uop", C.c_int32), ("vcomp", C.c_int32), ("vop", C.c_int32), ("coef", C.c_double)]


class LTerm(C.Structure):
    _fields_ = [("vcomp", C.c_int32), ("vop", C.c_int32), ("coef", C.c_double)]


def build():
    so = os.path.join(_ORACLE_DIR, "liboracle.so")
    src = [os.path.join(_ORACLE_DIR, f) for f in ("fforacle.c", "fforacle.h")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ffo_assemble_coo.restype = C.c_int64
        _LIB.ffo_assemble_coo_boundary.restype = C.c_int64
        _LIB.ffo_assemble_coo_qcoef.restype = C.c_int64
        _LIB.ffo_assemble_coo_boundary_qcoef.restype = C.c_int64
        _LIB.ffo_assemble_coo_rect.restype = C.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def bterms(lst):
    arr = (BTerm * len(lst))()
    for k, (uc, uo, vc, vo, c) in enumerate(lst):
        arr[k] = BTerm(uc, uo, vc, vo, c)
    return arr


def lterms(lst):
    arr = (LTerm * len(lst))()
    for k, (vc, vo, c) in enumerate(lst):
        arr[k] = LTerm(vc, vo, c)
    return arr


def quadrature(dim, name):
    pts = np.zeros(16 * dim)
    w = np.zeros(16)
    n = lib().ffo_quadrature(dim, name.encode(), _p(pts, C.c_double), _p(w, C.c_double))
    assert n > 0, name
    return pts[: n * dim].reshape(n, dim).copy(), w[:n].copy()


def nloc(dim, order):
    return lib().ffo_nloc(dim, order)


def cube(nx, ny, nz):
    nv, nt, nbe = C.c_int(), C.c_int(), C.c_int()
    lib().ffo_cube_sizes(nx, ny, nz, C.byref(nv), C.byref(nt), C.byref(nbe))
    nv, nt, nbe = nv.value, nt.value, nbe.value
    m = dict(dim=3, xyz=np.zeros((nv, 3)), conn=np.zeros((nt, 4), np.int32), elab=np.zeros(nt, np.int32),
             bconn=np.zeros((nbe, 3), np.int32), blab=np.zeros(nbe, np.int32), belem=np.zeros(nbe, np.int32),
             bface=np.zeros(nbe, np.int32))
    lib().ffo_cube(nx, ny, nz, _p(m["xyz"], C.c_double), _p(m["conn"], C.c_int32), _p(m["elab"], C.c_int32),
                   _p(m["bconn"], C.c_int32), _p(m["blab"], C.c_int32), _p(m["belem"], C.c_int32), _p(m["bface"], C.c_int32))
    return m


def square(nx, ny):
    nv, nt, nbe = C.c_int(), C.c_int(), C.c_int()
    lib().ffo_square_sizes(nx, ny, C.byref(nv), C.byref(nt), C.byref(nbe))
    nv, nt, nbe = nv.value, nt.value, nbe.value
    m = dict(dim=2, xyz=np.zeros((nv, 2)), conn=np.zeros((nt, 3), np.int32), elab=np.zeros(nt, np.int32),
             bconn=np.zeros((nbe, 2), np.int32), blab=np.zeros(nbe, np.int32), belem=np.zeros(nbe, np.int32),
             bface=np.zeros(nbe, np.int32))
    lib().ffo_square(nx, ny, _p(m["xyz"], C.c_double), _p(m["conn"], C.c_int32), _p(m["elab"], C.c_int32),
                     _p(m["bconn"], C.c_int32), _p(m["blab"], C.c_int32), _p(m["belem"], C.c_int32), _p(m["bface"], C.c_int32))
    return m


def buildlayers(mesh2, nlayer, ni=None, zmin=None, zmax=None, regmap=(), midmap=(), upmap=(), downmap=()):
    """layered mesh over the 2-D mesh dict (keys xyz, conn, elab, blab, belem, bface); maps are flat (old,new,...) lists"""
    xy, tri, trilab = _f64(mesh2["xyz"]), _i32(mesh2["conn"]), _i32(mesh2["elab"])
    bl, be, bf = _i32(mesh2["blab"]), _i32(mesh2["belem"]), _i32(mesh2["bface"])
    nv2, nt2, nbe2 = xy.shape[0], tri.shape[0], bl.shape[0]
    ni = np.full(nv2, nlayer, np.int32) if ni is None else _i32(ni)
    zmin = np.zeros(nv2) if zmin is None else _f64(zmin)
    zmax = np.ones(nv2) if zmax is None else _f64(zmax)
    maps = [_i32(np.asarray(m_, dtype=np.int32).ravel()) for m_ in (regmap, midmap, upmap, downmap)]
    nv, nt, nbe = C.c_int(), C.c_int(), C.c_int()
    lib().ffo_buildlayers_sizes(nv2, nt2, _p(tri, C.c_int32), nbe2, _p(be, C.c_int32), _p(bf, C.c_int32), nlayer,
                                _p(ni, C.c_int32), C.byref(nv), C.byref(nt), C.byref(nbe))
    nv, nt, nbe = nv.value, nt.value, nbe.value
    m = dict(dim=3, xyz=np.zeros((nv, 3)), conn=np.zeros((nt, 4), np.int32), elab=np.zeros(nt, np.int32),
             bconn=np.zeros((nbe, 3), np.int32), blab=np.zeros(nbe, np.int32), belem=np.zeros(nbe, np.int32),
             bface=np.zeros(nbe, np.int32))
    margs = []
    for m_ in maps:
        margs += [m_.size // 2, _p(m_, C.c_int32)]
    lib().ffo_buildlayers(nv2, _p(xy, C.c_double), nt2, _p(tri, C.c_int32), _p(trilab, C.c_int32), nbe2, _p(bl, C.c_int32),
                          _p(be, C.c_int32), _p(bf, C.c_int32), nlayer, _p(ni, C.c_int32), _p(zmin, C.c_double),
                          _p(zmax, C.c_double), *margs, _p(m["xyz"], C.c_double), _p(m["conn"], C.c_int32),
                          _p(m["elab"], C.c_int32), _p(m["bconn"], C.c_int32), _p(m["blab"], C.c_int32),
                          _p(m["belem"], C.c_int32), _p(m["bface"], C.c_int32))
    return m


def p2_nodes_3d(nv, conn):
    conn = _i32(conn)
    nt = conn.shape[0]
    e2n = np.zeros((nt, 10), np.int32)
    nn = lib().ffo_p2_nodes_3d(nv, nt, _p(conn, C.c_int32), _p(e2n, C.c_int32))
    return e2n, nn


def assemble_coo(mesh, order, ncomp, elem2node, terms, qpts, qw, labels=None):
    dim = mesh["dim"]
    xyz, conn, elab = _f64(mesh["xyz"]), _i32(mesh["conn"]), _i32(mesh["elab"])
    nt = conn.shape[0]
    nd = nloc(dim, order) * ncomp
    cap = nt * nd * nd
    ci, cj, ca = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    e2n = _i32(elem2node)
    lab = _i32(labels)
    bt = bterms(terms)
    qpts, qw = _f64(qpts), _f64(qw)
    nnz = lib().ffo_assemble_coo(dim, xyz.shape[0], _p(xyz, C.c_double), nt, _p(conn, C.c_int32), _p(elab, C.c_int32),
                                 order, ncomp, _p(e2n, C.c_int32), len(terms), bt, len(qw), _p(qpts, C.c_double),
                                 _p(qw, C.c_double), 0 if lab is None else len(lab), _p(lab, C.c_int32),
                                 _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double))
    return ci[:nnz].copy(), cj[:nnz].copy(), ca[:nnz].copy()


def assemble_coo_qcoef(mesh, order, ncomp, elem2node, terms, qpts, qw, cq):
    """assemble_coo with every term multiplied by the coefficient whose values at the quadrature nodes are cq[k, q]"""
    dim = mesh["dim"]
    xyz, conn, elab = _f64(mesh["xyz"]), _i32(mesh["conn"]), _i32(mesh["elab"])
    nt = conn.shape[0]
    nd = nloc(dim, order) * ncomp
    cap = nt * nd * nd
    ci, cj, ca = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    e2n = _i32(elem2node)
    bt = bterms(terms)
    qpts, qw, cq = _f64(qpts), _f64(qw), _f64(cq)
    nnz = lib().ffo_assemble_coo_qcoef(dim, xyz.shape[0], _p(xyz, C.c_double), nt, _p(conn, C.c_int32), _p(elab, C.c_int32),
                                       order, ncomp, _p(e2n, C.c_int32), len(terms), bt, len(qw), _p(qpts, C.c_double),
                                       _p(qw, C.c_double), _p(cq, C.c_double), _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double))
    return ci[:nnz].copy(), cj[:nnz].copy(), ca[:nnz].copy()


def assemble_coo_rect(mesh, order_v, ncomp_v, e2n_v, order_u, ncomp_u, e2n_u, terms, qpts, qw, labels=None):
    """COO (HashMatrix insertion order) of a rectangular matrix: rows = dofs of the test space, columns = dofs of the space
    of the unknown, both on the same mesh"""
    dim = mesh["dim"]
    xyz, conn, elab = _f64(mesh["xyz"]), _i32(mesh["conn"]), _i32(mesh["elab"])
    nt = conn.shape[0]
    cap = nt * nloc(dim, order_v) * ncomp_v * nloc(dim, order_u) * ncomp_u
    ci, cj, ca = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    ev, eu, lab = _i32(e2n_v), _i32(e2n_u), _i32(labels)
    bt = bterms(terms)
    qpts, qw = _f64(qpts), _f64(qw)
    nnz = lib().ffo_assemble_coo_rect(dim, _p(xyz, C.c_double), nt, _p(conn, C.c_int32), _p(elab, C.c_int32),
                                      order_v, ncomp_v, _p(ev, C.c_int32), order_u, ncomp_u, _p(eu, C.c_int32),
                                      len(terms), bt, len(qw), _p(qpts, C.c_double), _p(qw, C.c_double),
                                      0 if lab is None else len(lab), _p(lab, C.c_int32),
                                      _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double))
    return ci[:nnz].copy(), cj[:nnz].copy(), ca[:nnz].copy()


def coo_to_csr(n, ci, cj, ca):
    ci, cj, ca = _i32(ci), _i32(cj), _f64(ca)
    nnz = len(ci)
    rp, col, val = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
    lib().ffo_coo_to_csr(n, C.c_int64(nnz), _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double),
                         _p(rp, C.c_int32), _p(col, C.c_int32), _p(val, C.c_double))
    return rp, col, val


def assemble_rhs(mesh, order, ncomp, elem2node, ndof, terms, qpts, qw, labels=None):
    dim = mesh["dim"]
    xyz, conn, elab = _f64(mesh["xyz"]), _i32(mesh["conn"]), _i32(mesh["elab"])
    e2n, lab = _i32(elem2node), _i32(labels)
    b = np.zeros(ndof)
    lt = lterms(terms)
    qpts, qw = _f64(qpts), _f64(qw)
    lib().ffo_assemble_rhs(dim, xyz.shape[0], _p(xyz, C.c_double), conn.shape[0], _p(conn, C.c_int32), _p(elab, C.c_int32),
                           order, ncomp, _p(e2n, C.c_int32), ndof, len(terms), lt, len(qw), _p(qpts, C.c_double),
                           _p(qw, C.c_double), 0 if lab is None else len(lab), _p(lab, C.c_int32), _p(b, C.c_double))
    return b


def quad_points_xyz(mesh, qpts):
    """physical coordinates of the quadrature nodes of every element: (nt, nq, dim)"""
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    qpts = _f64(qpts)
    lam = np.concatenate([1.0 - qpts.sum(axis=1, keepdims=True), qpts], axis=1)      # (nq, dim+1) barycentric
    return np.einsum("qa,kad->kqd", lam, xyz[conn])


def assemble_rhs_qvalues(mesh, order, ncomp, elem2node, b, qpts, qw, fq):
    """adds int(f v) with f given at the quadrature nodes, fq[c, k, q], to b (returns a copy)"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    e2n = _i32(elem2node)
    b, qpts, qw, fq = _f64(b).copy(), _f64(qpts), _f64(qw), _f64(fq)
    lib().ffo_assemble_rhs_qvalues(dim, _p(xyz, C.c_double), conn.shape[0], _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32),
                                   len(qw), _p(qpts, C.c_double), _p(qw, C.c_double), _p(fq, C.c_double), _p(b, C.c_double))
    return b


def assemble_rhs_qterms(mesh, order, ncomp, elem2node, b, qpts, qw, fq):
    """adds int(sum_s f_s d^s v) with the f_s given at the quadrature nodes, fq[c, s, k, q] (s = 0 value, 1..dim derivatives)"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    e2n = _i32(elem2node)
    b, qpts, qw, fq = _f64(b).copy(), _f64(qpts), _f64(qw), _f64(fq)
    assert fq.shape[1] == dim + 1
    lib().ffo_assemble_rhs_qterms(dim, _p(xyz, C.c_double), conn.shape[0], _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32),
                                  len(qw), _p(qpts, C.c_double), _p(qw, C.c_double), _p(fq, C.c_double), _p(b, C.c_double))
    return b


def assemble_rhs_boundary(mesh, order, ncomp, elem2node, b, terms, qpts, qw, labels=None):
    """adds the boundary integrals int2d(Th3,labels)(...) / int1d(Th,labels)(...) of a linear form to b (returns a copy)"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    blab, belem, bface = _i32(mesh["blab"]), _i32(mesh["belem"]), _i32(mesh["bface"])
    e2n, lab = _i32(elem2node), _i32(labels)
    b = _f64(b).copy()
    lt = lterms(terms)
    qpts, qw = _f64(qpts), _f64(qw)
    lib().ffo_assemble_rhs_boundary(dim, _p(xyz, C.c_double), _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32), len(blab),
                                    _p(blab, C.c_int32), _p(belem, C.c_int32), _p(bface, C.c_int32), len(terms), lt, len(qw),
                                    _p(qpts, C.c_double), _p(qw, C.c_double), 0 if lab is None else len(lab), _p(lab, C.c_int32),
                                    _p(b, C.c_double))
    return b


def assemble_coo_boundary(mesh, order, ncomp, elem2node, terms, qpts, qw, labels=None):
    """COO of the boundary integrals int2d(Th3,labels)(c u v) / int1d(Th,labels)(c u v) of a bilinear form (Robin terms):
    one entry per distinct couple of the elements adjacent to the labelled boundary elements"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    blab, belem, bface = _i32(mesh["blab"]), _i32(mesh["belem"]), _i32(mesh["bface"])
    e2n, lab = _i32(elem2node), _i32(labels)
    nd = nloc(dim, order) * ncomp
    cap = max(len(blab), 1) * nd * nd
    ci, cj, ca = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    bt = bterms(terms)
    qpts, qw = _f64(qpts), _f64(qw)
    nnz = lib().ffo_assemble_coo_boundary(dim, _p(xyz, C.c_double), _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32),
                                          len(blab), _p(blab, C.c_int32), _p(belem, C.c_int32), _p(bface, C.c_int32),
                                          len(terms), bt, len(qw), _p(qpts, C.c_double), _p(qw, C.c_double),
                                          0 if lab is None else len(lab), _p(lab, C.c_int32),
                                          _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double))
    return ci[:nnz].copy(), cj[:nnz].copy(), ca[:nnz].copy()


def coo_add(n, a, b):
    """HashMatrix accumulation of two COO matrices with distinct couples each: couples of a keep their place, new couples
    of b are appended (the order of a COO list is irrelevant once sorted into CSR)"""
    (ai, aj, aa), (bi, bj, ba) = a, b
    ka = ai.astype(np.int64) * n + aj
    kb = bi.astype(np.int64) * n + bj
    order = np.argsort(ka, kind="stable")
    pos = np.searchsorted(ka[order], kb)
    pos = np.minimum(pos, len(ka) - 1) if len(ka) else pos
    hit = (ka[order][pos] == kb) if len(ka) else np.zeros(len(kb), bool)
    out = aa.copy()
    out[order[pos[hit]]] += ba[hit]
    return (np.concatenate([ai, bi[~hit]]).astype(np.int32), np.concatenate([aj, bj[~hit]]).astype(np.int32),
            np.concatenate([out, ba[~hit]]))


_NVFACE = np.array([[3, 2, 1], [0, 2, 3], [3, 1, 0], [0, 1, 2]])
_NVEDGE = np.array([[1, 2], [2, 0], [0, 1]])


def bquad_points_xyz(mesh, fqpts):
    """physical coordinates of the face quadrature nodes of every boundary element, PBord(ie, q): (nbe, nq, dim)"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    belem, bface = _i32(mesh["belem"]), _i32(mesh["bface"])
    fq = _f64(fqpts).reshape(len(fqpts), dim - 1)
    lam = np.concatenate([1.0 - fq.sum(axis=1, keepdims=True), fq], axis=1)          # (nq, dim) weights of the face vertices
    fv = (_NVFACE if dim == 3 else _NVEDGE)[bface]                                   # (nbe, dim) local vertices of the face
    verts = np.take_along_axis(conn[belem], fv, axis=1)                              # (nbe, dim) global vertices
    return np.einsum("qa,ead->eqd", lam, xyz[verts])


def assemble_rhs_boundary_qvalues(mesh, order, ncomp, elem2node, b, qpts, qw, gq):
    """adds the boundary integral of g v with g given at the face quadrature nodes, gq[c, ib, q], to b (returns a copy)"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    blab, belem, bface = _i32(mesh["blab"]), _i32(mesh["belem"]), _i32(mesh["bface"])
    e2n = _i32(elem2node)
    b, qpts, qw, gq = _f64(b).copy(), _f64(qpts), _f64(qw), _f64(gq)
    lib().ffo_assemble_rhs_boundary_qvalues(dim, _p(xyz, C.c_double), _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32), len(blab),
                                            _p(blab, C.c_int32), _p(belem, C.c_int32), _p(bface, C.c_int32), len(qw),
                                            _p(qpts, C.c_double), _p(qw, C.c_double), _p(gq, C.c_double), _p(b, C.c_double))
    return b


def assemble_coo_boundary_qcoef(mesh, order, ncomp, elem2node, terms, qpts, qw, cq, labels=None):
    """assemble_coo_boundary with every term multiplied by the coefficient given at the face quadrature nodes, cq[ib, q]"""
    dim = mesh["dim"]
    xyz, conn = _f64(mesh["xyz"]), _i32(mesh["conn"])
    blab, belem, bface = _i32(mesh["blab"]), _i32(mesh["belem"]), _i32(mesh["bface"])
    e2n, lab = _i32(elem2node), _i32(labels)
    nd = nloc(dim, order) * ncomp
    cap = max(len(blab), 1) * nd * nd
    ci, cj, ca = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    bt = bterms(terms)
    qpts, qw, cq = _f64(qpts), _f64(qw), _f64(cq)
    nnz = lib().ffo_assemble_coo_boundary_qcoef(dim, _p(xyz, C.c_double), _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32),
                                                len(blab), _p(blab, C.c_int32), _p(belem, C.c_int32), _p(bface, C.c_int32),
                                                len(terms), bt, len(qw), _p(qpts, C.c_double), _p(qw, C.c_double),
                                                0 if lab is None else len(lab), _p(lab, C.c_int32), _p(cq, C.c_double),
                                                _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double))
    return ci[:nnz].copy(), cj[:nnz].copy(), ca[:nnz].copy()


def face_quadrature(dim):
    """default rule of a boundary integral (qforder = 6): 7-point rule on the faces of tetrahedra, 3-point Gauss-Legendre
    on the edges of triangles (QuadratureFormular.cpp: QuadratureFormular_T_5, QF_GaussLegendre3)"""
    if dim == 3:
        return quadrature(2, "qf5pT")
    r = np.sqrt(3.0 / 5.0)
    return np.array([[(1 - r) / 2], [0.5], [(1 + r) / 2]]), np.array([5.0 / 18, 8.0 / 18, 5.0 / 18])


def bc_pairs(mesh, order, ncomp, elem2node, labels, compmask, values):
    dim = mesh["dim"]
    conn = _i32(mesh["conn"])
    blab, belem, bface = _i32(mesh["blab"]), _i32(mesh["belem"]), _i32(mesh["bface"])
    e2n, lab = _i32(elem2node), _i32(labels)
    vals = _f64(values)
    cap = len(blab) * ncomp * nloc(dim, order)
    od, ov = np.zeros(cap, np.int32), np.zeros(cap)
    n = lib().ffo_bc_pairs(dim, conn.shape[0], _p(conn, C.c_int32), order, ncomp, _p(e2n, C.c_int32), len(blab),
                           _p(blab, C.c_int32), _p(belem, C.c_int32), _p(bface, C.c_int32), len(lab), _p(lab, C.c_int32),
                           compmask, _p(vals, C.c_double), _p(od, C.c_int32), _p(ov, C.c_double))
    return od[:n].copy(), ov[:n].copy()


def bc_matrix_coo(ci, cj, ca, n, dofs, tgv):
    ci, cj, dofs = _i32(ci), _i32(cj), _i32(dofs)
    ca = _f64(ca).copy()
    lib().ffo_bc_matrix_coo(C.c_int64(len(ci)), _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double), n, len(dofs),
                            _p(dofs, C.c_int32), C.c_double(tgv))
    return ca


def bc_rhs(b, dofs, vals, tgv):
    b = _f64(b).copy()
    dofs, vals = _i32(dofs), _f64(vals)
    lib().ffo_bc_rhs(_p(b, C.c_double), len(dofs), _p(dofs, C.c_int32), _p(vals, C.c_double), C.c_double(tgv))
    return b


def spmv_coo(n, ci, cj, ca, x):
    ci, cj, ca, x = _i32(ci), _i32(cj), _f64(ca), _f64(x)
    y = np.zeros(n)
    lib().ffo_spmv_coo(n, C.c_int64(len(ci)), _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double), _p(x, C.c_double),
                       _p(y, C.c_double))
    return y


def cg(n, ci, cj, ca, b, x0, eps=1e-6, itmax=0, tgv=1e30):
    ci, cj, ca, b = _i32(ci), _i32(cj), _f64(ca), _f64(b)
    x = _f64(x0).copy()
    it, g = C.c_int(), C.c_double()
    ret = lib().ffo_cg(n, C.c_int64(len(ci)), _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double), _p(b, C.c_double),
                       _p(x, C.c_double), C.c_double(eps), itmax, C.c_double(tgv), C.byref(it), C.byref(g))
    return x, it.value, ret, g.value


def gmres(n, ci, cj, ca, b, x0, eps=1e-6, itmax=0, nbkrylov=1000, tgv=1e30):
    """SolverGMRES of the reference (fgmres, right Jacobi preconditioner): returns x, iterations, converged, relative residual"""
    ci, cj, ca, b = _i32(ci), _i32(cj), _f64(ca), _f64(b)
    x = _f64(x0).copy()
    it, r = C.c_int(), C.c_double()
    ret = lib().ffo_gmres(n, C.c_int64(len(ci)), _p(ci, C.c_int32), _p(cj, C.c_int32), _p(ca, C.c_double), _p(b, C.c_double),
                          _p(x, C.c_double), C.c_double(eps), itmax, nbkrylov, C.c_double(tgv), C.byref(it), C.byref(r))
    return x, it.value, ret, r.value
