"""CPU test of the ARITHMETIC of the rectangular assembly kernel: csrc/rect_row.cuh (the row routine k_asm_rect compiles) is
built for the host by g++ into a scratch library and run row by row on the 6 rectangular fixtures dumped from the reference;
incidence lists and node-level pattern are formed here with numpy.  Pattern bit-exact, values 1e-12 of the largest entry.
(The launch plumbing and the device symbolic phase around it are what tests/test_zz_gpu_rect.py checks on the GPU.)"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import ff_cases as fc
import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no g++ to build the host harness")
_LIB = None
SLOT = {fc.ID: 0, fc.DX: 1, fc.DY: 2, fc.DZ: 3}


def host_lib():
    global _LIB
    if _LIB is None:
        out = os.path.join(tempfile.mkdtemp(prefix="host_rect_"), "libhost_rect.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out, os.path.join(HERE, "host_rect.cpp")], check=True)
        _LIB = C.CDLL(out)
    return _LIB


def node_incidence(e2n, nnodes):
    """CSR lists of (element << 4 | local node) per node, each sorted by element: the transpose of the element -> node table"""
    nt, nloc = e2n.shape
    node = e2n.ravel()
    rec = (np.repeat(np.arange(nt, dtype=np.uint32), nloc) << np.uint32(4)) | np.tile(np.arange(nloc, dtype=np.uint32), nt)
    o = np.lexsort((rec, node))
    ptr = np.zeros(nnodes + 1, np.int32)
    np.add.at(ptr, node + 1, 1)
    return np.cumsum(ptr).astype(np.int32), np.ascontiguousarray(rec[o], dtype=np.uint32)


def node_pattern(incptr, inc, e2n_u):
    """sorted distinct column nodes of every row node"""
    rows = []
    for i in range(len(incptr) - 1):
        els = inc[incptr[i]:incptr[i + 1]] >> np.uint32(4)
        rows.append(np.unique(e2n_u[els].ravel()) if len(els) else np.zeros(0, np.int32))
    nrowptr = np.zeros(len(rows) + 1, np.int32)
    nrowptr[1:] = np.cumsum([len(r) for r in rows])
    return nrowptr, np.ascontiguousarray(np.concatenate(rows), dtype=np.int32)


def bary(dim, qp):
    lam = np.zeros((len(qp), 4))
    lam[:, 1:dim + 1] = np.asarray(qp).reshape(len(qp), dim)
    lam[:, 0] = 1.0 - lam[:, 1:dim + 1].sum(axis=1)
    return lam


def host_assemble(g, ov, cv, ev, ou, cu, eu, terms, qp, qw, labels=None, nnodes_v=None):
    """(rowptr, colind, vals) dof-level CSR as the device entry lays it out"""
    dim = int(g["dim"])
    nnv = int(ev.max()) + 1 if nnodes_v is None else int(nnodes_v)
    incptr, inc = node_incidence(ev, nnv)
    nrowptr, ncol = node_pattern(incptr, inc, eu)
    vals = np.zeros(cv * cu * len(ncol))
    lam = np.ascontiguousarray(bary(dim, qp))
    coef = np.array([t[4] for t in terms], dtype=np.float64)
    tdesc = np.array([[t[2], t[0], SLOT[t[3]], SLOT[t[1]]] for t in terms], dtype=np.int32)
    xyz, conn, elab = np.ascontiguousarray(g["xyz"], np.float64), np.ascontiguousarray(g["conn"], np.int32), np.ascontiguousarray(g["elab"], np.int32)
    lab = None if labels is None else np.ascontiguousarray(labels, np.int32)
    qw = np.ascontiguousarray(qw, np.float64)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t)) if a is not None else None  # noqa: E731
    rc = host_lib().host_rect_assemble(dim, nnv, p(incptr, C.c_int32), p(inc, C.c_uint32), p(conn, C.c_int32), p(elab, C.c_int32),
                                       p(xyz, C.c_double), dim, p(eu, C.c_int32), ov, cv, ou, cu, eu.shape[1], len(qw), p(lam, C.c_double),
                                       p(qw, C.c_double), len(terms), p(coef, C.c_double), p(tdesc, C.c_int32),
                                       0 if lab is None else len(lab), p(lab, C.c_int32), p(nrowptr, C.c_int32), p(ncol, C.c_int32),
                                       p(vals, C.c_double))
    assert rc == 0
    # dof-level CSR of the component-block layout: row (i, c) starts at cu * (cv * nrowptr[i] + c * L)
    L = np.diff(nrowptr)
    rowptr = np.zeros(nnv * cv + 1, np.int64)
    rowptr[1:] = np.cumsum(np.repeat(L * cu, cv))
    colind = np.concatenate([np.tile((ncol[nrowptr[i]:nrowptr[i + 1], None] * cu + np.arange(cu)).ravel(), cv) for i in range(nnv)])
    return rowptr, colind.astype(np.int32), vals


@pytest.mark.parametrize("name", sorted(fc.RECT_CASES))
def test_row_routine_on_the_host_against_the_reference(name):
    (ov, cv), (ou, cu), terms, qname = fc.RECT_CASES[name]
    g = fc.load(name)
    ev, eu = fc.rect_elem2node(g, "Vh", cv), fc.rect_elem2node(g, "Uh", cu)
    qp, qw = ol.quadrature(g["dim"], qname)
    rp, col, val = host_assemble(g, ov, cv, ev, ou, cu, eu, terms, qp, qw)
    n, m = int(g["n"]), int(g["m"])
    assert len(rp) == n + 1 and col.max() < m
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert np.array_equal(rows, g["coo_i"]) and np.array_equal(col, g["coo_j"])  # the fixture's [I,J,C] is sorted by (i, j)
    assert np.max(np.abs(val - g["coo_a"])) <= 1e-12 * np.abs(g["coo_a"]).max()


def test_row_routine_region_filter_and_params_size():
    """labels: only the listed regions contribute, the pattern keeps every couple (as the device entry documents)"""
    assert host_lib().host_rect_sizeof_params() % 8 == 0
    name = "rect3d_p1vec_p1"
    (ov, cv), (ou, cu), terms, qname = fc.RECT_CASES[name]
    g = fc.load(name)
    g["elab"] = (np.arange(len(g["elab"])) % 3).astype(np.int32)
    ev, eu = fc.rect_elem2node(g, "Vh", cv), fc.rect_elem2node(g, "Uh", cu)
    qp, qw = ol.quadrature(3, qname)
    rp, col, val = host_assemble(g, ov, cv, ev, ou, cu, eu, terms, qp, qw, labels=[0, 2])
    mesh = {k: g[k] for k in ("dim", "xyz", "conn", "elab")}
    ci, cj, ca = ol.assemble_coo_rect(mesh, ov, cv, ev, ou, cu, eu, terms, qp, qw, labels=[0, 2])
    n, m = int(g["n"]), int(g["m"])
    dense = np.zeros((n, m))
    dense[np.repeat(np.arange(n), np.diff(rp)), col] = val
    ref = np.zeros((n, m))
    ref[ci, cj] = ca
    assert np.max(np.abs(dense - ref)) <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("name", sorted(fc.MIXED_CASES))
def test_row_routine_on_the_blocks_of_a_mixed_order_space(name):
    """what the plugin does for [P2,P2,P1] / [P2,P2,P2,P1] spaces: one call of the rectangular entry per couple of components,
    scalar spaces whose node numbers are the global dofs of the component (most rows of a block have no element at all),
    blocks merged as COO"""
    orders, terms, qname = fc.MIXED_CASES[name]
    g = fc.load(name)
    n = int(g["n"])
    qp, qw = ol.quadrature(g["dim"], qname)
    I, J, A = [], [], []  # noqa: E741
    for ov, tv, ou, tu, bt in fc.mixed_blocks(g, orders, terms):
        rp, col, val = host_assemble(g, ov, 1, tv, ou, 1, tu, bt, qp, qw, nnodes_v=n)
        I.append(np.repeat(np.arange(n), np.diff(rp))), J.append(col), A.append(val)
    I, J, A = np.concatenate(I), np.concatenate(J), np.concatenate(A)  # noqa: E741
    o = np.argsort(I.astype(np.int64) * n + J, kind="stable")
    assert np.array_equal(I[o], g["coo_i"]) and np.array_equal(J[o], g["coo_j"])
    assert np.max(np.abs(A[o] - g["coo_a"])) <= 1e-12 * np.abs(g["coo_a"]).max()


def test_row_routine_region_filter_against_the_reference():
    """the fixture of a region-restricted form: the row routine's values on the reference's (sub-)pattern, zeros elsewhere"""
    g = fc.load("rect3d_region")
    ev, eu = fc.rect_elem2node(g, "Vh", 1), fc.rect_elem2node(g, "Uh", 1)
    terms = [(0, fc.DX, 0, fc.ID, 1.0), (0, fc.ID, 0, fc.ID, 1.0)]
    qp, qw = ol.quadrature(3, "qfV5")
    rp, col, val = host_assemble(g, 1, 1, ev, 2, 1, eu, terms, qp, qw, labels=[2])
    n, m = int(g["n"]), int(g["m"])
    dense = np.zeros((n, m))
    dense[np.repeat(np.arange(n), np.diff(rp)), col] = val
    ref = np.zeros((n, m))
    ref[g["coo_i"], g["coo_j"]] = g["coo_a"]
    assert np.max(np.abs(dense - ref)) <= 1e-12 * np.abs(ref).max()
