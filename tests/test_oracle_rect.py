"""CPU tests: the oracle's restatement of the rectangular assembly (`matrix B = vb(Uh,Vh)`, Element_Op with Ku != Kv,
fflib/problem.cpp:6337-6437) pinned on fixtures dumped from the unmodified reference (tests/golden/make_golden_rect.py):
HashMatrix insertion order and sorted pattern bit-exact, values 1e-12 of the largest entry."""
import numpy as np
import pytest

import ff_cases as fc
import oracle_lib as ol

RTOL = 1e-12


def oracle_rect(name):
    (ov, cv), (ou, cu), terms, qname = fc.RECT_CASES[name]
    g = fc.load(name)
    mesh = {k: g[k] for k in ("dim", "xyz", "conn", "elab")}
    ev, eu = fc.rect_elem2node(g, "Vh", cv), fc.rect_elem2node(g, "Uh", cu)
    qp, qw = ol.quadrature(g["dim"], qname)
    return g, ol.assemble_coo_rect(mesh, ov, cv, ev, ou, cu, eu, terms, qp, qw)


@pytest.mark.parametrize("name", sorted(fc.RECT_CASES))
def test_rectangular_matrix_against_the_reference(name):
    g, (ci, cj, ca) = oracle_rect(name)
    n, m = int(g["n"]), int(g["m"])
    assert ci.max() < n and cj.max() < m
    assert np.array_equal(ci, g["ins_i"]) and np.array_equal(cj, g["ins_j"])
    o = np.argsort(ci.astype(np.int64) * m + cj, kind="stable")
    assert np.array_equal(ci[o], g["coo_i"]) and np.array_equal(cj[o], g["coo_j"])
    assert np.max(np.abs(ca[o] - g["coo_a"])) <= RTOL * np.abs(g["coo_a"]).max()


def test_rectangular_with_equal_spaces_is_the_square_assembly():
    """Uh = Vh: the rectangular restatement gives what ffo_assemble_coo gives, bit for bit"""
    g = fc.load("lame3d_p2_cube2")
    order, ncomp, bt, _, qname, _ = fc.CASES["lame3d_p2_cube2"]
    mesh = {k: g[k] for k in ("dim", "xyz", "conn", "elab")}
    e2n = fc.elem2node(g, order, ncomp)
    qp, qw = ol.quadrature(3, qname)
    a = ol.assemble_coo(mesh, order, ncomp, e2n, bt, qp, qw)
    b = ol.assemble_coo_rect(mesh, order, ncomp, e2n, order, ncomp, e2n, bt, qp, qw)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name", sorted(fc.MIXED_CASES))
def test_mixed_order_space_as_scalar_blocks(name):
    """[P2,P2,P1] / [P2,P2,P2,P1] in one fespace: the sum of the scalar blocks, each assembled with the global dofs of its
    components as node numbers, is the matrix the reference assembles (sorted pattern bit-exact, values 1e-12)"""
    orders, terms, qname = fc.MIXED_CASES[name]
    g = fc.load(name)
    n = int(g["n"])
    mesh = {k: g[k] for k in ("dim", "xyz", "conn", "elab")}
    qp, qw = ol.quadrature(g["dim"], qname)
    I, J, A = [], [], []  # noqa: E741
    for ov, tv, ou, tu, bt in fc.mixed_blocks(g, orders, terms):
        ci, cj, ca = ol.assemble_coo_rect(mesh, ov, 1, tv, ou, 1, tu, bt, qp, qw)
        I.append(ci), J.append(cj), A.append(ca)
    I, J, A = np.concatenate(I), np.concatenate(J), np.concatenate(A)  # noqa: E741
    key = I.astype(np.int64) * n + J
    assert len(np.unique(key)) == len(key)  # the blocks are disjoint
    o = np.argsort(key, kind="stable")
    assert np.array_equal(I[o], g["coo_i"]) and np.array_equal(J[o], g["coo_j"])
    assert np.max(np.abs(A[o] - g["coo_a"])) <= RTOL * np.abs(g["coo_a"]).max()


def test_region_restricted_rectangular_form_has_the_sub_pattern():
    """int3d(Th,2)(...) on a mesh with two regions: HashMatrix only holds the couples of the visited elements - the oracle
    follows; the device entry documents that ITS pattern keeps every couple (the plugin therefore leaves such forms to
    FreeFEM unless the regions cover the mesh), its values on the sub-pattern are what the row routine's filter gives"""
    g = fc.load("rect3d_region")
    mesh = {k: g[k] for k in ("dim", "xyz", "conn", "elab")}
    ev, eu = fc.rect_elem2node(g, "Vh", 1), fc.rect_elem2node(g, "Uh", 1)
    terms = [(0, fc.DX, 0, fc.ID, 1.0), (0, fc.ID, 0, fc.ID, 1.0)]
    qp, qw = ol.quadrature(3, "qfV5")
    ci, cj, ca = ol.assemble_coo_rect(mesh, 1, 1, ev, 2, 1, eu, terms, qp, qw, labels=[2])
    assert np.array_equal(ci, g["ins_i"]) and np.array_equal(cj, g["ins_j"])
    m = int(g["m"])
    o = np.argsort(ci.astype(np.int64) * m + cj, kind="stable")
    assert np.array_equal(ci[o], g["coo_i"]) and np.array_equal(cj[o], g["coo_j"])
    assert np.max(np.abs(ca[o] - g["coo_a"])) <= RTOL * np.abs(g["coo_a"]).max()
    fi, fj, _ = ol.assemble_coo_rect(mesh, 1, 1, ev, 2, 1, eu, terms, qp, qw)
    assert len(fi) > len(ci)  # the full pattern is strictly larger
