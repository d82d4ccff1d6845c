"""CPU tests of the checker of ffcuda_fe_table (tests/fe_tables.py) and of the FE-data fixtures: (1) the numpy restatement
reproduces polynomials of the space's degree and their derivatives at volume and face quadrature nodes; (2) the oracle fed
with tables formed by it reproduces the matrices and right-hand sides FreeFEM assembled from forms whose data are FE
functions (tests/golden/fe*_data.npz, dumped from the unmodified reference): pattern bit-exact, values 1e-12."""
import numpy as np
import pytest

import fe_tables as ft
import ff_cases as fc
import oracle_lib as ol

RTOL = 1e-12


def _mesh(g):
    return {k: g[k] for k in ("dim", "xyz", "conn", "elab", "bconn", "blab", "belem", "bface")}


@pytest.mark.parametrize("dim,order", [(3, 0), (3, 1), (3, 2), (2, 0), (2, 1)])
def test_restatement_reproduces_polynomials(dim, order):
    g = ft.warped_mesh(dim)
    e2n, nn = ft.node_table(g, order)
    X = g["xyz"]
    co = np.array([0.3, -1.2, 0.7, 2.0][:dim])

    def poly(P, d=None):   # degree `order`; d: derivative axis
        lin = P @ co + 0.5
        if order == 0:
            return 0 * lin + (1.0 if d is None else 0.0)
        if order == 1:
            return lin if d is None else 0 * lin + co[d]
        q = lin * lin + P[..., 0] * P[..., -1]
        if d is None:
            return q
        return 2 * lin * co[d] + (P[..., -1] if d == 0 else 0) + (P[..., 0] if d == dim - 1 else 0)

    if order == 0:
        u = np.ones(nn)
    elif order == 1:
        u = poly(X)
    else:  # P2 nodes: vertices, then edge midpoints
        u = np.zeros(nn)
        u[e2n[:, :4]] = poly(X[g["conn"]])
        for e, (i, j) in enumerate(ft._EDGE3):
            u[e2n[:, 4 + e]] = poly(0.5 * (X[g["conn"][:, i]] + X[g["conn"][:, j]]))
    qp, _ = ol.quadrature(dim, "qfV5" if dim == 3 else "qf5pT")
    fq, _ = ol.face_quadrature(dim)
    ops = [fc.ID, fc.DX, fc.DY] + ([fc.DZ] if dim == 3 else [])
    for border, pts, P in ((False, qp, ol.quad_points_xyz(g, qp)), (True, fq, ol.bquad_points_xyz(g, fq))):
        for k, op in enumerate(ops):
            got = ft.fe_values(g, order, e2n, u, pts, op, border=border)
            want = poly(P, None if k == 0 else k - 1)
            assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.abs(want).max()), (border, op)


def _table(g, datum, pts, border=False):
    order, e2n, ncomp, comp, vals = fc.fe_function(g, datum)
    return ft.fe_values(g, order, e2n, vals[comp::ncomp], pts, datum[2], border=border)


@pytest.mark.parametrize("name", sorted(fc.FE_CASES))
def test_oracle_with_fe_tables_matches_reference(name):
    C = fc.FE_CASES[name]
    order, ncomp, bt, lt, qname, bcs = C["base"]
    g = fc.load(name)
    m = _mesh(g)
    dim, n = g["dim"], g["ndof"]
    e2n = fc.elem2node(g, order, ncomp)
    qp, qw = ol.quadrature(dim, qname)
    fq, fw = ol.face_quadrature(dim)
    nt, nbe, nq, nfq = g["conn"].shape[0], g["blab"].shape[0], len(qw), len(fw)
    ci, cj, ca = ol.assemble_coo(m, order, ncomp, e2n, bt, qp, qw)
    for datum, terms in C["qcoef"]:
        ci, cj, ca = ol.coo_add(n, (ci, cj, ca), ol.assemble_coo_qcoef(m, order, ncomp, e2n, terms, qp, qw, _table(g, datum, qp)))
    for labels, datum, terms in C["bbil"]:
        ci, cj, ca = ol.coo_add(n, (ci, cj, ca),
                                ol.assemble_coo_boundary_qcoef(m, order, ncomp, e2n, terms, fq, fw, _table(g, datum, fq, True), labels))
    o = np.argsort(ci.astype(np.int64) * n + cj, kind="stable")
    ci, cj, ca = ci[o], cj[o], ca[o]
    assert np.array_equal(ci, g["coo_i"]) and np.array_equal(cj, g["coo_j"])
    dofs, vals = [], []
    for labels, mask, values in bcs:
        d, v = ol.bc_pairs(m, order, ncomp, e2n, labels, mask, values)
        dofs.append(d)
        vals.append(v)
    dofs, vals = np.concatenate(dofs), np.concatenate(vals)
    ca = ol.bc_matrix_coo(ci, cj, ca, n, dofs, 1e30)
    reg = np.abs(g["coo_a"]) < 1e29
    assert np.array_equal(np.abs(ca) < 1e29, reg)
    assert np.max(np.abs(ca - g["coo_a"])[reg]) <= RTOL * np.abs(g["coo_a"][reg]).max()
    # right-hand side
    fqt = np.zeros((ncomp, dim + 1, nt, nq))
    for vcomp, vop, datum, scale in C["lin"]:
        fqt[vcomp, ft._SLOT[vop]] += scale * _table(g, datum, qp)
    b = ol.assemble_rhs(m, order, ncomp, e2n, n, lt, qp, qw)
    b = ol.assemble_rhs_qterms(m, order, ncomp, e2n, b, qp, qw, fqt)
    if C["blin"]:
        gq = np.zeros((ncomp, nbe, nfq))
        for labels, vcomp, datum, scale in C["blin"]:
            gq[vcomp] += scale * _table(g, datum, fq, True) * np.isin(g["blab"], labels)[:, None]
        b = ol.assemble_rhs_boundary_qvalues(m, order, ncomp, e2n, b, fq, fw, gq)
    b = ol.bc_rhs(b, dofs, vals, 1e30)
    big = np.abs(g["b"]) > 1e20
    assert np.array_equal(np.abs(b) > 1e20, big)
    assert np.max(np.abs(b - g["b"])[~big]) <= RTOL * np.abs(g["b"][~big]).max()
    # the solve of the fixture, converged to round-off
    x, it, ret, _ = ol.cg(n, ci, cj, ca, b, np.zeros(n), eps=1e-14, itmax=0, tgv=1e30)
    assert ret in (1, 2) and abs(it - int(g["cg_iters14"])) <= 3
    assert np.max(np.abs(x - g["u14"])) <= RTOL * np.abs(g["u14"]).max()
