"""The oracle against the reference RUN HERE, at sizes well above the committed fixtures (CPU; skipped where oracle/_ref is not
built): the unmodified FreeFem++ assembles a problem that combines every row of the path at once - coefficients depending on
the mesh point in the bilinear form, a Robin term alpha(x) u v, Neumann data g(x) v, a right-hand side with f(x) v and
derivatives of the test function, Dirichlet rows - and the C restatement must give the same pattern bit for bit and the same
values / right-hand side to 1e-12."""
import os
import sys

import numpy as np
import pytest

import ff_cases as fc
import oracle_lib as ol

sys.path.insert(0, fc.GOLDEN_DIR)
import make_golden as mg  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(mg.FF), reason="oracle/_ref/FreeFem++-nw not built (needs /root/reference at build time)")
RTOL = 1e-12


def _mesh(g):
    return {k: g[k] for k in ("dim", "xyz", "conn", "elab", "bconn", "blab", "belem", "bface")}


LIVE = {
    "p1_3d": dict(
        case=dict(dim=3, mesh="cube(7,6,8,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P1",
                  bil="(1+x*y+z*z)*(" + mg.LAP3 + ")+(2+sin(x))*u*v+0.5*dx(u)*v", lin="(x*y+sin(z))*v+(1+z)*dz(v)-y*dx(v)+2.*v",
                  blin="+int2d(Th,2,3)((1+x*z)*u*v)+int2d(Th,2,3)((y+sin(z))*v)+int2d(Th,6)(0.5*v)+int2d(Th,5)(3.*u*v)", bc="on(1,u=0.5)",
                  solve=False),
        order=1, ncomp=1, qname="qfV5", bcs=[([1], 1, [0.5])],
        const_b=[(0, fc.DX, 0, fc.ID, 0.5)], const_l=[(0, fc.ID, 2.0)],
        qcoef=[(lambda P: 1 + P[..., 0] * P[..., 1] + P[..., 2] ** 2, fc.LAP3), (lambda P: 2 + np.sin(P[..., 0]), [(0, fc.ID, 0, fc.ID, 1.0)])],
        fqt=lambda P: np.stack([np.stack([P[..., 0] * P[..., 1] + np.sin(P[..., 2]), -P[..., 1], 0 * P[..., 0], 1 + P[..., 2]])]),
        brobin_q=([2, 3], lambda P: 1 + P[..., 0] * P[..., 2], [(0, fc.ID, 0, fc.ID, 1.0)]), brobin_c=([5], [(0, fc.ID, 0, fc.ID, 3.0)]),
        bneu_q=([2, 3], lambda P: (P[..., 1] + np.sin(P[..., 2]))[None]), bneu_c=([6], [(0, fc.ID, 0.5)])),
    "p2_3d": dict(
        case=dict(dim=3, mesh="cube(4,3,4,[x+0.1*y*y,y+0.05*z,z*(1+0.2*x)])", fe="P2",
                  bil="(1+x*y+z*z)*(" + mg.LAP3 + ")+2.*u*v", lin="(1+z)*dz(v)+x*v-y*z*dx(v)",
                  blin="+int2d(Th,2,3)((1+x*z)*u*v)+int2d(Th,6)((y+sin(z))*v)", bc="on(1,u=0)", solve=False),
        order=2, ncomp=1, qname="qfV5", bcs=[([1], 1, [0.0])],
        const_b=[(0, fc.ID, 0, fc.ID, 2.0)], const_l=[],
        qcoef=[(lambda P: 1 + P[..., 0] * P[..., 1] + P[..., 2] ** 2, fc.LAP3)],
        fqt=lambda P: np.stack([np.stack([P[..., 0], -P[..., 1] * P[..., 2], 0 * P[..., 0], 1 + P[..., 2]])]),
        brobin_q=([2, 3], lambda P: 1 + P[..., 0] * P[..., 2], [(0, fc.ID, 0, fc.ID, 1.0)]), brobin_c=None,
        bneu_q=([6], lambda P: (P[..., 1] + np.sin(P[..., 2]))[None]), bneu_c=([1], [])),
    "lame_p1_3d": dict(
        case=dict(dim=3, mesh="cube(5,6,5)", fe="[P1,P1,P1]", unk="[u1,u2,u3]", tst="[v1,v2,v3]", pre=mg.LAME_PRE,
                  bil="(1+x)*(" + mg.LAME + ")", lin="x*dx(v1)+y*v2-0.05*(1+x)*dz(v3)+z*dy(v1)",
                  blin="+int2d(Th,3)(1e3*(1+x)*(u1*v1+u2*v2+u3*v3))+int2d(Th,2)(0.3*z*v1-0.2*(1+y)*v3)", bc="on(1,u1=0,u2=0,u3=0)",
                  solve=False),
        order=1, ncomp=3, qname="qfV5", bcs=[([1], 7, [0.0, 0.0, 0.0])],
        const_b=[], const_l=[],
        qcoef=[(lambda P: 1 + P[..., 0], fc.lame_terms())],
        fqt=lambda P: np.stack([np.stack([0 * P[..., 0], P[..., 0], P[..., 2], 0 * P[..., 0]]),
                                np.stack([P[..., 1], 0 * P[..., 0], 0 * P[..., 0], 0 * P[..., 0]]),
                                np.stack([0 * P[..., 0], 0 * P[..., 0], 0 * P[..., 0], -0.05 * (1 + P[..., 0])])]),
        brobin_q=([3], lambda P: 1 + P[..., 0], [(c, fc.ID, c, fc.ID, 1e3) for c in range(3)]), brobin_c=None,
        bneu_q=([2], lambda P: np.stack([0.3 * P[..., 2], 0 * P[..., 0], -0.2 * (1 + P[..., 1])])), bneu_c=([1], [])),
    "p2_2d": dict(
        case=dict(dim=2, mesh="square(14,12,[x+0.2*y*y,y*(1+0.3*x)])", fe="P2",
                  bil="(1+x*y)*(" + mg.LAP2 + ")+(1+sin(x)*y)*u*v", lin="exp(x)*y*v+sin(x)*dx(v)-y*dy(v)",
                  blin="+int1d(Th,2,3)((1+x*y)*u*v)+int1d(Th,2)(exp(y)*v)+int1d(Th,1)(0.25*v)", bc="on(4,u=0)", solve=False),
        order=2, ncomp=1, qname="qf5pT", bcs=[([4], 1, [0.0])],
        const_b=[], const_l=[],
        qcoef=[(lambda P: 1 + P[..., 0] * P[..., 1], fc.LAP2), (lambda P: 1 + np.sin(P[..., 0]) * P[..., 1], [(0, fc.ID, 0, fc.ID, 1.0)])],
        fqt=lambda P: np.stack([np.stack([np.exp(P[..., 0]) * P[..., 1], np.sin(P[..., 0]), -P[..., 1]])]),
        brobin_q=([2, 3], lambda P: 1 + P[..., 0] * P[..., 1], [(0, fc.ID, 0, fc.ID, 1.0)]), brobin_c=None,
        bneu_q=([2], lambda P: np.exp(P[..., 1])[None]), bneu_c=([1], [(0, fc.ID, 0.25)])),
}


@needs_ref
@pytest.mark.parametrize("name", sorted(LIVE))
def test_oracle_matches_the_reference_run_here(name):
    L = LIVE[name]
    g = mg.run_case(name, case=L["case"], save=False)
    g = {k: (v.item() if isinstance(v, np.ndarray) and v.ndim == 0 and k in ("dim", "ndof") else v) for k, v in g.items()}
    dim, n, order, ncomp = int(g["dim"]), int(g["ndof"]), L["order"], L["ncomp"]
    assert n > 400
    e2n = fc.elem2node(g, order, ncomp)
    m = _mesh(g)
    m["dim"] = dim
    qp, qw = ol.quadrature(dim, L["qname"])
    fq, fw = ol.face_quadrature(dim)
    # matrix: constant terms, coefficient groups, Robin terms
    coo = ol.assemble_coo(m, order, ncomp, e2n, L["const_b"], qp, qw)
    for cfun, terms in L["qcoef"]:
        coo = ol.coo_add(n, coo, ol.assemble_coo_qcoef(m, order, ncomp, e2n, terms, qp, qw, cfun(ol.quad_points_xyz(m, qp))))
    labs, cfun, terms = L["brobin_q"]
    coo = ol.coo_add(n, coo, ol.assemble_coo_boundary_qcoef(m, order, ncomp, e2n, terms, fq, fw, cfun(ol.bquad_points_xyz(m, fq)), labs))
    if L["brobin_c"]:
        coo = ol.coo_add(n, coo, ol.assemble_coo_boundary(m, order, ncomp, e2n, L["brobin_c"][1], fq, fw, L["brobin_c"][0]))
    ci, cj, ca = coo
    o = np.argsort(ci.astype(np.int64) * n + cj, kind="stable")
    ci, cj, ca = ci[o], cj[o], ca[o]
    assert np.array_equal(ci, g["coo_i"]) and np.array_equal(cj, g["coo_j"])                      # bit-exact pattern
    dofs, vals = [], []
    for labels, mask, values in L["bcs"]:
        d, v = ol.bc_pairs(m, order, ncomp, e2n, labels, mask, values)
        dofs.append(d)
        vals.append(v)
    dofs, vals = np.concatenate(dofs), np.concatenate(vals)
    ca = ol.bc_matrix_coo(ci, cj, ca, n, dofs, 1e30)
    big = np.abs(g["coo_a"]) > 1e29
    assert np.array_equal(np.abs(ca) > 1e29, big)
    assert np.max(np.abs(ca - g["coo_a"])[~big]) <= RTOL * np.abs(g["coo_a"][~big]).max()
    # right-hand side: constant terms, data at the nodes (value and derivative terms), Neumann data
    b = ol.assemble_rhs(m, order, ncomp, e2n, n, L["const_l"], qp, qw)
    b = ol.assemble_rhs_qterms(m, order, ncomp, e2n, b, qp, qw, L["fqt"](ol.quad_points_xyz(m, qp)))
    labs, gfun = L["bneu_q"]
    gq = gfun(ol.bquad_points_xyz(m, fq)) * np.isin(g["blab"], labs)[None, :, None]
    b = ol.assemble_rhs_boundary_qvalues(m, order, ncomp, e2n, b, fq, fw, gq)
    b = ol.assemble_rhs_boundary(m, order, ncomp, e2n, b, L["bneu_c"][1], fq, fw, L["bneu_c"][0])
    b = ol.bc_rhs(b, dofs, vals, 1e30)
    bbig = np.abs(g["b"]) > 1e20
    assert np.array_equal(np.abs(b) > 1e20, bbig)
    assert np.allclose(b[bbig], g["b"][bbig], rtol=1e-15, atol=0)
    assert np.max(np.abs(b - g["b"])[~bbig]) <= RTOL * np.abs(g["b"][~bbig]).max()
