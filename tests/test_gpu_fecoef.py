"""FE functions as data of a form, handed over as dof arrays: ffcuda_fe_table forms, on the device,
the values FreeFEM's interpreter would return at every quadrature node (pfer2R -> FElement::operator()(PHat,u,comp,op),
fflib/lgfem.cpp:2053-2088, femlib/FESpace.cpp:1078-1099, femlib/P012_3d.cpp:98-122), and the q-table entries take such a
table where it lies.  Checked against tests/fe_tables.py (numpy restatement, CPU-tested on its own) at 1e-13, and the entries
fed with a device table against the same entries fed with the same numbers from the host (bit-identical)."""
import numpy as np
import pytest

import fe_tables as ft
import ff_cases as fc
import oracle_lib as ol
from ffcuda_lib import ffcuda

pytestmark = pytest.mark.gpu

OPS3 = [fc.ID, fc.DX, fc.DY, fc.DZ]
OPS2 = [fc.ID, fc.DX, fc.DY]


@pytest.fixture(scope="module")
def ctx():
    c = ffcuda.Context(0)
    yield c
    c.close()


def _upload(ctx, g):
    return ctx.mesh_upload(g["dim"], g["xyz"], g["conn"], g["elab"], g["bconn"], g["blab"], g["belem"], g["bface"])


def _close(a, b, tol=1e-13):
    return np.max(np.abs(a - b)) <= tol * max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("dim,order", [(3, 1), (3, 2), (3, 0), (2, 1), (2, 0)])
def test_fe_table_matches_numpy(ctx, dim, order):
    g = ft.warped_mesh(dim)
    mesh = _upload(ctx, g)
    nt, nbe = g["conn"].shape[0], g["blab"].shape[0]
    e2n, nn = ft.node_table(g, order)
    rng = np.random.default_rng(7 + 10 * dim + order)
    u = rng.standard_normal(nn)
    du = ctx.vec_from(u)
    qp, qw = ffcuda.quadrature(dim, 6)
    fq, _ = ol.face_quadrature(dim)
    nq, nfq = len(qw), len(fq)
    for op in (OPS3 if dim == 3 else OPS2):
        # volume, every element; the node table is given for P2 only (P0 / P1: FreeFEM's default numbering)
        tab = ctx.vec(nt * nq)
        mesh.fe_table(order, du, tab, qp, op=op, e2n=e2n if order == 2 else None)
        want = ft.fe_values(g, order, e2n, u, qp, op)
        assert _close(tab.download().reshape(nt, nq), want), (dim, order, op)
        # boundary elements
        btab = ctx.vec(nbe * nfq)
        mesh.fe_table(order, du, btab, fq, op=op, border=True, e2n=e2n if order == 2 else None)
        bwant = ft.fe_values(g, order, e2n, u, fq, op, border=True)
        assert _close(btab.download().reshape(nbe, nfq), bwant), ("border", dim, order, op)
    # region filter, scale, offset and accumulation: table = [untouched | 2 f on region 1, 0 elsewhere] then += -0.5 dx f everywhere
    off = 5
    tab = ctx.vec(off + nt * nq)
    tab.upload(np.full(off + nt * nq, 9.0))
    mesh.fe_table(order, du, tab, qp, op=fc.ID, e2n=e2n if order == 2 else None, scale=2.0, labels=[1], offset=off)
    mesh.fe_table(order, du, tab, qp, op=fc.DX, e2n=e2n if order == 2 else None, scale=-0.5, offset=off, accumulate=True)
    want = 2.0 * ft.fe_values(g, order, e2n, u, qp, fc.ID) * (g["elab"] == 1)[:, None] - 0.5 * ft.fe_values(g, order, e2n, u, qp, fc.DX)
    got = tab.download()
    assert np.all(got[:off] == 9.0)
    assert _close(got[off:].reshape(nt, nq), want)
    # boundary labels
    btab = ctx.vec(nbe * nfq)
    mesh.fe_table(order, du, btab, fq, border=True, e2n=e2n if order == 2 else None, labels=[2, 4])
    bwant = ft.fe_values(g, order, e2n, u, fq, fc.ID, border=True) * np.isin(g["blab"], [2, 4])[:, None]
    assert _close(btab.download().reshape(nbe, nfq), bwant)


def test_fe_table_component_of_a_vector_function(ctx):
    """[P1,P1,P1] function: dof = node * 3 + c"""
    g = ft.warped_mesh(3)
    mesh = _upload(ctx, g)
    nt, nv = g["conn"].shape[0], g["xyz"].shape[0]
    u = np.random.default_rng(3).standard_normal(3 * nv)
    du = ctx.vec_from(u)
    qp, qw = ffcuda.quadrature(3, 6)
    for c in range(3):
        tab = ctx.vec(nt * len(qw))
        mesh.fe_table(1, du, tab, qp, op=fc.DZ, dstride=3, doff=c)
        assert _close(tab.download().reshape(nt, len(qw)), ft.fe_values(g, 1, None, u[c::3], qp, fc.DZ))


def test_fe_table_refuses_bad_arguments(ctx):
    g = ft.warped_mesh(3)
    mesh = _upload(ctx, g)
    nt, nv = g["conn"].shape[0], g["xyz"].shape[0]
    qp, qw = ffcuda.quadrature(3, 6)
    with pytest.raises(ffcuda.FfcudaError):   # dof vector too short for a P1 function
        mesh.fe_table(1, ctx.vec(nv - 1), ctx.vec(nt * len(qw)), qp)
    with pytest.raises(ffcuda.FfcudaError):   # table too short
        mesh.fe_table(1, ctx.vec(nv), ctx.vec(nt * len(qw) - 1), qp)
    with pytest.raises(ffcuda.FfcudaError):   # P2 without its node table
        mesh.fe_table(2, ctx.vec(10 * nv), ctx.vec(nt * len(qw)), qp)
    with pytest.raises(ffcuda.FfcudaError):   # dz on a 2-D mesh
        g2 = ft.warped_mesh(2)
        m2 = _upload(ctx, g2)
        q2, w2 = ffcuda.quadrature(2, 6)
        m2.fe_table(1, ctx.vec(g2["xyz"].shape[0]), ctx.vec(g2["conn"].shape[0] * len(w2)), q2, op=fc.DZ)


@pytest.mark.parametrize("order,ncomp", [(1, 1), (2, 1), (1, 3)])
def test_entries_take_device_tables(ctx, order, ncomp):
    """every entry that takes data at the quadrature nodes, fed with a device table formed from a dof array, against the same
    entry fed with the same numbers from the host: bit-identical; and against the oracle fed with the numpy table: 1e-12."""
    g = ft.warped_mesh(3)
    mesh = _upload(ctx, g)
    nt, nbe, nv = g["conn"].shape[0], g["blab"].shape[0], g["xyz"].shape[0]
    e2n, nn = ft.node_table(g, order)
    sp = mesh.space(order, ncomp, e2n if order == 2 else None, nn if order == 2 else 0)
    pat = sp.symbolic()
    n = pat.info()[0]
    qp, qw = ffcuda.quadrature(3, 6)
    fq, fw = ol.face_quadrature(3)
    nq, nfq = len(qw), len(fw)
    rng = np.random.default_rng(11)
    kappa = 1.5 + 0.5 * np.sin(3.0 * g["xyz"][:, 0]) * g["xyz"][:, 1]          # P1 coefficient, positive
    dk = ctx.vec_from(kappa)
    # --- bilinear: kappa (grad u . grad v + u v), and a Robin term kappa u v on two faces
    bt = []
    for c in range(ncomp):
        bt += [(c, fc.DX, c, fc.DX, 1.0), (c, fc.DY, c, fc.DY, 1.0), (c, fc.DZ, c, fc.DZ, 1.0), (c, fc.ID, c, fc.ID, 2.0)]
    rb = [(c, fc.ID, c, fc.ID, 3.0) for c in range(ncomp)]
    ctab, cbtab = ctx.vec(nt * nq), ctx.vec(nbe * nfq)
    mesh.fe_table(1, dk, ctab, qp)
    mesh.fe_table(1, dk, cbtab, fq, border=True)
    hc, hcb = ctab.download().reshape(nt, nq), cbtab.download().reshape(nbe, nfq)
    A0, A1 = pat.matrix(), pat.matrix()
    A0.assemble_qcoef(bt, qp, qw, hc)
    A0.assemble_boundary_qcoef(rb, fq, fw, hcb, [2, 5])
    A1.assemble_qcoef(bt, qp, qw, ctab)
    A1.assemble_boundary_qcoef(rb, fq, fw, cbtab, [2, 5])
    v0, v1 = A0.download(), A1.download()
    assert np.array_equal(v0, v1) and np.abs(v0).max() > 0
    oe2n = e2n if order == 2 else None
    ci, cj, ca = ol.assemble_coo_qcoef(g, order, ncomp, oe2n, bt, qp, qw, ft.fe_values(g, 1, None, kappa, qp, fc.ID))
    ci, cj, ca = ol.coo_add(n, (ci, cj, ca), ol.assemble_coo_boundary_qcoef(g, order, ncomp, oe2n, rb, fq, fw,
                                                                             ft.fe_values(g, 1, None, kappa, fq, fc.ID, border=True), [2, 5]))
    orp, ocol, oval = ol.coo_to_csr(n, ci, cj, ca)
    rp, col = pat.download()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert _close(v1, oval, 1e-12)
    # --- linear: value terms f_c v_c, derivative terms (the residual of a Newton step: grad uk . grad v), Neumann data
    f = rng.standard_normal(nv)
    uk = rng.standard_normal(nn)
    df, duk = ctx.vec_from(f), ctx.vec_from(uk)
    ftab = ctx.vec(ncomp * nt * nq)
    for c in range(ncomp):
        mesh.fe_table(1, df, ftab, qp, scale=1.0 + c, offset=c * nt * nq)
    b0, b1 = ctx.vec(n), ctx.vec(n)
    sp.assemble_linear_qvalues(b0, qp, qw, ftab.download().reshape(ncomp, nt, nq))
    sp.assemble_linear_qvalues(b1, qp, qw, ftab)
    h0, h1 = b0.download(), b1.download()
    assert np.array_equal(h0, h1) and np.abs(h0).max() > 0
    want = np.stack([(1.0 + c) * ft.fe_values(g, 1, None, f, qp, fc.ID) for c in range(ncomp)])
    assert _close(h1, ol.assemble_rhs_qvalues(g, order, ncomp, oe2n, np.zeros(n), qp, qw, want), 1e-12)
    ttab = ctx.vec(ncomp * 4 * nt * nq)   # fq[c, s, k, q]; component 0 carries grad uk (a function of the space itself), all carry f
    for c in range(ncomp):
        mesh.fe_table(1, df, ttab, qp, offset=(c * 4 + 0) * nt * nq)
    for s_, op in ((1, fc.DX), (2, fc.DY), (3, fc.DZ)):
        mesh.fe_table(order, duk, ttab, qp, op=op, e2n=e2n if order == 2 else None, offset=s_ * nt * nq)
    sp.assemble_linear_qterms(b0, qp, qw, ttab.download().reshape(ncomp, 4, nt, nq))
    sp.assemble_linear_qterms(b1, qp, qw, ttab)
    h0, h1 = b0.download(), b1.download()
    assert np.array_equal(h0, h1) and np.abs(h0).max() > 0
    want = np.zeros((ncomp, 4, nt, nq))
    want[:, 0] = ft.fe_values(g, 1, None, f, qp, fc.ID)
    for s_, op in ((1, fc.DX), (2, fc.DY), (3, fc.DZ)):
        want[0, s_] = ft.fe_values(g, order, e2n, uk, qp, op)
    assert _close(h1, ol.assemble_rhs_qterms(g, order, ncomp, oe2n, np.zeros(n), qp, qw, want), 1e-12)
    gtab = ctx.vec(ncomp * nbe * nfq)
    for c in range(ncomp):
        mesh.fe_table(1, df, gtab, fq, border=True, labels=[1, 6], scale=0.7, offset=c * nbe * nfq)
    sp.assemble_linear_boundary_qvalues(b0, fq, fw, gtab.download().reshape(ncomp, nbe, nfq), accumulate=False)
    sp.assemble_linear_boundary_qvalues(b1, fq, fw, gtab, accumulate=False)
    h0, h1 = b0.download(), b1.download()
    assert np.array_equal(h0, h1) and np.abs(h0).max() > 0
    gwant = 0.7 * ft.fe_values(g, 1, None, f, fq, fc.ID, border=True) * np.isin(g["blab"], [1, 6])[:, None]
    assert _close(h1, ol.assemble_rhs_boundary_qvalues(g, order, ncomp, oe2n, np.zeros(n), fq, fw, np.stack([gwant] * ncomp)), 1e-12)


def _dev_table(ctx, mesh, g, datum, pts, nunits, out=None, offset=0, border=False, labels=None, scale=1.0, cache=None):
    """device table of an FE datum of ff_cases.FE_CASES (the dof array goes up once per function)"""
    order, e2n, ncomp, comp, vals = fc.fe_function(g, datum)
    key = datum[0]
    if cache is not None and key in cache:
        dv = cache[key]
    else:
        dv = ctx.vec_from(vals)
        if cache is not None:
            cache[key] = dv
    dim = g["dim"]
    nq = np.asarray(pts).size // (dim - 1 if border else dim)
    tab = out if out is not None else ctx.vec(nunits * nq)
    # P0 / P1 scalar functions are numbered like elements / vertices by FreeFEM: the table is only needed otherwise
    default = order == 0 or (order == 1 and np.array_equal(e2n, g["conn"]))
    mesh.fe_table(order, dv, tab, pts, op=datum[2], border=border, e2n=None if default else e2n, dstride=ncomp, doff=comp, scale=scale,
                  labels=labels, offset=offset, accumulate=out is not None)
    return tab


@pytest.mark.parametrize("name", sorted(fc.FE_CASES))
def test_fe_cases_match_reference(ctx, name):
    """forms whose data are FE functions, assembled on the device from the dof arrays, against what FreeFEM assembled
    (fixtures dumped from the unmodified reference): pattern bit-exact, values / rhs / solution 1e-12"""
    C = fc.FE_CASES[name]
    order, ncomp, bt, lt, qname, bcs = C["base"]
    g = fc.load(name)
    dim, n = g["dim"], g["ndof"]
    e2n = fc.elem2node(g, order, ncomp)
    mesh = _upload(ctx, g)
    sp = mesh.space(order, ncomp, e2n, int(e2n.max()) + 1)
    pat = sp.symbolic()
    assert pat.info()[0] == n
    qp, qw = ol.quadrature(dim, qname)
    fq, fw = ol.face_quadrature(dim)
    nt, nbe, nq, nfq = g["conn"].shape[0], g["blab"].shape[0], len(qw), len(fw)
    cache = {}
    A = pat.matrix()
    A.assemble(bt, qp, qw)
    for datum, terms in C["qcoef"]:
        A.assemble_qcoef(terms, qp, qw, _dev_table(ctx, mesh, g, datum, qp, nt, cache=cache), accumulate=True)
    for labels, datum, terms in C["bbil"]:
        A.assemble_boundary_qcoef(terms, fq, fw, _dev_table(ctx, mesh, g, datum, fq, nbe, border=True, cache=cache), labels, accumulate=True)
    b = ctx.vec(n)
    sp.assemble_linear(b, lt, qp, qw)
    ftab = ctx.vec(ncomp * (dim + 1) * nt * nq)
    for vcomp, vop, datum, scale in C["lin"]:
        _dev_table(ctx, mesh, g, datum, qp, nt, out=ftab, offset=(vcomp * (dim + 1) + ft._SLOT[vop]) * nt * nq, scale=scale, cache=cache)
    sp.assemble_linear_qterms(b, qp, qw, ftab, accumulate=True)
    if C["blin"]:
        gtab = ctx.vec(ncomp * nbe * nfq)
        for labels, vcomp, datum, scale in C["blin"]:
            _dev_table(ctx, mesh, g, datum, fq, nbe, out=gtab, offset=vcomp * nbe * nfq, border=True, labels=labels, scale=scale, cache=cache)
        sp.assemble_linear_boundary_qvalues(b, fq, fw, gtab, accumulate=True)
    for labels, mask, values in bcs:
        bc = sp.bc_from_labels(labels, mask, values)
        A.apply_bc(bc, 1e30)
        b.apply_bc(bc, 1e30)
    rp, col = pat.download()
    grp, gcol, gval = fc.golden_csr(g)
    assert np.array_equal(rp, grp) and np.array_equal(col, gcol)
    val, hb = A.download(), b.download()
    reg = np.abs(gval) < 1e29
    assert np.array_equal(np.abs(val) < 1e29, reg)
    assert np.max(np.abs(val - gval)[reg]) <= 1e-12 * np.abs(gval[reg]).max()
    big = np.abs(g["b"]) > 1e20
    assert np.array_equal(np.abs(hb) > 1e20, big)
    assert np.max(np.abs(hb - g["b"])[~big]) <= 1e-12 * np.abs(g["b"][~big]).max()
    x = ctx.vec(n)
    it, conv, _ = A.cg(b, x, eps=1e-14, itmax=0, tgv=1e30)
    assert conv == 1 and abs(it - int(g["cg_iters14"])) <= 3
    assert np.max(np.abs(x.download() - g["u14"])) <= 1e-12 * np.abs(g["u14"]).max()
