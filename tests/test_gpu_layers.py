"""GPU parity tests of the device `buildlayers` (ffcuda_mesh_buildlayers, csrc/mesh.cu) — SURVEY.md §8 f-4:
  (1) against the fixtures dumped from the unmodified FreeFEM (tests/golden/layers_*.npz, fflib/msh3.cpp:895-1757),
  (2) against the CPU oracle at larger sizes with degenerate columns and label maps,
  (3) properties at the Heat3d.idp size (volume, conformity through the device adjacency),
  (4) the assembly + CG path on a layered mesh against the oracle.
Bar: every array bit-exact (vertex coordinates included: same operations in the same order)."""
import numpy as np
import pytest

import ff_cases as fc
import oracle_lib as ol
from ffcuda_lib import ffcuda
from test_oracle_golden import LAYER_CASES, layers_inputs

pytestmark = pytest.mark.gpu

KEYS = ("xyz", "conn", "elab", "bconn", "blab", "belem", "bface")


@pytest.fixture(scope="module")
def ctx():
    c = ffcuda.Context(0)
    yield c
    c.close()


def _upload2(ctx, m2):
    return ctx.mesh_upload(2, m2["xyz"], m2["conn"], m2["elab"], m2["bconn"], m2["blab"], m2["belem"], m2["bface"])


@pytest.mark.parametrize("name", LAYER_CASES)
def test_buildlayers_vs_reference_fixture(ctx, name):
    g = fc.load(name)
    m2, kw = layers_inputs(g)
    m = _upload2(ctx, m2).buildlayers(int(g["nlayer"]), **kw).download()
    for k in KEYS:
        assert m[k].shape == g[k].shape, k
        assert np.array_equal(m[k], g[k]), k


def _columns(xy, nlayer, seed):
    """per-vertex data with everything the generator has to cope with: columns of 0..nlayer layers (no triangle with three
    empty columns), curved bottom and top"""
    rng = np.random.default_rng(seed)
    x, y = xy[:, 0], xy[:, 1]
    zmin = 0.1 * np.sin(3 * x) * y
    zmax = 1.0 + 0.3 * np.cos(2 * y) + 0.2 * x
    ni = rng.integers(1, nlayer + 1, xy.shape[0]).astype(np.int32)
    ni[rng.random(xy.shape[0]) < 0.08] = 0
    return ni, zmin, zmax


@pytest.mark.parametrize("nx,ny,nlayer,seed", [(40, 30, 17, 1), (9, 14, 64, 2), (25, 25, 1, 3)])
def test_buildlayers_vs_oracle_degenerate_columns(ctx, nx, ny, nlayer, seed):
    m2 = ol.square(nx, ny)
    ni, zmin, zmax = _columns(m2["xyz"], nlayer, seed)
    # a triangle whose three columns are empty would have no element: FreeFEM stops there (msh3.cpp:4662-4673)
    tri = m2["conn"]
    empty = (ni[tri] == 0).all(axis=1)
    ni[tri[empty, 0]] = 1
    maps = dict(regmap=[0, 5], midmap=[1, 10, 2, 20, 3, 30, 1, 11], upmap=[0, 7], downmap=[0, 8, 3, 9])
    ref = ol.buildlayers(m2, nlayer, ni, zmin, zmax, **maps)
    for base in (_upload2(ctx, m2), ctx.mesh_square(nx, ny)):  # uploaded and device-generated 2-D mesh
        m = base.buildlayers(nlayer, ni, zmin, zmax, **maps).download()
        for k in KEYS:
            assert m[k].shape == ref[k].shape, k
            assert np.array_equal(m[k], ref[k]), k


def test_buildlayers_heat3d_size_properties(ctx):
    """the mesh of idp/Heat3d.idp:14-16 at nn = 64: counts, positive volumes summing to 1, conforming (every interior face
    has exactly one mate, the unmatched faces are the boundary triangles), labels as the script maps them"""
    nn = 64
    mesh = ctx.mesh_square(nn, nn).buildlayers(nn, midmap=[1, 1, 2, 1, 3, 1, 4, 1], upmap=[0, 1], downmap=[0, 1])
    dim, nv, nt, nbe = mesh.info()
    assert (dim, nv, nt, nbe) == (3, (nn + 1) ** 3, 6 * nn ** 3, 12 * nn * nn)
    m = mesh.download()
    p = m["xyz"][m["conn"]]
    vol = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0]) / 6.0
    assert vol.min() > 0
    assert abs(vol.sum() - 1.0) < 1e-12
    assert set(np.unique(m["blab"])) == {1}
    adj = mesh.adjacency()
    assert int((adj == -1).sum()) == nbe and int((adj == -2).sum()) == 0
    # the boundary links point at boundary faces holding the same vertices
    fv = np.array([[3, 2, 1], [0, 2, 3], [3, 1, 0], [0, 1, 2]])
    faces = m["conn"][m["belem"][:, None], fv[m["bface"]]]
    assert np.array_equal(np.sort(faces, axis=1), np.sort(m["bconn"], axis=1))
    assert (adj[4 * m["belem"] + m["bface"]] == -1).all()


def test_heat_form_on_layered_mesh_vs_oracle(ctx):
    """config 4 (the heat step of idp/Heat3d.idp on its buildlayers mesh, uh*vh + dt grad.grad, on(1,uh=0)):
    matrix, right-hand side and CG on the device-built mesh against the oracle on the oracle's mesh — pattern bit-exact,
    values 1e-12, same iteration count.  Two assemblies: the second one runs on the row tiles / fans of the space."""
    nn, dt, tgv = 10, 0.01, 1e30
    maps = dict(midmap=[1, 1, 2, 1, 3, 1, 4, 1], upmap=[0, 1], downmap=[0, 1])
    m = ol.buildlayers(ol.square(nn, nn), nn, **maps)
    mesh = ctx.mesh_square(nn, nn).buildlayers(nn, **maps)
    bt = [(0, fc.ID, 0, fc.ID, 1.0), (0, fc.DX, 0, fc.DX, dt), (0, fc.DY, 0, fc.DY, dt), (0, fc.DZ, 0, fc.DZ, dt)]
    lt = [(0, fc.ID, dt)]
    bcs = [([1], 1, [0.0])]
    qp, qw = ffcuda.quadrature(3, 6)
    sp = mesh.space(1, 1)
    n = sp.info()[0]
    assert n == (nn + 1) ** 3
    ci, cj, ca = ol.assemble_coo(m, 1, 1, None, bt, qp, qw)
    d, v = ol.bc_pairs(m, 1, 1, None, *bcs[0])
    ca = ol.bc_matrix_coo(ci, cj, ca, n, d, tgv)
    ob = ol.bc_rhs(ol.assemble_rhs(m, 1, 1, None, n, lt, qp, qw), d, v, tgv)
    orp, ocol, oval = ol.coo_to_csr(n, ci, cj, ca)
    pat = sp.symbolic()
    rp, col = pat.download()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    A = pat.matrix()
    for _ in range(2):
        A.assemble(bt, qp, qw)
        b = ctx.vec(n)
        sp.assemble_linear(b, lt, qp, qw)
        bc = sp.bc_from_labels(*bcs[0])
        A.apply_bc(bc, tgv)
        b.apply_bc(bc, tgv)
        val = A.download()
        big = np.abs(oval) > 1e29
        assert np.array_equal(np.abs(val) > 1e29, big)
        assert np.max(np.abs(val - oval)[~big]) <= 1e-12 * np.abs(oval[~big]).max()
        hb = b.download()
        bbig = np.abs(ob) > 1e20
        assert np.array_equal(np.abs(hb) > 1e20, bbig)
        assert np.max(np.abs(hb - ob)[~bbig]) <= 1e-12 * np.abs(ob[~bbig]).max()
    x = ctx.vec(n)
    it, conv, _ = A.cg(b, x, eps=1e-6, itmax=0, tgv=tgv)
    ox, oit, _, _ = ol.cg(n, ci, cj, ca, ob, np.zeros(n), eps=1e-6, itmax=0, tgv=tgv)
    assert conv == 1 and it == oit
    assert np.max(np.abs(x.download() - ox)) <= 1e-12 * np.abs(ox).max()
