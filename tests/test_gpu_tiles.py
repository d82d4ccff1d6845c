"""Row-tile assembly (csrc/tiles.cu) of scalar P1 forms against the oracle and against the thread-per-row kernel:
same CSR pattern (it is the symbolic phase's), values within 1e-12, bit-identical from run to run, for every tile
size, on device-generated meshes, on the reference's own (warped) golden meshes and on a numbering without locality."""
import numpy as np
import pytest

import ff_cases as fc
import oracle_lib as ol
from ffcuda_lib import ffcuda

pytestmark = pytest.mark.gpu
RTOL = 1e-12
TGV = 1e30


@pytest.fixture()
def ctx():
    c = ffcuda.Context(0)
    yield c
    c.close()


def _oracle_vals(m, n, terms, qp, qw):
    ci, cj, ca = ol.assemble_coo(m, 1, 1, None, terms, qp, qw)
    return ol.coo_to_csr(n, ci, cj, ca)


def _assemble(ctx, mesh, terms, qp, qw, policy, rows=64, accumulate_terms=None):
    ctx.set_option("tile_policy", policy)
    ctx.set_option("tile_rows", rows)
    sp = mesh.space(1, 1)
    pat = sp.symbolic()
    A = pat.matrix()
    A.assemble(terms, qp, qw)
    if accumulate_terms:
        A.assemble(accumulate_terms, qp, qw, accumulate=True)
    return sp, pat, A


HEAT3 = [(0, fc.ID, 0, fc.ID, 100.0)] + fc.LAP3
HEAT2 = [(0, fc.ID, 0, fc.ID, 7.5)] + fc.LAP2
MESHES = [("cube", (11, 9, 13), fc.LAP3), ("cube", (8, 8, 8), HEAT3), ("cube", (1, 1, 1), fc.LAP3), ("cube", (2, 3, 1), HEAT3),
          ("square", (37, 29), fc.LAP2), ("square", (16, 16), HEAT2), ("square", (1, 1), fc.LAP2)]


@pytest.mark.parametrize("rows", [8, 32, 64, 256])
@pytest.mark.parametrize("case", MESHES, ids=[f"{k}{'x'.join(map(str, s))}_{len(t)}t" for k, s, t in MESHES])
def test_tiles_against_oracle(ctx, case, rows):
    kind, size, terms = case
    dim = 3 if kind == "cube" else 2
    m = ol.cube(*size) if kind == "cube" else ol.square(*size)
    mesh = ctx.mesh_cube(*size) if kind == "cube" else ctx.mesh_square(*size)
    qp, qw = ffcuda.quadrature(dim, 6)
    n = m["xyz"].shape[0]
    orp, ocol, oval = _oracle_vals(m, n, terms, qp, qw)
    sp, pat, A = _assemble(ctx, mesh, terms, qp, qw, 2, rows)
    rp, col = pat.download()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    val = A.download()
    assert np.max(np.abs(val - oval)) <= RTOL * np.abs(oval).max()
    # the other kernel, same inputs
    sp0, pat0, A0 = _assemble(ctx, mesh, terms, qp, qw, 0)
    assert np.max(np.abs(A0.download() - val)) <= RTOL * np.abs(oval).max()
    # bit-reproducible: no atomics, fixed summation order
    ctx.set_option("tile_policy", 2)
    A.assemble(terms, qp, qw)
    assert np.array_equal(A.download(), val)


def test_tiles_policy_default_switches_on_second_assembly(ctx):
    mesh = ctx.mesh_cube(6, 6, 6)
    qp, qw = ffcuda.quadrature(3, 6)
    ctx.set_option("tile_policy", 1)
    sp = mesh.space(1, 1)
    A = sp.symbolic().matrix()
    A.assemble(fc.LAP3, qp, qw)
    v1 = A.download().copy()
    ctx.prof_enable(True)
    ctx.prof_reset()
    A.assemble(fc.LAP3, qp, qw)   # second assembly on the fespace: the tiles are built and used
    # (3-D, default policy: the fan descriptors only; the element-by-element ones are built with tile_policy 2 or tile_fans 0)
    assert ctx.prof_get("fan_build")[1] == 1 and ctx.prof_get("tile_build")[1] == 0
    A.assemble(fc.LAP3, qp, qw)   # built once
    assert ctx.prof_get("fan_build")[1] == 1
    ctx.prof_enable(False)
    assert np.max(np.abs(A.download() - v1)) <= RTOL * np.abs(v1).max()


def test_tiles_accumulate_and_bc_and_solve(ctx):
    """A = stiffness, A += mass (accumulate), Dirichlet, CG: the solve on the tile-assembled matrix equals the oracle's."""
    size = (7, 6, 5)
    m = ol.cube(*size)
    n = m["xyz"].shape[0]
    qp, qw = ffcuda.quadrature(3, 6)
    mass = [(0, fc.ID, 0, fc.ID, 3.0)]
    mesh = ctx.mesh_cube(*size)
    sp, pat, A = _assemble(ctx, mesh, fc.LAP3, qp, qw, 2, 32, accumulate_terms=mass)
    ci, cj, ca = ol.assemble_coo(m, 1, 1, None, fc.LAP3 + mass, qp, qw)
    d, v = ol.bc_pairs(m, 1, 1, None, fc.ALL6, 1, [0.0])
    ca = ol.bc_matrix_coo(ci, cj, ca, n, d, TGV)
    ob = ol.bc_rhs(ol.assemble_rhs(m, 1, 1, None, n, [(0, fc.ID, 1.0)], qp, qw), d, v, TGV)
    orp, ocol, oval = ol.coo_to_csr(n, ci, cj, ca)
    b = ctx.vec(n)
    sp.assemble_linear(b, [(0, fc.ID, 1.0)], qp, qw)
    bc = sp.bc_from_labels(fc.ALL6, 1, [0.0])
    A.apply_bc(bc, TGV)
    b.apply_bc(bc, TGV)
    val = A.download()
    big = np.abs(oval) > 1e29
    assert np.array_equal(np.abs(val) > 1e29, big)
    assert np.max(np.abs(val - oval)[~big]) <= RTOL * np.abs(oval[~big]).max()
    x = ctx.vec(n)
    it, conv, _ = A.cg(b, x, eps=1e-6, itmax=0, tgv=TGV)
    ox, oit, _, _ = ol.cg(n, ci, cj, ca, ob, np.zeros(n), eps=1e-6, itmax=0, tgv=TGV)
    assert conv == 1 and it == oit
    assert np.max(np.abs(x.download() - ox)) <= RTOL * np.abs(ox).max()


@pytest.mark.parametrize("name", ["lap3d_p1_warp", "lap2d_p1_warp", "lap3d_p1_cube342", "lap2d_p1_sq12x9", "heat3d_p1_cube3"])
def test_tiles_on_reference_meshes(ctx, name):
    """meshes and matrices dumped from the reference FreeFEM build (tests/golden): tile assembly vs FreeFEM's own values"""
    g = fc.load(name)
    order, ncomp, bt, lt, qname, bcs = fc.CASES[name]
    dim = g["dim"]
    qp, qw = ffcuda.quadrature(dim, 6)
    mesh = ctx.mesh_upload(dim, g["xyz"], g["conn"], g["elab"], g["bconn"], g["blab"], g["belem"], g["bface"])
    sp, pat, A = _assemble(ctx, mesh, bt, qp, qw, 2, 16)
    for labels, mask, values in bcs:
        A.apply_bc(sp.bc_from_labels(labels, mask, values), TGV)
    rp, col = pat.download()
    grp, gcol, gval = fc.golden_csr(g)
    assert np.array_equal(rp, grp) and np.array_equal(col, gcol)
    val = A.download()
    big = np.abs(gval) > 1e29
    assert np.array_equal(np.abs(val) > 1e29, big)
    assert np.max(np.abs(val - gval)[~big]) <= RTOL * np.abs(gval[~big]).max()


@pytest.mark.parametrize("kind", ["cube", "square"])
def test_tiles_scrambled_numbering(ctx, kind):
    """vertex numbering without locality: the tiles are clusters in space (Morton order of the coordinates), so the
    tile path does not depend on the numbering; rows come out in the mesh's own numbering."""
    m = ol.cube(9, 8, 10) if kind == "cube" else ol.square(40, 37)
    dim = m["dim"]
    nv = m["xyz"].shape[0]
    rng = np.random.default_rng(11)
    perm = rng.permutation(nv).astype(np.int32)
    inv = np.argsort(perm)
    m2 = dict(m, xyz=np.ascontiguousarray(m["xyz"][inv]), conn=perm[m["conn"]], bconn=perm[m["bconn"]])
    terms = (fc.LAP3 if dim == 3 else fc.LAP2) + [(0, fc.ID, 0, fc.ID, 2.5)]
    qp, qw = ffcuda.quadrature(dim, 6)
    orp, ocol, oval = _oracle_vals(m2, nv, terms, qp, qw)
    mesh = ctx.mesh_upload(dim, m2["xyz"], m2["conn"], m2["elab"], m2["bconn"], m2["blab"], m2["belem"], m2["bface"])
    sp, pat, A = _assemble(ctx, mesh, terms, qp, qw, 2, 64)
    rp, col = pat.download()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert np.max(np.abs(A.download() - oval)) <= RTOL * np.abs(oval).max()


def test_tiles_full_size_properties(ctx):
    """cube(64): row sums of the stiffness matrix vanish, symmetry, A.1 = 0, equality with the thread-per-row kernel"""
    import scipy.sparse as sps

    n = 64
    mesh = ctx.mesh_cube(n, n, n)
    qp, qw = ffcuda.quadrature(3, 6)
    sp, pat, A = _assemble(ctx, mesh, fc.LAP3, qp, qw, 2, 64)
    rp, col = pat.download()
    val = A.download()
    nd = (n + 1) ** 3
    M = sps.csr_matrix((val, col, rp), shape=(nd, nd))
    scale = np.abs(val).max()
    assert np.max(np.abs(M @ np.ones(nd))) <= 1e-12 * scale
    assert abs(M - M.T).max() <= 1e-13 * scale
    sp0, pat0, A0 = _assemble(ctx, mesh, fc.LAP3, qp, qw, 0)
    assert np.max(np.abs(A0.download() - val)) <= RTOL * scale


def test_lazy_positions_after_tiles_are_built(ctx):
    """Once a space has its row tiles, the symbolic phase skips the per-record positions (only the thread-per-row kernels
    read them); a later form that is not on the tile path (region filter) must get them on demand."""
    size = (9, 7, 8)
    m = ol.cube(*size)
    n = m["xyz"].shape[0]
    qp, qw = ffcuda.quadrature(3, 6)
    mesh = ctx.mesh_cube(*size)
    ctx.set_option("tile_policy", 1)
    sp = mesh.space(1, 1)
    pat0 = sp.symbolic()
    A0 = pat0.matrix()
    A0.assemble(fc.LAP3, qp, qw)
    A0.assemble(fc.LAP3, qp, qw)          # builds the tile set
    rp0, col0 = pat0.download()
    ctx.prof_enable(True)
    ctx.prof_reset()
    pat1 = sp.symbolic()                  # lazy: no positions
    rp1, col1 = pat1.download()
    assert np.array_equal(rp0, rp1) and np.array_equal(col0, col1)
    orp, ocol, oval = _oracle_vals(m, n, fc.LAP3, qp, qw)
    assert np.array_equal(rp1, orp) and np.array_equal(col1, ocol)
    A1 = pat1.matrix()
    A1.assemble(fc.LAP3, qp, qw)          # tiles
    assert ctx.prof_get("sym_p1_positions")[1] == 0
    assert np.max(np.abs(A1.download() - oval)) <= RTOL * np.abs(oval).max()
    A1.assemble(HEAT3, qp, qw, labels=[0])  # region filter: thread-per-row kernel, positions produced now (all elements are in region 0)
    assert ctx.prof_get("sym_p1_positions")[1] == 1
    ctx.prof_enable(False)
    _, _, hval = _oracle_vals(m, n, HEAT3, qp, qw)
    assert np.max(np.abs(A1.download() - hval)) <= RTOL * np.abs(hval).max()
    # Dirichlet through diagpos of the lazily built pattern
    bc = sp.bc_from_labels(fc.ALL6, 1, [0.0])
    A1.apply_bc(bc, TGV)
    v = A1.download()
    d, _ = ol.bc_pairs(m, 1, 1, None, fc.ALL6, 1, [0.0])
    diag_idx = np.array([rp1[i] + np.searchsorted(col1[rp1[i]:rp1[i + 1]], i) for i in np.unique(d)])
    assert np.all(v[diag_idx] == TGV)


RHS_CASES = [("cube", (9, 7, 8), [(0, fc.ID, 1.0)]), ("cube", (6, 6, 5), [(0, fc.ID, 2.0), (0, fc.DX, 0.5), (0, fc.DZ, -1.5)]),
             ("square", (23, 17), [(0, fc.ID, 1.0)]), ("square", (14, 19), [(0, fc.DY, 3.0), (0, fc.ID, -0.25)])]


@pytest.mark.parametrize("rows", [16, 96])
@pytest.mark.parametrize("case", RHS_CASES, ids=[f"{k}{'x'.join(map(str, s))}_{len(t)}t" for k, s, t in RHS_CASES])
def test_rhs_by_tiles_against_oracle(ctx, case, rows):
    """right-hand side through the row tiles (value and gradient terms) = oracle = thread-per-row kernel; accumulation"""
    kind, size, lt = case
    dim = 3 if kind == "cube" else 2
    m = ol.cube(*size) if kind == "cube" else ol.square(*size)
    mesh = ctx.mesh_cube(*size) if kind == "cube" else ctx.mesh_square(*size)
    qp, qw = ffcuda.quadrature(dim, 6)
    n = m["xyz"].shape[0]
    ob = ol.assemble_rhs(m, 1, 1, None, n, lt, qp, qw)
    sp, pat, A = _assemble(ctx, mesh, fc.LAP3 if dim == 3 else fc.LAP2, qp, qw, 2, rows)   # builds the tile set
    ctx.prof_enable(True)
    ctx.prof_reset()
    b = ctx.vec(n)
    sp.assemble_linear(b, lt, qp, qw)
    hb = b.download()
    assert np.max(np.abs(hb - ob)) <= RTOL * np.abs(ob).max()
    sp.assemble_linear(b, lt, qp, qw, accumulate=True)
    assert np.max(np.abs(b.download() - 2 * ob)) <= 2 * RTOL * np.abs(ob).max()
    ctx.prof_enable(False)
    ctx.set_option("tile_policy", 0)
    b0 = ctx.vec(n)
    sp.assemble_linear(b0, lt, qp, qw)
    assert np.max(np.abs(b0.download() - hb)) <= RTOL * np.abs(ob).max()
    # reproducible
    ctx.set_option("tile_policy", 2)
    b1 = ctx.vec(n)
    sp.assemble_linear(b1, lt, qp, qw)
    assert np.array_equal(b1.download(), hb)


def test_pattern_download_async_overlaps_and_matches(ctx):
    """the asynchronous pattern download (second stream, behind the symbolic phase) returns the same CSR as the blocking one"""
    import torch

    mesh = ctx.mesh_cube(14, 11, 9)
    qp, qw = ffcuda.quadrature(3, 6)
    sp = mesh.space(1, 1)
    pat = sp.symbolic()
    n, nnz = pat.info()
    rp = torch.empty(n + 1, dtype=torch.int32, pin_memory=True).numpy()
    ci = torch.empty(nnz, dtype=torch.int32, pin_memory=True).numpy()
    rp[:] = -1
    ci[:] = -1
    pat.download_async(rp, ci)
    A = pat.matrix()
    A.assemble(fc.LAP3, qp, qw)       # overlaps the copies
    ctx.sync()
    rp0, ci0 = pat.download()
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)


@pytest.mark.parametrize("size", [(7, 6, 9), (20, 17, 13)])
def test_fans_against_oracle_and_round1_tiles(ctx, size):
    """3-D: the fan kernels (stiffness, heat form with its mass term, value-only right-hand side) against the oracle and
    against the element-by-element tile kernels of round 1 (tile_fans = 0): same pattern, values 1e-12, bit-identical from
    run to run; the fan kernels are the ones that run (asm_rows_p1 / rhs_rows launches of a space with tiles)."""
    m = ol.cube(*size)
    n = m["xyz"].shape[0]
    qp, qw = ffcuda.quadrature(3, 6)
    heat = [(0, fc.ID, 0, fc.ID, 100.0)] + fc.LAP3
    rhs = [(0, fc.ID, 2.5)]
    ob = ol.assemble_rhs(m, 1, 1, None, n, rhs, qp, qw)
    res = {}
    for fans in (1, 0):
        ctx.set_option("tile_fans", fans)
        mesh = ctx.mesh_cube(*size)
        sp, pat, A = _assemble(ctx, mesh, fc.LAP3, qp, qw, 2, 96)
        rp, col = pat.download()
        orp, ocol, oval = _oracle_vals(m, n, fc.LAP3, qp, qw)
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
        v = A.download()
        assert np.max(np.abs(v - oval)) <= RTOL * np.abs(oval).max()
        A.assemble(heat, qp, qw)
        hv = A.download()
        _, _, ohv = _oracle_vals(m, n, heat, qp, qw)
        assert np.max(np.abs(hv - ohv)) <= RTOL * np.abs(ohv).max()
        A.assemble(heat, qp, qw)
        assert np.array_equal(A.download(), hv)          # bit-reproducible
        A.assemble(fc.LAP3, qp, qw, accumulate=True)      # accumulate on top of the heat matrix
        assert np.max(np.abs(A.download() - (ohv + oval))) <= RTOL * np.abs(ohv).max()
        b = ctx.vec(n)
        sp.assemble_linear(b, rhs, qp, qw)
        hb = b.download()
        assert np.max(np.abs(hb - ob)) <= RTOL * np.abs(ob).max()
        sp.assemble_linear(b, rhs, qp, qw, accumulate=True)
        assert np.max(np.abs(b.download() - 2 * ob)) <= RTOL * np.abs(ob).max()
        res[fans] = (v, hv, hb)
    ctx.set_option("tile_fans", 1)
    for a, b_ in zip(res[0], res[1]):
        assert np.max(np.abs(a - b_)) <= RTOL * np.abs(a).max()
