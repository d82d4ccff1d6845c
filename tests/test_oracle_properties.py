"""Internal consistency of the oracle's entries that take data at the quadrature nodes (CPU): with a constant in the table
they must reproduce the constant-coefficient restatements, on meshes larger than the golden fixtures.  (The same property is
checked on the device at larger sizes in tests/test_gpu_parity.py.)"""
import numpy as np
import pytest

import ff_cases as fc
import oracle_lib as ol


@pytest.mark.parametrize("order,ncomp", [(1, 1), (1, 3), (2, 1), (2, 2)])
def test_tables_of_a_constant_equal_the_constant_forms(order, ncomp):
    N = 5
    m = ol.cube(N, N, N)
    qp, qw = ol.quadrature(3, "qfV5")
    fq, fw = ol.face_quadrature(3)
    e2n = ol.p2_nodes_3d(m["xyz"].shape[0], m["conn"])[0] if order == 2 else None
    nn = (int(e2n.max()) + 1) if order == 2 else m["xyz"].shape[0]
    n = nn * ncomp
    nt, nbe = m["conn"].shape[0], len(m["blab"])
    bt = []
    for c in range(ncomp):
        bt += [(c, fc.DX, c, fc.DX, 1.0), (c, fc.DZ, c, fc.DZ, 1.0), (c, fc.ID, c, fc.ID, 2.0), (c, fc.DX, (c + 1) % ncomp, fc.ID, 0.5)]
    kappa = 1.75
    i0, j0, a0 = ol.assemble_coo(m, order, ncomp, e2n, [(uc, uo, vc, vo, kappa * v) for uc, uo, vc, vo, v in bt], qp, qw)
    i1, j1, a1 = ol.assemble_coo_qcoef(m, order, ncomp, e2n, bt, qp, qw, np.full((nt, len(qw)), kappa))
    assert np.array_equal(i0, i1) and np.array_equal(j0, j1)
    assert np.max(np.abs(a0 - a1)) <= 1e-14 * np.abs(a0).max()
    rb = [(c, fc.ID, c, fc.ID, 3.0) for c in range(ncomp)]
    i0, j0, a0 = ol.assemble_coo_boundary(m, order, ncomp, e2n, [(uc, uo, vc, vo, kappa * v) for uc, uo, vc, vo, v in rb], fq, fw, [2, 5])
    i1, j1, a1 = ol.assemble_coo_boundary_qcoef(m, order, ncomp, e2n, rb, fq, fw, np.full((nbe, len(fw)), kappa), [2, 5])
    assert np.array_equal(i0, i1) and np.array_equal(j0, j1)
    assert np.max(np.abs(a0 - a1)) <= 1e-14 * np.abs(a0).max()
    lt = [(c, fc.ID, 1.0 + c) for c in range(ncomp)] + [(0, fc.DX, 0.5), (ncomp - 1, fc.DZ, -2.0)]
    b0 = ol.assemble_rhs(m, order, ncomp, e2n, n, lt, qp, qw)
    fqt = np.zeros((ncomp, 4, nt, len(qw)))
    for c, op, v in lt:
        fqt[c, {fc.ID: 0, fc.DX: 1, fc.DY: 2, fc.DZ: 3}[op]] += v
    b1 = ol.assemble_rhs_qterms(m, order, ncomp, e2n, np.zeros(n), qp, qw, fqt)
    assert np.max(np.abs(b0 - b1)) <= 1e-14 * np.abs(b0).max()
    b0 = ol.assemble_rhs(m, order, ncomp, e2n, n, lt[:ncomp], qp, qw)
    b2 = ol.assemble_rhs_qvalues(m, order, ncomp, e2n, np.zeros(n), qp, qw, fqt[:, 0])
    assert np.max(np.abs(b0 - b2)) <= 1e-14 * np.abs(b0).max()
    g0 = ol.assemble_rhs_boundary(m, order, ncomp, e2n, np.zeros(n), [(c, fc.ID, 0.7) for c in range(ncomp)], fq, fw, [1, 6])
    gq = np.full((ncomp, nbe, len(fw)), 0.7) * np.isin(m["blab"], [1, 6])[None, :, None]
    g1 = ol.assemble_rhs_boundary_qvalues(m, order, ncomp, e2n, np.zeros(n), fq, fw, gq)
    assert np.abs(g0).max() > 0 and np.max(np.abs(g0 - g1)) <= 1e-14 * np.abs(g0).max()
    # the physical quadrature nodes the tables are evaluated at: affine images of the reference nodes, inside their elements
    P = ol.quad_points_xyz(m, qp)
    assert P.shape == (nt, len(qw), 3) and P.min() >= 0 and P.max() <= 1
    PB = ol.bquad_points_xyz(m, fq)
    onface = np.minimum(np.abs(PB), np.abs(1 - PB)).min(axis=2)      # every boundary node lies on a face of the unit cube
    assert PB.shape == (nbe, len(fw), 3) and onface.max() <= 1e-14


def test_gmres_restatement_on_a_random_nonsymmetric_matrix():
    """ffo_gmres against a dense solve: convergence to 1e-12 with and without restarts, iteration counter monotone in eps."""
    rng = np.random.default_rng(3)
    n = 60
    A = np.eye(n) * 4 + rng.standard_normal((n, n)) * 0.3
    ci, cj = np.nonzero(A)
    ca = A[ci, cj]
    b = rng.standard_normal(n)
    xe = np.linalg.solve(A, b)
    its = []
    for eps in (1e-4, 1e-8, 1e-12):
        x, it, ret, rel = ol.gmres(n, ci, cj, ca, b, np.zeros(n), eps=eps, nbkrylov=1000)
        assert ret == 1 and rel < eps
        its.append(it)
    assert its[0] <= its[1] <= its[2] and np.max(np.abs(x - xe)) <= 1e-10 * np.abs(xe).max()
    x, it, ret, rel = ol.gmres(n, ci, cj, ca, b, np.zeros(n), eps=1e-12, nbkrylov=7)
    assert ret == 1 and it >= its[2] and np.max(np.abs(x - xe)) <= 1e-10 * np.abs(xe).max()
    x, it, ret, rel = ol.gmres(n, ci, cj, ca, b, np.zeros(n), eps=1e-12, itmax=3, nbkrylov=1000)
    assert ret == 0 and it <= 6


def test_bench_closed_form_cube_pattern_matches_the_oracle():
    """bench.py's `parity` object compares every rank's rows with BuildCube's pattern in closed form (row lengths and an
    order-independent hash of the global column set): the closed form itself is pinned here against the oracle's CSR, on
    anisotropic cubes, with rows given in any order and columns in any order inside a row, and it notices a wrong column."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(3)
    for dims in [(1, 1, 1), (3, 2, 4), (5, 7, 3)]:
        m = ol.cube(*dims)
        n = m["xyz"].shape[0]
        qp, qw = ol.quadrature(3, "qfV5")
        ci, cj, ca = ol.assemble_coo(m, 1, 1, None, [(0, 1, 0, 1, 1.0)], qp, qw)
        rp, col, _ = ol.coo_to_csr(n, ci, cj, ca)
        assert bench.cube_pattern_mismatches(np, *dims, np.arange(n), rp, col) == 0
        # a rank's view: a subset of the rows in another order, columns shuffled inside every row
        rows = rng.permutation(n)[: max(1, n // 2)]
        lens = np.diff(rp)[rows]
        rp2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        col2 = np.concatenate([rng.permutation(col[rp[r]:rp[r + 1]]) for r in rows])
        assert bench.cube_pattern_mismatches(np, *dims, rows, rp2, col2) == 0
        bad = col2.copy()
        bad[0] = (bad[0] + 1) % n if (bad[0] + 1) % n not in col2[rp2[0]:rp2[1]] else (bad[0] + 2) % n
        assert bench.cube_pattern_mismatches(np, *dims, rows, rp2, bad) >= 1
