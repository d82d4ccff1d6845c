"""Case table shared by the oracle tests and the GPU parity tests: for every golden fixture, the
variational form exactly as FreeFEM stores it (term lists after LinearComb merging),
the quadrature rule the script selects, and the boundary conditions of the .edp that produced it
(tests/golden/make_golden.py)."""
import os
import numpy as np

ID, DX, DY, DZ = 0, 1, 2, 6
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LAP2 = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0)]
LAP3 = [(0, DX, 0, DX, 1.0), (0, DY, 0, DY, 1.0), (0, DZ, 0, DZ, 1.0)]

_E, _SIG = 21.5e4, 0.29
MU = _E / (2 * (1 + _SIG))
LAMBDA = _E * _SIG / ((1 + _SIG) * (1 - 2 * _SIG))


def lame_terms():
    """21 terms (ucomp,uop,vcomp,vop,coef) in the order FreeFEM holds them."""
    d = [DX, DY, DZ]
    order = [((0, 1), (0, 1)), ((0, 1), (1, 2)), ((0, 1), (2, 6)), ((1, 2), (0, 1)), ((1, 2), (1, 2)), ((1, 2), (2, 6)),
             ((2, 6), (0, 1)), ((2, 6), (1, 2)), ((2, 6), (2, 6)), ((1, 6), (1, 6)), ((1, 6), (2, 2)), ((2, 2), (1, 6)),
             ((2, 2), (2, 2)), ((0, 6), (0, 6)), ((0, 6), (2, 1)), ((2, 1), (0, 6)), ((2, 1), (2, 1)), ((0, 2), (0, 2)),
             ((0, 2), (1, 1)), ((1, 1), (0, 2)), ((1, 1), (1, 1))]
    out = []
    for (uc, uo), (vc, vo) in order:
        if uo == d[uc] and vo == d[vc]:  # div-div (+ diagonal of 2 mu eps:eps)
            c = LAMBDA + 2.0 * MU if uc == vc else LAMBDA
        else:
            c = MU
        out.append((uc, uo, vc, vo, c))
    return out


# name -> (order, ncomp, bilinear terms, linear terms, quadrature name, [ (labels, compmask, values) ... ])
ALL6 = [1, 2, 3, 4, 5, 6]
CASES = {
    "lap2d_p1_sq4": (1, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([1, 2, 3, 4], 1, [0.0])]),
    "lap2d_p1_sq12x9": (1, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([1, 2, 3, 4], 1, [0.0])]),
    "lap2d_p1_warp": (1, 1, LAP2, [(0, ID, 3.0)], "qf5pT", [([1], 1, [1.0]), ([3], 1, [2.0])]),
    "lap2d_p2_sq3": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([1, 2, 3, 4], 1, [0.0])]),
    "lap2d_p2_warp": (2, 1, LAP2 + [(0, ID, 0, ID, 2.0)], [(0, ID, 1.0)], "qf5pT", [([2, 4], 1, [0.0])]),
    "lap3d_p1_cube2": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "lap3d_p1_cube5": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "lap3d_p1_cube342": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "lap3d_p1_warp": (1, 1, LAP3, [(0, ID, 2.0)], "qfV5", [([1], 1, [1.0]), ([6], 1, [-1.0])]),
    "lap3d_p2_cube2": (2, 1, LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "heat3d_p1_cube3": (1, 1, [(0, ID, 0, ID, 1.0 / 0.01)] + LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "mass3d_p1_lump": (1, 1, [(0, ID, 0, ID, 1.0)], [(0, ID, 1.0)], "qfV1lump", []),
    "mass2d_p2_qf2": (2, 1, [(0, ID, 0, ID, 1.0)], [(0, ID, 1.0)], "qf2pT", []),
    "nonsym3d_p1": (1, 1, [(0, DX, 0, ID, 1.0), (0, ID, 0, DY, 2.0), (0, DZ, 0, DX, 0.5)],
                    [(0, DX, 1.0), (0, ID, 2.0)], "qfV5", []),
    "nonsym2d_p2": (2, 1, [(0, DX, 0, ID, 1.0), (0, ID, 0, DY, 2.0), (0, DY, 0, DX, 0.5)],
                    [(0, DY, 1.0), (0, ID, 2.0)], "qf5pT", []),
    "lame3d_p2_cube2": (2, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "lame3d_p1_cube3": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "lame3d_p2_warp": (2, 3, lame_terms(), [(2, ID, -0.05)], "qfV5",
                       [([1], 7, [0.0, 0.0, 0.0]), ([3], 7, [0.01, 0.0, -0.02])]),
    # exact elimination (tgv < 0, CASE_TGV below)
    "lap3d_p1_tgvm1": (1, 1, LAP3, [(0, ID, 2.0)], "qfV5", [([1], 1, [1.0]), ([6], 1, [-1.0])]),
    "lap3d_p1_tgvm2": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "lap2d_p2_tgvm2": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([1, 2, 3, 4], 1, [0.0])]),
    "lame3d_p1_tgvm1": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "lap3d_p1_tgvm3": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [([1, 3], 1, [0.0])]),
    # boundary integrals in the linear form (CASE_BLIN below)
    "lap3d_p1_neumann": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [([1], 1, [0.0])]),
    "lap2d_p2_neumann": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p1_traction": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "lap3d_p2_neumann": (2, 1, LAP3 + [(0, ID, 0, ID, 1.0)], [(0, ID, 1.0)], "qfV5", []),
    # Robin terms: boundary integrals in the bilinear form (CASE_BBIL below)
    "lap3d_p1_robin": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [([1], 1, [0.0])]),
    "lap2d_p2_robin": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p1_robin": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "lap3d_p2_robin": (2, 1, LAP3, [(0, ID, 1.0)], "qfV5", []),
    # non-symmetric forms, solved by GMRES in the fixture (CASE_GMRES below)
    "convdiff3d_p1_gmres": (1, 1, LAP3 + [(0, DX, 0, ID, 8.0), (0, DY, 0, ID, 3.0), (0, DZ, 0, ID, -2.0)], [(0, ID, 1.0)], "qfV5",
                            [(ALL6, 1, [0.0])]),
    "convdiff2d_p2_gmres": (2, 1, LAP2 + [(0, DX, 0, ID, 5.0), (0, ID, 0, ID, 1.0)], [(0, ID, 1.0)], "qf5pT", [([1, 3], 1, [0.0])]),
    # right-hand sides with data depending on the mesh point (CASE_FQ below gives the data)
    "lap3d_p1_fxyz": (1, 1, LAP3, [], "qfV5", [([1, 2], 1, [0.0])]),
    "lap2d_p2_fxy": (2, 1, LAP2, [], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p1_fvec": (1, 3, lame_terms(), [], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    # bilinear forms with coefficients depending on the mesh point: the constant part here, the rest in CASE_QCOEF below
    "diff3d_p1_kappa": (1, 1, [(0, ID, 0, ID, 2.0)], [(0, ID, 1.0)], "qfV5", [([1, 2], 1, [0.0])]),
    "reac2d_p1_rho": (1, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p1_evar": (1, 3, [], [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    # boundary integrals with data depending on the mesh point (CASE_BQ below)
    "lap3d_p1_bnd_g": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [([1], 1, [0.0])]),
    "lap2d_p2_bnd_g": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p1_bnd_g": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "diff3d_p2_kappa": (2, 1, [(0, ID, 0, ID, 2.0)], [(0, ID, 1.0)], "qfV5", [([1, 2], 1, [0.0])]),
    "reac2d_p2_rho": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p2_evar": (2, 3, [], [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    # right-hand sides with derivatives of the test function times data depending on the mesh point (CASE_FQT below)
    "resid2d_p1_grad": (1, 1, LAP2, [], "qf5pT", [([4], 1, [0.0])]),
    "resid3d_p2_grad": (2, 1, LAP3, [], "qfV5", [([1, 2], 1, [0.0])]),
    "resid3d_p1_vec_grad": (1, 3, lame_terms(), [], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    # half storage (sym=1, CASE_SYM below): the fixture holds the lower triangle
    "lap3d_p1_sym": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [(ALL6, 1, [0.0])]),
    "lap2d_p2_sym": (2, 1, LAP2 + [(0, ID, 0, ID, 2.0)], [(0, ID, 1.0)], "qf5pT", [([2, 4], 1, [0.0])]),
    "lame3d_p1_sym": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    # boundary integrals with derivatives of the unknown / test function (CASE_BBIL / CASE_BLIN / CASE_BQ below)
    "lap3d_p1_bnd_grad": (1, 1, LAP3, [(0, ID, 1.0)], "qfV5", [([1], 1, [0.0])]),
    "lap2d_p2_bnd_grad": (2, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([4], 1, [0.0])]),
    "lame3d_p1_bnd_grad": (1, 3, lame_terms(), [(2, ID, -0.05)], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
    "lap3d_p2_bnd_gradq": (2, 1, LAP3, [(0, ID, 1.0)], "qfV5", [([1], 1, [0.0])]),
}
# the reference's own regression problems (examples/tutorial/regtests.edp, values in ref.edp): Laplace.edp, LaplaceP1.edp (Robin +
# Neumann on label 1), beam.edp ([P1,P1] elasticity on the mesh buildmesh gives it); tgv = 1e5 as in the scripts
_E2, _S2 = 21.5, 0.29
_MU2, _LA2 = _E2 / (2 * (1 + _S2)), _E2 * _S2 / ((1 + _S2) * (1 - 2 * _S2))
BEAM2 = [(0, DX, 0, DX, _LA2 + 2 * _MU2), (1, DY, 1, DY, _LA2 + 2 * _MU2), (0, DX, 1, DY, _LA2), (1, DY, 0, DX, _LA2),
         (0, DY, 0, DY, _MU2), (1, DX, 1, DX, _MU2), (0, DY, 1, DX, _MU2), (1, DX, 0, DY, _MU2)]
# name -> (case tuple as in CASES, tgv, Robin item, Neumann item, what regtests.edp asserts on u'*u: reference value, tolerance)
TUTORIAL_CASES = {
    "tutorial_laplace": ((1, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([1, 2, 3, 4], 1, [0.0])]), 1e5, None, None, (0.167397, 1e-2)),
    "tutorial_laplace_p1": ((1, 1, LAP2, [(0, ID, 1.0)], "qf5pT", [([2, 3, 4], 1, [0.0])]), 1e5, ([1], [(0, ID, 0, ID, 1.0)]),
                            ([1], [(0, ID, 1.0)]), (2.34669, 1e-2)),
    "tutorial_beam": ((1, 2, BEAM2, [(1, ID, -0.05)], "qf5pT", [([1], 3, [0.0, 0.0])]), 1e30, None, None, (2.19089, 5e-2)),
}
# boundary integrals of the linear form: name -> (labels, terms)
CASE_BLIN = {"lap3d_p1_neumann": ([2, 3], [(0, ID, 2.5)]), "lap2d_p2_neumann": ([2], [(0, ID, 1.5)]),
             "lame3d_p1_traction": ([2], [(0, ID, 0.3), (2, ID, -0.2)]), "lap3d_p2_neumann": ([6], [(0, ID, -1.0)])}
CASE_BLIN["lap3d_p1_robin"] = ([2, 3], [(0, ID, 2.5)])
# boundary integrals of the bilinear form (Robin terms): name -> (labels, terms (ucomp, uop, vcomp, vop, c))
CASE_BBIL = {"lap3d_p1_robin": ([2, 3], [(0, ID, 0, ID, 1.5)]), "lap2d_p2_robin": ([2, 3], [(0, ID, 0, ID, 0.7)]),
             "lame3d_p1_robin": ([2], [(0, ID, 0, ID, 1e4), (1, ID, 1, ID, 1e4), (2, ID, 2, ID, 5e3), (0, ID, 2, ID, 2e3)]),
             "lap3d_p2_robin": ([6, 1], [(0, ID, 0, ID, 2.0)])}
# fixtures solved with solver=GMRES: name -> dimKrylov (FreeFEM's default is 1000)
CASE_GMRES = {"convdiff3d_p1_gmres": 1000, "convdiff2d_p2_gmres": 25}
# data of the linear form at the points P (..., dim) -> (ncomp, ...) values, for the fixtures whose rhs depends on the mesh point
CASE_FQ = {
    "lap3d_p1_fxyz": lambda P: (P[..., 0] * P[..., 1] + np.sin(P[..., 2]))[None],
    "lap2d_p2_fxy": lambda P: (np.exp(P[..., 0]) * P[..., 1])[None],
    "lame3d_p1_fvec": lambda P: np.stack([P[..., 0], 0 * P[..., 0], -0.05 * (1 + P[..., 1])]),
}
# groups of terms multiplied by a coefficient that depends on the mesh point: name -> [(coefficient at points P (..., dim), terms)]
CASE_QCOEF = {
    "diff3d_p1_kappa": [(lambda P: 1 + P[..., 0] * P[..., 1] + P[..., 2] ** 2, LAP3)],
    "reac2d_p1_rho": [(lambda P: 1 + np.sin(P[..., 0]) * P[..., 1], [(0, ID, 0, ID, 1.0)])],
    "lame3d_p1_evar": [(lambda P: 1 + P[..., 0], lame_terms())],
    "diff3d_p2_kappa": [(lambda P: 1 + P[..., 0] * P[..., 1] + P[..., 2] ** 2, LAP3)],
    "reac2d_p2_rho": [(lambda P: 1 + np.sin(P[..., 0]) * P[..., 1], [(0, ID, 0, ID, 1.0), (0, DX, 0, ID, 0.5), (0, ID, 0, DX, 0.5)])],
    "lame3d_p2_evar": [(lambda P: 1 + P[..., 0], lame_terms())],
}
CASE_BLIN["lap3d_p1_bnd_g"] = ([6], [(0, ID, 0.5)])
# boundary data depending on the mesh point: name -> dict(lin=(labels, g at points P (..., dim) -> (ncomp, ...)),
#                                                        bil=(labels, coefficient at P, terms))
CASE_BQ = {
    "lap3d_p1_bnd_g": dict(lin=([2, 3], lambda P: (P[..., 1] + np.sin(P[..., 2]))[None]),
                           bil=([2, 3], lambda P: 1 + P[..., 0] * P[..., 2], [(0, ID, 0, ID, 1.0)])),
    "lap2d_p2_bnd_g": dict(lin=([2], lambda P: np.exp(P[..., 1])[None]),
                           bil=([2, 3], lambda P: 1 + P[..., 0] * P[..., 1], [(0, ID, 0, ID, 1.0)])),
    "lame3d_p1_bnd_g": dict(lin=([2], lambda P: np.stack([0.3 * P[..., 2], 0 * P[..., 0], -0.2 * (1 + P[..., 1])])),
                            bil=([3], lambda P: 1 + P[..., 0], [(c, ID, c, ID, 1e3) for c in range(3)])),
}
# ... with derivatives on the faces: dx(u) v, u dy(v), dz(u) dx(v) (every node of the adjacent element is reached)
CASE_BBIL["lap3d_p1_bnd_grad"] = ([2, 3], [(0, DX, 0, ID, 1.5), (0, ID, 0, DY, 0.5), (0, DZ, 0, DX, 2.0), (0, ID, 0, ID, 0.25)])
CASE_BLIN["lap3d_p1_bnd_grad"] = ([6], [(0, DZ, 0.3), (0, ID, -1.0)])
CASE_BBIL["lap2d_p2_bnd_grad"] = ([2, 3], [(0, DX, 0, ID, 0.7), (0, DY, 0, DY, 0.2)])
CASE_BLIN["lap2d_p2_bnd_grad"] = ([2], [(0, DX, 1.5), (0, ID, 0.5)])
CASE_BBIL["lame3d_p1_bnd_grad"] = ([3], [(0, DX, 1, ID, 1e3), (2, ID, 0, DZ, 1e3), (1, ID, 1, ID, 1e3)])
CASE_BLIN["lame3d_p1_bnd_grad"] = ([2], [(0, DY, 0.3), (2, ID, -0.2)])
CASE_BQ["lap3d_p2_bnd_gradq"] = dict(lin=([2], lambda P: (0 * P[..., 0])[None]),
                                     bil=([2, 3], lambda P: 1 + P[..., 0] * P[..., 2], [(0, DX, 0, ID, 1.0), (0, ID, 0, DZ, 0.5)]))
# data of the linear form per (component, slot): P (..., dim) -> (ncomp, dim+1, ...); slot 0 = value, 1..dim = dx, dy, dz
def _z(P):
    return 0 * P[..., 0]


CASE_FQT = {
    "resid2d_p1_grad": lambda P: np.stack([np.stack([P[..., 0] * P[..., 1], np.sin(P[..., 0]), -P[..., 1]])]),
    "resid3d_p2_grad": lambda P: np.stack([np.stack([P[..., 0], -P[..., 1] * P[..., 2], _z(P), 1 + P[..., 2]])]),
    "resid3d_p1_vec_grad": lambda P: np.stack([np.stack([_z(P), P[..., 0], P[..., 2], _z(P)]),
                                               np.stack([P[..., 1], _z(P), _z(P), _z(P)]),
                                               np.stack([_z(P), _z(P), _z(P), -0.05 * (1 + P[..., 0])])]),
}
# cases assembled with sym=1: MatriceMorse keeps the entries (i, j) with j <= i only (HashMatrix.cpp:1319-1325)
CASE_SYM = {"lap3d_p1_sym", "lap2d_p2_sym", "lame3d_p1_sym"}
# Dirichlet treatment of a case: penalty tgv = 1e30 unless listed here (HashMatrix::SetBC with tgv < 0)
CASE_TGV = {"lap3d_p1_tgvm1": -1.0, "lap3d_p1_tgvm2": -2.0, "lap2d_p2_tgvm2": -2.0, "lame3d_p1_tgvm1": -1.0, "lap3d_p1_tgvm3": -3.0}


def load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    for k in ("dim", "ndof"):  # mesh-only fixtures (buildlayers) carry neither
        if k in g:
            g[k] = int(g[k])
    return g


def lower(n, rp, ci, val):
    """lower triangle (j <= i) of a CSR matrix with sorted rows: what sym=1 stores"""
    rows = np.repeat(np.arange(n), np.diff(rp))
    keep = ci <= rows
    hrp = np.zeros(n + 1, np.int64)
    np.add.at(hrp, rows[keep] + 1, 1)
    return np.cumsum(hrp).astype(np.int32), ci[keep], val[keep]


def expand_lower(n, rp, ci, val):
    """full symmetric CSR from its lower triangle"""
    rows = np.repeat(np.arange(n), np.diff(rp))
    off = ci < rows
    fi = np.concatenate([rows, ci[off]])
    fj = np.concatenate([ci, rows[off]])
    fv = np.concatenate([val, val[off]])
    o = np.argsort(fi.astype(np.int64) * n + fj, kind="stable")
    frp = np.zeros(n + 1, np.int64)
    np.add.at(frp, fi + 1, 1)
    return np.cumsum(frp).astype(np.int32), fj[o].astype(np.int32), fv[o]


def golden_csr(g):
    """(rowptr, colind, vals) of the reference matrix = its COO sorted by (i,j) (HashMatrix::CSR())."""
    n = g["ndof"]
    key = g["coo_i"].astype(np.int64) * n + g["coo_j"]
    o = np.argsort(key, kind="stable")
    rp = np.zeros(n + 1, np.int32)
    np.add.at(rp, g["coo_i"] + 1, 1)
    return np.cumsum(rp).astype(np.int32), g["coo_j"][o].astype(np.int32), g["coo_a"][o]


# Rectangular matrices `matrix B = vb(Uh,Vh)` (tests/golden/make_golden_rect.py): rows = test space, columns = space of the
# unknown.  name -> ((order_v, ncomp_v), (order_u, ncomp_u), terms (ucomp,uop,vcomp,vop,coef), quadrature name)
RECT_CASES = {
    "rect3d_div_p2p1": ((1, 1), (2, 3), [(0, DX, 0, ID, -1.0), (1, DY, 0, ID, -1.0), (2, DZ, 0, ID, -1.0), (1, ID, 0, DX, 0.5)], "qfV5"),
    "rect3d_grad_p1p2": ((2, 3), (1, 1), [(0, ID, 0, DX, -1.0), (0, ID, 1, DY, -1.0), (0, ID, 2, DZ, -1.0), (0, DX, 2, ID, 1.0)], "qfV5"),
    "rect2d_p1_to_p2": ((2, 1), (1, 1), [(0, ID, 0, ID, 1.0), (0, DX, 0, DY, 0.5)], "qf5pT"),
    "rect2d_div_p2p1": ((1, 1), (2, 2), [(0, DX, 0, ID, -1.0), (1, DY, 0, ID, -1.0)], "qf5pT"),
    "rect3d_p1vec_p1": ((1, 1), (1, 3), [(0, DX, 0, ID, 1.0), (2, ID, 0, DZ, 2.0)], "qfV5"),
    "rect2d_p2vec_p1vec": ((1, 2), (2, 2), [(0, ID, 0, ID, 1.0), (1, ID, 1, ID, 1.0), (0, DY, 1, ID, 0.25)], "qf2pT"),
}


# Mixed-order product spaces in one fespace (Taylor-Hood): name -> (orders of the components, terms, quadrature name).  The
# local dofs of an element are component-major (begin_dfcomp / end_dfcomp, femlib/FESpacen.hpp:454-455); the plugin assembles
# one scalar block per couple of components, with the GLOBAL dofs of the component as its "nodes" (mixed_blocks below)
def _stokes(dim):
    d = [DX, DY, DZ][:dim]
    t = [(c, o, c, o, 1.0) for c in range(dim) for o in d]                    # grad u_c . grad v_c
    t += [(dim, ID, c, d[c], -1.0) for c in range(dim)]                      # - p div v
    t += [(c, d[c], dim, ID, -1.0) for c in range(dim)]                      # - div u q
    return t + [(dim, ID, dim, ID, -1e-10)]


MIXED_CASES = {
    "mixed2d_stokes": ([2, 2, 1], _stokes(2), "qf5pT"),
    "mixed3d_stokes": ([2, 2, 2, 1], _stokes(3), "qfV5"),
}


def mixed_blocks(g, orders, terms):
    """[(order_v, table_v, order_u, table_u, terms of the block re-indexed to component 0)] for every couple of components;
    table_c[k, a] = global dof of local node a of component c in element k"""
    dim = int(g["dim"])
    nl = [(dim + 1) if o == 1 else (6 if dim == 2 else 10) for o in orders]
    beg = np.concatenate([[0], np.cumsum(nl)])
    dof = g["dof_Vh"]
    assert dof.shape[1] == beg[-1] and np.array_equal(dof, g["dof_Uh"])
    tab = [np.ascontiguousarray(dof[:, beg[c]:beg[c + 1]], dtype=np.int32) for c in range(len(orders))]
    out = []
    for cv in range(len(orders)):
        for cu in range(len(orders)):
            bt = [(0, uo, 0, vo, co) for (uc, uo, vc, vo, co) in terms if uc == cu and vc == cv]
            out.append((orders[cv], tab[cv], orders[cu], tab[cu], bt or [(0, ID, 0, ID, 0.0)]))  # no term: the couples, with zeros
    return out


def rect_elem2node(g, which, ncomp):
    """node table of one space of a rectangular fixture (which = "Uh" | "Vh"), see elem2node"""
    dof = g["dof_" + which]
    nl = dof.shape[1] // ncomp
    e2n = dof[:, :nl] // ncomp
    for c in range(ncomp):
        assert np.array_equal(dof[:, c * nl:(c + 1) * nl], e2n * ncomp + c)
    return np.ascontiguousarray(e2n, dtype=np.int32)


def elem2node(g, order, ncomp):
    """node table (nt x nloc) recovered from the reference dof table: dof(k, c*nloc+a) = node*ncomp + c."""
    dof = g["dof"]
    nl = dof.shape[1] // ncomp
    e2n = dof[:, :nl] // ncomp
    for c in range(ncomp):
        assert np.array_equal(dof[:, c * nl:(c + 1) * nl], e2n * ncomp + c)
    return np.ascontiguousarray(e2n, dtype=np.int32)
# cases whose script does not solve (non-symmetric after tgv = -1 / -3 elimination)
NO_SOLVE_TGV = {"lap3d_p1_tgvm1", "lame3d_p1_tgvm1", "lap3d_p1_tgvm3", "lame3d_p1_robin",  # (non-symmetric Robin coupling)
                "convdiff3d_p1_gmres", "convdiff2d_p2_gmres",  # (GMRES fixtures have their own solve tests)
                "lap3d_p1_bnd_grad", "lap2d_p2_bnd_grad", "lame3d_p1_bnd_grad", "lap3d_p2_bnd_gradq"}  # (non-symmetric, not solved)


def loose_iterate(name):
    """fixtures whose eps=1e-6 iterate is compared loosely (1e-6) because their right-hand side or matrix carries extra ulp
    differences (boundary terms, data evaluated with another libm); their eps=1e-14 solutions are held to 1e-12 like all others"""
    return name in CASE_BLIN or name in CASE_BBIL or name in CASE_FQ or name in CASE_QCOEF or name in CASE_BQ or name in CASE_FQT


def p2_node_partition(conn, e2n, part):
    """rank of every P2 node of a tetrahedral mesh for the vertex partition `part`: a vertex node goes with its vertex, an
    edge node with its end point of smaller global id (then every element around an owned node is local to its rank)"""
    nn = int(e2n.max()) + 1
    pn = np.empty(nn, np.int32)
    pn[e2n[:, :4]] = part[conn]
    for e, (p, r) in enumerate([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]):
        pn[e2n[:, 4 + e]] = part[np.minimum(conn[:, p], conn[:, r])]
    return pn


# ---------------------------------------------------------------------------------------------------------------
# FE functions as data of a form, shipped as dof arrays (fixtures fe*_data.npz).  An FE datum is
# (array name in the fixture, component, operator); name -> dict(base = the constant part as in CASES,
#   qcoef = [(datum, bilinear terms it multiplies)], lin = [(vcomp, vop, datum, scale)],
#   bbil = [(labels, datum, terms)], blin = [(labels, vcomp, datum, scale)])
# ---------------------------------------------------------------------------------------------------------------
FE_CASES = {
    "fe3d_p1_data": dict(base=(1, 1, [], [], "qfV5", [([1], 1, [0.0])]),
                         qcoef=[(("kap", 0, ID), LAP3), (("rho", 0, ID), [(0, ID, 0, ID, 1.0)])],
                         lin=[(0, ID, ("ff", 0, ID), 1.0), (0, DX, ("uk", 0, DX), 1.0), (0, DY, ("uk", 0, DY), 1.0), (0, DZ, ("uk", 0, DZ), 1.0)],
                         bbil=[([2, 3], ("kap", 0, ID), [(0, ID, 0, ID, 1.0)])], blin=[([2, 3], 0, ("ff", 0, ID), 1.0)]),
    "fe3d_p2_data": dict(base=(2, 1, [], [], "qfV5", [([1, 2], 1, [0.0])]),
                         qcoef=[(("kap", 0, ID), LAP3), (("m2", 0, ID), [(0, ID, 0, ID, 1.0)])],
                         lin=[(0, ID, ("uk", 0, ID), 1.0), (0, DX, ("uk", 0, DX), 1.0), (0, DY, ("uk", 0, DZ), 1.0)],
                         bbil=[], blin=[([6], 0, ("uk", 0, ID), 1.0)]),
    "fe2d_p1_data": dict(base=(1, 1, [(0, ID, 0, ID, 1.0)], [], "qf5pT", [([4], 1, [0.0])]),
                         qcoef=[(("kap", 0, ID), LAP2)],
                         lin=[(0, ID, ("ff", 0, ID), 1.0), (0, DX, ("uk", 0, DX), 1.0), (0, DY, ("uk", 0, DY), 1.0)],
                         bbil=[([2, 3], ("ff", 0, ID), [(0, ID, 0, ID, 1.0)])], blin=[([2], 0, ("uk", 0, DY), 1.0)]),
    "fe3d_lame_data": dict(base=(1, 3, [], [], "qfV5", [([1], 7, [0.0, 0.0, 0.0])]),
                           qcoef=[(("ee", 0, ID), lame_terms())],
                           lin=[(0, ID, ("f1", 0, ID), 1.0), (2, ID, ("f1", 2, ID), 1.0), (1, ID, ("f1", 1, DX), 1.0)],
                           bbil=[], blin=[]),
}


def fe_function(g, datum):
    """(order, element -> node table, dstride, doff, whole dof array) of the FE function behind a datum of FE_CASES: the space
    is recognised from its dofs per element, dof(k, c*nloc + a) = node*ncomp + c as everywhere in FreeFEM's Lagrange spaces"""
    name, comp, _ = datum
    dof, vals = g["fe_" + name + "_dof"], g["fe_" + name]
    dim = int(g["dim"])
    sizes = {1: (0, 1), dim + 1: (1, 1), (10 if dim == 3 else 6): (2, 1), 3 * (dim + 1): (1, 3), 3 * (10 if dim == 3 else 6): (2, 3)}
    order, ncomp = sizes[dof.shape[1]]
    nl = dof.shape[1] // ncomp
    e2n = np.ascontiguousarray(dof[:, :nl] // ncomp, dtype=np.int32)
    for c in range(ncomp):
        assert np.array_equal(dof[:, c * nl:(c + 1) * nl], e2n * ncomp + c)
    assert 0 <= comp < ncomp
    return order, e2n, ncomp, comp, vals
