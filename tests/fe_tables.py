"""numpy restatement of what FreeFEM's interpreter returns when a P0 / P1 / P2 Lagrange FE function is used as data of a form:
pfer2R / pf3r2R (fflib/lgfem.cpp:2053-2088) -> FElement::operator()(PHat, u, comp, op) (femlib/FESpace.cpp:1078-1099,
:1637-1654, femlib/P012_3d.cpp:98-122): sum_a u[K(a)] d^op phi_a(PHat) in the element that holds the quadrature node.  TEST
INFRASTRUCTURE (the checker of ffcuda_fe_table): pinned on the fixtures tests/golden/fe*_data.npz dumped from the reference
(tests/test_fe_tables.py, CPU suite); never imported by the product."""
import numpy as np

import oracle_lib as ol

ID, DX, DY, DZ = 0, 1, 2, 6
_SLOT = {ID: 0, DX: 1, DY: 2, DZ: 3}
_EDGE3 = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]      # femlib/P012_3d.cpp: node 4+e on edge e (Element::nvedge)
_EDGE2 = [(1, 2), (0, 2), (0, 1)]                              # 2-D: node 3+e on the edge opposite vertex e
_NVFACE = np.array([[3, 2, 1], [0, 2, 3], [3, 1, 0], [0, 1, 2]])
_NVEDGE = np.array([[1, 2], [2, 0], [0, 1]])


def warped_mesh(dim):
    """a small mesh with two region labels and curved coordinate lines (nothing axis-aligned, no two elements congruent)"""
    if dim == 3:
        g = ol.cube(4, 3, 3)
        x, y, z = g["xyz"].T.copy()
        g["xyz"] = np.ascontiguousarray(np.stack([x + 0.1 * y * y, y + 0.05 * z, z * (1 + 0.2 * x)], axis=1))
    else:
        g = ol.square(6, 5)
        x, y = g["xyz"].T.copy()
        g["xyz"] = np.ascontiguousarray(np.stack([x + 0.2 * y * y, y * (1 + 0.3 * x)], axis=1))
    c = g["xyz"][g["conn"]].mean(axis=1)
    g["elab"] = np.where(c[:, 0] > 0.55, 1, 0).astype(np.int32)
    return g


def node_table(g, order):
    """(element -> node table or None, number of nodes) as FreeFEM numbers a scalar space of that order on the mesh
    (2-D P2 is not needed by the tests that call this)"""
    nt, nv = g["conn"].shape[0], g["xyz"].shape[0]
    if order == 0:
        return None, nt
    if order == 1:
        return None, nv
    assert g["dim"] == 3
    return ol.p2_nodes_3d(nv, g["conn"])


def bary_gradients(xyz, conn):
    """(nt, dim+1, dim): gradient of every barycentric coordinate on every element"""
    X = xyz[conn]                                                  # (nt, d+1, d)
    nt, nvk, d = X.shape
    M = np.concatenate([np.ones((nt, nvk, 1)), X], axis=2)         # rows (1, x_a): M @ (c0, g) = e_a
    return np.transpose(np.linalg.inv(M)[:, 1:, :], (0, 2, 1))     # inv[:, 1+x, a] = d lambda_a / d x


def basis(dim, order, lam, grads=None):
    """values (grads None) or physical gradients of the local basis functions: lam (..., dim+1) barycentric coordinates of the
    points, grads (..., dim+1, dim) gradients of the barycentric coordinates -> (..., nloc) or (..., nloc, dim)"""
    nvk = dim + 1
    edges = _EDGE3 if dim == 3 else _EDGE2
    if order == 0:
        one = np.ones(lam.shape[:-1] + (1,))
        return one if grads is None else np.zeros(lam.shape[:-1] + (1, dim))
    if order == 1:
        return lam if grads is None else np.broadcast_to(grads, lam.shape[:-1] + (nvk, dim))
    if grads is None:
        v = [lam[..., a] * (2 * lam[..., a] - 1) for a in range(nvk)] + [4 * lam[..., i] * lam[..., j] for i, j in edges]
        return np.stack(v, axis=-1)
    v = [(4 * lam[..., a] - 1)[..., None] * grads[..., a, :] for a in range(nvk)]
    v += [4 * (lam[..., i, None] * grads[..., j, :] + lam[..., j, None] * grads[..., i, :]) for i, j in edges]
    return np.stack(v, axis=-2)


def fe_values(g, order, e2n, u, qpts, op, border=False):
    """(nunits, nq): d^op f at the quadrature nodes of every element (border: at the face nodes of every boundary element, in
    the adjacent element); u = dofs of the scalar function, e2n = its node table (None: element / vertex numbering)"""
    dim = g["dim"]
    xyz, conn = np.asarray(g["xyz"], float), np.asarray(g["conn"])
    qpts = np.asarray(qpts, float).reshape(-1, dim - 1 if border else dim)
    nq = qpts.shape[0]
    lq = np.concatenate([1.0 - qpts.sum(axis=1, keepdims=True), qpts], axis=1)
    if border:
        belem, bface = np.asarray(g["belem"]), np.asarray(g["bface"])
        fv = (_NVFACE if dim == 3 else _NVEDGE)[bface]             # local vertices of the face, PBord's order
        lam = np.zeros((len(belem), nq, dim + 1))
        for j in range(dim):
            np.put_along_axis(lam, np.broadcast_to(fv[:, None, j:j + 1], (len(belem), nq, 1)), lq[None, :, j:j + 1], axis=2)
        elems = belem
    else:
        elems = np.arange(conn.shape[0])
        lam = np.broadcast_to(lq[None], (len(elems), nq, dim + 1))
    if e2n is None:
        nodes = elems[:, None] if order == 0 else conn[elems]
    else:
        nodes = np.asarray(e2n)[elems]
    uk = np.asarray(u, float)[nodes]                               # (nunits, nloc)
    if op == ID:
        return np.einsum("uqa,ua->uq", basis(dim, order, lam), uk)
    G = bary_gradients(xyz, conn)[elems]                           # (nunits, d+1, d)
    dphi = basis(dim, order, lam, np.broadcast_to(G[:, None], lam.shape[:2] + G.shape[1:]))
    return np.einsum("uqa,ua->uq", dphi[..., _SLOT[op] - 1], uk)
