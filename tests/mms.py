"""Method-of-manufactured-solutions harness restating the reference's Poisson convergence tests
(tests/convergence_tests/poisson_mms_common.rs:67-230, poisson_2d_mms.rs, poisson_3d_mms.rs) around a GIVEN stiffness
matrix: load vector, homogeneous Dirichlet boundary, solve, L2 / H1-seminorm error.  The stiffness matrix comes from
the oracle (CPU test) or from the GPU path (gpu test); the reference's stored errors
(tests/convergence_tests/reference_values/*.json, 1 % tolerance, poisson_mms_common.rs:40-65) pin it end to end.
Test infrastructure only."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fenris_oracle as fo


def basis_values(et, xi):
    if et == fo.QUAD4:
        return np.array([(1.0 + a * xi[0]) * (1.0 + b * xi[1]) / 4.0 for a, b in fo._QUAD4_NODES])  # quadrilateral.rs:79-91
    if et == fo.HEX8:
        return fo.hex8_basis(xi)
    if et == fo.TET4:
        return fo.tet4_basis(xi)
    raise NotImplementedError


def duffy_tet_rule(n):
    """Gauss^3 collapsed onto the reference tet (-1,-1,-1),(1,-1,-1),(-1,1,-1),(-1,-1,1); exact to degree 2n-3."""
    w1, x1 = fo.gauss(n)
    w1, x1 = np.array(w1), (np.array(x1) + 1.0) / 2.0  # [0,1]
    W, P = [], []
    for wa, a in zip(w1, x1):
        for wb, b in zip(w1, x1):
            for wc, c in zip(w1, x1):
                # unit simplex: u = a, v = b(1-a), t = c(1-a)(1-b); jac = (1-a)^2 (1-b); gauss weights on [0,1] carry 1/2 each
                u, v, t = a, b * (1 - a), c * (1 - a) * (1 - b)
                W.append(wa * wb * wc / 8.0 * (1 - a) ** 2 * (1 - b) * 8.0)  # x8: unit simplex -> reference tet (edge 2)
                P.append([2 * u - 1, 2 * v - 1, 2 * t - 1])
    return np.array(W), np.array(P)


def exact(d):
    if d == 2:
        u = lambda x: np.sin(np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1])
        gu = lambda x: np.pi * np.stack([np.cos(np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1]),
                                         np.sin(np.pi * x[..., 0]) * np.cos(np.pi * x[..., 1])], axis=-1)
        f = lambda x: 2.0 * np.pi ** 2 * u(x)
    else:
        s, c = np.sin, np.cos
        u = lambda x: s(np.pi * x[..., 0]) * s(np.pi * x[..., 1]) * s(np.pi * x[..., 2])
        gu = lambda x: np.pi * np.stack([c(np.pi * x[..., 0]) * s(np.pi * x[..., 1]) * s(np.pi * x[..., 2]),
                                         s(np.pi * x[..., 0]) * c(np.pi * x[..., 1]) * s(np.pi * x[..., 2]),
                                         s(np.pi * x[..., 0]) * s(np.pi * x[..., 1]) * c(np.pi * x[..., 2])], axis=-1)
        f = lambda x: 3.0 * np.pi ** 2 * u(x)
    return u, gu, f


def _per_point(et, verts, conn, xi):
    """x(xi), |det J|, physical gradients G[e, i, a] and basis N[a] for all elements at one reference point."""
    N = basis_values(et, xi)
    Gref = fo.reference_gradients(et, xi)  # d x n
    X = verts[conn]  # E, n, d
    xq = np.einsum("a,ead->ed", N, X)
    J = np.einsum("eai,ja->eij", X, Gref)
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)
    G = np.einsum("eji,ja->eia", Jinv, Gref)
    return xq, np.abs(det), G, N


def solve_poisson(et, verts, conn, K, quad_rule, error_rule):
    """K: scipy CSR stiffness (Laplace) on the mesh. Returns (L2 error, H1 seminorm error)."""
    d = verts.shape[1]
    u_ex, gu_ex, f = exact(d)
    conn = conn.astype(np.int64)
    nn = len(verts)
    b = np.zeros(nn)
    for w, xi in zip(*quad_rule):
        xq, adet, _, N = _per_point(et, verts, conn, xi)
        np.add.at(b, conn, (w * adet * f(xq))[:, None] * N[None, :])  # source.rs:217-278
    dirichlet = np.abs(verts - 0.5).max(axis=1) > 0.4999  # poisson_mms_common.rs:126-134
    free = np.nonzero(~dirichlet)[0]
    uh = np.zeros(nn)
    if len(free):
        Kff = K[free][:, free].tocsc()
        uh[free] = spla.spsolve(Kff, b[free])
    l2 = h1 = 0.0
    for w, xi in zip(*error_rule):
        xq, adet, G, N = _per_point(et, verts, conn, xi)
        ue = uh[conn]  # E, n
        l2 += np.sum(w * adet * (ue @ N - u_ex(xq)) ** 2)
        gh = np.einsum("eia,ea->ei", G, ue)
        h1 += np.sum(w * adet * np.sum((gh - gu_ex(xq)) ** 2, axis=1))
    return float(np.sqrt(l2)), float(np.sqrt(h1))


def csr_from(ro, ci, vals):
    n = len(ro) - 1
    return sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(n, n))


CASES = {
    # name: (element type, mesh producer, stiffness rule, error rule, golden key, resolutions used here)
    "quad4": (fo.QUAD4, fo.create_unit_square_uniform_quad_mesh_2d, lambda: fo.quadrilateral_gauss(2), lambda: fo.quadrilateral_gauss(6),
              "poisson2d_mms_quad4", [1, 2, 4, 8, 16, 32]),
    "hex8": (fo.HEX8, fo.create_unit_box_uniform_hex_mesh_3d, lambda: fo.hexahedron_gauss(2), lambda: fo.hexahedron_gauss(6),
             "poisson3d_mms_hex8", [1, 2, 4, 8, 16]),
    "tet4": (fo.TET4, fo.create_unit_box_uniform_tet_mesh_3d, lambda: fo.tetrahedron_rule(1), lambda: duffy_tet_rule(6),
             "poisson3d_mms_tet4", [1, 2, 4, 8]),
}
