"""The numpy oracle of the mass matrix / source vector path (oracle/fenris_oracle.py) pinned against the
reference's own tests: the analytic Quad4 mass matrix (tests/unit_tests/assembly/local.rs:38-69), the quadratic-form identity
of tests/unit_tests/assembly/local/mass.rs:15-118 and the inner-product identity of local/source.rs:19-160 (restated for the
elements of this scope), plus literal-vs-vectorised cross checks.  No GPU."""
import numpy as np
import pytest

from oracle import fenris_oracle as fo


def test_quad4_reference_element_mass_matrix_kat():
    X = np.array(fo._QUAD4_NODES)
    w, p = fo.quadrilateral_gauss(3)  # the reference uses total_order::quadrilateral(5): any rule exact for degree 4 per direction
    M = fo.element_mass_matrix(fo.QUAD4, X, w, p, [3.0] * len(w), 2)
    expected = np.kron(3.0 / 9.0 * np.array([[4, 2, 1, 2], [2, 4, 2, 1], [1, 2, 4, 2], [2, 1, 2, 4.0]]), np.eye(2))
    assert np.abs(M - expected).max() < 1e-14


@pytest.mark.parametrize("et", [fo.QUAD4, fo.TET4, fo.TET10, fo.HEX8, fo.HEX27])
def test_basis_is_a_nodal_partition_of_unity(et):
    n, _, d = fo.element_info(et)
    rng = np.random.default_rng(0)
    for xi in rng.uniform(-0.9, -0.1, size=(5, d)):  # inside every reference domain
        assert abs(fo.basis_values(et, xi).sum() - 1.0) < 1e-14
    nodes = {fo.QUAD4: fo._QUAD4_NODES, fo.HEX8: fo._HEX8_NODES, fo.HEX27: fo._HEX27_NODES}.get(et)
    if nodes is not None:  # phi_I(xi_J) = delta_IJ
        assert np.abs(np.array([fo.basis_values(et, x) for x in nodes]) - np.eye(n)).max() < 1e-14


def test_mass_quadratic_form_reproduces_weighted_l2_norm():
    # mass.rs:15-118 restated for a trilinear hex: with f in the element's space, f_h^T M f_h = int rho f^2; both sides by the same
    # (exact enough) Gauss rule, rho evaluated at the physical points
    X = np.array(fo._HEX8_NODES) * [1.0, 0.7, 1.3] + [2.0, 0.0, 1.0]
    w, p = fo.hexahedron_gauss(4)
    f = lambda x: 3.0 * x[..., 0] - x[..., 1] * x[..., 2] + 2.0 * x[..., 0] * x[..., 1] * x[..., 2] + 5.0  # trilinear
    rho = lambda x: (3.0 * x[..., 0] + 2.0 * x[..., 1] - 4.0 * x[..., 2] + 2.0) ** 2
    xq = np.array([fo.map_reference_coords(fo.HEX8, X, xi) for xi in p])
    M = fo.element_mass_matrix(fo.HEX8, X, w, p, rho(xq), 1)
    fh = f(X)
    detj = np.array([abs(fo.det_small(fo.reference_jacobian(fo.HEX8, X.T, xi))) for xi in p])
    assert abs(fh @ M @ fh - np.sum(w * detj * rho(xq) * f(xq) ** 2)) < 1e-10 * abs(fh @ M @ fh)


def test_source_vector_reproduces_inner_product():
    # source.rs:19-160: for u in the element's space, int f . u = u_K . f_K  (s = 2, a density-like parameter per point)
    a, b, c, d = np.array([2.0, 0, 1]), np.array([3.0, 4, 1]), np.array([1.0, 1, 2]), np.array([3.0, 1, 4])
    X4 = np.stack([a, b, c, d])
    X = np.concatenate([X4, [(X4[i] + X4[j]) / 2 for i, j in fo._TET10_EDGES]])  # Tet10 vertices (mesh_convert semantics)
    u = lambda x: np.stack([3 * x[..., 0] ** 2 - 4 * x[..., 0] * x[..., 1] + 3 * x[..., 0] * x[..., 2] - x[..., 2] ** 2 + 5,
                            3 * x[..., 0] + 3 * x[..., 1] * x[..., 2] - 2 * x[..., 1] + x[..., 2] * x[..., 1] - 3], axis=-1)
    f = lambda x: np.stack([6 * x[..., 0] ** 2 - 4 * x[..., 0] * x[..., 2] + 3 - x[..., 2] ** 2 - x[..., 0] + x[..., 1] - 3,
                            2 * x[..., 0] + 3 * x[..., 0] * (x[..., 1] - x[..., 2]) - x[..., 0] * x[..., 1] - 2 * x[..., 2] ** 2 + 5], axis=-1)
    from tests.mms import duffy_tet_rule
    w, p = duffy_tet_rule(5)
    xq = np.array([fo.map_reference_coords(fo.TET10, X, xi) for xi in p])
    dens = np.sum(xq ** 2, axis=1)  # the reference's artificial density |x|^2 (local.rs:72-78)
    fK = fo.element_source_vector(fo.TET10, X, w, p, dens[:, None] * f(xq))
    detj = abs(fo.det_small(fo.reference_jacobian(fo.TET10, X.T, p[0])))  # affine
    lhs = np.sum(w * detj * dens * np.sum(f(xq) * u(xq), axis=1))
    assert abs(lhs - u(X).reshape(-1) @ fK) < 1e-10 * abs(lhs)


@pytest.mark.parametrize("et,mesh,s", [(fo.QUAD4, lambda: fo.create_unit_square_uniform_quad_mesh_2d(4), 2),
                                      (fo.HEX8, lambda: fo.create_unit_box_uniform_hex_mesh_3d(2), 3),
                                      (fo.TET4, lambda: fo.create_unit_box_uniform_tet_mesh_3d(1), 1)])
def test_vectorised_variants_equal_the_literal_ones(et, mesh, s):
    v, c = mesh()
    v = fo.jitter_vertices(v, 0.25, amp=0.15)
    w, p = fo.quadrilateral_gauss(3) if et == fo.QUAD4 else (fo.hexahedron_gauss(3) if et == fo.HEX8 else fo.tetrahedron_rule(2))
    rho = np.linspace(1.0, 2.0, len(w))
    a, b = fo.assemble_mass_serial(et, v, c, w, p, rho, s), fo.assemble_mass_fast(et, v, c, w, p, rho, s)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and fo.rel_frobenius(b[2], a[2]) < 1e-14
    x = fo.physical_quadrature_points(et, v, c, p)
    f = np.stack([x[..., 0] ** 2, 1.0 + x[..., -1]], axis=-1)
    assert np.abs(fo.assemble_vector_serial(et, v, c, w, p, f) - fo.assemble_vector_fast(et, v, c, w, p, f)).max() < 1e-15
    # mass of the whole mesh: sum of all entries = s * int rho-weighted volume; for rho = 1 on the (jittered) unit domain
    one = fo.assemble_mass_fast(et, v, c, w, p, np.ones(len(w)), 1)[2].sum()
    vol = fo.assemble_vector_fast(et, v, c, w, p, np.ones((len(w), 1))).sum()
    assert abs(one - vol) < 1e-13


def test_dirichlet_oracle_small_known_answer():
    # 1-D chain of 4 nodes, s = 1, tridiagonal [2 -1; -1 2 -1; ...]; Dirichlet at node 0: row 0 -> (scale, 0), entry (1, 0) -> 0
    ro = np.array([0, 2, 5, 8, 10])
    ci = np.array([0, 1, 0, 1, 2, 1, 2, 3, 2, 3])
    vals = np.array([2.0, -1, -1, 2, -1, -1, 2, -1, -1, 2])
    scale = fo.apply_homogeneous_dirichlet_bc_csr(ro, ci, vals, [0], 1)
    assert scale == 2.0 and vals.tolist() == [2.0, 0, 0, 2, -1, -1, 2, -1, -1, 2]
    # first diagonal entry zero -> the next non-zero one gives the scale (skip_while, global.rs:393)
    vals = np.array([0.0, -1, -1, -3, -1, -1, 2, -1, -1, 2])
    assert fo.apply_homogeneous_dirichlet_bc_csr(ro, ci, vals, [3], 1) == 3.0 and vals.tolist() == [0.0, -1, -1, -3, -1, -1, 2, 0, 0, 3.0]
    rhs = np.arange(8.0)
    fo.apply_homogeneous_dirichlet_bc_rhs(rhs, [1, 3], 2)
    assert rhs.tolist() == [0, 1, 0, 0, 4, 5, 0, 0]


def test_conjugate_gradient_restatement():
    # the literal restatement of fenris-sparse/src/cg.rs:364-480 used as the checker of the device CG
    import scipy.sparse as sp
    n = 40
    A = sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(n, n)).tocsr()
    b = np.linspace(1.0, 2.0, n)
    x, its, status = fo.conjugate_gradient(lambda q: A @ q, b, rel_tol=1e-12)
    assert status == "ok" and its <= n and np.abs(A @ x - b).max() < 1e-10
    d = A.diagonal()
    xj, itsj, _ = fo.conjugate_gradient(lambda q: A @ q, b, rel_tol=1e-12, apply_p=lambda r: r / d)
    assert np.abs(xj - x).max() < 1e-10
    assert fo.conjugate_gradient(lambda q: A @ q, np.zeros(n), x0=np.ones(n))[1:] == (0, "ok")
    assert fo.conjugate_gradient(lambda q: A @ q, b, rel_tol=1e-14, max_iter=2)[2] == "max_iter"
    assert fo.conjugate_gradient(lambda q: -(A @ q), b)[2] == "indefinite_operator"


def test_hex20_serendipity_element_restatement():
    # hexahedron.rs:369-563: nodal basis (delta property, partition of unity), gradients = derivative of the basis, and the mesh
    # converter of mesh_convert.rs:168-217 (vertices + 12 edge midpoints, shared nodes merged)
    nodes = np.array(fo._HEX27_NODES[:20])
    assert np.abs(np.array([fo.basis_values(fo.HEX20, x) for x in nodes]) - np.eye(20)).max() == 0.0
    rng = np.random.default_rng(1)
    for xi in rng.uniform(-1, 1, size=(4, 3)):
        G = fo.reference_gradients(fo.HEX20, xi)
        h = 1e-6
        fd = np.stack([(fo.basis_values(fo.HEX20, xi + h * e) - fo.basis_values(fo.HEX20, xi - h * e)) / (2 * h) for e in np.eye(3)])
        assert abs(fo.basis_values(fo.HEX20, xi).sum() - 1.0) < 1e-14 and np.abs(G.sum(axis=1)).max() < 1e-14 and np.abs(G - fd).max() < 1e-8
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(3)
    v20, c20 = fo.hex20_mesh_from_hex8(v, c)
    assert c20.shape == (27, 20) and len(v20) == 4 ** 3 + 3 * 3 * 4 * 4  # vertices + one node per edge of the 3^3 grid
    assert np.array_equal(v20[c20[:, :8]], v[c])  # the first 8 nodes are the Hex8 vertices (the geometry, hexahedron.rs:552-554)
    # a quadratic field is reproduced exactly by the serendipity space: K u = 0 for u linear (stiffness annihilates constants + linears' curl-free part)
    prob = fo.Problem(fo.HEX20, v20, c20, fo.LAPLACE)
    ro, ci, vals = fo.assemble_fast(prob)
    import scipy.sparse as sp
    A = sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(len(v20),) * 2)
    assert np.abs(A @ np.ones(len(v20))).max() < 1e-12 and abs(v20[:, 0] @ (A @ v20[:, 0]) - 1.0) < 1e-12  # int |grad x|^2 = 1


def test_elliptic_vector_and_energy_restatement_tie_to_the_stiffness_matrix():
    # elliptic.rs:440-605 restated; for the linear operators f(u) = K u and psi(u) = u.K u / 2 with the K of elliptic.rs:361-439
    import scipy.sparse as sp
    mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
    for et, mesh, op in [(fo.HEX8, fo.create_unit_box_uniform_hex_mesh_3d(2), fo.LINEAR_ELASTIC),
                         (fo.QUAD4, fo.create_unit_square_uniform_quad_mesh_2d(3), fo.LAPLACE),
                         (fo.TET4, fo.create_unit_box_uniform_tet_mesh_3d(1), fo.LINEAR_ELASTIC)]:
        v, c = mesh
        v = fo.jitter_vertices(v, 0.3, amp=0.1)
        prob = fo.Problem(et, v, c.astype(np.int64), op, params=() if op == fo.LAPLACE else (mu, lam))
        ro, ci, vals = fo.assemble_fast(prob)
        n = len(ro) - 1
        A = sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(n, n))
        u = np.random.default_rng(0).normal(size=n)
        f, e = fo.assemble_elliptic_vector_serial(prob, u), fo.assemble_elliptic_scalar(prob, u)
        assert np.abs(f - A @ u).max() < 1e-13 * np.abs(f).max() and abs(e - 0.5 * u @ (A @ u)) < 1e-13 * abs(e)
    # the reference's energy known answers at a fixed deformation gradient (fenris-solid tests; materials KAT of tests/golden)
    F = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert abs(fo.elliptic_energy_density(fo.LINEAR_ELASTIC, (F - np.eye(2)).T, (384.0, 577.0)) - fo.linear_elastic_energy_density(F, 384.0, 577.0)) < 1e-9
