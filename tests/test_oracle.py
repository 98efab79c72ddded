"""Pins the numpy oracle against the reference's own golden vectors
(tests/golden/reference_kats.json, extracted by tests/golden/make_golden.py) and
against analytic values.  CPU only."""
import math

import numpy as np
import pytest

from oracle import fenris_oracle as fo


# ------------------------------------------------------------------ pattern KATs
def test_pattern_kats_exact(kats):
    # reference: tests/unit_tests/assembly/global.rs:70-216 (serial == parallel)
    for case in kats["pattern"]:
        offs, cols = fo.assemble_pattern(case["sdim"], case["num_nodes"], case["elements"])
        assert offs.tolist() == case["offsets"]
        assert cols.tolist() == case["indices"]
        assert len(offs) == case["nrows"] + 1


def test_pattern_fast_equals_literal():
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(3)
    for sdim in (1, 3):
        a = fo.assemble_pattern(sdim, len(v), c.tolist())
        b = fo.assemble_pattern_fast(sdim, len(v), c)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(2)
    a = fo.assemble_pattern(3, len(v), c.tolist())
    b = fo.assemble_pattern_fast(3, len(v), c)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_pattern_counts_match_survey():
    # hex n^3 cells -> P = (3n+1)^3 ; BCC tets -> P = 30n^3+21n^2+9n+1
    for n in (2, 3, 5):
        v, c = fo.create_unit_box_uniform_hex_mesh_3d(n)
        offs, cols = fo.assemble_pattern_fast(1, len(v), c)
        assert int(offs[-1]) == (3 * n + 1) ** 3
        v, c = fo.create_unit_box_uniform_tet_mesh_3d(n)
        assert len(c) == 12 * n ** 3 and len(v) == (n + 1) ** 3 + n ** 3
        offs, cols = fo.assemble_pattern_fast(1, len(v), c)
        assert int(offs[-1]) == 30 * n ** 3 + 21 * n ** 2 + 9 * n + 1


# ------------------------------------------------------------------ meshes
@pytest.mark.parametrize("res", [1, 2])
def test_bcc_tet_mesh_snapshots_exact(kats, res):
    # reference: tests/unit_tests/mesh/procedural.rs:18-30 + insta snapshots mesh_1 / mesh_2
    snap = kats[f"bcc_tet_mesh_{res}"]
    v, c = fo.create_rectangular_uniform_tet_mesh(1.0, 1, 1, 1, res)
    assert v.tolist() == snap["vertices"]
    assert c.tolist() == snap["connectivity"]


def test_tet_mesh_positive_jacobians():
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(3)
    for nodes in c:
        X = v[nodes].T
        J = fo.reference_jacobian(fo.TET4, X, (-0.5, -0.5, -0.5))
        assert fo.det_small(J) > 0
    vol = sum(4.0 / 3.0 * fo.det_small(fo.reference_jacobian(fo.TET4, v[n].T, (0, 0, 0))) for n in c)
    assert abs(vol - 1.0) < 1e-13


def test_hex_mesh_layout():
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(2)
    assert v.shape == (27, 3) and c.shape == (8, 8)
    assert v[1].tolist() == [0.5, 0.0, 0.0] and v[3].tolist() == [0.0, 0.5, 0.0] and v[9].tolist() == [0.0, 0.0, 0.5]
    assert c[0].tolist() == [0, 1, 4, 3, 9, 10, 13, 12]
    for nodes in c:
        J = fo.reference_jacobian(fo.HEX8, v[nodes].T, (0.1, -0.2, 0.3))
        assert abs(fo.det_small(J) - 0.25 ** 3) < 1e-15


def test_quad_mesh_layout():
    v, c = fo.create_unit_square_uniform_quad_mesh_2d(2)
    assert v.shape == (9, 2) and c.shape == (4, 4)
    assert v[0].tolist() == [0.0, 1.0] and v[8].tolist() == [1.0, 0.0]
    # counter-clockwise from the lower-left corner
    assert c[0].tolist() == [3, 4, 1, 0]
    for nodes in c:
        J = fo.reference_jacobian(fo.QUAD4, v[nodes].T, (0.3, 0.1))
        assert fo.det_small(J) > 0


def test_hex27_single_element(kats):
    # reference: tests/unit_tests/fe_mesh.rs:60-129
    k = kats["hex27_single"]
    V = np.array(k["vertices"])
    v27, c27 = fo.hex27_mesh_from_hex8(V, np.array([[0, 1, 2, 3, 4, 5, 6, 7]]))
    assert c27.shape == (1, 27) and c27[0, :8].tolist() == list(range(8))
    assert c27[0].tolist() == list(range(27))
    tol = k["abstol"]
    for i in range(8):
        assert np.allclose(v27[i], V[i], atol=tol, rtol=0)
    for n, (a, b) in enumerate(k["edge_pairs"]):
        assert np.allclose(v27[8 + n], (V[a] + V[b]) / 2.0, atol=tol, rtol=0)
    for n, fs in enumerate(k["face_sets"]):
        assert np.allclose(v27[20 + n], V[fs].sum(axis=0) / len(fs), atol=tol, rtol=0)
    assert np.allclose(v27[26], V[k["center_set"]].sum(axis=0) / 8.0, atol=tol, rtol=0)


def test_hex27_mesh_counts_and_sharing():
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(2)
    v27, c27 = fo.hex27_mesh_from_hex8(v, c)
    assert len(v27) == 5 ** 3 and c27.shape == (8, 27)
    # every node position is unique => shared nodes were merged
    assert len({tuple(np.round(p, 12)) for p in v27.tolist()}) == len(v27)
    # node coordinates equal the trilinear image of the reference node positions
    for e in range(len(c27)):
        X8 = v27[c27[e, :8]]
        for l, ref in enumerate(fo._HEX27_NODES):
            assert np.allclose(X8.T @ fo.hex8_basis(ref), v27[c27[e, l]], atol=1e-14)


def test_tet10_mesh():
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(1)
    v10, c10 = fo.tet10_mesh_from_tet4(v, c)
    assert c10.shape == (12, 10)
    for e in range(12):
        for n, (a, b) in enumerate(fo._TET_EDGES):
            assert np.allclose(v10[c10[e, 4 + n]], 0.5 * (v10[c10[e, a]] + v10[c10[e, b]]))


# ------------------------------------------------------------------ quadrature
def test_gauss_points_order_and_exactness():
    w, p = fo.gauss(2)
    assert p[0] > 0 > p[1] and abs(p[0] - 1 / math.sqrt(3)) < 1e-15 and w == [1.0, 1.0] or abs(w[0] - 1) < 1e-15
    w, p = fo.gauss(3)
    assert abs(p[0] - math.sqrt(0.6)) < 1e-15 and abs(p[1]) < 1e-15 and abs(p[2] + math.sqrt(0.6)) < 1e-15
    assert np.allclose(w, [5 / 9, 8 / 9, 5 / 9], atol=1e-15)
    for n in range(1, 8):
        w, p = fo.gauss(n)
        for k in range(2 * n):
            exact = 0.0 if k % 2 else 2.0 / (k + 1)
            assert abs(sum(wi * pi ** k for wi, pi in zip(w, p)) - exact) < 1e-13


def test_tensor_rule_ordering():
    w, p = fo.hexahedron_gauss(2)
    g = 1 / math.sqrt(3)
    assert np.allclose(p[0], [g, g, g]) and np.allclose(p[1], [g, g, -g]) and np.allclose(p[4], [-g, g, g])
    assert abs(w.sum() - 8.0) < 1e-14
    w, p = fo.quadrilateral_gauss(2)
    assert np.allclose(p[1], [g, -g]) and abs(w.sum() - 4.0) < 1e-14


def test_tet_rules():
    for s in (1, 2):
        w, p = fo.tetrahedron_rule(s)
        assert abs(w.sum() - 4.0 / 3.0) < 1e-15
    w, p = fo.tetrahedron_rule(2)
    # strength 2 integrates (x+1)^2 exactly: x = 2u-1 maps the unit simplex (volume scale 8),
    # int_T0 u^2 = 2!/5! = 1/60  =>  int (x+1)^2 = 8 * 4 / 60
    val = float(np.sum(w * (p[:, 0] + 1.0) ** 2))
    assert abs(val - 32.0 / 60.0) < 1e-14


# ------------------------------------------------------------------ elements
@pytest.mark.parametrize("et,pts", [
    (fo.QUAD4, [(0.3, -0.2)]), (fo.TET4, [(-0.5, -0.4, -0.7)]), (fo.HEX8, [(0.3, -0.2, 0.7)]),
    (fo.HEX27, [(0.3, -0.2, 0.7)]), (fo.TET10, [(-0.5, -0.4, -0.7)]),
])
def test_gradients_partition_of_unity(et, pts):
    # reference: tests/unit_tests/element.rs partition-of-unity property
    for xi in pts:
        G = fo.reference_gradients(et, xi)
        assert np.allclose(G.sum(axis=1), 0.0, atol=1e-14)


def test_hex27_gradients_match_finite_differences():
    def basis(xi):
        return np.array([fo._phi_quad(a, xi[0]) * fo._phi_quad(b, xi[1]) * fo._phi_quad(c, xi[2])
                         for a, b, c in fo._HEX27_NODES])
    xi = np.array([0.31, -0.42, 0.13])
    G = fo.reference_gradients(fo.HEX27, xi)
    h = 1e-6
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd = (basis(xi + e) - basis(xi - e)) / (2 * h)
        assert np.allclose(G[d], fd, atol=1e-8)
    # Lagrange property
    for k, node in enumerate(fo._HEX27_NODES):
        b = basis(node)
        assert abs(b[k] - 1.0) < 1e-14 and np.allclose(np.delete(b, k), 0.0, atol=1e-14)


def test_small_matrix_closed_forms():
    rng = np.random.default_rng(0)
    for d in (2, 3):
        for _ in range(20):
            m = rng.normal(size=(d, d))
            assert abs(fo.det_small(m) - np.linalg.det(m)) < 1e-12
            assert np.allclose(fo.try_inverse_small(m), np.linalg.inv(m), atol=1e-10)
    assert fo.try_inverse_small(np.zeros((3, 3))) is None


# ------------------------------------------------------------------ materials
def test_lame_conversion(kats):
    # fenris-solid/tests/unit_tests/materials.rs:74-84 (comp = float)
    m = kats["materials"]
    mu, lam = fo.lame_from_young_poisson(m["young"], m["poisson"])
    assert abs(mu - m["lame_mu"]) <= 4 * np.spacing(m["lame_mu"])
    assert abs(lam - m["lame_lambda"]) <= 4 * np.spacing(m["lame_lambda"])


def test_linear_elastic_energy_goldens(kats):
    # fenris-solid/tests/unit_tests/materials.rs:245-262
    m = kats["materials"]
    assert fo.linear_elastic_energy_density(np.array(m["F2"]), m["mu"], m["lambda"]) == pytest.approx(m["psi_linear_2d"], rel=1e-15)
    assert fo.linear_elastic_energy_density(np.array(m["F3"]), m["mu"], m["lambda"]) == pytest.approx(m["psi_linear_3d"], rel=1e-15)


@pytest.mark.parametrize("dim", [2, 3])
def test_stress_is_derivative_of_energy_and_contraction_of_stress(kats, dim):
    # fenris-solid/tests/unit_tests/materials.rs macros :76-238 (FD, tol 1e-9*amax... we use analytic-friendly tol)
    m = kats["materials"]
    mu, lam = m["mu"], m["lambda"]
    F = np.array(m["F2"] if dim == 2 else m["F3"])
    h = 1e-5
    P = fo.linear_elastic_stress(F, mu, lam)
    Pfd = np.zeros_like(P)
    for i in range(dim):
        for j in range(dim):
            E = np.zeros_like(F)
            E[i, j] = h
            Pfd[i, j] = (fo.linear_elastic_energy_density(F + E, mu, lam) - fo.linear_elastic_energy_density(F - E, mu, lam)) / (2 * h)
    assert np.allclose(P, Pfd, atol=1e-5 * np.abs(P).max())
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=dim), rng.normal(size=dim)
    C = fo.contract(fo.LINEAR_ELASTIC, a, b, (mu, lam))
    # contraction[(i,j)] = sum_{k,l} a[k] dP_ik/dF_jl b[l]   (materials.rs:40-71)
    Cfd = np.zeros((dim, dim))
    for j in range(dim):
        for l in range(dim):
            E = np.zeros_like(F)
            E[j, l] = h
            dP = (fo.linear_elastic_stress(F + E, mu, lam) - fo.linear_elastic_stress(F - E, mu, lam)) / (2 * h)
            for i in range(dim):
                for k in range(dim):
                    Cfd[i, j] += a[k] * dP[i, k] * b[l]
    assert np.allclose(C, Cfd, atol=1e-9 * np.abs(C).max() + 1e-6)


# ------------------------------------------------------------------ element matrices
def test_quad4_reference_element_laplace(kats):
    # tests/unit_tests/assembly.rs:159-162
    expected = np.array(kats["quad4_laplace_reference_element"]["sixth_times"]) / 6.0
    X = np.array(fo._QUAD4_NODES)
    w, p = fo.quadrilateral_gauss(2)
    K = fo.element_matrix(fo.QUAD4, X, fo.LAPLACE, w, p, [()] * 4)
    assert np.allclose(K, expected, atol=1e-15)


def test_tet4_reference_element_laplace():
    X = np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=float)
    w, p = fo.tetrahedron_rule(1)
    K = fo.element_matrix(fo.TET4, X, fo.LAPLACE, w, p, [()])
    exp = np.array([[1, -1 / 3, -1 / 3, -1 / 3], [-1 / 3, 1 / 3, 0, 0], [-1 / 3, 0, 1 / 3, 0], [-1 / 3, 0, 0, 1 / 3]])
    assert np.allclose(K, exp, atol=1e-15)


def test_hex8_unit_cube_elasticity_values():
    # SURVEY A.8 analytic: K[0,0] = mu/3 + (lam+mu)/9, K[0,1] = (lam+mu)/12
    mu, lam = 384.0, 577.0
    X = (np.array(fo._HEX8_NODES) + 1.0) / 2.0
    w, p = fo.hexahedron_gauss(2)
    K = fo.element_matrix(fo.HEX8, X, fo.LINEAR_ELASTIC, w, p, [(mu, lam)] * 8)
    assert abs(K[0, 0] - (mu / 3 + (lam + mu) / 9)) < 1e-12
    assert abs(K[0, 1] - (lam + mu) / 12) < 1e-12 and abs(K[0, 2] - (lam + mu) / 12) < 1e-12
    assert np.allclose(K, K.T, atol=0)
    assert abs(np.linalg.norm(K) - 1722.593114184845) < 1e-9
    # rigid body modes
    for t in np.eye(3):
        assert np.abs(K @ np.tile(t, 8)).max() < 1e-11
    for ax in range(3):
        w_ = np.zeros(3)
        w_[ax] = 1.0
        rot = np.concatenate([np.cross(w_, x) for x in X])
        assert np.abs(K @ rot).max() < 1e-11


def test_canonical_rule_matches_high_order_rule():
    # tests/unit_tests/quadrature/canonical.rs:41-99 (<= 64 ulp-ish): canonical == richer rule on the reference element
    for et, ref_nodes, hi in ((fo.HEX8, fo._HEX8_NODES, fo.hexahedron_gauss(5)),
                              (fo.HEX27, fo._HEX27_NODES, fo.hexahedron_gauss(6)),
                              (fo.QUAD4, fo._QUAD4_NODES, fo.quadrilateral_gauss(5))):
        X = np.array(ref_nodes, dtype=float)
        w, p = fo.canonical_stiffness_rule(et)
        K1 = fo.element_matrix(et, X, fo.LAPLACE, w, p, [()] * len(w))
        K2 = fo.element_matrix(et, X, fo.LAPLACE, hi[0], hi[1], [()] * len(hi[0]))
        assert np.allclose(K1, K2, atol=1e-13 * np.abs(K2).max() * 64)


def test_singular_jacobian_raises():
    X = np.zeros((8, 3))
    w, p = fo.hexahedron_gauss(2)
    with pytest.raises(fo.SingularJacobian):
        fo.element_matrix(fo.HEX8, X, fo.LAPLACE, w, p, [()] * 8)


def test_element_matrices_fast_equals_literal():
    rng = np.random.default_rng(7)
    mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
    cases = [(fo.HEX8, fo.create_unit_box_uniform_hex_mesh_3d(2)), (fo.TET4, fo.create_unit_box_uniform_tet_mesh_3d(1)),
             (fo.QUAD4, fo.create_unit_square_uniform_quad_mesh_2d(3))]
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(1)
    cases.append((fo.HEX27, fo.hex27_mesh_from_hex8(v, c)))
    for et, (v, c) in cases:
        v = v + rng.uniform(-0.03, 0.03, size=v.shape)
        for op in (fo.LAPLACE, fo.LINEAR_ELASTIC):
            prob = fo.Problem(et, v, c, op, params=(mu, lam) if op == fo.LINEAR_ELASTIC else ())
            Kf = fo.element_matrices_fast(prob)
            for e in range(min(3, len(c))):
                Kl = prob.element_matrix(e)
                assert fo.rel_frobenius(Kf[e], Kl) < 1e-14


# ------------------------------------------------------------------ global assembly
def _problems():
    mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
    v, c = fo.create_unit_square_uniform_quad_mesh_2d(4)
    yield "quad4-laplace", fo.Problem(fo.QUAD4, v, c, fo.LAPLACE)
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(2)
    yield "tet4-laplace", fo.Problem(fo.TET4, v, c, fo.LAPLACE)
    yield "tet4-elastic", fo.Problem(fo.TET4, v, c, fo.LINEAR_ELASTIC, params=(mu, lam))
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(3)
    yield "hex8-elastic", fo.Problem(fo.HEX8, v, c, fo.LINEAR_ELASTIC, params=(mu, lam))
    yield "hex8-laplace", fo.Problem(fo.HEX8, v, c, fo.LAPLACE)


@pytest.mark.parametrize("name,prob", list(_problems()))
def test_serial_equals_colored_equals_fast(name, prob):
    # reference's own check: par == serial, tests/convergence_tests/poisson_mms_common.rs:117-121
    ro, ci, vals = fo.assemble_serial(prob)
    colors = fo.sequential_greedy_coloring(prob.connectivity.tolist())
    ro2, ci2, vals2 = fo.assemble_colored(prob, colors)
    assert np.array_equal(ro, ro2) and np.array_equal(ci, ci2)
    assert fo.rel_frobenius(vals2, vals) < 1e-15
    ro3, ci3, vals3 = fo.assemble_fast(prob)
    assert np.array_equal(ro, ro3) and np.array_equal(ci, ci3)
    assert fo.rel_frobenius(vals3, vals) < 1e-14


def test_global_elasticity_invariants():
    import scipy.sparse as sp
    mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(4)
    v = fo.jitter_vertices(v, 0.25, amp=0.15)
    prob = fo.Problem(fo.HEX8, v, c, fo.LINEAR_ELASTIC, params=(mu, lam))
    ro, ci, vals = fo.assemble_fast(prob)
    n = 3 * len(v)
    A = sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(n, n))
    assert abs(A - A.T).max() < 1e-9 * abs(A).max()
    scale = abs(A).max()
    for t in np.eye(3):
        assert np.abs(A @ np.tile(t, len(v))).max() < 1e-10 * scale
    rot = np.cross(np.array([0.0, 0.0, 1.0]), v).ravel()
    assert np.abs(A @ rot).max() < 1e-10 * scale


def test_accumulate_semantics():
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(1)
    prob = fo.Problem(fo.TET4, v, c, fo.LAPLACE)
    ro, ci, vals = fo.assemble_serial(prob)
    ro, ci, vals2 = fo.assemble_serial(prob, pattern=(ro, ci), values=vals.copy())
    assert np.allclose(vals2, 2 * vals, rtol=1e-15)


def test_scatter_missing_column_raises():
    with pytest.raises(IndexError):
        fo.add_element_row_to_csr_row(np.zeros(2), np.array([0, 1]), [0, 5], [0, 1], 1, np.ones(2))


# ------------------------------------------------------------------ colouring
def test_greedy_coloring_invariants_and_structured_counts():
    # fenris-paradis/src/coloring.rs:83-108
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(4)
    colors = fo.sequential_greedy_coloring(c.tolist())
    assert len(colors) == 8 and all(len(col) == 8 for col in colors)
    seen = sorted(e for col in colors for e in col)
    assert seen == list(range(len(c)))
    for col in colors:
        nodes = np.concatenate([c[e] for e in col])
        assert len(np.unique(nodes)) == len(nodes)
    # first colour in element order: cells with even (i, j, k)
    assert colors[0] == [0, 2, 8, 10, 32, 34, 40, 42]
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(3)
    colors = fo.sequential_greedy_coloring(c.tolist())
    for col in colors:
        nodes = np.concatenate([c[e] for e in col])
        assert len(np.unique(nodes)) == len(nodes)
    assert sorted(e for col in colors for e in col) == list(range(len(c)))


def test_coloring_mock_with_repeats_and_empty():
    colors = fo.sequential_greedy_coloring([[0, 1, 2], [2, 3], [], [3, 4]])
    assert colors == [[0, 2, 3], [1]]
