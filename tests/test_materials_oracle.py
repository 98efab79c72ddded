"""Pins the StVK restatement (oracle/fenris_oracle.py, SURVEY 8(f) rank 4 - a non-linear material) to the reference's own tests:
golden energies of fenris-solid/tests/unit_tests/materials.rs:299-315 (fixtures of tests/unit_tests/mod.rs:11-29), stress = dpsi/dF and
contraction = a.dP/dF.b by central differences (materials.rs:10-120, 317-330), and the assembled tangent K(u) = df/du."""
import json
import os

import numpy as np
import pytest

from oracle import fenris_oracle as fo

# the reference's own fixtures and golden values, extracted by tests/golden/make_golden.py
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["materials"]
MU, LAM = KATS["mu"], KATS["lambda"]  # lame_parameters(), mod.rs:11-16
F2 = np.array(KATS["F2"])  # deformation_gradient_2d, mod.rs:18-22
F3 = np.array(KATS["F3"])  # deformation_gradient_3d, mod.rs:24-29


def test_stvk_golden_strain_energies():
    assert KATS["psi_stvk_2d"] == 132578.0 and KATS["psi_stvk_3d"] == 9136789.125
    assert fo.stvk_energy_density(F2, MU, LAM) == KATS["psi_stvk_2d"]  # materials.rs:300-306
    assert fo.stvk_energy_density(F3, MU, LAM) == KATS["psi_stvk_3d"]  # materials.rs:309-315


@pytest.mark.parametrize("F", [F2, F3])
def test_stvk_stress_and_contraction_by_finite_differences(F):
    d, h = F.shape[0], 1e-5
    P = fo.stvk_stress(F, MU, LAM)
    Pfd = np.zeros((d, d))
    for i in range(d):
        for j in range(d):
            D = np.zeros((d, d))
            D[i, j] = h
            Pfd[i, j] = (fo.stvk_energy_density(F + D, MU, LAM) - fo.stvk_energy_density(F - D, MU, LAM)) / (2 * h)
    assert np.abs(P - Pfd).max() < 1e-6 * np.abs(P).max()
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=d), rng.normal(size=d)
    # contraction C_ij = a_k dP_ik / dF_jm b_m  (HyperelasticMaterial::compute_stress_contraction, fenris-solid lib.rs)
    Cfd = np.zeros((d, d))
    for j in range(d):
        for m in range(d):
            D = np.zeros((d, d))
            D[j, m] = h
            dP = (fo.stvk_stress(F + D, MU, LAM) - fo.stvk_stress(F - D, MU, LAM)) / (2 * h)
            Cfd[:, j] += (dP @ a) * b[m]
    C = fo.stvk_contraction(F, a, b, MU, LAM)
    assert np.abs(C - Cfd).max() < 1e-6 * np.abs(C).max()
    # contraction(b, a) = contraction(a, b)^T: what clone_upper_to_lower relies on (elliptic.rs:433-436)
    assert np.abs(fo.stvk_contraction(F, b, a, MU, LAM) - C.T).max() < 1e-12 * np.abs(C).max()


def test_stvk_at_the_reference_state_is_linear_elasticity():
    rng = np.random.default_rng(5)
    for d in (2, 3):
        a, b = rng.normal(size=d), rng.normal(size=d)
        assert np.allclose(fo.stvk_contraction(np.eye(d), a, b, MU, LAM), fo.contract(fo.LINEAR_ELASTIC, a, b, (MU, LAM)), rtol=1e-14, atol=1e-12)
    v, c = fo.create_unit_square_uniform_quad_mesh_2d(3)
    prob = fo.Problem(fo.QUAD4, v, c, fo.STVK, params=(MU, LAM))
    lin = fo.Problem(fo.QUAD4, v, c, fo.LINEAR_ELASTIC, params=(MU, LAM))
    _, _, k0 = fo.assemble_matrix_u_serial(fo.QUAD4, v, c, fo.STVK, np.zeros(2 * len(v)), prob.weights, prob.points, prob.params_per_point)
    _, _, kl = fo.assemble_serial(lin)
    assert fo.rel_frobenius(k0, kl) < 1e-14


@pytest.mark.parametrize("et,mesh", [(fo.QUAD4, "quad"), (fo.HEX8, "hex"), (fo.TET4, "tet")])
def test_stvk_tangent_is_the_derivative_of_the_element_vector(et, mesh):
    # the commented-out reference tests (tests/unit_tests/assembly.rs:385-470) check exactly this with h = 1e-6
    v, c = {"quad": fo.create_unit_square_uniform_quad_mesh_2d, "hex": fo.create_unit_box_uniform_hex_mesh_3d,
            "tet": fo.create_unit_box_uniform_tet_mesh_3d}[mesh](1)
    prob = fo.Problem(et, v, c, fo.STVK, params=(2.0, 3.0))
    n, _, d = fo.element_info(et)
    u = 0.2 * np.random.default_rng(9).normal(size=n * d)
    X = v[c[0]]
    K = fo.element_matrix_u(et, X, fo.STVK, u, prob.weights, prob.points, prob.params_per_point)
    Kfd = np.zeros_like(K)
    h = 1e-6
    for k in range(n * d):
        D = np.zeros(n * d)
        D[k] = h
        Kfd[:, k] = (fo.element_elliptic_vector(et, X, fo.STVK, u + D, prob.weights, prob.points, prob.params_per_point)
                     - fo.element_elliptic_vector(et, X, fo.STVK, u - D, prob.weights, prob.points, prob.params_per_point)) / (2 * h)
    assert np.abs(K - Kfd).max() < 1e-7 * np.abs(K).max()
    assert np.array_equal(K, K.T)
    # the element vector is the derivative of the element energy
    f = fo.element_elliptic_vector(et, X, fo.STVK, u, prob.weights, prob.points, prob.params_per_point)
    ffd = np.zeros_like(f)
    for k in range(n * d):
        D = np.zeros(n * d)
        D[k] = h
        ffd[k] = (fo.element_elliptic_energy(et, X, fo.STVK, u + D, prob.weights, prob.points, prob.params_per_point)
                  - fo.element_elliptic_energy(et, X, fo.STVK, u - D, prob.weights, prob.points, prob.params_per_point)) / (2 * h)
    assert np.abs(f - ffd).max() < 1e-7 * np.abs(f).max()


# ---- NeoHookeanMaterial (fenris-solid/src/materials.rs:232-353), same test method (materials.rs:335-375 of the reference's unit tests)
def test_neo_hookean_golden_strain_energies():
    # compute_energy_density(F) = compute_energy_density_du((F - I)^T)  (materials.rs:246-249)
    assert abs(fo.neo_hookean_energy_density_du((F2 - np.eye(2)).T, MU, LAM) - KATS["psi_neo_hookean_2d"]) < 1e-12 * 5505.0
    assert abs(fo.neo_hookean_energy_density_du((F3 - np.eye(3)).T, MU, LAM) - KATS["psi_neo_hookean_3d"]) < 1e-12 * 48833.0
    assert fo.neo_hookean_energy_density_du(-2.0 * np.eye(3), MU, LAM) == np.inf  # det F <= 0 (materials.rs:262-264)
    assert np.isnan(fo.neo_hookean_stress(-np.eye(3), MU, LAM)).all()  # materials.rs:277-279


@pytest.mark.parametrize("F", [F2, F3])
def test_neo_hookean_stress_and_contraction_by_finite_differences(F):
    d, h = F.shape[0], 1e-5
    energy = lambda G: fo.neo_hookean_energy_density_du((G - np.eye(d)).T, MU, LAM)
    P = fo.neo_hookean_stress(F, MU, LAM)
    Pfd = np.zeros((d, d))
    for i in range(d):
        for j in range(d):
            D = np.zeros((d, d))
            D[i, j] = h
            Pfd[i, j] = (energy(F + D) - energy(F - D)) / (2 * h)
    assert np.abs(P - Pfd).max() < 1e-6 * np.abs(P).max()
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=d), rng.normal(size=d)
    Cfd = np.zeros((d, d))
    for j in range(d):
        for m in range(d):
            D = np.zeros((d, d))
            D[j, m] = h
            dP = (fo.neo_hookean_stress(F + D, MU, LAM) - fo.neo_hookean_stress(F - D, MU, LAM)) / (2 * h)
            Cfd[:, j] += (dP @ a) * b[m]
    C = fo.neo_hookean_contraction(F, a, b, MU, LAM)
    assert np.abs(C - Cfd).max() < 1e-6 * np.abs(C).max()
    assert np.abs(fo.neo_hookean_contraction(F, b, a, MU, LAM) - C.T).max() < 1e-12 * np.abs(C).max()


def test_neo_hookean_at_the_reference_state_is_linear_elasticity():
    rng = np.random.default_rng(5)
    for d in (2, 3):
        a, b = rng.normal(size=d), rng.normal(size=d)
        assert np.allclose(fo.neo_hookean_contraction(np.eye(d), a, b, MU, LAM), fo.contract(fo.LINEAR_ELASTIC, a, b, (MU, LAM)), rtol=1e-14, atol=1e-12)
    # log det through log1p keeps the small-strain energy accurate where log(det F) would cancel (logdet.rs:57-86)
    U = 1e-9 * np.array([[1.0, 2.0, 0.5], [0.3, -1.0, 0.7], [0.2, 0.1, 0.4]])
    assert abs(fo.log_det_F(U) - np.trace(U)) < 1e-17


@pytest.mark.parametrize("et,mesh", [(fo.QUAD4, "quad"), (fo.HEX8, "hex")])
def test_neo_hookean_tangent_is_the_derivative_of_the_element_vector(et, mesh):
    v, c = {"quad": fo.create_unit_square_uniform_quad_mesh_2d, "hex": fo.create_unit_box_uniform_hex_mesh_3d}[mesh](1)
    prob = fo.Problem(et, v, c, fo.NEO_HOOKEAN, params=(2.0, 3.0))
    n, _, d = fo.element_info(et)
    u = 0.1 * np.random.default_rng(9).normal(size=n * d)
    X = v[c[0]]
    args = (prob.weights, prob.points, prob.params_per_point)
    K = fo.element_matrix_u(et, X, fo.NEO_HOOKEAN, u, *args)
    Kfd, ffd = np.zeros_like(K), np.zeros(n * d)
    h = 1e-6
    for k in range(n * d):
        D = np.zeros(n * d)
        D[k] = h
        Kfd[:, k] = (fo.element_elliptic_vector(et, X, fo.NEO_HOOKEAN, u + D, *args) - fo.element_elliptic_vector(et, X, fo.NEO_HOOKEAN, u - D, *args)) / (2 * h)
        ffd[k] = (fo.element_elliptic_energy(et, X, fo.NEO_HOOKEAN, u + D, *args) - fo.element_elliptic_energy(et, X, fo.NEO_HOOKEAN, u - D, *args)) / (2 * h)
    assert np.abs(K - Kfd).max() < 1e-7 * np.abs(K).max() and np.array_equal(K, K.T)
    f = fo.element_elliptic_vector(et, X, fo.NEO_HOOKEAN, u, *args)
    assert np.abs(f - ffd).max() < 1e-7 * np.abs(f).max()


def _rotation(d, rng):
    q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


@pytest.mark.parametrize("d", [2, 3])
def test_hyperelastic_materials_are_frame_indifferent(d):
    # psi(Q F) = psi(F) and P(Q F) = Q P(F) for rotations Q: holds for StVK and NeoHookean, not for the linear material
    rng = np.random.default_rng(31)
    F = np.eye(d) + 0.2 * rng.normal(size=(d, d))
    Q = _rotation(d, rng)
    for energy, stress in ((lambda G: fo.stvk_energy_density(G, MU, LAM), lambda G: fo.stvk_stress(G, MU, LAM)),
                           (lambda G: fo.neo_hookean_energy_density_du((G - np.eye(d)).T, MU, LAM), lambda G: fo.neo_hookean_stress(G, MU, LAM))):
        assert abs(energy(Q @ F) - energy(F)) < 1e-11 * abs(energy(F))
        assert np.abs(stress(Q @ F) - Q @ stress(F)).max() < 1e-11 * np.abs(stress(F)).max()


@pytest.mark.parametrize("mat", [fo.STVK, fo.NEO_HOOKEAN])
def test_rigid_rotation_produces_no_internal_forces_and_no_energy(mat):
    # u(X) = (Q - I) X + t on a jittered Hex8 mesh: F = Q exactly for the trilinear element, so E = 0, P = 0, psi = 0
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(2)
    v = fo.jitter_vertices(v, 0.5, amp=0.15)
    rng = np.random.default_rng(8)
    Q = _rotation(3, rng)
    u = (v @ (Q - np.eye(3)).T + np.array([0.3, -0.2, 0.1])).reshape(-1)
    prob = fo.Problem(fo.HEX8, v, c, mat, params=(MU, LAM))
    f = fo.assemble_elliptic_vector_serial(prob, u)
    scale = MU * np.abs(u).max()
    assert np.abs(f).max() < 1e-11 * scale
    assert abs(fo.assemble_elliptic_scalar(prob, u)) < 1e-11 * scale
    # the linear material does see a finite rotation
    lin = fo.Problem(fo.HEX8, v, c, fo.LINEAR_ELASTIC, params=(MU, LAM))
    assert np.abs(fo.assemble_elliptic_vector_serial(lin, u)).max() > 1e-3 * scale
    # and the tangent at the rotated state has the infinitesimal rigid motions of the CURRENT configuration in its null space
    _, _, k = fo.assemble_matrix_u_serial(fo.HEX8, v, c, mat, u, prob.weights, prob.points, prob.params_per_point)
    ro, ci = fo.assemble_pattern(3, len(v), c.tolist())
    import scipy.sparse as sp
    K = sp.csr_matrix((k, ci, ro), shape=(3 * len(v),) * 2)
    x = v @ Q.T  # current positions (up to the translation)
    W = np.array([[0.0, -1.0, 0.5], [1.0, 0.0, -0.3], [-0.5, 0.3, 0.0]])  # skew: w(x) = W x
    w = (x @ W.T).reshape(-1)
    assert np.abs(K @ w).max() < 1e-10 * np.abs(k).max() * np.abs(w).max()
    assert np.abs(K @ np.tile([1.0, 2.0, 3.0], len(v))).max() < 1e-10 * np.abs(k).max()
