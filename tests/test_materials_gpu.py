"""GPU parity of the state-dependent elliptic operators (MaterialEllipticOperator<StVKMaterial> / <NeoHookeanMaterial>,
fenris-solid/src/materials.rs:232-469) against the oracle pinned in tests/test_materials_oracle.py: tangent stiffness matrix through fb200_assemble_into_csr with
u, element vector (internal forces) and energy, and the reference-API mirror."""
import numpy as np
import pytest

import fenris_b200 as fb
from oracle import fenris_oracle as fo

pytestmark = pytest.mark.gpu
TOL = 1e-12
MU, LAM = fo.lame_from_young_poisson(1e6, 0.2)


@pytest.fixture(scope="module")
def ctx():
    c = fb.Context(0)
    yield c
    c.close()


def _case(kind, n):
    if kind == "quad4":
        et, (v, c) = fo.QUAD4, fo.create_unit_square_uniform_quad_mesh_2d(n)
    elif kind == "tet4":
        et, (v, c) = fo.TET4, fo.create_unit_box_uniform_tet_mesh_3d(n)
    elif kind == "hex8":
        et, (v, c) = fo.HEX8, fo.create_unit_box_uniform_hex_mesh_3d(n)
    elif kind == "hex27":
        et, (v, c) = fo.HEX27, fo.hex27_mesh_from_hex8(*fo.create_unit_box_uniform_hex_mesh_3d(n))
    elif kind == "hex20":
        et, (v, c) = fo.HEX20, fo.hex20_mesh_from_hex8(*fo.create_unit_box_uniform_hex_mesh_3d(n))
    else:
        et, (v, c) = fo.TET10, fo.tet10_mesh_from_tet4(*fo.create_unit_box_uniform_tet_mesh_3d(n))
    if kind in ("quad4", "tet4", "hex8"):
        v = fo.jitter_vertices(v, 1.0 / n, amp=0.15)
    return et, np.ascontiguousarray(v), np.ascontiguousarray(c)


MATERIALS = [fo.STVK, fo.NEO_HOOKEAN]


@pytest.mark.parametrize("mat", MATERIALS)
@pytest.mark.parametrize("kind,n", [("quad4", 5), ("tet4", 2), ("hex8", 3), ("hex27", 1), ("tet10", 1), ("hex20", 2)])
def test_material_matrix_vector_energy_equal_oracle(ctx, kind, n, mat):
    et, v, c = _case(kind, n)
    prob = fo.Problem(et, v, c.astype(np.int64), mat, params=(MU, LAM))
    s = prob.sdim
    u = 0.05 * np.random.default_rng(21).normal(size=s * len(v))
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.assemble_pattern(s)
    ctx.color_nodes()
    oro, oci, ovals = fo.assemble_matrix_u_serial(et, v, c.astype(np.int64), mat, u, prob.weights, prob.points, prob.params_per_point)
    ro, ci = ctx.pattern_download()
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_COLORED):
        ctx.assemble_into_csr_device(mat, prob.weights, prob.points, (MU, LAM), scatter_mode=mode, u=u)
        ctx.synchronize()
        vals = ctx.values_download().copy()
        assert fo.rel_frobenius(vals, ovals) < TOL, (kind, mode)
    import scipy.sparse as sp
    A = sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(len(ro) - 1,) * 2)
    assert abs(A - A.T).max() == 0.0  # coloured: every block pair written from the same registers (clone_upper_to_lower, util.rs:38-50)
    # element vector (internal forces) and energy of the same operator
    ref = fo.assemble_elliptic_vector_serial(prob, u)
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_COLORED):
        f = ctx.assemble_elliptic_vector(mat, prob.weights, prob.points, (MU, LAM), u, scatter_mode=mode)
        assert np.abs(f - ref).max() < TOL * np.abs(ref).max()
    e = ctx.assemble_elliptic_scalar(mat, prob.weights, prob.points, (MU, LAM), u)
    oe = fo.assemble_elliptic_scalar(prob, u)
    assert abs(e - oe) < TOL * abs(oe)


@pytest.mark.parametrize("mat", MATERIALS)
def test_material_without_state_is_the_linear_elastic_matrix(ctx, mat):
    # F = I: the contraction reduces to LinearElasticMaterial's (materials.rs:110-121 vs :417-437); u = NULL means zeros
    et, v, c = _case("hex8", 4)
    prob = fo.Problem(et, v, c.astype(np.int64), fo.LINEAR_ELASTIC, params=(MU, LAM))
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.assemble_pattern(3)
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, prob.weights, prob.points, (MU, LAM))
    ctx.synchronize()
    lin = ctx.values_download().copy()
    ctx.assemble_into_csr_device(mat, prob.weights, prob.points, (MU, LAM))
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), lin) < TOL
    # accumulate semantics (global.rs:534)
    ctx.assemble_into_csr_device(mat, prob.weights, prob.points, (MU, LAM), accumulate=True)
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), 2.0 * lin) < TOL


@pytest.mark.parametrize("mat", MATERIALS)
def test_material_tangent_is_the_derivative_of_the_device_vector(ctx, mat):
    # size-independent property at a size the oracle does not touch: K(u) w = d/dt f(u + t w) on a 12^3 mesh, all on the device
    et, v, c = _case("hex8", 12)
    w8, p8 = fb.canonical_stiffness_quadrature(fb.HEX8)
    rng = np.random.default_rng(2)
    u = 0.003 * rng.normal(size=3 * len(v))  # small against the cell size 1 / 12: no inverted element (NeoHookean would give NaN)
    w = rng.normal(size=3 * len(v))
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.assemble_pattern(3)
    ctx.assemble_into_csr_device(mat, w8, p8, (2.0, 3.0), u=u)
    ctx.synchronize()
    Kw = ctx.spmv(w)
    h = 1e-6
    fp = ctx.assemble_elliptic_vector(mat, w8, p8, (2.0, 3.0), u + h * w).copy()
    fm = ctx.assemble_elliptic_vector(mat, w8, p8, (2.0, 3.0), u - h * w).copy()
    assert np.abs(Kw - (fp - fm) / (2 * h)).max() < 1e-7 * np.abs(Kw).max()


def test_stvk_singular_element_is_reported(ctx):
    cube = (np.array(fo._HEX8_NODES) + 1.0) / 2.0
    v = np.concatenate([cube + [2.0 * k, 0, 0] for k in range(4)])
    v[8:16] = 0.5
    c = np.arange(32, dtype=np.uint64).reshape(4, 8)
    ctx.space_upload(fb.HEX8, v, c)
    ctx.assemble_pattern(3)
    w8, p8 = fb.canonical_stiffness_quadrature(fb.HEX8)
    ctx.assemble_into_csr_device(fo.STVK, w8, p8, (MU, LAM), u=np.zeros(96))
    with pytest.raises(fb.SingularJacobianError) as ei:
        ctx.synchronize()
    assert ei.value.element_index == 1


def test_neo_hookean_inverted_element_gives_nan_and_infinite_energy(ctx):
    # det F <= 0: the reference returns NaN stress / contraction (materials.rs:277-279, 301-303) and +inf energy (:262-264), no error
    et, v, c = _case("hex8", 2)
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.assemble_pattern(3)
    w8, p8 = fb.canonical_stiffness_quadrature(fb.HEX8)
    u = (-2.0 * v).reshape(-1)  # F = I - 2 I = -I everywhere: det F = -1
    ctx.assemble_into_csr_device(fo.NEO_HOOKEAN, w8, p8, (MU, LAM), u=u)
    ctx.synchronize()
    assert np.isnan(ctx.values_download()).all()
    assert np.isnan(ctx.assemble_elliptic_vector(fo.NEO_HOOKEAN, w8, p8, (MU, LAM), u)).all()
    assert ctx.assemble_elliptic_scalar(fo.NEO_HOOKEAN, w8, p8, (MU, LAM), u) == np.inf


@pytest.mark.parametrize("material,mat", [(fb.StVKMaterial, fo.STVK), (fb.NeoHookeanMaterial, fo.NEO_HOOKEAN)])
def test_materials_through_the_reference_api_mirror(material, mat):
    m = fb.create_unit_square_uniform_quad_mesh_2d(4)
    w, p = fb.canonical_stiffness_quadrature(fb.QUAD4)
    qt = fb.UniformQuadratureTable.from_points_and_weights(p, w).with_uniform_data(fb.LameParameters(2.0, 3.0))
    u = 0.02 * np.random.default_rng(4).normal(size=2 * m.num_nodes())  # det F > 0 everywhere (checked with the oracle)
    ea = (fb.ElementEllipticAssemblerBuilder().with_finite_element_space(m).with_operator(fb.MaterialEllipticOperator(material()))
          .with_quadrature_table(qt).with_u(u).build())
    A = fb.CsrAssembler().assemble(ea)
    v, c = m.vertices(), m.connectivity().astype(np.int64)
    _, _, ovals = fo.assemble_matrix_u_serial(fo.QUAD4, v, c, mat, u, w, p, [(2.0, 3.0)] * len(w))
    assert fo.rel_frobenius(A.values, ovals) < TOL
    f = fb.VectorAssembler().assemble_vector(ea)
    prob = fo.Problem(fo.QUAD4, v, c, mat, weights=w, points=p, params=(2.0, 3.0))
    ref = fo.assemble_elliptic_vector_serial(prob, u)
    assert np.abs(f - ref).max() < TOL * np.abs(ref).max()
    assert abs(fb.assemble_scalar(ea) - fo.assemble_elliptic_scalar(prob, u)) < TOL * abs(fo.assemble_elliptic_scalar(prob, u))


@pytest.mark.parametrize("mat", MATERIALS)
@pytest.mark.parametrize("kind,n", [("quad4", 3), ("tet4", 2), ("hex8", 2), ("tet10", 1), ("hex20", 1), ("hex27", 1)])
def test_material_element_matrices_equal_oracle(ctx, kind, n, mat):
    # ElementMatrixAssembler::assemble_element_matrix (local.rs:78-80) for the state-dependent operators: dense K_e(u), no pattern needed
    et, v, c = _case(kind, n)
    prob = fo.Problem(et, v, c.astype(np.int64), mat, params=(MU, LAM))
    s = prob.sdim
    nn = c.shape[1]
    u = 0.05 * np.random.default_rng(5).normal(size=s * len(v))
    ctx.space_upload(et, v, c.astype(np.uint64))
    first, count = 1 if len(c) > 2 else 0, min(len(c), 7)
    count = min(count, len(c) - first)
    K = ctx.element_matrices(mat, prob.weights, prob.points, (MU, LAM), first, count, s * nn, u=u)
    for k in range(count):
        nodes = c[first + k].astype(np.int64)
        ue = u.reshape(-1, s)[nodes].ravel()
        ref = fo.element_matrix_u(et, v[nodes], mat, ue, prob.weights, prob.points, prob.params_per_point)
        assert fo.rel_frobenius(K[k], ref) < TOL, (kind, k)
        assert np.array_equal(K[k], K[k].T)  # upper triangle mirrored (clone_upper_to_lower, util.rs:38-50)
    # without a state: the linear-elastic element matrix (F = I)
    K0 = ctx.element_matrices(mat, prob.weights, prob.points, (MU, LAM), first, count, s * nn)
    Kl = ctx.element_matrices(fo.LINEAR_ELASTIC, prob.weights, prob.points, (MU, LAM), first, count, s * nn)
    assert fo.rel_frobenius(K0, Kl) < TOL
