"""GPU parity of the elliptic operator's element VECTOR and ENERGY (ElementEllipticAssembler as ElementVectorAssembler /
ElementScalarAssembler, src/assembly/local/elliptic.rs:342-359, 440-605 - row a7 of SURVEY 8, grad u, in use) against the literal
oracle, plus the identities f(u) = K u and psi(u) = u.K u / 2 that tie them to the assembled stiffness matrix."""
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
    else:
        et, (v, c) = fo.TET10, fo.tet10_mesh_from_tet4(*fo.create_unit_box_uniform_tet_mesh_3d(n))
    if kind in ("quad4", "tet4", "hex8"):
        v = fo.jitter_vertices(v, 1.0 / n, amp=0.15)
    return et, np.ascontiguousarray(v), np.ascontiguousarray(c)


@pytest.mark.parametrize("kind,n,op", [("quad4", 5, fo.LAPLACE), ("quad4", 4, fo.LINEAR_ELASTIC), ("tet4", 2, fo.LINEAR_ELASTIC), ("hex8", 3, fo.LINEAR_ELASTIC),
                                      ("hex8", 3, fo.LAPLACE), ("hex27", 2, fo.LINEAR_ELASTIC), ("tet10", 1, fo.LAPLACE)])
def test_elliptic_vector_and_energy_equal_oracle(ctx, kind, n, op):
    et, v, c = _case(kind, n)
    prob = fo.Problem(et, v, c.astype(np.int64), op, params=() if op == fo.LAPLACE else (MU, LAM))
    data = None if op == fo.LAPLACE else (MU, LAM)
    s = prob.sdim
    u = np.random.default_rng(11).normal(size=s * len(v))
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.color_nodes()
    ref = fo.assemble_elliptic_vector_serial(prob, u)
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_COLORED):
        f = ctx.assemble_elliptic_vector(op, prob.weights, prob.points, data, u, scatter_mode=mode)
        assert np.abs(f - ref).max() < TOL * np.abs(ref).max()
    e = ctx.assemble_elliptic_scalar(op, prob.weights, prob.points, data, u)
    oe = fo.assemble_elliptic_scalar(prob, u)
    assert abs(e - oe) < TOL * abs(oe)
    # ties to the stiffness path: for these linear operators f(u) = K u and psi(u) = u.K u / 2
    ctx.assemble_pattern(s)
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, data)
    ctx.synchronize()
    Ku = ctx.spmv(u)
    assert np.abs(f - Ku).max() < 1e-11 * np.abs(Ku).max() and abs(e - 0.5 * u @ Ku) < 1e-11 * abs(e)
    # accumulate semantics of assemble_vector_into (global.rs:611)
    start = np.linspace(-1.0, 1.0, len(u)) * np.abs(ref).max()
    out = ctx.assemble_elliptic_vector(op, prob.weights, prob.points, data, u, out=start.copy(), accumulate=True)
    assert np.abs(out - start - ref).max() < TOL * np.abs(ref).max()


def test_elliptic_vector_singular_element_and_zero_u(ctx):
    cube = (np.array(fo._HEX8_NODES) + 1.0) / 2.0
    v = np.concatenate([cube + [2.0 * k, 0, 0] for k in range(3)])
    c = np.arange(24, dtype=np.uint64).reshape(3, 8)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    ctx.space_upload(fb.HEX8, v, c)
    assert not ctx.assemble_elliptic_vector(fo.LAPLACE, w, p, None, np.zeros(24)).any()  # u = 0 -> f = 0
    assert ctx.assemble_elliptic_scalar(fo.LINEAR_ELASTIC, w, p, (MU, LAM), np.zeros(72)) == 0.0
    v[8:16] = 0.5  # element 1 collapsed: "Singular element Jacobian encountered" (elliptic.rs:493-497)
    ctx.space_upload(fb.HEX8, v, c)
    with pytest.raises(fb.SingularJacobianError) as ei:
        ctx.assemble_elliptic_vector(fo.LAPLACE, w, p, None, np.ones(24))
    assert ei.value.element_index == 1


def test_reference_api_vector_and_scalar_of_the_elliptic_assembler():
    m = fb.create_unit_box_uniform_hex_mesh_3d(3)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(1e6, 0.2))
    qt = fb.UniformQuadratureTable.from_points_and_weights(p, w).with_uniform_data(lame)
    u = np.random.default_rng(4).normal(size=3 * m.num_nodes())
    ea = fb.ElementEllipticAssemblerBuilder().with_finite_element_space(m).with_operator(fb.MaterialEllipticOperator(fb.LinearElasticMaterial())) \
        .with_quadrature_table(qt).with_u(u).build()
    f = fb.VectorAssembler().assemble_vector(ea)
    f_par = fb.VectorParAssembler().assemble_vector(fb.color_nodes(m), ea)
    energy = fb.assemble_scalar(ea)
    prob = fo.Problem(fo.HEX8, m.vertices(), m.connectivity().astype(np.int64), fo.LINEAR_ELASTIC, params=(lame.mu, lame.lambda_))
    ref = fo.assemble_elliptic_vector_serial(prob, u)
    assert np.abs(f - ref).max() < TOL * np.abs(ref).max() and np.abs(f_par - ref).max() < TOL * np.abs(ref).max()
    assert abs(energy - fo.assemble_elliptic_scalar(prob, u)) < TOL * abs(energy)
