"""GPU parity of the mass matrix / source vector / global vector path (src/assembly/local/mass.rs,
source.rs, global.rs:569-686) against the numpy oracle and the reference's own known answers.  Tolerance 1e-12 (fp64)."""
import numpy as np
import pytest

import fenris_b200 as fb
from oracle import fenris_oracle as fo

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = fb.Context(0)
    yield c
    c.close()


def _mesh(kind, n, jitter=0.0):
    if kind == "quad4":
        v, c = fo.create_unit_square_uniform_quad_mesh_2d(n)
        et = fo.QUAD4
    elif kind == "tet4":
        v, c = fo.create_unit_box_uniform_tet_mesh_3d(n)
        et = fo.TET4
    elif kind == "tet10":
        v, c = fo.tet10_mesh_from_tet4(*fo.create_unit_box_uniform_tet_mesh_3d(n))
        et = fo.TET10
    elif kind == "hex8":
        v, c = fo.create_unit_box_uniform_hex_mesh_3d(n)
        et = fo.HEX8
    else:
        v, c = fo.hex27_mesh_from_hex8(*fo.create_unit_box_uniform_hex_mesh_3d(n))
        et = fo.HEX27
    if jitter and kind in ("quad4", "tet4", "hex8"):
        v = fo.jitter_vertices(v, 1.0 / n, amp=jitter)
    return et, np.ascontiguousarray(v), np.ascontiguousarray(c).astype(np.uint64)


def _rule(et):
    # mass needs more strength than the canonical stiffness rule; Gauss 3^d integrates the (bi/tri)linear products exactly
    if et == fo.QUAD4:
        return fo.quadrilateral_gauss(3)
    if et in (fo.HEX8, fo.HEX27):
        return fo.hexahedron_gauss(3)
    return fo.tetrahedron_rule(2)


def test_reference_quad4_mass_matrix_kat(ctx):
    # tests/unit_tests/assembly/local.rs:38-69: density 3 on the reference quad, s = 2:
    # M = (rho / 9) [[4,2,1,2],[2,4,2,1],[1,2,4,2],[2,1,2,4]] (x) I_2
    v = np.array(fo._QUAD4_NODES, dtype=np.float64)
    c = np.array([[0, 1, 2, 3]], dtype=np.uint64)
    w, p = fo.quadrilateral_gauss(3)
    ctx.space_upload(fb.QUAD4, v, c)
    ctx.assemble_pattern(2)
    ctx.assemble_mass_into_csr_device(w, p, 3.0)
    ctx.synchronize()
    ro, ci = ctx.pattern_download()
    M = np.zeros((8, 8))
    vals = ctx.values_download()
    for r in range(8):
        M[r, ci[ro[r]:ro[r + 1]].astype(np.int64)] = vals[ro[r]:ro[r + 1]]
    expected = np.kron(3.0 / 9.0 * np.array([[4, 2, 1, 2], [2, 4, 2, 1], [1, 2, 4, 2], [2, 1, 2, 4.0]]), np.eye(2))
    assert np.abs(M - expected).max() < 1e-14


@pytest.mark.parametrize("kind,n,s,jit", [("quad4", 7, 1, 0.2), ("quad4", 5, 2, 0.0), ("tet4", 3, 1, 0.15), ("tet4", 2, 3, 0.0),
                                          ("hex8", 4, 3, 0.2), ("hex8", 5, 1, 0.0), ("hex27", 2, 3, 0.0), ("tet10", 2, 1, 0.0)])
@pytest.mark.parametrize("mode", [fb.SCATTER_ATOMIC, fb.SCATTER_COLORED])
def test_mass_matrix_equals_oracle(ctx, kind, n, s, jit, mode):
    et, v, c = _mesh(kind, n, jit)
    w, p = _rule(et)
    rho = np.linspace(0.5, 2.0, len(w))  # a different density at every point: the Parameters really are per point
    ctx.space_upload(et, v, c)
    ctx.assemble_pattern(s)
    ctx.color_nodes()
    ctx.assemble_mass_into_csr_device(w, p, rho, scatter_mode=mode)
    ctx.synchronize()
    oro, oci, ovals = fo.assemble_mass_fast(et, v, c, w, p, rho, s)
    ro, ci = ctx.pattern_download()
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)
    vals = ctx.values_download()
    assert fo.rel_frobenius(vals, ovals) < TOL
    # total mass = s * int rho: with rho = 1 the entries of a scalar mass matrix sum to the volume (partition of unity)
    ctx.assemble_mass_into_csr_device(w, p, 1.0, scatter_mode=mode)
    ctx.synchronize()
    one = fo.assemble_mass_fast(et, v, c, w, p, np.ones(len(w)), s)[2]
    assert abs(ctx.values_download().sum() - one.sum()) < 1e-12 * abs(one.sum())


def test_mass_literal_serial_oracle_and_accumulate(ctx):
    et, v, c = _mesh("hex8", 2, 0.2)
    w, p = _rule(et)
    rho = np.linspace(1.0, 3.0, len(w))
    _, _, ovals = fo.assemble_mass_serial(et, v, c, w, p, rho, 3)  # the literal restatement of mass.rs:218-286 + global.rs:133-182
    ctx.space_upload(et, v, c)
    ctx.assemble_pattern(3)
    start = np.linspace(-1.0, 1.0, ctx.nnz)
    vals = start.copy()
    ctx.assemble_mass_into_csr(w, p, rho, vals, accumulate=True)
    assert fo.rel_frobenius(vals - start, ovals) < TOL
    ctx.assemble_mass_into_csr(w, p, rho, vals, accumulate=False)
    assert fo.rel_frobenius(vals, ovals) < TOL


@pytest.mark.parametrize("kind,n,jit", [("quad4", 6, 0.2), ("tet4", 3, 0.1), ("hex8", 4, 0.2), ("hex27", 2, 0.0), ("tet10", 2, 0.0)])
def test_physical_points_and_source_vector_equal_oracle(ctx, kind, n, jit):
    et, v, c = _mesh(kind, n, jit)
    w, p = _rule(et)
    d = v.shape[1]
    ctx.space_upload(et, v, c)
    ctx.color_nodes()
    x = ctx.physical_quadrature_points(w, p, len(c))
    ox = fo.physical_quadrature_points(et, v, c, p)
    assert np.abs(x - ox).max() < 1e-14
    f = np.stack([np.sin(x[..., 0]) + x[..., 1] ** 2, 1.0 + x[..., d - 1]], axis=-1)  # s = 2, evaluated at the physical points
    ref = fo.assemble_vector_fast(et, v, c, w, p, f)
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_COLORED):
        out = ctx.assemble_vector(w, p, f, len(v), scatter_mode=mode)
        assert np.abs(out - ref).max() < TOL * np.abs(ref).max()
    # uniform source (e.g. gravity) + accumulate semantics (assemble_vector_into adds, global.rs:611)
    g = np.tile(np.array([0.0, -9.81, 2.0])[:d], (len(w), 1))
    ref_g = fo.assemble_vector_fast(et, v, c, w, p, g)
    start = np.linspace(0.0, 1.0, d * len(v))
    out = ctx.assemble_vector(w, p, g, len(v), out=start.copy(), accumulate=True)
    assert np.abs(out - start - ref_g).max() < TOL * np.abs(ref_g).max()
    # int 1 dV: the source vector of f = 1 sums to the volume
    ones = ctx.assemble_vector(w, p, np.ones((len(w), 1)), len(v))
    assert abs(ones.sum() - fo.assemble_vector_fast(et, v, c, w, p, np.ones((len(w), 1))).sum()) < 1e-13


def test_source_vector_literal_serial_oracle(ctx):
    et, v, c = _mesh("tet4", 2, 0.1)
    w, p = _rule(et)
    ctx.space_upload(et, v, c)
    x = ctx.physical_quadrature_points(w, p, len(c))
    f = np.stack([x[..., 0] * x[..., 1], x[..., 2] - 1.0, np.cos(x[..., 0])], axis=-1)
    ref = fo.assemble_vector_serial(et, v, c, w, p, f)  # literal source.rs:217-278 + add_local_to_global
    out = ctx.assemble_vector(w, p, f, len(v))
    assert np.abs(out - ref).max() < TOL * np.abs(ref).max()


def test_reference_api_mass_and_vector_assemblers():
    # the reference's call pattern: ElementMassAssembler through CsrAssembler::assemble (tests/unit_tests/assembly/local/mass.rs),
    # ElementSourceAssembler through VectorAssembler / VectorParAssembler (examples/poisson2d.rs:62-80)
    m = fb.create_unit_square_uniform_quad_mesh_2d(5)
    w, p = fo.quadrilateral_gauss(3)
    qt = fb.UniformQuadratureTable.from_points_weights_and_data(p, w, [fb.Density(2.0)] * len(w))
    mass = fb.ElementMassAssembler.with_finite_element_space(m).with_quadrature_table(qt).with_solution_dim(1).build()
    M = fb.CsrAssembler().assemble(mass)
    _, _, ovals = fo.assemble_mass_fast(fo.QUAD4, m.vertices(), m.connectivity(), w, p, np.full(len(w), 2.0), 1)
    assert fo.rel_frobenius(M.values, ovals) < TOL and abs(M.values.sum() - 2.0) < 1e-12  # int 2 dA over the unit square
    src = fb.ElementSourceAssembler(m, qt, lambda x, data: (float(data) * x[..., :1] ** 2), 1)
    b = fb.VectorAssembler().assemble_vector(src)
    colors = fb.color_nodes(m)
    b_par = fb.VectorParAssembler().assemble_vector(colors, src)
    x = fo.physical_quadrature_points(fo.QUAD4, m.vertices(), m.connectivity(), p)
    ref = fo.assemble_vector_fast(fo.QUAD4, m.vertices(), m.connectivity(), w, p, 2.0 * x[..., :1] ** 2)
    assert np.abs(b - ref).max() < 1e-13 and np.abs(b_par - ref).max() < 1e-13
    assert abs(b.sum() - 2.0 / 3.0) < 1e-12  # int 2 x^2 over the unit square


@pytest.mark.parametrize("kind,n,s", [("quad4", 6, 1), ("hex8", 4, 3), ("tet4", 3, 3), ("hex27", 2, 1)])
def test_dirichlet_bc_csr_equals_reference_semantics(ctx, kind, n, s):
    # apply_homogeneous_dirichlet_bc_csr (global.rs:379-451) on the device-resident stiffness matrix vs the literal restatement
    et, v, c = _mesh(kind, n, 0.0)
    op = fo.LAPLACE if s == 1 else fo.LINEAR_ELASTIC
    prob = fo.Problem(et, v, c.astype(np.int64), op, params=() if s == 1 else fo.lame_from_young_poisson(1e6, 0.2))
    ctx.space_upload(et, v, c)
    ctx.assemble_pattern(s)
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, None if s == 1 else fo.lame_from_young_poisson(1e6, 0.2))
    ctx.synchronize()
    ro, ci = ctx.pattern_download()
    ref = ctx.values_download().copy()
    boundary = np.nonzero(np.abs(v - 0.5).max(axis=1) > 0.4999)[0]  # the unit-domain boundary, as poisson_mms_common.rs:126-134
    scale = ctx.apply_homogeneous_dirichlet_bc_csr(boundary)
    oscale = fo.apply_homogeneous_dirichlet_bc_csr(ro, ci, ref, boundary, s)
    assert scale == oscale
    assert np.array_equal(ctx.values_download(), ref)  # only copies, zeros and one scale value: bit-exact
    # no Dirichlet nodes: nothing changes
    before = ctx.values_download().copy()
    ctx.apply_homogeneous_dirichlet_bc_csr(np.zeros(0, dtype=np.uint64))
    assert np.array_equal(ctx.values_download(), before)


def test_poisson_system_end_to_end_on_the_reference_api():
    # examples/poisson2d.rs:33-86: stiffness + source vector + homogeneous Dirichlet, then solve; u = sin(pi x) sin(pi y)
    import scipy.sparse.linalg as spla
    m = fb.create_unit_square_uniform_quad_mesh_2d(24)
    w, p = fo.quadrilateral_gauss(2)
    qt = fb.UniformQuadratureTable.from_points_and_weights(p, w)
    stiffness = fb.ElementEllipticAssemblerBuilder().with_finite_element_space(m).with_operator(fb.LaplaceOperator()).with_quadrature_table(qt) \
        .with_u(np.zeros(m.num_nodes())).build()
    asm = fb.CsrAssembler()
    A = asm.assemble(stiffness)
    f = lambda x, _data: 2.0 * np.pi ** 2 * np.sin(np.pi * x[..., :1]) * np.sin(np.pi * x[..., 1:2])
    b = fb.VectorAssembler().assemble_vector(fb.ElementSourceAssembler(m, qt, f, 1))
    v = m.vertices()
    boundary = np.nonzero(np.abs(v - 0.5).max(axis=1) > 0.4999)[0]
    fb.apply_homogeneous_dirichlet_bc_csr(A, boundary, 1)
    fb.apply_homogeneous_dirichlet_bc_rhs(b, boundary, 1)
    u = spla.spsolve(A.to_scipy().tocsc(), b)
    exact = np.sin(np.pi * v[:, 0]) * np.sin(np.pi * v[:, 1])
    assert np.abs(u - exact).max() < 5e-3  # O(h^2) with h = 1/24
