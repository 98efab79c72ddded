"""GPU parity of the Hex20 (serendipity hexahedron, hexahedron.rs:369-563) path - the remaining element of north_star's high-order list:
element matrices, global stiffness in all scatter modes, mass matrix and source vector against the oracle (1e-12)."""
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


def _hex20(n, jitter=0.0):
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(n)
    if jitter:
        v = fo.jitter_vertices(v, 1.0 / n, amp=jitter)  # jitter the Hex8 vertices, THEN refine: edge nodes stay midpoints
    v20, c20 = fo.hex20_mesh_from_hex8(v, c)
    return np.ascontiguousarray(v20), np.ascontiguousarray(c20).astype(np.uint64)


def test_hex20_mesh_and_rule_equal_oracle():
    m = fb.hex20_mesh_from(fb.create_unit_box_uniform_hex_mesh_3d(3))
    v20, c20 = _hex20(3)
    assert np.array_equal(m.vertices(), v20) and np.array_equal(m.connectivity(), c20)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX20)
    ow, op = fo.canonical_stiffness_rule(fo.HEX20)  # hexahedron_gauss(3), canonical.rs:108-111
    assert np.array_equal(w, ow) and np.array_equal(p, op)


@pytest.mark.parametrize("op", [fo.LAPLACE, fo.LINEAR_ELASTIC])
def test_hex20_element_matrices(ctx, op):
    v, c = _hex20(2, 0.15)
    prob = fo.Problem(fo.HEX20, v, c.astype(np.int64), op, params=() if op == fo.LAPLACE else (MU, LAM))
    ctx.space_upload(fb.HEX20, v, c)
    dofs = prob.sdim * 20
    K = ctx.element_matrices(op, prob.weights, prob.points, None if op == fo.LAPLACE else (MU, LAM), 0, len(c), dofs)
    for e in range(len(c)):
        assert fo.rel_frobenius(K[e], prob.element_matrix(e)) < 1e-13


@pytest.mark.parametrize("op,n,jit", [(fo.LAPLACE, 3, 0.2), (fo.LINEAR_ELASTIC, 3, 0.0), (fo.LINEAR_ELASTIC, 2, 0.2)])
@pytest.mark.parametrize("mode", [fb.SCATTER_ATOMIC, fb.SCATTER_COLORED, fb.SCATTER_GATHER])
def test_hex20_global_assembly_equals_oracle(ctx, op, n, jit, mode):
    v, c = _hex20(n, jit)
    prob = fo.Problem(fo.HEX20, v, c.astype(np.int64), op, params=() if op == fo.LAPLACE else (MU, LAM))
    oro, oci, ovals = fo.assemble_fast(prob)
    ctx.space_upload(fb.HEX20, v, c)
    ctx.assemble_pattern(prob.sdim)
    ro, ci = ctx.pattern_download()
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)
    ctx.color_nodes()
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, None if op == fo.LAPLACE else (MU, LAM), scatter_mode=mode)
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), ovals) < TOL


def test_hex20_mass_and_source(ctx):
    v, c = _hex20(2, 0.1)
    w, p = fo.hexahedron_gauss(3)
    rho = np.linspace(1.0, 2.0, len(w))
    ctx.space_upload(fb.HEX20, v, c)
    ctx.assemble_pattern(3)
    ctx.assemble_mass_into_csr_device(w, p, rho)
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), fo.assemble_mass_fast(fo.HEX20, v, c, w, p, rho, 3)[2]) < TOL
    x = ctx.physical_quadrature_points(w, p, len(c))
    assert np.abs(x - fo.physical_quadrature_points(fo.HEX20, v, c, p)).max() < 1e-14
    f = np.stack([x[..., 0] * x[..., 1], 1.0 + x[..., 2]], axis=-1)
    ref = fo.assemble_vector_fast(fo.HEX20, v, c, w, p, f)
    assert np.abs(ctx.assemble_vector(w, p, f, len(v)) - ref).max() < TOL * np.abs(ref).max()
