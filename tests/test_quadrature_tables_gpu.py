"""GPU parity of assembly with a quadrature rule PER ELEMENT (CompactQuadratureTable / GeneralQuadratureTable,
src/assembly/local/quadrature_table.rs:57-210, 312-439 - SURVEY 8f rank 4) against the literal oracle loop, for the linear operators and the
state-dependent materials (StVK, NeoHookean) alike."""
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


def _rules(et, op):
    """three rules of different size; for elasticity also different Lame data per rule and per point"""
    if et == fo.TET4:
        base = [fo.tetrahedron_rule(1), fo.tetrahedron_rule(2), fo.tetrahedron_rule(2)]
    elif et == fo.QUAD4:
        base = [fo.quadrilateral_gauss(2), fo.quadrilateral_gauss(3), fo.quadrilateral_gauss(2)]
    else:
        base = [fo.hexahedron_gauss(2), fo.hexahedron_gauss(3), fo.hexahedron_gauss(2)]
    rules = []
    for r, (w, p) in enumerate(base):
        if op == fo.LAPLACE:
            rules.append((w, p, [()] * len(w), None))
        else:
            par = [(MU * (1.0 + 0.1 * r + (0.05 * k if r == 2 else 0.0)), LAM * (1.0 - 0.2 * r)) for k in range(len(w))]
            rules.append((w, p, par, np.array(par)))
    return rules


@pytest.mark.parametrize("kind,n,op", [("hex8", 4, fo.LINEAR_ELASTIC), ("hex8", 3, fo.LAPLACE), ("tet4", 2, fo.LINEAR_ELASTIC), ("quad4", 6, fo.LAPLACE),
                                       ("hex8", 3, fo.STVK), ("quad4", 5, fo.NEO_HOOKEAN), ("tet4", 2, fo.NEO_HOOKEAN)])
@pytest.mark.parametrize("mode", [fb.SCATTER_ATOMIC, fb.SCATTER_COLORED])
def test_compact_table_assembly_equals_oracle(ctx, kind, n, op, mode):
    if kind == "hex8":
        et, (v, c) = fo.HEX8, fo.create_unit_box_uniform_hex_mesh_3d(n)
    elif kind == "tet4":
        et, (v, c) = fo.TET4, fo.create_unit_box_uniform_tet_mesh_3d(n)
    else:
        et, (v, c) = fo.QUAD4, fo.create_unit_square_uniform_quad_mesh_2d(n)
    v = fo.jitter_vertices(v, 1.0 / n, amp=0.15)
    rules = _rules(et, op)
    emap = (np.arange(len(c)) * 7 + 1) % 3  # every rule is used, in no particular order
    s = 1 if op == fo.LAPLACE else v.shape[1]
    # the state of the non-linear materials (elliptic.rs:393-399), small against the cell size so that no element inverts
    u = (0.05 / n) * np.random.default_rng(17).normal(size=s * len(v)) if op in (fo.STVK, fo.NEO_HOOKEAN) else None
    oro, oci, ovals = fo.assemble_with_quadrature_table(et, v, c, op, [(w, p, par) for w, p, par, _ in rules], emap, u=u)
    assert not np.isnan(ovals).any()
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.assemble_pattern(s)
    ctx.color_nodes()
    ctx.assemble_into_csr_table_device(op, [(w, p, d) for w, p, _, d in rules], emap, scatter_mode=mode, u=u)
    ctx.synchronize()
    ro, ci = ctx.pattern_download()
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)
    vals = ctx.values_download().copy()
    assert fo.rel_frobenius(vals, ovals) < TOL
    # accumulate semantics, and a map that uses one rule only == the uniform-table path
    ctx.assemble_into_csr_table_device(op, [(w, p, d) for w, p, _, d in rules], emap, scatter_mode=mode, accumulate=True, u=u)
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), 2.0 * ovals) < TOL
    ctx.assemble_into_csr_table_device(op, [(w, p, d) for w, p, _, d in rules], np.zeros(len(c), dtype=np.uint32), scatter_mode=mode, u=u)
    ctx.synchronize()
    table_uniform = ctx.values_download().copy()
    ctx.assemble_into_csr_device(op, rules[0][0], rules[0][1], rules[0][3], scatter_mode=mode, u=u)
    ctx.synchronize()
    assert fo.rel_frobenius(table_uniform, ctx.values_download()) < 1e-13


def test_table_errors(ctx):
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(2)
    w, p = fo.hexahedron_gauss(2)
    ctx.space_upload(fo.HEX8, v, c.astype(np.uint64))
    ctx.assemble_pattern(1)
    with pytest.raises(fb.Fb200Error) as ei:  # rule index out of bounds (quadrature_table.rs:361-366 panics)
        ctx.assemble_into_csr_table_device(fo.LAPLACE, [(w, p, None)], np.array([0, 0, 0, 1, 0, 0, 0, 0]))
    assert ei.value.status == fb.ERR_INDEX_OOB and ei.value.element_index == 3
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.assemble_into_csr_table_device(fo.LAPLACE, [(w, p, None)], np.zeros(8), scatter_mode=fb.SCATTER_GATHER)
    assert ei.value.status == fb.ERR_UNSUPPORTED


def test_reference_api_general_and_compact_tables():
    # GeneralQuadratureTable (a rule per element) and the equivalent CompactQuadratureTable through CsrAssembler / CsrParAssembler
    m = fb.create_unit_box_uniform_hex_mesh_3d(3)
    rules = [fo.hexahedron_gauss(2), fo.hexahedron_gauss(3)]
    emap = np.arange(m.num_elements()) % 2
    lame = [fb.LameParameters(MU, LAM), fb.LameParameters(2.0 * MU, 0.5 * LAM)]
    data = [[lame[r]] * len(rules[r][0]) for r in range(2)]
    compact = fb.CompactQuadratureTable.from_quadrature_rules_and_map([r[1] for r in rules], [r[0] for r in rules], data, emap)
    general = fb.GeneralQuadratureTable.from_points_weights_and_data([rules[r][1] for r in emap], [rules[r][0] for r in emap], [data[r] for r in emap])
    assert len(general.weights) == 2  # identical rules were merged
    op = fb.MaterialEllipticOperator(fb.LinearElasticMaterial())
    u = np.zeros(3 * m.num_nodes())
    build = lambda qt: fb.ElementEllipticAssemblerBuilder().with_finite_element_space(m).with_operator(op).with_quadrature_table(qt).with_u(u).build()
    A = fb.CsrAssembler().assemble(build(compact))
    B = fb.CsrAssembler().assemble(build(general))
    Cp = fb.CsrParAssembler().assemble(fb.color_nodes(m), build(compact))
    _, _, ovals = fo.assemble_with_quadrature_table(fo.HEX8, m.vertices(), m.connectivity(), fo.LINEAR_ELASTIC,
                                                    [(rules[r][0], rules[r][1], [(lame[r].mu, lame[r].lambda_)] * len(rules[r][0])) for r in range(2)], emap)
    for M in (A, B, Cp):
        assert fo.rel_frobenius(M.values, ovals) < TOL


@pytest.mark.parametrize("kind,n,op", [("hex8", 3, fo.LINEAR_ELASTIC), ("quad4", 6, fo.LAPLACE), ("hex8", 3, fo.NEO_HOOKEAN), ("tet4", 2, fo.STVK)])
def test_elliptic_vector_and_energy_with_a_rule_per_element(ctx, kind, n, op):
    # ElementEllipticAssembler as vector / scalar assembler (elliptic.rs:342-359, 440-605) over a Compact table
    if kind == "hex8":
        et, (v, c) = fo.HEX8, fo.create_unit_box_uniform_hex_mesh_3d(n)
    elif kind == "tet4":
        et, (v, c) = fo.TET4, fo.create_unit_box_uniform_tet_mesh_3d(n)
    else:
        et, (v, c) = fo.QUAD4, fo.create_unit_square_uniform_quad_mesh_2d(n)
    v = fo.jitter_vertices(v, 1.0 / n, amp=0.15)
    rules = _rules(et, op)
    emap = (np.arange(len(c)) * 5 + 2) % 3
    s = 1 if op == fo.LAPLACE else v.shape[1]
    u = (0.05 / n) * np.random.default_rng(23).normal(size=s * len(v))
    orules = [(w, p, par) for w, p, par, _ in rules]
    ref = fo.assemble_elliptic_vector_with_table(et, v, c, op, orules, emap, u)
    oe = fo.assemble_elliptic_scalar_with_table(et, v, c, op, orules, emap, u)
    assert np.isfinite(ref).all() and np.isfinite(oe)
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.color_nodes()
    drules = [(w, p, d) for w, p, _, d in rules]
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_COLORED):
        f = ctx.assemble_elliptic_vector_table(op, drules, emap, u, scatter_mode=mode)
        assert np.abs(f - ref).max() < TOL * np.abs(ref).max(), mode
    f2 = ctx.assemble_elliptic_vector_table(op, drules, emap, u, out=f.copy(), accumulate=True)
    assert np.abs(f2 - 2.0 * ref).max() < 2 * TOL * np.abs(ref).max()
    e = ctx.assemble_elliptic_scalar_table(op, drules, emap, u)
    assert abs(e - oe) < TOL * abs(oe)
    # one rule for every element == the uniform-table entry points
    zero = np.zeros(len(c), dtype=np.uint32)
    fu = ctx.assemble_elliptic_vector(op, rules[0][0], rules[0][1], rules[0][3], u)
    assert np.abs(ctx.assemble_elliptic_vector_table(op, drules, zero, u) - fu).max() < 1e-13 * np.abs(fu).max()
    eu = ctx.assemble_elliptic_scalar(op, rules[0][0], rules[0][1], rules[0][3], u)
    assert abs(ctx.assemble_elliptic_scalar_table(op, drules, zero, u) - eu) < 1e-13 * abs(eu)
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.assemble_elliptic_vector_table(op, drules, np.where(np.arange(len(c)) == 2, 9, emap), u)
    assert ei.value.status == fb.ERR_INDEX_OOB and ei.value.element_index == 2
