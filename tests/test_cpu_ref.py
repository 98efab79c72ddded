"""Pins the C restatement (oracle/cpu_ref.c, the CPU baseline) against the numpy oracle,
which is itself pinned to the reference's golden vectors (tests/test_oracle.py). CPU only."""
import numpy as np
import pytest

from oracle import cpu_ref as cr
from oracle import fenris_oracle as fo


def test_pattern_kats_exact(kats):
    for case in kats["pattern"]:
        offs, cols = cr.pattern(case["sdim"], case["num_nodes"], case["elements"])
        assert offs.tolist() == case["offsets"]
        assert cols.tolist() == case["indices"]


def test_meshes_bitwise_equal_oracle(kats):
    for n in (1, 2, 3):
        v, c = cr.gen_tet_mesh(n)
        vo, co = fo.create_unit_box_uniform_tet_mesh_3d(n)
        assert np.array_equal(v, vo) and np.array_equal(c.astype(np.int64), co)
    assert cr.gen_tet_mesh(2)[1].tolist() == kats["bcc_tet_mesh_2"]["connectivity"]
    for n in (1, 3, 7):
        v, c = cr.gen_hex_mesh(n)
        vo, co = fo.create_unit_box_uniform_hex_mesh_3d(n)
        assert np.array_equal(v, vo) and np.array_equal(c.astype(np.int64), co)
        v, c = cr.gen_quad_mesh(n)
        vo, co = fo.create_unit_square_uniform_quad_mesh_2d(n)
        assert np.array_equal(v, vo) and np.array_equal(c.astype(np.int64), co)


def test_coloring_equals_oracle():
    for conn, nn in ((fo.create_unit_box_uniform_hex_mesh_3d(4)[1], 125), (fo.create_unit_box_uniform_tet_mesh_3d(3)[1], 91)):
        coffs, celems = cr.color_greedy(conn.astype(np.uint64), nn)
        ref = fo.sequential_greedy_coloring(conn.tolist())
        assert len(coffs) - 1 == len(ref)
        for k, col in enumerate(ref):
            assert celems[coffs[k]:coffs[k + 1]].tolist() == col
    coffs, celems = cr.color_greedy([[0, 1, 2], [2, 3], [], [3, 4]], 5)
    assert coffs.tolist() == [0, 3, 4] and celems.tolist() == [0, 2, 3, 1]


@pytest.mark.parametrize("et", [fo.QUAD4, fo.TET4, fo.HEX8, fo.HEX27, fo.TET10])
@pytest.mark.parametrize("op", [fo.LAPLACE, fo.LINEAR_ELASTIC])
def test_element_matrix_equals_oracle(et, op):
    rng = np.random.default_rng(et * 10 + op)
    n, ng, d = fo.element_info(et)
    ref_nodes = {fo.QUAD4: fo._QUAD4_NODES, fo.HEX8: fo._HEX8_NODES, fo.HEX27: fo._HEX27_NODES}.get(et)
    if ref_nodes is None:
        X = np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=float)
        if et == fo.TET10:
            X = np.vstack([X] + [0.5 * (X[a] + X[b]) for a, b in fo._TET_EDGES])
    else:
        X = np.array(ref_nodes, dtype=float)
    X = X + rng.uniform(-0.1, 0.1, size=X.shape)
    w, p = fo.canonical_stiffness_rule(et)
    mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
    params = (mu, lam) if op == fo.LINEAR_ELASTIC else ()
    Ko = fo.element_matrix(et, X, op, w, p, [params] * len(w))
    Kc = cr.element_matrix(et, op, w, p, params, X)
    assert Kc.shape == Ko.shape
    assert fo.rel_frobenius(Kc, Ko) < 1e-15


def _cases():
    mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
    v, c = fo.create_unit_square_uniform_quad_mesh_2d(5)
    yield fo.Problem(fo.QUAD4, v, c, fo.LAPLACE)
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(2)
    yield fo.Problem(fo.TET4, fo.jitter_vertices(v, 0.5, amp=0.05), c, fo.LAPLACE)
    yield fo.Problem(fo.TET4, v, c, fo.LINEAR_ELASTIC, params=(mu, lam))
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(3)
    yield fo.Problem(fo.HEX8, fo.jitter_vertices(v, 1 / 3, amp=0.1), c, fo.LINEAR_ELASTIC, params=(mu, lam))
    v27, c27 = fo.hex27_mesh_from_hex8(*fo.create_unit_box_uniform_hex_mesh_3d(2))
    yield fo.Problem(fo.HEX27, v27, c27, fo.LINEAR_ELASTIC, params=(mu, lam))


@pytest.mark.parametrize("prob", list(_cases()))
def test_global_assembly_equals_oracle(prob):
    ro, ci, vals = fo.assemble_serial(prob)
    ro_c, ci_c = cr.pattern(prob.sdim, prob.num_nodes, prob.connectivity.astype(np.uint64))
    assert np.array_equal(ro, ro_c) and np.array_equal(ci, ci_c)
    par = prob.params_per_point[0]
    serial = cr.assemble(prob.elem_type, prob.op, prob.weights, prob.points, par, prob.vertices, prob.connectivity, ro, ci)
    assert fo.rel_frobenius(serial, vals) < 1e-15
    colors = cr.color_greedy(prob.connectivity.astype(np.uint64), prob.num_nodes)
    for nthreads in (1, 4):
        colored = cr.assemble(prob.elem_type, prob.op, prob.weights, prob.points, par, prob.vertices, prob.connectivity,
                              ro, ci, colors=colors, nthreads=nthreads)
        assert np.array_equal(colored, serial) or fo.rel_frobenius(colored, serial) < 1e-15


def test_singular_and_accumulate():
    v, c = fo.create_unit_box_uniform_hex_mesh_3d(2)
    ro, ci = cr.pattern(1, len(v), c.astype(np.uint64))
    w, p = fo.hexahedron_gauss(2)
    vals = cr.assemble(fo.HEX8, fo.LAPLACE, w, p, (), v, c, ro, ci)
    vals2 = cr.assemble(fo.HEX8, fo.LAPLACE, w, p, (), v, c, ro, ci, values=vals.copy())
    assert np.allclose(vals2, 2 * vals, rtol=1e-15)
    with pytest.raises(ArithmeticError):
        cr.assemble(fo.HEX8, fo.LAPLACE, w, p, (), np.zeros_like(v), c, ro, ci)
