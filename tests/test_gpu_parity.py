"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same inputs.
Run on the B200 box with `pytest -m gpu`.  Tolerance: 1e-12 relative Frobenius norm (north_star),
patterns / colours / meshes bit-exact."""
import numpy as np
import pytest

import fenris_b200 as fb
from oracle import cpu_ref as cr
from oracle import fenris_oracle as fo

pytestmark = pytest.mark.gpu

TOL = 1e-12  # north_star: relative Frobenius norm, fp64
MODES = [fb.SCATTER_ATOMIC, fb.SCATTER_COLORED, fb.SCATTER_GATHER]
MU, LAM = fo.lame_from_young_poisson(1e6, 0.2)


@pytest.fixture(scope="module")
def ctx():
    c = fb.Context(0)
    yield c
    c.close()


def _mesh(kind, n, jitter=0.0):
    if kind == "quad4":
        m = fb.create_unit_square_uniform_quad_mesh_2d(n)
    elif kind == "tet4":
        m = fb.create_unit_box_uniform_tet_mesh_3d(n)
    elif kind == "hex8":
        m = fb.create_unit_box_uniform_hex_mesh_3d(n)
    elif kind == "hex27":
        m = fb.hex27_mesh_from(fb.create_unit_box_uniform_hex_mesh_3d(n))
    else:
        raise ValueError(kind)
    if jitter:
        m = fb.Mesh(fo.jitter_vertices(m.vertices(), 1.0 / n, amp=jitter), m.connectivity(), m.element_type)
    return m


def _oracle_problem(mesh, op, params=None):
    par = () if op == fo.LAPLACE else (params if params is not None else (MU, LAM))
    return fo.Problem(mesh.element_type, mesh.vertices(), mesh.connectivity().astype(np.int64), op, params=par)


def _data(op, nq, params=None):
    if op == fo.LAPLACE:
        return None
    return params if params is not None else (MU, LAM)


# ------------------------------------------------------------------ patterns / colours
def test_pattern_kats_exact(ctx, kats):
    # the reference's own vectors: tests/unit_tests/assembly/global.rs:70-216
    for case in kats["pattern"]:
        ctx.connectivity_upload(case["num_nodes"], case["elements"])
        nrows, nnz = ctx.assemble_pattern(case["sdim"])
        ro, ci = ctx.pattern_download()
        assert nrows == case["nrows"]
        assert ro.tolist() == case["offsets"]
        assert ci.tolist() == case["indices"]


@pytest.mark.parametrize("kind,n", [("quad4", 7), ("tet4", 4), ("hex8", 5), ("hex27", 3)])
@pytest.mark.parametrize("sdim", [1, 3])
def test_pattern_equals_oracle(ctx, kind, n, sdim):
    m = _mesh(kind, n)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(sdim)
    ro, ci = ctx.pattern_download()
    oro, oci = fo.assemble_pattern_fast(sdim, m.num_nodes(), m.connectivity().astype(np.int64))
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)


def test_pattern_many_incident_elements(ctx):
    # a fan of 700 triangles-as-ragged-elements around node 0: exercises the global-scratch path (> 1024 candidates)
    elements = [[0, i, i + 1] for i in range(1, 701)]
    ctx.connectivity_upload(702, elements)
    ctx.assemble_pattern(2)
    ro, ci = ctx.pattern_download()
    oro, oci = fo.assemble_pattern(2, 702, elements)
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)


@pytest.mark.parametrize("kind,n", [("tet4", 3), ("hex8", 6)])
def test_colors_equal_reference_algorithm(ctx, kind, n):
    m = _mesh(kind, n)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ncol = ctx.color_nodes()
    offs, elems = ctx.colors_download(m.num_elements())
    ref = fo.sequential_greedy_coloring(m.connectivity().astype(np.int64).tolist())
    assert ncol == len(ref)
    for c, col in enumerate(ref):
        assert elems[int(offs[c]):int(offs[c + 1])].tolist() == col
    if kind == "hex8":
        assert ncol == 8


def test_colors_adopt_rejects_overlap(ctx):
    m = _mesh("hex8", 2)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.colors_adopt(np.array([0, 8], dtype=np.uint64), np.arange(8, dtype=np.uint64))  # all 8 cells share the centre node
    assert ei.value.status == fb.ERR_COLORING


# ------------------------------------------------------------------ element matrices
@pytest.mark.parametrize("kind,n", [("quad4", 3), ("tet4", 2), ("hex8", 2), ("hex27", 2)])
@pytest.mark.parametrize("op", [fo.LAPLACE, fo.LINEAR_ELASTIC])
def test_element_matrices_equal_oracle(ctx, kind, n, op):
    m = _mesh(kind, n, jitter=0.1)
    prob = _oracle_problem(m, op)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    dofs = prob.sdim * prob.n
    count = min(m.num_elements(), 24)
    K = ctx.element_matrices(op, prob.weights, prob.points, _data(op, len(prob.weights)), 0, count, dofs)
    for e in range(count):
        Ko = prob.element_matrix(e)
        assert fo.rel_frobenius(K[e], Ko) < 1e-13, (kind, op, e)
        assert np.array_equal(K[e], K[e].T), "device K_e must be exactly symmetric (implicit mirror)"


def test_reference_element_kats(ctx, kats):
    # reference Quad4 Laplace K = 1/6 [[4,-1,-2,-1],...] (tests/unit_tests/assembly.rs:159-162)
    X = np.array(fo._QUAD4_NODES)
    ctx.space_upload(fb.QUAD4, X, np.array([[0, 1, 2, 3]], dtype=np.uint64))
    w, p = fb.canonical_stiffness_quadrature(fb.QUAD4)
    K = ctx.element_matrices(fo.LAPLACE, w, p, None, 0, 1, 4)[0]
    assert np.allclose(K, np.array(kats["quad4_laplace_reference_element"]["sixth_times"]) / 6.0, atol=1e-15)
    # unit-cube Hex8 elasticity with mu=384, lambda=577 (fenris-solid fixtures): K00 = mu/3 + (lam+mu)/9
    mu, lam = kats["materials"]["mu"], kats["materials"]["lambda"]
    X = (np.array(fo._HEX8_NODES) + 1.0) / 2.0
    ctx.space_upload(fb.HEX8, X, np.arange(8, dtype=np.uint64)[None, :])
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    K = ctx.element_matrices(fo.LINEAR_ELASTIC, w, p, (mu, lam), 0, 1, 24)[0]
    assert abs(K[0, 0] - (mu / 3 + (lam + mu) / 9)) < 1e-11 and abs(K[0, 1] - (lam + mu) / 12) < 1e-11
    assert abs(np.linalg.norm(K) - 1722.593114184845) < 1e-9


# ------------------------------------------------------------------ global assembly
CASES = [("quad4", 32, fo.LAPLACE, 0.0),         # config C1 of BASELINE.json
         ("quad4", 9, fo.LINEAR_ELASTIC, 0.15),
         ("tet4", 6, fo.LAPLACE, 0.0),
         ("tet4", 5, fo.LINEAR_ELASTIC, 0.1),
         ("hex8", 8, fo.LINEAR_ELASTIC, 0.0),
         ("hex8", 7, fo.LINEAR_ELASTIC, 0.2),
         ("hex8", 6, fo.LAPLACE, 0.2),
         ("hex27", 3, fo.LINEAR_ELASTIC, 0.0),
         ("hex27", 2, fo.LAPLACE, 0.1)]


@pytest.mark.parametrize("kind,n,op,jit", CASES)
@pytest.mark.parametrize("mode", MODES)
def test_global_assembly_equals_oracle(ctx, kind, n, op, jit, mode):
    m = _mesh(kind, n, jitter=jit)
    prob = _oracle_problem(m, op)
    oro, oci, ovals = fo.assemble_fast(prob)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(prob.sdim)
    ro, ci = ctx.pattern_download()
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)
    ctx.color_nodes()
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, _data(op, len(prob.weights)), scatter_mode=mode, accumulate=False)
    ctx.synchronize()
    vals = ctx.values_download()
    assert fo.rel_frobenius(vals, ovals) < TOL, (kind, n, op, mode)


def test_literal_serial_oracle_small(ctx):
    # against the LITERAL restatement of CsrAssembler::assemble (not the vectorised one)
    m = _mesh("hex8", 3, jitter=0.1)
    prob = _oracle_problem(m, fo.LINEAR_ELASTIC)
    oro, oci, ovals = fo.assemble_serial(prob)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    for mode in MODES:
        ctx.color_nodes()
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, prob.weights, prob.points, (MU, LAM), scatter_mode=mode)
        ctx.synchronize()
        assert fo.rel_frobenius(ctx.values_download(), ovals) < TOL


@pytest.mark.parametrize("mode", MODES)
def test_accumulate_semantics_and_host_call(ctx, mode):
    # assemble_into_csr accumulates into existing values (global.rs:534); assemble() starts from zeros (:126)
    m = _mesh("tet4", 3)
    prob = _oracle_problem(m, fo.LINEAR_ELASTIC)
    _, _, ovals = fo.assemble_fast(prob)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    ctx.color_nodes()
    start = np.linspace(1.0, 2.0, ctx.nnz)
    vals = start.copy()
    ctx.assemble_into_csr(fo.LINEAR_ELASTIC, prob.weights, prob.points, (MU, LAM), vals, scatter_mode=mode, accumulate=True)
    assert fo.rel_frobenius(vals - start, ovals) < TOL
    ctx.assemble_into_csr(fo.LINEAR_ELASTIC, prob.weights, prob.points, (MU, LAM), vals, scatter_mode=mode, accumulate=False)
    assert fo.rel_frobenius(vals, ovals) < TOL


def test_per_point_parameters(ctx):
    # UniformQuadratureTable::from_points_weights_and_data: different Lame data at each quadrature point
    m = _mesh("hex8", 4, jitter=0.1)
    rng = np.random.default_rng(5)
    per_point = [(MU * (1 + 0.3 * rng.random()), LAM * (1 + 0.3 * rng.random())) for _ in range(8)]
    prob = fo.Problem(fo.HEX8, m.vertices(), m.connectivity().astype(np.int64), fo.LINEAR_ELASTIC, params=per_point)
    _, _, ovals = fo.assemble_fast(prob)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    ctx.color_nodes()
    for mode in MODES:
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, prob.weights, prob.points, np.array(per_point), scatter_mode=mode)
        ctx.synchronize()
        assert fo.rel_frobenius(ctx.values_download(), ovals) < TOL


def test_pattern_adopt_roundtrip_and_missing_column(ctx):
    m = _mesh("hex8", 4)
    prob = _oracle_problem(m, fo.LINEAR_ELASTIC)
    oro, oci, ovals = fo.assemble_fast(prob)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.pattern_adopt(3, oro, oci)
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, prob.weights, prob.points, (MU, LAM))
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), ovals) < TOL
    # a pattern built from fewer elements lacks couplings: "Could not find column index ..." (global.rs:533)
    pro, pci = fo.assemble_pattern_fast(3, m.num_nodes(), m.connectivity().astype(np.int64)[:-1])
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.pattern_adopt(3, pro, pci)
    assert ei.value.status == fb.ERR_COLUMN_NOT_IN_PATTERN and ei.value.element_index == m.num_elements() - 1


def test_error_paths(ctx):
    m = _mesh("hex8", 3)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    # singular Jacobian -> Err("Singular element Jacobian encountered") with the first offending element:
    # four disconnected unit cubes, cubes 1 and 3 collapsed to a point (det J == 0 exactly)
    cube = (np.array(fo._HEX8_NODES) + 1.0) / 2.0
    v = np.concatenate([cube + [2.0 * k, 0, 0] for k in range(4)])
    v[8:16] = 0.5
    v[24:32] = 0.25
    c = np.arange(32, dtype=np.uint64).reshape(4, 8)
    ctx.space_upload(fb.HEX8, v, c)
    ctx.assemble_pattern(1)
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_GATHER):
        ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None, scatter_mode=mode)
        with pytest.raises(fb.SingularJacobianError) as ei:
            ctx.synchronize()
        assert ei.value.element_index == 1
    # index out of bounds
    c = m.connectivity().copy()
    c[7, 3] = m.num_nodes()
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.space_upload(fb.HEX8, m.vertices(), c)
    assert ei.value.status == fb.ERR_INDEX_OOB and ei.value.element_index == 7
    # state errors
    ctx.space_upload(fb.HEX8, m.vertices(), m.connectivity())
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None)
    assert ei.value.status == fb.ERR_STATE
    ctx.assemble_pattern(3)
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None)  # pattern sdim 3 != Laplace sdim 1
    assert ei.value.status == fb.ERR_SHAPE
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_COLORED)  # no colours yet
    assert ei.value.status == fb.ERR_STATE
    # after the errors the context still works
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM))
    ctx.synchronize()


def test_empty_mesh(ctx):
    ctx.space_upload(fb.HEX8, np.zeros((0, 3)), np.zeros((0, 8), dtype=np.uint64))
    assert ctx.assemble_pattern(3) == (0, 0)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    ctx.color_nodes()
    for mode in MODES:
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=mode)
        ctx.synchronize()


# ------------------------------------------------------------------ reference-shaped API
def test_reference_api_serial_vs_parallel():
    # mirrors tests/convergence_tests/poisson_mms_common.rs:88-121 (CsrParAssembler == CsrAssembler)
    mesh = fb.create_unit_box_uniform_hex_mesh_3d(5)
    quadrature = fb.UniformQuadratureTable.from_quadrature(fb.canonical_stiffness_quadrature(fb.HEX8))
    u = np.zeros(mesh.num_nodes())
    colors = fb.color_nodes(mesh)
    laplace_assembler = (fb.ElementEllipticAssemblerBuilder().with_finite_element_space(mesh).with_operator(fb.LaplaceOperator())
                         .with_quadrature_table(quadrature).with_u(u).build())
    a_global = fb.CsrAssembler().assemble(laplace_assembler)
    par_a_global = fb.CsrParAssembler().assemble(colors, laplace_assembler)
    assert np.array_equal(a_global.row_offsets, par_a_global.row_offsets) and np.array_equal(a_global.col_indices, par_a_global.col_indices)
    assert fo.rel_frobenius(par_a_global.values, a_global.values) < 1e-14
    prob = fo.Problem(fo.HEX8, mesh.vertices(), mesh.connectivity().astype(np.int64), fo.LAPLACE)
    _, _, ovals = fo.assemble_fast(prob)
    assert fo.rel_frobenius(a_global.values, ovals) < TOL
    # assemble_into_csr accumulates into the caller's matrix
    fb.CsrAssembler().assemble_into_csr(a_global, laplace_assembler)
    assert fo.rel_frobenius(a_global.values, 2 * ovals) < TOL
    # elasticity through MaterialEllipticOperator + with_uniform_data
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(1e6, 0.2))
    qt = quadrature.with_uniform_data(lame)
    ea = (fb.ElementEllipticAssemblerBuilder().with_finite_element_space(mesh).with_operator(fb.MaterialEllipticOperator(fb.LinearElasticMaterial()))
          .with_quadrature_table(qt).with_u(np.zeros(3 * mesh.num_nodes())).build())
    A = fb.CsrAssembler(scatter_mode=fb.SCATTER_GATHER).assemble(ea)
    prob = fo.Problem(fo.HEX8, mesh.vertices(), mesh.connectivity().astype(np.int64), fo.LINEAR_ELASTIC, params=(lame.mu, lame.lambda_))
    _, _, ovals = fo.assemble_fast(prob)
    assert fo.rel_frobenius(A.values, ovals) < TOL
    Ke = ea.assemble_element_matrix(3)
    assert fo.rel_frobenius(Ke, prob.element_matrix(3)) < 1e-13


# ------------------------------------------------------------------ larger sizes: C oracle + size-independent properties
@pytest.mark.parametrize("mode", MODES)
def test_hex8_elasticity_40_cubed_vs_c_oracle(ctx, mode):
    n = 40
    m = _mesh("hex8", n, jitter=0.1)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    ro, ci = ctx.pattern_download()
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    colors = cr.color_greedy(m.connectivity(), m.num_nodes())
    ref = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (MU, LAM), m.vertices(), m.connectivity(), ro, ci, colors=colors)
    ctx.color_nodes()
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=mode)
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), ref) < TOL


def _operator_checks(ro, ci, vals, X):
    """Rigid-body null space (K t = 0, K (w x X) = 0) and symmetry (x.Ay == y.Ax) of an elasticity operator."""
    import scipy.sparse as sp
    n = len(ro) - 1
    A = sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(n, n))
    scale = np.abs(vals).max()
    res = 0.0
    for t in np.eye(3):
        res = max(res, np.abs(A @ np.tile(t, len(X))).max() / scale)
    for ax in range(3):
        wv = np.zeros(3)
        wv[ax] = 1.0
        res = max(res, np.abs(A @ np.cross(wv, X).ravel()).max() / scale)
    rng = np.random.default_rng(1)
    x, y = rng.normal(size=n), rng.normal(size=n)
    Ay, Ax = A @ y, A @ x
    sym = abs(x @ Ay - y @ Ax) / (np.linalg.norm(x) * np.linalg.norm(Ay))
    return res, sym


def test_full_size_c3_entrywise(ctx):
    """BASELINE config C3 (Hex8 elasticity, 126^3 = 2 000 376 elements, nnz 489 959 451) - the configuration the bench line is quoted
    on.  Every scatter mode (ATOMIC = the tile kernel, COLORED, GATHER) is compared ENTRYWISE with the C restatement of the reference's
    CsrParAssembler (oracle/cpu_ref.c, global.rs:314-376) at 1e-12 relative Frobenius norm, plus the size-independent properties:
    counts, rigid-body null space, symmetry, checksum."""
    n = 126
    m = _mesh("hex8", n)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    nrows, nnz = ctx.assemble_pattern(3)
    assert m.num_elements() == 2000376 and nrows == 6145149 and nnz == 489959451
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    assert ctx.color_nodes() == 8
    ro, ci = ctx.pattern_download()
    colors = cr.color_greedy(m.connectivity(), m.num_nodes())
    ref = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (MU, LAM), m.vertices(), m.connectivity(), ro, ci, colors=colors)
    res, sym = _operator_checks(ro, ci, ref, m.vertices())
    assert res < 1e-12 and sym < 1e-13
    buf = np.zeros(nnz)
    for mode in (fb.SCATTER_ATOMIC, fb.SCATTER_COLORED, fb.SCATTER_GATHER):
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=mode)
        ctx.synchronize()
        ctx.values_download(buf)
        assert fo.rel_frobenius(buf, ref) < TOL, mode
        assert np.abs(buf - ref).max() <= 1e-12 * np.abs(ref).max(), mode
        # sum of all entries = 1^T K 1 = 0 (1 is a rigid translation)
        assert abs(buf.sum()) < 1e-9 * np.abs(buf).sum()
    # the atomic path twice more: accumulate onto the assembled values (values += contributions, global.rs:133-182) and overwrite again
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_ATOMIC, accumulate=False)
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_ATOMIC, accumulate=True)
    ctx.synchronize()
    ctx.values_download(buf)
    assert fo.rel_frobenius(buf, 2.0 * ref) < TOL
    res, sym = _operator_checks(ro, ci, buf, m.vertices())
    assert res < 1e-12 and sym < 1e-13


def test_full_size_c4_hex27_entrywise(ctx):
    """BASELINE config C4: Hex27 elasticity on the 63^3 cube (250 047 elements, dense 81 x 81 K_e, nnz 1 159 088 625) through the
    FP64 tensor-pipe kernel (assemble_hex27_mma_kernel), entrywise against the C restatement of the reference CPU path."""
    m = _mesh("hex27", 63)
    assert m.num_elements() == 250047 and m.num_nodes() == 2048383
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    nrows, nnz = ctx.assemble_pattern(3)
    assert nrows == 6145149 and nnz == 1159088625
    ro, ci = ctx.pattern_download()
    w, p = fb.canonical_stiffness_quadrature(fb.HEX27)
    assert len(w) == 27
    colors = cr.color_greedy(m.connectivity(), m.num_nodes())
    ref = cr.assemble(fo.HEX27, fo.LINEAR_ELASTIC, w, p, (MU, LAM), m.vertices(), m.connectivity(), ro, ci, colors=colors)
    del ro, ci
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_ATOMIC)
    ctx.synchronize()
    vals = ctx.values_download()
    assert fo.rel_frobenius(vals, ref) < TOL
    assert np.abs(vals - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(vals.sum()) < 1e-9 * np.abs(vals).sum()


def test_full_size_c2_tet_poisson(ctx):
    """BASELINE config C2: Tet4 Poisson, 44^3 cells -> 1 022 208 tets, nnz 2 596 573; compared entrywise with the C oracle."""
    m = _mesh("tet4", 44)
    assert m.num_elements() == 1022208 and m.num_nodes() == 176309
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    nrows, nnz = ctx.assemble_pattern(1)
    assert nnz == 2596573
    ro, ci = ctx.pattern_download()
    w, p = fb.canonical_stiffness_quadrature(fb.TET4)
    ref = cr.assemble(fo.TET4, fo.LAPLACE, w, p, (), m.vertices(), m.connectivity(), ro, ci)
    ctx.color_nodes()
    for mode in MODES:
        ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None, scatter_mode=mode)
        ctx.synchronize()
        assert fo.rel_frobenius(ctx.values_download(), ref) < TOL


# ------------------------------------------------------------------ the reference's own convergence goldens, end to end on the GPU matrix
@pytest.mark.parametrize("name", ["quad4", "hex8", "tet4"])
def test_mms_goldens_with_gpu_matrix(ctx, kats, name):
    """tests/convergence_tests/poisson_{2d,3d}_mms.rs: assemble (GPU) -> Dirichlet -> solve -> L2/H1 errors within 1 % of the
    reference's stored values (reference_values/*.json)."""
    from tests import mms
    et, producer, qrule, erule, key, resolutions = mms.CASES[name]
    golden = kats["mms_summaries"][key]
    for k, res in enumerate(resolutions):
        v, c = producer(res)
        w, p = qrule()
        ctx.space_upload(et, v, c.astype(np.uint64))
        ctx.assemble_pattern(1)
        ro, ci = ctx.pattern_download()
        ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None, scatter_mode=fb.SCATTER_GATHER)
        ctx.synchronize()
        vals = ctx.values_download().copy()
        l2, h1 = mms.solve_poisson(et, v, c, mms.csr_from(ro, ci, vals), qrule(), erule())
        assert abs(l2 - golden["L2_errors"][k]) / golden["L2_errors"][k] < 0.01, (name, res, l2)
        assert abs(h1 - golden["H1_seminorm_errors"][k]) / golden["H1_seminorm_errors"][k] < 0.01, (name, res, h1)
