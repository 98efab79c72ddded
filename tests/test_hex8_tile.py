"""GPU parity tests of the Hex8 tile-accumulating scatter (hex8_tile_kernel.cuh + tiles.cpp) against the oracle.

The tile kernel replaces the per-element row scatter of global.rs:155-178, 504-537 for Hex8 + ATOMIC; the per-element
kernel (hex8_tile = 0), the coloured and the gather scatter are independent implementations of the same sums.
Tolerance: 1e-12 relative Frobenius norm (north_star); symmetry to rounding."""
import numpy as np
import pytest

import fenris_b200 as fb
from oracle import cpu_ref as cr
from oracle import fenris_oracle as fo

pytestmark = pytest.mark.gpu

TOL = 1e-12
MU, LAM = fo.lame_from_young_poisson(1e6, 0.2)
# "own" = tile kernel with first-writer stores (default, no zero-fill of the values), "zero" = tile kernel after a zero-fill, 0 = per-element kernel
TILES = ["own", "zero", 0]


@pytest.fixture(scope="module")
def ctx():
    c = fb.Context(0)
    yield c
    c.close()


def _hex(n, jitter=0.0, scramble=False, seed=7):
    m = fb.create_unit_box_uniform_hex_mesh_3d(n)
    v, c = m.vertices(), m.connectivity()
    if jitter:
        v = fo.jitter_vertices(v, 1.0 / n, amp=jitter)
    if scramble:  # arbitrary node numbering and element order: nothing structured is left but the geometry
        rng = np.random.default_rng(seed)
        perm = rng.permutation(len(v))
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(v))
        v = v[perm]
        c = inv[c.astype(np.int64)].astype(np.uint64)
        c = c[rng.permutation(len(c))]
    return fb.Mesh(np.ascontiguousarray(v), np.ascontiguousarray(c), fb.HEX8)


def _tune(ctx, tile):
    ctx.set_tuning("hex8_tile", 0 if tile == 0 else 64)
    ctx.set_tuning("hex8_owner_stores", 0 if tile == "zero" else 1)


def _assemble(ctx, m, op, tile, accumulate=False):
    prob = fo.Problem(fo.HEX8, m.vertices(), m.connectivity().astype(np.int64), op, params=() if op == fo.LAPLACE else (MU, LAM))
    _tune(ctx, tile)
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, None if op == fo.LAPLACE else (MU, LAM),
                                 scatter_mode=fb.SCATTER_ATOMIC, accumulate=accumulate)
    ctx.synchronize()
    return prob


@pytest.mark.parametrize("tile", TILES)
@pytest.mark.parametrize("n,op,jit,scramble", [(9, fo.LINEAR_ELASTIC, 0.2, False), (8, fo.LINEAR_ELASTIC, 0.0, False),
                                              (10, fo.LAPLACE, 0.15, False), (7, fo.LINEAR_ELASTIC, 0.1, True),
                                              (1, fo.LINEAR_ELASTIC, 0.0, False), (2, fo.LAPLACE, 0.1, False)])
def test_tile_scatter_equals_oracle(ctx, tile, n, op, jit, scramble):
    m = _hex(n, jit, scramble)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    sdim = 1 if op == fo.LAPLACE else 3
    ctx.assemble_pattern(sdim)
    prob = _assemble(ctx, m, op, tile)
    oro, oci, ovals = fo.assemble_fast(prob)
    ro, ci = ctx.pattern_download()
    assert np.array_equal(ro, oro) and np.array_equal(ci, oci)
    vals = ctx.values_download()
    assert fo.rel_frobenius(vals, ovals) < TOL, (tile, n, op)
    # upper-triangle-then-mirror semantics (operators.rs:177-180, util.rs:38-50): (u, v) and (v, u) get the same per-tile sums;
    # only the order in which different tiles' partial sums reach a shared row differs
    import scipy.sparse as sp
    A = sp.csr_matrix((vals, ci.astype(np.int64), ro.astype(np.int64)), shape=(len(ro) - 1,) * 2)
    assert abs(A - A.T).max() <= 4e-16 * np.abs(vals).max()


@pytest.mark.parametrize("tile", TILES)
def test_tile_accumulate_semantics(ctx, tile):
    # assemble_into_csr adds to the existing values (global.rs:534): complete rows must not be overwritten then
    m = _hex(6, 0.1)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    prob = _assemble(ctx, m, fo.LINEAR_ELASTIC, tile)
    once = ctx.values_download().copy()
    start = np.linspace(-1.0, 3.0, ctx.nnz) * np.abs(once).max()
    ctx.values_upload(start)
    _assemble(ctx, m, fo.LINEAR_ELASTIC, tile, accumulate=True)
    assert fo.rel_frobenius(ctx.values_download() - start, once) < 1e-12
    _assemble(ctx, m, fo.LINEAR_ELASTIC, tile, accumulate=False)
    assert np.array_equal(ctx.values_download(), once) or fo.rel_frobenius(ctx.values_download(), once) < 1e-15


@pytest.mark.parametrize("tile", ["own"])
def test_tile_degenerate_elements_fall_back(ctx, tile):
    # an element with a repeated node cannot use the tile accumulators (two lanes would share one): the library must fall back to
    # the per-element kernel and still produce the reference's sums (the collapsed element itself is singular)
    m = _hex(3)
    c = m.connectivity().copy()
    c[5, 1] = c[5, 0]
    ctx.space_upload(fb.HEX8, m.vertices(), c)
    ctx.assemble_pattern(1)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    _tune(ctx, tile)
    ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None, scatter_mode=fb.SCATTER_ATOMIC)
    a = None
    try:
        ctx.synchronize()
        a = ctx.values_download().copy()
    except fb.SingularJacobianError:
        pass
    ctx.set_tuning("hex8_tile", 0)
    ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None, scatter_mode=fb.SCATTER_ATOMIC)
    try:
        ctx.synchronize()
        b = ctx.values_download()
        assert a is not None and fo.rel_frobenius(a, b) < TOL
    except fb.SingularJacobianError:
        assert a is None


def test_tile_singular_element_index(ctx):
    # the deferred error reports the smallest offending ELEMENT id, not a schedule position (elliptic.rs:401-404)
    cube = (np.array(fo._HEX8_NODES) + 1.0) / 2.0
    v = np.concatenate([cube + [2.0 * k, 0, 0] for k in range(6)])
    v[16:24] = 0.5
    v[40:48] = 0.25
    c = np.arange(48, dtype=np.uint64).reshape(6, 8)
    ctx.space_upload(fb.HEX8, v, c)
    ctx.assemble_pattern(3)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    for tile in TILES:
        _tune(ctx, tile)
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_ATOMIC)
        with pytest.raises(fb.SingularJacobianError) as ei:
            ctx.synchronize()
        assert ei.value.element_index == 2


@pytest.mark.parametrize("tile", ["own", "zero"])
def test_tile_40_cubed_vs_c_oracle_and_modes(ctx, tile):
    n = 40
    m = _hex(n, 0.1)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    ro, ci = ctx.pattern_download()
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    colors = cr.color_greedy(m.connectivity(), m.num_nodes())
    ref = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (MU, LAM), m.vertices(), m.connectivity(), ro, ci, colors=colors)
    _assemble(ctx, m, fo.LINEAR_ELASTIC, tile)
    vals = ctx.values_download().copy()
    assert fo.rel_frobenius(vals, ref) < TOL
    ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_GATHER)
    ctx.synchronize()
    assert fo.rel_frobenius(vals, ctx.values_download()) < 1e-14


@pytest.mark.parametrize("tile", TILES)
@pytest.mark.parametrize("n", [5, 12])
def test_tile_overwrite_ignores_previous_values(ctx, tile, n):
    # assemble() starts from zeros (global.rs:126): whatever the value array held before must not leak into the result - the
    # tile path clears only the rows that receive reductions and overwrites the rows of tile-complete nodes with plain stores
    m = _hex(n, 0.1)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    prob = fo.Problem(fo.HEX8, m.vertices(), m.connectivity().astype(np.int64), fo.LINEAR_ELASTIC, params=(MU, LAM))
    _, _, ovals = fo.assemble_fast(prob)
    ctx.values_upload(np.full(ctx.nnz, 1e300))
    _assemble(ctx, m, fo.LINEAR_ELASTIC, tile)
    assert fo.rel_frobenius(ctx.values_download(), ovals) < TOL


def test_tile_repeatable(ctx):
    # within a tile the sums are formed in a fixed order; only the reductions into rows shared between tiles commute
    m = _hex(12, 0.1)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(3)
    for tile in ("own", "zero"):
        _assemble(ctx, m, fo.LINEAR_ELASTIC, tile)
        a = ctx.values_download().copy()
        for _ in range(5):
            _assemble(ctx, m, fo.LINEAR_ELASTIC, tile)
            assert fo.rel_frobenius(ctx.values_download(), a) < 1e-15


@pytest.mark.parametrize("n,op,scramble", [(9, fo.LINEAR_ELASTIC, False), (10, fo.LAPLACE, False), (7, fo.LINEAR_ELASTIC, True), (24, fo.LINEAR_ELASTIC, True)])
def test_tile_owner_stores_equal_zero_fill(ctx, n, op, scramble):
    # first-writer stores (no zero-fill: the owner of a shared row writes all of its entries, the other tiles wait for its flag and
    # reduce) against zero-fill + reductions: the same per-tile partial sums, added in a different order
    m = _hex(n, 0.15, scramble)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(1 if op == fo.LAPLACE else 3)
    _assemble(ctx, m, op, "zero")
    ref = ctx.values_download().copy()
    for fill in (1e300, np.nan, 0.0):
        ctx.values_upload(np.full(ctx.nnz, fill))
        prob = _assemble(ctx, m, op, "own")
        vals = ctx.values_download().copy()
        assert fo.rel_frobenius(vals, ref) < 1e-15
    assert fo.rel_frobenius(vals, fo.assemble_fast(prob)[2]) < TOL


def test_tile_owner_stores_with_ghost_elements(ctx):
    # a partition: rows that ghost elements touch are never stored (the neighbouring rank adds to them): they are cleared instead,
    # and the owned contributions are reduced into them; every other row is stored by its owner tile
    from fenris_b200.partition import structured_hex_slab
    verts, conn, n_owned, _ = structured_hex_slab(8, 8, 12, 1.0 / 8, 1, 3)
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    ctx.space_upload(fb.HEX8, verts, conn)
    ctx.set_num_owned_elements(n_owned)
    ctx.assemble_pattern(3)
    out = {}
    for tile in ("zero", "own"):
        _tune(ctx, tile)
        ctx.values_upload(np.full(ctx.nnz, 1e300))
        ctx.assemble_into_csr_device(fo.LINEAR_ELASTIC, w, p, (MU, LAM), scatter_mode=fb.SCATTER_ATOMIC)
        ctx.synchronize()
        out[tile] = ctx.values_download().copy()
    _tune(ctx, "own")
    assert np.isfinite(out["own"]).all() and fo.rel_frobenius(out["own"], out["zero"]) < 1e-15


@pytest.mark.parametrize("n,op,scramble", [(9, fo.LINEAR_ELASTIC, False), (10, fo.LAPLACE, False), (7, fo.LINEAR_ELASTIC, True), (20, fo.LINEAR_ELASTIC, False)])
def test_colored_scatter_by_tile_colours_is_bitwise_reproducible(ctx, n, op, scramble):
    # FB200_SCATTER_COLORED on a Hex8 space: one launch of the tile kernel per TILE colour (CsrParAssembler's colouring idea,
    # global.rs:322-373, at tile granularity).  Fixed order inside a tile, fixed colour order: identical bits on every run.
    m = _hex(n, 0.15, scramble)
    ctx.space_upload(m.element_type, m.vertices(), m.connectivity())
    ctx.assemble_pattern(1 if op == fo.LAPLACE else 3)
    ctx.color_nodes()
    prob = fo.Problem(fo.HEX8, m.vertices(), m.connectivity().astype(np.int64), op, params=() if op == fo.LAPLACE else (MU, LAM))
    data = None if op == fo.LAPLACE else (MU, LAM)
    _tune(ctx, "own")
    runs = []
    for fill in (1e300, 0.0, -3.0):
        ctx.values_upload(np.full(ctx.nnz, fill))
        ctx.assemble_into_csr_device(op, prob.weights, prob.points, data, scatter_mode=fb.SCATTER_COLORED, accumulate=False)
        ctx.synchronize()
        runs.append(ctx.values_download().copy())
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    assert fo.rel_frobenius(runs[0], fo.assemble_fast(prob)[2]) < TOL
    # the per-element colours (one launch of the element kernel per colour) give the same sums in another order
    ctx.set_tuning("hex8_colored_tiles", 0)
    try:
        ctx.assemble_into_csr_device(op, prob.weights, prob.points, data, scatter_mode=fb.SCATTER_COLORED, accumulate=False)
        ctx.synchronize()
        assert fo.rel_frobenius(ctx.values_download(), runs[0]) < 1e-14
    finally:
        ctx.set_tuning("hex8_colored_tiles", 1)
    # accumulate semantics (global.rs:534)
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, data, scatter_mode=fb.SCATTER_COLORED, accumulate=True)
    ctx.synchronize()
    assert fo.rel_frobenius(ctx.values_download(), 2.0 * runs[0]) < TOL
