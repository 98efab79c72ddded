"""GPU tests of the consumer side: y = A x and the conjugate gradient of fenris-sparse/src/cg.rs:364-480 on the
device-resident CSR, against scipy and the literal numpy restatement; the reference's Poisson MMS goldens end to end on the device
(stiffness, source vector, Dirichlet rows and the solve never leave the library)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import fenris_b200 as fb
from oracle import fenris_oracle as fo
from tests import mms

pytestmark = pytest.mark.gpu
MU, LAM = fo.lame_from_young_poisson(1e6, 0.2)


@pytest.fixture(scope="module")
def ctx():
    c = fb.Context(0)
    yield c
    c.close()


def _csr(ctx):
    ro, ci = ctx.pattern_download()
    n = len(ro) - 1
    return sp.csr_matrix((ctx.values_download().copy(), ci.astype(np.int64), ro.astype(np.int64)), shape=(n, n))


@pytest.mark.parametrize("et,mesh,op,s", [(fo.QUAD4, lambda: fo.create_unit_square_uniform_quad_mesh_2d(9), fo.LAPLACE, 1),
                                         (fo.QUAD4, lambda: fo.create_unit_square_uniform_quad_mesh_2d(6), fo.LINEAR_ELASTIC, 2),
                                         (fo.HEX8, lambda: fo.create_unit_box_uniform_hex_mesh_3d(6), fo.LINEAR_ELASTIC, 3),
                                         (fo.TET4, lambda: fo.create_unit_box_uniform_tet_mesh_3d(3), fo.LAPLACE, 1)])
def test_spmv_equals_scipy(ctx, et, mesh, op, s):
    v, c = mesh()
    prob = fo.Problem(et, v, c.astype(np.int64), op, params=() if op == fo.LAPLACE else (MU, LAM))
    ctx.space_upload(et, v, c.astype(np.uint64))
    ctx.assemble_pattern(s)
    ctx.assemble_into_csr_device(op, prob.weights, prob.points, None if op == fo.LAPLACE else (MU, LAM))
    ctx.synchronize()
    A = _csr(ctx)
    x = np.random.default_rng(5).normal(size=A.shape[0])
    y = ctx.spmv(x)
    assert np.abs(y - A @ x).max() < 1e-12 * np.abs(A @ x).max()


def _poisson_system(ctx, n):
    v, c = fo.create_unit_square_uniform_quad_mesh_2d(n)
    w, p = fo.quadrilateral_gauss(2)
    ctx.space_upload(fo.QUAD4, v, c.astype(np.uint64))
    ctx.assemble_pattern(1)
    ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None)
    x = ctx.physical_quadrature_points(w, p, len(c))
    f = (2.0 * np.pi ** 2 * np.sin(np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1]))[..., None]
    b = ctx.assemble_vector(w, p, f, len(v))
    boundary = np.nonzero(np.abs(v - 0.5).max(axis=1) > 0.4999)[0]
    ctx.apply_homogeneous_dirichlet_bc_csr(boundary)
    b[boundary] = 0.0
    return v, b


@pytest.mark.parametrize("jacobi", [False, True])
def test_cg_equals_the_reference_algorithm(ctx, jacobi):
    v, b = _poisson_system(ctx, 20)
    A = _csr(ctx)
    x, its, res = ctx.cg_solve(b, rel_tol=1e-10, jacobi=jacobi)
    d = A.diagonal()
    ox, oits, status = fo.conjugate_gradient(lambda q: A @ q, b, rel_tol=1e-10, apply_p=(lambda r: r / d) if jacobi else None)
    assert status == "ok" and abs(its - oits) <= 1 and res <= 1e-10
    assert np.abs(x - ox).max() < 1e-8 * np.abs(ox).max()
    assert np.abs(x - spla.spsolve(A.tocsc(), b)).max() < 1e-8
    exact = np.sin(np.pi * v[:, 0]) * np.sin(np.pi * v[:, 1])
    assert np.abs(x - exact).max() < 1e-2  # discretisation error, h = 1/20
    # warm start from the solution: converged before the first update (cg.rs:408-424)
    x2, its2, _ = ctx.cg_solve(b, x0=x, rel_tol=1e-8, jacobi=jacobi)
    assert its2 == 0 and np.array_equal(x2, x)


def test_cg_special_cases_and_errors(ctx):
    v, b = _poisson_system(ctx, 8)
    x, its, res = ctx.cg_solve(np.zeros_like(b), x0=np.ones_like(b))
    assert its == 0 and not x.any()  # b = 0 -> x = 0 (cg.rs:403-406)
    b = np.random.default_rng(2).normal(size=len(b))  # (the MMS load is a discrete eigenvector: CG would finish in one step)
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.cg_solve(b, rel_tol=1e-14, max_iter=3)
    assert ei.value.status == fb.ERR_NOT_CONVERGED
    vals = ctx.values_download().copy()
    ctx.values_upload(-vals)  # -A is negative definite: p.Ap <= 0 at the first step
    with pytest.raises(fb.Fb200Error) as ei:
        ctx.cg_solve(b, jacobi=False)
    assert ei.value.status == fb.ERR_INDEFINITE
    ctx.values_upload(vals)


@pytest.mark.parametrize("name", ["quad4", "hex8"])
def test_mms_goldens_end_to_end_on_the_device(ctx, kats, name):
    # tests/convergence_tests/poisson_mms_common.rs:67-230 with every step on the device: stiffness, load vector, Dirichlet, CG;
    # the stored L2 / H1 errors of the reference (1 % tolerance, :40-65) pin the whole chain
    et, make_mesh, quad_rule, error_rule, key, resolutions = mms.CASES[name]
    golden = kats["mms_summaries"][key]
    u_ex, gu_ex, f_ex = mms.exact(2 if et == fo.QUAD4 else 3)
    for k, res in enumerate(resolutions):
        if res < 2:
            continue  # a single cell has no free node
        v, c = make_mesh(res)
        c64 = c.astype(np.int64)
        w, p = quad_rule()
        ctx.space_upload(et, v, c.astype(np.uint64))
        ctx.assemble_pattern(1)
        ctx.assemble_into_csr_device(fo.LAPLACE, w, p, None)
        xq = ctx.physical_quadrature_points(w, p, len(c))
        b = ctx.assemble_vector(w, p, f_ex(xq)[..., None], len(v))
        boundary = np.nonzero(np.abs(v - 0.5).max(axis=1) > 0.4999)[0]
        ctx.apply_homogeneous_dirichlet_bc_csr(boundary)
        b[boundary] = 0.0
        uh, _, _ = ctx.cg_solve(b, rel_tol=1e-12)
        l2 = h1 = 0.0
        for wq, xi in zip(*error_rule()):
            xqe, adet, G, N = mms._per_point(et, v, c64, xi)
            l2 += np.sum(wq * adet * (uh[c64] @ N - u_ex(xqe)) ** 2)
            h1 += np.sum(wq * adet * np.sum((np.einsum("eia,ea->ei", G, uh[c64]) - gu_ex(xqe)) ** 2, axis=1))
        assert abs(np.sqrt(l2) - golden["L2_errors"][k]) < 0.01 * golden["L2_errors"][k], (name, res)
        assert abs(np.sqrt(h1) - golden["H1_seminorm_errors"][k]) < 0.01 * golden["H1_seminorm_errors"][k], (name, res)
