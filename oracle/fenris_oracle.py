"""CPU oracle: a numpy restatement of fenris's global operator assembly path.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product (``fenris_b200``) never does and
fails loudly when its CUDA library is missing.

Every function restates one piece of the reference (InteractiveComputerGraphics/
fenris @ 7181b15, v0.0.33) and cites the ``file:line`` it follows.  Paths are
relative to the reference checkout.  The restatement is deliberately literal
(same loop order, same accumulation order, same upper-triangle-then-mirror
rule) so that it is the arbiter for parity; ``assemble_fast`` is a vectorised
variant used only to reach larger meshes and is itself checked against the
literal path in ``tests/test_oracle.py``.

Parity pinning: the reference cannot be compiled here (no Rust toolchain), so
the oracle is pinned against the reference's own golden vectors (pattern KATs,
BCC tet mesh insta snapshots, Hex8->Hex27 single element test, Lame
conversion, linear-elastic energy densities, reference Quad4 Laplace matrix)
held in ``tests/golden/`` - see ``tests/golden/make_golden.py``.

Third-party arithmetic that is NOT in the reference tree: nalgebra 0.32.1
(``Cargo.toml:90``) ``determinant()`` / ``try_inverse()`` for 2x2 and 3x3
matrices.  Its published closed forms (nalgebra ``src/linalg/determinant.rs``,
``src/linalg/inverse.rs``) are restated in ``det_small`` / ``try_inverse_small``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# Element type ids (shared with include/fenris_b200.h)
# --------------------------------------------------------------------------
QUAD4 = 1
TET4 = 2
HEX8 = 3
HEX27 = 4
TET10 = 5
HEX20 = 6

LAPLACE = 1
LINEAR_ELASTIC = 2

_ELEMENT_INFO = {
    # type: (nodes, geometry nodes, dim)
    QUAD4: (4, 4, 2),
    TET4: (4, 4, 3),
    HEX8: (8, 8, 3),
    HEX27: (27, 8, 3),
    TET10: (10, 4, 3),
    HEX20: (20, 8, 3),
}


def element_info(elem_type: int) -> Tuple[int, int, int]:
    return _ELEMENT_INFO[elem_type]


# --------------------------------------------------------------------------
# Quadrature  (fenris-quadrature)
# --------------------------------------------------------------------------
def _legendre(n: int, x: float) -> Tuple[float, float]:
    """p_n(x), p_{n-1}(x) by the three-term recurrence.

    fenris-quadrature/src/univariate.rs:22-36 (LegendreRecurrence::evaluate).
    """
    p1, p2 = 1.0, 0.0
    for m in range(1, n + 1):
        mf = float(m)
        p3 = p2
        p2 = p1
        p1 = ((2.0 * mf - 1.0) * x * p2 - (mf - 1.0) * p3) / mf
    return p1, p2


def gauss(n: int) -> Tuple[List[float], List[float]]:
    """Gauss rule on [-1, 1]: (weights, points), positive roots first.

    fenris-quadrature/src/univariate.rs:66-117.
    """
    assert n > 0
    m = (n + 1) // 2
    points: List[float] = []
    weights: List[float] = []
    for i in range(m):
        x = math.cos(math.pi * (i + 0.75) / (n + 0.5))
        p1, p2 = _legendre(n, x)
        dp = n * (x * p1 - p2) / (x * x - 1.0)
        p = p1
        while True:
            dx = -p / dp
            x += dx
            p1, p2 = _legendre(n, x)
            p = p1
            dp = n * (x * p1 - p2) / (x * x - 1.0)
            if abs(dx) <= 1e-15:
                break
        w = 2.0 / ((1.0 - x * x) * dp * dp)
        points.append(x)
        weights.append(w)
    for i in range(m, n):
        mirror = n - i - 1
        points.append(-points[mirror])
        weights.append(weights[mirror])
    return weights, points


def quadrilateral_gauss(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """fenris-quadrature/src/tensor.rs:13-32 (x outer, y inner)."""
    w1, p1 = gauss(n)
    w, p = [], []
    for wx, x in zip(w1, p1):
        for wy, y in zip(w1, p1):
            w.append(wx * wy)
            p.append([x, y])
    return np.array(w), np.array(p)


def hexahedron_gauss(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """fenris-quadrature/src/tensor.rs:36-58 (x outer, z inner, w = wx*wy*wz)."""
    w1, p1 = gauss(n)
    w, p = [], []
    for wx, x in zip(w1, p1):
        for wy, y in zip(w1, p1):
            for wz, z in zip(w1, p1):
                w.append(wx * wy * wz)
                p.append([x, y, z])
    return np.array(w), np.array(p)


def tetrahedron_rule(strength: int) -> Tuple[np.ndarray, np.ndarray]:
    """Smallest polyquad tet rule with at least the given strength (only 1, 2).

    fenris-quadrature/src/polyquad.rs:54-56 + build.rs:172-197, data
    fenris-quadrature/rules/polyquad/expanded/tet/{1-1,2-4}.txt (38-digit
    decimals rounded to f64 here exactly as Rust's f64 literal parser does).
    """
    if strength <= 1:
        return np.array([1.3333333333333333333333333333333333333]), np.array([[-0.5, -0.5, -0.5]])
    if strength == 2:
        a = -0.72360679774997896964091736687312762354
        b = 0.17082039324993690892275210061938287063
        w = 0.33333333333333333333333333333333333333
        return np.array([w, w, w, w]), np.array([[a, a, b], [a, b, a], [b, a, a], [a, a, a]])
    raise NotImplementedError("only tet strengths 1 and 2 are on the hot path")


def canonical_stiffness_rule(elem_type: int) -> Tuple[np.ndarray, np.ndarray]:
    """src/quadrature/canonical.rs:95,102-104,110-112."""
    if elem_type == QUAD4:
        return quadrilateral_gauss(2)
    if elem_type == HEX8:
        return hexahedron_gauss(2)
    if elem_type in (HEX27, HEX20):
        return hexahedron_gauss(3)
    if elem_type == TET4:
        return tetrahedron_rule(1)
    if elem_type == TET10:
        return tetrahedron_rule(2)
    raise ValueError(elem_type)


# --------------------------------------------------------------------------
# Reference elements (src/element*.rs)
# --------------------------------------------------------------------------
def _phi_lin(alpha: float, x: float) -> float:  # src/element.rs:246-253
    return (1.0 + alpha * x) / 2.0


def _dphi_lin(alpha: float) -> float:  # src/element.rs:258-263
    return alpha / 2.0


def _phi_quad(alpha: float, x: float) -> float:  # src/element.rs:272-281
    a2 = alpha * alpha
    return (3.0 / 2.0 * a2 - 1.0) * (x * x) + 0.5 * alpha * x + 1.0 - a2


def _dphi_quad(alpha: float, x: float) -> float:  # src/element.rs:289-298
    a2 = alpha * alpha
    return 2.0 * (3.0 / 2.0 * a2 - 1.0) * x + 0.5 * alpha


# node sign tables
_QUAD4_NODES = [(-1.0, -1.0), (1.0, -1.0), (1.0, 1.0), (-1.0, 1.0)]  # quadrilateral.rs:84-89
_HEX8_NODES = [  # hexahedron.rs:50-59
    (-1.0, -1.0, -1.0), (1.0, -1.0, -1.0), (1.0, 1.0, -1.0), (-1.0, 1.0, -1.0),
    (-1.0, -1.0, 1.0), (1.0, -1.0, 1.0), (1.0, 1.0, 1.0), (-1.0, 1.0, 1.0),
]
_HEX27_NODES = _HEX8_NODES + [  # hexahedron.rs:176-212 / :236-268
    (0.0, -1.0, -1.0), (-1.0, 0.0, -1.0), (-1.0, -1.0, 0.0), (1.0, 0.0, -1.0),
    (1.0, -1.0, 0.0), (0.0, 1.0, -1.0), (1.0, 1.0, 0.0), (-1.0, 1.0, 0.0),
    (0.0, -1.0, 1.0), (-1.0, 0.0, 1.0), (1.0, 0.0, 1.0), (0.0, 1.0, 1.0),
    (0.0, 0.0, -1.0), (0.0, -1.0, 0.0), (-1.0, 0.0, 0.0), (1.0, 0.0, 0.0),
    (0.0, 1.0, 0.0), (0.0, 0.0, 1.0),
    (0.0, 0.0, 0.0),
]
_TET4_GRADS = np.array([[-0.5, -0.5, -0.5], [0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 0.5]]).T  # tetrahedron.rs:561-568
_TET10_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (2, 3), (1, 3)]  # tetrahedron.rs:236-241


def tet4_basis(xi: Sequence[float]) -> np.ndarray:
    """src/element/tetrahedron.rs:551-558."""
    x, y, z = xi
    return np.array([-0.5 * x - 0.5 * y - 0.5 * z - 0.5, 0.5 * x + 0.5, 0.5 * y + 0.5, 0.5 * z + 0.5])


def hex8_basis(xi: Sequence[float]) -> np.ndarray:
    """src/element/hexahedron.rs:43-60."""
    return np.array([_phi_lin(a, xi[0]) * _phi_lin(b, xi[1]) * _phi_lin(c, xi[2]) for a, b, c in _HEX8_NODES])


def reference_gradients(elem_type: int, xi: Sequence[float]) -> np.ndarray:
    """Reference-basis gradients G_ref(xi), shape (d, n), column = node.

    Quad4 src/element/quadrilateral.rs:94-107; Tet4 tetrahedron.rs:561-568;
    Tet10 tetrahedron.rs:198-223; Hex8 hexahedron.rs:63-83; Hex27
    hexahedron.rs:269-315.
    """
    if elem_type == QUAD4:
        g = np.empty((2, 4))
        for k, (a, b) in enumerate(_QUAD4_NODES):
            g[0, k] = a * (1.0 + b * xi[1]) / 4.0
            g[1, k] = b * (1.0 + a * xi[0]) / 4.0
        return g
    if elem_type == TET4:
        return _TET4_GRADS.copy()
    if elem_type == TET10:
        psi = tet4_basis(xi)
        g4 = _TET4_GRADS
        cols = [g4[:, i] * (4.0 * psi[i] - 1.0) for i in range(4)]
        cols += [g4[:, i] * (4.0 * psi[j]) + g4[:, j] * (4.0 * psi[i]) for i, j in _TET10_EDGES]
        return np.stack(cols, axis=1)
    if elem_type == HEX8:
        g = np.empty((3, 8))
        for k, (a, b, c) in enumerate(_HEX8_NODES):
            g[0, k] = _dphi_lin(a) * _phi_lin(b, xi[1]) * _phi_lin(c, xi[2])
            g[1, k] = _phi_lin(a, xi[0]) * _dphi_lin(b) * _phi_lin(c, xi[2])
            g[2, k] = _phi_lin(a, xi[0]) * _phi_lin(b, xi[1]) * _dphi_lin(c)
        return g
    if elem_type == HEX27:
        g = np.empty((3, 27))
        for k, (a, b, c) in enumerate(_HEX27_NODES):
            g[0, k] = _dphi_quad(a, xi[0]) * _phi_quad(b, xi[1]) * _phi_quad(c, xi[2])
            g[1, k] = _phi_quad(a, xi[0]) * _dphi_quad(b, xi[1]) * _phi_quad(c, xi[2])
            g[2, k] = _phi_quad(a, xi[0]) * _phi_quad(b, xi[1]) * _dphi_quad(c, xi[2])
        return g
    if elem_type == HEX20:  # serendipity hexahedron, hexahedron.rs:467-545 (corner / edge formulas kept as written there)
        g = np.empty((3, 20))
        x0, x1, x2 = xi
        for k, (al, be, ga) in enumerate(_HEX27_NODES[:20]):
            gg = (1.0 + al * x0) * (1.0 + be * x1) * (1.0 + ga * x2)
            if k < 8:
                f = al * x0 + be * x1 + ga * x2 - 2.0
                s = 1.0 / 8.0
                g[0, k] = s * (al * gg + f * al * (1.0 + be * x1) * (1.0 + ga * x2))
                g[1, k] = s * (be * gg + f * be * (1.0 + al * x0) * (1.0 + ga * x2))
                g[2, k] = s * (ga * gg + f * ga * (1.0 + al * x0) * (1.0 + be * x1))
            else:
                a2, b2, c2 = al * al, be * be, ga * ga
                h = (1.0 - (1.0 - a2) * x0 * x0) * (1.0 - (1.0 - b2) * x1 * x1) * (1.0 - (1.0 - c2) * x2 * x2)
                s = 1.0 / 4.0
                dh0 = -2.0 * (1.0 - a2) * x0 * (1.0 - (1.0 - b2) * x1 * x1) * (1.0 - (1.0 - c2) * x2 * x2)
                dh1 = -2.0 * (1.0 - b2) * x1 * (1.0 - (1.0 - a2) * x0 * x0) * (1.0 - (1.0 - c2) * x2 * x2)
                dh2 = -2.0 * (1.0 - c2) * x2 * (1.0 - (1.0 - a2) * x0 * x0) * (1.0 - (1.0 - b2) * x1 * x1)
                g[0, k] = s * (dh0 * gg + h * al * (1.0 + be * x1) * (1.0 + ga * x2))
                g[1, k] = s * (dh1 * gg + h * be * (1.0 + al * x0) * (1.0 + ga * x2))
                g[2, k] = s * (dh2 * gg + h * ga * (1.0 + al * x0) * (1.0 + be * x1))
        return g
    raise NotImplementedError(elem_type)


def geometry_gradients(elem_type: int, xi: Sequence[float]) -> np.ndarray:
    """Gradients of the GEOMETRY basis: Hex27/Hex20 use the embedded Hex8
    (hexahedron.rs:318-335,546-563), Tet10 the embedded Tet4
    (tetrahedron.rs:226-246); the others are isoparametric."""
    if elem_type in (HEX27, HEX20):
        return reference_gradients(HEX8, xi)
    if elem_type == TET10:
        return reference_gradients(TET4, xi)
    return reference_gradients(elem_type, xi)


def reference_jacobian(elem_type: int, X: np.ndarray, xi: Sequence[float]) -> np.ndarray:
    """J = X * G^T with X (d x n_geom) holding the geometry vertices as columns.

    hexahedron.rs:101-107, tetrahedron.rs:584-590, quadrilateral.rs:125-132.
    The products are accumulated in node order, like nalgebra's small gemm.
    """
    G = geometry_gradients(elem_type, xi)
    d, ng = G.shape
    J = np.zeros((d, d))
    for i in range(d):
        for j in range(d):
            acc = 0.0
            for a in range(ng):
                acc += X[i, a] * G[j, a]
            J[i, j] = acc
    return J


# --------------------------------------------------------------------------
# nalgebra 0.32.1 closed forms (third party; see module docstring)
# --------------------------------------------------------------------------
def det_small(m: np.ndarray) -> float:
    """nalgebra src/linalg/determinant.rs (2x2, 3x3 closed forms)."""
    if m.shape == (2, 2):
        return m[0, 0] * m[1, 1] - m[1, 0] * m[0, 1]
    if m.shape == (3, 3):
        m11, m12, m13 = m[0]
        m21, m22, m23 = m[1]
        m31, m32, m33 = m[2]
        minor_m12_m23 = m22 * m33 - m32 * m23
        minor_m11_m23 = m21 * m33 - m31 * m23
        minor_m11_m22 = m21 * m32 - m31 * m22
        return m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22
    raise ValueError(m.shape)


def try_inverse_small(m: np.ndarray) -> Optional[np.ndarray]:
    """nalgebra src/linalg/inverse.rs (2x2, 3x3 closed forms); None if det == 0."""
    det = det_small(m)
    if det == 0.0:
        return None
    if m.shape == (2, 2):
        return np.array([[m[1, 1] / det, -m[0, 1] / det], [-m[1, 0] / det, m[0, 0] / det]])
    m11, m12, m13 = m[0]
    m21, m22, m23 = m[1]
    m31, m32, m33 = m[2]
    inv = np.empty((3, 3))
    inv[0, 0] = (m22 * m33 - m32 * m23) / det
    inv[0, 1] = (m13 * m32 - m33 * m12) / det
    inv[0, 2] = (m12 * m23 - m22 * m13) / det
    inv[1, 0] = -(m21 * m33 - m31 * m23) / det
    inv[1, 1] = (m11 * m33 - m31 * m13) / det
    inv[1, 2] = (m13 * m21 - m23 * m11) / det
    inv[2, 0] = (m21 * m32 - m31 * m22) / det
    inv[2, 1] = (m12 * m31 - m32 * m11) / det
    inv[2, 2] = (m11 * m22 - m21 * m12) / det
    return inv


# --------------------------------------------------------------------------
# Operators
# --------------------------------------------------------------------------
def lame_from_young_poisson(young: float, poisson: float) -> Tuple[float, float]:
    """(mu, lambda). fenris-solid/src/materials.rs:31-43."""
    mu = 0.5 * young / (1.0 + poisson)
    lam = 2.0 * mu * poisson / (1.0 - 2.0 * poisson)
    return mu, lam


def linear_elastic_energy_density(F: np.ndarray, mu: float, lam: float) -> float:
    """fenris-solid/src/materials.rs:80-84 (eps = sym(F) - I)."""
    eps = 0.5 * (F + F.T) - np.eye(F.shape[0])
    return mu * float(np.sum(eps * eps)) + 0.5 * lam * float(np.trace(eps)) ** 2


def linear_elastic_stress(F: np.ndarray, mu: float, lam: float) -> np.ndarray:
    """fenris-solid/src/materials.rs:86-95."""
    d = F.shape[0]
    eps = 0.5 * (F + F.T) - np.eye(d)
    return eps * 2.0 * mu + np.eye(d) * (lam * np.trace(eps))


def contract(op: int, a: np.ndarray, b: np.ndarray, params: Tuple[float, ...]) -> np.ndarray:
    """The s x s contraction C(a, b) of one pair of physical basis gradients.

    Laplace: a . b (1x1), src/assembly/operators/laplace.rs:60-68.
    Linear elasticity: (I (a.b) + b a^T) mu + a b^T lambda,
    fenris-solid/src/materials.rs:108-122 (F is ignored).
    """
    if op == LAPLACE:
        return np.array([[float(np.dot(a, b))]])
    if op == LINEAR_ELASTIC:
        mu, lam = params
        d = a.shape[0]
        return (np.eye(d) * float(np.dot(a, b)) + np.outer(b, a)) * mu + np.outer(a, b) * lam
    raise ValueError(op)


def solution_dim(op: int, d: int) -> int:
    return 1 if op == LAPLACE else d


# --------------------------------------------------------------------------
# Element matrix  (src/assembly/local/elliptic.rs:361-439)
# --------------------------------------------------------------------------
class SingularJacobian(Exception):
    """`Err("Singular element Jacobian encountered")`, elliptic.rs:401-404."""


def element_matrix(
    elem_type: int,
    X_elem: np.ndarray,
    op: int,
    weights: np.ndarray,
    points: np.ndarray,
    params_per_point: Sequence[Tuple[float, ...]],
) -> np.ndarray:
    """K_e, shape (s n, s n), local dof = s*I + i.

    X_elem: (n, d) coordinates of ALL element nodes in local order (the
    geometry uses the first n_geom of them).  Literal restatement of
    assemble_element_elliptic_matrix (elliptic.rs:361-439) with the default
    symmetric block loop (operators.rs:176-188 / fenris-solid lib.rs:381-391)
    and the element-level mirror (util.rs:38-50).
    """
    n, ng, d = element_info(elem_type)
    s = solution_dim(op, d)
    X = np.asarray(X_elem, dtype=np.float64)[:ng].T  # d x n_geom
    K = np.zeros((s * n, s * n))
    for w, xi, par in zip(weights, points, params_per_point):
        J = reference_jacobian(elem_type, X, xi)
        j_det = det_small(J)
        j_inv = try_inverse_small(J)
        if j_inv is None:
            raise SingularJacobian("Singular element Jacobian encountered")
        j_inv_t = j_inv.T
        G = reference_gradients(elem_type, xi)
        G = j_inv_t @ G  # elliptic.rs:415-418, column by column
        scale = w * abs(j_det)  # elliptic.rs:422
        for Jn in range(n):
            for In in range(min(Jn + 1, n)):
                c = contract(op, G[:, In], G[:, Jn], par)
                K[s * In:s * In + s, s * Jn:s * Jn + s] += c * scale
    # clone_upper_to_lower, util.rs:38-50
    for j in range(s * n):
        for i in range(j + 1, s * n):
            K[i, j] = K[j, i]
    return K


# --------------------------------------------------------------------------
# Pattern + serial scatter  (src/assembly/global.rs)
# --------------------------------------------------------------------------
def assemble_pattern(sdim: int, num_nodes: int, elements: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """(row_offsets, col_indices) exactly as CsrAssembler::assemble_pattern.

    src/assembly/global.rs:65-120.  `elements` may be ragged, may contain
    empty elements and repeated nodes (KAT tests/unit_tests/assembly/global.rs:100-138).
    """
    node_sets: List[set] = [set() for _ in range(num_nodes)]
    for nodes in elements:
        for ni in nodes:
            for nj in nodes:
                node_sets[ni].add(nj)
    offsets = [0]
    cur = 0
    for ns in node_sets:
        for _ in range(sdim):
            cur += sdim * len(ns)
            offsets.append(cur)
    cols: List[int] = []
    for ns in node_sets:
        buf = sorted(ns)
        for _ in range(sdim):
            for nj in buf:
                for j in range(sdim):
                    cols.append(sdim * nj + j)
    assert offsets[-1] == len(cols)
    return np.array(offsets, dtype=np.uint64), np.array(cols, dtype=np.uint64)


def add_element_row_to_csr_row(row_values, row_cols, nodes, perm, dim, local_row) -> None:
    """src/assembly/global.rs:504-537: cursor walk through the sorted CSR row."""
    cursor = 0
    ncols = len(row_cols)
    for node_local in perm:
        node_global = nodes[node_local]
        for i in range(dim):
            local_col = dim * node_local + i
            global_col = dim * node_global + i
            while cursor < ncols and row_cols[cursor] != global_col:
                cursor += 1
            if cursor >= ncols:
                raise IndexError("Could not find column index associated with node in CSR row")
            row_values[cursor] += local_row[local_col]
            cursor += 1


def scatter_element(values, row_offsets, col_indices, sdim, nodes, K) -> None:
    """Per-element body of assemble_into_csr, src/assembly/global.rs:155-178."""
    perm = sorted(range(len(nodes)), key=lambda i: nodes[i])
    for local_node, global_node in enumerate(nodes):
        for i in range(sdim):
            lrow = sdim * local_node + i
            grow = sdim * global_node + i
            b, e = int(row_offsets[grow]), int(row_offsets[grow + 1])
            add_element_row_to_csr_row(values[b:e], col_indices[b:e], nodes, perm, sdim, K[lrow, :])


class Problem:
    """The four borrowed ingredients of ElementEllipticAssembler (elliptic.rs:153-158):
    space (mesh), operator, uniform quadrature table (+ per-point data); u = 0."""

    def __init__(self, elem_type, vertices, connectivity, op, weights=None, points=None, params=None):
        self.elem_type = elem_type
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.connectivity = np.ascontiguousarray(connectivity, dtype=np.int64)
        self.op = op
        if weights is None:
            weights, points = canonical_stiffness_rule(elem_type)
        self.weights = np.asarray(weights, dtype=np.float64)
        self.points = np.asarray(points, dtype=np.float64)
        n, ng, d = element_info(elem_type)
        self.n, self.ng, self.d = n, ng, d
        self.sdim = solution_dim(op, d)
        if params is None:
            params = ()
        params = tuple(params)
        # UniformQuadratureTable::with_uniform_data: same data at every point
        # (quadrature_table.rs:264-266).  A list of per-point tuples is also accepted.
        if len(params) > 0 and isinstance(params[0], (tuple, list)):
            self.params_per_point = [tuple(p) for p in params]
        else:
            self.params_per_point = [params] * len(self.weights)

    @property
    def num_nodes(self):
        return self.vertices.shape[0]

    @property
    def num_elements(self):
        return self.connectivity.shape[0]

    def element_matrix(self, e: int) -> np.ndarray:
        nodes = self.connectivity[e]
        return element_matrix(self.elem_type, self.vertices[nodes], self.op, self.weights, self.points,
                              self.params_per_point)


def assemble_serial(problem: Problem, pattern=None, values=None):
    """CsrAssembler::assemble / assemble_into_csr (global.rs:124-182), literal."""
    if pattern is None:
        pattern = assemble_pattern(problem.sdim, problem.num_nodes, problem.connectivity.tolist())
    row_offsets, col_indices = pattern
    if values is None:
        values = np.zeros(len(col_indices))
    for e in range(problem.num_elements):
        K = problem.element_matrix(e)
        scatter_element(values, row_offsets, col_indices, problem.sdim, problem.connectivity[e].tolist(), K)
    return row_offsets, col_indices, values


def assemble_colored(problem: Problem, colors, pattern=None, values=None):
    """CsrParAssembler::assemble_into_csr (global.rs:314-376): colours in order,
    elements of a colour in label order (any order gives the same sums because
    a colour's elements touch disjoint rows)."""
    if pattern is None:
        pattern = assemble_pattern(problem.sdim, problem.num_nodes, problem.connectivity.tolist())
    row_offsets, col_indices = pattern
    if values is None:
        values = np.zeros(len(col_indices))
    for color in colors:
        for e in color:
            K = problem.element_matrix(e)
            scatter_element(values, row_offsets, col_indices, problem.sdim, problem.connectivity[e].tolist(), K)
    return row_offsets, col_indices, values


# --------------------------------------------------------------------------
# Vectorised variant (same math, batched over elements) for larger meshes.
# --------------------------------------------------------------------------
def element_matrices_fast(problem: Problem) -> np.ndarray:
    """All K_e at once, shape (E, s n, s n).  Same formula and the same
    upper-triangle-then-mirror rule; only the summation order over nodes inside
    J = X G^T and over quadrature points may differ in rounding."""
    et, n, ng, d, s = problem.elem_type, problem.n, problem.ng, problem.d, problem.sdim
    conn = problem.connectivity
    E = conn.shape[0]
    X = problem.vertices[conn[:, :ng]]  # E, ng, d
    K = np.zeros((E, s * n, s * n))
    for w, xi, par in zip(problem.weights, problem.points, problem.params_per_point):
        Gg = geometry_gradients(et, xi)  # d, ng
        J = np.einsum("eai,ja->eij", X, Gg)  # J[i,j] = sum_a X[a,i] G[j,a]
        if d == 2:
            det = J[:, 0, 0] * J[:, 1, 1] - J[:, 1, 0] * J[:, 0, 1]
        else:
            det = (J[:, 0, 0] * (J[:, 1, 1] * J[:, 2, 2] - J[:, 2, 1] * J[:, 1, 2])
                   - J[:, 0, 1] * (J[:, 1, 0] * J[:, 2, 2] - J[:, 2, 0] * J[:, 1, 2])
                   + J[:, 0, 2] * (J[:, 1, 0] * J[:, 2, 1] - J[:, 2, 0] * J[:, 1, 1]))
        if np.any(det == 0.0):
            raise SingularJacobian("Singular element Jacobian encountered")
        Jinv = np.linalg.inv(J)
        Gr = reference_gradients(et, xi)  # d, n
        G = np.einsum("eji,jn->ein", Jinv, Gr)  # J^{-T} G_ref
        scale = w * np.abs(det)
        dots = np.einsum("eia,eib->eab", G, G)
        if problem.op == LAPLACE:
            K += scale[:, None, None] * dots
        else:
            mu, lam = par
            # block (a,b) entry (i,j): mu (g_a.g_b delta_ij + g_b^i g_a^j) + lam g_a^i g_b^j
            outer = np.einsum("eia,ejb->eaibj", G, G)  # g_a^i g_b^j
            Kq = lam * outer + mu * np.transpose(outer, (0, 1, 4, 3, 2))
            eye = np.eye(d)
            Kq = Kq + mu * dots[:, :, None, :, None] * eye[None, None, :, None, :]
            K += scale[:, None, None] * Kq.reshape(E, s * n, s * n)
    iu = np.triu_indices(s * n, 1)
    # block-upper-triangle rule: take everything on/above the scalar diagonal, mirror it
    Ku = np.triu(K)
    K = Ku + np.transpose(np.triu(K, 1), (0, 2, 1))
    del iu
    return K


def assemble_fast(problem: Problem, pattern=None):
    """Vectorised global assembly on the exact reference pattern."""
    import scipy.sparse as sp

    if pattern is None:
        pattern = assemble_pattern_fast(problem.sdim, problem.num_nodes, problem.connectivity)
    row_offsets, col_indices = pattern
    s, n = problem.sdim, problem.n
    K = element_matrices_fast(problem)
    dofs = (problem.connectivity[:, :, None] * s + np.arange(s)[None, None, :]).reshape(-1, s * n)
    rows = np.repeat(dofs, s * n, axis=1).ravel()
    cols = np.tile(dofs, (1, s * n)).ravel()
    nrows = s * problem.num_nodes
    A = sp.coo_matrix((K.ravel(), (rows, cols)), shape=(nrows, nrows)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    # put onto the reference pattern (which may contain entries A lacks only if K has exact zeros dropped - it does not drop)
    P = sp.csr_matrix((np.zeros(len(col_indices)), col_indices.astype(np.int64), row_offsets.astype(np.int64)),
                      shape=(nrows, nrows))
    assert np.array_equal(A.indptr, P.indptr) and np.array_equal(A.indices, P.indices), "pattern mismatch"
    return row_offsets, col_indices, A.data.copy()


def assemble_pattern_fast(sdim: int, num_nodes: int, connectivity: np.ndarray):
    """Same output as assemble_pattern for uniform connectivity, via sort-unique."""
    conn = np.asarray(connectivity, dtype=np.int64)
    n = conn.shape[1]
    I = np.repeat(conn, n, axis=1).ravel()
    J = np.tile(conn, (1, n)).ravel()
    key = np.unique(I * np.int64(num_nodes) + J)
    bi = key // num_nodes
    bj = key % num_nodes
    counts = np.bincount(bi, minlength=num_nodes)
    row_len = np.repeat(counts * sdim, sdim)
    row_offsets = np.concatenate([[0], np.cumsum(row_len)]).astype(np.uint64)
    # per node: block cols, expanded by sdim, repeated sdim times
    starts = np.concatenate([[0], np.cumsum(counts)])
    cols = np.empty(int(row_offsets[-1]), dtype=np.uint64)
    exp = (bj[:, None] * sdim + np.arange(sdim)[None, :]).ravel()  # per block sdim entries, node-major
    pos = 0
    estarts = starts * sdim
    for node in range(num_nodes):
        seg = exp[estarts[node]:estarts[node + 1]]
        L = len(seg)
        for _ in range(sdim):
            cols[pos:pos + L] = seg
            pos += L
    return row_offsets, cols


# --------------------------------------------------------------------------
# Greedy colouring  (fenris-paradis/src/coloring.rs:6-70)
# --------------------------------------------------------------------------
def sequential_greedy_coloring(elements: Sequence[Sequence[int]]) -> List[List[int]]:
    """Returns, per colour, the element labels in the order the reference stores them."""
    colors: List[List[int]] = []
    current = list(range(len(elements)))
    last_visited = {}
    c = 0
    while current:
        postponed: List[int] = []
        members: List[int] = []
        for e in current:
            nodes = elements[e]
            blocked = any(last_visited.get(nd, -1) == c for nd in nodes)
            if blocked:
                postponed.append(e)
            else:
                for nd in nodes:
                    last_visited[nd] = c
                members.append(e)
        colors.append(members)
        current = postponed
        c += 1
    return colors


# --------------------------------------------------------------------------
# Procedural meshes  (src/mesh/procedural.rs, src/mesh_convert.rs)
# --------------------------------------------------------------------------
def create_unit_square_uniform_quad_mesh_2d(cells_per_dim: int):
    """src/mesh/procedural.rs:15-20,46-93 (top_left = (0, 1), rows go down)."""
    if cells_per_dim == 0:
        return np.zeros((0, 2)), np.zeros((0, 4), dtype=np.int64)
    cell_size = 1.0 / cells_per_dim
    nx = ny = cells_per_dim
    verts = []
    for j in range(ny + 1):
        for i in range(nx + 1):
            verts.append([0.0 + float(i) * cell_size, 1.0 + (-float(j)) * cell_size])
    idx = lambda i, j: (nx + 1) * j + i
    cells = []
    for j in range(ny):
        for i in range(nx):
            cells.append([idx(i, j + 1), idx(i + 1, j + 1), idx(i + 1, j), idx(i, j)])
    return np.array(verts), np.array(cells, dtype=np.int64)


def create_rectangular_uniform_hex_mesh(unit_length: float, ux: int, uy: int, uz: int, cells_per_unit: int):
    """src/mesh/procedural.rs:216-277."""
    if cells_per_unit == 0 or ux == 0 or uy == 0:
        return np.zeros((0, 3)), np.zeros((0, 8), dtype=np.int64)
    h = unit_length / cells_per_unit
    cx, cy, cz = ux * cells_per_unit, uy * cells_per_unit, uz * cells_per_unit
    vx, vy, vz = cx + 1, cy + 1, cz + 1
    k, j, i = np.meshgrid(np.arange(vz), np.arange(vy), np.arange(vx), indexing="ij")
    verts = np.stack([i.ravel() * h, j.ravel() * h, k.ravel() * h], axis=1).astype(np.float64)
    idx = lambda i, j, k: (vx * vy) * k + vx * j + i
    k, j, i = np.meshgrid(np.arange(cz), np.arange(cy), np.arange(cx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cells = np.stack([idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k),
                      idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)],
                     axis=1).astype(np.int64)
    return verts, cells


def create_unit_box_uniform_hex_mesh_3d(cells_per_dim: int):
    """src/mesh/procedural.rs:30-35."""
    return create_rectangular_uniform_hex_mesh(1.0, 1, 1, 1, cells_per_dim)


_POS_FACE_DELTAS = [  # src/mesh/procedural.rs:331-335
    [[1, 0, 1], [1, 1, 1], [1, 1, 0], [1, 0, 0]],
    [[0, 1, 0], [1, 1, 0], [1, 1, 1], [0, 1, 1]],
    [[0, 1, 1], [1, 1, 1], [1, 0, 1], [0, 0, 1]],
]


def create_rectangular_uniform_tet_mesh(unit_length: float, ux: int, uy: int, uz: int, cells_per_unit: int):
    """BCC tet mesh, 12 tets per cell.  src/mesh/procedural.rs:286-403."""
    if ux == 0 or uy == 0 or uz == 0 or cells_per_unit == 0:
        return np.zeros((0, 3)), np.zeros((0, 4), dtype=np.int64)
    h = unit_length / float(cells_per_unit)
    cx, cy, cz = ux * cells_per_unit, uy * cells_per_unit, uz * cells_per_unit
    vx, vy, vz = cx + 1, cy + 1, cz + 1
    verts = []
    for k in range(vz):
        for j in range(vy):
            for i in range(vx):
                verts.append([h * float(i), h * float(j), h * float(k)])
    center_offset = len(verts)
    for k in range(cz):
        for j in range(cy):
            for i in range(cx):
                verts.append([h * (0.5 + float(i)), h * (0.5 + float(j)), h * (0.5 + float(k))])
    vidx = lambda v: (vx * vy) * v[2] + vx * v[1] + v[0]
    cidx = lambda c: (cx * cy) * c[2] + cx * c[1] + c[0] + center_offset
    conn = []
    ncells = [cx, cy, cz]
    for k in range(cz):
        for j in range(cy):
            for i in range(cx):
                cell = [i, j, k]
                for axis in (0, 1, 2):
                    if cell[axis] + 1 < ncells[axis]:
                        # connect_centers_with_tets, :337-357
                        face = [vidx([cell[0] + d[0], cell[1] + d[1], cell[2] + d[2]])
                                for d in _POS_FACE_DELTAS[axis]]
                        c1 = cidx(cell)
                        nb = list(cell)
                        nb[axis] += 1
                        c2 = cidx(nb)
                        cyc = face + [face[0]]
                        for v1, v2 in zip(cyc[:-1], cyc[1:]):
                            conn.append([c1, c2, v2, v1])
                    for positive in (False, True):
                        if (not positive and cell[axis] == 0) or (positive and cell[axis] + 1 == ncells[axis]):
                            # make_pyramid, :359-384
                            fv = [[cell[0] + d[0], cell[1] + d[1], cell[2] + d[2]] for d in _POS_FACE_DELTAS[axis]]
                            if not positive:
                                fv.reverse()
                                for c in fv:
                                    c[axis] -= 1
                            a, b, c, dd = [vidx(v) for v in fv]
                            center = cidx(cell)
                            if (i + j + k) % 2 == 0:
                                conn.append([a, b, c, center])
                                conn.append([a, c, dd, center])
                            else:
                                conn.append([a, b, dd, center])
                                conn.append([b, c, dd, center])
    return np.array(verts, dtype=np.float64), np.array(conn, dtype=np.int64)


def create_unit_box_uniform_tet_mesh_3d(cells_per_dim: int):
    """src/mesh/procedural.rs:37-42."""
    return create_rectangular_uniform_tet_mesh(1.0, 1, 1, 1, cells_per_dim)


_HEX_EDGES = [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)]  # mesh_convert.rs:118-129
_HEX_FACES = [((0, 1, 2, 3), (0.0, 0.0, -1.0)), ((0, 1, 4, 5), (0.0, -1.0, 0.0)), ((0, 3, 4, 7), (-1.0, 0.0, 0.0)),
              ((1, 2, 5, 6), (1.0, 0.0, 0.0)), ((2, 3, 6, 7), (0.0, 1.0, 0.0)), ((4, 5, 6, 7), (0.0, 0.0, 1.0))]  # :145-150
_TET_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (2, 3), (1, 3)]  # mesh_convert.rs:74-79


def _relabel(local_children, vertices_out):
    """Global labelling in first-seen order keyed by the sorted parent set.
    src/mesh_convert.rs:227-330 (child index is always 0 for these refinements)."""
    label = {}
    final_vertices = []
    conn = []
    for elem in local_children:
        row = []
        for parents, coord in elem:
            key = tuple(sorted(parents))
            if key not in label:
                label[key] = len(final_vertices)
                final_vertices.append(coord)
            row.append(label[key])
        conn.append(row)
    return np.array(final_vertices, dtype=np.float64), np.array(conn, dtype=np.int64)


def hex27_mesh_from_hex8(vertices: np.ndarray, hex8: np.ndarray):
    """Hex27Mesh::from(&hex8_mesh).  src/mesh_convert.rs:85-166 + :227-330."""
    out = []
    for nodes in hex8.tolist():
        Xe = vertices[nodes]  # 8 x 3
        elem = [((g,), Xe[l].copy()) for l, g in enumerate(nodes)]
        for a, b in _HEX_EDGES:
            # lerp(t=0.5): nalgebra axpy -> 0.5*b + 0.5*a
            elem.append(((nodes[a], nodes[b]), 0.5 * Xe[b] + 0.5 * Xe[a]))
        for face, ref in _HEX_FACES:
            N = hex8_basis(ref)
            elem.append((tuple(nodes[f] for f in face), Xe.T @ N))
        elem.append((tuple(nodes), Xe.T @ hex8_basis((0.0, 0.0, 0.0))))
        out.append(elem)
    return _relabel(out, None)


def hex20_mesh_from_hex8(vertices: np.ndarray, hex8: np.ndarray):
    """Hex20Mesh::from(&hex8_mesh).  src/mesh_convert.rs:168-217 + :227-330 (vertices, then the 12 edge midpoints)."""
    out = []
    for nodes in hex8.tolist():
        Xe = vertices[nodes]
        elem = [((g,), Xe[l].copy()) for l, g in enumerate(nodes)]
        for a, b in _HEX_EDGES:
            elem.append(((nodes[a], nodes[b]), 0.5 * Xe[b] + 0.5 * Xe[a]))
        out.append(elem)
    return _relabel(out, None)


def tet10_mesh_from_tet4(vertices: np.ndarray, tet4: np.ndarray):
    """Tet10Mesh::from(&tet4_mesh).  src/mesh_convert.rs:42-83 + :227-330."""
    out = []
    for nodes in tet4.tolist():
        Xe = vertices[nodes]
        elem = [((g,), Xe[l].copy()) for l, g in enumerate(nodes)]
        for a, b in _TET_EDGES:
            elem.append(((nodes[a], nodes[b]), 0.5 * Xe[b] + 0.5 * Xe[a]))
        out.append(elem)
    return _relabel(out, None)


# --------------------------------------------------------------------------
# Helpers used by tests / bench
# --------------------------------------------------------------------------
def rel_frobenius(a: np.ndarray, b: np.ndarray) -> float:
    """||a - b||_F / ||b||_F on identical patterns (SURVEY 8d parity metric)."""
    nb = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / (nb if nb > 0 else 1.0)


def jitter_vertices(vertices: np.ndarray, h: float, seed: int = 12345, amp: float = 0.2) -> np.ndarray:
    """Robustness input of SURVEY 8d: every vertex moved by U(-amp h, amp h) (PCG64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return vertices + rng.uniform(-amp * h, amp * h, size=vertices.shape)


# ------------------------------------------------------------------------------------------------------------------------
# element mass matrix, source vector and the global vector assembler (test infrastructure, like the rest)
# ------------------------------------------------------------------------------------------------------------------------
def basis_values(elem_type: int, xi: Sequence[float]) -> np.ndarray:
    """Reference basis phi(xi), shape (n,).  Quad4 quadrilateral.rs:79-91, Tet4 tetrahedron.rs:551-558,
    Tet10 tetrahedron.rs:179-196, Hex8 hexahedron.rs:43-60, Hex27 hexahedron.rs:223-268."""
    if elem_type == QUAD4:
        return np.array([_phi_lin(a, xi[0]) * _phi_lin(b, xi[1]) for a, b in _QUAD4_NODES])
    if elem_type == TET4:
        return tet4_basis(xi)
    if elem_type == TET10:
        psi = tet4_basis(xi)
        return np.array([psi[i] * (2.0 * psi[i] - 1.0) for i in range(4)] + [4.0 * psi[i] * psi[j] for i, j in _TET10_EDGES])
    if elem_type == HEX8:
        return hex8_basis(xi)
    if elem_type == HEX27:
        return np.array([_phi_quad(a, xi[0]) * _phi_quad(b, xi[1]) * _phi_quad(c, xi[2]) for a, b, c in _HEX27_NODES])
    if elem_type == HEX20:  # hexahedron.rs:414-465
        x0, x1, x2 = xi
        out = []
        for k, (al, be, ga) in enumerate(_HEX27_NODES[:20]):
            if k < 8:
                out.append((1.0 / 8.0) * (1.0 + al * x0) * (1.0 + be * x1) * (1.0 + ga * x2) * (al * x0 + be * x1 + ga * x2 - 2.0))
            else:
                out.append((1.0 / 4.0) * (1.0 - (1.0 - al * al) * x0 * x0) * (1.0 - (1.0 - be * be) * x1 * x1) * (1.0 - (1.0 - ga * ga) * x2 * x2)
                           * (1.0 + al * x0) * (1.0 + be * x1) * (1.0 + ga * x2))
        return np.array(out)
    raise NotImplementedError(elem_type)


def geometry_basis_values(elem_type: int, xi: Sequence[float]) -> np.ndarray:
    """Basis of the GEOMETRY map: Hex27 -> embedded Hex8 (hexahedron.rs:318-335), Tet10 -> Tet4 (tetrahedron.rs:226-246)."""
    if elem_type in (HEX27, HEX20):
        return basis_values(HEX8, xi)
    if elem_type == TET10:
        return basis_values(TET4, xi)
    return basis_values(elem_type, xi)


def map_reference_coords(elem_type: int, X: np.ndarray, xi: Sequence[float]) -> np.ndarray:
    """x(xi) = sum_a phi_a^geo(xi) X_a (FiniteElement::map_reference_coords, e.g. quadrilateral.rs:117-122); X: (n, d) rows."""
    _, ng, _ = element_info(elem_type)
    return geometry_basis_values(elem_type, xi) @ np.asarray(X)[:ng]


def element_mass_matrix(elem_type: int, X: np.ndarray, weights, points, density, s: int) -> np.ndarray:
    """assemble_element_mass_matrix, src/assembly/local/mass.rs:218-286 (literal): per point scale = w |det J| rho, upper
    triangle of blocks m_IJ I_s with m_IJ += scale * phi_I * phi_J, then clone_upper_to_lower (util.rs:38-50)."""
    n, _, _ = element_info(elem_type)
    M = np.zeros((s * n, s * n))
    for w, xi, rho in zip(weights, points, density):
        j_det = det_small(reference_jacobian(elem_type, np.asarray(X).T, xi))  # X: (n, d) rows
        phi = basis_values(elem_type, xi)
        scale = w * abs(j_det) * rho
        for I in range(n):
            for J in range(I, n):
                m = scale * phi[I] * phi[J]
                for i in range(s):
                    M[s * I + i, s * J + i] += m
    iu = np.triu_indices(s * n, 1)
    M[(iu[1], iu[0])] = M[iu]
    return M


def element_source_vector(elem_type: int, X: np.ndarray, weights, points, fvals: np.ndarray) -> np.ndarray:
    """assemble_element_source_vector, src/assembly/local/source.rs:217-278: output (s x n, column = node) += w |det J| f phi^T;
    fvals[q] = source.evaluate(x_q, data_q) (shape (q, s)).  Returned as the length s*n vector (node-major)."""
    n, _, _ = element_info(elem_type)
    fvals = np.asarray(fvals, dtype=np.float64)
    s = fvals.shape[1]
    out = np.zeros((s, n))
    for w, xi, f in zip(weights, points, fvals):
        phi = basis_values(elem_type, xi)
        j = reference_jacobian(elem_type, np.asarray(X).T, xi)  # X: (n, d) rows
        out += (w * abs(det_small(j))) * np.outer(f, phi)
    return out.T.reshape(-1)


def assemble_mass_serial(elem_type: int, vertices, connectivity, weights, points, density, s: int):
    """CsrAssembler::assemble with an ElementMassAssembler (global.rs:124-182, mass.rs:127-159): (row_offsets, col_indices, values)."""
    conn = np.asarray(connectivity, dtype=np.int64)
    ro, ci = assemble_pattern(s, len(vertices), conn.tolist())
    values = np.zeros(len(ci))
    for e in range(len(conn)):
        M = element_mass_matrix(elem_type, np.asarray(vertices)[conn[e]], weights, points, density, s)
        scatter_element(values, ro, ci, s, conn[e].tolist(), M)
    return ro, ci, values


def assemble_mass_fast(elem_type: int, vertices, connectivity, weights, points, density, s: int):
    """Vectorised variant of assemble_mass_serial (same sums up to fp reassociation); cross-checked against it in the tests."""
    conn = np.asarray(connectivity, dtype=np.int64)
    V = np.asarray(vertices, dtype=np.float64)
    E, n = conn.shape
    X = V[conn]
    m = np.zeros((E, n, n))
    for w, xi, rho in zip(weights, points, density):
        G = geometry_gradients(elem_type, xi)
        ng = G.shape[1]
        J = np.einsum("eai,ja->eij", X[:, :ng], G)
        phi = basis_values(elem_type, xi)
        m += (w * rho * np.abs(np.linalg.det(J)))[:, None, None] * np.outer(phi, phi)[None]
    ro, ci = assemble_pattern_fast(s, len(V), conn)
    rows = (s * conn[:, :, None, None] + np.arange(s)[None, None, None, :]) + 0 * conn[:, None, :, None]
    cols = (s * conn[:, None, :, None] + np.arange(s)[None, None, None, :]) + 0 * conn[:, :, None, None]
    vals = np.broadcast_to(m[:, :, :, None], rows.shape)
    # place the entries on the pattern: (row, col) keys are strictly increasing along the CSR arrays
    ncols = s * len(V)
    row_of = np.repeat(np.arange(len(ro) - 1, dtype=np.int64), np.diff(ro.astype(np.int64)))
    keys = row_of * ncols + ci.astype(np.int64)
    idx = np.searchsorted(keys, rows.ravel().astype(np.int64) * ncols + cols.ravel().astype(np.int64))
    assert np.array_equal(keys[idx], rows.ravel() * ncols + cols.ravel())
    values = np.zeros(len(ci))
    np.add.at(values, idx, vals.ravel())
    return ro, ci, values


def physical_quadrature_points(elem_type: int, vertices, connectivity, points) -> np.ndarray:
    """x_q of every element (E, q, d): what an ElementSourceAssembler hands to SourceFunction::evaluate (source.rs:253-255)."""
    conn = np.asarray(connectivity, dtype=np.int64)
    _, ng, _ = element_info(elem_type)
    X = np.asarray(vertices, dtype=np.float64)[conn[:, :ng]]
    N = np.stack([geometry_basis_values(elem_type, xi) for xi in points])  # q, ng
    return np.einsum("qa,ead->eqd", N, X)


def assemble_vector_serial(elem_type: int, vertices, connectivity, weights, points, fvals: np.ndarray) -> np.ndarray:
    """VectorAssembler::assemble_vector (global.rs:569-617) with an ElementSourceAssembler: fvals (E, q, s) or (q, s) (uniform);
    add_local_to_global (global.rs:779-796): global[s I + i] += local[s a + i]."""
    conn = np.asarray(connectivity, dtype=np.int64)
    fvals = np.asarray(fvals, dtype=np.float64)
    s = fvals.shape[-1]
    out = np.zeros(s * len(vertices))
    for e in range(len(conn)):
        fe = fvals[e] if fvals.ndim == 3 else fvals
        local = element_source_vector(elem_type, np.asarray(vertices)[conn[e]], weights, points, fe)
        for a, I in enumerate(conn[e]):
            out[s * I:s * I + s] += local[s * a:s * a + s]
    return out


def assemble_vector_fast(elem_type: int, vertices, connectivity, weights, points, fvals: np.ndarray) -> np.ndarray:
    conn = np.asarray(connectivity, dtype=np.int64)
    V = np.asarray(vertices, dtype=np.float64)
    fvals = np.asarray(fvals, dtype=np.float64)
    s = fvals.shape[-1]
    E, n = conn.shape
    X = V[conn]
    out = np.zeros((len(V), s))
    for q, (w, xi) in enumerate(zip(weights, points)):
        G = geometry_gradients(elem_type, xi)
        J = np.einsum("eai,ja->eij", X[:, :G.shape[1]], G)
        scale = w * np.abs(np.linalg.det(J))  # E
        f = fvals[:, q] if fvals.ndim == 3 else np.broadcast_to(fvals[q], (E, s))
        phi = basis_values(elem_type, xi)
        np.add.at(out, conn, scale[:, None, None] * phi[None, :, None] * f[:, None, :])
    return out.reshape(-1)


def apply_homogeneous_dirichlet_bc_csr(row_offsets, col_indices, values, nodes, solution_dim: int) -> float:
    """Literal restatement of apply_homogeneous_dirichlet_bc_csr (src/assembly/global.rs:379-451); modifies `values` in place and
    returns the diagonal scale.  (The reference sizes its two flag vectors d * nrows, :410-411 - more than needed; nrows here.)"""
    ro = np.asarray(row_offsets, dtype=np.int64)
    ci = np.asarray(col_indices, dtype=np.int64)
    d = solution_dim
    nrows = len(ro) - 1
    scale = 1.0
    for i in range(nrows):  # triplet_iter is row-major: the first non-zero DIAGONAL entry (:388-397)
        hit = False
        for k in range(ro[i], ro[i + 1]):
            if ci[k] == i and values[k] != 0.0:
                scale, hit = abs(values[k]), True
                break
        if hit:
            break
    member = np.zeros(nrows, dtype=bool)
    visit = np.zeros(nrows, dtype=bool)
    for node in nodes:
        for i in range(d):
            r = d * int(node) + i
            member[r] = True
            for k in range(ro[r], ro[r + 1]):
                if ci[k] == r:
                    values[k] = scale
                else:
                    values[k] = 0.0
                    visit[ci[k]] = True
    for r in np.nonzero(visit)[0]:
        if not member[r]:
            for k in range(ro[r], ro[r + 1]):
                if member[ci[k]]:
                    values[k] = 0.0
    return scale


def apply_homogeneous_dirichlet_bc_rhs(rhs, nodes, solution_dim: int) -> None:
    """src/assembly/global.rs:479-495."""
    for node in nodes:
        rhs[solution_dim * int(node):solution_dim * int(node) + solution_dim] = 0.0


def conjugate_gradient(apply_a, b, x0=None, rel_tol: float = 1e-8, max_iter: int = 0, apply_p=None):
    """Literal restatement of ConjugateGradient::solve_with_guess (fenris-sparse/src/cg.rs:364-480) with
    RelativeResidualCriterion (cg.rs:85-124).  apply_a / apply_p: callables y = A x / z = P r (identity if None).
    Returns (x, iterations, status) with status in {"ok", "max_iter", "indefinite_operator", "indefinite_preconditioner"}."""
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
    apply_p = apply_p or (lambda r: r.copy())
    r = b - apply_a(x)
    z = apply_p(r)
    p = z.copy()
    zTr = float(z @ r)
    b_norm = float(np.linalg.norm(b))
    if b_norm == 0.0:
        return np.zeros_like(b), 0, "ok"
    it = 0
    while True:
        if np.linalg.norm(r) <= rel_tol * b_norm:
            return x, it, "ok"
        if max_iter and it >= max_iter:
            return x, it, "max_iter"
        Ap = apply_a(p)
        pAp = float(p @ Ap)
        if pAp <= 0.0:
            return x, it, "indefinite_operator"
        if zTr <= 0.0:
            return x, it, "indefinite_preconditioner"
        alpha = zTr / pAp
        x = x + alpha * p
        r = r - alpha * Ap
        it += 1
        z = apply_p(r)
        zTr_next = float(z @ r)
        p = p * (zTr_next / zTr) + z
        zTr = zTr_next


def assemble_with_quadrature_table(elem_type: int, vertices, connectivity, op: int, rules, element_to_rule, u=None):
    """CsrAssembler::assemble with an ElementEllipticAssembler whose QuadratureTable has a rule per element
    (CompactQuadratureTable / GeneralQuadratureTable, quadrature_table.rs:57-210, 312-439; the element loop of elliptic.rs:299-340 calls
    populate_element_quadrature_from_table per element).  rules[r] = (weights, points, params_per_point); u: the state of a
    state-dependent operator (STVK, NEO_HOOKEAN), zeros when omitted."""
    conn = np.asarray(connectivity, dtype=np.int64)
    V = np.asarray(vertices, dtype=np.float64)
    n, _, d = element_info(elem_type)
    s = solution_dim(op, d)
    ro, ci = assemble_pattern(s, len(V), conn.tolist())
    values = np.zeros(len(ci))
    for e in range(len(conn)):
        w, p, params = rules[int(element_to_rule[e])]
        if op in (LAPLACE, LINEAR_ELASTIC):
            K = element_matrix(elem_type, V[conn[e]], op, w, p, params)
        else:
            ue = np.zeros(n * s) if u is None else np.asarray(u, dtype=np.float64).reshape(-1, s)[conn[e]].reshape(-1)
            K = element_matrix_u(elem_type, V[conn[e]], op, ue, w, p, params)
        scatter_element(values, ro, ci, s, conn[e].tolist(), K)
    return ro, ci, values


# ------------------------------------------------------------------------------------------------------------------------
# ElementEllipticAssembler as vector / scalar assembler (elliptic.rs:342-359, 440-605) - row a7 of SURVEY 8 (grad u) in use
# ------------------------------------------------------------------------------------------------------------------------
def volume_u_grad(j_inv_t: np.ndarray, G_ref: np.ndarray, U: np.ndarray) -> np.ndarray:
    """compute_volume_u_grad (elliptic.rs:25-59): grad u = J^{-T} sum_I grad_ref phi_I (x) u_I, shape (d, s).
    G_ref: (d, n) reference gradients as columns, U: (s, n) nodal values as columns."""
    d, s = G_ref.shape[0], U.shape[0]
    acc = np.zeros((d, s))
    for I in range(G_ref.shape[1]):
        acc += np.outer(G_ref[:, I], U[:, I])  # ger, node by node
    return j_inv_t @ acc


def elliptic_operator_transpose(op: int, u_grad: np.ndarray, params) -> np.ndarray:
    """g^T (s x d).  Laplace: g = grad u (laplace.rs:44-51).  MaterialEllipticOperator<LinearElasticMaterial>: g^T = P(F) with
    F = I + (grad u)^T (compute_stress_tensor_du, fenris-solid lib.rs:20-29, 98-104, 460-468; materials.rs:86-95)."""
    if op == LAPLACE:
        return u_grad.T.copy()
    mu, lam = params
    d = u_grad.shape[0]
    return linear_elastic_stress(np.eye(d) + u_grad.T, mu, lam)


def elliptic_energy_density(op: int, u_grad: np.ndarray, params) -> float:
    """psi(grad u): Laplace 0.5 grad u . grad u (laplace.rs:33-35); LinearElastic through F = I + (grad u)^T (lib.rs:78-84, 438-440)."""
    if op == LAPLACE:
        return 0.5 * float(np.sum(u_grad * u_grad))
    mu, lam = params
    d = u_grad.shape[0]
    return linear_elastic_energy_density(np.eye(d) + u_grad.T, mu, lam)


def element_elliptic_vector(elem_type: int, X_elem: np.ndarray, op: int, u_element: np.ndarray, weights, points, params_per_point) -> np.ndarray:
    """assemble_element_elliptic_vector (elliptic.rs:456-526): output (s x n) += w |det J| (g^T J^{-T}) G_ref; returned node-major."""
    n, ng, d = element_info(elem_type)
    s = solution_dim(op, d)
    X = np.asarray(X_elem, dtype=np.float64)[:ng].T
    U = np.asarray(u_element, dtype=np.float64).reshape(n, s).T  # s x n
    out = np.zeros((s, n))
    for w, xi, par in zip(weights, points, params_per_point):
        J = reference_jacobian(elem_type, X, xi)
        det = det_small(J)
        inv = try_inverse_small(J)
        if inv is None:
            raise SingularJacobian()
        j_inv_t = inv.T
        G_ref = reference_gradients(elem_type, xi)
        u_grad = volume_u_grad(j_inv_t, G_ref, U)
        g_t = elliptic_operator_transpose(op, u_grad, par)
        out += (w * abs(det)) * ((g_t @ j_inv_t) @ G_ref)
    return out.T.reshape(-1)


def element_elliptic_energy(elem_type: int, X_elem: np.ndarray, op: int, u_element: np.ndarray, weights, points, params_per_point) -> float:
    """compute_element_elliptic_energy (elliptic.rs:545-605)."""
    n, ng, d = element_info(elem_type)
    s = solution_dim(op, d)
    X = np.asarray(X_elem, dtype=np.float64)[:ng].T
    U = np.asarray(u_element, dtype=np.float64).reshape(n, s).T
    integral = 0.0
    for w, xi, par in zip(weights, points, params_per_point):
        J = reference_jacobian(elem_type, X, xi)
        det = det_small(J)
        inv = try_inverse_small(J)
        if inv is None:
            raise SingularJacobian()
        u_grad = volume_u_grad(inv.T, reference_gradients(elem_type, xi), U)
        integral += w * abs(det) * elliptic_energy_density(op, u_grad, par)
    return integral


def assemble_elliptic_vector_serial(problem: "Problem", u: np.ndarray) -> np.ndarray:
    """VectorAssembler::assemble_vector (global.rs:569-617) over an ElementEllipticAssembler."""
    s = problem.sdim
    U = np.asarray(u, dtype=np.float64).reshape(-1, s)
    out = np.zeros(s * len(problem.vertices))
    for e in range(len(problem.connectivity)):
        nodes = problem.connectivity[e]
        local = element_elliptic_vector(problem.elem_type, problem.vertices[nodes], problem.op, U[nodes].reshape(-1), problem.weights,
                                        problem.points, problem.params_per_point)
        for a, I in enumerate(nodes):
            out[s * I:s * I + s] += local[s * a:s * a + s]
    return out


def assemble_elliptic_scalar(problem: "Problem", u: np.ndarray) -> float:
    """assemble_scalar (global.rs:697-722): the element energies summed in element order."""
    s = problem.sdim
    U = np.asarray(u, dtype=np.float64).reshape(-1, s)
    total = 0.0
    for e in range(len(problem.connectivity)):
        nodes = problem.connectivity[e]
        total += element_elliptic_energy(problem.elem_type, problem.vertices[nodes], problem.op, U[nodes].reshape(-1), problem.weights,
                                         problem.points, problem.params_per_point)
    return total


def assemble_elliptic_vector_with_table(elem_type: int, vertices, connectivity, op: int, rules, element_to_rule, u) -> np.ndarray:
    """VectorAssembler::assemble_vector over an ElementEllipticAssembler whose table has a rule per element (elliptic.rs:456-526 with
    populate_element_quadrature_from_table).  rules[r] = (weights, points, params_per_point)."""
    conn = np.asarray(connectivity, dtype=np.int64)
    V = np.asarray(vertices, dtype=np.float64)
    _, _, d = element_info(elem_type)
    s = solution_dim(op, d)
    U = np.asarray(u, dtype=np.float64).reshape(-1, s)
    out = np.zeros(s * len(V))
    for e in range(len(conn)):
        w, p, params = rules[int(element_to_rule[e])]
        local = element_elliptic_vector(elem_type, V[conn[e]], op, U[conn[e]].reshape(-1), w, p, params)
        for a, I in enumerate(conn[e]):
            out[s * I:s * I + s] += local[s * a:s * a + s]
    return out


def assemble_elliptic_scalar_with_table(elem_type: int, vertices, connectivity, op: int, rules, element_to_rule, u) -> float:
    """assemble_scalar (global.rs:697-722) with a rule per element."""
    conn = np.asarray(connectivity, dtype=np.int64)
    V = np.asarray(vertices, dtype=np.float64)
    _, _, d = element_info(elem_type)
    s = solution_dim(op, d)
    U = np.asarray(u, dtype=np.float64).reshape(-1, s)
    total = 0.0
    for e in range(len(conn)):
        w, p, params = rules[int(element_to_rule[e])]
        total += element_elliptic_energy(elem_type, V[conn[e]], op, U[conn[e]].reshape(-1), w, p, params)
    return total


# ------------------------------------------------------------------------------------------------------------------------
# a non-linear material - St. Venant-Kirchhoff (fenris-solid/src/materials.rs:355-469) with u != 0
# ------------------------------------------------------------------------------------------------------------------------
STVK = 3


def _green_strain(F: np.ndarray) -> np.ndarray:
    """materials.rs:379-388: E = (F^T F - I) / 2."""
    return (F.T @ F - np.eye(F.shape[0])) * 0.5


def stvk_energy_density(F: np.ndarray, mu: float, lam: float) -> float:
    """materials.rs:400-404."""
    E = _green_strain(F)
    return mu * float(np.sum(E * E)) + 0.5 * lam * float(np.trace(E)) ** 2


def stvk_stress(F: np.ndarray, mu: float, lam: float) -> np.ndarray:
    """materials.rs:406-415: P = 2 mu F E + lambda tr(E) F."""
    E = _green_strain(F)
    return F @ E * 2.0 * mu + F * lam * np.trace(E)


def stvk_contraction(F: np.ndarray, a: np.ndarray, b: np.ndarray, mu: float, lam: float) -> np.ndarray:
    """materials.rs:417-437: C = I (2 mu a.E b + lambda tr(E) a.b) + mu Fb Fa^T + lambda Fa Fb^T + mu (a.b) F F^T."""
    d = F.shape[0]
    E = _green_strain(F)
    a_dot_b = float(a @ b)
    Fa, Fb, Eb = F @ a, F @ b, E @ b
    return (np.eye(d) * (2.0 * mu * float(a @ Eb) + lam * np.trace(E) * a_dot_b) + np.outer(Fb, Fa) * mu + np.outer(Fa, Fb) * lam
            + F @ F.T * mu * a_dot_b)


NEO_HOOKEAN = 4


def log_det_F(du_dX: np.ndarray):
    """fenris-solid/src/logdet.rs:17-86: log det(I + du_dX) as log1p(gamma) (accurate for small displacement gradients); None when
    det F <= 0."""
    U = du_dX
    if U.shape[0] == 2:
        gamma = U[0, 0] * U[1, 1] + U[0, 0] + U[1, 1] - U[0, 1] * U[1, 0]
    else:
        u11, u22, u33 = U[0, 0], U[1, 1], U[2, 2]
        a, e, i = 1.0 + u11, 1.0 + u22, 1.0 + u33
        b, c, d, f, g, h = U[0, 1], U[0, 2], U[1, 0], U[1, 2], U[2, 0], U[2, 1]
        gamma = (u11 * u22 * u33 + u11 * u22 + u11 * u33 + u22 * u33 + u11 + u22 + u33 + b * f * g + c * d * h - c * e * g - b * d * i
                 - a * f * h)
    return math.log1p(gamma) if gamma > -1.0 else None


def neo_hookean_energy_density_du(u_grad: np.ndarray, mu: float, lam: float) -> float:
    """materials.rs:251-265: psi = mu tr(E) - mu log J + lambda (log J)^2 / 2 with tr(E) = tr(U) + |U|^2 / 2, U = (grad u)^T."""
    U = u_grad.T
    logJ = log_det_F(U)
    if logJ is None:
        return math.inf
    tr_E = float(np.trace(U)) + 0.5 * float(np.sum(U * U))
    return mu * tr_E - mu * logJ + 0.5 * lam * logJ ** 2


def neo_hookean_stress(F: np.ndarray, mu: float, lam: float) -> np.ndarray:
    """materials.rs:267-289: P = F^-T (-mu + lambda log J) + mu F; NaN for J <= 0."""
    J = det_small(F)
    if J <= 0.0:
        return np.full(F.shape, math.nan)
    F_inv_T = try_inverse_small(F).T
    return F_inv_T * (-mu + lam * math.log(J)) + F * mu


def neo_hookean_contraction(F: np.ndarray, a: np.ndarray, b: np.ndarray, mu: float, lam: float) -> np.ndarray:
    """materials.rs:291-318: C = lambda (F^-T a)(F^-T b)^T - alpha (F^-T b)(F^-T a)^T + mu (a.b) I, alpha = -mu + lambda log J."""
    J = det_small(F)
    if J <= 0.0:
        return np.full(F.shape, math.nan)
    F_inv_T = try_inverse_small(F).T
    Ta, Tb = F_inv_T @ a, F_inv_T @ b
    alpha = -mu + lam * math.log(J)
    return np.outer(Ta, Tb) * lam - np.outer(Tb, Ta) * alpha + np.eye(F.shape[0]) * (mu * float(a @ b))


_elliptic_operator_transpose_linear = elliptic_operator_transpose
_elliptic_energy_density_linear = elliptic_energy_density


def elliptic_operator_transpose(op: int, u_grad: np.ndarray, params) -> np.ndarray:  # noqa: F811 (extends the linear version with STVK)
    if op == STVK:
        return stvk_stress(np.eye(u_grad.shape[0]) + u_grad.T, params[0], params[1])
    if op == NEO_HOOKEAN:
        return neo_hookean_stress(np.eye(u_grad.shape[0]) + u_grad.T, params[0], params[1])
    return _elliptic_operator_transpose_linear(op, u_grad, params)


def elliptic_energy_density(op: int, u_grad: np.ndarray, params) -> float:  # noqa: F811
    if op == STVK:
        return stvk_energy_density(np.eye(u_grad.shape[0]) + u_grad.T, params[0], params[1])
    if op == NEO_HOOKEAN:  # MaterialEllipticOperator -> compute_energy_density_du (fenris-solid lib.rs:492-499)
        return neo_hookean_energy_density_du(u_grad, params[0], params[1])
    return _elliptic_energy_density_linear(op, u_grad, params)


def contract_u(op: int, u_grad: np.ndarray, a: np.ndarray, b: np.ndarray, params) -> np.ndarray:
    """EllipticContraction::contract with the state: MaterialEllipticOperator::contract = compute_stress_contraction_du
    (fenris-solid lib.rs:480-489, 140-151: F = I + (grad u)^T); the linear operators ignore u."""
    if op == STVK:
        return stvk_contraction(np.eye(u_grad.shape[0]) + u_grad.T, a, b, params[0], params[1])
    if op == NEO_HOOKEAN:
        return neo_hookean_contraction(np.eye(u_grad.shape[0]) + u_grad.T, a, b, params[0], params[1])
    return contract(op, a, b, params)


def element_matrix_u(elem_type: int, X_elem: np.ndarray, op: int, u_element: np.ndarray, weights, points, params_per_point) -> np.ndarray:
    """assemble_element_elliptic_matrix (elliptic.rs:361-439) INCLUDING the state: grad u per point (compute_volume_u_grad), upper
    blocks I <= J of K += scale * contract(grad u, grad phi_I, grad phi_J) (operators.rs:176-188), then clone_upper_to_lower."""
    n, ng, d = element_info(elem_type)
    s = d if op in (LINEAR_ELASTIC, STVK, NEO_HOOKEAN) else 1
    X = np.asarray(X_elem, dtype=np.float64)[:ng].T
    U = np.asarray(u_element, dtype=np.float64).reshape(n, s).T
    K = np.zeros((s * n, s * n))
    for w, xi, par in zip(weights, points, params_per_point):
        J = reference_jacobian(elem_type, X, xi)
        det = det_small(J)
        inv = try_inverse_small(J)
        if inv is None:
            raise SingularJacobian()
        j_inv_t = inv.T
        G_ref = reference_gradients(elem_type, xi)
        u_grad = volume_u_grad(j_inv_t, G_ref, U)
        G = j_inv_t @ G_ref
        scale = w * abs(det)
        for I in range(n):
            for Jn in range(I, n):
                K[s * I:s * I + s, s * Jn:s * Jn + s] += scale * contract_u(op, u_grad, G[:, I], G[:, Jn], par)
    iu = np.triu_indices(s * n, 1)
    K[(iu[1], iu[0])] = K[iu]
    return K


def assemble_matrix_u_serial(elem_type: int, vertices, connectivity, op: int, u: np.ndarray, weights, points, params_per_point):
    """CsrAssembler::assemble (global.rs:124-182) with a state-dependent operator."""
    conn = np.asarray(connectivity, dtype=np.int64)
    V = np.asarray(vertices, dtype=np.float64)
    _, _, d = element_info(elem_type)
    s = d if op in (LINEAR_ELASTIC, STVK, NEO_HOOKEAN) else 1
    Uall = np.asarray(u, dtype=np.float64).reshape(-1, s)
    ro, ci = assemble_pattern(s, len(V), conn.tolist())
    values = np.zeros(len(ci))
    for e in range(len(conn)):
        K = element_matrix_u(elem_type, V[conn[e]], op, Uall[conn[e]].reshape(-1), weights, points, params_per_point)
        scatter_element(values, ro, ci, s, conn[e].tolist(), K)
    return ro, ci, values
