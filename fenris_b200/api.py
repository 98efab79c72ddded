"""Host-side mirror of the reference's operator/assembler interface for the assembly hot path.

Same names, argument meaning and error behaviour as the Rust reference so that the parity tests read like the
reference's own tests (file:line = InteractiveComputerGraphics/fenris @ 7181b15):

    Mesh / procedural generators            src/mesh.rs:23-40, src/mesh/procedural.rs
    UniformQuadratureTable                  src/assembly/local/quadrature_table.rs:213-298
    LaplaceOperator                         src/assembly/operators/laplace.rs:14
    LameParameters / YoungPoisson           fenris-solid/src/materials.rs:9-43
    MaterialEllipticOperator(LinearElasticMaterial)   fenris-solid/src/lib.rs:412-508
    ElementEllipticAssemblerBuilder         src/assembly/local/elliptic.rs:63-150
    ElementConnectivityAssembler            src/assembly/local.rs:18-47
    color_nodes                             src/assembly/global.rs:540-551
    CsrAssembler / CsrParAssembler          src/assembly/global.rs:27-182 / 186-376

Everything numerical happens in libfenris_b200.so on the GPU; this module only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _native as nat
from ._native import Fb200Error, SingularJacobianError  # noqa: F401
from .context import Context

_NODES = {nat.QUAD4: 4, nat.TET4: 4, nat.HEX8: 8, nat.HEX27: 27, nat.TET10: 10, nat.HEX20: 20}
_DIM = {nat.QUAD4: 2, nat.TET4: 3, nat.HEX8: 3, nat.HEX27: 3, nat.TET10: 3, nat.HEX20: 3}


# ----------------------------------------------------------------------------- meshes
class Mesh:
    """Mesh<f64, D, C>: `vertices` (N x D, AoS) and `connectivity` (E x n usize)."""

    def __init__(self, vertices: np.ndarray, connectivity: np.ndarray, element_type: int):
        self.vertices_ = np.ascontiguousarray(vertices, dtype=np.float64)
        self.connectivity_ = np.ascontiguousarray(connectivity, dtype=np.uint64)
        self.element_type = element_type
        assert self.connectivity_.ndim == 2 and self.connectivity_.shape[1] == _NODES[element_type]
        assert self.vertices_.ndim == 2 and self.vertices_.shape[1] == _DIM[element_type]

    @staticmethod
    def from_vertices_and_connectivity(vertices, connectivity, element_type):
        return Mesh(vertices, connectivity, element_type)

    def vertices(self):
        return self.vertices_

    def connectivity(self):
        return self.connectivity_

    # ElementConnectivityAssembler for Mesh (local.rs:49-75): solution_dim = 1
    def solution_dim(self):
        return 1

    def num_elements(self):
        return self.connectivity_.shape[0]

    def num_nodes(self):
        return self.vertices_.shape[0]

    def element_node_count(self, i):
        return self.connectivity_.shape[1]

    def populate_element_nodes(self, output, i):
        output[:] = self.connectivity_[i]


def _gen(fn, dims, h, n_per_elem, dim, element_type):
    L = nat.lib()
    nv, ne = C.c_uint64(0), C.c_uint64(0)
    st = fn(*dims, C.c_double(h), C.byref(nv), C.byref(ne), None, None)
    if st != nat.OK:
        raise Fb200Error(st, "mesh generator failed")
    v = np.zeros((nv.value, dim))
    c = np.zeros((ne.value, n_per_elem), dtype=np.uint64)
    if nv.value and ne.value:
        st = fn(*dims, C.c_double(h), C.byref(nv), C.byref(ne), nat.ptr(v), nat.ptr(c))
        if st != nat.OK:
            raise Fb200Error(st, "mesh generator failed")
    del L
    return Mesh(v, c, element_type)


def create_rectangular_uniform_hex_mesh(unit_length: float, units_x: int, units_y: int, units_z: int, cells_per_unit: int) -> Mesh:
    """src/mesh/procedural.rs:216-277"""
    if cells_per_unit == 0:
        return Mesh(np.zeros((0, 3)), np.zeros((0, 8), dtype=np.uint64), nat.HEX8)
    h = unit_length / cells_per_unit
    return _gen(nat.lib().fb200_gen_hex_mesh, (units_x * cells_per_unit, units_y * cells_per_unit, units_z * cells_per_unit), h, 8, 3, nat.HEX8)


def create_unit_box_uniform_hex_mesh_3d(cells_per_dim: int) -> Mesh:
    """src/mesh/procedural.rs:30-35"""
    return create_rectangular_uniform_hex_mesh(1.0, 1, 1, 1, cells_per_dim)


def create_rectangular_uniform_tet_mesh(unit_length: float, units_x: int, units_y: int, units_z: int, cells_per_unit: int) -> Mesh:
    """src/mesh/procedural.rs:286-403"""
    if cells_per_unit == 0:
        return Mesh(np.zeros((0, 3)), np.zeros((0, 4), dtype=np.uint64), nat.TET4)
    h = unit_length / float(cells_per_unit)
    return _gen(nat.lib().fb200_gen_tet_mesh, (units_x * cells_per_unit, units_y * cells_per_unit, units_z * cells_per_unit), h, 4, 3, nat.TET4)


def create_unit_box_uniform_tet_mesh_3d(cells_per_dim: int) -> Mesh:
    """src/mesh/procedural.rs:37-42"""
    return create_rectangular_uniform_tet_mesh(1.0, 1, 1, 1, cells_per_dim)


def create_unit_square_uniform_quad_mesh_2d(cells_per_dim: int) -> Mesh:
    """src/mesh/procedural.rs:15-20"""
    if cells_per_dim == 0:
        return Mesh(np.zeros((0, 2)), np.zeros((0, 4), dtype=np.uint64), nat.QUAD4)
    return _gen(nat.lib().fb200_gen_quad_mesh, (cells_per_dim, cells_per_dim), 1.0 / cells_per_dim, 4, 2, nat.QUAD4)


def hex27_mesh_from(hex8_mesh: Mesh) -> Mesh:
    """Hex27Mesh::from(&hex8_mesh), src/mesh_convert.rs:85-166,227-330"""
    assert hex8_mesh.element_type == nat.HEX8
    L = nat.lib()
    v, c = hex8_mesh.vertices_, hex8_mesh.connectivity_
    n27 = C.c_uint64(0)
    st = L.fb200_hex27_from_hex8(len(v), nat.ptr(v), len(c), nat.ptr(c), C.byref(n27), None, None)
    if st != nat.OK:
        raise Fb200Error(st, "hex27 conversion failed")
    v27 = np.zeros((n27.value, 3))
    c27 = np.zeros((len(c), 27), dtype=np.uint64)
    st = L.fb200_hex27_from_hex8(len(v), nat.ptr(v), len(c), nat.ptr(c), C.byref(n27), nat.ptr(v27), nat.ptr(c27))
    if st != nat.OK:
        raise Fb200Error(st, "hex27 conversion failed")
    return Mesh(v27, c27, nat.HEX27)


def hex20_mesh_from(hex8_mesh: Mesh) -> Mesh:
    """Hex20Mesh::from(&hex8_mesh), src/mesh_convert.rs:168-217,481-488"""
    assert hex8_mesh.element_type == nat.HEX8
    L = nat.lib()
    v, c = hex8_mesh.vertices_, hex8_mesh.connectivity_
    n20 = C.c_uint64(0)
    st = L.fb200_hex20_from_hex8(len(v), nat.ptr(v), len(c), nat.ptr(c), C.byref(n20), None, None)
    if st != nat.OK:
        raise Fb200Error(st, "hex20 conversion failed")
    v20 = np.zeros((n20.value, 3))
    c20 = np.zeros((len(c), 20), dtype=np.uint64)
    st = L.fb200_hex20_from_hex8(len(v), nat.ptr(v), len(c), nat.ptr(c), C.byref(n20), nat.ptr(v20), nat.ptr(c20))
    if st != nat.OK:
        raise Fb200Error(st, "hex20 conversion failed")
    return Mesh(v20, c20, nat.HEX20)


def tet10_mesh_from(tet4_mesh: Mesh) -> Mesh:
    """Tet10Mesh::from(&tet4_mesh), src/mesh_convert.rs:42-83,444-452"""
    assert tet4_mesh.element_type == nat.TET4
    L = nat.lib()
    v, c = tet4_mesh.vertices_, tet4_mesh.connectivity_
    n10 = C.c_uint64(0)
    st = L.fb200_tet10_from_tet4(len(v), nat.ptr(v), len(c), nat.ptr(c), C.byref(n10), None, None)
    if st != nat.OK:
        raise Fb200Error(st, "tet10 conversion failed")
    v10 = np.zeros((n10.value, 3))
    c10 = np.zeros((len(c), 10), dtype=np.uint64)
    st = L.fb200_tet10_from_tet4(len(v), nat.ptr(v), len(c), nat.ptr(c), C.byref(n10), nat.ptr(v10), nat.ptr(c10))
    if st != nat.OK:
        raise Fb200Error(st, "tet10 conversion failed")
    return Mesh(v10, c10, nat.TET10)


# ----------------------------------------------------------------------------- quadrature tables / operators
def canonical_stiffness_quadrature(element_type: int):
    """(weights, points) of the element's CanonicalStiffnessQuadrature, src/quadrature/canonical.rs:95-112"""
    L = nat.lib()
    n = C.c_int32(0)
    st = L.fb200_canonical_quadrature(element_type, C.byref(n), None, None)
    if st != nat.OK:
        raise Fb200Error(st, "no canonical quadrature for this element type")
    w = np.zeros(n.value)
    p = np.zeros((n.value, _DIM[element_type]))
    L.fb200_canonical_quadrature(element_type, C.byref(n), nat.ptr(w), nat.ptr(p))
    return w, p


@dataclass
class LameParameters:
    mu: float = 0.0
    lambda_: float = 0.0

    @staticmethod
    def from_young_poisson(yp: "YoungPoisson") -> "LameParameters":
        mu, lam = C.c_double(0), C.c_double(0)
        nat.lib().fb200_lame_from_young_poisson(yp.young, yp.poisson, C.byref(mu), C.byref(lam))
        return LameParameters(mu.value, lam.value)


@dataclass
class YoungPoisson:
    young: float
    poisson: float


class UniformQuadratureTable:
    def __init__(self, points, weights, data=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64)
        assert len(self.points) == len(self.weights), "points and weights must have the same length"
        self.data = data  # None (= ()) or list of LameParameters per point

    @staticmethod
    def from_points_and_weights(points, weights):
        return UniformQuadratureTable(points, weights, None)

    @staticmethod
    def from_quadrature(rule):
        w, p = rule
        return UniformQuadratureTable(p, w, None)

    @staticmethod
    def from_points_weights_and_data(points, weights, data):
        assert len(data) == len(weights)
        return UniformQuadratureTable(points, weights, list(data))

    def with_uniform_data(self, data):
        return UniformQuadratureTable(self.points, self.weights, [data] * len(self.weights))


class CompactQuadratureTable:
    """src/assembly/local/quadrature_table.rs:312-439: a set of rules and a map element -> rule."""

    def __init__(self, points, weights, data, element_to_rule_map):
        assert len(points) == len(weights), "Quadrature point and weight tables must have the same number of rules."
        self.points = [np.ascontiguousarray(p, dtype=np.float64) for p in points]
        self.weights = [np.ascontiguousarray(w, dtype=np.float64) for w in weights]
        self.data = [None] * len(weights) if data is None else list(data)
        assert len(self.data) == len(self.weights), "Quadrature point and data tables must have the same number of rules."
        for r, (p, w, d) in enumerate(zip(self.points, self.weights, self.data)):
            assert len(p) == len(w) and (d is None or len(d) == len(w)), f"rule {r} has mismatched numbers of points, weights and data"
        self.element_to_rule_map = np.ascontiguousarray(element_to_rule_map, dtype=np.uint32)
        assert self.element_to_rule_map.size == 0 or int(self.element_to_rule_map.max()) < len(self.weights), \
            "element to rule map contains out-of-bounds rule indices"  # quadrature_table.rs:361-366

    @staticmethod
    def from_points_weights_and_map(points, weights, element_to_rule_map):
        return CompactQuadratureTable(points, weights, None, element_to_rule_map)

    @staticmethod
    def from_quadrature_rules_and_map(points, weights, data, element_to_rule_map):
        return CompactQuadratureTable(points, weights, data, element_to_rule_map)


class GeneralQuadratureTable(CompactQuadratureTable):
    """src/assembly/local/quadrature_table.rs:57-210: one rule per element (identical rules are merged before they reach the device)."""

    def __init__(self, points, weights, data=None):
        data = [None] * len(weights) if data is None else list(data)
        keys, rules, emap = {}, [], []
        for p, w, d in zip(points, weights, data):
            p = np.ascontiguousarray(p, dtype=np.float64)
            w = np.ascontiguousarray(w, dtype=np.float64)
            dk = None if d is None else tuple((x.mu, x.lambda_) if hasattr(x, "mu") else float(x) for x in d)
            key = (p.tobytes(), w.tobytes(), dk)
            if key not in keys:
                keys[key] = len(rules)
                rules.append((p, w, d))
            emap.append(keys[key])
        super().__init__([r[0] for r in rules], [r[1] for r in rules], [r[2] for r in rules], emap)

    @staticmethod
    def from_points_and_weights(points, weights):
        return GeneralQuadratureTable(points, weights, None)

    @staticmethod
    def from_points_weights_and_data(points, weights, data):
        return GeneralQuadratureTable(points, weights, data)


class LaplaceOperator:
    kind = nat.LAPLACE

    def solution_dim(self, geometry_dim):
        return 1


class LinearElasticMaterial:
    pass


class StVKMaterial:
    """fenris-solid/src/materials.rs:355-469 (state-dependent: the assembler's u enters through F = I + (grad u)^T)."""


class NeoHookeanMaterial:
    """fenris-solid/src/materials.rs:232-353 (state-dependent; det F <= 0 gives NaN / +inf as in the reference)."""


class MaterialEllipticOperator:
    """Wraps a hyperelastic material as an elliptic operator (fenris-solid/src/lib.rs:412-508).
    LinearElasticMaterial, StVKMaterial and NeoHookeanMaterial have device specialisations; anything else raises (no CPU fallback)."""

    def __init__(self, material):
        if isinstance(material, LinearElasticMaterial) or material is LinearElasticMaterial:
            self.kind = nat.LINEAR_ELASTIC
        elif isinstance(material, StVKMaterial) or material is StVKMaterial:
            self.kind = nat.STVK
        elif isinstance(material, NeoHookeanMaterial) or material is NeoHookeanMaterial:
            self.kind = nat.NEO_HOOKEAN
        else:
            raise Fb200Error(nat.ERR_UNSUPPORTED, "material has no device specialisation (no CPU fallback)")

    def solution_dim(self, geometry_dim):
        return geometry_dim


class ElementConnectivityAssembler:
    """A bare connectivity view (like the reference tests' MockElementAssembler)."""

    def __init__(self, solution_dim: int, num_nodes: int, element_connectivities: Sequence[Sequence[int]]):
        self._sdim, self._n, self._conn = solution_dim, num_nodes, [list(e) for e in element_connectivities]

    def solution_dim(self):
        return self._sdim

    def num_elements(self):
        return len(self._conn)

    def num_nodes(self):
        return self._n

    def element_node_count(self, i):
        return len(self._conn[i])

    def populate_element_nodes(self, output, i):
        output[:] = self._conn[i]


class ElementEllipticAssembler:
    def __init__(self, space: Mesh, op, qtable: UniformQuadratureTable, u):
        self.space, self.op, self.qtable, self.u = space, op, qtable, u

    # ElementConnectivityAssembler (elliptic.rs:160-187)
    def solution_dim(self):
        return self.op.solution_dim(_DIM[self.space.element_type])

    def num_elements(self):
        return self.space.num_elements()

    def num_nodes(self):
        return self.space.num_nodes()

    def element_node_count(self, i):
        return self.space.element_node_count(i)

    def populate_element_nodes(self, output, i):
        self.space.populate_element_nodes(output, i)

    def _data(self):
        if self.op.kind == nat.LAPLACE:
            return None
        if self.qtable.data is None:
            raise Fb200Error(nat.ERR_SHAPE, "quadrature table carries no LameParameters")
        return np.array([[d.mu, d.lambda_] for d in self.qtable.data], dtype=np.float64)

    def _rules(self):
        """(weights, points, data) per rule of a Compact / General table."""
        out = []
        for p, w, d in zip(self.qtable.points, self.qtable.weights, self.qtable.data):
            if self.op.kind == nat.LAPLACE:
                out.append((w, p, None))
            else:
                if d is None:
                    raise Fb200Error(nat.ERR_SHAPE, "quadrature table carries no LameParameters")
                out.append((w, p, np.array([[x.mu, x.lambda_] for x in d], dtype=np.float64)))
        return out

    # ElementMatrixAssembler::assemble_element_matrix (local.rs:77-103)
    def assemble_element_matrix(self, element_index: int, ctx: Optional[Context] = None) -> np.ndarray:
        own = ctx is None
        ctx = ctx or Context()
        try:
            ctx.space_upload(self.space.element_type, self.space.vertices_, self.space.connectivity_)
            dofs = self.solution_dim() * _NODES[self.space.element_type]
            return ctx.element_matrices(self.op.kind, self.qtable.weights, self.qtable.points, self._data(), element_index, 1, dofs,
                                        u=getattr(self, "u", None))[0]
        finally:
            if own:
                ctx.close()


class ElementEllipticAssemblerBuilder:
    """src/assembly/local/elliptic.rs:63-150"""

    def __init__(self):
        self._space = self._op = self._qt = self._u = None

    def with_finite_element_space(self, space):
        self._space = space
        return self

    def with_operator(self, op):
        self._op = op
        return self

    def with_quadrature_table(self, qt):
        self._qt = qt
        return self

    def with_u(self, u):
        self._u = u
        return self

    def build(self) -> ElementEllipticAssembler:
        assert self._space is not None and self._op is not None and self._qt is not None
        if self._u is not None:
            expected = self._op.solution_dim(_DIM[self._space.element_type]) * self._space.num_nodes()
            assert len(self._u) == expected, "u has the wrong length"  # elliptic.rs:378-383
        return ElementEllipticAssembler(self._space, self._op, self._qt, self._u)


class Density(float):
    """Parameters of the mass matrix: a density per quadrature point (src/assembly/local/mass.rs:23-31)."""


class ElementMassAssembler:
    """src/assembly/local/mass.rs:33-159: M_IJ = I_s int rho phi_I phi_J.  qtable.data = one Density per point."""

    def __init__(self, space: Mesh, qtable: UniformQuadratureTable, solution_dim: int):
        self.space, self.qtable, self._s = space, qtable, int(solution_dim)

    @staticmethod
    def with_finite_element_space(space):  # builder-style spelling of the reference (mass.rs:52-107)
        return _MassBuilder(space)

    def solution_dim(self):
        return self._s

    def num_elements(self):
        return self.space.num_elements()

    def num_nodes(self):
        return self.space.num_nodes()

    def element_node_count(self, i):
        return self.space.element_node_count(i)

    def populate_element_nodes(self, output, i):
        self.space.populate_element_nodes(output, i)

    def _density(self):
        if self.qtable.data is None:
            raise Fb200Error(nat.ERR_SHAPE, "quadrature table carries no Density data")
        return np.array([float(d) for d in self.qtable.data], dtype=np.float64)


class _MassBuilder:
    def __init__(self, space):
        self._space, self._qt, self._s = space, None, None

    def with_quadrature_table(self, qt):
        self._qt = qt
        return self

    def with_solution_dim(self, s):
        self._s = s
        return self

    def build(self) -> ElementMassAssembler:
        assert self._qt is not None and self._s is not None
        return ElementMassAssembler(self._space, self._qt, self._s)


class ElementSourceAssembler:
    """src/assembly/local/source.rs:24-190: f_I = int f(x) phi_I.  `source` is a callable (x: (..., d) array, data) -> (..., s) array
    (SourceFunction::evaluate, vectorised); it runs on the host at the physical quadrature points the device computes."""

    def __init__(self, space: Mesh, qtable: UniformQuadratureTable, source, solution_dim: int):
        self.space, self.qtable, self.source, self._s = space, qtable, source, int(solution_dim)

    def solution_dim(self):
        return self._s

    def num_elements(self):
        return self.space.num_elements()

    def num_nodes(self):
        return self.space.num_nodes()

    def element_node_count(self, i):
        return self.space.element_node_count(i)

    def populate_element_nodes(self, output, i):
        self.space.populate_element_nodes(output, i)


class VectorAssembler:
    """src/assembly/global.rs:569-617 (serial semantics); on the device element-parallel with f64 atomics."""

    def __init__(self, device: int = 0, scatter_mode: int = nat.SCATTER_ATOMIC):
        self.ctx = Context(device)
        self.scatter_mode = scatter_mode

    def _colors(self):
        return None

    def assemble_vector_into(self, output: np.ndarray, ea: ElementSourceAssembler):
        ctx = self.ctx
        ctx.space_upload(ea.space.element_type, ea.space.vertices_, ea.space.connectivity_)
        s, n = ea.solution_dim(), ea.num_nodes()
        assert len(output) == s * n, "Output dimensions mismatch"  # global.rs:592
        mode = self.scatter_mode
        colors = self._colors()
        if colors is not None:
            offs = np.zeros(len(colors) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(c.labels()) for c in colors])
            elems = np.concatenate([c.labels() for c in colors]) if len(colors) else np.zeros(0, dtype=np.uint64)
            ctx.colors_adopt(offs, elems)
            mode = nat.SCATTER_COLORED
        qt = ea.qtable
        if isinstance(ea, ElementEllipticAssembler):  # elliptic.rs:342-359: the operator's vector at ea.u
            assert ea.u is not None and len(ea.u) == s * n, "u has the wrong length"
            if isinstance(qt, CompactQuadratureTable):  # a rule per element (quadrature_table.rs:57-210, 312-439)
                ctx.assemble_elliptic_vector_table(ea.op.kind, ea._rules(), qt.element_to_rule_map, np.asarray(ea.u, dtype=np.float64), out=output,
                                                   scatter_mode=mode, accumulate=True)
                return output
            ctx.assemble_elliptic_vector(ea.op.kind, qt.weights, qt.points, ea._data(), np.asarray(ea.u, dtype=np.float64), out=output,
                                         scatter_mode=mode, accumulate=True)
            return output
        x = ctx.physical_quadrature_points(qt.weights, qt.points, ea.num_elements())
        data = qt.data if qt.data is not None else [None] * len(qt.weights)
        f = np.stack([np.asarray(ea.source(x[:, q], data[q]), dtype=np.float64).reshape(ea.num_elements(), s) for q in range(len(qt.weights))], axis=1)
        ctx.assemble_vector(qt.weights, qt.points, np.ascontiguousarray(f), n, out=output, scatter_mode=mode, accumulate=True)
        return output

    def assemble_vector(self, ea) -> np.ndarray:
        return self.assemble_vector_into(np.zeros(ea.solution_dim() * ea.num_nodes()), ea)


class VectorParAssembler(VectorAssembler):
    """src/assembly/global.rs:619-686: coloured; one launch per colour, plain read-modify-write."""

    def __init__(self, device: int = 0):
        super().__init__(device, nat.SCATTER_COLORED)
        self._cols = None

    def _colors(self):
        return self._cols

    def assemble_vector(self, colors, ea):  # type: ignore[override]
        self._cols = list(colors)
        return super().assemble_vector(ea)

    def assemble_vector_into(self, output, colors=None, ea=None):  # type: ignore[override]
        if ea is None:  # called through the base class with (output, ea)
            return super().assemble_vector_into(output, colors)
        self._cols = list(colors)
        return super().assemble_vector_into(output, ea)


# ----------------------------------------------------------------------------- CSR containers / colours
class SparsityPattern:
    def __init__(self, major_offsets, minor_indices, nrows):
        self.major_offsets, self.minor_indices, self.nrows = major_offsets, minor_indices, nrows

    def nnz(self):
        return len(self.minor_indices)


class CsrMatrix:
    """nalgebra_sparse::CsrMatrix<f64> as three arrays (usize, usize, f64)."""

    def __init__(self, row_offsets, col_indices, values):
        self.row_offsets, self.col_indices, self.values = row_offsets, col_indices, values

    @staticmethod
    def try_from_pattern_and_values(pattern: SparsityPattern, values):
        assert len(values) == pattern.nnz()
        return CsrMatrix(pattern.major_offsets, pattern.minor_indices, np.ascontiguousarray(values, dtype=np.float64))

    def nrows(self):
        return len(self.row_offsets) - 1

    def to_scipy(self):
        import scipy.sparse as sp
        n = self.nrows()
        return sp.csr_matrix((self.values, self.col_indices.astype(np.int64), self.row_offsets.astype(np.int64)), shape=(n, n))


class DisjointSubsets:
    """One colour: element labels whose node subsets are pairwise disjoint (fenris-paradis/src/lib.rs:171-181)."""

    def __init__(self, labels):
        self.labels_ = np.asarray(labels, dtype=np.uint64)

    def labels(self):
        return self.labels_


def _upload(ctx: Context, assembler):
    if isinstance(assembler, (ElementEllipticAssembler, ElementMassAssembler, ElementSourceAssembler)):
        ctx.space_upload(assembler.space.element_type, assembler.space.vertices_, assembler.space.connectivity_)
    elif isinstance(assembler, Mesh):
        ctx.space_upload(assembler.element_type, assembler.vertices_, assembler.connectivity_)
    else:  # any ElementConnectivityAssembler
        conn = []
        for i in range(assembler.num_elements()):
            buf = np.zeros(assembler.element_node_count(i), dtype=np.uint64)
            assembler.populate_element_nodes(buf, i)
            conn.append(buf.tolist())
        ctx.connectivity_upload(assembler.num_nodes(), conn)


def color_nodes(connectivity, ctx: Optional[Context] = None) -> List[DisjointSubsets]:
    """src/assembly/global.rs:540-551 (sequential_greedy_coloring)."""
    own = ctx is None
    ctx = ctx or Context()
    try:
        _upload(ctx, connectivity)
        ctx.color_nodes()
        offs, elems = ctx.colors_download(connectivity.num_elements())
        return [DisjointSubsets(elems[int(offs[c]):int(offs[c + 1])]) for c in range(len(offs) - 1)]
    finally:
        if own:
            ctx.close()


class CsrAssembler:
    """Serial reference semantics (global.rs:27-182); on the device the element loop is parallel with
    f64 atomics (default) or the row-owner gather (scatter_mode)."""

    def __init__(self, device: int = 0, scatter_mode: int = nat.SCATTER_ATOMIC):
        self.ctx = Context(device)
        self.scatter_mode = scatter_mode

    def assemble_pattern(self, element_assembler) -> SparsityPattern:
        _upload(self.ctx, element_assembler)
        nrows, _ = self.ctx.assemble_pattern(element_assembler.solution_dim())
        ro, ci = self.ctx.pattern_download()
        return SparsityPattern(ro, ci, nrows)

    def assemble(self, element_assembler: ElementEllipticAssembler) -> CsrMatrix:
        pattern = self.assemble_pattern(element_assembler)
        matrix = CsrMatrix.try_from_pattern_and_values(pattern, np.zeros(pattern.nnz()))
        self._assemble_values(matrix, element_assembler, adopt=False)
        return matrix

    def assemble_into_csr(self, csr: CsrMatrix, element_assembler: ElementEllipticAssembler):
        _upload(self.ctx, element_assembler)
        self.ctx.pattern_adopt(element_assembler.solution_dim(), csr.row_offsets, csr.col_indices)
        self._assemble_values(csr, element_assembler, adopt=True)

    def _colors(self):
        return None

    def _assemble_values(self, csr: CsrMatrix, ea: ElementEllipticAssembler, adopt: bool):
        mode = self.scatter_mode
        colors = self._colors()
        if colors is not None:
            offs = np.zeros(len(colors) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(c.labels()) for c in colors])
            elems = np.concatenate([c.labels() for c in colors]) if len(colors) else np.zeros(0, dtype=np.uint64)
            self.ctx.colors_adopt(offs, elems)
            mode = nat.SCATTER_COLORED
        if len(csr.values) == 0:
            return
        if isinstance(getattr(ea, "qtable", None), CompactQuadratureTable):  # a rule per element (quadrature_table.rs:57-210, 312-439)
            if mode == nat.SCATTER_GATHER:
                mode = nat.SCATTER_ATOMIC
            self.ctx.values_upload(csr.values)
            u = np.asarray(ea.u, dtype=np.float64) if ea.op.kind in (nat.STVK, nat.NEO_HOOKEAN) and ea.u is not None else None
            self.ctx.assemble_into_csr_table_device(ea.op.kind, ea._rules(), ea.qtable.element_to_rule_map, scatter_mode=mode, accumulate=True, u=u)
            self.ctx.synchronize()
            self.ctx.values_download(csr.values)
            return
        if isinstance(ea, ElementMassAssembler):
            if mode == nat.SCATTER_GATHER:
                mode = nat.SCATTER_ATOMIC
            self.ctx.assemble_mass_into_csr(ea.qtable.weights, ea.qtable.points, ea._density(), csr.values, scatter_mode=mode, accumulate=True)
            return
        u = None
        if ea.op.kind in (nat.STVK, nat.NEO_HOOKEAN):  # the state enters the contraction (elliptic.rs:393-399)
            if mode == nat.SCATTER_GATHER:
                mode = nat.SCATTER_ATOMIC
            u = None if ea.u is None else np.asarray(ea.u, dtype=np.float64)
        self.ctx.assemble_into_csr(ea.op.kind, ea.qtable.weights, ea.qtable.points, ea._data(), csr.values, scatter_mode=mode, accumulate=True, u=u)


class CsrParAssembler(CsrAssembler):
    """Coloured assembly (global.rs:186-376): one kernel launch per colour, plain read-modify-write."""

    def __init__(self, device: int = 0):
        super().__init__(device, nat.SCATTER_COLORED)
        self._cols = None

    def _colors(self):
        return self._cols

    def assemble(self, colors: Sequence[DisjointSubsets], element_assembler) -> CsrMatrix:  # type: ignore[override]
        self._cols = list(colors)
        return super().assemble(element_assembler)

    def assemble_into_csr(self, csr: CsrMatrix, colors: Sequence[DisjointSubsets], element_assembler):  # type: ignore[override]
        self._cols = list(colors)
        super().assemble_into_csr(csr, element_assembler)


def assemble_scalar(ea: ElementEllipticAssembler, ctx: Optional[Context] = None) -> float:
    """src/assembly/global.rs:697-722 over an ElementEllipticAssembler (ElementScalarAssembler, elliptic.rs:352-359): the elliptic energy of ea.u."""
    own = ctx is None
    ctx = ctx or Context()
    try:
        ctx.space_upload(ea.space.element_type, ea.space.vertices_, ea.space.connectivity_)
        if isinstance(ea.qtable, CompactQuadratureTable):
            return ctx.assemble_elliptic_scalar_table(ea.op.kind, ea._rules(), ea.qtable.element_to_rule_map, np.asarray(ea.u, dtype=np.float64))
        return ctx.assemble_elliptic_scalar(ea.op.kind, ea.qtable.weights, ea.qtable.points, ea._data(), np.asarray(ea.u, dtype=np.float64))
    finally:
        if own:
            ctx.close()


# ----------------------------------------------------------------------------- Dirichlet conditions (global.rs:379-495)
def apply_homogeneous_dirichlet_bc_csr(matrix: CsrMatrix, nodes, solution_dim: int, assembler: Optional[CsrAssembler] = None) -> float:
    """src/assembly/global.rs:379-451.  With `assembler` (the CsrAssembler that produced `matrix`) the values still resident on its
    device context are modified in place there and copied back; otherwise the matrix is uploaded to a temporary context first."""
    nodes = np.asarray(nodes, dtype=np.uint64)
    if assembler is not None:
        ctx, own = assembler.ctx, False
    else:
        ctx, own = Context(), True
        nrows = matrix.nrows()
        assert nrows % solution_dim == 0
        ctx.connectivity_upload(nrows // solution_dim, [])
        ctx.pattern_adopt(solution_dim, matrix.row_offsets, matrix.col_indices)
    try:
        ctx.values_upload(matrix.values)
        scale = ctx.apply_homogeneous_dirichlet_bc_csr(nodes)
        ctx.values_download(matrix.values)
        return scale
    finally:
        if own:
            ctx.close()


def apply_homogeneous_dirichlet_bc_rhs(rhs: np.ndarray, nodes, solution_dim: int) -> None:
    """src/assembly/global.rs:479-495 (host loop; the vector lives on the host)."""
    for node in np.asarray(nodes, dtype=np.int64):
        rhs[solution_dim * node:solution_dim * node + solution_dim] = 0.0

