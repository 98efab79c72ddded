"""Thin object wrapper over the C ABI (one fb200_ctx = one GPU + one stream)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from ._native import Fb200Error, SingularJacobianError


class Context:
    def __init__(self, device: int = 0):
        self._lib = nat.lib()
        h = C.c_void_p()
        st = self._lib.fb200_create(device, C.byref(h))
        if st != nat.OK:
            raise Fb200Error(st, "fb200_create failed: no usable CUDA device (fenris_b200 has no CPU fallback)")
        self._h = h
        self.device = device
        self._keep = []  # host arrays referenced by the last quadrature struct

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.fb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st: int):
        if st == nat.OK:
            return
        buf = C.create_string_buffer(512)
        elem = C.c_int64(-1)
        self._lib.fb200_last_error(self._h, buf, 512, C.byref(elem))
        msg = buf.value.decode() or self._lib.fb200_status_string(st).decode()
        if st == nat.ERR_SINGULAR_JACOBIAN:
            raise SingularJacobianError(st, msg, elem.value)
        raise Fb200Error(st, msg, elem.value)

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        self._check(self._lib.fb200_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        self._check(self._lib.fb200_synchronize(self._h))

    def timer_begin(self):
        self._check(self._lib.fb200_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float(0)
        self._check(self._lib.fb200_timer_end(self._h, C.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.fb200_launch_count(self._h))

    def set_tuning(self, name: str, value: int):
        """Kernel-selection knobs, e.g. ("hex8_tile", 64 | 0)."""
        self._check(self._lib.fb200_set_tuning(self._h, name.encode(), int(value)))

    # -- space
    def space_upload(self, element_type: int, vertices: np.ndarray, connectivity: np.ndarray):
        v = nat.as_f64(vertices)
        c = nat.as_u64(connectivity)
        self._check(self._lib.fb200_space_upload(self._h, element_type, v.shape[0], nat.ptr(v), c.shape[0], nat.ptr(c)))

    def space_update_vertices(self, vertices: np.ndarray):
        v = nat.as_f64(vertices)
        self._check(self._lib.fb200_space_update_vertices(self._h, nat.ptr(v)))

    def connectivity_upload(self, num_nodes: int, elements: Sequence[Sequence[int]]):
        offs = np.zeros(len(elements) + 1, dtype=np.uint64)
        if len(elements):
            offs[1:] = np.cumsum([len(e) for e in elements])
        flat = np.array([x for e in elements for x in e], dtype=np.uint64)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint64)
        self._check(self._lib.fb200_connectivity_upload(self._h, num_nodes, len(elements), nat.ptr(offs), nat.ptr(flat)))

    def set_num_owned_elements(self, n: int):
        self._check(self._lib.fb200_set_num_owned_elements(self._h, n))

    # -- pattern
    def assemble_pattern(self, solution_dim: int) -> Tuple[int, int]:
        nrows, nnz = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.fb200_assemble_pattern(self._h, solution_dim, C.byref(nrows), C.byref(nnz)))
        self.nrows, self.nnz = int(nrows.value), int(nnz.value)
        return self.nrows, self.nnz

    def pattern_download(self) -> Tuple[np.ndarray, np.ndarray]:
        ro = np.zeros(self.nrows + 1, dtype=np.uint64)
        ci = np.zeros(max(self.nnz, 1), dtype=np.uint64)
        self._check(self._lib.fb200_pattern_download(self._h, nat.ptr(ro), nat.ptr(ci)))
        return ro, ci[: self.nnz]

    def row_offsets_download(self) -> np.ndarray:
        """Only the row offsets of the CSR pattern (the column indices of C3 - C5 are 4 - 9 GB as u64)."""
        ro = np.zeros(self.nrows + 1, dtype=np.uint64)
        self._check(self._lib.fb200_pattern_download(self._h, nat.ptr(ro), None))
        return ro

    def pattern_adopt(self, solution_dim: int, row_offsets: np.ndarray, col_indices: np.ndarray):
        ro, ci = nat.as_u64(row_offsets), nat.as_u64(col_indices)
        if ci.size == 0:
            ci = np.zeros(1, dtype=np.uint64)
        self._check(self._lib.fb200_pattern_adopt(self._h, solution_dim, len(ro) - 1, nat.ptr(ro), nat.ptr(ci)))
        self.nrows, self.nnz = len(ro) - 1, int(ro[-1])

    # -- colours
    def color_nodes(self) -> int:
        n = C.c_uint64(0)
        self._check(self._lib.fb200_color_nodes(self._h, C.byref(n)))
        self.num_colors = int(n.value)
        return self.num_colors

    def colors_download(self, num_elements: int) -> Tuple[np.ndarray, np.ndarray]:
        offs = np.zeros(self.num_colors + 1, dtype=np.uint64)
        self._check(self._lib.fb200_colors_download(self._h, nat.ptr(offs), None))
        elems = np.zeros(max(int(offs[-1]), 1), dtype=np.uint64)
        self._check(self._lib.fb200_colors_download(self._h, None, nat.ptr(elems)))
        return offs, elems[: int(offs[-1])]

    def colors_adopt(self, color_offsets: np.ndarray, element_ids: np.ndarray):
        o, e = nat.as_u64(color_offsets), nat.as_u64(element_ids)
        if e.size == 0:
            e = np.zeros(1, dtype=np.uint64)
        self._check(self._lib.fb200_colors_adopt(self._h, len(o) - 1, nat.ptr(o), nat.ptr(e)))
        self.num_colors = len(o) - 1

    # -- assembly
    def _structs(self, op_kind: int, weights, points, data):
        w = nat.as_f64(weights)
        p = nat.as_f64(points).reshape(len(w), -1)
        d = None
        if data is not None:
            d = nat.as_f64(data)
            if d.ndim == 1:
                d = np.ascontiguousarray(np.tile(d, (len(w), 1)))
            assert d.shape == (len(w), 2)
        self._keep = [w, p, d]
        q = nat.Quadrature(len(w), p.shape[1], w.ctypes.data_as(C.POINTER(C.c_double)), p.ctypes.data_as(C.POINTER(C.c_double)),
                           d.ctypes.data_as(C.POINTER(C.c_double)) if d is not None else None)
        return nat.Operator(op_kind), q

    def assemble_into_csr_device(self, op_kind: int, weights, points, data=None, scatter_mode: int = nat.SCATTER_ATOMIC,
                                 accumulate: bool = False, u: Optional[np.ndarray] = None):
        """u: the state of a non-linear operator (STVK; elliptic.rs:361-439 with u_grad), ignored by the linear ones."""
        op, q = self._structs(op_kind, weights, points, data)
        uv = None if u is None else nat.as_f64(u).reshape(-1)
        self._check(self._lib.fb200_assemble_into_csr_device(self._h, C.byref(op), C.byref(q), None if uv is None else nat.ptr(uv), scatter_mode,
                                                             int(accumulate)))

    def assemble_into_csr(self, op_kind: int, weights, points, data, values: np.ndarray, scatter_mode: int = nat.SCATTER_ATOMIC,
                          accumulate: bool = True, u: Optional[np.ndarray] = None):
        assert values.dtype == np.float64 and values.flags["C_CONTIGUOUS"] and values.size >= self.nnz
        op, q = self._structs(op_kind, weights, points, data)
        uv = None if u is None else nat.as_f64(u).reshape(-1)
        self._check(self._lib.fb200_assemble_into_csr(self._h, C.byref(op), C.byref(q), None if uv is None else nat.ptr(uv), scatter_mode,
                                                      int(accumulate), nat.ptr(values)))
        return values

    # -- mass matrix / source vector / physical points
    def _quad_only(self, weights, points, density=None):
        w = nat.as_f64(weights)
        p = nat.as_f64(points).reshape(len(w), -1)
        d = None
        if density is not None:
            d = nat.as_f64(density).reshape(-1)
            if d.size == 1:
                d = np.full(len(w), float(d[0]))
            assert d.shape == (len(w),)
        self._keep = [w, p, d]
        return nat.Quadrature(len(w), p.shape[1], w.ctypes.data_as(C.POINTER(C.c_double)), p.ctypes.data_as(C.POINTER(C.c_double)),
                              d.ctypes.data_as(C.POINTER(C.c_double)) if d is not None else None)

    def assemble_mass_into_csr_device(self, weights, points, density, scatter_mode: int = nat.SCATTER_ATOMIC, accumulate: bool = False):
        q = self._quad_only(weights, points, density)
        self._check(self._lib.fb200_assemble_mass_into_csr_device(self._h, C.byref(q), scatter_mode, int(accumulate)))

    def assemble_mass_into_csr(self, weights, points, density, values: np.ndarray, scatter_mode: int = nat.SCATTER_ATOMIC,
                               accumulate: bool = True):
        assert values.dtype == np.float64 and values.flags["C_CONTIGUOUS"] and values.size >= self.nnz
        q = self._quad_only(weights, points, density)
        self._check(self._lib.fb200_assemble_mass_into_csr(self._h, C.byref(q), scatter_mode, int(accumulate), nat.ptr(values)))
        return values

    def assemble_vector(self, weights, points, source_values, num_nodes: int, out: Optional[np.ndarray] = None,
                        scatter_mode: int = nat.SCATTER_ATOMIC, accumulate: bool = False) -> np.ndarray:
        """source_values: (q, s) shared by all elements, or (E, q, s)."""
        f = nat.as_f64(source_values)
        assert f.ndim in (2, 3)
        s = f.shape[-1]
        if out is None:
            out = np.zeros(s * num_nodes)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == s * num_nodes
        q = self._quad_only(weights, points)
        self._check(self._lib.fb200_assemble_vector(self._h, C.byref(q), s, nat.ptr(f), int(f.ndim == 3), scatter_mode, int(accumulate), nat.ptr(out)))
        return out

    def assemble_elliptic_vector(self, op_kind: int, weights, points, data, u: np.ndarray, out: Optional[np.ndarray] = None,
                                 scatter_mode: int = nat.SCATTER_ATOMIC, accumulate: bool = False) -> np.ndarray:
        """elliptic.rs:456-526 through VectorAssembler (global.rs:569-686): internal-force-like vector of the elliptic operator at u."""
        uv = nat.as_f64(u)
        if out is None:
            out = np.zeros_like(uv)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == uv.size
        op, q = self._structs(op_kind, weights, points, data)
        self._check(self._lib.fb200_assemble_elliptic_vector(self._h, C.byref(op), C.byref(q), nat.ptr(uv), scatter_mode, int(accumulate), nat.ptr(out)))
        return out

    def assemble_elliptic_scalar(self, op_kind: int, weights, points, data, u: np.ndarray) -> float:
        """elliptic.rs:545-605 through assemble_scalar (global.rs:697-722): the elliptic energy of u."""
        uv = nat.as_f64(u)
        op, q = self._structs(op_kind, weights, points, data)
        e = C.c_double(0.0)
        self._check(self._lib.fb200_assemble_elliptic_scalar(self._h, C.byref(op), C.byref(q), nat.ptr(uv), C.byref(e)))
        return float(e.value)

    def _rule_structs(self, op_kind: int, rules, element_rule):
        keep, qs = [], (nat.Quadrature * len(rules))()
        for r, (weights, points, data) in enumerate(rules):
            _, q = self._structs(op_kind, weights, points, data)
            keep.append(self._keep)
            qs[r] = q
        er = np.ascontiguousarray(element_rule, dtype=np.uint32)
        self._keep = keep
        return nat.Operator(op_kind), qs, er

    def assemble_elliptic_vector_table(self, op_kind: int, rules, element_rule, u: np.ndarray, out: Optional[np.ndarray] = None,
                                       scatter_mode: int = nat.SCATTER_ATOMIC, accumulate: bool = False) -> np.ndarray:
        """assemble_elliptic_vector with a rule per element (rules / element_rule as assemble_into_csr_table_device)."""
        uv = nat.as_f64(u)
        if out is None:
            out = np.zeros_like(uv)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == uv.size
        op, qs, er = self._rule_structs(op_kind, rules, element_rule)
        self._check(self._lib.fb200_assemble_elliptic_vector_table(self._h, C.byref(op), len(rules), qs, nat.ptr(er), nat.ptr(uv), scatter_mode,
                                                                   int(accumulate), nat.ptr(out)))
        return out

    def assemble_elliptic_scalar_table(self, op_kind: int, rules, element_rule, u: np.ndarray) -> float:
        uv = nat.as_f64(u)
        op, qs, er = self._rule_structs(op_kind, rules, element_rule)
        e = C.c_double(0.0)
        self._check(self._lib.fb200_assemble_elliptic_scalar_table(self._h, C.byref(op), len(rules), qs, nat.ptr(er), nat.ptr(uv), C.byref(e)))
        return float(e.value)

    def physical_quadrature_points(self, weights, points, num_elements: int) -> np.ndarray:
        q = self._quad_only(weights, points)
        d = q.dim
        out = np.zeros((max(num_elements, 1), len(self._keep[0]), d))
        self._check(self._lib.fb200_physical_quadrature_points(self._h, C.byref(q), nat.ptr(out)))
        return out[:num_elements]

    def apply_homogeneous_dirichlet_bc_csr(self, nodes) -> float:
        """global.rs:379-451 on the device-resident values; returns the diagonal scale that was used."""
        n = nat.as_u64(nodes)
        cnt = len(n)
        if cnt == 0:
            n = np.zeros(1, dtype=np.uint64)
        scale = C.c_double(0.0)
        self._check(self._lib.fb200_apply_homogeneous_dirichlet_bc_csr(self._h, cnt, nat.ptr(n), C.byref(scale)))
        return float(scale.value)

    def spmv(self, x: np.ndarray) -> np.ndarray:
        """y = A x with the device-resident matrix."""
        xv = nat.as_f64(x)
        y = np.zeros_like(xv)
        self._check(self._lib.fb200_spmv(self._h, nat.ptr(xv), nat.ptr(y)))
        return y

    def cg_solve(self, b: np.ndarray, x0: Optional[np.ndarray] = None, rel_tol: float = 1e-8, max_iter: int = 0, jacobi: bool = True):
        """ConjugateGradient::solve_with_guess (fenris-sparse/src/cg.rs:364-480) on the device-resident matrix.
        Returns (x, iterations, relative residual); raises Fb200Error(ERR_NOT_CONVERGED | ERR_INDEFINITE)."""
        bv = nat.as_f64(b)
        x = np.zeros_like(bv) if x0 is None else nat.as_f64(x0).copy()
        it, res = C.c_uint64(0), C.c_double(0.0)
        self._check(self._lib.fb200_cg_solve(self._h, nat.ptr(bv), nat.ptr(x), float(rel_tol), int(max_iter), int(jacobi), C.byref(it), C.byref(res)))
        return x, int(it.value), float(res.value)

    def assemble_into_csr_table_device(self, op_kind: int, rules, element_rule, scatter_mode: int = nat.SCATTER_ATOMIC, accumulate: bool = False,
                                       u: Optional[np.ndarray] = None):
        """rules: sequence of (weights, points, data) - a CompactQuadratureTable; element_rule: rule index per element; u: the state of a
        non-linear operator (STVK, NEO_HOOKEAN)."""
        op, qs, er = self._rule_structs(op_kind, rules, element_rule)
        uv = None if u is None else nat.as_f64(u).reshape(-1)
        self._check(self._lib.fb200_assemble_into_csr_table_device(self._h, C.byref(op), len(rules), qs, nat.ptr(er), None if uv is None else nat.ptr(uv),
                                                                   scatter_mode, int(accumulate)))

    def values_download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.zeros(max(self.nnz, 1))
        self._check(self._lib.fb200_values_download(self._h, nat.ptr(out)))
        return out[: self.nnz]

    def values_upload(self, values: np.ndarray):
        v = nat.as_f64(values)
        self._check(self._lib.fb200_values_upload(self._h, nat.ptr(v)))

    def values_device_ptr(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64(0)
        self._check(self._lib.fb200_values_device(self._h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def element_matrices(self, op_kind: int, weights, points, data, first: int, count: int, dofs: int, u=None) -> np.ndarray:
        """Dense K_e of elements [first, first + count); u = the state for StVK / NeoHookean (None = zeros), ignored by linear operators."""
        op, q = self._structs(op_kind, weights, points, data)
        out = np.zeros((max(count, 1), dofs, dofs))
        uv = None if u is None else nat.as_f64(u)
        self._check(self._lib.fb200_element_matrices_u(self._h, C.byref(op), C.byref(q), None if uv is None else nat.ptr(uv), first, count, nat.ptr(out)))
        # column-major per element -> numpy [e][row][col]
        return np.transpose(out[:count], (0, 2, 1)).copy()

    # -- multi GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        st = nat.lib().fb200_comm_unique_id(buf)
        if st != nat.OK:
            raise Fb200Error(st, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, num_ranks: int):
        assert len(unique_id) == 128
        self._check(self._lib.fb200_comm_init(self._h, unique_id, rank, num_ranks))

    def interface_set(self, local_nodes: np.ndarray, packed_offsets: np.ndarray, packed_len: int):
        n, o = nat.as_u64(local_nodes), nat.as_u64(packed_offsets)
        cnt = len(n)
        if cnt == 0:
            n = np.zeros(1, dtype=np.uint64)
            o = np.zeros(1, dtype=np.uint64)
        self._check(self._lib.fb200_interface_set(self._h, cnt, nat.ptr(n), nat.ptr(o), packed_len))

    def interface_set_peers(self, peers):
        """peers: sequence of (peer_rank, local_node_ids) - the neighbour-exchange form of the interface (partition.py)."""
        ranks = np.ascontiguousarray([int(r) for r, _ in peers] or [0], dtype=np.int32)
        lists = [nat.as_u64(n) for _, n in peers]
        begin = np.zeros(len(peers) + 1, dtype=np.uint64)
        if lists:
            begin[1:] = np.cumsum([len(x) for x in lists])
        nodes = np.concatenate(lists) if lists and int(begin[-1]) else np.zeros(1, dtype=np.uint64)
        self._check(self._lib.fb200_interface_set_peers(self._h, len(peers), nat.ptr(ranks), nat.ptr(begin), nat.ptr(nat.as_u64(nodes))))

    def interface_allreduce(self):
        self._check(self._lib.fb200_interface_allreduce(self._h))

    def interface_enable_p2p(self) -> bool:
        """Fuse the interface exchange into the tile kernel's flush (peer-mapped values over NVLink).  Collective over the neighbours.
        Returns False when the partition / the machine cannot use it (the packed ncclSend/ncclRecv exchange then stays in use)."""
        st = self._lib.fb200_interface_enable_p2p(self._h)
        if st == nat.ERR_UNSUPPORTED:
            return False
        self._check(st)
        return True
