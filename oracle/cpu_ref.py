"""ctypes wrapper around oracle/cpu_ref.c (C restatement of the reference CPU path).

TEST INFRASTRUCTURE / CPU BASELINE ONLY - see the header of cpu_ref.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfenris_cpu_ref.so")
_lib = None

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _host_tag() -> str:
    """The library is built with -march=native; a copy built on another host (the .so travels to the
    GPU box with the repo snapshot) must be rebuilt there, so the build is stamped with the CPU flags."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False) -> str:
    """Compile cpu_ref.c when the library is missing, older than the source or was built on another host.  Several processes may call this
    at once (one bench rank per GPU, each checking its rows): the build runs under a file lock, into a private file that is renamed into
    place, so nobody ever loads a half-written library."""
    import fcntl
    src = os.path.join(_HERE, "cpu_ref.c")
    bdir = os.path.join(_HERE, "_build")
    stamp = os.path.join(bdir, "host.tag")
    tag = _host_tag()

    def stale() -> bool:
        old = open(stamp).read().strip() if os.path.exists(stamp) else ""
        return (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src) or old != tag

    if force or stale():
        os.makedirs(bdir, exist_ok=True)
        with open(os.path.join(bdir, ".lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():  # (another process may have built it while this one waited)
                    tmp = os.path.join("_build", f"libfenris_cpu_ref.{os.getpid()}.so")
                    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", f"OUT={tmp}"])
                    os.replace(os.path.join(_HERE, tmp), _SO)
                    with open(stamp + f".{os.getpid()}", "w") as f:
                        f.write(tag)
                    os.replace(stamp + f".{os.getpid()}", stamp)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oref_gen_tet_mesh.restype = C.c_uint64
    return _lib


def max_threads() -> int:
    return int(lib().oref_max_threads())


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def gen_hex_mesh(cells: int, cz: Optional[int] = None, cell_size: Optional[float] = None):
    cx = cy = cells
    cz = cells if cz is None else cz
    h = (1.0 / cells) if cell_size is None else cell_size
    v = np.empty(((cx + 1) * (cy + 1) * (cz + 1), 3))
    c = np.empty((cx * cy * cz, 8), dtype=np.uint64)
    lib().oref_gen_hex_mesh(C.c_uint64(cx), C.c_uint64(cy), C.c_uint64(cz), C.c_double(h), _ptr(v, C.c_double), _ptr(c, C.c_uint64))
    return v, c


def gen_quad_mesh(cells: int):
    v = np.empty(((cells + 1) ** 2, 2))
    c = np.empty((cells * cells, 4), dtype=np.uint64)
    lib().oref_gen_quad_mesh(C.c_uint64(cells), C.c_double(1.0 / cells), _ptr(v, C.c_double), _ptr(c, C.c_uint64))
    return v, c


def gen_tet_mesh(cells: int):
    n = cells
    v = np.empty(((n + 1) ** 3 + n ** 3, 3))
    c = np.empty((12 * n ** 3, 4), dtype=np.uint64)
    cnt = lib().oref_gen_tet_mesh(C.c_uint64(n), C.c_uint64(n), C.c_uint64(n), C.c_double(1.0 / float(n)), _ptr(v, C.c_double), _ptr(c, C.c_uint64))
    assert cnt == len(c)
    return v, c


def _ragged(conn):
    if isinstance(conn, np.ndarray) and conn.ndim == 2:
        E, n = conn.shape
        offs = (np.arange(E + 1, dtype=np.uint64) * np.uint64(n))
        return offs, np.ascontiguousarray(conn, dtype=np.uint64).ravel()
    offs = np.zeros(len(conn) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(e) for e in conn])
    flat = np.array([x for e in conn for x in e], dtype=np.uint64)
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.uint64)
    return offs, flat


def pattern(sdim: int, num_nodes: int, conn) -> Tuple[np.ndarray, np.ndarray]:
    offs, flat = _ragged(conn)
    row_offsets = np.zeros(sdim * num_nodes + 1, dtype=np.uint64)
    cols_p = C.POINTER(C.c_uint64)()
    nnz = C.c_uint64(0)
    st = lib().oref_pattern(C.c_int(sdim), C.c_uint64(num_nodes), C.c_uint64(len(offs) - 1), _ptr(offs, C.c_uint64),
                            _ptr(flat, C.c_uint64), _ptr(row_offsets, C.c_uint64), C.byref(cols_p), C.byref(nnz))
    if st:
        raise RuntimeError(f"oref_pattern failed: {st}")
    cols = np.ctypeslib.as_array(cols_p, shape=(max(nnz.value, 1),))[: nnz.value].copy()
    lib().oref_free(cols_p)
    return row_offsets, cols


def color_greedy(conn, num_nodes: int):
    offs, flat = _ragged(conn)
    E = len(offs) - 1
    elems = np.zeros(max(E, 1), dtype=np.uint64)
    offs_p = C.POINTER(C.c_uint64)()
    ncol = C.c_uint64(0)
    lib().oref_color_greedy(C.c_uint64(E), _ptr(offs, C.c_uint64), _ptr(flat, C.c_uint64), C.c_uint64(num_nodes),
                            C.byref(offs_p), _ptr(elems, C.c_uint64), C.byref(ncol))
    coffs = np.ctypeslib.as_array(offs_p, shape=(ncol.value + 1,)).copy()
    lib().oref_free(offs_p)
    return coffs, elems[:E]


def _params(op, q, params):
    if op == 1:
        return None
    p = np.asarray(params, dtype=np.float64)
    if p.ndim == 1:
        p = np.tile(p, (q, 1))
    return np.ascontiguousarray(p)


def element_matrix(elem_type, op, weights, points, params, X):
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = lib().oref_elem_nodes(elem_type)
    d = lib().oref_elem_dim(elem_type)
    s = 1 if op == 1 else d
    K = np.zeros((s * n, s * n), order="F")
    par = _params(op, len(w), params)
    Xc = np.ascontiguousarray(X, dtype=np.float64)
    st = lib().oref_element_matrix(C.c_int(elem_type), C.c_int(op), C.c_int(len(w)), _ptr(w, C.c_double), _ptr(pts, C.c_double),
                                   _ptr(par, C.c_double) if par is not None else None, _ptr(Xc, C.c_double), _ptr(K, C.c_double))
    if st:
        raise RuntimeError(f"oref_element_matrix: status {st}")
    return np.ascontiguousarray(K)


def assemble(elem_type, op, weights, points, params, vertices, conn, row_offsets, col_indices, values=None,
             colors=None, nthreads: int = 0):
    """Serial (colors=None) or coloured-threaded assembly, accumulating into `values`."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pts = np.ascontiguousarray(points, dtype=np.float64)
    par = _params(op, len(w), params)
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    c = np.ascontiguousarray(conn, dtype=np.uint64)
    ro = np.ascontiguousarray(row_offsets, dtype=np.uint64)
    ci = np.ascontiguousarray(col_indices, dtype=np.uint64)
    if values is None:
        values = np.zeros(len(ci))
    bad = C.c_int64(-1)
    parp = _ptr(par, C.c_double) if par is not None else None
    if colors is None:
        st = lib().oref_assemble_serial(C.c_int(elem_type), C.c_int(op), C.c_int(len(w)), _ptr(w, C.c_double), _ptr(pts, C.c_double), parp,
                                        _ptr(v, C.c_double), C.c_uint64(len(c)), _ptr(c, C.c_uint64), _ptr(ro, C.c_uint64),
                                        _ptr(ci, C.c_uint64), _ptr(values, C.c_double), C.byref(bad))
    else:
        coffs, celems = colors
        coffs = np.ascontiguousarray(coffs, dtype=np.uint64)
        celems = np.ascontiguousarray(celems, dtype=np.uint64)
        st = lib().oref_assemble_colored(C.c_int(elem_type), C.c_int(op), C.c_int(len(w)), _ptr(w, C.c_double), _ptr(pts, C.c_double), parp,
                                         _ptr(v, C.c_double), _ptr(c, C.c_uint64), C.c_uint64(len(coffs) - 1), _ptr(coffs, C.c_uint64),
                                         _ptr(celems, C.c_uint64), _ptr(ro, C.c_uint64), _ptr(ci, C.c_uint64), _ptr(values, C.c_double),
                                         C.c_int(nthreads), C.byref(bad))
    if st == 1:
        raise ArithmeticError(f"Singular element Jacobian encountered (element {bad.value})")
    if st:
        raise RuntimeError(f"cpu_ref assembly failed with status {st} at element {bad.value}")
    return values
