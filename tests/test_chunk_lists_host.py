"""Host-side checks of the Tet4 chunk lists (fenris_b200/csrc/chunks.cpp) - no GPU: the lists decide which contributions the chunk kernel
(tet4_chunk_kernel.cuh, config C5) sums into which CSR block, which rows it may write with plain stores and which rows the fused
peer-memory exchange forwards to the neighbouring rank.  fb200_chunk_lists_selftest verifies their invariants exhaustively on the
reference's BCC tet meshes (procedural.rs:286-403), whole and cut into element-range slabs with ghost elements (partition.py)."""
import ctypes as C

import numpy as np
import pytest

import fenris_b200 as fb
from fenris_b200 import _native as nat
from fenris_b200 import partition
from oracle import fenris_oracle as fo


def _selftest(v, c, owned=None, chunk=1024, sdim=3):
    v = np.ascontiguousarray(v, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.uint64)
    stats = (C.c_uint64 * 4)()
    failed = C.c_int32(0)
    st = nat.lib().fb200_chunk_lists_selftest(len(v), nat.ptr(v), len(c), nat.ptr(c), len(c) if owned is None else owned, chunk, sdim, stats,
                                              C.byref(failed))
    return st, failed.value, list(stats)


@pytest.mark.parametrize("n,chunk,sdim", [(1, 1024, 3), (3, 512, 3), (5, 1024, 1), (6, 256, 3), (8, 2048, 3)])
def test_whole_meshes(n, chunk, sdim):
    m = fb.create_unit_box_uniform_tet_mesh_3d(n)
    st, failed, (chunks, slots, complete, iface) = _selftest(m.vertices(), m.connectivity(), chunk=chunk, sdim=sdim)
    assert st == nat.OK, f"check {failed} failed"
    assert chunks == -(-m.num_elements() // chunk) and iface == 0
    assert slots >= m.num_elements()  # at least the 16 blocks of an element's nodes, shared between neighbours
    if chunks == 1:
        # a single chunk holds every element: every row is complete, and there is one slot per node block of the pattern
        ro, _ = fo.assemble_pattern_fast(1, m.num_nodes(), m.connectivity().astype(np.int64))
        assert complete == slots == int(ro[-1])


def test_jittered_renumbered_mesh():
    m = fb.create_unit_box_uniform_tet_mesh_3d(5)
    v = fo.jitter_vertices(m.vertices(), 1.0 / 5, amp=0.15)
    rng = np.random.default_rng(11)
    perm = rng.permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    c = inv[m.connectivity().astype(np.int64)][rng.permutation(m.num_elements())]
    st, failed, s = _selftest(v[perm], c, chunk=512)
    assert st == nat.OK, f"check {failed} failed"


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_with_ghost_elements(world):
    # config C5's partition: rows that ghost elements touch are flagged as interface rows and are never complete (no plain stores):
    # the neighbouring rank adds to them as well, by the packed exchange or by reductions over peer memory
    n = 6
    m = fb.create_unit_box_uniform_tet_mesh_3d(n)
    v, c = m.vertices(), m.connectivity()
    layer_starts = partition.tet_box_layer_starts(n, n, n)
    starts = layer_starts[partition.split_layers(n, world)]
    total_iface = 0
    for rank in range(world):
        part = partition.element_range_partition(v, c, starts, rank)
        st, failed, (chunks, slots, complete, iface) = _selftest(part["vertices"], part["connectivity"], owned=part["num_owned"], chunk=256)
        assert st == nat.OK, f"rank {rank}: check {failed} failed"
        assert iface > 0 and complete > 0
        total_iface += iface
        # without the ghosts the same owned elements see more complete rows (the interface rows would be stored blindly)
        st2, _, s2 = _selftest(part["vertices"], part["connectivity"][:part["num_owned"]], chunk=256)
        assert st2 == nat.OK and s2[2] > complete and s2[3] == 0
    assert total_iface > 0


def test_bad_arguments():
    m = fb.create_unit_box_uniform_tet_mesh_3d(2)
    c = m.connectivity().copy()
    c[0, 0] = m.num_nodes()
    assert _selftest(m.vertices(), c)[0] == nat.ERR_INDEX_OOB
    assert _selftest(m.vertices(), m.connectivity(), chunk=0)[0] == nat.ERR_SHAPE
    assert _selftest(m.vertices(), m.connectivity(), owned=m.num_elements() + 1)[0] == nat.ERR_SHAPE
