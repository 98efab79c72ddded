"""Host-side checks of the Hex8 tile lists (fenris_b200/csrc/tiles.cpp) - no GPU: the lists decide where every element block of
the tile kernel lands, so their invariants are verified exhaustively by fb200_tile_lists_selftest on several meshes
(structured, renumbered + reordered, a partition with ghost elements, degenerate elements)."""
import ctypes as C

import numpy as np
import pytest

import fenris_b200 as fb
from fenris_b200 import _native as nat
from fenris_b200.partition import structured_hex_slab
from oracle import fenris_oracle as fo


def _selftest(v, c, owned=None):
    v = np.ascontiguousarray(v, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.uint64)
    stats = (C.c_uint64 * 8)()
    failed = C.c_int32(0)
    st = nat.lib().fb200_tile_lists_selftest(len(v), nat.ptr(v), len(c), nat.ptr(c), len(c) if owned is None else owned, stats, C.byref(failed))
    return st, failed.value, list(stats)


def _selftest_ex(v, c, flush_rot, owned=None):
    v = np.ascontiguousarray(v, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.uint64)
    stats = (C.c_uint64 * 10)()
    failed = C.c_int32(0)
    st = nat.lib().fb200_tile_lists_selftest_ex(len(v), nat.ptr(v), len(c), nat.ptr(c), len(c) if owned is None else owned, flush_rot, stats,
                                                C.byref(failed))
    return st, failed.value, list(stats)


@pytest.mark.parametrize("n", [1, 3, 4, 8, 13])
def test_structured_cubes(n):
    m = fb.create_unit_box_uniform_hex_mesh_3d(n)
    st, failed, s = _selftest(m.vertices(), m.connectivity())
    assert st == nat.OK, f"check {failed} failed"
    tiles, max_nodes, max_pos, flush, complete, conflict_ppm, positions, max_rounds = s
    assert tiles == ((n + 3) // 4) ** 3 and max_nodes <= 125 and max_pos <= 1216
    if n % 4 == 0:
        # 4 x 4 x 4 tiles: 125 nodes and 2197 node blocks each, 8 colours of 8 elements; complete = nodes off the inner tile faces
        assert flush == tiles * 2197 and complete == (n + 2 - n // 4) ** 3 and max_rounds == 8 and positions == n ** 3
    assert conflict_ppm < 30000  # the bank-aware accumulator positions keep collisions rare


def test_renumbered_reordered_jittered_mesh():
    m = fb.create_unit_box_uniform_hex_mesh_3d(9)
    v = fo.jitter_vertices(m.vertices(), 1.0 / 9, amp=0.2)
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    c = inv[m.connectivity().astype(np.int64)][rng.permutation(m.num_elements())]
    st, failed, s = _selftest(v[perm], c)
    assert st == nat.OK, f"check {failed} failed"
    assert s[6] >= m.num_elements()  # every element has a schedule position (plus padding)


def test_partition_with_ghost_elements():
    # rank 1 of 3 z-slabs: owned elements first, ghost layers after them (pattern only): ghost-touched nodes are never complete
    verts, conn, n_owned, _ = structured_hex_slab(8, 8, 12, 1.0 / 8, 1, 3)
    st, failed, s = _selftest(verts, conn, owned=n_owned)
    assert st == nat.OK, f"check {failed} failed"
    st2, _, s2 = _selftest(verts, conn[:n_owned])  # the same elements without the ghosts: more complete nodes
    assert st2 == nat.OK and s2[4] > s[4]


def test_elements_with_repeated_nodes_are_refused():
    m = fb.create_unit_box_uniform_hex_mesh_3d(3)
    c = m.connectivity().copy()
    c[4, 6] = c[4, 2]
    st, _, _ = _selftest(m.vertices(), c)
    assert st == nat.ERR_UNSUPPORTED  # the library then keeps the per-element kernel (tests/test_hex8_tile.py)


def test_index_out_of_bounds():
    m = fb.create_unit_box_uniform_hex_mesh_3d(2)
    c = m.connectivity().copy()
    c[0, 0] = m.num_nodes()
    assert _selftest(m.vertices(), c)[0] == nat.ERR_INDEX_OOB


def test_rotated_flush_lists_keep_every_invariant_and_spread_the_banks():
    # opt-in fb200_set_tuning("hex8_flush_rot"): the flush words carry a rotation (bits 30-31) of the block row a lane reads first.
    # Same lists otherwise (all 21 checks), and the modelled shared-memory wavefronts per flush load drop (the model reproduces the
    # ncu source counters of the shipped order: 6.2 vs 5.6-6.6 measured, profiles/r01/README.md)
    m = fb.create_unit_box_uniform_hex_mesh_3d(12)
    st0, failed0, s0 = _selftest_ex(m.vertices(), m.connectivity(), 0)
    st1, failed1, s1 = _selftest_ex(m.vertices(), m.connectivity(), 1)
    assert st0 == nat.OK and st1 == nat.OK, (failed0, failed1)
    assert s0[:8] == s1[:8] and s0[8] == s1[8] > 0
    per_load0, per_load1 = s0[9] / s0[8], s1[9] / s1[8]
    assert 5.5 < per_load0 < 6.8 and per_load1 < 0.72 * per_load0 and per_load1 >= 2.0
    # unstructured numbering, ghosts
    v = fo.jitter_vertices(m.vertices(), 1.0 / 12, amp=0.2)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    c = inv[m.connectivity().astype(np.int64)][rng.permutation(m.num_elements())]
    st, failed, s = _selftest_ex(v[perm], c, 1)
    assert st == nat.OK, f"check {failed} failed"
    verts, conn, n_owned, _ = structured_hex_slab(8, 8, 12, 1.0 / 8, 1, 3)
    st, failed, s = _selftest_ex(verts, conn, 1, owned=n_owned)
    assert st == nat.OK, f"check {failed} failed"
