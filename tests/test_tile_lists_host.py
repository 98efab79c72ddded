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


def _selftest_ex(v, c, owner_stores, owned=None):
    v = np.ascontiguousarray(v, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.uint64)
    stats = (C.c_uint64 * 10)()
    failed = C.c_int32(0)
    st = nat.lib().fb200_tile_lists_selftest_ex(len(v), nat.ptr(v), len(c), nat.ptr(c), len(c) if owned is None else owned, owner_stores, stats,
                                                C.byref(failed))
    return st, failed.value, list(stats)


@pytest.mark.parametrize("n", [1, 3, 4, 8, 13])
def test_structured_cubes(n):
    m = fb.create_unit_box_uniform_hex_mesh_3d(n)
    st, failed, s = _selftest(m.vertices(), m.connectivity())
    assert st == nat.OK, f"check {failed} failed"
    tiles, max_nodes, max_pos, flush, complete, conflict_ppm, positions, max_rounds = s
    assert tiles == ((n + 3) // 4) ** 3 and max_nodes <= 125 and max_pos <= 1216
    for owner in (0, 1):
        st, failed, sx = _selftest_ex(m.vertices(), m.connectivity(), owner)
        assert st == nat.OK, f"owner_stores={owner}: check {failed} failed"
        flush, store, zeros = sx[3], sx[8], sx[9]
        if n % 4 == 0:
            # 4 x 4 x 4 tiles: 125 nodes and 2197 node blocks each, 8 colours of 8 elements; complete = nodes off the inner tile faces
            assert flush == tiles * 2197 + zeros and complete == (n + 2 - n // 4) ** 3 and max_rounds == 8 and positions == n ** 3
        nnz_blocks = (3 * n + 1) ** 3  # node-block pattern of the structured mesh: (3n + 1)^3 coupled pairs
        if owner:
            # every row is stored exactly once, zeros included: the STORE segments cover the whole block pattern
            assert store == nnz_blocks and (zeros > 0) == (tiles > 1)
        else:
            assert zeros == 0 and store < nnz_blocks or tiles == 1
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


def test_owner_lists_on_unstructured_numbering_and_partitions():
    # first-writer ownership (fb200_set_tuning "hex8_owner_stores"): every row stored completely by exactly one tile, reducers wait for
    # a lower-numbered tile, rows that ghost elements touch are listed for clearing instead (checks 22-32 of the self test)
    m = fb.create_unit_box_uniform_hex_mesh_3d(12)
    v = fo.jitter_vertices(m.vertices(), 1.0 / 12, amp=0.2)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    c = inv[m.connectivity().astype(np.int64)][rng.permutation(m.num_elements())]
    for owner in (0, 1):
        st, failed, s = _selftest_ex(v[perm], c, owner)
        assert st == nat.OK, f"check {failed} failed"
        assert (s[8] == 37 ** 3) == bool(owner)
    for rank in range(3):
        verts, conn, n_owned, _ = structured_hex_slab(8, 8, 12, 1.0 / 8, rank, 3)
        for owner in (0, 1):
            st, failed, s = _selftest_ex(verts, conn, owner, owned=n_owned)
            assert st == nat.OK, f"rank {rank}: check {failed} failed"


@pytest.mark.parametrize("seed,keep", [(0, 0.7), (1, 0.4), (2, 0.9), (3, 0.995)])
def test_perforated_meshes_many_local_connectivities(seed, keep):
    """Cubes with a random subset of the elements removed (unused nodes stay in the space): partial tiles of every shape, i.e. many
    different tile-local connectivities - some repeated, most not.  Besides the list invariants this exercises the memo of local
    connectivities both ways (check 38 of the self test rebuilds every tile without it and compares all arrays)."""
    m = fb.create_unit_box_uniform_hex_mesh_3d(14)
    rng = np.random.default_rng(seed)
    c = m.connectivity()[rng.random(m.num_elements()) < keep]
    for owner in (0, 1):
        st, failed, s = _selftest_ex(m.vertices(), c, owner)
        assert st == nat.OK, f"check {failed} failed"
        assert s[6] >= len(c)
    # and as a partition: the removed elements' neighbours of the upper half are ghosts
    half = c[:, 0] < np.median(c[:, 0])
    cc = np.concatenate([c[half], c[~half]])
    st, failed, s = _selftest_ex(m.vertices(), cc, 1, owned=int(half.sum()))
    assert st == nat.OK, f"check {failed} failed"
