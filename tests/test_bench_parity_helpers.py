"""CPU checks of bench.py's own parity helpers (the `parity` field of every bench line): they must accept a correct matrix and reject a
wrong one, for the single-GPU layout and for the slab layouts with ghost planes - otherwise a green `parity` in a scaling run proves nothing.
The "device result" is emulated with the oracle: all local elements (owned + ghost) assembled on the rank-local mesh, which is what the
owned assembly + interface exchange leaves in the rows the helpers look at."""
import numpy as np
import pytest

import bench
from fenris_b200 import partition
from oracle import cpu_ref as cr
from oracle import fenris_oracle as fo

MU, LAM = fo.lame_from_young_poisson(bench.YOUNG, bench.POISSON)


def _local_matrix(verts, conn):
    ro, ci = cr.pattern(3, len(verts), conn)
    w, p = fo.hexahedron_gauss(2)
    vals = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (MU, LAM), verts, conn, ro, ci)
    return ro, vals


@pytest.mark.parametrize("world", [1, 2, 3])
def test_parity_c3_accepts_the_right_matrix_and_rejects_a_wrong_one(world):
    cells = 6
    h = 1.0 / cells
    for rank in range(world):
        if world == 1:
            verts, conn = cr.gen_hex_mesh(cells)
        else:
            verts, conn, _, _ = partition.structured_hex_slab(cells, cells, cells * world, h, rank, world)
        ro, vals = _local_matrix(verts, conn)
        err, rows = bench.parity_c3(None, vals, cells, h, rank, world, (MU, LAM))
        planes = 1 + (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
        assert rows == planes * 3 * (cells + 1) ** 2
        assert err < 1e-13, (world, rank, err)
        # a perturbation of one entry of a checked row must show
        g0 = 1 if rank > 0 else 0
        node = (cells + 1) ** 2 * (g0 + cells // 2) + 5
        bad = vals.copy()
        bad[int(ro[3 * node]) + 2] *= 1.0 + 1e-6
        err_bad, _ = bench.parity_c3(None, bad, cells, h, rank, world, (MU, LAM))
        assert err_bad > 1e-10


def test_parity_c3_sees_a_missing_interface_contribution():
    # without the exchange the interface plane of a slab lacks the neighbour's elements: the helper must flag it
    cells, world, rank = 4, 2, 0
    h = 1.0 / cells
    verts, conn, n_owned, _ = partition.structured_hex_slab(cells, cells, cells * world, h, rank, world)
    ro, ci = cr.pattern(3, len(verts), conn)
    w, p = fo.hexahedron_gauss(2)
    owned_only = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (MU, LAM), verts, conn[:n_owned], ro, ci)
    err, _ = bench.parity_c3(None, owned_only, cells, h, rank, world, (MU, LAM))
    assert err > 1e-2


class _FakeCtx:
    def __init__(self, ro):
        self._ro = ro

    def row_offsets_download(self):
        return self._ro


@pytest.mark.parametrize("world", [1, 2])
def test_parity_sampled_rows_c5_layout(world):
    # config C5's check: sampled owned + interface rows of a rank-local Tet4 matrix against the oracle on the sub-mesh of the global
    # elements around them; emulated device result = all local elements (owned + ghost) assembled on the local mesh
    n = 5
    gv, gc = cr.gen_tet_mesh(n)
    w, p = fo.tetrahedron_rule(1)
    layer_starts = partition.tet_box_layer_starts(n, n, n)
    starts = layer_starts[partition.split_layers(n, world)]
    for rank in range(world):
        if world == 1:
            part = {"vertices": gv, "connectivity": gc, "num_owned": len(gc), "global_nodes": np.arange(len(gv)), "peers": []}
        else:
            part = partition.element_range_partition(gv, gc, starts, rank)
        lv, lc = part["vertices"], np.ascontiguousarray(part["connectivity"], dtype=np.uint64)
        ro, ci = cr.pattern(3, len(lv), lc)
        vals = cr.assemble(fo.TET4, fo.LINEAR_ELASTIC, w, p, (MU, LAM), lv, lc, ro, ci)
        err, rows = bench.parity_sampled_rows(_FakeCtx(ro), vals, fo.TET4, gv, gc, part, (MU, LAM), (w, p), 3 + rank)
        assert rows > 0 and err < 1e-13, (world, rank, err)
        bad = vals.copy()
        owned_nodes = np.unique(lc[:part["num_owned"]].astype(np.int64))
        bad[int(ro[3 * owned_nodes[len(owned_nodes) // 2]]):int(ro[3 * owned_nodes[len(owned_nodes) // 2] + 3])] *= 1.001
        # (the sample covers every owned node of such a small mesh, so the perturbed row is among the checked ones)
        err_bad, _ = bench.parity_sampled_rows(_FakeCtx(ro), bad, fo.TET4, gv, gc, part, (MU, LAM), (w, p), 3 + rank)
        assert err_bad > 1e-8
