"""World-size-2 (and 3) `gloo` tests of the multi-GPU host logic on CPU: element partition with ghost layers, interface
packing layout and the all-reduce step.  Each rank assembles its OWNED elements with the oracle (the device kernels are
covered by the gpu tests), packs its interface rows exactly as fb200_interface_allreduce does, all-reduces the packed
buffer over gloo, unpacks, and the stitched rows must equal a single-process assembly of the whole mesh."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fenris_b200 import partition
from oracle import fenris_oracle as fo

MU, LAM = fo.lame_from_young_poisson(1e6, 0.2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _assemble_owned(verts, conn, n_owned, sdim_op):
    prob_all = fo.Problem(fo.HEX8, verts, conn.astype(np.int64), sdim_op, params=(MU, LAM))
    ro, ci = fo.assemble_pattern_fast(prob_all.sdim, len(verts), conn.astype(np.int64))  # ghosts complete the pattern
    values = np.zeros(len(ci))
    K = fo.element_matrices_fast(prob_all)
    for e in range(n_owned):  # ghosts are NOT assembled
        fo.scatter_element(values, ro, ci, prob_all.sdim, conn[e].astype(np.int64).tolist(), K[e])
    return ro, ci, values


def _worker(rank, world, port, cx, cy, cz, exchange, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = 1.0 / cx
        verts, conn, n_owned, iface = partition.structured_hex_slab(cx, cy, cz, h, rank, world)
        ro, ci, values = _assemble_owned(verts, conn, n_owned, fo.LINEAR_ELASTIC)
        s = 3
        # pack (same arithmetic as iface_copy_kernel<PACK>): node row block = values[s*s*blk_off[node] ...]
        if exchange == "peers":
            # neighbour exchange (fb200_interface_set_peers): pack the rows shared with each peer, swap, add what arrives
            sends, recvs, reqs = [], [], []
            for peer, nodes in iface["peers"]:
                seg = np.concatenate([values[int(ro[s * n]):int(ro[s * n + s])] for n in nodes.astype(np.int64)])
                sends.append(torch.from_numpy(seg.copy()))
                recvs.append(torch.zeros(len(seg), dtype=torch.float64))
                reqs.append(dist.isend(sends[-1], dst=peer))
                reqs.append(dist.irecv(recvs[-1], src=peer))
            for r in reqs:
                r.wait()
            for (peer, nodes), got in zip(iface["peers"], recvs):
                got, o = got.numpy(), 0
                for n in nodes.astype(np.int64):
                    b, e = int(ro[s * n]), int(ro[s * n + s])
                    values[b:e] += got[o:o + (e - b)]
                    o += e - b
        else:
            packed = np.zeros(iface["packed_len"])
            for node, off in zip(iface["local_nodes"].astype(np.int64), iface["packed_offsets"].astype(np.int64)):
                b, e = int(ro[s * node]), int(ro[s * node + s])
                packed[off:off + (e - b)] = values[b:e]
            t = torch.from_numpy(packed)
            dist.all_reduce(t)
            for node, off in zip(iface["local_nodes"].astype(np.int64), iface["packed_offsets"].astype(np.int64)):
                b, e = int(ro[s * node]), int(ro[s * node + s])
                values[b:e] = packed[off:off + (e - b)]
        # rows of the nodes in the planes this rank's owned cells touch, with global ids
        per = cz // world
        plane = (cx + 1) * (cy + 1)
        first = iface["first_global_node"]
        g_lo, g_hi = plane * rank * per, plane * ((rank + 1) * per + 1)
        rows = {}
        for g in range(g_lo, g_hi):
            l = g - first
            for i in range(s):
                b, e = int(ro[s * l + i]), int(ro[s * l + i + 1])
                rows[s * g + i] = (ci[b:e].astype(np.int64) + s * first, values[b:e].copy())
        q.put((rank, rows, int(n_owned)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cz,exchange", [(2, 4, "allreduce"), (3, 6, "allreduce"), (2, 4, "peers"), (3, 6, "peers")])
def test_slab_partition_allreduce_equals_global(world, cz, exchange):
    cx = cy = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cx, cy, cz, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v, c = fo.create_rectangular_uniform_hex_mesh(1.0 / cx * cx, 1, 1, 1, cx) if cz == cx else (None, None)
    # global reference: cx x cy x cz box with cell size 1/cx
    from oracle import cpu_ref as cr
    vg, cg = cr.gen_hex_mesh(cx, cz=cz, cell_size=1.0 / cx)
    prob = fo.Problem(fo.HEX8, vg, cg.astype(np.int64), fo.LINEAR_ELASTIC, params=(MU, LAM))
    ro, ci, vals = fo.assemble_fast(prob)
    assert sum(r[2] for r in results) == len(cg)
    checked = 0
    for rank, rows, _ in results:
        for grow, (cols, v) in rows.items():
            b, e = int(ro[grow]), int(ro[grow + 1])
            assert np.array_equal(cols, ci[b:e].astype(np.int64)), (rank, grow)
            assert np.allclose(v, vals[b:e], rtol=1e-13, atol=1e-9 * np.abs(vals).max()), (rank, grow)
            checked += 1
    assert checked >= len(ro) - 1  # every global row is held (complete) by at least one rank


def test_general_partition_peer_lists_are_symmetric():
    # the neighbour-exchange form: rank r's list for peer q and q's list for r name the same global nodes in the same order,
    # and together the lists of a rank cover exactly its interface nodes
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(3)
    ro, _ = fo.assemble_pattern_fast(1, len(v), c)
    centroid = v[c].mean(axis=1)
    part = (centroid[:, 0] > 0.5).astype(np.int64) + 2 * (centroid[:, 1] > 0.5).astype(np.int64)  # 4 parts meeting along a line
    layout = partition.general_partition(c, part, len(v), 4, 1, np.diff(ro.astype(np.int64)))
    for r, l in enumerate(layout):
        union = np.zeros(0, dtype=np.int64)
        for q, ids in l["peers"].items():
            assert q != r and np.array_equal(ids, layout[q]["peers"][r]) and np.all(np.diff(ids) > 0)
            union = np.union1d(union, ids)
        assert np.array_equal(union, l["iface_global"])
    assert any(len(l["peers"]) == 3 for l in layout)  # nodes on the common line are shared by all four ranks


def test_general_partition_layout_is_consistent():
    v, c = fo.create_unit_box_uniform_tet_mesh_3d(3)
    ro, ci = fo.assemble_pattern_fast(1, len(v), c)
    blocks = np.diff(ro.astype(np.int64))
    centroid_x = v[c].mean(axis=1)[:, 0]
    part = (centroid_x > 0.34).astype(np.int64) + (centroid_x > 0.67).astype(np.int64)
    layout = partition.general_partition(c, part, len(v), 3, 1, blocks)
    owned_all = np.sort(np.concatenate([l["owned"] for l in layout]))
    assert np.array_equal(owned_all, np.arange(len(c)))
    seen = {}
    for r, l in enumerate(layout):
        assert l["packed_len"] == layout[0]["packed_len"]
        lv, lc, gids = partition.localize(v, c, l["owned"], l["ghosts"])
        # every element touching an interface node of this rank is present locally (owned or ghost)
        loc_ro, _ = fo.assemble_pattern_fast(1, len(lv), lc.astype(np.int64))
        lookup = {g: i for i, g in enumerate(gids.tolist())}
        for g, off in zip(l["iface_global"].tolist(), l["packed_offsets"].tolist()):
            li = lookup[g]
            assert int(loc_ro[li + 1] - loc_ro[li]) == int(blocks[g]), "ghosts must complete interface rows"
            assert seen.setdefault(g, off) == off, "all sharing ranks use the same packed offset"
    assert len(seen) > 0


# ------------------------------------------------------------------------------------------------ element-range partitions (C5: Tet4 slabs)
def _range_mesh(kind):
    if kind == "tet4":
        v, c = fo.create_unit_box_uniform_tet_mesh_3d(3)
        layers = partition.tet_box_layer_starts(3, 3, 3)
        return fo.TET4, v, np.asarray(c, dtype=np.int64), layers
    v8, c8 = fo.create_unit_box_uniform_hex_mesh_3d(3)
    v, c = fo.hex27_mesh_from_hex8(v8, np.asarray(c8, dtype=np.int64))
    return fo.HEX27, np.asarray(v), np.asarray(c, dtype=np.int64), np.arange(4, dtype=np.int64) * 9  # 9 cells per z-layer


def _range_worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        et, v, c, layer_starts = _range_mesh(kind)
        starts = layer_starts[partition.split_layers(len(layer_starts) - 1, world)]
        part = partition.element_range_partition(v, c, starts, rank)
        lv, lc, n_owned, gids = part["vertices"], part["connectivity"].astype(np.int64), part["num_owned"], part["global_nodes"]
        prob = fo.Problem(et, lv, lc, fo.LINEAR_ELASTIC, params=(MU, LAM))
        s = prob.sdim
        ro, ci = fo.assemble_pattern_fast(s, len(lv), lc)  # ghosts complete the pattern
        values = np.zeros(len(ci))
        K = fo.element_matrices_fast(prob)
        for e in range(n_owned):  # ghosts are NOT assembled
            fo.scatter_element(values, ro, ci, s, lc[e].tolist(), K[e])
        # neighbour exchange exactly as fb200_interface_set_peers / interface_exchange_peers order the rows
        sends, recvs, reqs = [], [], []
        for peer, nodes in part["peers"]:
            seg = np.concatenate([values[int(ro[s * n]):int(ro[s * n + s])] for n in nodes.astype(np.int64)])
            sends.append(torch.from_numpy(seg.copy()))
            recvs.append(torch.zeros(len(seg), dtype=torch.float64))
            reqs.append(dist.isend(sends[-1], dst=peer))
            reqs.append(dist.irecv(recvs[-1], src=peer))
        for r in reqs:
            r.wait()
        for (peer, nodes), got in zip(part["peers"], recvs):
            got, o = got.numpy(), 0
            for n in nodes.astype(np.int64):
                b, e = int(ro[s * n]), int(ro[s * n + s])
                values[b:e] += got[o:o + (e - b)]
                o += e - b
        # rows of the nodes of this rank's OWNED elements, with global column ids
        rows = {}
        for l in np.unique(lc[:n_owned]).tolist():
            for i in range(s):
                b, e = int(ro[s * l + i]), int(ro[s * l + i + 1])
                cols = ci[b:e].astype(np.int64)
                rows[s * int(gids[l]) + i] = (s * gids[cols // s] + cols % s, values[b:e].copy())
        q.put((rank, rows, int(n_owned)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("tet4", 2), ("tet4", 3), ("hex27", 2)])
def test_element_range_partition_equals_global(kind, world):
    """partition.element_range_partition (config C5's slabs of the BCC tet mesh; Hex27 slabs): owned + ghost elements, neighbour lists in
    the same order on both sides; the exchanged rows equal a single-process assembly of the whole mesh."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_range_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    et, v, c, _ = _range_mesh(kind)
    prob = fo.Problem(et, v, c, fo.LINEAR_ELASTIC, params=(MU, LAM))
    ro, ci, vals = fo.assemble_fast(prob)
    assert sum(r[2] for r in results) == len(c)
    covered = set()
    for rank, rows, _ in results:
        for grow, (cols, val) in rows.items():
            b, e = int(ro[grow]), int(ro[grow + 1])
            assert np.array_equal(cols, ci[b:e].astype(np.int64)), (rank, grow)
            assert np.allclose(val, vals[b:e], rtol=1e-13, atol=1e-9 * np.abs(vals).max()), (rank, grow)
            covered.add(grow)
    assert len(covered) == len(ro) - 1
