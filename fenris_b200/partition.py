"""Element partitioning for the multi-GPU path (host-side preprocessing, numpy only - no device work here).

The mesh is split by ELEMENTS; a rank assembles only the elements it owns.  Rows of nodes that touch elements of
more than one rank ("interface nodes") receive partial sums on each of them and are completed by
`fb200_interface_allreduce`: a neighbour exchange of the packed interface rows (ncclSend/ncclRecv per peer, summed on arrival) when the
interface is given as peer lists, else an ncclAllReduce over a packed buffer of all interface rows.  To make the CSR row layout
of an interface node identical on every sharing rank, each rank also receives the other ranks' elements that touch
its interface nodes as GHOST elements: they take part in the sparsity pattern but are not assembled
(`fb200_set_num_owned_elements`).

The reference has no distributed mode (README.md:58); this is the north-star extension of its element loop
(src/assembly/global.rs:314-376), whose colouring-based race avoidance becomes, across GPUs, ownership + one exchange.
"""
from __future__ import annotations

from typing import List

import numpy as np


def structured_hex_slab(cx: int, cy: int, cz: int, h: float, rank: int, nranks: int):
    """z-slab partition of the reference's structured hex mesh (src/mesh/procedural.rs:216-277).

    Returns (vertices, connectivity_local_u64, num_owned, iface) for `rank`:
      * local nodes = the contiguous node planes the slab (+ ghost layers) touches, numbered in the global order
        (so local id = global id - first_plane * (cx+1)(cy+1));
      * elements: owned cells first (global order), then the ghost cell layer(s);
      * iface: local ids of the interface-plane nodes, their offsets into the packed exchange buffer and its length -
        identical layout on all ranks.
    """
    assert cz % nranks == 0, "cells in z must divide evenly over the ranks"
    per = cz // nranks
    k0, k1 = rank * per, (rank + 1) * per
    g0 = k0 - (1 if rank > 0 else 0)
    g1 = k1 + (1 if rank < nranks - 1 else 0)
    vx, vy = cx + 1, cy + 1
    plane = vx * vy
    z0, z1 = g0, g1 + 1  # node planes [z0, z1)
    kk, jj, ii = np.meshgrid(np.arange(z0, z1), np.arange(vy), np.arange(vx), indexing="ij")
    verts = np.stack([ii.ravel() * h, jj.ravel() * h, kk.ravel() * h], axis=1).astype(np.float64)

    def cells(ka, kb):
        k, j, i = np.meshgrid(np.arange(ka, kb), np.arange(cy), np.arange(cx), indexing="ij")
        i, j, k = i.ravel(), j.ravel(), k.ravel()
        idx = lambda a, b, c: plane * (c - z0) + vx * b + a
        return np.stack([idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k),
                         idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)], axis=1)

    parts = [cells(k0, k1)]
    if rank > 0:
        parts.append(cells(k0 - 1, k0))
    if rank < nranks - 1:
        parts.append(cells(k1, k1 + 1))
    conn = np.concatenate(parts).astype(np.uint64)
    n_owned = cx * cy * per

    # interface planes: plane index p (1..nranks-1) lies at z = p*per, shared by ranks p-1 and p.
    # row block of node (i,j) on an interior z-plane couples to nx*ny*3 nodes, 9 values each (s = 3).
    nx = np.full(vx, 3)
    nx[[0, -1]] = 2
    ny = np.full(vy, 3)
    ny[[0, -1]] = 2
    if cx == 1:
        nx[:] = 2
    if cy == 1:
        ny[:] = 2
    blk = (ny[:, None] * nx[None, :] * 3 * 9).ravel().astype(np.uint64)  # doubles per node, plane order (x fastest)
    plane_len = int(blk.sum())
    within = np.concatenate([[0], np.cumsum(blk)[:-1]]).astype(np.uint64)
    local_nodes: List[np.ndarray] = []
    offsets: List[np.ndarray] = []
    for pidx in range(1, nranks):
        if pidx - 1 == rank or pidx == rank:
            zplane = pidx * per
            local_nodes.append(np.arange(plane, dtype=np.uint64) + np.uint64(plane * (zplane - z0)))
            offsets.append(within + np.uint64((pidx - 1) * plane_len))
    peers = []  # neighbour-exchange form: (peer rank, local ids of the shared plane in plane order - the same order on both sides)
    if rank > 0:
        peers.append((rank - 1, np.arange(plane, dtype=np.uint64) + np.uint64(plane * (k0 - z0))))
    if rank < nranks - 1:
        peers.append((rank + 1, np.arange(plane, dtype=np.uint64) + np.uint64(plane * (k1 - z0))))
    iface = {
        "peers": peers,
        "local_nodes": np.concatenate(local_nodes) if local_nodes else np.zeros(0, dtype=np.uint64),
        "packed_offsets": np.concatenate(offsets) if offsets else np.zeros(0, dtype=np.uint64),
        "packed_len": plane_len * max(nranks - 1, 0),
        "first_global_node": plane * z0,
    }
    return verts, conn, n_owned, iface


def general_partition(connectivity: np.ndarray, part_of_element: np.ndarray, num_nodes: int, nranks: int, sdim: int,
                      row_blocks_of_node: np.ndarray):
    """Generic element partition -> per-rank (owned element ids, ghost element ids, interface nodes, packed layout).

    row_blocks_of_node[g] = number of coupled nodes of global node g in the GLOBAL pattern (row length / sdim).
    Interface nodes are ordered by global id; every rank gets the same packed layout.
    """
    conn = np.asarray(connectivity, dtype=np.int64)
    part = np.asarray(part_of_element, dtype=np.int64)
    E, n = conn.shape
    touched = np.zeros((nranks, num_nodes), dtype=bool)
    for r in range(nranks):
        touched[r, np.unique(conn[part == r])] = True
    shared = touched.sum(axis=0) > 1
    iface_global = np.nonzero(shared)[0]
    sizes = (row_blocks_of_node[iface_global].astype(np.uint64) * np.uint64(sdim * sdim))
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) if len(sizes) else np.zeros(0, dtype=np.uint64)
    packed_len = int(sizes.sum())
    out = []
    for r in range(nranks):
        owned = np.nonzero(part == r)[0]
        mine = shared & touched[r]
        ghost_mask = (part != r) & mine[conn].any(axis=1)
        ghosts = np.nonzero(ghost_mask)[0]
        sel = np.isin(iface_global, np.nonzero(mine)[0])
        peers = {}  # peer rank -> global ids (ascending) of the nodes shared with it: the neighbour-exchange form
        for q in range(nranks):
            if q != r:
                both = np.nonzero(mine & touched[q])[0]
                if len(both):
                    peers[q] = both
        out.append({"owned": owned, "ghosts": ghosts, "iface_global": iface_global[sel], "packed_offsets": offs[sel], "packed_len": packed_len,
                    "peers": peers})
    return out


def localize(vertices: np.ndarray, connectivity: np.ndarray, owned: np.ndarray, ghosts: np.ndarray):
    """Build a rank-local mesh (owned elements first, then ghosts) with local node ids in ascending global order.
    Returns (local_vertices, local_connectivity_u64, global_ids_of_local_nodes)."""
    conn = np.asarray(connectivity, dtype=np.int64)
    elems = np.concatenate([owned, ghosts])
    sub = conn[elems]
    gids = np.unique(sub)
    lookup = np.full(int(conn.max()) + 1 if conn.size else 1, -1, dtype=np.int64)
    lookup[gids] = np.arange(len(gids))
    return np.ascontiguousarray(vertices[gids]), lookup[sub].astype(np.uint64), gids
