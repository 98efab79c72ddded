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


def element_range_partition(vertices: np.ndarray, connectivity: np.ndarray, starts, rank: int):
    """Partition by CONTIGUOUS ELEMENT RANGES of any uniform mesh: rank r owns elements [starts[r], starts[r+1]).

    The reference's generators emit cells layer by layer (procedural.rs:216-403), so ranges of element ids are z-slabs of the
    structured hex and BCC tet meshes (config C5: 50 M tets over 8 ranks); for any other mesh the ranges are whatever the caller's
    ordering makes them.  Returns a dict for `rank`:
      vertices, connectivity  rank-local mesh: owned elements first, then the GHOST elements = elements of other ranks that touch a node
                              of an owned element (pattern only: they make the rows of interface nodes identical on all sharing ranks)
      num_owned               owned elements
      global_nodes            global id of every local node (ascending, so local order = global order)
      peers                   [(peer rank, local ids of the nodes shared with it, ascending global id - the same order on both sides)]
    Only O(own slab) memory beyond the global arrays passed in."""
    conn = np.asarray(connectivity)
    starts = np.asarray(starts, dtype=np.int64)
    nranks = len(starts) - 1
    e0, e1 = int(starts[rank]), int(starts[rank + 1])
    num_nodes = len(vertices)
    mine = np.zeros(num_nodes, dtype=bool)
    mine[conn[e0:e1].ravel().astype(np.int64)] = True
    # elements of other ranks that touch one of my nodes (chunked: the gather mine[conn] is the only O(E) temporary)
    ghost_ids = []
    chunk = 1 << 22
    for a in range(0, len(conn), chunk):
        b = min(a + chunk, len(conn))
        hit = mine[conn[a:b].astype(np.int64)].any(axis=1)
        lo, hi = max(a, e0), min(b, e1)
        if lo < hi:
            hit[lo - a:hi - a] = False
        ghost_ids.append(np.nonzero(hit)[0] + a)
    ghosts = np.concatenate(ghost_ids) if ghost_ids else np.zeros(0, dtype=np.int64)
    owner_of_ghost = np.searchsorted(starts, ghosts, side="right") - 1
    owned = np.arange(e0, e1, dtype=np.int64)
    lverts, lconn, gids = localize(vertices, conn, owned, ghosts)
    peers = []
    for q in np.unique(owner_of_ghost):
        assert 0 <= q < nranks and q != rank
        touched_by_q = np.unique(conn[ghosts[owner_of_ghost == q]].ravel().astype(np.int64))
        shared = touched_by_q[mine[touched_by_q]]  # ascending global ids: nodes of my owned elements that q's elements touch too
        peers.append((int(q), np.searchsorted(gids, shared).astype(np.uint64)))
    return {"vertices": lverts, "connectivity": lconn, "num_owned": e1 - e0, "global_nodes": gids, "peers": peers}


def tet_box_layer_starts(cx: int, cy: int, cz: int) -> np.ndarray:
    """First element id of every z-layer of cells of the reference's BCC tet box mesh (create_rectangular_uniform_tet_mesh,
    procedural.rs:286-403), as fb200_gen_tet_mesh / the oracle emit it: cells in k-major order, a cell emits 4 tets per face shared with
    its +axis neighbour and 2 per face on the box boundary.  Returns cz + 1 offsets."""
    def per_axis(n):
        c = np.full(n, 4, dtype=np.int64)   # interior +face octahedron
        c[-1] = 0
        c[0] += 2                           # low boundary pyramid
        c[-1] += 2                          # high boundary pyramid
        return c
    ax, ay, az = per_axis(cx), per_axis(cy), per_axis(cz)
    per_layer = cy * ax.sum() + cx * ay.sum() + cx * cy * az  # [cz]
    return np.concatenate([[0], np.cumsum(per_layer)]).astype(np.int64)


def split_layers(num_layers: int, nranks: int) -> np.ndarray:
    """Layer boundaries of nranks slabs, as even as possible (the first num_layers % nranks slabs get one more)."""
    base, extra = divmod(num_layers, nranks)
    sizes = np.array([base + (1 if r < extra else 0) for r in range(nranks)], dtype=np.int64)
    return np.concatenate([[0], np.cumsum(sizes)])
