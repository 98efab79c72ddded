"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): z-slab element partition with ghost layers, one
process per GPU, interface rows summed by fb200_interface_allreduce (ncclAllReduce over NVLink).  Each rank compares the
complete rows it holds with a single-process C-oracle assembly of the whole mesh."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _num_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, uid, cx, cy, cz, scatter_mode, exchange, q):
    try:
        import fenris_b200 as fb
        from fenris_b200 import partition
        from oracle import cpu_ref as cr
        from oracle import fenris_oracle as fo
        mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
        h = 1.0 / cx
        verts, conn, n_owned, iface = partition.structured_hex_slab(cx, cy, cz, h, rank, world)
        w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
        ctx = fb.Context(rank)
        ctx.space_upload(fb.HEX8, verts, conn)
        ctx.set_num_owned_elements(n_owned)
        ctx.assemble_pattern(3)
        ctx.color_nodes()
        ctx.comm_init(uid, rank, world)
        if exchange in ("peers", "p2p", "p2p_zero"):
            ctx.interface_set_peers(iface["peers"])
            if exchange != "peers":  # the exchange fused into the tile kernel's flush over peer-mapped memory (comm.cu)
                assert ctx.interface_enable_p2p(), "fused p2p exchange refused on a slab partition"
            if exchange == "p2p_zero":  # ... with zero-fill + reductions instead of owner stores (the other list flavour)
                ctx.set_tuning("hex8_owner_stores", 0)
        else:
            ctx.interface_set(iface["local_nodes"], iface["packed_offsets"], iface["packed_len"])
        ctx.values_upload(np.full(ctx.nnz, 1e300))  # overwrite semantics: nothing of this may survive
        for _ in range(3):  # several times: the exchange must be repeatable
            ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, (mu, lam), scatter_mode=scatter_mode, accumulate=False)
            ctx.interface_allreduce()
        ctx.synchronize()
        ro, ci = ctx.pattern_download()
        vals = ctx.values_download().copy()
        ctx.close()
        # global reference
        vg, cg = cr.gen_hex_mesh(cx, cz=cz, cell_size=h)
        gro, gci = cr.pattern(3, len(vg), cg)
        ref = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (mu, lam), vg, cg, gro, gci)
        per = cz // world
        plane = (cx + 1) * (cy + 1)
        first = iface["first_global_node"]
        g_lo, g_hi = plane * rank * per, plane * ((rank + 1) * per + 1)
        num = den = 0.0
        for g in range(g_lo, g_hi):
            l = g - first
            for i in range(3):
                b, e = int(ro[3 * l + i]), int(ro[3 * l + i + 1])
                gb, ge = int(gro[3 * g + i]), int(gro[3 * g + i + 1])
                assert e - b == ge - gb, "interface row layout differs from the global pattern"
                assert np.array_equal(ci[b:e].astype(np.int64) + 3 * first, gci[gb:ge].astype(np.int64))
                num += float(np.sum((vals[b:e] - ref[gb:ge]) ** 2))
                den += float(np.sum(ref[gb:ge] ** 2))
        q.put((rank, float(np.sqrt(num / den)), None))
    except Exception as exc:  # noqa
        import traceback
        q.put((rank, None, traceback.format_exc()))


def _run(world, target, args):
    import torch.multiprocessing as mp

    import fenris_b200 as fb
    uid = fb.Context.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, uid) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = [q.get(timeout=300) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    for rank, err, tb in results:
        assert tb is None, tb
        assert err < 1e-12, (rank, err)


@pytest.mark.skipif(_num_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("scatter_mode,exchange", [(0, "p2p"), (0, "p2p_zero"), (0, "peers"), (0, "allreduce"), (2, "peers"), (1, "p2p")])
def test_slab_partition_nccl_equals_global(scatter_mode, exchange):
    _run(2, _worker, (12, 12, 16, scatter_mode, exchange))


@pytest.mark.skipif(_num_gpus() < 4, reason="needs at least four GPUs")
@pytest.mark.parametrize("exchange", ["p2p", "peers"])
def test_slab_partition_four_ranks(exchange):
    _run(4, _worker, (10, 10, 16, 0, exchange))  # inner ranks have two neighbours (the reference generator wants cx == cy)


def _range_worker(rank, world, uid, kind, exchange, q):
    """Tet4 (config C5's kernel: the chunk kernel behind a partition with ghost elements) and Hex27 (the DMMA kernel) slabs from
    partition.element_range_partition, neighbour exchange of the packed interface rows."""
    try:
        import fenris_b200 as fb
        from fenris_b200 import partition
        from oracle import cpu_ref as cr
        from oracle import fenris_oracle as fo
        mu, lam = fo.lame_from_young_poisson(1e6, 0.2)
        if kind == "tet4":
            n = 10
            m = fb.create_unit_box_uniform_tet_mesh_3d(n)
            layer_starts = partition.tet_box_layer_starts(n, n, n)
            et, oet = fb.TET4, fo.TET4
        else:
            n = 6
            m = fb.hex27_mesh_from(fb.create_unit_box_uniform_hex_mesh_3d(n))
            layer_starts = np.arange(n + 1, dtype=np.int64) * n * n
            et, oet = fb.HEX27, fo.HEX27
        vg, cg = m.vertices(), m.connectivity()
        vg = fo.jitter_vertices(vg, 1.0 / n, amp=0.1)
        starts = layer_starts[partition.split_layers(len(layer_starts) - 1, world)]
        part = partition.element_range_partition(vg, cg, starts, rank)
        w, p = fb.canonical_stiffness_quadrature(et)
        ctx = fb.Context(rank)
        ctx.space_upload(et, part["vertices"], part["connectivity"])
        ctx.set_num_owned_elements(part["num_owned"])
        ctx.assemble_pattern(3)
        ctx.comm_init(uid, rank, world)
        ctx.interface_set_peers(part["peers"])
        if exchange == "p2p":
            assert ctx.interface_enable_p2p(), "fused p2p exchange refused on a slab partition"
        ctx.values_upload(np.full(ctx.nnz, 1e300))
        for _ in range(3):
            ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, (mu, lam), scatter_mode=fb.SCATTER_ATOMIC, accumulate=False)
            ctx.interface_allreduce()
        ctx.synchronize()
        ro, ci = ctx.pattern_download()
        vals = ctx.values_download().copy()
        ctx.close()
        gro, gci = cr.pattern(3, len(vg), cg)
        ref = cr.assemble(oet, fo.LINEAR_ELASTIC, w, p, (mu, lam), vg, cg, gro, gci)
        gids = part["global_nodes"].astype(np.int64)
        num = den = 0.0
        owned_nodes = np.unique(part["connectivity"][:part["num_owned"]].astype(np.int64))
        for l in owned_nodes.tolist():
            g = int(gids[l])
            for i in range(3):
                b, e = int(ro[3 * l + i]), int(ro[3 * l + i + 1])
                gb, ge = int(gro[3 * g + i]), int(gro[3 * g + i + 1])
                assert e - b == ge - gb, "row layout differs from the global pattern"
                cols = ci[b:e].astype(np.int64)
                assert np.array_equal(3 * gids[cols // 3] + cols % 3, gci[gb:ge].astype(np.int64))
                num += float(np.sum((vals[b:e] - ref[gb:ge]) ** 2))
                den += float(np.sum(ref[gb:ge] ** 2))
        q.put((rank, float(np.sqrt(num / den)), None))
    except Exception:  # noqa
        import traceback
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.skipif(_num_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("kind,exchange", [("tet4", "p2p"), ("tet4", "peers"), ("hex27", "peers"), ("hex27", "p2p")])
def test_element_range_partition_equals_global(kind, exchange):
    # ("hex27", "p2p"): the Hex27 kernel has no fused exchange - with p2p enabled the packed neighbour exchange must still be used
    _run(2, _range_worker, (kind, exchange))
