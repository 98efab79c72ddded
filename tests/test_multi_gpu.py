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
        if exchange == "peers":
            ctx.interface_set_peers(iface["peers"])
        else:
            ctx.interface_set(iface["local_nodes"], iface["packed_offsets"], iface["packed_len"])
        for _ in range(2):  # twice: the exchange must be repeatable (overwrite semantics)
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


@pytest.mark.skipif(_num_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("scatter_mode,exchange", [(0, "peers"), (0, "allreduce"), (2, "peers")])
def test_slab_partition_nccl_equals_global(scatter_mode, exchange):
    import torch.multiprocessing as mp

    import fenris_b200 as fb
    world = 2
    cx = cy = 12
    cz = 16
    uid = fb.Context.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, cx, cy, cz, scatter_mode, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, tb in results:
        assert tb is None, tb
        assert err < 1e-12, (rank, err)
