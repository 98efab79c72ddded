#!/usr/bin/env python3
"""bench.py - headline benchmark: elements/s assembled into a pre-built global CSR.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W   prints ONE JSON line.
  * --workload c3 (default; BASELINE.json config C3, the one the metric is quoted on): Hex8 3-D linear elasticity, unit cube,
    126^3 cells per GPU (2 000 376 elements, 6 145 149 dofs, nnz 489 959 451), canonical 2x2x2 Gauss rule, Lame from Young 1e6 /
    Poisson 0.2, u = 0, fp64.  N > 1 (torchrun, one rank per GPU): WEAK scaling - every rank owns a 126^3-cell z-slab of a
    126 x 126 x (126 N) box; the rows of the N-1 interface node planes are completed by the interface exchange:
      --exchange p2p (default)  fused into the tile kernel's flush: reductions straight into the neighbour's rows over NVLink (CUDA IPC
                                peer memory) + one neighbour barrier - no pack / send / add passes (fenris_b200/csrc/comm.cu)
      --exchange peers          packed rows, ncclSend / ncclRecv per neighbour, added on arrival
      --exchange allreduce      one world ncclAllReduce over a packed buffer of all interface rows
  * --workload c5 (BASELINE.json config C5): Tet4 linear elasticity, create_unit_box_uniform_tet_mesh_3d(161) = 50 079 372 elements,
    STRONG scaling: the mesh is cut into N z-slabs of element ranges (fenris_b200/partition.py); the same three exchanges (the fused one
    lives in the Tet4 chunk kernel's slot scatter).
  * a "step" = one assemble()-equivalent (values = all element contributions, interface exchange included), mesh / pattern / scatter
    lists resident in HBM and built outside the timed region, exactly like benches/assembly.rs:131-141 of the reference keeps the
    pattern outside the timed closure.
  * "parity": after the timed region every rank compares rows of its assembled matrix - the interface planes (N > 1) and a plane in
    the middle of its slab - ENTRYWISE with the C restatement of the reference CPU path on the sub-mesh around those rows (the oracle as
    checker, never as the thing measured).
  * --impl reference: the reference's CPU path (C restatement, OpenMP over colours = fenris-paradis semantics; the Rust
    original cannot be built in this image) timed on the host cores on the same workload when the host has the memory for it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CELLS = 126
C5_CELLS = 161
YOUNG, POISSON = 1e6, 0.2
MODE_NAMES = {"atomic": 0, "colored": 1, "gather": 2}
SAMPLE_CELLS = 64  # CPU baseline sample inside the GPU arm: 64^3 Hex8 elasticity cube (262 144 elements)


# ------------------------------------------------------------------------------------------------ helpers
def algorithmic_bytes(n, d, E, N, nnz):
    """SURVEY 8(d) / BASELINE.md 4: connectivity int32 once + coordinates f64 once + node-block map int32 once +
    one f64 read-modify-write per CSR value."""
    return 4 * n * E + 8 * d * N + 4 * n * n * E + 16 * nnz


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_memory_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs ~0.1-0.5 s to start: block until it delivers, so that samples exist for a short timed region."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = [r for (t, r) in self.rows if t_begin is None or (t_begin <= t <= t_end + 0.03)]
        if not rows and self.rows:  # region shorter than the sampling period: take the sample closest to it
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - (t_begin or 0)))[1]]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def best_thread_count(cr, et, op, w, p, data, v, c, ro, ci, vals, colors):
    """All the host threads the CPU path can USE: the coloured loop is memory/NUMA bound and gets slower when
    hyper-threads are oversubscribed, so try max, max/2, max/4 and keep the fastest."""
    mx = cr.max_threads()
    cands = sorted({mx, max(mx // 2, 1), max(mx // 4, 1), min(os.cpu_count() or mx, mx)}, reverse=True)
    best, best_t = mx, None
    for t in cands:
        vals[:] = 0
        t0 = time.perf_counter()
        cr.assemble(et, op, w, p, data, v, c, ro, ci, values=vals, colors=colors, nthreads=t)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = t, dt
    return best


def cpu_problem(workload, cells):
    """The CPU arm's problem: mesh, pattern, colours, rule (oracle side only)."""
    from oracle import cpu_ref as cr
    from oracle import fenris_oracle as fo
    mu, lam = fo.lame_from_young_poisson(YOUNG, POISSON)
    if workload == "c5":
        v, c = cr.gen_tet_mesh(cells)
        et = fo.TET4
        w, p = fo.tetrahedron_rule(1)
    else:
        v, c = cr.gen_hex_mesh(cells)
        et = fo.HEX8
        w, p = fo.hexahedron_gauss(2)
    ro, ci = cr.pattern(3, len(v), c)
    colors = cr.color_greedy(c, len(v))
    return cr, et, fo.LINEAR_ELASTIC, w, p, (mu, lam), v, c, ro, ci, colors


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full = C5_CELLS if args.workload == "c5" else CELLS
    # full size needs the CSR (values + u64 column indices) and the mesh on the host: ~9 GB for C3, ~22 GB for C5
    need_gb = 30.0 if args.workload == "c5" else 14.0
    small = 64 if args.workload != "c5" else 40
    n = full if (host_memory_gb() > need_gb and not args.reference_sample) else small
    cr, et, op, w, p, data, v, c, ro, ci, colors = cpu_problem(args.workload, n)
    vals = np.zeros(len(ci))
    cores = best_thread_count(cr, et, op, w, p, data, v, c, ro, ci, vals, colors)
    for _ in range(max(min(args.warmup, 3), 1)):
        vals[:] = 0
        cr.assemble(et, op, w, p, data, v, c, ro, ci, values=vals, colors=colors, nthreads=cores)
    times = []
    for _ in range(args.steps):
        vals[:] = 0
        t0 = time.perf_counter()
        cr.assemble(et, op, w, p, data, v, c, ro, ci, values=vals, colors=colors, nthreads=cores)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    value = len(c) / dt
    name = "Tet4" if args.workload == "c5" else "Hex8"
    sample = f"{name} elasticity {n}^3 cells ({len(c)} elements) per step, coloured OpenMP assembly on {cores} threads" + \
             (" = the full workload" if n == full else f" (sample of the {full}^3 workload: host memory)")
    out = {
        "impl": "reference", "metric": f"elements/sec into global CSR ({name} 3D linear elasticity, fp64)", "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, full), "reference_cells": n, "same_config": n == full,
                   "note": "C restatement of CsrParAssembler (fenris-paradis colouring); the Rust reference cannot be built here (no rustc/cargo)"},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_name(workload, cells):
    if workload == "c5":
        return f"C5: Tet4 linear elasticity, create_unit_box_uniform_tet_mesh_3d({cells}), one-point rule, Lame(E=1e6, nu=0.2), u=0"
    return f"C3: Hex8 linear elasticity (fenris-solid), unit cube {cells}^3 cells per GPU, Gauss 2^3, Lame(E=1e6, nu=0.2), u=0"


# ------------------------------------------------------------------------------------------------ parity (oracle = checker)
def parity_c3(ctx, vals, cells, h, rank, world, data):
    """Rows of whole node planes of this rank's matrix against the C oracle on the two cell layers around each plane: the interface
    planes (completed by the exchange) and the middle plane of the slab.  K_e is translation invariant, so the sub-mesh may start at z = 0."""
    from oracle import cpu_ref as cr
    from oracle import fenris_oracle as fo
    vx = cells + 1
    plane = vx * vx
    g0 = 1 if rank > 0 else 0                      # ghost cell layers below
    nplanes = cells + 1 + g0 + (1 if rank < world - 1 else 0)
    cnt1 = np.full(vx, 3)
    cnt1[[0, -1]] = 2
    cz = np.full(nplanes, 3)
    cz[[0, -1]] = 2
    blocks = (cz[:, None, None] * cnt1[None, :, None] * cnt1[None, None, :]).reshape(-1)  # coupled nodes per node, plane-major
    row_start = np.concatenate([[0], np.cumsum(np.repeat(3 * blocks, 3))]).astype(np.int64)
    assert int(row_start[-1]) == len(vals), "row layout of the structured slab"
    sv, sc = cr.gen_hex_mesh(cells, cz=2, cell_size=h)
    sro, sci = cr.pattern(3, len(sv), sc)
    w, p = fo.hexahedron_gauss(2)
    ref = cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, data, sv, sc, sro, sci)
    rb, re = int(sro[3 * plane]), int(sro[3 * 2 * plane])
    planes = [g0 + cells // 2]
    if rank > 0:
        planes.append(g0)
    if rank < world - 1:
        planes.append(g0 + cells)
    worst, rows = 0.0, 0
    for kp in planes:
        b, e = int(row_start[3 * plane * kp]), int(row_start[3 * plane * (kp + 1)])
        assert e - b == re - rb
        worst = max(worst, float(np.linalg.norm(vals[b:e] - ref[rb:re]) / np.linalg.norm(ref[rb:re])))
        rows += 3 * plane
    return worst, rows


def parity_sampled_rows(ctx, vals, et, gverts, gconn, part, data, rule, seed):
    """Unstructured form (C5): sample nodes of this rank's owned elements (half of them on the partition interface when there is one),
    gather every global element that touches them, assemble that sub-mesh with the C oracle and compare the sampled rows entrywise."""
    from oracle import cpu_ref as cr
    from oracle import fenris_oracle as fo
    rng = np.random.default_rng(seed)
    gids = part["global_nodes"].astype(np.int64)
    owned_nodes = np.unique(part["connectivity"][:part["num_owned"]].astype(np.int64))
    take = [rng.choice(owned_nodes, size=min(4000, len(owned_nodes)), replace=False)]
    for _, ids in part["peers"]:
        ids = ids.astype(np.int64)
        take.append(rng.choice(ids, size=min(2000, len(ids)), replace=False))
    sample_local = np.unique(np.concatenate(take))
    sample_global = gids[sample_local]
    mark = np.zeros(len(gverts), dtype=bool)
    mark[sample_global] = True
    hit = []
    chunk = 1 << 22
    for a in range(0, len(gconn), chunk):
        hit.append(np.nonzero(mark[gconn[a:a + chunk].astype(np.int64)].any(axis=1))[0] + a)
    elems = np.concatenate(hit)
    sub = gconn[elems].astype(np.int64)
    sg = np.unique(sub)
    sconn = np.searchsorted(sg, sub).astype(np.uint64)
    sverts = np.ascontiguousarray(gverts[sg])
    sro, sci = cr.pattern(3, len(sverts), sconn)
    ref = cr.assemble(et, fo.LINEAR_ELASTIC, rule[0], rule[1], data, sverts, sconn, sro, sci)
    ro = ctx.row_offsets_download().astype(np.int64)
    spos = np.searchsorted(sg, sample_global)
    num = den = 0.0
    for l, sp in zip(sample_local.tolist(), spos.tolist()):
        b, e = int(ro[3 * l]), int(ro[3 * l + 3])
        rb, re = int(sro[3 * sp]), int(sro[3 * sp + 3])
        assert e - b == re - rb, "sampled row has a different length in the sub-mesh"
        d = vals[b:e] - ref[rb:re]
        num += float(d @ d)
        den += float(ref[rb:re] @ ref[rb:re])
    return float(np.sqrt(num / den)), 3 * len(sample_local)


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"])
    ap.add_argument("--scatter", default=os.environ.get("FB200_SCATTER", "atomic"), choices=list(MODE_NAMES))
    ap.add_argument("--cells", type=int, default=0, help="cells per edge (c3: per GPU, default 126; c5: of the whole box, default 161)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "peers", "allreduce"],
                    help="N > 1: interface rows fused into the kernel over peer memory (default), neighbour ncclSend/ncclRecv, or one world ncclAllReduce")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--reference-sample", action="store_true", help="--impl reference: time the small sample even if the host could hold the full workload")
    ap.add_argument("--all-modes", action="store_true", help="also time the other scatter modes (extra JSON field)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import fenris_b200 as fb
    from fenris_b200 import partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch with torchrun --nproc-per-node N for N > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    c5 = args.workload == "c5"
    cells = args.cells or (C5_CELLS if c5 else CELLS)
    h = 1.0 / cells
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(YOUNG, POISSON))
    et = fb.TET4 if c5 else fb.HEX8
    n_el_nodes = 4 if c5 else 8
    w, p = fb.canonical_stiffness_quadrature(et)
    data = (lame.mu, lame.lambda_)
    mode = MODE_NAMES[args.scatter]

    ctx = fb.Context(local_rank)
    t_setup = time.perf_counter()
    part = gverts = gconn = iface = None
    if c5:
        gmesh = fb.create_unit_box_uniform_tet_mesh_3d(cells)
        gverts, gconn = gmesh.vertices(), gmesh.connectivity()
        layer_starts = partition.tet_box_layer_starts(cells, cells, cells)
        starts = layer_starts[partition.split_layers(cells, world)]
        if world == 1:
            part = {"vertices": gverts, "connectivity": gconn, "num_owned": len(gconn), "global_nodes": np.arange(len(gverts)), "peers": []}
        else:
            part = partition.element_range_partition(gverts, gconn, starts, rank)
        verts, conn, n_owned = part["vertices"], part["connectivity"], part["num_owned"]
        peers = part["peers"]
    elif world == 1:
        mesh = fb.create_rectangular_uniform_hex_mesh(1.0, 1, 1, 1, cells)
        verts, conn, n_owned, peers = mesh.vertices(), mesh.connectivity(), mesh.num_elements(), []
    else:
        verts, conn, n_owned, iface = partition.structured_hex_slab(cells, cells, cells * world, h, rank, world)
        peers = iface["peers"]
    t_mesh = time.perf_counter() - t_setup
    ctx.space_upload(et, verts, conn)
    ctx.set_num_owned_elements(n_owned)
    nrows, nnz = ctx.assemble_pattern(3)
    if mode == MODE_NAMES["colored"] or args.all_modes:
        ctx.color_nodes()
    exchange = args.exchange
    if world > 1:
        uid = [fb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        if exchange == "allreduce" and not c5:
            ctx.interface_set(iface["local_nodes"], iface["packed_offsets"], iface["packed_len"])
        else:
            ctx.interface_set_peers(peers)
            if exchange == "allreduce":
                exchange = "peers"
            if exchange == "p2p" and (mode != MODE_NAMES["atomic"] or not ctx.interface_enable_p2p()):
                exchange = "peers"  # the fused exchange lives in the Hex8 tile kernel and the Tet4 chunk kernel (ATOMIC scatter)
    # the first assembly builds the scatter lists of the kernel (tile / chunk lists): part of the set-up, like the pattern
    ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, data, scatter_mode=mode, accumulate=False)
    if world > 1:
        ctx.interface_allreduce()
    ctx.synchronize()
    setup_s = time.perf_counter() - t_setup

    def step(m=mode):
        ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, data, scatter_mode=m, accumulate=False)
        if world > 1:
            ctx.interface_allreduce()

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(nsteps, fn):
        barrier()
        ctx.timer_begin()
        for _ in range(nsteps):
            fn()
        ms = ctx.timer_end()
        ctx.synchronize()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    t_begin = sampler.mark()
    ms_total = timed(args.steps, step)
    t_end = sampler.mark()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    launches = ctx.launch_count - launches0
    ms_per_step = ms_total / args.steps
    total_owned = n_owned
    if world > 1:
        t = torch.tensor([n_owned], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total_owned = int(t.item())
    value = total_owned / (ms_per_step * 1e-3)

    # ---- parity of what the timed steps left in HBM (every rank; max over ranks)
    parity = None
    vals_host = None
    if not args.no_parity or not args.no_e2e:
        vals_host = torch.empty(max(nnz, 1), dtype=torch.float64, pin_memory=True).numpy()
    if not args.no_parity:
        ctx.values_download(vals_host)
        if c5:
            from oracle import fenris_oracle as fo
            err, rows = parity_sampled_rows(ctx, vals_host[:nnz], fo.TET4, gverts, gconn, part, data, (w, p), 17 + rank)
        else:
            err, rows = parity_c3(ctx, vals_host[:nnz], cells, h, rank, world, data)
        if world > 1:
            t = torch.tensor([err], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r = torch.tensor([rows], dtype=torch.int64, device="cuda")
            dist.all_reduce(r)
            err, rows = float(t.item()), int(r.item())
        parity = {"rel_frobenius": err, "rows_checked": rows, "tolerance": 1e-12, "ok": bool(err < 1e-12),
                  "against": "oracle/cpu_ref.c on the sub-mesh around the checked rows" +
                             ("; interface planes + the middle plane of every rank's slab" if not c5 else "; sampled owned + interface nodes of every rank")}

    # ---- roofline of the dominant kernel (this rank's launch; algorithmic bytes of SURVEY 8d): the step's assembly launch timed alone
    E_loc, N_loc = n_owned, verts.shape[0]
    b_algo = algorithmic_bytes(n_el_nodes, 3, E_loc, N_loc, nnz)
    kernel_ms = timed(args.steps, lambda: ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, data, scatter_mode=mode, accumulate=False)) / args.steps
    if world > 1:
        ctx.interface_allreduce()
    peak, peak_src = measured_peak_gbs()
    achieved = b_algo / (kernel_ms * 1e-3) / 1e9
    kernel_name = {"atomic": ("assemble_tet4_chunk_kernel<LINEAR_ELASTIC>" if c5 else
                              "assemble_hex8_tile_kernel<LINEAR_ELASTIC> (owner stores: the overwriting launch of the step, no zero-fill pass)"
                              if os.environ.get("FB200_HEX8_TILE", "64") != "0" else "assemble_hex8_mma_kernel<LINEAR_ELASTIC, ATOMIC>"),
                   "colored": "element kernel, COLORED, one launch per colour",
                   "gather": "assemble_gather_kernel"}[args.scatter]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{args.workload}:{args.scatter}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel_name, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": b_algo,
                "bytes_per_element": b_algo / max(E_loc, 1), "peak_source": peak_src,
                "frac_of_step": (b_algo / (ms_per_step * 1e-3) / 1e9) / peak}

    other = None
    if args.all_modes:
        other = {}
        for name, m in MODE_NAMES.items():
            for _ in range(2):
                step(m)
            other[name] = total_owned / (timed(max(args.steps // 2, 3), lambda m=m: step(m)) / max(args.steps // 2, 3) * 1e-3)

    # ---- end to end through the host-buffer path: H2D of the vertex coordinates + D2H of the CSR values every step
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 5))
        verts_host = torch.from_numpy(np.ascontiguousarray(verts)).pin_memory().numpy()

        def e2e_step():
            ctx.space_update_vertices(verts_host)
            if world == 1:
                ctx.assemble_into_csr(fb.LINEAR_ELASTIC, w, p, data, vals_host, scatter_mode=mode, accumulate=False)
            else:  # the host must receive the COMPLETED rows: exchange before the download
                ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, data, scatter_mode=mode, accumulate=False)
                ctx.interface_allreduce()
                ctx.values_download(vals_host)

        e2e_step()
        ms = timed(e2e_steps, e2e_step) / e2e_steps
        e2e = {"value": total_owned / (ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": int(verts_host.nbytes), "d2h_bytes_per_step": int(nnz * 8),
               "ms_per_step": ms, "steps": e2e_steps}
    del vals_host

    # ---- the pipeline in which the matrix never leaves the device (N = 1, C3): assemble -> homogeneous Dirichlet rows on the resident CSR
    # (global.rs:379-451) -> Jacobi-PCG on the device (fenris-sparse/src/cg.rs:364-480) -> only the solution crosses PCIe.  Informational:
    # the PCIe-bound `e2e` above ships 3.9 GB of values per step; a solver that runs where the matrix is assembled ships 49 MB.
    e2e_solve = None
    if world == 1 and not c5 and not args.no_e2e and mode == MODE_NAMES["atomic"]:
        try:
            vx = cells + 1
            fixed = np.arange(vx * vx, dtype=np.uint64)          # the node plane z = 0 is clamped
            b = np.zeros(3 * verts.shape[0])
            b[2::3] = -1.0 / verts.shape[0]                       # a uniform downward nodal load
            b[: 3 * len(fixed)] = 0.0
            iters = 60
            t0 = time.perf_counter()
            ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, data, scatter_mode=mode, accumulate=False)
            ctx.apply_homogeneous_dirichlet_bc_csr(fixed)
            ctx.synchronize()
            t1 = time.perf_counter()
            res = None
            try:
                _, its, res = ctx.cg_solve(b, rel_tol=1e-30, max_iter=iters)  # a fixed number of iterations: time per iteration, not convergence
            except fb.Fb200Error as exc:
                if exc.status != fb.ERR_NOT_CONVERGED:
                    raise
                its = iters
            t2 = time.perf_counter()
            e2e_solve = {"assemble_dirichlet_ms": (t1 - t0) * 1e3, "pcg_iterations": int(its), "pcg_ms": (t2 - t1) * 1e3,
                         "pcg_ms_per_iteration": (t2 - t1) * 1e3 / max(int(its), 1), "h2d_bytes": int(b.nbytes), "d2h_bytes": int(b.nbytes),
                         "note": "wall clock incl. the H2D of b and the D2H of x; the 3.9 GB of CSR values stay in HBM"}
        except Exception as exc:  # informational only: never break the bench line
            e2e_solve = {"error": str(exc)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = 40 if c5 else SAMPLE_CELLS
        cr, oet, oop, sw, sp_, sdata, sv, sc, sro, sci, colors = cpu_problem(args.workload, n)
        vals = np.zeros(len(sci))
        cores = best_thread_count(cr, oet, oop, sw, sp_, sdata, sv, sc, sro, sci, vals, colors)
        cr.assemble(oet, oop, sw, sp_, sdata, sv, sc, sro, sci, values=vals, colors=colors, nthreads=cores)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 200):
            vals[:] = 0
            cr.assemble(oet, oop, sw, sp_, sdata, sv, sc, sro, sci, values=vals, colors=colors, nthreads=cores)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": len(sc) / dt, "unit": "elements/s", "cores": cores, "kind": "port",
               "sample": f"{'Tet4' if c5 else 'Hex8'} elasticity {n}^3 cells ({len(sc)} elements) x {reps} reps, coloured OpenMP C restatement of CsrParAssembler"}

    if rank == 0:
        exch = {"p2p": "fused into the assembly kernel's scatter over NVLink peer memory (red.global.add.f64 on the neighbour's rows) + neighbour barrier",
                "peers": "neighbour ncclSend/ncclRecv of the packed rows, summed on arrival", "allreduce": "world ncclAllReduce"}[exchange]
        out = {
            "metric": f"elements/sec into global CSR ({'Tet4' if c5 else 'Hex8'} 3D linear elasticity, fp64)", "value": value, "unit": "elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if c5 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, cells),
                       "elements_total": int(total_owned), "elements_rank0": int(n_owned), "nnz_rank0": int(nnz), "scatter": args.scatter,
                       "parallelism": "1 GPU" if world == 1 else f"z-slab element partition x{world}, interface rows: {exch}",
                       "exchange": exchange if world > 1 else None,
                       "l2": "inputs+outputs >> L2 (values 3.9 - 9 GB per GPU); no L2 flush needed", "setup_s": setup_s, "mesh_s": t_mesh},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        }
        if e2e_solve:
            out["e2e_solve"] = e2e_solve
        if other:
            out["other_modes_elements_per_s"] = other
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
