#!/usr/bin/env python3
"""bench.py - headline benchmark: Hex8 3-D linear-elasticity elements/s assembled into a pre-built global CSR.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W   prints ONE JSON line.
  * N = 1 workload = BASELINE.json config C3: unit cube, 126^3 Hex8 cells (2 000 376 elements, 6 145 149 dofs,
    nnz 489 959 451), canonical 2x2x2 Gauss rule, Lame from Young 1e6 / Poisson 0.2, u = 0, fp64.
  * N > 1 (torchrun, one rank per GPU): weak scaling - every rank owns a 126 x 126 x 126-cell z-slab of a
    126 x 126 x (126 N) box; rows of the N-1 interface node planes are summed with ncclAllReduce (packed, interface rows only).
  * a "step" = one assemble_into_csr-equivalent (values zeroed/overwritten + all element contributions + interface exchange),
    mesh / pattern / scatter map / colours resident in HBM and built outside the timed region, exactly like
    benches/assembly.rs:131-141 of the reference keeps the pattern outside the timed closure.
  * --impl reference: the reference's CPU path (C restatement, OpenMP over colours = fenris-paradis semantics; the Rust
    original cannot be built in this image) timed on the host cores on a bounded sample of the same workload.
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
YOUNG, POISSON = 1e6, 0.2
MODE_NAMES = {"atomic": 0, "colored": 1, "gather": 2}
SAMPLE_CELLS = 64  # CPU baseline sample: 64^3 Hex8 elasticity cube (262 144 elements)


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


def slab_local_mesh(cells_xy: int, cells_z_per_rank: int, rank: int, nranks: int, h: float):
    """Element partition by z-slabs with one ghost cell layer per neighbour (see fenris_b200/partition.py)."""
    from fenris_b200 import partition
    return partition.structured_hex_slab(cells_xy, cells_xy, cells_z_per_rank * nranks, h, rank, nranks)


def best_thread_count(cr, fo, w, p, data, v, c, ro, ci, vals, colors):
    """All the host threads the CPU path can USE: the coloured loop is memory/NUMA bound and gets slower when
    hyper-threads are oversubscribed, so try max, max/2, max/4 and keep the fastest."""
    mx = cr.max_threads()
    cands = sorted({mx, max(mx // 2, 1), max(mx // 4, 1), min(os.cpu_count() or mx, mx)}, reverse=True)
    best, best_t = mx, None
    for t in cands:
        dt = None
        for _ in range(2):
            vals[:] = 0
            t0 = time.perf_counter()
            cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, data, v, c, ro, ci, values=vals, colors=colors, nthreads=t)
            d = time.perf_counter() - t0
            dt = d if dt is None else min(dt, d)
        if best_t is None or dt < best_t:
            best, best_t = t, dt
    return best


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    from oracle import cpu_ref as cr
    from oracle import fenris_oracle as fo
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = SAMPLE_CELLS
    v, c = cr.gen_hex_mesh(n)
    ro, ci = cr.pattern(3, len(v), c)
    colors = cr.color_greedy(c, len(v))
    w, p = fo.hexahedron_gauss(2)
    mu, lam = fo.lame_from_young_poisson(YOUNG, POISSON)
    vals = np.zeros(len(ci))
    cores = best_thread_count(cr, fo, w, p, (mu, lam), v, c, ro, ci, vals, colors)
    for _ in range(max(args.warmup, 1)):
        vals[:] = 0
        cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (mu, lam), v, c, ro, ci, values=vals, colors=colors, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals[:] = 0
        cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, w, p, (mu, lam), v, c, ro, ci, values=vals, colors=colors, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    value = len(c) / dt
    sample = f"Hex8 elasticity {n}^3 cube ({len(c)} elements) per step, coloured OpenMP assembly on {cores} threads"
    out = {
        "impl": "reference", "metric": "elements/sec into global CSR (Hex8 3D linear elasticity, fp64)", "value": value, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: Hex8 linear elasticity, unit cube 126^3 (reference arm timed on a 64^3 sample of the same problem)",
                   "note": "C restatement of CsrParAssembler (fenris-paradis colouring); the Rust reference cannot be built here (no rustc/cargo)"},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scatter", default=os.environ.get("FB200_SCATTER", "atomic"), choices=list(MODE_NAMES))
    ap.add_argument("--cells", type=int, default=CELLS, help="cells per edge per GPU (126 = BASELINE config C3)")
    ap.add_argument("--exchange", default="peers", choices=["peers", "allreduce"],
                    help="N > 1: interface rows by neighbour ncclSend/ncclRecv (default) or one world ncclAllReduce")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--all-modes", action="store_true", help="also time the other scatter modes (extra JSON field)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import fenris_b200 as fb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch with torchrun --nproc-per-node N for N > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cells = args.cells
    h = 1.0 / cells
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(YOUNG, POISSON))
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)
    data = (lame.mu, lame.lambda_)
    mode = MODE_NAMES[args.scatter]

    ctx = fb.Context(local_rank)
    t_setup = time.perf_counter()
    if world == 1:
        mesh = fb.create_rectangular_uniform_hex_mesh(1.0, 1, 1, 1, cells)
        verts, conn, n_owned, iface = mesh.vertices(), mesh.connectivity(), mesh.num_elements(), None
    else:
        verts, conn, n_owned, iface = slab_local_mesh(cells, cells, rank, world, h)
    ctx.space_upload(fb.HEX8, verts, conn)
    ctx.set_num_owned_elements(n_owned)
    nrows, nnz = ctx.assemble_pattern(3)
    ctx.color_nodes()
    if world > 1:
        uid = [fb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        if args.exchange == "peers":
            ctx.interface_set_peers(iface["peers"])
        else:
            ctx.interface_set(iface["local_nodes"], iface["packed_offsets"], iface["packed_len"])
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
    total_owned = n_owned * world if world == 1 else None
    if world > 1:
        t = torch.tensor([n_owned], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total_owned = int(t.item())
    value = total_owned / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launch; algorithmic bytes of SURVEY 8d)
    E_loc, N_loc = n_owned, verts.shape[0]
    b_algo = algorithmic_bytes(8, 3, E_loc, N_loc, nnz)
    kernel_ms = timed(args.steps, lambda: ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, data, scatter_mode=mode,
                                                                      accumulate=(mode != MODE_NAMES["gather"]))) / args.steps
    peak, peak_src = measured_peak_gbs()
    achieved = b_algo / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.scatter)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": {"atomic": "assemble_hex8_tile_kernel<LINEAR_ELASTIC> (accumulate launch: every CSR value read-modify-written)"
                           if os.environ.get("FB200_HEX8_TILE", "64") != "0" else "assemble_hex8_mma_kernel<LINEAR_ELASTIC, ATOMIC>",
                           "colored": "assemble_hex8_mma_kernel<LINEAR_ELASTIC, COLORED> x colours",
                           "gather": "assemble_gather_kernel"}[args.scatter],
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": b_algo, "bytes_per_element": b_algo / max(E_loc, 1), "peak_source": peak_src}

    other = None
    if args.all_modes:
        other = {}
        for name, m in MODE_NAMES.items():
            for _ in range(2):
                step(m)
            other[name] = total_owned / (timed(max(args.steps // 2, 3), lambda m=m: step(m)) / max(args.steps // 2, 3) * 1e-3)

    # ---- end to end through the host-buffer C-ABI call: H2D of the vertex coordinates + D2H of the CSR values every step
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 5))
        vals_host = torch.empty(max(nnz, 1), dtype=torch.float64, pin_memory=True).numpy()
        verts_host = torch.from_numpy(np.ascontiguousarray(verts)).pin_memory().numpy()

        def e2e_step():
            ctx.space_update_vertices(verts_host)
            ctx.assemble_into_csr(fb.LINEAR_ELASTIC, w, p, data, vals_host, scatter_mode=mode, accumulate=False)
            if world > 1:
                ctx.interface_allreduce()

        e2e_step()
        ms = timed(e2e_steps, e2e_step) / e2e_steps
        e2e = {"value": total_owned / (ms * 1e-3), "unit": "elements/s", "h2d_bytes_per_step": int(verts_host.nbytes), "d2h_bytes_per_step": int(nnz * 8),
               "ms_per_step": ms, "steps": e2e_steps}
        del vals_host

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_ref as cr
        from oracle import fenris_oracle as fo
        n = SAMPLE_CELLS
        sv, sc = cr.gen_hex_mesh(n)
        sro, sci = cr.pattern(3, len(sv), sc)
        colors = cr.color_greedy(sc, len(sv))
        sw, sp_ = fo.hexahedron_gauss(2)
        vals = np.zeros(len(sci))
        cores = best_thread_count(cr, fo, sw, sp_, data, sv, sc, sro, sci, vals, colors)
        cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, sw, sp_, data, sv, sc, sro, sci, values=vals, colors=colors, nthreads=cores)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 200):
            vals[:] = 0
            cr.assemble(fo.HEX8, fo.LINEAR_ELASTIC, sw, sp_, data, sv, sc, sro, sci, values=vals, colors=colors, nthreads=cores)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": len(sc) / dt, "unit": "elements/s", "cores": cores, "kind": "port",
               "sample": f"Hex8 elasticity {n}^3 cube ({len(sc)} elements) x {reps} reps, coloured OpenMP C restatement of CsrParAssembler"}

    if rank == 0:
        out = {
            "metric": "elements/sec into global CSR (Hex8 3D linear elasticity, fp64)", "value": value, "unit": "elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: Hex8 linear elasticity (fenris-solid), unit cube {cells}^3 cells per GPU, Gauss 2^3, Lame(E=1e6, nu=0.2), u=0",
                       "elements_per_gpu": int(n_owned), "nnz_per_gpu": int(nnz), "scatter": args.scatter,
                       "parallelism": "1 GPU" if world == 1 else f"z-slab element partition x{world}, interface rows: " +
                                      ("neighbour ncclSend/ncclRecv, summed on arrival" if args.exchange == "peers" else "world ncclAllReduce"),
                       "l2": "inputs+outputs >> L2 (values 3.9 GB per GPU); no L2 flush needed", "setup_s": setup_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if other:
            out["other_modes_elements_per_s"] = other
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
