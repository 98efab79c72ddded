#!/usr/bin/env python3
"""Times the BASELINE.json configs that are NOT the headline bench line (C2 Tet4 Poisson, C4 Hex27 elasticity, a per-GPU share
of C5 Tet4 elasticity) - they are parity-test cases in tests/, this script only records where their kernels stand.
One JSON line per (config, scatter mode)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fenris_b200 as fb  # noqa: E402

MODES = {"atomic": 0, "colored": 1, "gather": 2}


def measure_mass_or_vector(what, steps, peak):
    mesh = fb.create_unit_box_uniform_hex_mesh_3d(126)
    E, N = mesh.num_elements(), mesh.num_nodes()
    with fb.Context(0) as ctx:
        ctx.space_upload(mesh.element_type, mesh.vertices(), mesh.connectivity())
        if what == "mass":
            w, p = fb.canonical_stiffness_quadrature(fb.HEX27)  # Gauss 3^3 (canonical.rs:102-112)
            nrows, nnz = ctx.assemble_pattern(3)
            fn = lambda: ctx.assemble_mass_into_csr_device(w, p, 1000.0, accumulate=False)
            b_algo = 4 * 8 * E + 8 * 3 * N + 4 * 64 * E + 16 * nnz
            name = "Hex8 mass matrix (s = 3, Gauss 3^3) on the C3 mesh, device-resident CSR"
        elif what == "stvk":  # tangent stiffness of StVKMaterial at u (u is copied host -> device inside the step)
            w, p = fb.canonical_stiffness_quadrature(fb.HEX8)  # Gauss 2^3
            nrows, nnz = ctx.assemble_pattern(3)
            u = 0.01 * np.random.default_rng(0).normal(size=3 * N)
            fn = lambda: ctx.assemble_into_csr_device(fb.STVK, w, p, (3.0e5, 2.0e5), accumulate=False, u=u)
            b_algo = 4 * 8 * E + 8 * 3 * N + 8 * 3 * N + 4 * 64 * E + 16 * nnz
            name = "Hex8 StVK tangent stiffness at u (Gauss 2^3) on the C3 mesh, device-resident CSR"
        else:
            w, p = fb.canonical_stiffness_quadrature(fb.HEX8)  # Gauss 2^3
            nnz = 0
            out = np.zeros(3 * N)
            g = np.tile([0.0, -9.81, 0.0], (len(w), 1))
            fn = lambda: ctx.assemble_vector(w, p, g, N, out=out)
            b_algo = 4 * 8 * E + 8 * 3 * N + 16 * 3 * N
            name = "Hex8 source vector (uniform body force, s = 3) on the C3 mesh, host output (D2H of 3 N doubles inside the step)"
        for _ in range(3):
            fn()
        ctx.synchronize()
        ctx.timer_begin()
        for _ in range(steps):
            fn()
        ms = ctx.timer_end() / steps
        ctx.synchronize()
        print(json.dumps({"config": name, "scatter": "atomic", "elements": E, "nnz": nnz, "ms_per_step": ms, "elements_per_s": E / (ms * 1e-3),
                          "algorithmic_GBps": b_algo / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": b_algo / (ms * 1e-3) / 1e9 / peak}), flush=True)


def measure_cg(peak):
    import time
    mesh = fb.create_unit_box_uniform_hex_mesh_3d(126)
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(1e6, 0.2))
    w, p = fb.canonical_stiffness_quadrature(fb.HEX8)  # Gauss 2^3
    N = mesh.num_nodes()
    with fb.Context(0) as ctx:
        ctx.space_upload(mesh.element_type, mesh.vertices(), mesh.connectivity())
        nrows, nnz = ctx.assemble_pattern(3)
        ctx.assemble_into_csr_device(fb.LINEAR_ELASTIC, w, p, (lame.mu, lame.lambda_))
        b = ctx.assemble_vector(w, p, np.tile([0.0, -9.81, 0.0], (len(w), 1)), N)
        clamped = np.nonzero(mesh.vertices()[:, 1] < 1e-12)[0]
        ctx.apply_homogeneous_dirichlet_bc_csr(clamped)
        for node in clamped:
            b[3 * node:3 * node + 3] = 0.0
        times = {}
        for its in (10, 10, 110, 10, 110):
            ctx.synchronize()
            t0 = time.perf_counter()
            try:
                ctx.cg_solve(b, rel_tol=1e-30, max_iter=its, jacobi=True)
                raise RuntimeError("converged to 1e-30?")
            except fb.Fb200Error as exc:
                assert exc.status == fb.ERR_NOT_CONVERGED, exc
            times.setdefault(its, []).append(time.perf_counter() - t0)
        ms_it = (min(times[110]) - min(times[10])) / 100 * 1e3
        bytes_it = 8 * nnz + 4 * (nnz // 9) + 8 * nrows * 14  # values + node-block columns + the vector passes of one iteration
        print(json.dumps({"config": "Jacobi-PCG on the device-resident C3 Hex8 elasticity matrix (6.1 M unknowns, nnz 4.9e8)", "ms_per_iteration": ms_it,
                          "algorithmic_GBps": bytes_it / (ms_it * 1e-3) / 1e9, "frac_of_measured_hbm": bytes_it / (ms_it * 1e-3) / 1e9 / peak,
                          "note": "wall clock difference of 110 and 10 iterations (host-driven loop, three scalar read-backs per iteration)", "raw_s": times}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c4,c5")
    ap.add_argument("--modes", default="atomic,gather")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    lame = fb.LameParameters.from_young_poisson(fb.YoungPoisson(1e6, 0.2))
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for cfg in args.configs.split(","):
        if cfg == "c2":
            mesh, op, data, sdim, name = fb.create_unit_box_uniform_tet_mesh_3d(44), fb.LAPLACE, None, 1, "C2 Tet4 Poisson 44^3 cells"
        elif cfg == "c4":
            mesh, op, data, sdim, name = fb.hex27_mesh_from(fb.create_unit_box_uniform_hex_mesh_3d(63)), fb.LINEAR_ELASTIC, (lame.mu, lame.lambda_), 3, "C4 Hex27 elasticity 63^3 cells"
        elif cfg == "c5":
            mesh, op, data, sdim, name = fb.create_unit_box_uniform_tet_mesh_3d(80), fb.LINEAR_ELASTIC, (lame.mu, lame.lambda_), 3, "C5 share: Tet4 elasticity 80^3 cells (1/8 of 161^3)"
        elif cfg == "hex20":  # north_star's high-order tensor path: Hex20 (60 x 60 K_e) on the 63^3 cube
            mesh, op, data, sdim, name = fb.hex20_mesh_from(fb.create_unit_box_uniform_hex_mesh_3d(63)), fb.LINEAR_ELASTIC, (lame.mu, lame.lambda_), 3, "Hex20 elasticity 63^3 cells"
        elif cfg == "tet10":  # Tet10 (30 x 30 K_e) on the 30^3 BCC box (324 000 tets)
            mesh, op, data, sdim, name = fb.tet10_mesh_from(fb.create_unit_box_uniform_tet_mesh_3d(30)), fb.LINEAR_ELASTIC, (lame.mu, lame.lambda_), 3, "Tet10 elasticity 30^3 cells"
        elif cfg == "c3":
            mesh, op, data, sdim, name = fb.create_unit_box_uniform_hex_mesh_3d(126), fb.LINEAR_ELASTIC, (lame.mu, lame.lambda_), 3, "C3 Hex8 elasticity 126^3"
        elif cfg == "cg":  # Jacobi-PCG iterations on the device-resident C3 elasticity matrix
            measure_cg(peak)
            continue
        elif cfg in ("mass", "vector", "stvk"):  # SURVEY 8f rank 1 on the C3 mesh: mass matrix (s = 3, Gauss 3^3) / source vector (Gauss 2^3)
            measure_mass_or_vector(cfg, args.steps, peak)
            continue
        else:
            continue
        w, p = fb.canonical_stiffness_quadrature(mesh.element_type)
        n, d = mesh.connectivity().shape[1], mesh.vertices().shape[1]
        with fb.Context(0) as ctx:
            ctx.space_upload(mesh.element_type, mesh.vertices(), mesh.connectivity())
            nrows, nnz = ctx.assemble_pattern(sdim)
            ctx.color_nodes()
            E, N = mesh.num_elements(), mesh.num_nodes()
            b_algo = 4 * n * E + 8 * d * N + 4 * n * n * E + 16 * nnz
            for mname in args.modes.split(","):
                m = MODES[mname]
                try:
                    for _ in range(3):
                        ctx.assemble_into_csr_device(op, w, p, data, scatter_mode=m, accumulate=False)
                    ctx.synchronize()
                    ctx.timer_begin()
                    for _ in range(args.steps):
                        ctx.assemble_into_csr_device(op, w, p, data, scatter_mode=m, accumulate=False)
                    ms = ctx.timer_end() / args.steps
                    ctx.synchronize()
                    print(json.dumps({"config": name, "scatter": mname, "elements": E, "nnz": nnz, "ms_per_step": ms, "elements_per_s": E / (ms * 1e-3),
                                      "algorithmic_GBps": b_algo / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": b_algo / (ms * 1e-3) / 1e9 / peak}), flush=True)
                except fb.Fb200Error as exc:
                    print(json.dumps({"config": name, "scatter": mname, "error": str(exc)}), flush=True)


if __name__ == "__main__":
    main()
