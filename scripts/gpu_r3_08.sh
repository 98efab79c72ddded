#!/bin/bash
# ncu captures of the widened-row kernels on the C3 mesh (SpMV of the PCG, mass matrix)
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 5 -c 1 -f -o gpurun_out/prof_spmv python scripts/bench_configs.py --configs cg > gpurun_out/ncu_spmv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mass_source_kernel -s 2 -c 1 -f -o gpurun_out/prof_mass python scripts/bench_configs.py --configs mass --steps 2 > gpurun_out/ncu_mass.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
