#!/bin/bash
# round 2, call 13: ncu of the Tet4 chunk kernel with slot records (C5 share: 80^3 cells = 6.1 M tets)
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tet4_chunk -s 4 -c 1 -o gpurun_out/prof_tet4_rec python bench.py --workload c5 --cells 80 --no-e2e --no-cpu --no-parity --steps 2 --warmup 3 > gpurun_out/r2b_13_ncu.log 2>&1
timeout 300 python bench.py --workload c5 --cells 80 --no-e2e --no-cpu --no-parity --steps 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 80^3', d['ms_per_step'], d['value'], d['roofline']['frac'])"
ls -la gpurun_out/prof_tet4_rec.ncu-rep
