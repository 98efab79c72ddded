#!/bin/bash
# round 2, call 14: Tet4 chunk kernel variants on the C5 share (80^3 cells); tile suite after the latest changes
mkdir -p gpurun_out
for ch in 1024 512 2048; do
FB200_TET4_CHUNK=$ch timeout 300 python bench.py --workload c5 --cells 80 --no-e2e --no-cpu --steps 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 80^3 chunk $ch', round(d['ms_per_step'],4), '%.4g' % d['value'], round(d['roofline']['frac'],3), d['parity']['rel_frobenius'])" | tee -a gpurun_out/r2b_14_tet.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tet" > gpurun_out/r2b_14_parity.log 2>&1; tail -n 2 gpurun_out/r2b_14_parity.log
