#!/bin/bash
# round 2, call 15: tile-coloured deterministic scatter, restored Tet4 kernel, full GPU suite, bench with all modes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_15_pytest.log 2>&1; tail -n 5 gpurun_out/r2b_15_pytest.log
timeout 600 python bench.py --all-modes --steps 10 > gpurun_out/r2b_15_bench.json 2> gpurun_out/r2b_15_bench.err
python -c "
import json; d = json.load(open('gpurun_out/r2b_15_bench.json')); print(d['ms_per_step'], d['value'], d['other_modes_elements_per_s'], d['e2e'], d.get('e2e_solve'), d['config']['setup_s'])" || tail -5 gpurun_out/r2b_15_bench.err
timeout 300 python bench.py --workload c5 --cells 80 --no-e2e --no-cpu --steps 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 80^3', round(d['ms_per_step'],4), '%.4g' % d['value'], round(d['roofline']['frac'],3), d['parity']['rel_frobenius'])"
