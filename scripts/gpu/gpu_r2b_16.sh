#!/bin/bash
# round 2, call 16: where does the set-up time go (FB200_DEBUG_SETUP), tile tests after the list-builder changes
mkdir -p gpurun_out
nproc
FB200_DEBUG_SETUP=1 timeout 600 python bench.py --no-e2e --no-cpu --steps 5 2> gpurun_out/r2b_16_setup.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['config']['setup_s'], d['config']['mesh_s'], d['parity']['ok'])"
grep "fb200 setup" gpurun_out/r2b_16_setup.log | grep -v worker
timeout 600 python -m pytest tests/test_hex8_tile.py -q -x > gpurun_out/r2b_16_tile.log 2>&1; tail -n 2 gpurun_out/r2b_16_tile.log
