#!/bin/bash
# round 2, call 18: 64-bit row table of the flush, set-up time after the staged parallel D2H; tile + parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_tile.py -q -x > gpurun_out/r2b_18_tile.log 2>&1; tail -n 2 gpurun_out/r2b_18_tile.log
for i in 1 2; do
FB200_DEBUG_SETUP=1 timeout 600 python bench.py --no-e2e --no-cpu --steps 20 2> gpurun_out/r2b_18_setup.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], d['roofline']['kernel_ms'], 'setup', d['config']['setup_s'], d['parity']['rel_frobenius'])"
done
grep "fb200 setup" gpurun_out/r2b_18_setup.log | grep -v "worker\|  tile"
FB200_TILE_STATIC=0 FB200_HEX8_OWNER=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ticket+memset', d['ms_per_step'], d['roofline']['kernel_ms'])"
