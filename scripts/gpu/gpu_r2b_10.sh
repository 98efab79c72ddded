#!/bin/bash
# round 2, call 10: early publish (owned shared rows first, first poll hidden behind the complete rows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_tile.py -q -x > gpurun_out/r2b_10_tile.log 2>&1; tail -n 3 gpurun_out/r2b_10_tile.log
for st in 1 0; do
  FB200_TILE_STATIC=$st timeout 300 python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('static=$st', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['parity']['rel_frobenius'])" | tee -a gpurun_out/r2b_10_bench.log
done
FB200_DEBUG=64 timeout 300 python bench.py --no-e2e --no-cpu --no-parity --steps 3 --warmup 3 2> gpurun_out/r2b_10_waits.log | cut -c1-100
head -3 gpurun_out/r2b_10_waits.log
