#!/bin/bash
mkdir -p gpurun_out
for st in 1 0; do for own in 1 0; do
  FB200_TILE_STATIC=$st FB200_HEX8_OWNER=$own timeout 300 python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('static=$st owner=$own', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])" | tee -a gpurun_out/r2b_05_bench.log
done; done
FB200_TILE_STATIC=1 FB200_DEBUG=64 timeout 300 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 2> gpurun_out/r2b_05_waits.log | cut -c1-100
head -4 gpurun_out/r2b_05_waits.log
