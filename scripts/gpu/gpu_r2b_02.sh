#!/bin/bash
# round 2, call 2: owner-store tile lists (no zero-fill) - parity of the tile tests, full-size C3 / C4 entrywise parity, bench both ways
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hex8_tile.py -q -x > gpurun_out/r2b_02_tile.log 2>&1; tail -n 5 gpurun_out/r2b_02_tile.log
for own in 1 0; do
  FB200_HEX8_OWNER=$own timeout 300 python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('owner=$own', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])" | tee -a gpurun_out/r2b_02_bench.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "full_size" > gpurun_out/r2b_02_full.log 2>&1; tail -n 5 gpurun_out/r2b_02_full.log
