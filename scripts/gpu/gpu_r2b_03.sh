#!/bin/bash
# round 2, call 3: where does the owner-store step lose its time?  knobs of the flush (FB200_DEBUG: 16 no wait, 8 no publish + no wait, 40 also no barrier)
mkdir -p gpurun_out
for dbg in 0 16 8 40; do
  FB200_DEBUG=$dbg timeout 300 python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dbg=$dbg', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])" | tee -a gpurun_out/r2b_03_bench.log
done
