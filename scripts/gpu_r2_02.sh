#!/bin/bash
# session 2, call 2: where does the Hex8 DMMA kernel's time go?  debug knobs: 1 skip compute, 2 skip scatter, 4 st instead of red, 8 L2 prefetch
mkdir -p gpurun_out; rm -f gpurun_out/*.log
for cap in 2 3; do for dbg in 0 1 2 3 4 8; do
FB200_DEBUG=$dbg FB200_GRID_CAP=$cap timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_dbg${dbg}_cap$cap.log 2>&1
done; done
for cap in 4 6; do for dbg in 0 8; do
FB200_DEBUG=$dbg FB200_GRID_CAP=$cap timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_dbg${dbg}_cap$cap.log 2>&1
done; done
for f in gpurun_out/b_dbg*.log; do echo -n "$f "; tail -n 1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"; done
