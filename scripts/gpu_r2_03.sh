#!/bin/bash
# session 2, call 3: is the scatter bound per lane or per sector?  16 = only the first double of every node block (1/3 of the lanes, ~3/4 of the sectors),
# 32 = only the first of the three rows of a node (1/3 of the instructions, lanes and sectors)
mkdir -p gpurun_out; rm -f gpurun_out/*.log
for cap in 2 3; do for dbg in 0 16 32 48; do
FB200_DEBUG=$dbg FB200_GRID_CAP=$cap timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_dbg${dbg}_cap$cap.log 2>&1
done; done
for f in gpurun_out/b_dbg*.log; do echo -n "$f "; tail -n 1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"; done
