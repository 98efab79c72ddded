#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
for chunk in 1 4 8 64 512; do for cap in 2 3 6; do
  FB200_CHUNK=$chunk FB200_GRID_CAP=$cap timeout 200 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_chunk${chunk}_cap${cap}.log 2>&1
done; done
FB200_STATIC_SCHED=1 FB200_GRID_CAP=2 timeout 200 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_static_cap2.log 2>&1
