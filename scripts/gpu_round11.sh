#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
for mb in 5 6; do for cap in 3 4 5; do
  FB200_MINB=$mb FB200_CHUNK=8 FB200_GRID_CAP=$cap timeout 200 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_c8_mb${mb}_cap${cap}.log 2>&1
done; done
FB200_CHUNK=8 FB200_GRID_CAP=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/prof_atomic_v5 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_full_atomic.log 2>&1
