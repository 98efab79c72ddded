#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
for c in 16 24 32 40 48 64 96; do
timeout 300 python bench.py --steps 30 --warmup 3 --scatter atomic --no-e2e --no-cpu --cells $c > gpurun_out/bench_cells$c.log 2>&1
done
tail -n 1 gpurun_out/*.log | cut -c1-200
