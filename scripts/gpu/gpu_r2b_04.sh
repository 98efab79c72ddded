#!/bin/bash
mkdir -p gpurun_out
FB200_DEBUG=64 timeout 300 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 2> gpurun_out/r2b_04_waits.log | cut -c1-200
head -12 gpurun_out/r2b_04_waits.log
