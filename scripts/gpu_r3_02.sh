#!/bin/bash
# session 3, call 3: ncu full capture (with source) of the Hex8 tile kernel on C3
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8_tile -s 2 -c 1 -f -o gpurun_out/prof_tile64_v6 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_tile.log 2>&1
tail -n 3 gpurun_out/ncu_tile.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep | tail -3
