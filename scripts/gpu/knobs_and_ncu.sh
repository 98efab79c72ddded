#!/bin/bash
# knob timings of the new tile kernel (which side bounds it); full ncu capture of the step's launch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_materials_gpu.py tests/test_host.py -q -x -m gpu > gpurun_out/knobs_mat.log 2>&1; tail -n 3 gpurun_out/knobs_mat.log
for dbg in 0 1 2 3 4 5 6 7; do
  FB200_DEBUG=$dbg timeout 300 python bench.py --no-e2e --no-cpu --no-parity --steps 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dbg=$dbg', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))" | tee -a gpurun_out/knobs_knobs.log
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hex8_tile -s 4 -c 1 -o gpurun_out/prof_tile_owner python bench.py --no-e2e --no-cpu --no-parity --steps 2 --warmup 3 > gpurun_out/knobs_ncu.log 2>&1
ls -la gpurun_out/prof_tile_owner.ncu-rep
