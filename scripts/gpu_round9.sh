#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/m_*.csv
nvidia-smi --query-gpu=name,serial,uuid,clocks.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "hex8 or smoke or literal or accumulate or error or api" > gpurun_out/pytest_hex8.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_hex8.log
for rep in 1 2; do
for sched in dyn static; do for hint in hint nohint; do for cap in 2 3 4 6; do
  E=""; [ $sched = static ] && E="$E FB200_STATIC_SCHED=1"; [ $hint = nohint ] && E="$E FB200_NO_L2_HINTS=1"
  env $E FB200_GRID_CAP=$cap timeout 200 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_${sched}_${hint}_cap${cap}_r$rep.log 2>&1
done; done; done; done
tail -n 3 gpurun_out/pytest_hex8.log
