#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/m_*.csv
timeout 600 python -m pytest tests -m gpu -q -x -k "hex8 or smoke or literal or accumulate or error or api" > gpurun_out/pytest_hex8.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_hex8.log
for cap in 2 3 4 6; do FB200_GRID_CAP=$cap timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/bench_dyn_cap$cap.log 2>&1; done
for cap in 2 3; do FB200_STATIC_SCHED=1 FB200_GRID_CAP=$cap timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/bench_static_cap$cap.log 2>&1; done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active"
for cap in 3 6; do FB200_GRID_CAP=$cap timeout 300 ncu --metrics $M --clock-control none -k regex:assemble_hex8 -s 3 -c 1 --csv --log-file gpurun_out/m_dyn_cap$cap.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/r_dyn_cap$cap.log 2>&1; done
tail -n 3 gpurun_out/pytest_hex8.log
