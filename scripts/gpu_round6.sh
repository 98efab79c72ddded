#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/m_*.csv
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum,sm__warps_active.avg.pct_of_peak_sustained_active"
run() { # name, env..., cells
  name=$1; cells=$2; shift 2
  env "$@" timeout 300 ncu --metrics $M --clock-control none -k regex:assemble_hex8 -s 3 -c 1 --csv --log-file gpurun_out/m_$name.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic --cells $cells > gpurun_out/r_$name.log 2>&1
}
run c64 64 A=1
run c96 96 A=1
run c126_mb5 126 FB200_MINB=5
run c126_mb6 126 FB200_MINB=6
run c126_cap1 126 FB200_GRID_CAP=1
run c126_cap2 126 FB200_GRID_CAP=2
run c126_cap3 126 FB200_GRID_CAP=3
run c126_cap4 126 FB200_GRID_CAP=4
run c126_noorder 126 FB200_NO_ORDER=1
for cap in 2 3 4; do FB200_GRID_CAP=$cap timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/bench_cap$cap.log 2>&1; done
grep -h "assemble_hex8" gpurun_out/m_*.csv | cut -d, -f1,13- | head -80
