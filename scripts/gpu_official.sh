#!/bin/bash
# the numbers committed under profiles/: default bench line, reference arm, launch list, one full capture of the top kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/final_*
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/final_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/final_prof_atomic python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py --all-modes --no-e2e --no-cpu --steps 10 > gpurun_out/final_bench_all_modes.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
tail -n 2 gpurun_out/final_bench.json gpurun_out/final_smoke.log | cut -c1-300
