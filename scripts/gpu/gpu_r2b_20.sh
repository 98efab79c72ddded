#!/bin/bash
# round 2, call 20: final single-GPU validation - full GPU suite, smoke, default bench line, reference arm, launch list of the default command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_20_pytest.log 2>&1; tail -n 3 gpurun_out/r2b_20_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-200
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2b_20_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r2b_20_bench.json 2> gpurun_out/r2b_20_bench.err
kill $SMI
python -c "
import json; d = json.load(open('gpurun_out/r2b_20_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['frac_of_step'], d['parity'], d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['gpu_launches'], d['clocks'], d.get('e2e_solve'))" || tail -5 gpurun_out/r2b_20_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2b_20_reference.json 2> gpurun_out/r2b_20_reference.err; cut -c1-400 gpurun_out/r2b_20_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r2b_20_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2b_20_ncu_launch.log 2>&1
grep -c "assemble_hex8_tile" gpurun_out/r2b_20_launches.csv
