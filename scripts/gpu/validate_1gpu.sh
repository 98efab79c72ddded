#!/bin/bash
# final validation of a tree - full GPU suite, smoke, default bench line, ncu summary of the step's launch, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/final_pytest.log 2>&1; tail -n 3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-200
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python -c "
import json; d = json.load(open('gpurun_out/final_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['parity']['rel_frobenius'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])" || tail -5 gpurun_out/final_bench.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:hex8_tile -s 4 -c 1 -o gpurun_out/prof_tile_final python bench.py --no-e2e --no-cpu --no-parity --steps 2 --warmup 3 > gpurun_out/final_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/final_ncu_launch.log 2>&1
ls -la gpurun_out/prof_tile_final.ncu-rep
