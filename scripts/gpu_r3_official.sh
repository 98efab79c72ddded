#!/bin/bash
# session 3: full GPU suite + the numbers committed under profiles/r01 (tile kernel)
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/final_*
timeout 1200 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/final_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 python bench.py --all-modes --no-e2e --no-cpu --steps 10 > gpurun_out/final_bench_all_modes.json 2>/dev/null
FB200_HEX8_TILE=0 timeout 300 python bench.py --no-e2e --no-cpu --steps 10 > gpurun_out/final_bench_tile0.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 600 python scripts/bench_configs.py --configs c2,c4,c5 --modes atomic > gpurun_out/final_configs.log 2>&1
tail -n 2 gpurun_out/final_bench.json gpurun_out/final_smoke.log | cut -c1-600; cat gpurun_out/final_configs.log | cut -c1-300
