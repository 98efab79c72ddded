#!/bin/bash
# first GPU round trip: tests (small), sanitizer on a slice, bench (all modes), launch list + one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "not full_size" > gpurun_out/pytest_small.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_small.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "pattern_kats or (global_assembly and hex8 and 6) or error_paths" > gpurun_out/sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer.log
timeout 600 python bench.py --steps 5 --warmup 3 --cells 64 --all-modes --no-cpu > gpurun_out/bench64.log 2>&1; echo "rc=$?" >> gpurun_out/bench64.log
timeout 900 python bench.py --steps 10 --warmup 3 --all-modes > gpurun_out/bench126.log 2>&1; echo "rc=$?" >> gpurun_out/bench126.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "full_size" > gpurun_out/pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_full.log
for m in gather atomic; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$m.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter $m > gpurun_out/ncu_launch_$m.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_gather -s 3 -c 1 -o gpurun_out/prof_gather python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter gather > gpurun_out/ncu_full_gather.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_elements -s 3 -c 1 -o gpurun_out/prof_atomic python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_full_atomic.log 2>&1
tail -3 gpurun_out/*.log
