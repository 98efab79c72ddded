#!/bin/bash
# round-trip 2: full gpu tests, sanitizer slice, bench variants (v2 kernel, Morton order on/off), ncu of the atomic v2 kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "element_matrices or (global_assembly and hex8) or error_paths or literal" > gpurun_out/sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer.log
timeout 600 python bench.py --steps 10 --warmup 3 --all-modes --scatter atomic > gpurun_out/bench_atomic.log 2>&1; echo "rc=$?" >> gpurun_out/bench_atomic.log
FB200_NO_ORDER=1 timeout 600 python bench.py --steps 10 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/bench_atomic_noorder.log 2>&1
FB200_HEX8_V1=1 timeout 600 python bench.py --steps 10 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/bench_atomic_v1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_atomic.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_launch_atomic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/prof_atomic_v2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_full_atomic.log 2>&1
FB200_NO_ORDER=1 timeout 900 ncu --set full --clock-control none -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/prof_atomic_v2_noorder python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_full_atomic_noorder.log 2>&1
tail -2 gpurun_out/*.log
