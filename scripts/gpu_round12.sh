#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "(global_assembly and hex8) or literal or accumulate" > gpurun_out/sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -q -x -k "global_assembly and hex8 and 8" > gpurun_out/racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/racecheck.log
timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_zfuse.log 2>&1
FB200_NO_ZFUSE=1 timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_nozfuse.log 2>&1
for cap in 2 4 6; do FB200_GRID_CAP=$cap timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_zfuse_cap$cap.log 2>&1; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/prof_atomic_v6 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_full_atomic.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/sanitizer.log gpurun_out/racecheck.log
