#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
export FB200_ZFUSE=1
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "(global_assembly and hex8) or literal or accumulate" > gpurun_out/sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer.log
for cap in 2 3 4; do for look in 512 2048 8192; do FB200_ZERO_LOOK=$look FB200_GRID_CAP=$cap timeout 120 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_z_cap${cap}_look$look.log 2>&1; done; done
unset FB200_ZFUSE
timeout 300 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/b_nozfuse.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/sanitizer.log
