#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
for mb in 5 6 8; do
FB200_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 3 --scatter atomic --no-e2e --no-cpu > gpurun_out/bench_atomic_mb$mb.log 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 --scatter colored --no-e2e --no-cpu > gpurun_out/bench_colored.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/prof_atomic_v4 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --scatter atomic > gpurun_out/ncu_full_atomic.log 2>&1
timeout 900 python scripts/bench_configs.py --configs c2,c5,c4 --modes atomic,gather > gpurun_out/bench_configs.log 2>&1
tail -n 2 gpurun_out/*.log
