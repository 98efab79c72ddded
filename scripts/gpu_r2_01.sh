#!/bin/bash
# session 2, call 1: DMMA Hex8 kernel - parity, fp64 pipe microbenchmark, A/B against the DFMA kernel, CTA sweep, one full ncu capture
mkdir -p gpurun_out; rm -f gpurun_out/*.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log
timeout 60 build/microbench_fp64 > gpurun_out/microbench_fp64.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
for cap in 2 3 4; do FB200_GRID_CAP=$cap timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_mma_cap$cap.log 2>&1; done
FB200_HEX8_DFMA=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_dfma.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex8 -s 3 -c 1 -o gpurun_out/prof_mma_v6 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/microbench_fp64.log; for f in gpurun_out/b_*.log; do echo $f; tail -n 1 $f | cut -c1-400; done
