#!/bin/bash
# session 2, call 5: Hex27 DMMA kernel - parity, C4 timing against the generic kernel, one full ncu capture
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_configs.py --configs c4 --modes atomic,colored > gpurun_out/c4_mma.log 2>&1
FB200_HEX27_V1=1 timeout 300 python scripts/bench_configs.py --configs c4 --modes atomic > gpurun_out/c4_v1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex27 -s 3 -c 1 -o gpurun_out/prof_hex27_mma python scripts/bench_configs.py --configs c4 --modes atomic --steps 2 > gpurun_out/ncu_hex27.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/c4_mma.log gpurun_out/c4_v1.log | cut -c1-400
