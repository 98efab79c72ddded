#!/bin/bash
# session 2, call 7: pipelined Hex27 kernel, Tet4 chunk-size sweep
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_configs.py --configs c4 --modes atomic > gpurun_out/cfg_c4.log 2>&1
for ch in 512 1024 2048; do FB200_TET4_CHUNK=$ch timeout 600 python scripts/bench_configs.py --configs c2,c5 --modes atomic > gpurun_out/cfg_tet_$ch.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex27 -s 3 -c 1 -o gpurun_out/prof_hex27_mma3 python scripts/bench_configs.py --configs c4 --modes atomic --steps 2 > gpurun_out/ncu_hex27.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/cfg_c4.log gpurun_out/cfg_tet_*.log | cut -c1-330
