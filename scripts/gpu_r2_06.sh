#!/bin/bash
# session 2, call 6: Tet4 chunk-local kernel + Hex27 scatter rewrite - parity, C2/C4/C5-share timings (new vs generic kernel), ncu of both
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_configs.py --configs c2,c4,c5 --modes atomic > gpurun_out/cfg_new.log 2>&1
FB200_TET4_V1=1 timeout 600 python scripts/bench_configs.py --configs c2,c5 --modes atomic > gpurun_out/cfg_v1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_tet4 -s 3 -c 1 -o gpurun_out/prof_tet4_chunk python scripts/bench_configs.py --configs c5 --modes atomic --steps 2 > gpurun_out/ncu_tet4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_hex27 -s 3 -c 1 -o gpurun_out/prof_hex27_mma2 python scripts/bench_configs.py --configs c4 --modes atomic --steps 2 > gpurun_out/ncu_hex27.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/cfg_new.log gpurun_out/cfg_v1.log | cut -c1-330
