#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests/test_quadrature_tables_gpu.py tests/test_hex8_tile.py -m gpu -q --maxfail=6 > gpurun_out/pytest_tab.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tab.log
tail -n 30 gpurun_out/pytest_tab.log | cut -c1-220
timeout 150 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_tile64.log 2>&1
tail -n 1 gpurun_out/b_tile64.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['gpu_launches'])"
