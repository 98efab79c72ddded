#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests/test_hex20_gpu.py -m gpu -q --maxfail=6 > gpurun_out/pytest_h20.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_h20.log
tail -n 40 gpurun_out/pytest_h20.log | cut -c1-220
