#!/bin/bash
# round 2, call 22: Hex20 / Tet10 on the templated DMMA kernel: parity suites, timings against the generic element kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hex20_gpu.py tests/test_gpu_parity.py tests/test_materials_gpu.py tests/test_quadrature_tables_gpu.py tests/test_mass_source_gpu.py -q -x -k "not full_size_c3" > gpurun_out/r2b_22_pytest.log 2>&1; tail -n 3 gpurun_out/r2b_22_pytest.log
timeout 600 python scripts/bench_configs.py --configs hex20,tet10,c4 --modes atomic --steps 5 > gpurun_out/r2b_22_dmma.jsonl 2>&1; cut -c1-330 gpurun_out/r2b_22_dmma.jsonl
FB200_HEX27_V1=1 timeout 600 python scripts/bench_configs.py --configs hex20,tet10 --modes atomic --steps 5 > gpurun_out/r2b_22_generic.jsonl 2>&1; cut -c1-330 gpurun_out/r2b_22_generic.jsonl
