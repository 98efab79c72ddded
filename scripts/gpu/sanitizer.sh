#!/bin/bash
# compute-sanitizer on the new kernels (small meshes): memcheck of the tile tests and a Tet4 case, racecheck of one tile case
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_hex8_tile.py -q -x -k "owner or colored or equals_oracle" > gpurun_out/sanitizer_memcheck_tile.log 2>&1; echo "memcheck tile rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_tile.log | tail -n 3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -q -x -k "tet4 and not full_size" > gpurun_out/sanitizer_memcheck_tet.log 2>&1; echo "memcheck tet rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_tet.log | tail -n 3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_hex8_tile.py -q -x -k "test_tile_owner_stores_equal_zero_fill and 9" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | tail -n 8
