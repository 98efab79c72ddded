#!/bin/bash
# full GPU suite with the owner-store + round-robin default, smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_06_pytest.log 2>&1; tail -n 5 gpurun_out/r2b_06_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2b_06_bench.json 2> gpurun_out/r2b_06_bench.err; cut -c1-1500 gpurun_out/r2b_06_bench.json
