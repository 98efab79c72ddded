#!/bin/bash
# session 4: full GPU suite (with the StVK rows), smoke, the default bench line and the StVK timing
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_full_r4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_full_r4.log
tail -n 4 gpurun_out/pytest_full_r4.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; tail -n 2 gpurun_out/r4_smoke.log | cut -c1-300
timeout 200 python scripts/bench_configs.py --configs stvk --steps 5 > gpurun_out/r4_stvk.log 2>&1; cut -c1-400 gpurun_out/r4_stvk.log
timeout 400 python bench.py > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; cut -c1-900 gpurun_out/r4_bench.json
