#!/bin/bash
# quick validation of a tree on one GPU: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/quick_pytest.log 2>&1; tail -n 3 gpurun_out/quick_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-200
timeout 900 python bench.py > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python -c "
import json; d = json.load(open('gpurun_out/quick_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['parity']['rel_frobenius'], d['e2e']['value'], d['config']['setup_s'], d['gpu_launches'])" || tail -5 gpurun_out/quick_bench.err
