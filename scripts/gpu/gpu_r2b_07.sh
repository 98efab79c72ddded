#!/bin/bash
# round 2, call 7 (2 GPUs): multi-GPU parity tests (p2p fused exchange, Tet4 / Hex27 slabs), weak-scaling bench N = 2 with both exchanges
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2b_07_multi.log 2>&1; tail -n 15 gpurun_out/r2b_07_multi.log
for ex in p2p peers; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --exchange $ex > gpurun_out/r2b_07_bench_n2_$ex.json 2> gpurun_out/r2b_07_bench_n2_$ex.err
  tail -n 1 gpurun_out/r2b_07_bench_n2_$ex.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$ex', d['ms_per_step'], d['value'], d['config']['exchange'], d['parity'], d['roofline']['kernel_ms'])" || tail -n 5 gpurun_out/r2b_07_bench_n2_$ex.err
done
timeout 600 python bench.py --steps 20 --no-e2e > gpurun_out/r2b_07_bench_n1.json 2> gpurun_out/r2b_07_bench_n1.err; python -c "
import json; d = json.load(open('gpurun_out/r2b_07_bench_n1.json')); print('n1', d['ms_per_step'], d['value'], d['parity'], d['roofline']['kernel_ms'], d['config']['setup_s'])"
