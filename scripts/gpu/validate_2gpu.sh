#!/bin/bash
# 2 GPUs: multi-GPU tests (incl. p2p with zero-fill lists), bench N = 2, N = 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/v2_multi.log 2>&1; tail -n 2 gpurun_out/v2_multi.log
for i in 1 2; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --no-e2e 2>/dev/null | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('n2', d['ms_per_step'], d['value'], d['config']['exchange'], d['parity']['rel_frobenius'])"; done
timeout 600 python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], d['roofline']['kernel_ms'], d['parity']['rel_frobenius'])"
