#!/bin/bash
# round 2, call 12 (2 GPUs): Tet4 chunk kernel with slot records + fused p2p exchange: parity tests, C5 at N = 1 / 2, single-GPU Tet4 tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x -k "element_range or four" > gpurun_out/r2b_12_multi.log 2>&1; tail -n 4 gpurun_out/r2b_12_multi.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "not full_size_c4 and not full_size_c3" > gpurun_out/r2b_12_parity.log 2>&1; tail -n 3 gpurun_out/r2b_12_parity.log
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 900 python bench.py "$@" > gpurun_out/r2b_12_$name.json 2> gpurun_out/r2b_12_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $np "$@" > gpurun_out/r2b_12_$name.json 2> gpurun_out/r2b_12_$name.err; fi
  tail -n 1 gpurun_out/r2b_12_$name.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$name', d['n_gpus'], round(d['ms_per_step'], 4), '%.4g' % d['value'], d['config'].get('exchange'), 'parity %.2e rows %d' % (d['parity']['rel_frobenius'], d['parity']['rows_checked']), 'kernel', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3), 'setup', round(d['config']['setup_s'], 2))" || tail -n 8 gpurun_out/r2b_12_$name.err
}
run c5_n1 1 --workload c5 --no-e2e --no-cpu --steps 10
run c5_n2_p2p 2 --workload c5 --no-e2e --steps 10
run c5_n2_peers 2 --workload c5 --no-e2e --steps 10 --exchange peers
