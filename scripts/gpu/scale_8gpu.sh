#!/bin/bash
# scaling numbers on one 8-GPU box - 4-rank parity tests, C3 weak (default line at N = 8), C5 strong with the fused exchange
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x -k "four or p2p" > gpurun_out/scale_multi.log 2>&1; tail -n 3 gpurun_out/scale_multi.log
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 900 python bench.py "$@" > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $np "$@" > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err; fi
  tail -n 1 gpurun_out/scale_$name.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); e = d.get('e2e') or {}
print('$name', d['n_gpus'], round(d['ms_per_step'], 4), '%.4g' % d['value'], d['config'].get('exchange'), 'parity %.2e rows %d' % (d['parity']['rel_frobenius'], d['parity']['rows_checked']), 'kernel', round(d['roofline']['kernel_ms'], 4), 'e2e', e.get('ms_per_step'), 'setup', round(d['config']['setup_s'], 2))" || tail -n 8 gpurun_out/scale_$name.err
}
run c3_n8 8
run c5_n8_p2p 8 --workload c5 --no-e2e --steps 10
run c5_n4_p2p 4 --workload c5 --no-e2e --steps 10
run c5_n8_peers 8 --workload c5 --no-e2e --steps 10 --exchange peers
run c3_n1 1 --no-e2e --no-cpu
