#!/bin/bash
# round 2, call 17 (2 GPUs): multi-GPU suite with the restored Tet4 kernel + fused exchange, C5 at N = 2 (p2p / peers), C3 N = 2, set-up time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2b_17_multi.log 2>&1; tail -n 3 gpurun_out/r2b_17_multi.log
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 900 python bench.py "$@" > gpurun_out/r2b_17_$name.json 2> gpurun_out/r2b_17_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $np "$@" > gpurun_out/r2b_17_$name.json 2> gpurun_out/r2b_17_$name.err; fi
  tail -n 1 gpurun_out/r2b_17_$name.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$name', d['n_gpus'], round(d['ms_per_step'], 4), '%.4g' % d['value'], d['config'].get('exchange'), 'parity %.2e rows %d' % (d['parity']['rel_frobenius'], d['parity']['rows_checked']), 'kernel', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 3), 'setup', round(d['config']['setup_s'], 2))" || tail -n 8 gpurun_out/r2b_17_$name.err
}
run c5_n2_p2p 2 --workload c5 --no-e2e --steps 10
run c5_n2_peers 2 --workload c5 --no-e2e --steps 10 --exchange peers
run c3_n2_p2p 2 --no-e2e
FB200_DEBUG_SETUP=1 timeout 600 python bench.py --no-e2e --no-cpu --steps 5 2> gpurun_out/r2b_17_setup.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], 'setup', d['config']['setup_s'], d['parity']['ok'])"
grep "fb200 setup" gpurun_out/r2b_17_setup.log | grep -v "worker\|  tile"
