#!/bin/bash
# round 2, call 8 (2 GPUs): config C5 (Tet4 elasticity) - small box first, then the full 161^3 box at N = 1 and N = 2
mkdir -p gpurun_out
run() { # name, nproc, extra args
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 900 python bench.py --workload c5 --no-e2e --no-cpu "$@" > gpurun_out/r2b_08_$name.json 2> gpurun_out/r2b_08_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --gpus $np --no-e2e --no-cpu "$@" > gpurun_out/r2b_08_$name.json 2> gpurun_out/r2b_08_$name.err; fi
  tail -n 1 gpurun_out/r2b_08_$name.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$name', d['n_gpus'], d['ms_per_step'], d['value'], d['parity']['rel_frobenius'], d['parity']['rows_checked'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['setup_s'], d['config']['mesh_s'])" || tail -n 8 gpurun_out/r2b_08_$name.err
}
run small_n1 1 --cells 40 --steps 5
run small_n2 2 --cells 40 --steps 5
run full_n1 1 --steps 10
run full_n2 2 --steps 10
