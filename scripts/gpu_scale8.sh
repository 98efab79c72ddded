#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sc_*.log
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/sc_bench$N.log 2>&1; echo "rc=$?" >> gpurun_out/sc_bench$N.log
tail -n 3 gpurun_out/sc_bench$N.log | cut -c1-700
