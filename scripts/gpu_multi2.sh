#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/mg_*.log
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/mg_gpus.log 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/mg_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/mg_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg_bench2.log 2>&1; echo "rc=$?" >> gpurun_out/mg_bench2.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/mg_bench1.log 2>&1
tail -n 4 gpurun_out/mg_*.log
